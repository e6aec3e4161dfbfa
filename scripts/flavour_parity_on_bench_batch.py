"""Both library flavours against the fp32 CPU oracle on utterances 0..R-1 of bench.py's timed batch (seed 1000, B = 64):
encoder states, aggregator output (plain and centred per utterance) and index agreement, plus the fp64 margin of every
differing token.  python scripts/flavour_parity_on_bench_batch.py [R] [B] > gpurun_out/flavour_parity.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from taste_spokenlm_b200 import synth
from taste_spokenlm_b200.tower import TasteAudioTowerB200
from oracle import taste_oracle as O

R = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = 64
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
cfg = synth.FULL
W = synth.random_weights(cfg, 1234)
batch = bench.make_batch(1000, B, T, dev)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


torch.set_num_threads(os.cpu_count() or 1)
ref = []
for i in range(R):
    wav, ids, wid = batch["wav"][i].cpu(), batch["ids"][i].cpu(), batch["wid"][i].cpu()
    feats, _ = O.log_mel(wav[None])
    out = O.tower_forward(W, ids[None], torch.tensor([T], dtype=torch.int32), feats, wid[None], cfg.heads, cfg.enc_layers,
                          cfg.dec_layers, cfg.num_quantizers, cfg.target_hidden_layer, stages=True)
    ref.append({k: out[k][0] for k in ("_h_last", "_h_target", "_aggregated", "quantized_indices")})
    print("oracle", i, file=sys.stderr, flush=True)

Win = W["vq.rvq.project_in.weight"].double()
bin_ = W["vq.rvq.project_in.bias"].double()
codes = [W[f"vq.rvq.layers.{q}._codebook.embed"][0].double() for q in range(cfg.num_quantizers)]

report = {"utterances": R, "batch": B, "tokens": T}
for prec in ("bf16", "fp16"):
    tower = TasteAudioTowerB200.from_config(cfg, precision=prec).eval()
    tower.load_state_dict(W, strict=True)
    tower = tower.to(dev)
    eng = tower.engine()
    _, feats = eng.logmel(batch["wav"], batch["n_samples"], want_f32=False, want_bf16=True)
    h_last, h_t = eng.encode(feats)
    h_last, h_t = h_last[:R].float().cpu(), h_t[:R].float().cpu()
    agg, _ = eng.segment_and_quantize(*eng.encode(feats), batch["ids"], batch["wid"], batch["lengths_host"], skip_vq=True)
    _, idx = eng.tokenize_device(batch["wav"], batch["n_samples"], batch["ids"], batch["wid"], batch["lengths_host"])
    agg, idx = agg[:R].float().cpu(), idx[:R].cpu()
    rows = []
    for i in range(R):
        a_ref, a_got = ref[i]["_aggregated"][:T], agg[i, :T]
        ri, gi = ref[i]["quantized_indices"][:T], idx[i, :T]
        misses = []
        for t in np.nonzero((ri != gi).any(-1).numpy())[0]:
            q = int(np.argmax((ri[t] != gi[t]).numpy()))
            r = a_ref[t].double() @ Win.T + bin_                       # the reference's residual entering level q
            for qq in range(q):
                r = r - codes[qq][ri[t, qq]]
            dA, dB = float((r - codes[q][ri[t, q]]).norm()), float((r - codes[q][gi[t, q]]).norm())
            misses.append({"t": int(t), "level": q, "margin_rel": (dB - dA) / dA})
        rows.append({
            "h_last_rel": rel(h_last[i], ref[i]["_h_last"]), "h_target_rel": rel(h_t[i], ref[i]["_h_target"]),
            "aggregator_rel": rel(a_got, a_ref),
            "aggregator_centred_rel": rel(a_got - a_got.mean(0, keepdim=True), a_ref - a_ref.mean(0, keepdim=True)),
            "token_varying_share_of_norm": float((a_ref - a_ref.mean(0, keepdim=True)).norm() / a_ref.norm()),
            "index_agreement": float((ri == gi).float().mean()), "tokens_differing": len(misses), "misses": misses})
    report[prec] = rows
    del tower, eng
    torch.cuda.empty_cache()
print(json.dumps(report, indent=1))
