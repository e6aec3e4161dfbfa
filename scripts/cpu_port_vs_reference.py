"""One-off, build container only: the CPU oracle port timed beside the REAL reference on the same cores and the same
utterance (the bench's CPU arm is the port because /root/reference does not exist on the GPU box).
    PYTHONPATH=. python scripts/cpu_port_vs_reference.py > profiles/r2_cpu_port_vs_reference.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_shim, taste_oracle as O
from taste_spokenlm_b200 import synth
torch.set_grad_enabled(False)
cores = os.cpu_count() or 1
torch.set_num_threads(cores)
cfg = synth.FULL
W = synth.random_weights(cfg, 1234)
b = synth.synth_batch(11, [30.0], [64])
fe = ref_shim.build_reference_frontend()
tower = ref_shim.build_reference_tower()
tower.load_state_dict(W, strict=True)


def ref_step():
    n = int(b["n_samples"][0])
    feats, _ = fe(b["wav"][:, :n], torch.tensor([n]))                                  # WF:87-113
    return tower(b["asr_token_ids"], b["asr_token_lengths"], feats, torch.tensor([3000]),
                 asr_word_ids=b["asr_word_ids"])["quantized_indices"]                  # MT:108-211


def port_step():
    feats, _ = O.log_mel(b["wav"])
    return O.tower_forward(W, b["asr_token_ids"], b["asr_token_lengths"], feats, b["asr_word_ids"], cfg.heads,
                           cfg.enc_layers)["quantized_indices"]


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


t_ref, i_ref = timed(ref_step)
t_port, i_port = timed(port_step)
print(json.dumps({"cores": cores, "utterance": "30 s, 64 tokens, FULL geometry, fp32",
                  "reference_s_per_utt": t_ref, "port_s_per_utt": t_port,
                  "reference_audio_s_per_s": 30.0 / t_ref, "port_audio_s_per_s": 30.0 / t_port,
                  "index_agreement_port_vs_reference": float((i_ref == i_port).float().mean()),
                  "note": "same container, same threads, same utterance: the bench's CPU arm (the port) stands in for the "
                          "reference's own CPU path to within the ratio above"}, indent=1))
