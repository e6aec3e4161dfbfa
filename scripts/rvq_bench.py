"""RVQ encode micro-benchmark (realistic aggregator outputs and adversarially scaled inputs).  python scripts/rvq_bench.py"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from taste_spokenlm_b200 import synth
from taste_spokenlm_b200.tower import TasteAudioTowerB200
torch.set_grad_enabled(False)
cfg = synth.TowerConfig(enc_layers=8)
W = synth.random_weights(cfg, 1234)
tower = TasteAudioTowerB200.from_config(cfg).eval(); tower.load_state_dict(W, strict=True); tower = tower.to("cuda")
eng = tower.engine()
zf = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "tower_full_b34.npz"))
agg = torch.from_numpy(zf["aggregated_packed"]).cuda()                   # realistic aggregator outputs [2068, 1280]
for name, z in (("fixture x2 (4136 rows)", torch.cat([agg, agg])[None].contiguous()),
                ("randn (4096 rows)", torch.randn(64, 64, 1280, device="cuda")),
                ("randn*13 (4096 rows)", 13 * torch.randn(64, 64, 1280, device="cuda"))):
    lens = torch.full((z.shape[0],), z.shape[1], dtype=torch.int32, device="cuda")
    for wq in (True, False):
        for _ in range(2): eng.rvq_encode(z, lens, want_quantized=wq)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): qz, idx = eng.rvq_encode(z, lens, want_quantized=wq)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} quantized={wq}: {e0.elapsed_time(e1)/10:.3f} ms", flush=True)
