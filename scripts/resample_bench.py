"""(f)2 ingest kernel alone: taste_resample_mean_f32 on a batch of 64 x 30 s decoded-PCM arrays (device-resident), against
its HBM contract 4 * (channels * n_in + n_out) bytes per utterance.  python scripts/resample_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from taste_spokenlm_b200 import ingest
rs = ingest.ResampleMeanB200("cuda:0")
peak = 6542.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
res = []
B = 64
for sr, ch in ((24000, 1), (24000, 2), (44100, 1), (48000, 2), (8000, 1)):
    n = sr * 30
    x = torch.randn(B * ch * n, device="cuda") * 0.1
    off = np.arange(B + 1, dtype=np.int64) * ch * n
    chs, nin = np.full(B, ch, np.int64), np.full(B, n, np.int64)
    out = torch.empty(B, 480000, device="cuda")
    for _ in range(3):
        rs.run_device(x, off, chs, nin, sr, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        rs.run_device(x, off, chs, nin, sr, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    byts = 4.0 * B * (ch * n + 480000)
    res.append(dict(orig_sr=sr, channels=ch, batch=B, ms=ms, algorithmic_GBps=byts / ms / 1e6, frac_of_hbm_peak=byts / ms / 1e6 / peak,
                    audio_s_per_s=B * 30.0 / (ms / 1e3)))
    print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(hbm_peak_GBps=peak, results=res), open("gpurun_out/r2_resample_bench.json", "w"), indent=1)
