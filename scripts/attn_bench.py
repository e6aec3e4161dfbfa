"""Encoder attention micro-benchmark (both kernels).  python scripts/attn_bench.py [B]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taste_spokenlm_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S, H, D = 1500, 20, 1280
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
qkv = (torch.randn(B * S, 3 * D, device="cuda") * 0.7).bfloat16()
o = torch.zeros(B * S, D, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
outs = []
for mode in (0, 1):
    lib.taste_attention_set_mode(mode)
    for _ in range(2):
        _lib.check(lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), 3 * D, 3 * D, 3 * D, D,
                                            None, None, S, S, B, H, 0, st), "attn")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), 3 * D, 3 * D, 3 * D, D, None, None,
                                 S, S, B, H, 0, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if mode == 0 and os.environ.get("ATTN_CLOCKS"):
        # SM clock and board power under a sustained run of this kernel alone (is the micro-benchmark power-capped?)
        import pynvml, time
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        e0.record()
        for _ in range(400):
            lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), 3 * D, 3 * D, 3 * D, D, None,
                                     None, S, S, B, H, 0, st)
        e1.record()
        smp = []
        while not e1.query():
            smp.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.02)
        torch.cuda.synchronize()
        print(f"sustained: {e0.elapsed_time(e1) / 400:.3f} ms/launch; clocks MHz {sorted(c for c, _ in smp)[len(smp) // 2]}"
              f" (min {min(c for c, _ in smp)}, max {max(c for c, _ in smp)}), power W max {max(w for _, w in smp):.0f}", flush=True)
    print(f"mode {mode} ({'tcgen05' if mode == 0 else 'mma.sync'}): {ms:.3f} ms  {4.0*B*H*S*S*64/ms/1e9:.1f} TF/s", flush=True)
    outs.append(o.float().clone())
lib.taste_attention_set_mode(0)
print("rel diff between kernels", float((outs[0] - outs[1]).norm() / outs[1].norm()))
