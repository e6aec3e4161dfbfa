"""LayerNorm-folded GEMM variants vs the plain ones (CUDA events).  python scripts/gemm_ln_bench.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taste_spokenlm_b200 import _lib
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
M, D, FF = 96000, 1280, 5120
ONLY = os.environ.get("ONLY", "")
QUICK = bool(ONLY)
def run(name, g):
    if ONLY and not name.startswith(ONLY): return
    if QUICK:
        _lib.check(lib.taste_gemm_ex(C.byref(g), st()), name); _lib.check(lib.taste_gemm_ex(C.byref(g), st()), name)
        torch.cuda.synchronize(); print(name); return
    for _ in range(3): _lib.check(lib.taste_gemm_ex(C.byref(g), st()), name)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): lib.taste_gemm_ex(C.byref(g), st())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:28s} {ms:7.3f} ms  {2.0*g.m*g.n*g.k/ms/1e9:7.1f} TF/s", flush=True)
def t(*shape, dt=torch.bfloat16, scale=0.5): return (torch.randn(*shape, device="cuda") * scale).to(dt)
stats = torch.rand(M, D // 128, 2, device="cuda") * 128 + 100
stats[..., 1] = stats[..., 0] ** 2 / 128 + 128
hb = t(M, D); h = t(M, D, dt=torch.float32)
for name, n, k, epi in (("qkv", 3 * D, D, 0), ("fc1_gelu", FF, D, 1)):
    w = t(n, k, scale=0.03); b = t(n, dt=torch.float32, scale=0.1); c = t(n, dt=torch.float32, scale=0.1)
    out = torch.zeros(M, n, device="cuda", dtype=torch.bfloat16)
    run(name + " plain", _lib.GemmEx(a=hb.data_ptr(), w=w.data_ptr(), bias=b.data_ptr(), out=out.data_ptr(), m=M, n=n, k=k, epilogue=epi))
    run(name + " LN-in", _lib.GemmEx(a=hb.data_ptr(), w=w.data_ptr(), bias=b.data_ptr(), out=out.data_ptr(), m=M, n=n, k=k, epilogue=epi,
                                     ln_stats=stats.data_ptr(), ln_nseg=D // 128, ln_colsum=c.data_ptr()))
    del w, out
for name, k in (("out_proj", D), ("fc2", FF)):
    a = t(M, k); w = t(D, k, scale=0.03); b = t(D, dt=torch.float32, scale=0.1)
    so = torch.zeros(M, D // 128, 2, device="cuda"); ob = torch.zeros(M, D, device="cuda", dtype=torch.bfloat16)
    run(name + " resid plain", _lib.GemmEx(a=a.data_ptr(), w=w.data_ptr(), bias=b.data_ptr(), out=h.data_ptr(), m=M, n=D, k=k, epilogue=2))
    run(name + " resid LN-out", _lib.GemmEx(a=a.data_ptr(), w=w.data_ptr(), bias=b.data_ptr(), out=h.data_ptr(), m=M, n=D, k=k, epilogue=2,
                                            stats_out=so.data_ptr(), out_bf16=ob.data_ptr()))
    del a, w
