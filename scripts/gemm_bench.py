"""GEMM micro-benchmark on the encoder's shapes (CUDA events, both tile modes).  python scripts/gemm_bench.py"""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taste_spokenlm_b200 import _lib
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
M = int(os.environ.get("M", 96000))
shapes = [("qkv", M, 3840, 1280, 0), ("out_proj", M, 1280, 1280, 2), ("fc1_gelu", M, 5120, 1280, 1), ("fc2", M, 1280, 5120, 2),
          ("f32out", M, 1280, 1280, 3)]
res = []
for name, m, n, k, epi in shapes:
    a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.03).bfloat16()
    bias = torch.randn(n, device="cuda") * 0.1
    out = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
    for mode in (0, 1):
        lib.taste_gemm_set_mode(mode)
        for _ in range(3):
            _lib.check(lib.taste_gemm_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), m, n, k, epi, st()), "gemm")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            lib.taste_gemm_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), m, n, k, epi, st())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = 2.0 * m * n * k / ms / 1e9
        res.append(dict(shape=name, m=m, n=n, k=k, epi=epi, mode="pair" if mode == 0 else "single", ms=ms, tflops=tf))
        print(f"{name:10s} {m}x{n}x{k} epi{epi} {'pair  ' if mode == 0 else 'single'} {ms:7.3f} ms {tf:7.1f} TF/s", flush=True)
    del a, w, out
lib.taste_gemm_set_mode(0)
# cuBLAS reference point for the same shape (library baseline, not the product path)
for name, m, n, k, epi in shapes[:4]:
    a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16(); w = (torch.randn(n, k, device="cuda") * 0.03).bfloat16()
    for _ in range(3): torch.matmul(a, w.T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): torch.matmul(a, w.T)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:10s} cuBLAS bf16 (no epilogue) {ms:7.3f} ms {2.0*m*n*k/ms/1e9:7.1f} TF/s", flush=True)
    res.append(dict(shape=name, mode="cublas", ms=ms, tflops=2.0*m*n*k/ms/1e9))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)
