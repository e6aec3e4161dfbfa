"""One GEMM shape, one mode, a few launches (for ncu).  python scripts/gemm_one.py <mode> <m> <n> <k> <epi> [reps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taste_spokenlm_b200 import _lib
lib = _lib.load()
mode, m, n, k, epi = [int(x) for x in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 4
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
w = (torch.randn(n, k, device="cuda") * 0.03).bfloat16()
bias = torch.randn(n, device="cuda") * 0.1
out = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
lib.taste_gemm_set_mode(mode)
for _ in range(reps):
    _lib.check(lib.taste_gemm_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), m, n, k, epi, st), "gemm")
torch.cuda.synchronize()
ref = a[:512].float() @ w.float().T + bias
print("ok", float((out[:512].float() - ref).norm() / ref.norm()))
