"""tcgen05 attention: time vs number of key blocks (fixed overhead vs per-block cost).  python scripts/attn_sweep.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taste_spokenlm_b200 import _lib
lib = _lib.load()
B, H, D, SQ = 64, 20, 1280, 1536
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
q = (torch.randn(B * SQ, D, device="cuda") * 0.7).bfloat16()
o = torch.zeros(B * SQ, D, device="cuda", dtype=torch.bfloat16)
for SK in (128, 256, 512, 1024, 1536):
    k = (torch.randn(B * SK, D, device="cuda") * 0.7).bfloat16()
    v = torch.randn(B * SK, D, device="cuda").bfloat16()
    f = lambda: lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), D, D, D, D, None, None, SQ, SK, B, H, 0, st)
    for _ in range(2): _lib.check(f(), "attn")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ctas = B * H * SQ // 256
    print(f"kv {SK:5d} blocks {SK//128:2d}: {ms:.3f} ms  per-CTA {ms*1e3/(ctas/148):.2f} us  {4.0*B*H*SQ*SK*64/ms/1e9:.1f} TF/s", flush=True)
