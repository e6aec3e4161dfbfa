"""Encoder attention micro-benchmark.  python scripts/attn_bench.py [B]
Times taste_attention_bf16 on B x 20 heads x 1500 x 64 (the encoder shape) and checks it against the mma.sync kernel.
(Round-2 A/B recorded in DESIGN.md: computing all of a block's exponentials before waiting for the previous block's
P V - 16 more live registers - measured 1.008 ms against 0.970 ms for the shipped order, same box, same run.)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taste_spokenlm_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
exps = [0]
S, H, D = 1500, 20, 1280
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
qkv = (torch.randn(B * S, 3 * D, device="cuda") * 0.7).bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]


def run(o):
    _lib.check(lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), 3 * D, 3 * D, 3 * D, D,
                                        None, None, S, S, B, H, 0, st), "attn")


ref = torch.zeros(B * S, D, device="cuda", dtype=torch.bfloat16)
lib.taste_attention_set_mode(1)
run(ref)
lib.taste_attention_set_mode(0)
torch.cuda.synchronize()
for rep in range(2):                       # two passes: the order of the variants must not matter
    for e in exps:
        o = torch.zeros_like(ref)
        for _ in range(3):
            run(o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run(o)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        rel = float((o.float() - ref.float()).norm() / ref.float().norm())
        print(f"pass {rep} variant {e}: {ms:.4f} ms  {4.0 * B * H * S * S * 64 / ms / 1e9:.1f} TFLOP/s  rel diff vs mma.sync {rel:.2e}", flush=True)
