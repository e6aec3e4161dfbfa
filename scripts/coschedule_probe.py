"""Power-cap probe: does running the (power-capped) GEMMs on part of the SMs next to the (not power-capped) attention
kernel on the rest raise the aggregate throughput?  The step alternates ~1000 W GEMM phases at ~1.4 GHz with ~530 W
attention phases at ~1.85 GHz; co-scheduling would spend the attention phases' power headroom on GEMM clocks.
  python scripts/coschedule_probe.py [gemm_pairs] [n_gemm] [n_attn]
Sequential: n_gemm fc1-shaped GEMMs + n_attn encoder attention launches on one stream at full size.
Concurrent: the same work on two streams, GEMM limited to `gemm_pairs` CTA pairs, attention to the other SMs."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taste_spokenlm_b200 import _lib
lib = _lib.load()
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 50
n_gemm = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n_attn = int(sys.argv[3]) if len(sys.argv) > 3 else 20
B, S, H, D, FF = 64, 1500, 20, 1280, 5120
M = B * S
a = (torch.randn(M, D, device="cuda") * 0.5).bfloat16()
w = (torch.randn(FF, D, device="cuda") * 0.03).bfloat16()
bias = torch.randn(FF, device="cuda") * 0.1
mid = torch.zeros(M, FF, device="cuda", dtype=torch.bfloat16)
qkv = (torch.randn(M, 3 * D, device="cuda") * 0.7).bfloat16()
o = torch.zeros(M, D, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]


def gemm(stream):
    _lib.check(lib.taste_gemm_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(mid), M, FF, D, 1,
                                   C.c_void_p(stream.cuda_stream)), "gemm")


def attn(stream):
    _lib.check(lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), 3 * D, 3 * D, 3 * D, D, None,
                                        None, S, S, B, H, 0, C.c_void_p(stream.cuda_stream)), "attn")


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(concurrent):
    if concurrent:
        os.environ["TASTE_GEMM_MAX_PAIRS"] = str(pairs)
        os.environ["TASTE_FA_MAX_CTAS"] = str(148 - 2 * pairs)
    else:
        os.environ.pop("TASTE_GEMM_MAX_PAIRS", None)
        os.environ.pop("TASTE_FA_MAX_CTAS", None)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(torch.cuda.current_stream())
    s1.wait_event(e0)
    s2.wait_event(e0)
    if concurrent:
        for _ in range(n_gemm):
            gemm(s1)
        for _ in range(n_attn):
            attn(s2)
    else:
        done = 0
        for i in range(n_attn):              # interleaved like the encoder layers
            want = (i + 1) * n_gemm // n_attn
            for _ in range(want - done):
                gemm(s1)
            done = want
            attn(s1)
    e1.record(s1)
    e2.record(s2)
    torch.cuda.synchronize()
    return max(e0.elapsed_time(e1), e0.elapsed_time(e2)), e0.elapsed_time(e1), e0.elapsed_time(e2)


for c in (False, True, False, True):
    run(c)                                   # warm-up (power state)
    t, t1, t2 = run(c)
    print(f"{'concurrent' if c else 'sequential'}: {t:.2f} ms  (gemm stream {t1:.2f}, attention stream {t2:.2f})"
          f"  pairs {pairs if c else 74} / attention CTAs {148 - 2 * pairs if c else 148}", flush=True)
