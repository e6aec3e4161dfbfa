"""Same-box yardstick for the encoder attention kernel (comparison only, never on the product path): torch SDPA with the
cuDNN / flash / efficient backends and flash-attn 2.x on B x 20 heads x 1500 x 64, bf16, non-causal, beside
taste_attention_bf16.  python scripts/attn_yardstick.py [B] > profiles/r2_attn_yardstick.json"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taste_spokenlm_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S, H, hd = 1500, 20, 64
D = H * hd
flops = 4.0 * B * H * S * S * hd
torch.manual_seed(0)
qkv = (torch.randn(B, S, 3, H, hd, device="cuda") * 0.7).bfloat16()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {"shape": dict(batch=B, heads=H, seq=S, head_dim=hd, dtype="bf16", causal=False), "flops": flops, "kernels": {}}


def record(name, fn):
    try:
        ms = timeit(fn)
        res["kernels"][name] = dict(ms=ms, tflops=flops / ms / 1e9)
    except Exception as e:  # noqa: BLE001
        res["kernels"][name] = dict(error=f"{type(e).__name__}: {str(e)[:200]}")
    print(name, res["kernels"][name], file=sys.stderr, flush=True)


lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flat = qkv.view(B * S, 3 * D)
q2, k2, v2 = flat[:, :D], flat[:, D:2 * D], flat[:, 2 * D:]
o2 = torch.empty(B * S, D, device="cuda", dtype=torch.bfloat16)
record("taste_attention_bf16 (this repo, tcgen05)",
       lambda: _lib.check(lib.taste_attention_bf16(_lib.ptr(q2), _lib.ptr(k2), _lib.ptr(v2), _lib.ptr(o2), 3 * D, 3 * D,
                                                   3 * D, D, None, None, S, S, B, H, 0, st), "attn"))
q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))         # [B, H, S, hd] views
from torch.nn.attention import SDPBackend, sdpa_kernel
for nm, be in (("torch SDPA cuDNN", SDPBackend.CUDNN_ATTENTION), ("torch SDPA flash", SDPBackend.FLASH_ATTENTION),
               ("torch SDPA efficient", SDPBackend.EFFICIENT_ATTENTION)):
    def f(be=be):
        with sdpa_kernel(be):
            return torch.nn.functional.scaled_dot_product_attention(q, k, v)
    record(nm, f)
try:
    from flash_attn import flash_attn_func
    import flash_attn
    qf, kf, vf = (qkv[:, :, i] for i in range(3))                   # [B, S, H, hd]
    record(f"flash_attn {flash_attn.__version__}", lambda: flash_attn_func(qf, kf, vf))
except Exception as e:  # noqa: BLE001
    res["kernels"]["flash_attn"] = dict(error=str(e)[:200])
# correctness of ours against SDPA (math in fp32)
ref = torch.nn.functional.scaled_dot_product_attention(q[:2].float(), k[:2].float(), v[:2].float())
# taste_attention expects q pre-scaled (CW:342): rerun on pre-scaled q for the comparison
qs = (flat[: 2 * S, :D].float() * hd ** -0.5).bfloat16().contiguous()
o3 = torch.empty(2 * S, D, device="cuda", dtype=torch.bfloat16)
kk, vv = flat[: 2 * S, D:2 * D], flat[: 2 * S, 2 * D:]
_lib.check(lib.taste_attention_bf16(_lib.ptr(qs), _lib.ptr(kk), _lib.ptr(vv), _lib.ptr(o3), D, 3 * D, 3 * D, D, None, None,
                                    S, S, 2, H, 0, st), "attn")
got = o3.view(2, S, H, hd).transpose(1, 2).float()
res["taste_vs_sdpa_fp32_rel_l2"] = float((got - ref).norm() / ref.norm())
print(json.dumps(res, indent=1))
