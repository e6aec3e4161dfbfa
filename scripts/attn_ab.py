"""A/B of the encoder attention kernel variants (TASTE_FA_VAR: 0 = 168 registers, free-running tiles; 10 = + setmaxnreg
232/40; 26 = + strict ping-pong; unset = default: ping-pong with early hand-over; TASTE_FA_POLY=0: all exp2 on the MUFU).
python scripts/attn_ab.py [B]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taste_spokenlm_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S, H, D = 1500, 20, 1280
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
qkv = (torch.randn(B * S, 3 * D, device="cuda") * 0.7).bfloat16()
# a high-variance case too: rows whose maximum keeps growing exercise the rescale path
qkv2 = qkv.clone()
qkv2[:, :D] *= 4.0
o = torch.zeros(B * S, D, device="cuda", dtype=torch.bfloat16)


def run(x, n):
    q, k, v = x[:, :D], x[:, D:2 * D], x[:, 2 * D:]
    for _ in range(n):
        _lib.check(lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), 3 * D, 3 * D, 3 * D, D,
                                            None, None, S, S, B, H, 0, st), "attn")


lib.taste_attention_set_mode(1)
refs = []
for x in (qkv, qkv2):
    run(x, 1)
    torch.cuda.synchronize()
    refs.append(o.float().clone())
lib.taste_attention_set_mode(0)
for var, poly in ((0, None), (10, None), (26, None), (-1, None), (-1, 0), (0, None), (-1, None)):
    if var < 0:
        os.environ.pop("TASTE_FA_VAR", None)          # the default variant
    else:
        os.environ["TASTE_FA_VAR"] = str(var)
    if poly is None:
        os.environ.pop("TASTE_FA_POLY", None)
    else:
        os.environ["TASTE_FA_POLY"] = str(poly)
    errs = []
    for x, ref in zip((qkv, qkv2), refs):
        o.zero_()
        run(x, 1)
        torch.cuda.synchronize()
        errs.append(float((o.float() - ref).norm() / ref.norm()))
    run(qkv, 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(qkv, 10)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"var {var} poly {poly}: {ms:.3f} ms  {4.0*B*H*S*S*64/ms/1e9:.1f} TF/s  rel diff vs mma.sync {errs[0]:.2e} {errs[1]:.2e}", flush=True)
