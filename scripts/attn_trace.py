import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# usage: attn_trace.py [VAR [POLY]]   (VAR must include bit 2 = timeline trace; default 68 = split phases + trace)
os.environ["TASTE_FA_VAR"] = sys.argv[1] if len(sys.argv) > 1 else "68"
os.environ["TASTE_FA_POLY"] = sys.argv[2] if len(sys.argv) > 2 else "5"
import torch
from taste_spokenlm_b200 import _lib
lib = _lib.load()
B, S, H, D = 8, 1500, 20, 1280
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
qkv = (torch.randn(B * S, 3 * D, device="cuda") * 0.7).bfloat16()
o = torch.zeros(B * S, D, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
dbg = torch.zeros(6 * 256, dtype=torch.int64, device="cuda")
lib.taste_dbg_attention_trace(C.c_void_p(dbg.data_ptr()))
for _ in range(3):
    _lib.check(lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), 3 * D, 3 * D, 3 * D, D, None, None, S, S, B, H, 0, st), "attn")
torch.cuda.synchronize()
t = dbg.cpu().view(6, 32, 8)
t0 = int(t[0, 0, 0])
names_wg = ["enter", "s_full", "max_done", "s_free", "pre_o_wait", "o_full", "block_end", "out_o_full"]
names_mma = ["s_free", "qk_issued", "p_full", "pv_issued"]
for j in range(12, 26):      # second work item: steady state
    print(f"blk {j:2d} WG0:", " ".join(f"{n}={int(t[0,j,e])-t0}" for e, n in enumerate(names_wg)))
    print(f"       WG1:", " ".join(f"{n}={int(t[1,j,e])-t0}" for e, n in enumerate(names_wg)))
    print(f"      MMA0:", " ".join(f"{n}={int(t[2,j,e])-t0}" for e, n in enumerate(names_mma)))
    print(f"      MMA1:", " ".join(f"{n}={int(t[3,j,e])-t0}" for e, n in enumerate(names_mma)))
names_out = ["o_full", "setup", "ld0", "st0", "ld1", "st1"]
for it in range(1, 4):       # output phase of work items 1..3 (rows 4 / 5 of the trace buffer)
    for i in range(2):
        print(f"out it={it} WG{i}:", " ".join(f"{n}={int(t[4+i,it,e])-t0}" for e, n in enumerate(names_out)))
