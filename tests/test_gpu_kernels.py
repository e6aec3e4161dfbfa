"""Kernel-level parity (-m gpu): every CUDA kernel, called through the C ABI, against the CPU oracle / a plain torch
fp32 statement of the same op / the committed reference fixtures.  Tolerances are stated per test."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import taste_oracle as O                      # checker only
from taste_spokenlm_b200 import _lib, synth


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.fixture(scope="module")
def lib(built_lib):
    assert torch.cuda.is_available()
    return built_lib


@pytest.fixture(scope="module")
def rvq_engine(lib):
    from taste_spokenlm_b200.engine import TowerEngine
    cfg = synth.TowerConfig(enc_layers=0, dec_layers=0, vocab=8)     # full-size RVQ, no encoder/decoder weights
    W = synth.random_weights(cfg, 77)                                # vq.* values do not depend on the other keys
    eng = TowerEngine(cfg, "cuda:0")
    eng.pack(W)
    return eng, W


# ---------------------------------------------------------------------------------------------------------------
# GEMM (tcgen05)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (256, 256, 128), (1500, 1280, 1280), (35, 384, 128),
                                   (3000, 3840, 1280), (777, 5120, 1280), (20000, 1280, 5120), (129, 256, 192),
                                   (19000, 256, 64), (9601, 2560, 320)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
@pytest.mark.parametrize("mode", [0, 1])
def test_gemm(lib, m, n, k, epi, mode):
    """mode 0: automatic tile shape (CTA pairs for the large cases); mode 1: single-CTA kernel everywhere."""
    if mode == 1 and m < 3000:
        pytest.skip("small shapes already use the single-CTA kernel in mode 0")
    assert lib.taste_gemm_set_mode(mode) == 0
    torch.manual_seed(m * 7 + n * 3 + k + epi)
    a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(n, k, device="cuda") / math.sqrt(k)).bfloat16()
    bias = torch.randn(n, device="cuda") * 0.1
    ref = a.float() @ w.float().T + bias
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    # guard bands of 32 rows around the output catch out-of-bounds stores of the partial last tile (compute-sanitizer is
    # closed on this pool)
    G = 32
    if epi in (0, 1):
        buf = torch.full((m + 2 * G, n), float("nan"), device="cuda").bfloat16()
    else:
        buf = torch.full((m + 2 * G, n), float("nan"), device="cuda")
    buf[:G] = 7.0
    buf[G + m:] = 7.0
    out = buf[G:G + m]
    if epi == 2:
        res = torch.randn(m, n, device="cuda")
        out.copy_(res)
        ref = ref + res
    _lib.check(lib.taste_gemm_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), m, n, k, epi, _stream()),
               "gemm")
    torch.cuda.synchronize()
    lib.taste_gemm_set_mode(0)
    assert torch.isfinite(out.float()).all()
    assert bool((buf[:G] == 7.0).all()) and bool((buf[G + m:] == 7.0).all()), "store outside the output rows"
    tol = 4e-3 if epi in (0, 1) else 2e-5 * math.sqrt(k)       # bf16 output rounding vs fp32 accumulate-order noise
    assert _rel(out.float(), ref) < tol


@pytest.mark.parametrize("m,d,n,dist", [(2048, 1280, 3840, "normal"), (3000, 256, 512, "normal"), (5000, 1280, 5120, "normal"),
                                        (2048, 1280, 3840, "outliers")])
@pytest.mark.parametrize("gelu", [0, 1])
def test_gemm_layernorm_folding(lib, m, d, n, gelu, dist):
    """Producer GEMM (residual epilogue) emits h, bf16(h) and per-128-column row statistics; consumer GEMM on the RAW
    bf16 rows with gamma-folded weights reproduces Linear(LayerNorm(h)) (include/taste_b200.h: taste_gemm_ex).

    `outliers`: a Whisper-like residual stream - a few massive-activation channels (+-300 against a unit
    bulk) and a common-mode row offset of half a standard deviation.  The fold rounds the RAW rows to bf16 and removes
    the mean after the GEMM, so its error grows with |row mean| / row std; it must stay within the same budget as
    LayerNorm -> bf16 -> GEMM here.  (Rows whose mean dwarfs their spread need `taste_encoder_set_mode(1)`.)"""
    torch.manual_seed(m + d + n + gelu)
    kk = 256
    a = (torch.randn(m, kk, device="cuda") * 0.5).bfloat16()
    wo = (torch.randn(d, kk, device="cuda") / math.sqrt(kk)).bfloat16()
    bo = torch.randn(d, device="cuda") * 0.1
    h0 = torch.randn(m, d, device="cuda") * 2.0 + 0.3
    if dist == "outliers":
        h0 = torch.randn(m, d, device="cuda")
        for c, v in ((7, 300.0), (500, -260.0), (777, 120.0), (1200, -90.0)):
            h0[:, c] = v * (1.0 + 0.05 * torch.randn(m, device="cuda"))
        h0 = h0 + 0.5 * h0.std(dim=1, keepdim=True)
    h = h0.clone()
    hb_buf = torch.full((m + 64, d), 7.0, device="cuda").bfloat16()
    hb = hb_buf[32:32 + m]
    hb.fill_(float("nan"))
    st_buf = torch.full((m + 64, d // 128, 2), 7.0, device="cuda")
    stats = st_buf[32:32 + m]
    stats.fill_(float("nan"))
    g = _lib.GemmEx(a=a.data_ptr(), w=wo.data_ptr(), bias=bo.data_ptr(), out=h.data_ptr(), m=m, n=d, k=kk, epilogue=2,
                    stats_out=stats.data_ptr(), out_bf16=hb.data_ptr())
    _lib.check(lib.taste_gemm_ex(C.byref(g), _stream()), "gemm_ex producer")
    torch.cuda.synchronize()
    h_ref = h0 + a.float() @ wo.float().T + bo
    assert _rel(h, h_ref) < 1e-5
    assert torch.equal(hb, h.bfloat16())
    for gb in (hb_buf, st_buf):
        assert bool((gb[:32] == 7.0).all()) and bool((gb[32 + m:] == 7.0).all()), "store outside the output rows"
    seg = h.view(m, d // 128, 128)
    assert _rel(stats[..., 0], seg.sum(-1)) < 1e-5 and _rel(stats[..., 1], (seg * seg).sum(-1)) < 1e-5
    # consumer
    gamma = 1.0 + 0.2 * torch.randn(d, device="cuda")
    beta = 0.1 * torch.randn(d, device="cuda")
    w = torch.randn(n, d, device="cuda") / math.sqrt(d)
    b = torch.randn(n, device="cuda") * 0.1
    wf = (w * gamma[None, :]).bfloat16()
    colsum = wf.float().sum(1)
    bp = b + w @ beta
    out = torch.full((m, n), float("nan"), device="cuda").bfloat16()
    g2 = _lib.GemmEx(a=hb.data_ptr(), w=wf.data_ptr(), bias=bp.data_ptr(), out=out.data_ptr(), m=m, n=n, k=d,
                     epilogue=gelu, ln_stats=stats.data_ptr(), ln_nseg=d // 128, ln_colsum=colsum.data_ptr())
    _lib.check(lib.taste_gemm_ex(C.byref(g2), _stream()), "gemm_ex consumer")
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(h.double(), (d,), gamma.double(), beta.double(), 1e-5) @ w.double().T + b.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert torch.isfinite(out.float()).all()
    assert _rel(out.float(), ref) < 6e-3          # bf16 A / W' / output rounding, same budget as LayerNorm -> bf16 -> GEMM


def test_gemm_rejects_bad_shapes(lib):
    a = torch.zeros(128, 96, device="cuda").bfloat16()
    w = torch.zeros(128, 96, device="cuda").bfloat16()
    out = torch.zeros(128, 128, device="cuda").bfloat16()
    rc = lib.taste_gemm_bf16(_lib.ptr(a), _lib.ptr(w), None, _lib.ptr(out), 128, 128, 96, 0, _stream())
    assert rc == -2 and b"gemm" in lib.taste_last_error()
    assert lib.taste_gemm_bf16(None, _lib.ptr(w), None, _lib.ptr(out), 128, 128, 64, 0, _stream()) == -1


# ---------------------------------------------------------------------------------------------------------------
# LayerNorm
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,d", [(1, 128), (37, 384), (3000, 1280), (5, 2048)])
@pytest.mark.parametrize("bf16", [0, 1])
def test_layernorm(lib, rows, d, bf16):
    torch.manual_seed(rows + d)
    x = torch.randn(rows, d, device="cuda") * 3 + 1.5
    w = torch.randn(d, device="cuda")
    b = torch.randn(d, device="cuda")
    y = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16 if bf16 else torch.float32)
    _lib.check(lib.taste_layernorm_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), rows, d, bf16, _stream()), "ln")
    ref = torch.nn.functional.layer_norm(x.double(), (d,), w.double(), b.double(), 1e-5)
    assert _rel(y, ref) < (4e-3 if bf16 else 2e-6)


# ---------------------------------------------------------------------------------------------------------------
# attention (segmented flash attention)
# ---------------------------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, causal):
    s = q.double() @ k.double().transpose(-1, -2)
    if causal:
        Tq, Tk = s.shape[-2:]
        s = s + torch.full((Tq, Tk), float("-inf"), device=s.device, dtype=s.dtype).triu(1)
    return torch.softmax(s, -1) @ v.double()


@pytest.mark.parametrize("B,S,H", [(1, 64, 1), (2, 1500, 3), (3, 200, 2), (2, 65, 2), (1, 256, 1), (2, 300, 2),
                                   (1, 1536, 2), (3, 1000, 4), (1, 257, 1),
                                   (40, 400, 5), (9, 777, 6)])      # more work items than SMs: the persistent walk
@pytest.mark.parametrize("mode", [0, 1])
def test_attention_fixed(lib, B, S, H, mode):
    """mode 0: tcgen05/TMEM kernel when S >= 256 (encoder shape), mode 1: mma.sync kernel everywhere."""
    if mode == 1 and S < 256:
        pytest.skip("already the mma.sync kernel in mode 0")
    assert lib.taste_attention_set_mode(mode) == 0
    torch.manual_seed(B * 100 + S)
    D = H * 64
    qkv = (torch.randn(B * S, 3 * D, device="cuda") * 0.7).bfloat16()
    obuf = torch.full((B * S + 64, D), float("nan"), device="cuda").bfloat16()
    obuf[:32] = 7.0
    obuf[32 + B * S:] = 7.0
    o = obuf[32:32 + B * S]
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    _lib.check(lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), 3 * D, 3 * D, 3 * D, D,
                                        None, None, S, S, B, H, 0, _stream()), "attn")
    torch.cuda.synchronize()
    lib.taste_attention_set_mode(0)
    assert bool((obuf[:32] == 7.0).all()) and bool((obuf[32 + B * S:] == 7.0).all()), "store outside the output rows"
    qh = q.float().view(B, S, H, 64).transpose(1, 2)
    kh = k.float().view(B, S, H, 64).transpose(1, 2)
    vh = v.float().view(B, S, H, 64).transpose(1, 2)
    ref = _attn_ref(qh, kh, vh, False).transpose(1, 2).reshape(B * S, D)
    assert torch.isfinite(o.float()).all()
    assert _rel(o.float(), ref) < 6e-3          # bf16 P and bf16 output rounding


@pytest.mark.parametrize("S", [1500, 300])                   # 3 x 64 free-running tiles / 2 x 128 ping-pong tiles
@pytest.mark.parametrize("jump", [0.0, 30.0, 250.0])
def test_attention_score_range(lib, S, jump):
    """Encoder kernel on rows whose scores GROW along the key axis: later key blocks exceed the running maximum by `jump`
    natural units (lazy rescale: the stale maximum may lag by up to 2^8).  Some queries also see strongly negative scores
    (argument range of the polynomial exp2)."""
    torch.manual_seed(11)
    B, H = 2, 2
    D = H * 64
    qkv = (torch.randn(B * S, 3 * D, device="cuda") * 0.7)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    # a shared direction: score offset = qa * kb[key]; kb steps up across the 128-key blocks
    pos = torch.arange(S, device="cuda").repeat(B)
    step = torch.zeros(B * S, device="cuda")
    step[pos >= S // 5] = 0.2
    step[pos >= S // 2] = 0.5
    step[pos >= (3 * S) // 4] = 1.0
    step[pos >= (14 * S) // 15] = 0.6         # and down again: the last block must not dominate
    qa = torch.full((B * S,), 4.0, device="cuda")
    qa[::7] = -4.0                            # these queries prefer the EARLY keys (later scores strongly negative)
    for h in range(H):
        q[:, h * 64] = qa
        k[:, h * 64] = step * (jump / 4.0)
    qkv = qkv.bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    o = torch.full((B * S, D), float("nan"), device="cuda").bfloat16()
    assert lib.taste_attention_set_mode(0) == 0
    _lib.check(lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), 3 * D, 3 * D, 3 * D, D,
                                        None, None, S, S, B, H, 0, _stream()), "attn")
    torch.cuda.synchronize()
    qh = q.float().view(B, S, H, 64).transpose(1, 2)
    kh = k.float().view(B, S, H, 64).transpose(1, 2)
    vh = v.float().view(B, S, H, 64).transpose(1, 2)
    ref = _attn_ref(qh, kh, vh, False).transpose(1, 2).reshape(B * S, D)
    assert torch.isfinite(o.float()).all()
    assert _rel(o.float(), ref) < 6e-3
    row_err = (o.float() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-6)
    assert float(row_err.max()) < 3e-2, int(row_err.argmax())       # no single row may be off (a missed rescale)


@pytest.mark.parametrize("causal", [0, 1])
def test_attention_ragged(lib, causal):
    torch.manual_seed(5 + causal)
    H, D = 2, 128
    qlens = [35, 1, 130, 64]
    kvlens = qlens if causal else [1500, 77, 64, 129]
    cu_q = torch.tensor([0] + list(np.cumsum(qlens)), dtype=torch.int32, device="cuda")
    cu_kv = torch.tensor([0] + list(np.cumsum(kvlens)), dtype=torch.int32, device="cuda")
    q = (torch.randn(sum(qlens), D, device="cuda") * 0.7).bfloat16()
    k = (torch.randn(sum(kvlens), D, device="cuda") * 0.7).bfloat16()
    v = (torch.randn(sum(kvlens), D, device="cuda")).bfloat16()
    o = torch.full((sum(qlens), D), float("nan"), device="cuda").bfloat16()
    _lib.check(lib.taste_attention_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), D, D, D, D, _lib.ptr(cu_q),
                                        _lib.ptr(cu_kv), max(qlens), max(kvlens), len(qlens), H, causal, _stream()), "attn")
    assert torch.isfinite(o.float()).all()
    for b in range(len(qlens)):
        qs, ks = int(cu_q[b]), int(cu_kv[b])
        qb = q[qs:qs + qlens[b]].float().view(-1, H, 64).transpose(0, 1)
        kb = k[ks:ks + kvlens[b]].float().view(-1, H, 64).transpose(0, 1)
        vb = v[ks:ks + kvlens[b]].float().view(-1, H, 64).transpose(0, 1)
        ref = _attn_ref(qb, kb, vb, bool(causal)).transpose(0, 1).reshape(qlens[b], D)
        assert _rel(o[qs:qs + qlens[b]].float(), ref) < 6e-3, b


@pytest.mark.parametrize("qlens,kv", [([69, 5, 130, 448, 1, 128, 129], 1500), ([70] * 40, 1500), ([33, 200], 200),
                                      ([448, 447, 3], 129)])
def test_attention_ragged_cross_tcgen05(lib, qlens, kv):
    """The aggregator's cross-attention (ragged packed queries, fixed-length keys, no mask) on the tcgen05 / TMA kernel
    (one 128-query tile per work item), against an fp64 statement and the mma.sync kernel; guard rows around the packed
    output catch stores past an utterance's last row."""
    torch.manual_seed(len(qlens) * 7 + kv)
    H = 3
    D = H * 64
    B, total = len(qlens), sum(qlens)
    cu_q = torch.tensor([0] + list(np.cumsum(qlens)), dtype=torch.int32, device="cuda")
    q = (torch.randn(total, D, device="cuda") * 0.7).bfloat16()
    k = (torch.randn(B * kv, D, device="cuda") * 0.7).bfloat16()
    v = torch.randn(B * kv, D, device="cuda").bfloat16()
    obuf = torch.full((total + 64, D), float("nan"), device="cuda").bfloat16()
    obuf[:32] = 7.0
    obuf[32 + total:] = 7.0
    o = obuf[32:32 + total]
    n0 = lib.taste_launch_count()
    _lib.check(lib.taste_attention_ragged_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), D, D, D, D, _lib.ptr(cu_q),
                                               total, max(qlens), kv, B, H, _stream()), "attn ragged")
    torch.cuda.synchronize()
    assert lib.taste_launch_count() == n0 + 1
    assert bool((obuf[:32] == 7.0).all()) and bool((obuf[32 + total:] == 7.0).all()), "store outside the packed rows"
    assert torch.isfinite(o.float()).all()
    o2 = torch.full_like(o, float("nan"))
    assert lib.taste_attention_set_mode(1) == 0
    _lib.check(lib.taste_attention_ragged_bf16(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o2), D, D, D, D, _lib.ptr(cu_q),
                                               total, max(qlens), kv, B, H, _stream()), "attn ragged mma")
    lib.taste_attention_set_mode(0)
    for b in range(B):
        qs = int(cu_q[b])
        qb = q[qs:qs + qlens[b]].float().view(-1, H, 64).transpose(0, 1)
        kb = k[b * kv:(b + 1) * kv].float().view(-1, H, 64).transpose(0, 1)
        vb = v[b * kv:(b + 1) * kv].float().view(-1, H, 64).transpose(0, 1)
        ref = _attn_ref(qb, kb, vb, False).transpose(0, 1).reshape(qlens[b], D)
        assert _rel(o[qs:qs + qlens[b]].float(), ref) < 6e-3, b
        assert _rel(o2[qs:qs + qlens[b]].float(), ref) < 6e-3, b


# ---------------------------------------------------------------------------------------------------------------
# log-mel front-end:  <= 1e-4 relative L2 (north star), checked against the oracle AND the reference fixtures
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(params=[0, 1], ids=["tensor_dft", "fma_dft"])
def logmel_mode(lib, request):
    """Both formulations of the front-end DFT (taste_logmel_set_mode): split-bf16 GEMM on tcgen05, fp32 FMA kernel."""
    assert lib.taste_logmel_set_mode(request.param) == 0
    yield request.param
    lib.taste_logmel_set_mode(0)


def test_logmel_vs_oracle_and_reference(lib, golden_dir, logmel_mode):
    from taste_spokenlm_b200.frontend import WhisperFrontendB200
    fe = WhisperFrontendB200(whisper_model="large-v3", do_pad_trim=True, permute=True)
    z = np.load(os.path.join(golden_dir, "frontend.npz"))
    for nm, seed, n in json.loads(str(z["meta"])):
        wav = synth.synth_waveform(seed, n)[None]
        feats, lens = fe(wav, torch.tensor([n]))
        assert feats.device.type == "cpu" and feats.shape == (1, 3000, 128)
        ref, rlens = O.log_mel(wav, [n])
        assert _rel(feats, ref) < 1e-4, nm
        assert _rel(feats[0, ::25, :], torch.from_numpy(z[f"{nm}_sub"])) < 1e-4, nm
        assert int(lens[0]) == int(rlens[0]) == int(z[f"{nm}_len"][0])
    feats, _ = fe(torch.zeros(2, 8000), torch.tensor([8000, 8000]))
    assert torch.equal(feats, torch.full_like(feats, -1.5))                  # silence: every bin at the 1e-10 clamp


def test_logmel_batch_ragged_and_bf16(lib, logmel_mode):
    from taste_spokenlm_b200.engine import FrontendEngine
    eng = FrontendEngine("cuda:0")
    b = synth.synth_batch(3, [1.0, 30.0, 12.34, 0.01], [1, 1, 1, 1], pad_wave_to=480000)
    f32, b16 = eng.logmel(b["wav"].cuda(), b["n_samples"].cuda(), True, True)
    ref, _ = O.log_mel(b["wav"])
    assert _rel(f32, ref) < 1e-4
    assert _rel(b16.float(), ref) < 4e-3
    # batch independence: each row equals its solo run bit for bit
    for i in range(4):
        solo, _ = eng.logmel(b["wav"][i:i + 1].cuda(), b["n_samples"][i:i + 1].cuda(), True, False)
        assert torch.equal(solo[0], f32[i])


# ---------------------------------------------------------------------------------------------------------------
# RVQ: indices bit-exact against the fp32 reference (fixtures) except fp64-verified near ties
# ---------------------------------------------------------------------------------------------------------------
def _near_tie(W, x_in, idx_a, idx_b, q_level, rel_gap=2e-6):
    """fp64: are the two candidate codes equidistant from the level-q residual up to fp32 rounding?"""
    r = x_in.double()
    for q in range(q_level):
        r = r - W[f"vq.rvq.layers.{q}._codebook.embed"][0].double()[idx_a[q]]
    e = W[f"vq.rvq.layers.{q_level}._codebook.embed"][0].double()
    da, db = (r - e[idx_a[q_level]]).norm(), (r - e[idx_b[q_level]]).norm()
    return abs(float(da - db)) <= rel_gap * float(max(da, db))


def test_rvq_bit_exact_vs_reference(lib, rvq_engine, golden_dir):
    eng, W = rvq_engine
    z = np.load(os.path.join(golden_dir, "rvq.npz"))
    meta = json.loads(str(z["meta"]))
    g = torch.Generator().manual_seed(meta["z_seed"])
    x = torch.randn(3, 50, 1280, generator=g)
    x = x + 0.7 * torch.randn(1, 1, 1280, generator=g)
    lens = torch.tensor(meta["lens"], dtype=torch.int32)
    qz, idx = eng.rvq_encode(x.cuda(), lens.cuda())
    idx, qz = idx.cpu(), qz.cpu()
    ridx = torch.from_numpy(z["quantized_indices"])
    assert torch.equal(idx < 0, ridx < 0)
    x_in = x @ W["vq.rvq.project_in.weight"].T + W["vq.rvq.project_in.bias"]
    mism = (idx != ridx).any(-1).nonzero()
    ties = []
    for b, t in mism.tolist():
        ql = int((idx[b, t] != ridx[b, t]).nonzero()[0])
        assert _near_tie(W, x_in[b, t], ridx[b, t], idx[b, t], ql), (b, t, idx[b, t], ridx[b, t])
        ties.append((b, t, ql))
    print("fp64-verified near ties:", ties)
    assert len(ties) <= 2
    same = (idx == ridx).all(-1)
    assert _rel(qz[same], torch.from_numpy(z["quantized_feats"])[same]) < 1e-5
    # decode paths (RVQ:183-242) on the reference's own indices
    out = eng.rvq_decode(ridx.cuda(), True).cpu()
    assert _rel(out, torch.from_numpy(z["output_from_indices"])) < 1e-6
    code = eng.rvq_decode(ridx.cuda(), False).cpu()
    assert _rel(code, torch.from_numpy(z["code_from_indices"])) < 1e-6
    # get_indices_from_code (RVQ:258-357)
    _, idx2 = eng.rvq_encode(x_in.cuda().contiguous(), lens.cuda(), want_quantized=False)
    r2 = torch.from_numpy(z["indices_from_code"])
    idx2 = idx2.cpu()
    assert torch.equal(idx2 < 0, r2 < 0)
    for b, t in (idx2 != r2).any(-1).nonzero().tolist():          # exact, or an fp64-verified tie at the first divergence
        ql = int((idx2[b, t] != r2[b, t]).nonzero()[0])
        assert _near_tie(W, x_in[b, t], r2[b, t], idx2[b, t], ql), (b, t, idx2[b, t], r2[b, t])


def test_rvq_random_vs_oracle(lib, rvq_engine):
    eng, W = rvq_engine
    g = torch.Generator().manual_seed(123)
    x = torch.randn(4, 33, 1280, generator=g) * 1.3
    lens = torch.tensor([33, 1, 20, 0], dtype=torch.int32)
    mask = torch.arange(33)[None] < lens[:, None]
    qz, idx = eng.rvq_encode(x.cuda(), lens.cuda())
    rq, ridx = O.rvq_encode(W, x, mask)
    assert torch.equal(idx.cpu() < 0, ridx < 0)
    agree = (idx.cpu() == ridx).float().mean().item()
    assert agree > 0.995, agree
    assert torch.allclose(qz.cpu()[3], W["vq.rvq.project_out.bias"].expand(33, -1))     # fully padded row


@pytest.mark.parametrize("scale,rows", [(1.0, 4096), (13.0, 1500), (0.02, 700)])
def test_rvq_search_survivor_paths_vs_oracle(lib, rvq_engine, scale, rows):
    """The tensor-core search keeps every code whose error-bounded score could be the minimum; one survivor is the answer,
    two are re-evaluated exactly by the token's thread, three or more by the whole warp.  |r| >> |e| (scale 13) widens the
    bound relative to the score spacing and drives hundreds of tokens through both exact paths; |r| << |e| (scale 0.02)
    makes the codes' own norms decide.  Indices must equal the fp32 oracle except fp64-verified ties."""
    eng, W = rvq_engine
    g = torch.Generator().manual_seed(int(scale * 100) + rows)
    code = torch.randn(1, rows, 256, generator=g) * scale               # get_indices_from_code path: the search alone
    _, idx = eng.rvq_encode(code.cuda(), None, want_quantized=False)
    _, ridx = O.rvq_encode(W, code, None, project_in=False)
    idx = idx.cpu()
    bad = (idx != ridx).any(-1).nonzero()
    for b, t in bad.tolist():
        ql = int((idx[b, t] != ridx[b, t]).nonzero()[0])
        assert _near_tie(W, code[b, t], ridx[b, t], idx[b, t], ql), (scale, t, idx[b, t], ridx[b, t])
    assert len(bad) <= max(2, rows // 500), len(bad)
    # with project_in / project_out around it
    z = torch.randn(3, 40, 1280, generator=g) * scale
    lens = torch.tensor([40, 7, 23], dtype=torch.int32)
    mask = torch.arange(40)[None] < lens[:, None]
    qz, idx2 = eng.rvq_encode(z.cuda(), lens.cuda())
    rq, ridx2 = O.rvq_encode(W, z, mask)
    assert torch.equal(idx2.cpu() < 0, ridx2 < 0)
    same = (idx2.cpu() == ridx2).all(-1)
    assert same.float().mean() > 0.97
    assert _rel(qz.cpu()[same], rq[same]) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# word pooling (JES:418-458 incl. the padded-row quirk) and the llm-token mapping
# ---------------------------------------------------------------------------------------------------------------
def test_word_pool_vs_oracle_and_reference_cases(lib, golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "word_pooling.json")))
    cases.append(dict(word_ids=[[0, 1, 1, 2, 3, 3, 3, 4, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                                [1, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 1, 2, 3, 4, 5, 6, 7, 8, 8]], lengths=[9, 3, 2, 11]))
    D = 128
    for c in cases:
        wid = torch.tensor(c["word_ids"], dtype=torch.int32)
        lens1 = torch.tensor(c["lengths"])                    # = T_b + 1 (JES:396)
        T = (lens1 - 1).clamp_min(0).to(torch.int32)
        B, Tmax = wid.shape
        if int(T.max()) > Tmax:
            continue
        rows = [int(t) + 5 for t in T]
        cu = torch.tensor([0] + list(np.cumsum(rows)), dtype=torch.int32)
        torch.manual_seed(B * 31 + Tmax)
        dec = torch.randn(int(cu[-1]), D)
        z = torch.empty(B, Tmax, D, device="cuda")
        dec_d, cu_d, wid_d, T_d = dec.cuda(), cu.cuda(), wid.cuda(), T.cuda()      # keep the device copies alive
        _lib.check(lib.taste_word_pool_f32(_lib.ptr(dec_d), _lib.ptr(cu_d), _lib.ptr(wid_d), _lib.ptr(T_d), B, Tmax, D,
                                           _lib.ptr(z), _stream()), "pool")
        torch.cuda.synchronize()
        # oracle on the padded [B, Tmax+1, D] view the reference sees (rows past T_b: whatever the decoder produced;
        # only row T_b can ever be pooled, and it exists in the packed layout)
        x = torch.zeros(B, Tmax + 1, D)
        for b in range(B):
            n = min(int(T[b]) + 1, Tmax + 1)
            x[b, :n] = dec[int(cu[b]) + 4: int(cu[b]) + 4 + n]
        ref = O.word_pool(x, wid, lens1)[:, :-1]
        for b in range(B):
            tb = int(T[b])
            assert torch.allclose(z[b, :tb].cpu(), ref[b, :tb], atol=1e-6), (c, b)
            assert float(z[b, tb:].abs().sum()) == 0.0


def test_map_to_llm_tokens(lib, golden_dir):
    z = np.load(os.path.join(golden_dir, "llm_mapping.npz"))
    idx = torch.from_numpy(z["asr_indices"]).cuda()
    B, T, Q = idx.shape
    L = z["llm_wid"].shape[1]
    out = torch.empty(B, L, Q, dtype=torch.int64, device="cuda")
    d = {k: torch.from_numpy(z[k].astype(np.int32)).cuda() for k in ("asr_wid", "asr_len", "llm_wid", "llm_len")}
    _lib.check(lib.taste_map_to_llm_tokens(_lib.ptr(idx), _lib.ptr(d["asr_wid"]), _lib.ptr(d["asr_len"]),
                                           _lib.ptr(d["llm_wid"]), _lib.ptr(d["llm_len"]), B, T, L, Q,
                                           _lib.ptr(out), _stream()), "map")
    assert np.array_equal(out.cpu().numpy(), z["llm_indices"])
