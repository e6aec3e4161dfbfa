"""Size-independent properties of the CUDA path at BASELINE.json's FULL size (batch 64 x 30 s, distil-large-v3 geometry),
where the CPU oracle is too slow to be the checker: permutation equivariance, determinism, RVQ encode/decode round
trip, index range / padding, log-mel gain shift."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from taste_spokenlm_b200 import synth
from taste_spokenlm_b200.frontend import WhisperFrontendB200
from taste_spokenlm_b200.tower import TasteAudioTowerB200

torch.set_grad_enabled(False)
B, T = 64, 64


@pytest.fixture(scope="module")
def full(built_lib):
    cfg = synth.FULL
    tower = TasteAudioTowerB200.from_config(cfg).eval()
    tower.load_state_dict(synth.random_weights(cfg, 1234), strict=True)
    tower = tower.to("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(7)
    n = 480000
    t = torch.arange(n, device="cuda", dtype=torch.float32) / 16000.0
    wav = 0.01 * torch.randn(B, n, device="cuda", generator=g)
    f = 80.0 + 7520.0 * torch.rand(B, 4, device="cuda", generator=g)
    for i in range(4):
        wav += 0.08 * torch.sin(6.2831853 * f[:, i:i + 1] * t)
    lens = np.array([T - (b % 5) * 7 for b in range(B)])                     # ragged transcripts, longest = T
    rows = [synth.synth_transcript(900 + b, int(lens[b]), T) for b in range(B)]
    batch = dict(wav=wav, n_samples=torch.full((B,), n, dtype=torch.int32, device="cuda"),
                 ids=torch.stack([r[0] for r in rows]).cuda(), wid=torch.stack([r[1] for r in rows]).cuda(), lens=lens)
    eng = tower.engine()
    qz, idx = eng.tokenize_device(batch["wav"], batch["n_samples"], batch["ids"], batch["wid"], batch["lens"])
    torch.cuda.synchronize()
    return tower, eng, batch, qz, idx


def test_index_range_padding_and_determinism(full):
    tower, eng, b, qz, idx = full
    assert idx.shape == (B, T, 4) and idx.dtype == torch.int64 and qz.shape == (B, T, 1280)
    mask = torch.arange(T, device="cuda")[None, :] < torch.as_tensor(b["lens"], device="cuda")[:, None]
    assert bool((idx[mask] >= 0).all()) and bool((idx[mask] < 512).all())
    assert bool((idx[~mask] == -1).all())
    assert torch.isfinite(qz).all()
    # every codebook level is in use (the synthetic weights are conditioned for that, synth.py)
    for q in range(4):
        assert idx[..., q][mask].unique().numel() > 64
    qz2, idx2 = eng.tokenize_device(b["wav"], b["n_samples"], b["ids"], b["wid"], b["lens"])
    assert torch.equal(idx, idx2) and torch.equal(qz, qz2)                   # no atomics on the path: bit-reproducible


def test_permutation_equivariance(full):
    """Utterances are independent: permuting the batch permutes the result, bit for bit."""
    tower, eng, b, qz, idx = full
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    pc = perm.cuda()
    qzp, idxp = eng.tokenize_device(b["wav"][pc].contiguous(), b["n_samples"][pc].contiguous(), b["ids"][pc].contiguous(),
                                    b["wid"][pc].contiguous(), b["lens"][perm.numpy()])
    assert torch.equal(idxp, idx[pc])
    assert torch.equal(qzp, qz[pc])


def test_rvq_round_trip(full):
    """get_output_from_indices(encode(z).indices) == encode(z).quantized (RVQ:239-242 vs RVQ:470), padded rows = bias."""
    tower, eng, b, qz, idx = full
    out = eng.rvq_decode(idx, project_out=True)
    assert float((out - qz).abs().max()) <= 1e-5 * float(qz.abs().max())
    code = eng.rvq_decode(idx, project_out=False)                             # sum of the 4 selected codes
    bias = tower.vq.rvq.project_out.bias
    mask = torch.arange(T, device="cuda")[None, :] < torch.as_tensor(b["lens"], device="cuda")[:, None]
    assert torch.allclose(qz[~mask], bias.expand(int((~mask).sum()), -1), atol=1e-6)
    assert float(code[~mask].abs().max()) == 0.0


def test_logmel_gain_shift(full):
    """log-mel of g * x equals log-mel of x + 2 log10(g) / 4 on every bin (the max - 8 floor moves with the signal)."""
    tower, eng, b, qz, idx = full
    fe = WhisperFrontendB200(whisper_model="large-v3", do_pad_trim=True, permute=True).to("cuda:0")
    a, _ = fe.forward_device(b["wav"][:8].contiguous(), b["n_samples"][:8].contiguous(), True, False)
    c, _ = fe.forward_device((b["wav"][:8] * 4.0).contiguous(), b["n_samples"][:8].contiguous(), True, False)
    shift = 2.0 * math.log10(4.0) / 4.0
    assert float((c - a - shift).abs().max()) < 2e-5
