"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, the drop-in modules
carry the reference's state_dict layout, the log-mel tables reproduce the DFT, and the product refuses to run on CPU."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import taste_oracle as O                      # checker only
from taste_spokenlm_b200 import _lib, mel, synth

torch.set_grad_enabled(False)


def test_public_header_is_plain_c(tmp_path):
    """include/taste_b200.h is the binding surface for any host language: it must compile as strict C99 on its own and
    link against the built library from a C translation unit (no C++ types, no torch, no CUDA headers)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include "taste_b200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '  taste_weights_t w; (void)w;\n'
                   '  printf("%d %d %d\\n", taste_abi_version(), TASTE_ABI_VERSION, taste_operand_dtype());\n'
                   '  return taste_abi_version() == TASTE_ABI_VERSION ? 0 : 1;\n}\n')
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                    "-fsyntax-only", str(src)], check=True)
    for flavour, code in (("bf16", 0), ("fp16", 1)):
        so = _lib.LIB_PATHS[flavour]
        if not os.path.exists(so):
            pytest.skip("library not built")
        exe = tmp_path / f"abi_{flavour}"
        subprocess.run([gcc, "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), so,
                        f"-Wl,-rpath,{os.path.dirname(so)}"], check=True)
        out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
        assert out == ["4", "4", str(code)]


def test_library_exports_every_declared_symbol(built_lib):
    declared = _lib.declared_symbols()
    assert len(declared) >= 18
    assert set(declared) == set(_lib._SIGS), "ctypes signature table out of sync with include/taste_b200.h"
    for flavour, code in (("bf16", 0), ("fp16", 1)):                 # both library flavours export the same ABI
        lib = _lib.load(flavour)
        for name in declared:
            assert hasattr(lib, name), f"{flavour} library does not export {name}"
        assert lib.taste_abi_version() == 4
        assert lib.taste_operand_dtype() == code
        assert isinstance(lib.taste_launch_count(), int)
    assert _lib.load("bf16") is built_lib or _lib.resolve_precision() == "fp16"
    with pytest.raises(_lib.TasteError):
        _lib.load("int8")


def test_argument_errors_without_a_gpu(built_lib):
    # argument validation happens before any CUDA call, so these are safe on a CPU box
    assert built_lib.taste_handle_create(None, None) == -1
    assert b"null" in built_lib.taste_last_error()
    assert built_lib.taste_gemm_bf16(None, None, None, None, 1, 128, 64, 9, None) == -1
    assert built_lib.taste_ws_bytes(None, 4, 100) == 0
    assert built_lib.taste_logmel_set_mode(2) == -1 and b"logmel_set_mode" in built_lib.taste_last_error()
    assert built_lib.taste_logmel_set_mode(0) == 0
    assert built_lib.taste_logmel_f32(None, None, None, 1, 0, None, None, None, 0, None) == -1


def test_tower_module_state_layout_matches_reference_keys():
    from taste_spokenlm_b200.tower import TasteAudioTowerB200
    for cfg in (synth.TINY, synth.SMALL):
        t = TasteAudioTowerB200.from_config(cfg)
        sd = t.state_dict()
        spec = synth.state_dict_spec(cfg)
        assert list(sd.keys()) == list(spec.keys()) or set(sd.keys()) == set(spec.keys())
        for k, shape in spec.items():
            assert tuple(sd[k].shape) == tuple(shape), k
        # identity cross-attention v_proj at construction (JES:320-322)
        v = sd["audio_joint_encoder_segmenter.audio_segmenter.decoder.layers.0.encoder_attn.v_proj.weight"]
        assert torch.equal(v, torch.eye(cfg.d_model))
        t.load_state_dict(synth.random_weights(cfg, 3), strict=True)


def test_full_size_spec_matches_golden_reference_key_list(golden_dir):
    # the 759.0 M parameter count of SURVEY section 6 (audio tower incl. RVQ and the unused QINCo MLPs)
    spec = synth.state_dict_spec(synth.FULL)
    n = sum(int(np.prod(s)) for k, s in spec.items()
            if not any(k.endswith(b) for b in ("initted", "cluster_size", "embed_avg", "_codebook.embed")))
    assert abs(n / 1e6 - 759.0) < 1.0, n / 1e6


def test_product_has_no_cpu_fallback():
    from taste_spokenlm_b200.tower import TasteAudioTowerB200
    from taste_spokenlm_b200.frontend import WhisperFrontendB200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    t = TasteAudioTowerB200.from_config(synth.TINY).eval()
    b = synth.synth_batch(0, [1.0], [4])
    with pytest.raises(_lib.TasteError):
        t(b["asr_token_ids"], b["asr_token_lengths"], torch.zeros(1, 3000, 128), torch.tensor([3000]),
          asr_word_ids=b["asr_word_ids"])
    with pytest.raises(_lib.TasteError):
        WhisperFrontendB200(whisper_model="large-v3", permute=True)(b["wav"], b["n_samples"])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "taste_spokenlm_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_token_assembly_host_matches_oracle():
    from taste_spokenlm_b200.engine import TowerEngine
    ids = torch.tensor([[5, 6, 7, 8], [9, 10, 0, 0], [11, 0, 0, 0]])
    lens = np.array([4, 2, 1])
    tok, cu = TowerEngine.assemble_tokens_host(ids.numpy(), lens)
    full = O.assemble_tokens(ids)                                      # MT:144-151, padded
    for b in range(3):
        assert tok[cu[b]:cu[b + 1]].tolist() == full[b, : lens[b] + 5].tolist()
    assert cu.tolist() == [0, 9, 16, 22]


def test_mel_tables_match_oracle_filterbank():
    fb = mel.slaney_filterbank()
    np.testing.assert_allclose(fb, O.mel_filterbank(), rtol=2e-6, atol=1e-9)
    st, cnt, wt = mel.sparse_filterbank()
    dense = np.zeros_like(fb)
    for m in range(128):
        dense[m, st[m]: st[m] + cnt[m]] = wt[m, : cnt[m]]
    assert np.array_equal(dense, fb)
    assert int(cnt.sum()) >= 394 and int(np.count_nonzero(fb)) == 394


def test_folded_dft_tables_reproduce_rfft():
    """The kernel's arithmetic (logmel.cu stage 1-2) in numpy fp64: X[k] = sum_{n=1..199} E[n] cos - i O[n] sin
    + (-1)^k y[200], with E/O the folded windowed frame; must equal rfft(frame * hann)."""
    c, s = mel.dft_tables()
    w = mel.hann_periodic().astype(np.float64)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(400)
    y = w * x
    n = np.arange(1, 200)
    E = y[n] + y[400 - n]
    Od = y[n] - y[400 - n]
    k = np.arange(201)
    re = E @ c[:199, :201].astype(np.float64) + np.where(k % 2 == 1, -1.0, 1.0) * y[200]      # y[0] = 0 (periodic Hann)
    im = Od @ s[:199, :201].astype(np.float64)
    ref = np.fft.rfft(y)
    np.testing.assert_allclose(re, ref.real, atol=2e-5)
    np.testing.assert_allclose(im, -ref.imag, atol=2e-5)
    assert w[0] == 0.0


def test_split_bf16_dft_weights_reproduce_the_oracle_logmel():
    """The tensor-core front-end's arithmetic (logmel.cu mode 0) in numpy: frames split into bf16 halves against the
    [512 x 1344] twiddle matrix {hi, lo, hi} of mel.dft_gemm_weights(), power, mel, log, floor, scale - must equal the
    oracle's log-mel within the north star's 1e-4 relative L2 (the GPU test checks the kernels themselves)."""
    W = mel.dft_gemm_weights()
    assert W.shape == (mel.DFT_N, 3 * mel.DFT_K) and W.dtype == np.uint16
    Wf = (W.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    assert not Wf[201:256].any() and not Wf[457:].any() and not Wf[:, 400:448].any()      # padding rows / columns
    n = 16000 * 3
    wav = synth.synth_waveform(5, n)[None]
    ref, _ = O.log_mel(wav, [n])
    x = np.zeros(480000, np.float32)
    x[:n] = wav[0].numpy()
    xp = np.concatenate([np.pad(x, (200, 200), mode="reflect"), np.zeros(mel.DFT_K, np.float32)])
    frames = xp[np.arange(3000)[:, None] * 160 + np.arange(mel.DFT_K)[None, :]]
    hi = mel._bf16_round(frames)
    lo = mel._bf16_round(frames - hi)
    spec = (np.concatenate([hi, hi, lo], 1).astype(np.float64) @ Wf.T).astype(np.float32)
    power = spec[:, :201] ** 2 + spec[:, 256:457] ** 2
    st, cnt, wt = mel.sparse_filterbank()
    M = np.zeros((128, 201), np.float32)
    for m in range(128):
        M[m, st[m]:st[m] + cnt[m]] = wt[m, :cnt[m]]
    lg = np.log10(np.maximum(power @ M.T, 1e-10))
    feats = (np.maximum(lg, lg.max() - 8.0) + 4.0) / 4.0
    err = np.linalg.norm(feats - ref[0].numpy()) / np.linalg.norm(ref[0].numpy())
    assert err < 1e-4, err


def test_word_runs_padded_row_quirk():
    # a final word with id 0 merges with the zero padding and is not pooled
    assert O.word_runs(torch.tensor([0, 0, 0, 0]), 2) == []
    assert O.word_runs(torch.tensor([0, 0]), 2) == [(0, 2)]
    assert O.word_runs(torch.tensor([0, 1, 1, 2, 2, 2, 0, 0]), 7) == [(1, 3), (3, 6)]


def test_tower_deepcopy_and_pickle_rebind_children():
    """The RVQ / segmenter reach the kernels through a weak reference to their owning tower: a copied or unpickled tower
    must own its children's references (not share the original's engine) and drops the packed weights."""
    import copy
    import io
    from taste_spokenlm_b200.tower import TasteAudioTowerB200
    t = TasteAudioTowerB200.from_config(synth.TINY)
    t2 = copy.deepcopy(t)
    assert t2.vq.rvq._tower_ref() is t2 and t2.audio_joint_encoder_segmenter._tower_ref() is t2
    assert t.vq.rvq._tower_ref() is t
    buf = io.BytesIO()
    torch.save(t, buf)
    buf.seek(0)
    t3 = torch.load(buf, weights_only=False)
    assert t3.vq.rvq._tower_ref() is t3 and t3._engine is None
    assert set(t3.state_dict()) == set(t.state_dict())
