"""ctypes binding of libtaste_b200.so (include/taste_b200.h).  There is no fallback: if the library is missing or a
declared symbol is absent, load() raises."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtaste_b200.so")
# Library flavours (csrc/common.cuh, TASTE_F16): the 16-bit type of every tensor-core operand.  bf16 is BASELINE config 2's
# dtype and the default; fp16 is the reference's own GPU dtype (autocast, JES:133) with 3 more mantissa bits at the same
# tensor-core rate.  Select with `precision=` on the modules / engines or TASTE_PRECISION in the environment.
LIB_PATHS = {"bf16": LIB_PATH, "fp16": os.path.join(_HERE, "libtaste_b200_f16.so")}
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "taste_b200.h")

DFT_LD = 224
MEL_MAXW = 16
N_FRAMES = 3000
N_MELS = 128
ENC_FRAMES = 1500
N_SAMPLES = 480000

EPI_BF16, EPI_GELU_BF16, EPI_RESID_F32, EPI_F32 = 0, 1, 2, 3

p = C.c_void_p


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "d_model", "heads", "ffn", "enc_layers", "dec_layers", "vocab", "max_target_pos", "codebook_dim",
        "codebook_size", "num_quantizers", "target_layer", "reserved")]


class EncLayer(C.Structure):
    _fields_ = [(n, p) for n in ("ln1_w", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2",
                                 "wqkv_ln", "bqkv_ln", "cqkv_ln", "w1_ln", "b1_ln", "c1_ln")]


class GemmEx(C.Structure):
    _fields_ = [("a", p), ("w", p), ("bias", p), ("out", p), ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
                ("epilogue", C.c_int32), ("ln_stats", p), ("ln_nseg", C.c_int32), ("reserved", C.c_int32),
                ("ln_colsum", p), ("stats_out", p), ("out_bf16", p)]


class DecLayer(C.Structure):
    _fields_ = [(n, p) for n in (
        "ln1_w", "ln1_b", "wqkv", "bqkv", "wo", "bo", "lnx_w", "lnx_b", "wq_x", "bq_x", "wk_x", "wv_x", "bv_x",
        "wo_x", "bo_x", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2")]


class Weights(C.Structure):
    _fields_ = [("dims", Dims)] + [(n, p) for n in (
        "dft_cos", "dft_sin", "hann", "mel_start", "mel_count", "mel_weight", "dft_w_bf16",
        "conv1_w", "conv1_b", "conv2_w", "conv2_b", "enc_pos", "enc", "enc_ln_w", "enc_ln_b",
        "tok_emb", "dec_pos", "dec", "dec_ln_w", "dec_ln_b",
        "rvq_win_t", "rvq_bin", "rvq_code_split", "rvq_code", "rvq_code_sq", "rvq_wout_t", "rvq_bout")]


class ProfEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_longlong), ("total_ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


_SIGS = {
    "taste_launch_count": (C.c_ulonglong, []),
    "taste_prof_enable": (C.c_int, [C.c_int]),
    "taste_prof_reset": (C.c_int, []),
    "taste_prof_collect": (C.c_int, [C.POINTER(ProfEntry), C.c_int, C.POINTER(C.c_int)]),
    "taste_abi_version": (C.c_int, []),
    "taste_operand_dtype": (C.c_int, []),
    "taste_last_error": (C.c_char_p, []),
    "taste_handle_create": (C.c_int, [C.POINTER(Weights), C.POINTER(p)]),
    "taste_handle_destroy": (C.c_int, [p]),
    "taste_ws_bytes": (C.c_size_t, [p, C.c_int, C.c_int]),
    "taste_logmel_f32": (C.c_int, [p, p, p, C.c_int, C.c_int64, p, p, p, C.c_size_t, p]),
    "taste_encoder_fwd": (C.c_int, [p, p, p, C.c_int, p, p, p, C.c_size_t, p]),
    "taste_aggregator_fwd": (C.c_int, [p, p, p, p, p, C.c_int, C.c_int, C.c_int, p, p, C.c_size_t, p]),
    "taste_assemble_tokens": (C.c_int, [p, p, p, C.c_int, C.c_int, p, p]),
    "taste_word_pool_f32": (C.c_int, [p, p, p, p, C.c_int, C.c_int, C.c_int, p, p]),
    "taste_rvq_ws_bytes": (C.c_size_t, [C.c_int]),
    "taste_rvq_encode_f32": (C.c_int, [p, p, p, C.c_int, C.c_int, C.c_int, p, p, p, C.c_size_t, p]),
    "taste_rvq_decode_f32": (C.c_int, [p, p, C.c_int, C.c_int, p, p]),
    "taste_map_to_llm_tokens": (C.c_int, [p, p, p, p, p, C.c_int, C.c_int, C.c_int, C.c_int, p, p]),
    "taste_resample_mean_f32": (C.c_int, [p, p, p, p, C.c_int, C.c_int, C.c_int, C.c_int, p, p, C.c_int, C.c_int, C.c_int,
                                          C.c_int64, C.c_int64, p, C.c_int64, p, p]),
    "taste_gemm_bf16": (C.c_int, [p, p, p, p, C.c_int, C.c_int, C.c_int, C.c_int, p]),
    "taste_gemm_ex": (C.c_int, [C.POINTER(GemmEx), p]),
    "taste_encoder_set_mode": (C.c_int, [C.c_int]),
    "taste_logmel_set_mode": (C.c_int, [C.c_int]),
    "taste_attention_set_mode": (C.c_int, [C.c_int]),
    "taste_gemm_set_mode": (C.c_int, [C.c_int]),
    "taste_layernorm_f32": (C.c_int, [p, p, p, p, C.c_int, C.c_int, C.c_int, p]),
    "taste_attention_ragged_bf16": (C.c_int, [p, p, p, p, C.c_int, C.c_int, C.c_int, C.c_int, p, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int, p]),
    "taste_attention_bf16": (C.c_int, [p, p, p, p, C.c_int, C.c_int, C.c_int, C.c_int, p, p, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, p]),
}

_libs = {}


def resolve_precision(precision=None) -> str:
    p = (precision or os.environ.get("TASTE_PRECISION") or "bf16").lower()
    p = {"bfloat16": "bf16", "float16": "fp16", "f16": "fp16", "half": "fp16"}.get(p, p)
    if p not in LIB_PATHS:
        raise TasteError(f"unknown precision {precision!r}: expected 'bf16' or 'fp16'")
    return p


def torch_dtype(precision=None):
    import torch
    return torch.float16 if resolve_precision(precision) == "fp16" else torch.bfloat16


class TasteError(RuntimeError):
    pass


def declared_symbols(header_path: str = HEADER_PATH):
    """Every function the public header declares (used by the symbol-coverage test)."""
    txt = open(header_path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(taste_[a-z0-9_]+)\s*\(", txt)))


def load(precision=None):
    """dlopen the library of the given flavour (default: TASTE_PRECISION or bf16) and bind every declared symbol."""
    precision = resolve_precision(precision)
    if precision in _libs:
        return _libs[precision]
    path = LIB_PATHS[precision]
    if not os.path.exists(path):
        raise TasteError(f"{path} not found: run `python __graft_entry__.py` (build()) first; there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise TasteError(f"{os.path.basename(path)} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.taste_operand_dtype() != (1 if precision == "fp16" else 0):
        raise TasteError(f"{os.path.basename(path)} was not built for {precision} operands")
    lib.precision = precision
    _libs[precision] = lib
    return lib


def check(rc: int, what: str = "", lib=None):
    if rc != 0:
        msg = (lib or load()).taste_last_error().decode(errors="replace")
        raise TasteError(f"{what} failed (code {rc}): {msg}")


def prof_collect(lib=None):
    """[{name, launches, total_ms, flops, bytes}] for every kernel class launched since the last taste_prof_reset()."""
    lib = lib or load()
    arr = (ProfEntry * 32)()
    n = C.c_int(0)
    check(lib.taste_prof_collect(arr, 32, C.byref(n)), "taste_prof_collect", lib)
    return [dict(name=arr[i].name.decode(), launches=int(arr[i].launches), total_ms=float(arr[i].total_ms),
                 flops=float(arr[i].flops), bytes=float(arr[i].bytes)) for i in range(n.value)]


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
