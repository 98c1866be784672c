"""ctypes binding of libvilbert_b200.so (C ABI: include/vilbert_b200.h).

The library is the product; there is NO fallback. If the shared object is missing or a call
returns a non-zero status this module raises — nothing here ever routes to a CPU/PyTorch path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvilbert_b200.so")

VB_ACT_NONE, VB_ACT_GELU, VB_ACT_RELU, VB_ACT_DGELU = 0, 1, 2, 3


class VBError(RuntimeError):
    pass


class Dropout(C.Structure):
    """Mirror of ``struct vb_dropout``."""

    _fields_ = [("step", C.c_void_p), ("site", C.c_uint32), ("p", C.c_float)]


class GemmArgs(C.Structure):
    """Mirror of ``struct vb_gemm_args`` (include/vilbert_b200.h)."""

    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
        ("alpha", C.c_float),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ld_res", C.c_int64),
        ("aux", C.c_void_p), ("ld_aux", C.c_int64),
        ("act", C.c_int32),
        ("out_f32", C.c_void_p), ("ld_out_f32", C.c_int64),
        ("out_bf16", C.c_void_p), ("ld_out_bf16", C.c_int64),
        ("out_pre", C.c_void_p), ("ld_out_pre", C.c_int64),
        ("atomic_out", C.c_int32), ("out_colsum", C.c_void_p), ("dropout", Dropout), ("split_k", C.c_int32), ("block_n", C.c_int32), ("max_ctas", C.c_int32),
        ("dbg_lbo_a", C.c_uint32), ("dbg_sbo_a", C.c_uint32), ("dbg_lbo_b", C.c_uint32), ("dbg_sbo_b", C.c_uint32),
        ("dbg_timeline", C.c_void_p), ("cluster_m", C.c_int32),
        ("a_fp16", C.c_int32), ("b_fp16", C.c_int32), ("out_fp16", C.c_int32),
        ("A_lo", C.c_void_p), ("B_lo", C.c_void_p), ("out_lo", C.c_void_p), ("out_b16", C.c_void_p),
    ]


class AttnArgs(C.Structure):
    """Mirror of ``struct vb_attn_args``."""

    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("Nq", C.c_int32), ("Nk", C.c_int32), ("D", C.c_int32),
        ("Q", C.c_void_p), ("ldq", C.c_int64), ("K", C.c_void_p), ("ldk", C.c_int64), ("V", C.c_void_p), ("ldv", C.c_int64),
        ("mask", C.c_void_p), ("scale", C.c_float),
        ("O", C.c_void_p), ("ldo", C.c_int64), ("lse", C.c_void_p),
        ("dO", C.c_void_p), ("lddo", C.c_int64), ("dQ", C.c_void_p), ("lddq", C.c_int64),
        ("dK", C.c_void_p), ("lddk", C.c_int64), ("dV", C.c_void_p), ("lddv", C.c_int64),
        ("delta", C.c_void_p),
        ("dbias_q", C.c_void_p), ("dbias_k", C.c_void_p), ("dbias_v", C.c_void_p),
        ("dropout", Dropout),
        ("qkv_fp16", C.c_int32), ("Q_lo", C.c_void_p), ("K_lo", C.c_void_p), ("V_lo", C.c_void_p), ("O_lo", C.c_void_p),
        ("O_b16", C.c_void_p),
    ]


class AdamWGroup(C.Structure):
    """Mirror of ``struct vb_adamw_group``."""

    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("correct_bias", C.c_int32)]


_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
# argument types of every entry point of include/vilbert_b200.h (the trailing void* is the stream)
_SIGNATURES = {
    "vb_device_info": [C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "vb_gemm_bf16": [C.POINTER(GemmArgs), _P],
    "vb_gemm_plan": [C.POINTER(GemmArgs), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "vb_attention_fwd": [C.POINTER(AttnArgs), _P],
    "vb_attention_bwd": [C.POINTER(AttnArgs), _P],
    "vb_attention_probs": [C.POINTER(AttnArgs), _P, _P],
    "vb_layernorm_fwd": [_P, _I64, _P, _P, _F, _P, _P, _I64, _P, _P, _I32, _I32, _P, _I32, _P, _P, _P],
    "vb_layernorm_bwd": [_P, _I64, _P, _I64, _P, _P, _P, _P, _P, _I64, _P, _I64, _P, _P, _P, _I32, _I32, _P, _P, _P],
    "vb_cast_f32_to_bf16": [_P, _P, _I64, _I32, _P, _P, _P],
    "vb_cast2d_f32_to_bf16": [_P, _I64, _P, _I64, _I32, _I32, _F, _P],
    "vb_embed_text_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P],
    "vb_embed_text_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P],
    "vb_loc_proj_fwd": [_P, _P, _P, _P, _I32, _I32, _P],
    "vb_loc_proj_bwd": [_P, _P, _P, _P, _I32, _I32, _P],
    "vb_colsum": [_P, _I32, _I64, _P, _I32, _I32, _P],
    "vb_small_linear_fwd": [_P, _I64, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P],
    "vb_small_linear_bwd": [_P, _P, _I64, _P, _P, _I64, _I32, _P, _P, _I32, _I32, _I32, _P, _P],
    "vb_fuse_pooled_fwd": [_P, _P, _P, _P, _I64, _I32, _P, _I32, _P, _P, _P],
    "vb_fuse_pooled_bwd": [_P, _P, _P, _P, _P, _I64, _I32, _P, _P],
    "vb_step_counter_bump": [_P, _P],
    "vb_broadcast_rows": [_P, _P, _I64, _I32, _P],
    "vb_repeat_rows": [_P, _P, _I64, _I64, _I32, _P],
    "vb_sum_strided": [_P, _P, _I64, _I32, _I64, _I32, _I64, _I32, _P],
    "vb_relu_bwd": [_P, _P, _P, _P, _I64, _P],
    "vb_axpy_f32": [_P, _P, _I64, _F, _P],
    "vb_bce_logits_loss": [_P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _P],
    "vb_mask_to_additive": [_P, _P, _I32, _I32, _I32, _P],
    "vb_memset_zero": [_P, _I64, _P],
    "vb_ce_loss": [_P, _I64, _P, _I64, _P, _P, _I64, _P, _I64, _I32, _I32, _F, _I32, _P],
    "vb_kl_masked_loss": [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _F, _I32, _P],
    "vb_masked_mean_fwd": [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "vb_masked_mean_bwd": [_P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "vb_gate_scale_fwd": [_P, _P, _I64, _P, _I32, _I32, _I32, _I32, _P],
    "vb_gate_scale_bwd": [_P, _I64, _P, _P, _I64, _P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "vb_compact_rows": [_P, _I64, _I32, _I32, _P, _P, _P, _P],
    "vb_gather_rows16": [_P, _P, _P, _P, _P, _I32, _I32, _P],
    "vb_scatter_rows_f32": [_P, _P, _P, _I32, _I32, _P, _P, _P],
    "vb_adamw_step": [_P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _I32, _P, _P, _F, _I32, _P],
}

_lib = None


def lib():
    """Loads the shared library once; raises VBError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VBError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the ViLBERT B200 kernels)")
        _lib = C.CDLL(LIB_PATH)
        _lib.vb_last_error.restype = C.c_char_p
        _lib.vb_version.restype = C.c_int
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(_lib, name)   # AttributeError here = stale build: fail loudly
            fn.argtypes = argtypes
            fn.restype = C.c_int
    return _lib


def check(status, what=""):
    if status != 0:
        msg = lib().vb_last_error().decode("utf-8", "replace")
        raise VBError(f"{what or 'libvilbert_b200'} failed with status {status}: {msg}")


def exported_symbols():
    """Names declared in include/vilbert_b200.h (parsed), used by the ABI test."""
    import re
    hdr = os.path.join(_HERE, "..", "include", "vilbert_b200.h")
    with open(hdr) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", text)))
