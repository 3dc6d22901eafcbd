"""ctypes binding of libidiff_sm100.so (the C ABI declared in include/idiff.h).

There is NO fallback: if the shared library is missing (``python __graft_entry__.py`` /
``make -C instancediff_b200/csrc`` builds it in-tree) every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IDIFF_LIB_PATH") or os.path.join(_HERE, "libidiff_sm100.so")   # override: A/B builds

c_f32p = C.c_void_p      # device pointers travel as integers (tensor.data_ptr())
c_ptr = C.c_void_p


GN_SLOTS = 16            # IDIFF_GN_SLOTS


class GnFuse(C.Structure):
    """Mirror of ``idiff_gn_fuse`` (include/idiff.h): GroupNorm finalize folded into the producing launch."""

    _fields_ = [
        ("sums", c_ptr), ("arrivals", c_ptr), ("gamma", c_ptr), ("beta", c_ptr), ("t_scale", c_ptr), ("t_shift", c_ptr),
        ("scale_out", c_ptr), ("shift_out", c_ptr),
        ("t_ld", C.c_int32), ("count_per_group", C.c_int32), ("eps", C.c_float), ("reserved", C.c_int32),
    ]


class GemmParams(C.Structure):
    """Mirror of ``idiff_gemm_params`` (include/idiff.h)."""

    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32),
        ("cin0", C.c_int32), ("cin1", C.c_int32), ("up0", C.c_int32),
        ("N", C.c_int32), ("NT", C.c_int32), ("a_silu", C.c_int32), ("epi", C.c_int32),
        ("gn_groups", C.c_int32), ("out_ld", C.c_int32), ("dbg_swap_lbo_sbo", C.c_int32),
        ("src0_ld", C.c_int32), ("src1_ld", C.c_int32), ("reserved0", C.c_int32),
        ("src0", c_ptr), ("src1", c_ptr), ("a_scale", c_ptr), ("a_shift", c_ptr),
        ("w", c_ptr), ("w_image_stride", C.c_int64),
        ("bias", c_ptr), ("bias_img", c_ptr), ("row_stats", c_ptr), ("wsum", c_ptr),
        ("res0", c_ptr), ("res1", c_ptr), ("res0_scale", c_ptr), ("res0_shift", c_ptr),
        ("ln_g", c_ptr), ("out", c_ptr), ("gn_partial", c_ptr), ("out_row_stats", c_ptr),
        ("qscale", C.c_float), ("ln_eps", C.c_float),
        ("gn_fuse", C.POINTER(GnFuse)),
    ]


class UNetCfg(C.Structure):
    """Mirror of ``idiff_unet_cfg`` (include/idiff.h)."""

    _fields_ = [("in_nc", C.c_int32), ("out_nc", C.c_int32), ("nf", C.c_int32), ("n_levels", C.c_int32),
                ("ch_mult", C.c_int32 * 8), ("context_dim", C.c_int32), ("down_kernel", C.c_int32)]


EPI_PLAIN, EPI_QSOFTMAX, EPI_GEGLU, EPI_LN_OUT = 0, 1, 2, 3

# name -> (restype, argtypes); every symbol of include/idiff.h
SIGNATURES = {
    "idiff_abi_version": (C.c_int, []),
    "idiff_last_error": (C.c_char_p, []),
    "idiff_watchdog_status": (C.c_int, [C.c_int]),
    "idiff_sde_step": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_uint64,
                                 C.c_uint64, C.c_size_t, c_ptr]),
    "idiff_sde_step_rng": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, c_ptr, C.c_int, C.c_size_t, c_ptr]),
    "idiff_sde_pack_table": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int, C.c_float, C.c_double, c_ptr]),
    "idiff_noise_state": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_float, C.c_int, C.c_uint64, C.c_uint64,
                                    C.c_size_t, c_ptr]),
    "idiff_random_states": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_uint64, C.c_uint64,
                                      C.c_uint32, C.c_size_t, C.c_size_t, c_ptr]),
    "idiff_philox_normal": (C.c_int, [c_ptr, C.c_uint64, C.c_uint64, C.c_uint32, C.c_size_t, c_ptr]),
    "idiff_step_select": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_float, c_ptr]),
    "idiff_step_select_ss": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_float, c_ptr, c_ptr, C.c_int, c_ptr]),
    "idiff_set_debug_flags": (C.c_int, [C.c_int]),
    "idiff_debug_read_prof": (C.c_int, [c_ptr]),
    "idiff_conv_gemm": (C.c_int, [C.POINTER(GemmParams), c_ptr]),
    "idiff_sizeof_gemm_params": (C.c_int, []),
    "idiff_conv_gemm_smem_bytes": (C.c_int, [C.POINTER(GemmParams)]),
    "idiff_conv_gemm_gn_rows": (C.c_int, [C.c_int, C.c_int]),
    "idiff_conv3_rowpair": (C.c_int, [C.POINTER(GemmParams), c_ptr]),
    "idiff_conv3_rowpair_supported": (C.c_int, [C.POINTER(GemmParams)]),
    "idiff_conv3_rowpair_gn_rows": (C.c_int, [C.c_int, C.c_int]),
    "idiff_conv_ref": (C.c_int, [C.POINTER(GemmParams), c_ptr, c_ptr, c_ptr]),
    "idiff_stem_conv7": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr]),
    "idiff_stem_conv7_tc": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr]),
    "idiff_stem_packed_bytes": (C.c_int, []),
    "idiff_head_conv3": (C.c_int, [c_ptr, c_ptr, C.c_float, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr]),
    "idiff_time_embed": (C.c_int, [c_ptr, C.c_float, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                   C.c_int, C.c_int, C.c_int, c_ptr]),
    "idiff_gn_finalize": (C.c_int, [c_ptr, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, c_ptr, c_ptr,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_ptr]),
    "idiff_gn_stats": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr]),
    "idiff_gn_stats_ntile": (C.c_int, [C.c_int]),
    "idiff_block_tail": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_float, C.c_int, C.c_int,
                                   C.c_int, c_ptr]),
    "idiff_add_rows": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_float, C.c_size_t, C.c_int, c_ptr]),
    "idiff_chan_ln": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_float, C.c_size_t, C.c_int, c_ptr]),
    "idiff_chan_ln_gn": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(GnFuse),
                                   c_ptr]),
    "idiff_sizeof_gn_fuse": (C.c_int, []),
    "idiff_linattn_context": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr]),
    "idiff_linattn_scratch_floats": (C.c_size_t, [C.c_int, C.c_int]),
    "idiff_linattn_fused": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                      C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, c_ptr]),
    "idiff_linattn_fused_scratch_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "idiff_self_attention": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_float, c_ptr]),
    "idiff_cross_vec": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr]),
    "idiff_unet_create": (C.c_int, [C.POINTER(UNetCfg), C.POINTER(c_ptr)]),
    "idiff_unet_destroy": (None, [c_ptr]),
    "idiff_unet_load_weight": (C.c_int, [c_ptr, C.c_char_p, c_ptr, C.c_int, C.POINTER(C.c_int64)]),
    "idiff_unet_finalize": (C.c_int, [c_ptr]),
    "idiff_unet_set_context": (C.c_int, [c_ptr, c_ptr, C.c_int, c_ptr]),
    "idiff_unet_forward": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_float, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr]),
    "idiff_unet_reverse_sde": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_float, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                         C.c_int, c_ptr]),
    "idiff_unet_num_launches": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int]),
    "idiff_f32_to_bf16": (C.c_int, [c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "idiff_bf16_to_f32": (C.c_int, [c_ptr, c_ptr, C.c_size_t, c_ptr]),
}

_lib = None


class IdiffError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IdiffError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C instancediff_b200/csrc`). "
                "There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if handle.idiff_sizeof_gemm_params() != C.sizeof(GemmParams):
            raise IdiffError("idiff_gemm_params layout mismatch between include/idiff.h and _lib.GemmParams")
        if handle.idiff_sizeof_gn_fuse() != C.sizeof(GnFuse):
            raise IdiffError("idiff_gn_fuse layout mismatch between include/idiff.h and _lib.GnFuse")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().idiff_last_error().decode("utf-8", "replace")
        raise IdiffError(f"{what or 'idiff call'} failed ({rc}): {msg}")


def watchdog(clear: bool = True) -> None:
    """Synchronise and raise if any device-side pipeline wait timed out."""
    v = lib().idiff_watchdog_status(1 if clear else 0)
    if v != 0:
        raise IdiffError(f"device watchdog tripped at site {v}: {lib().idiff_last_error().decode()}")
