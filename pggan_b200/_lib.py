"""ctypes binding of libpggan_b200.so — the thin C-ABI boundary (include/pggan_b200.h).

There is NO CPU fallback: if the shared library is missing or a call fails, the error is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

PG_BF16, PG_F32, PG_F32_RAW, PG_F16 = 0, 1, 2, 3
PG_BACKEND_AUTO, PG_BACKEND_SIMT = 0, 1
PG_CONV_KIND_SIMT, PG_CONV_KIND_PLAIN, PG_CONV_KIND_SWAPPED, PG_CONV_KIND_ROWS = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpggan_b200.so")


class ConvParams(Structure):
    _fields_ = [
        ("x", c_void_p), ("wp", c_void_p), ("bias", c_void_p), ("mask_src", c_void_p), ("y", c_void_p),
        ("B", c_int32), ("H", c_int32), ("W", c_int32), ("C", c_int32), ("N", c_int32), ("taps", c_int32),
        ("dtype", c_int32), ("out_dtype", c_int32),
        ("scale", c_float), ("slope", c_float), ("mask_slope", c_float),
        ("ws_x", c_void_p), ("ws_w", c_void_p),
        ("pn_rnorm", c_void_p), ("pn_eps", c_float),
        ("up2", c_int32),
        ("ws_k", c_void_p), ("ws_k_bytes", ctypes.c_int64),
    ]


class WgradParams(Structure):
    _fields_ = [
        ("g", c_void_p), ("x", c_void_p), ("acc", c_void_p),
        ("B", c_int32), ("H", c_int32), ("W", c_int32), ("C", c_int32), ("N", c_int32), ("taps", c_int32),
        ("dtype", c_int32),
        ("ws_g", c_void_p), ("ws_x", c_void_p),
        ("direct", c_void_p), ("direct_scale", c_float), ("reserved_", c_int32),
    ]


class PackMap(Structure):
    _fields_ = [("n_total", c_int32), ("n_valid", c_int32), ("n2", c_int32), ("taps", c_int32),
                ("c_total", c_int32), ("c_valid", c_int32), ("c2", c_int32),
                ("base", c_int64), ("s_n1", c_int64), ("s_n2", c_int64), ("s_t", c_int64), ("s_c1", c_int64),
                ("s_c2", c_int64)]


class Coef(Structure):
    """pg_coef: value = c0 + c1 * (*alpha); alpha = device pointer to one fp32 (or NULL -> c0)"""
    _fields_ = [("c0", c_float), ("c1", c_float), ("alpha", c_void_p)]


class AdamEntry(Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("numel", c_int64), ("shadow", c_void_p), ("shadow_dtype", c_int32), ("pad_", c_int32)]


# name -> (restype, argtypes); exactly the symbols include/pggan_b200.h declares
SIGNATURES = {
    "pggan_version": (c_int, []),
    "pggan_last_error": (c_char_p, []),
    "pggan_set_conv_backend": (c_int, [c_int]),
    "pggan_launch_counts": (c_int, [POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    "pggan_sizeof_conv_params": (c_int, []),
    "pggan_sizeof_wgrad_params": (c_int, []),
    "pggan_pack_weight": (c_int, [c_void_p, c_void_p, c_int, POINTER(PackMap), c_void_p]),
    "pggan_unpack_wgrad": (c_int, [c_void_p, c_void_p, POINTER(PackMap), c_int, c_float, c_int, c_void_p]),
    "pggan_convert": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "pggan_split3": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "pggan_conv_fprop": (c_int, [POINTER(ConvParams), c_void_p]),
    "pggan_conv_kernel_kind": (c_int, [POINTER(ConvParams)]),
    "pggan_conv_pn_fusable": (c_int, [POINTER(ConvParams)]),
    "pggan_conv_up2_fusable": (c_int, [POINTER(ConvParams)]),
    "pggan_conv_wgrad": (c_int, [POINTER(WgradParams), c_void_p]),
    "pggan_conv_wgrad_slices": (c_int, [POINTER(WgradParams)]),
    "pggan_conv_wgrad_direct_ok": (c_int, [POINTER(WgradParams)]),
    "pggan_lrelu_mask_mul": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_int, c_void_p]),
    "pggan_upsample2x": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, Coef, c_int, c_void_p]),
    "pggan_upsample2x_mask": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, Coef, c_float, c_int,
                                      c_void_p]),
    "pggan_upsample2x_mask_colsum": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, Coef,
                                             c_float, c_float, c_int, c_void_p]),
    "pggan_pool2x2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, Coef, Coef, c_int,
                              c_void_p]),
    "pggan_axpby": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, Coef, Coef, c_int, c_void_p]),
    "pggan_pixelnorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_int, c_void_p]),
    "pggan_pixelnorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_float,
                                    c_int, c_int, c_void_p]),
    "pggan_mask_colsum": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_int,
                                  c_void_p]),
    "pggan_colsum": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_int, c_void_p]),
    "pggan_rgb_expand": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_float, c_float, c_int, c_void_p]),
    "pggan_rgb_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_float, c_float, c_int, c_void_p]),
    "pggan_rgb_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_float, c_int, c_void_p]),
    "pggan_image_fade": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, Coef, c_void_p]),
    "pggan_image_pool_sum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, Coef, c_void_p]),
    "pggan_resize_aa": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "pggan_jitter_normalize_fade": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, Coef,
                                            c_void_p]),
    "pggan_mbstd_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p]),
    "pggan_mbstd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_void_p]),
    "pggan_mbstd_bwd2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int,
                                 c_void_p]),
    "pggan_rows_sumsq": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "pggan_rows_scale": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "pggan_gp_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p]),
    "pggan_adam_chunk_size": (c_int, []),
    "pggan_adam_step": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                c_float, c_float, c_void_p]),
}

_lib = None


class PgganLibraryError(RuntimeError):
    pass


def load(path: str | None = None) -> ctypes.CDLL:
    """Load the shared library (building nothing: run `python -m pggan_b200.build` or
    `__graft_entry__.build()` first).  Raises if it is absent — there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("PGGAN_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise PgganLibraryError(
            f"{path} not found: the CUDA extension is not built (python -m pggan_b200.build). "
            "pggan_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    # the parameter structs of this binding must be the ones the library was compiled with
    for struct, fn in ((ConvParams, lib.pggan_sizeof_conv_params), (WgradParams, lib.pggan_sizeof_wgrad_params)):
        if ctypes.sizeof(struct) != fn():
            raise PgganLibraryError(f"{path}: sizeof({struct.__name__}) is {fn()} in the library but "
                                    f"{ctypes.sizeof(struct)} in this binding - rebuild (python -m pggan_b200.build)")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pggan_last_error()
        raise PgganLibraryError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def launch_counts():
    a, b, c = c_int64(0), c_int64(0), c_int64(0)
    load().pggan_launch_counts(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return {"tcgen05": a.value, "simt": b.value, "other": c.value}
