"""Loader of the native C-ABI library (include/gemmul8_c.h).  There is NO fallback: if the sm_100a
extension is missing the import of any compute entry point fails loudly."""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libg8core.so"

c_size_t, c_int, c_uint, c_void_p, c_double = ctypes.c_size_t, ctypes.c_int, ctypes.c_uint, ctypes.c_void_p, ctypes.c_double


class GemmDesc(ctypes.Structure):
    """struct g8_gemm_desc (include/gemmul8_c.h) == parameter list of gemmul8::gemm/gemmLt (reference include/gemmul8.hpp:41-94)."""
    _fields_ = [
        ("dtype", c_int), ("backend", c_int), ("op_A", c_int), ("op_B", c_int),
        ("m", c_size_t), ("n", c_size_t), ("k", c_size_t),
        ("alpha", c_void_p), ("A", c_void_p), ("lda", c_size_t), ("B", c_void_p), ("ldb", c_size_t),
        ("beta", c_void_p), ("C", c_void_p), ("ldc", c_size_t),
        ("num_moduli", c_uint), ("fastmode", c_int),
        ("work", c_void_p), ("workA", c_void_p), ("workB", c_void_p),
        ("enable_skip_scalA", c_int), ("enable_skip_scalB", c_int), ("skip_scalA", c_int), ("skip_scalB", c_int),
        ("stream", c_void_p),
    ]


# every symbol include/gemmul8_c.h declares: (restype, argtypes)
SYMBOLS = {
    "g8_work_size": (c_size_t, [c_int, c_int, c_size_t, c_size_t, c_size_t, c_uint, c_int, c_int,
                                ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t)]),
    "g8_gemm": (c_int, [ctypes.POINTER(GemmDesc), ctypes.POINTER(c_double)]),
    "g8_stage_split": (c_int, [c_int, c_int, c_int, c_size_t, c_size_t, c_void_p, c_size_t, c_uint, c_int, c_void_p,
                               c_void_p, c_size_t, c_size_t, c_void_p]),
    "g8_stage_finalize_shift": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, c_void_p]),
    "g8_stage_gemm": (c_int, [c_int, c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_int, c_int,
                              ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_void_p, c_size_t, c_size_t, c_void_p, c_void_p,
                              c_void_p]),
    "g8_stage_crt": (c_int, [c_int, c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_uint, c_void_p, c_size_t, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "g8_stage_crt_parts": (c_int, [c_int, c_void_p, c_int, c_size_t, c_size_t, c_size_t, c_size_t, c_size_t, c_uint, c_void_p, c_size_t, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "g8_stage_requant_i32": (c_int, [c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_int, c_int, c_void_p, c_size_t, c_size_t, c_void_p]),
    "g8_stage_residue_sum": (c_int, [c_void_p, c_int, c_size_t, c_size_t, c_size_t, c_size_t, c_size_t, c_int, c_int, c_void_p, c_size_t,
                                     c_size_t, c_void_p]),
    "g8_stage_maxabs_i32": (c_int, [c_void_p, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "g8_stage_maxabs_i32_parts": (c_int, [c_void_p, c_int, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "g8_stage_gemm_bound_chain": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_int, c_void_p, c_void_p, c_void_p]),
    "g8_stage_gemm_scatter": (c_int, [c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_int, c_int,
                                      ctypes.POINTER(c_void_p), c_int, c_int, c_size_t, c_size_t, c_void_p]),
    "g8_peer_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p), c_void_p]),
    "g8_peer_open": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "g8_peer_close": (c_int, [c_void_p]),
    "g8_peer_free": (c_int, [c_void_p]),
    "g8_stage_stats": (c_int, [c_int, c_int, c_int, c_size_t, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "g8_stage_shift_from_stats": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, c_int, c_void_p, c_void_p]),
    "g8_host_plan_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_size_t, c_size_t, c_size_t, c_uint, c_int, c_size_t]),
    "g8_gemm_host": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p]),
    "g8_host_plan_destroy": (c_int, [c_void_p]),
    "g8_mg_comm_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_size_t, c_void_p]),
    "g8_mg_comm_connect": (c_int, [c_void_p, c_void_p]),
    "g8_mg_comm_barrier": (c_int, [c_void_p, c_void_p]),
    "g8_mg_comm_status": (c_int, [c_void_p]),
    "g8_mg_comm_destroy": (c_int, [c_void_p]),
    "g8_mg_plan_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_int, c_int, c_int, c_size_t, c_size_t, c_size_t, c_uint, c_int]),
    "g8_mg_plan_create_backend": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_int, c_int, c_int, c_int, c_size_t, c_size_t, c_size_t, c_uint, c_int]),
    "g8_gemm_mg": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p]),
    "g8_mg_plan_destroy": (c_int, [c_void_p]),
    "g8_randmat": (c_int, [c_int, c_void_p, c_size_t, c_size_t, c_double, ctypes.c_ulonglong, c_void_p]),
    "g8_version": (ctypes.c_char_p, []),
    "g8_device_supported": (c_int, [c_int]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """dlopen gemmul8_b200/lib/libg8core.so and bind every declared entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("GEMMUL8_B200_LIB", LIB_PATH))
    if not path.exists():
        raise NativeLibraryMissing(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C gemmul8_b200/csrc`. gemmul8_b200 has no CPU / PyTorch fallback path.")
    lib = ctypes.CDLL(str(path), mode=ctypes.RTLD_LOCAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib
