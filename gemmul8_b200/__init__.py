"""gemmul8_b200: B200-native (sm_100a, tcgen05) Ozaki-II GEMM emulator -- a drop-in for the hot path of
RIKEN-RCCS/GEMMul8 (split -> num_moduli INT8 GEMMs -> CRT).  See DESIGN.md / INTEGRATION.md."""
from . import tables  # noqa: F401  (pure python, no GPU needed)

__all__ = ["tables", "Backend", "Op", "gemm", "matmul", "work_size", "layout", "randmat", "gemm_host", "HostGemm", "NativeHostGemm"]


def __getattr__(name):
    # torch / the native library are imported lazily so that `import gemmul8_b200.tables` stays light
    if name in ("gemm_host", "HostGemm", "NativeHostGemm"):
        from . import host_pipeline

        return getattr(host_pipeline, name)
    if name in __all__:
        from . import api

        return getattr(api, name)
    raise AttributeError(name)
