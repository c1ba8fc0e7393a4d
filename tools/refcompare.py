#!/usr/bin/env python3
"""TEST / MEASUREMENT INFRASTRUCTURE: one emulated GEMM, ours next to the unmodified reference library (oracle/_ref) on the same GPU,
same inputs (the reference harness' generator), protocol of testing/test_flops.hpp:169-216 (warm-up + timed calls bracketed by CUDA
events, median; TFLOPS = 2mnk (x4 complex) / t).  Used by tools/config_table.py, tools/odd_shapes.py and tests/test_gpu_configs.py.
Importing this module has no side effects."""
import ctypes
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        from bench import RefLib
        _ref = RefLib()
    return _ref


def run_case(tag, t, m, n, k, N, be, fast, phi=-1.0, warm=5, reps=10, corner=256, opA="N", opB="N", keep_outputs=False):
    """Returns a row dict: per-implementation ms / TFLOPS / corner errors, `bit_identical` (complete C) and the speed-up."""
    import gemmul8_b200 as g8
    from gemmul8_b200 import api
    ref = ref_lib()
    st = torch.cuda.current_stream()
    dt = DT[t]
    cplx = dt.is_complex
    ra, ca = (m, k) if opA == "N" else (k, m)
    rb, cb = (k, n) if opB == "N" else (n, k)
    A = g8.randmat(ra, ca, dt, phi=phi, seed=12345)
    B = g8.randmat(rb, cb, dt, phi=phi, seed=54321)
    wide = torch.complex128 if cplx else torch.float64
    cm, cn = min(corner, m), min(corner, n)

    def opview(X, r, c, op):
        M = X.view(c, r).t()  # column-major r x c
        return M if op == "N" else (M.t() if op == "T" else M.conj().t())
    Am = opview(A, ra, ca, opA)[:cm, :].to(wide)
    Bm = opview(B, rb, cb, opB)[:, :cn].to(wide)
    want = Am @ Bm
    keep = []
    one, zero = api._scalar_ptr(1.0, dt, keep), api._scalar_ptr(0.0, dt, keep)
    out = {}
    for impl in ("ours", "reference"):
        C = torch.zeros(m * n, dtype=dt, device="cuda")
        if impl == "ours":
            tot = g8.work_size(m, n, k, N, is_complex=cplx, backend=be)[0]
        else:
            tot = ref.L.ref_work_size(int(cplx), be, m, n, k, N, 0, 0, None, None)
        work = torch.empty(tot, dtype=torch.uint8, device="cuda")

        def step():
            if impl == "ours":
                g8.gemm(opA, opB, m, n, k, 1.0, A, ra, B, rb, 0.0, C, m, N, fast, work, backend=be)
            else:
                code = ref.L.ref_gemm(api._DTYPES[dt], be, 1, api._op(opA), api._op(opB), m, n, k, one, A.data_ptr(), ra, B.data_ptr(), rb, zero,
                                      C.data_ptr(), m, N, int(fast), work.data_ptr(), None, None, 0, 0, 0, 0, ctypes.c_void_p(st.cuda_stream), None)
                assert code == 0, code
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        got = C.view(n, m).t()[:cm, :cn].to(wide)
        rel = (got - want).abs() / want.abs().clamp_min(1e-300)
        out[impl] = dict(ms=round(ms, 4), tflops=round(2.0 * m * n * k * (4 if cplx else 1) / ms * 1e-9, 2), err_max=float(rel.max()),
                         err_med=float(rel.median()), err_abs_over_max=float((got - want).abs().max() / want.abs().max().clamp_min(1e-300)))
        out[impl + "_C"] = C
        del work
    same = bool(torch.equal(out["ours_C"].view(torch.uint8), out["reference_C"].view(torch.uint8)))
    d = (out["ours_C"] - out["reference_C"]).abs().max()
    row = dict(tag=tag, type=t, m=m, n=n, k=k, N=N, backend="INT8" if be == 0 else "FP8", fast=bool(fast), phi=phi, opA=opA, opB=opB,
               ours=out["ours"], reference=out["reference"], bit_identical=same,
               max_abs_diff_over_max=float(d / out["reference_C"].abs().max().clamp_min(1e-300)),
               speedup=round(out["reference"]["ms"] / out["ours"]["ms"], 3))
    if keep_outputs:
        row["_C"] = (out["ours_C"], out["reference_C"])
    return row
