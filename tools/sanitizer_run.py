#!/usr/bin/env python3
"""Small shapes through every product code path, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):
INT8 real + complex, FP8 real + complex, fast + accurate, ragged sizes, alpha/beta general."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import helpers as H

rng = np.random.default_rng(3)
for dtype, N, be in ((np.float64, 14, 0), (np.float32, 6, 0), (np.complex128, 7, 0), (np.float64, 9, 1), (np.complex64, 5, 1)):
    for fast in (False, True):
        m, n, k = 130, 70, 300
        A = H.rand_matrix(rng, (m, k), dtype); B = H.rand_matrix(rng, (k, n), dtype); C0 = H.rand_matrix(rng, (m, n), dtype)
        C = H.run_gemm(A, B, "N", "T" if False else "N", N, fast, alpha=0.5, beta=-1.0, C0=C0, backend=be)
        wide = np.complex128 if np.dtype(dtype).kind == "c" else np.float64
        ref = 0.5 * (A.astype(wide) @ B.astype(wide)) - C0.astype(wide)
        err = np.abs(C - ref).max() / np.abs(ref).max()
        print(f"{np.dtype(dtype).name} N={N} be={be} fast={fast} relerr={err:.2e}", flush=True)
print("sanitizer workload done")
