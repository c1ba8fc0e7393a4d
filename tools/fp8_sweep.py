#!/usr/bin/env python3
"""BASELINE.json config 5: DGEMM 8192^3, FP8 backend, num_moduli sweep -> (err_max, err_med, TFLOPS) for ours and the reference.
Errors are measured against a float64 torch.matmul on a 512 x 512 corner with the harness' relative-error definition."""
import ctypes, json, statistics, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gemmul8_b200 as g8
sys.argv = [sys.argv[0]] + sys.argv[1:]
from bench import RefLib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
Ns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 10, 12, 14, 16, 18, 20]
dt = torch.float64
A = g8.randmat(S, S, dt, seed=12345); B = g8.randmat(S, S, dt, seed=54321)
C = torch.zeros(S * S, dtype=dt, device="cuda")
Am = A.view(S, S).t()[:512, :]; Bm = B.view(S, S).t()[:, :512]
ref_corner = Am @ Bm
ref = RefLib()
one = (ctypes.c_double * 1)(1.0); zero = (ctypes.c_double * 1)(0.0)
st = torch.cuda.current_stream()
rows = []
for N in Ns:
    for fast in (False, True):
        for impl in ("ours", "reference"):
            tot = g8.work_size(S, S, S, N, backend=1)[0]
            work = torch.empty(tot, dtype=torch.uint8, device="cuda")
            def step():
                if impl == "ours":
                    g8.gemm("N", "N", S, S, S, 1.0, A, S, B, S, 0.0, C, S, N, fast, work, backend=1)
                else:
                    ref.L.ref_gemm(1, 1, 1, 0, 0, S, S, S, ctypes.addressof(one), A.data_ptr(), S, B.data_ptr(), S, ctypes.addressof(zero), C.data_ptr(), S,
                                   N, int(fast), work.data_ptr(), None, None, 0, 0, 0, 0, ctypes.c_void_p(st.cuda_stream), None)
            for _ in range(3): step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            K = 6
            e0.record()
            for _ in range(K): step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            got = C.view(S, S).t()[:512, :512]
            rel = ((got - ref_corner).abs() / ref_corner.abs().clamp_min(1e-300))
            rows.append(dict(N=N, fast=fast, impl=impl, ms=round(ms, 3), tflops=round(2 * S ** 3 / ms * 1e-9, 1), err_max=float(rel.max()), err_med=float(rel.median())))
            print(json.dumps(rows[-1]), flush=True)
            del work
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/fp8_sweep.json").write_text(json.dumps(rows, indent=1))
