#!/usr/bin/env python3
"""All single-GPU configurations of BASELINE.json, ours next to the unmodified reference library on the same B200:
  [S6]  SGEMM 1024^3 INT8 N=6        [D14] DGEMM 8192^3 INT8 N=14       [Z18] ZGEMM 4096^3 INT8 N=18
  [F8]  DGEMM 8192^3 FP8 N=8..20     + the reference's accuracy shape (testing/test_accuracy.hpp: m=n=128, k sweep, phi sweep)
Protocol of testing/test_flops.hpp:169-216 (warm-up + timed calls bracketed by CUDA events, median; TFLOPS = 2mnk (x4 complex) / t).
Errors: max / median relative error of a corner block against a wide reference (float64 for S, float64 matmul for D/Z -- a
double-double reference is not needed to see the emulated precision of N <= 14; larger N report the float64 noise floor).
Writes gpurun_out/config_table.json and a markdown table on stdout."""
import ctypes, json, statistics, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gemmul8_b200 as g8
from gemmul8_b200 import api
from bench import RefLib

quick = "--quick" in sys.argv
ref = RefLib()
st = torch.cuda.current_stream()
DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
rows = []

def run_case(tag, t, m, n, k, N, be, fast, phi=-1.0, warm=5, reps=10, corner=256):
    dt = DT[t]
    cplx = dt.is_complex
    A = g8.randmat(m, k, dt, phi=phi, seed=12345); B = g8.randmat(k, n, dt, phi=phi, seed=54321)
    wide = torch.complex128 if cplx else torch.float64
    cm, cn = min(corner, m), min(corner, n)
    Am = A.view(k, m).t()[:cm, :].to(wide); Bm = B.view(n, k).t()[:, :cn].to(wide)
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
                g8.gemm("N", "N", m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, fast, work, backend=be)
            else:
                code = ref.L.ref_gemm(api._DTYPES[dt], be, 1, 0, 0, m, n, k, one, A.data_ptr(), m, B.data_ptr(), k, zero, C.data_ptr(), m, N, int(fast),
                                      work.data_ptr(), None, None, 0, 0, 0, 0, ctypes.c_void_p(st.cuda_stream), None)
                assert code == 0, code
        for _ in range(warm): step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        got = C.view(n, m).t()[:cm, :cn].to(wide)
        rel = (got - want).abs() / want.abs().clamp_min(1e-300)
        out[impl] = dict(ms=round(ms, 4), tflops=round(2.0 * m * n * k * (4 if cplx else 1) / ms * 1e-9, 2), err_max=float(rel.max()), err_med=float(rel.median()))
        out[impl + "_C"] = C
    same = bool(torch.equal(out["ours_C"].view(torch.uint8), out["reference_C"].view(torch.uint8)))
    row = dict(tag=tag, type=t, m=m, n=n, k=k, N=N, backend="INT8" if be == 0 else "FP8", fast=bool(fast), phi=phi, ours=out["ours"], reference=out["reference"],
               bit_identical=same, speedup=round(out["reference"]["ms"] / out["ours"]["ms"], 3))
    rows.append(row)
    print(json.dumps(row), flush=True)
    del out

for fast in (False, True):
    run_case("S6", "s", 1024, 1024, 1024, 6, 0, fast)
    run_case("D14", "d", 8192, 8192, 8192, 14, 0, fast)
    run_case("Z18", "z", 4096, 4096, 4096, 18, 0, fast)
    for N in ((8, 14, 20) if quick else (8, 10, 12, 14, 16, 18, 20)):
        run_case("F8", "d", 8192, 8192, 8192, N, 1, fast)
# testing/test_accuracy.hpp: m = n = 128, k in 1024..65536, phi in {-1, 0, 0.5, 1, 2, 4}, SGEMM N=6 and DGEMM N=14 (INT8)
for t, N in (("s", 6), ("d", 14)):
    for phi in ((-1.0, 1.0, 4.0) if quick else (-1.0, 0.0, 0.5, 1.0, 2.0, 4.0)):
        for k in ((1024, 65536) if quick else (1024, 4096, 16384, 65536)):
            for fast in (False, True):
                run_case("ACC", t, 128, 128, k, N, 0, fast, phi=phi, warm=2, reps=3, corner=128)

Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/config_table.json").write_text(json.dumps(rows, indent=1))
print("\n| config | type | m=n | k | N | backend | mode | phi | ours TFLOPS | ref TFLOPS | speed-up | ours err_max | ref err_max | bit-identical |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r['tag']} | {r['type']} | {r['m']} | {r['k']} | {r['N']} | {r['backend']} | {'fast' if r['fast'] else 'accu'} | {r['phi']} | {r['ours']['tflops']} | "
          f"{r['reference']['tflops']} | {r['speedup']} | {r['ours']['err_max']:.2e} | {r['reference']['err_max']:.2e} | {r['bit_identical']} |")
