#!/usr/bin/env python3
"""All single-GPU configurations of BASELINE.json, ours next to the unmodified reference library on the same B200:
  [S6]  SGEMM 1024^3 INT8 N=6        [D14] DGEMM 8192^3 INT8 N=14       [Z18] ZGEMM 4096^3 INT8 N=18
  [F8]  DGEMM 8192^3 FP8 N=8..20     + the reference's accuracy shape (testing/test_accuracy.hpp: m=n=128, k sweep, phi sweep)
Protocol of testing/test_flops.hpp:169-216 (warm-up + timed calls bracketed by CUDA events, median; TFLOPS = 2mnk (x4 complex) / t).
Errors: max / median relative error of a corner block against a wide reference (float64 for S, float64 matmul for D/Z -- a
double-double reference is not needed to see the emulated precision of N <= 14; larger N report the float64 noise floor).
Writes gpurun_out/config_table.json and a markdown table on stdout."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import refcompare

quick = "--quick" in sys.argv
rows = []

def run_case(*a, **kw):
    row = refcompare.run_case(*a, **kw)
    rows.append(row)
    print(json.dumps(row), flush=True)

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
