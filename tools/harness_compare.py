#!/usr/bin/env python3
"""Compare the two CSVs written by the reference's OWN benchmark harness (testing/test.cu, unmodified, compiled by oracle/Makefile):
once linked against the reference library (oracle/_ref/harness_ref) and once against this repo's libgemmul8.a (oracle/_ref/harness_ours).
Same inputs, same timing protocol, same error evaluation (double-double reference product): the accuracy columns must be IDENTICAL for
the emulated rows (bit-identical C), the TFLOPS column gives the speed-up under the reference's own protocol.
usage: harness_compare.py ref.csv ours.csv [out.json]"""
import csv
import json
import statistics
import sys


def load(path):
    rows = {}
    for r in csv.reader(open(path)):
        if len(r) < 9 or r[0] == "phi":
            continue
        try:
            key = (int(r[1]), int(r[2]), int(r[3]), r[4])
            rows[key] = dict(err_max=r[5], err_med=r[6], tflops=float(r[7]), t=float(r[8]), phases=[float(x) if x else None for x in r[9:13]])
        except ValueError:
            continue
    return rows


ref, ours = load(sys.argv[1]), load(sys.argv[2])
common = [k for k in ref if k in ours]
emu = [k for k in common if k[3].startswith("OS2")]
same_err = [k for k in emu if ref[k]["err_max"] == ours[k]["err_max"] and ref[k]["err_med"] == ours[k]["err_med"]]
diff = [k for k in emu if k not in same_err]
sp = [ours[k]["tflops"] / ref[k]["tflops"] for k in emu]
big = [k for k in emu if k[0] == 8192]
out = {
    "rows_ref": len(ref), "rows_ours": len(ours), "common_rows": len(common), "emulated_rows_compared": len(emu),
    "accuracy_columns_identical": len(same_err), "accuracy_columns_differ": [list(k) + [ref[k]["err_max"], ours[k]["err_max"]] for k in diff[:20]],
    "speedup_tflops_ours_over_ref": {"min": round(min(sp), 3), "median": round(statistics.median(sp), 3), "max": round(max(sp), 3)} if sp else None,
    "rows_m8192": [{"m": k[0], "k": k[2], "fn": k[3], "ref_tflops": round(ref[k]["tflops"], 1), "ours_tflops": round(ours[k]["tflops"], 1),
                    "err_max": ours[k]["err_max"]} for k in sorted(big, key=lambda k: (k[2], k[3])) if k[3] in ("OS2-fast-14", "OS2-accu-14")],
    "native_dgemm_rows_tflops": {f"{k[0]}x{k[2]}": [round(ref[k]["tflops"], 1), round(ours[k]["tflops"], 1)] for k in common if k[3] == "DGEMM" and k[0] >= 4096},
}
print(json.dumps(out, indent=1))
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(json.dumps(out, indent=1))
