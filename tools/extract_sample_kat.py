#!/usr/bin/env python3
"""Parse the reference's only known-answer vector (GEMMul8/sample/dgemm_cuBLASLt_int8.cu:26-40:
A 4x5, B 5x3, hC_exact 4x3 as hex floats; N=15, accurate mode) into tests/golden/sample_kat.json."""
import json
import re
import sys
from pathlib import Path

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/GEMMul8/sample/dgemm_cuBLASLt_int8.cu").read_text()
HEXF = r"-?0x[01]\.[0-9a-fA-F]+p[+-]?\d+"


def vec(name):
    m = re.search(name + r"\s*=\s*\{(.*?)\};", src, re.S)
    return re.findall(HEXF, m.group(1))


out = dict(source="GEMMul8/sample/dgemm_cuBLASLt_int8.cu:26-40", m=4, n=3, k=5, num_moduli=15, fastmode=False,
           A=vec("hA"), B=vec("hB"), C_exact=vec("hC_exact"))
assert len(out["A"]) == 20 and len(out["B"]) == 15 and len(out["C_exact"]) == 12
dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "sample_kat.json"
dst.write_text(json.dumps(out, indent=1))
print("wrote", dst)
