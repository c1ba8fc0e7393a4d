#!/usr/bin/env python3
"""tests/golden/worksize.json: gemmul8::workSize<is_Complex,backend> of the UNMODIFIED reference
(oracle/_ref/libgemmul8_ref.so -> ref_work_size), host-only call, for a grid of shapes."""
import ctypes
import itertools
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
R = ctypes.CDLL(str(ROOT / "oracle/_ref/libgemmul8_ref.so"))
R.ref_work_size.restype = ctypes.c_size_t
R.ref_work_size.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_size_t] * 3 + [ctypes.c_uint, ctypes.c_int, ctypes.c_int,
                                                                                    ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
rows = []
shapes = ((4, 3, 5), (1024, 1024, 1024), (8192, 8192, 8192), (16384, 16384, 2048), (4096, 4096, 4096), (100, 7, 3000), (1, 1, 1), (257, 255, 256))
for cplx, be, (m, n, k), N, ea, eb in itertools.product((0, 1), (0, 1), shapes, (2, 6, 7, 13, 14, 18, 20), (0, 1), (0, 1)):
    wa, wb = ctypes.c_size_t(), ctypes.c_size_t()
    t = R.ref_work_size(cplx, be, m, n, k, N, ea, eb, ctypes.byref(wa), ctypes.byref(wb))
    rows.append([cplx, be, m, n, k, N, ea, eb, t, wa.value, wb.value])
(ROOT / "tests/golden/worksize.json").write_text(json.dumps({"columns": "is_complex,backend,m,n,k,num_moduli,enA,enB,total,workSizeA,workSizeB", "rows": rows}))
print(len(rows), "rows")
