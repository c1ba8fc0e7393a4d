#!/usr/bin/env python3
"""FP8 backend, accurate mode: how often do our shift exponents differ from the reference's, and by how much does C differ then?
(The FP8 bound product is accumulated in binary32 by different kernels -- cuBLASLt vs tcgen05 -- and then inflated, so a shift can
land on the other side of a floor(); INT8 shifts are integer-exact and always identical.)  DGEMM S^3 FP8 for num_moduli = 8 .. 20, same
inputs for both libraries (harness generator); per N: number of differing sftA / sftB entries, the largest difference, whether C is
bit-identical, and max |C_ours - C_ref| / max |C| in units of the emulated precision (the reference's own error against float64).
Writes gpurun_out/fp8_shift_stats.json.   usage: fp8_shift_stats.py [S] [phi ...]"""
import ctypes
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import gemmul8_b200 as g8
from gemmul8_b200 import api
import refcompare

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
phis = [float(x) for x in sys.argv[2:]] or [-1.0, 1.0]
ref = refcompare.ref_lib()
st = torch.cuda.current_stream()
dt = torch.float64
rows = []
for phi in phis:
    A = g8.randmat(S, S, dt, phi=phi, seed=12345)
    B = g8.randmat(S, S, dt, phi=phi, seed=54321)
    want = A.view(S, S).t()[:256, :] @ B.view(S, S).t()[:, :256]
    for N in range(8, 21, 2):
        keep = []
        one, zero = api._scalar_ptr(1.0, dt, keep), api._scalar_ptr(0.0, dt, keep)
        L = api.layout(S, S, S, N, False, backend=1)
        res = {}
        for impl in ("ours", "ref"):
            C = torch.zeros(S * S, dtype=dt, device="cuda")
            tot = g8.work_size(S, S, S, N, backend=1)[0]
            work = torch.empty(tot, dtype=torch.uint8, device="cuda")
            if impl == "ours":
                g8.gemm("N", "N", S, S, S, 1.0, A, S, B, S, 0.0, C, S, N, False, work, backend=1)
            else:
                code = ref.L.ref_gemm(1, 1, 1, 0, 0, S, S, S, one, A.data_ptr(), S, B.data_ptr(), S, zero, C.data_ptr(), S, N, 0, work.data_ptr(), None, None,
                                      0, 0, 0, 0, ctypes.c_void_p(st.cuda_stream), None)
                assert code == 0
            torch.cuda.synchronize()
            w = api.aligned_view(work)
            res[impl] = (C, w[L.sftA:L.sftA + 2 * S].view(torch.int16).clone(), w[L.sftB:L.sftB + 2 * S].view(torch.int16).clone())
            del work
        (Co, sAo, sBo), (Cr, sAr, sBr) = res["ours"], res["ref"]
        dA, dB = (sAo.int() - sAr.int()), (sBo.int() - sBr.int())
        scale = Cr.abs().max()
        err_ref = float(((Cr.view(S, S).t()[:256, :256] - want).abs().max() / want.abs().max()))
        diff = float((Co - Cr).abs().max() / scale)
        row = dict(S=S, phi=phi, N=N, sftA_diff=int((dA != 0).sum()), sftB_diff=int((dB != 0).sum()), max_abs_shift_diff=int(max(dA.abs().max(), dB.abs().max())),
                   ours_minus_ref_sign=[int((dA > 0).sum() + (dB > 0).sum()), int((dA < 0).sum() + (dB < 0).sum())],
                   C_bit_identical=bool(torch.equal(Co.view(torch.int64), Cr.view(torch.int64))), max_C_diff_over_max=diff,
                   ref_err_vs_fp64_over_max=err_ref, diff_in_units_of_ref_err=(diff / err_ref if err_ref > 0 else 0.0))
        rows.append(row)
        print(json.dumps(row), flush=True)
        del res, Co, Cr
        torch.cuda.empty_cache()
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/fp8_shift_stats.json").write_text(json.dumps(rows, indent=1))
