#!/usr/bin/env python3
"""Per-stage kernel timings (CUDA events, back-to-back launches) for DGEMM 8192^3 N=14 shapes: split of A (row-strided) and
B (row-contiguous) in modes 0 / 1 / 2, the bound GEMM, the contraction and the CRT.  Prints one JSON line; environment switches of the
kernels (G8_SPLIT_TILES_PER_BLOCK, G8_ACCU_STAGE1_CTAS_PER_SM, G8_SPLIT_ROW_CACHE, G8_SPLIT_FUSED_ACCU, G8_GEMM_GROUP ...) are read
once per process, so sweeps run this script several times.   usage: stage_bench.py [S] [N] [reps]"""
import ctypes
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gemmul8_b200 as g8
from gemmul8_b200 import _lib, api

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lib = _lib.load()
dt = torch.float64
m = n = k = S
A = g8.randmat(m, k, dt, seed=12345)
B = g8.randmat(k, n, dt, seed=54321)
C = torch.zeros(m * n, dtype=dt, device="cuda")
kp, mp = api.pad256(k), api.pad256(m)
A_lo = torch.empty(N * kp * mp, dtype=torch.int8, device="cuda")
B_lo = torch.empty(N * kp * n, dtype=torch.int8, device="cuda")
C_mid = torch.empty(N * mp * n, dtype=torch.int8, device="cuda")
sftA = torch.zeros(mp, dtype=torch.int16, device="cuda")
sftB = torch.zeros(api.pad256(n), dtype=torch.int16, device="cuda")
rowmax = torch.zeros(mp, dtype=torch.int32, device="cuda")
colmax = torch.zeros(api.pad256(n), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
one, zero = (ctypes.c_double * 1)(1.0), (ctypes.c_double * 1)(0.0)


def split(is_A, mode):
    X, rows, ld, sft, planes, stride = (A, m, m, sftA, A_lo, kp * mp) if is_A else (B, n, k, sftB, B_lo, kp * n)
    return lambda: api._check(lib.g8_stage_split(1, int(is_A), 0, rows, k, X.data_ptr(), ld, N, mode, sft.data_ptr(), planes.data_ptr(), stride, N, st), "split")


def timeit(fn, r=reps):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(r):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / r, 4)


out = {"S": S, "N": N, "env": {k_: v for k_, v in os.environ.items() if k_.startswith("G8_")}}
out["splitA_fast_stats+split"] = timeit(split(True, 1))
out["splitB_fast"] = timeit(split(False, 1))
out["stage1_A_accu"] = timeit(split(True, 2))
out["stage1_B_accu"] = timeit(split(False, 2))
out["bound_gemm"] = timeit(lambda: api._check(lib.g8_stage_gemm(2, 0, A_lo.data_ptr(), kp * mp, B_lo.data_ptr(), kp * n, m, n, kp, 1, 0, None, None, None, 0, mp,
                                                                   rowmax.data_ptr(), colmax.data_ptr(), st), "bound"))
# sensible shifts for the mode-0 splits: run the fast path once, keep its shifts
split(True, 1)(); split(False, 1)()
out["splitA_mode0"] = timeit(split(True, 0))
out["splitB_mode0"] = timeit(split(False, 0))
out["gemm_all_moduli"] = timeit(lambda: api._check(lib.g8_stage_gemm(0, 0, A_lo.data_ptr(), kp * mp, B_lo.data_ptr(), kp * n, m, n, kp, N, 0, None, None, C_mid.data_ptr(),
                                                                        mp * n, mp, None, None, st), "gemm"), max(3, reps // 4))
out["crt"] = timeit(lambda: api._check(lib.g8_stage_crt(1, C_mid.data_ptr(), mp, mp * n, m, n, N, C.data_ptr(), m, sftA.data_ptr(), sftB.data_ptr(),
                                                         ctypes.addressof(one), ctypes.addressof(zero), st), "crt"))
tot, _, _ = g8.work_size(m, n, k, N)
work = torch.empty(tot, dtype=torch.uint8, device="cuda")
for fast in (False, True):
    out["gemm_call_" + ("fast" if fast else "accu")] = timeit(lambda: g8.gemm("N", "N", m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, fast, work), max(5, reps // 2))
gb = lambda bytes_, ms: round(bytes_ / ms * 1e-6, 1)
out["GBps"] = {"splitA_mode0": gb(m * k * (8 + N), out["splitA_mode0"]), "splitB_mode0": gb(n * k * (8 + N), out["splitB_mode0"]),
               "crt": gb(m * n * (8 + N), out["crt"]), "stage1_A_accu": gb(m * k * 9, out["stage1_A_accu"]), "stage1_B_accu": gb(n * k * 9, out["stage1_B_accu"])}
print(json.dumps(out))
