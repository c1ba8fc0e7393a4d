#!/usr/bin/env python3
"""Round-2 code paths under compute-sanitizer (small shapes): the native host-buffer pipeline (g8_gemm_host, all types / backends), the
native multi-GPU driver with a world of one rank (mailbox all-reduce, flag barrier, peer copies, chained bound GEMM, scatter epilogue,
shard-sum CRT, complex recombination), the skip-scaling sequence and the chained bound GEMM."""
import ctypes
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import torch
import helpers as H
import gemmul8_b200 as g8
from gemmul8_b200 import _lib, api

lib = _lib.load()
rng = np.random.default_rng(4)
for dtype, N, be in ((np.float64, 14, 0), (np.complex64, 6, 0), (np.float64, 9, 1), (np.complex128, 7, 1)):
    for fast in (False, True):
        m, n, k = 130, 520, 300
        A = H.rand_matrix(rng, (k, m), dtype); B = H.rand_matrix(rng, (k, n), dtype); C0 = H.rand_matrix(rng, (m, n), dtype)
        want = H.run_gemm(A, B, "T", "N", N, fast, alpha=0.5, beta=-1.0, C0=C0, backend=be)
        dA, lda = H.to_dev_colmajor(A); dB, ldb = H.to_dev_colmajor(B); dC, ldc = H.to_dev_colmajor(C0)
        hA, hB, hC = dA.cpu().pin_memory(), dB.cpu().pin_memory(), dC.cpu().pin_memory()
        plan = g8.NativeHostGemm(m, n, k, H.NP2T[np.dtype(dtype)], N, fast, "T", "N", chunk=256, backend=be)
        plan.run(hA, hB, hC, 0.5, -1.0, lda, ldb, ldc)
        torch.cuda.synchronize()
        plan.close()
        got = hC.numpy().reshape(n, ldc)[:, :m].T
        print(f"host pipeline {np.dtype(dtype).name} N={N} be={be} fast={fast} bit-identical={H.bits_equal(np.ascontiguousarray(got), want)}", flush=True)
for dtype, N in ((np.float64, 14), (np.complex128, 8)):
    for fast in (False, True):
        m, n, k = 100, 256, 300
        tdt = H.NP2T[np.dtype(dtype)]
        A = H.rand_matrix(rng, (m, k), dtype); B = H.rand_matrix(rng, (k, n), dtype)
        want = H.run_gemm(A, B, "N", "N", N, fast)
        dA, lda = H.to_dev_colmajor(A); dB, ldb = H.to_dev_colmajor(B); dC, ldc = H.to_dev_colmajor(np.zeros((m, n), dtype=dtype))
        comm, plan = ctypes.c_void_p(), ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        api._check(lib.g8_mg_comm_create(ctypes.byref(comm), 1, 0, 8 * (m + n) + 4096, handle), "comm")
        api._check(lib.g8_mg_comm_connect(comm, handle), "connect")
        api._check(lib.g8_mg_plan_create(ctypes.byref(plan), comm, api._DTYPES[tdt], 0, 0, m, n, k, N, int(fast)), "plan")
        keep = []
        api._check(lib.g8_gemm_mg(plan, api._scalar_ptr(1.0, tdt, keep), dA.data_ptr(), lda, dB.data_ptr(), ldb, api._scalar_ptr(0.0, tdt, keep), dC.data_ptr(), ldc,
                                  torch.cuda.current_stream().cuda_stream), "gemm_mg")
        torch.cuda.synchronize()
        got = H.from_dev_colmajor(dC, m, n, ldc)
        lib.g8_mg_plan_destroy(plan); lib.g8_mg_comm_destroy(comm)
        print(f"native mg world=1 {np.dtype(dtype).name} N={N} fast={fast} maxdiff={np.abs(got - want).max() / np.abs(want).max():.1e}", flush=True)
print("sanitizer extra workload done")
