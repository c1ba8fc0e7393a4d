"""End-to-end emulated GEMM on HOST buffers: `gemm_host` / `HostGemm`.

The reference's API takes device pointers only; an application whose matrices live in host memory has to wrap the
call in three bulk copies (A, B in; C out) and the PCIe time dominates (1.5 GB at ~55 GB/s = 28 ms against ~7 ms of
GPU work for DGEMM 8192^3).  Because the emulation is separable along the columns of B / C -- the shift of column c
of B needs only that column (plus all of A), and C[:, c] needs only column c of the residue planes -- we stream B and
C in column chunks and overlap the copies with the stage kernels on three CUDA streams:

    h2d   : A .......| B[:,0] | B[:,1] | B[:,2] | ...
    comp  :          | splitA | chunk 0: splitB, GEMM(all moduli), CRT | chunk 1 ... |
    d2h   :                                                       | C[:,0] | C[:,1] | ...

Accurate mode needs the row maxima of the bound product over ALL columns before A can be split, so it runs the B side
(bound planes, bound GEMM, final B shifts, B split) chunk-wise while B streams in, then splits A and pipelines
GEMM / CRT / D2H.  Everything numerical is done by the same stage kernels as `g8_gemm` (C ABI `g8_stage_*`), so the
result is bit-identical to the monolithic call (tests/test_gpu_parity.py::test_host_pipeline_bitwise).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, api


class HostGemm:
    def __init__(self, m, n, k, dtype=torch.float64, num_moduli=14, fastmode=False, op_A="N", op_B="N", chunk=1024, device=None):
        self.lib = _lib.load()
        self.m, self.n, self.k = m, n, k
        self.dtype, self.N, self.fast = dtype, num_moduli, bool(fastmode)
        self.opA, self.opB = api._op(op_A), api._op(op_B)
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.cplx = dtype.is_complex
        self.dt = api._DTYPES[dtype]
        self.G = 3 if self.cplx else 1
        self.k_pad, self.m_pad, self.n_pad = api.pad256(k), api.pad256(m), api.pad256(n)
        self.sizeA, self.sizeB, self.sizeC = self.k_pad * self.m_pad, self.k_pad * n, self.m_pad * n
        self.chunk = max(128, min(chunk, n))
        self.pipelined = self.opB == 0  # column chunks of op(B) are contiguous in the stored B only for op N
        N, G = num_moduli, self.G
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.A_lo = torch.empty(self.sizeA * N * G, **u8)
        self.B_lo = torch.empty(self.sizeB * N * G, **u8)
        self.C_mid = torch.empty(self.sizeC * N * (2 if self.cplx else 1), **u8)
        self.sftA = torch.zeros(self.m_pad, dtype=torch.int16, device=self.dev)
        self.sftB = torch.zeros(self.n_pad, dtype=torch.int16, device=self.dev)
        self.maxes = torch.zeros(self.m_pad + self.n_pad, dtype=torch.int32, device=self.dev)
        rowsA, colsA = (m, k) if self.opA == 0 else (k, m)
        self.shapeA = (rowsA, colsA)
        self.dA = None  # allocated on first run (needs lda)
        self.dBc = [None, None]
        self.dCc = [None, None]
        self.s_h2d, self.s_comp, self.s_d2h = (torch.cuda.Stream(self.dev) for _ in range(3))
        self._ga = (ctypes.c_int * 3)(*((0, N, 2 * N) if self.cplx else (0, 0, 0)))
        self._gb = (ctypes.c_int * 3)(*((0, 1, 2) if self.cplx else (0, 0, 0)))

    # ---- thin stage wrappers (all on self.s_comp) ----
    def _chk(self, code, what):
        api._check(code, what)

    def _split(self, is_A, op, rows, X_ptr, ld, mode, sft_ptr, planes_ptr, plane_stride, group_stride):
        self._chk(self.lib.g8_stage_split(self.dt, int(is_A), op, rows, self.k, X_ptr, ld, self.N, mode, sft_ptr, planes_ptr,
                                          plane_stride, group_stride, self.s_comp.cuda_stream), "g8_stage_split")

    def _gemm(self, epi, A_ptr, B_ptr, ncols, units, ga, gb, out_ptr, rowmax_ptr, colmax_ptr):
        self._chk(self.lib.g8_stage_gemm(epi, 0, A_ptr, self.sizeA, B_ptr, self.sizeB, self.m, ncols, self.k_pad, units, 0, ga, gb,
                                         out_ptr, self.sizeC, self.m_pad, rowmax_ptr, colmax_ptr, self.s_comp.cuda_stream), "g8_stage_gemm")

    def run(self, hA, hB, hC, alpha=1.0, beta=0.0, lda=None, ldb=None, ldc=None):
        """hA, hB, hC: pinned host tensors holding the column-major matrices (flat, leading dimensions lda/ldb/ldc).
        Returns a CUDA event recorded after the last D2H copy (hC is valid once it has completed)."""
        m, n, k, N = self.m, self.n, self.k, self.N
        lda = lda or self.shapeA[0]
        ldb = ldb or (k if self.opB == 0 else n)
        ldc = ldc or m
        esz = hA.element_size()
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_h2d, self.s_comp, self.s_d2h):
            s.wait_stream(cur)
        keep = []
        pa, pb = api._scalar_ptr(alpha, self.dtype, keep), api._scalar_ptr(beta, self.dtype, keep)
        need_c_in = not (isinstance(beta, (int, float, complex)) and beta == 0)

        nA = lda * self.shapeA[1]
        if self.dA is None or self.dA.numel() < nA:
            self.dA = torch.empty(nA, dtype=self.dtype, device=self.dev)
        W = self.chunk if self.pipelined else n
        for i in range(2):
            nb = (ldb * W) if self.opB == 0 else (ldb * k)
            if self.dBc[i] is None or self.dBc[i].numel() < nb:
                self.dBc[i] = torch.empty(nb, dtype=self.dtype, device=self.dev)
            if self.dCc[i] is None or self.dCc[i].numel() < ldc * W:
                self.dCc[i] = torch.empty(ldc * W, dtype=self.dtype, device=self.dev)

        A_lo, B_lo, C_mid = self.A_lo.data_ptr(), self.B_lo.data_ptr(), self.C_mid.data_ptr()
        sftA, sftB = self.sftA.data_ptr(), self.sftB.data_ptr()
        rowmax = self.maxes.data_ptr()
        colmax = rowmax + 4 * self.m_pad
        mid = 2 if self.cplx else 1
        bound_ga = (ctypes.c_int * 3)(0, 1, 2)

        # ---- A: copy, then (fast) shift+split or (accurate) s0 + bound planes ----
        with torch.cuda.stream(self.s_h2d):
            self.dA[:nA].copy_(hA[:nA], non_blocking=True)
            evA = torch.cuda.Event(); evA.record(self.s_h2d)
        self.s_comp.wait_event(evA)
        if self.fast:
            self._split(True, self.opA, m, self.dA.data_ptr(), lda, 1, sftA, A_lo, self.sizeA, N)
        else:
            with torch.cuda.stream(self.s_comp):
                self.maxes.zero_()
            self._split(True, self.opA, m, self.dA.data_ptr(), lda, 2, sftA, A_lo, self.sizeA, N)

        chunks = [(c0, min(c0 + W, n)) for c0 in range(0, n, W)]
        ev_b_free = [None, None]   # compute finished reading dBc[i]
        ev_c_free = [None, None]   # d2h finished reading dCc[i]
        done = None

        def b_side(ci, c0, c1, mode):
            """H2D of B chunk + its split (mode 1 fast / accurate: bound -> bound GEMM -> final shift -> split)."""
            i = ci & 1
            nc = c1 - c0
            with torch.cuda.stream(self.s_h2d):
                if ev_b_free[i] is not None:
                    self.s_h2d.wait_event(ev_b_free[i])
                if self.opB == 0:
                    self.dBc[i][:ldb * nc].copy_(hB[c0 * ldb:c0 * ldb + ldb * nc], non_blocking=True)
                else:
                    self.dBc[i][:ldb * k].copy_(hB[:ldb * k], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(self.s_h2d)
            self.s_comp.wait_event(ev)
            X = self.dBc[i].data_ptr() + (0 if self.opB == 0 else c0 * esz)
            planes = B_lo + c0 * self.k_pad
            if mode == 1:
                self._split(False, self.opB, nc, X, ldb, 1, sftB + 2 * c0, planes, self.sizeB, N)
            else:
                self._split(False, self.opB, nc, X, ldb, 2, sftB + 2 * c0, planes, self.sizeB, N)
                self._gemm(4 if self.cplx else 2, A_lo, planes, nc, 1, bound_ga if self.cplx else self._gb, bound_ga if self.cplx else self._gb,
                           None, rowmax, colmax + 4 * c0)
                self._chk(self.lib.g8_stage_finalize_shift(sftB + 2 * c0, colmax + 4 * c0, nc, N, self.s_comp.cuda_stream), "finalize")
                self._split(False, self.opB, nc, X, ldb, 0, sftB + 2 * c0, planes, self.sizeB, N)
            e = torch.cuda.Event(); e.record(self.s_comp)
            ev_b_free[i] = e

        def c_side(ci, c0, c1):
            """GEMM over all moduli for the chunk, CRT into a device chunk, D2H."""
            nonlocal done
            i = ci & 1
            nc = c1 - c0
            self._gemm(3 if self.cplx else 0, A_lo, B_lo + c0 * self.k_pad, nc, N, self._ga, self._ga, C_mid + c0 * self.m_pad * mid, None, None)
            if ev_c_free[i] is not None:
                self.s_comp.wait_event(ev_c_free[i])
            if need_c_in:
                with torch.cuda.stream(self.s_h2d):
                    if ev_c_free[i] is not None:
                        self.s_h2d.wait_event(ev_c_free[i])
                    self.dCc[i][:ldc * nc].copy_(hC[c0 * ldc:c0 * ldc + ldc * nc], non_blocking=True)  # padding rows travel too (harmless)
                    e = torch.cuda.Event(); e.record(self.s_h2d)
                self.s_comp.wait_event(e)
            self._chk(self.lib.g8_stage_crt(self.dt, C_mid + c0 * self.m_pad * mid, self.m_pad, self.sizeC, m, nc, N, self.dCc[i].data_ptr(), ldc,
                                            sftA, sftB + 2 * c0, pa, pb, self.s_comp.cuda_stream), "g8_stage_crt")
            e = torch.cuda.Event(); e.record(self.s_comp)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(e)
                if ldc == m:
                    hC[c0 * ldc:c0 * ldc + ldc * nc].copy_(self.dCc[i][:ldc * nc], non_blocking=True)
                else:  # keep the caller's padding rows m..ldc untouched: strided (2-D) copy of the m valid rows per column
                    hC[c0 * ldc:c0 * ldc + ldc * nc].view(nc, ldc)[:, :m].copy_(self.dCc[i][:ldc * nc].view(nc, ldc)[:, :m], non_blocking=True)
                e2 = torch.cuda.Event(); e2.record(self.s_d2h)
            ev_c_free[i] = e2
            done = e2

        if self.fast:
            for ci, (c0, c1) in enumerate(chunks):
                b_side(ci, c0, c1, 1)
                c_side(ci, c0, c1)
        else:
            for ci, (c0, c1) in enumerate(chunks):
                b_side(ci, c0, c1, 2)
            self._chk(self.lib.g8_stage_finalize_shift(sftA, rowmax, m, N, self.s_comp.cuda_stream), "finalize")
            self._split(True, self.opA, m, self.dA.data_ptr(), lda, 0, sftA, A_lo, self.sizeA, N)
            for ci, (c0, c1) in enumerate(chunks):
                c_side(ci, c0, c1)
        cur.wait_event(done)
        self._keep = keep
        return done


class NativeHostGemm:
    """The NATIVE host-buffer pipeline (C ABI g8_host_plan_create / g8_gemm_host / g8_host_plan_destroy, csrc/g8_host.cu): the same
    chunked three-stream schedule as HostGemm above, driven entirely from C++, for all four types, both backends and every
    op_A / op_B.  This class only holds the opaque plan; `run` enqueues one call on torch's current stream (stream-ordered, does not
    block the host) -- what bench.py's `e2e` leg measures."""

    def __init__(self, m, n, k, dtype=torch.float64, num_moduli=14, fastmode=False, op_A="N", op_B="N", chunk=1024, device=None, backend=0):
        self.lib = _lib.load()
        self.dtype = dtype
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.plan = ctypes.c_void_p()
        with torch.cuda.device(self.dev):
            api._check(self.lib.g8_host_plan_create(ctypes.byref(self.plan), api._DTYPES[dtype], int(backend), api._op(op_A), api._op(op_B), m, n, k,
                                                    num_moduli, int(bool(fastmode)), chunk), "g8_host_plan_create")
        rowsA = m if api._op(op_A) == 0 else k
        self.defaults = (rowsA, k if api._op(op_B) == 0 else n, m)

    def run(self, hA, hB, hC, alpha=1.0, beta=0.0, lda=None, ldb=None, ldc=None):
        keep = []
        pa, pb = api._scalar_ptr(alpha, self.dtype, keep), api._scalar_ptr(beta, self.dtype, keep)
        with torch.cuda.device(self.dev):
            api._check(self.lib.g8_gemm_host(self.plan, pa, hA.data_ptr(), lda or self.defaults[0], hB.data_ptr(), ldb or self.defaults[1], pb,
                                             hC.data_ptr(), ldc or self.defaults[2], torch.cuda.current_stream(self.dev).cuda_stream), "g8_gemm_host")

    def close(self):
        if self.plan:
            self.lib.g8_host_plan_destroy(self.plan)
            self.plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gemm_host(op_A, op_B, m, n, k, alpha, hA, lda, hB, ldb, beta, hC, ldc, num_moduli=14, fastmode=False, chunk=1024, device=None,
              plan=None, backend=0, native=True):
    """One-shot convenience (pass `plan=` to reuse the device buffers across calls).  Synchronises.  native=True (default) runs the
    C++ pipeline (g8_gemm_host); native=False the Python-orchestrated one (INT8 backend only)."""
    if plan is None:
        plan = (NativeHostGemm(m, n, k, hC.dtype, num_moduli, fastmode, op_A, op_B, chunk, device, backend) if native else
                HostGemm(m, n, k, hC.dtype, num_moduli, fastmode, op_A, op_B, chunk, device))
    ev = plan.run(hA, hB, hC, alpha, beta, lda, ldb, ldc)
    if ev is not None:
        ev.synchronize()
    else:
        torch.cuda.current_stream(plan.dev).synchronize()
    return plan
