"""Shared helpers of the parity tests: drive the C ABI (include/gemmul8_c.h) with numpy in / numpy out."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

import gemmul8_b200 as g8
from gemmul8_b200 import _lib, api

NP2T = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
        np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}


def to_dev_colmajor(X: np.ndarray, ld=None):
    """numpy (r x c) -> flat CUDA tensor holding the column-major matrix with leading dimension ld."""
    r, c = X.shape
    ld = ld or max(r, 1)
    buf = np.zeros((c, ld), dtype=X.dtype)
    buf[:, :r] = X.T
    return torch.from_numpy(buf.reshape(-1)).cuda(), ld


def from_dev_colmajor(t: torch.Tensor, r, c, ld):
    return t.cpu().numpy().reshape(c, ld)[:, :r].T.copy()


def stored_shape(op, rows, cols):
    """shape of the stored matrix X when op(X) is rows x cols"""
    return (rows, cols) if str(op).upper() in ("N", "0") else (cols, rows)


def rand_matrix(rng, shape, dtype, phi=0.5):
    dtype = np.dtype(dtype)
    def real(s):
        return ((rng.random(s) - 0.5) * np.exp(rng.standard_normal(s) * phi))
    if dtype.kind == "c":
        return (real(shape) + 1j * real(shape)).astype(dtype)
    return real(shape).astype(dtype)


def run_gemm(A, B, op_A="N", op_B="N", num_moduli=14, fastmode=False, alpha=1.0, beta=0.0, C0=None, lda=None, ldb=None,
             ldc=None, device_scalars=False, return_work=False, backend=0):
    """Run g8_gemm on numpy inputs. A, B are the STORED matrices. Returns C (m x n) [, dict of workspace pieces]."""
    dtype = np.result_type(A.dtype, B.dtype)
    m = A.shape[0] if op_A.upper() == "N" else A.shape[1]
    k = A.shape[1] if op_A.upper() == "N" else A.shape[0]
    n = B.shape[1] if op_B.upper() == "N" else B.shape[0]
    dA, lda = to_dev_colmajor(A, lda)
    dB, ldb = to_dev_colmajor(B, ldb)
    C0 = np.zeros((m, n), dtype=dtype) if C0 is None else C0.astype(dtype)
    dC, ldc = to_dev_colmajor(C0, ldc)
    cplx = dtype.kind == "c"
    tot, _, _ = g8.work_size(m, n, k, num_moduli, is_complex=cplx, backend=backend)
    work = torch.zeros(tot, dtype=torch.uint8, device="cuda")
    tdt = NP2T[np.dtype(dtype)]
    if device_scalars:
        alpha_t = torch.tensor([alpha], dtype=tdt, device="cuda")
        beta_t = torch.tensor([beta], dtype=tdt, device="cuda")
        g8.gemm(op_A, op_B, m, n, k, alpha_t, dA, lda, dB, ldb, beta_t, dC, ldc, num_moduli, fastmode, work, backend=backend)
    else:
        g8.gemm(op_A, op_B, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc, num_moduli, fastmode, work, backend=backend)
    torch.cuda.synchronize()
    C = from_dev_colmajor(dC, m, n, ldc)
    if not return_work:
        return C
    return C, read_workspace(work, m, n, k, num_moduli, cplx, backend=backend)


def read_workspace(work, m, n, k, num_moduli, cplx, enable_skip_scalA=False, enable_skip_scalB=False, backend=0):
    """Pull sftA/sftB, the residue planes and C_mid out of a (single-buffer) workspace."""
    L = api.layout(m, n, k, num_moduli, cplx, enable_skip_scalA, enable_skip_scalB, backend=backend)
    w = api.aligned_view(work).cpu().numpy()
    N, G = num_moduli, L.groups
    if backend == 1:  # FP8 backend: num_mat e4m3 planes per operand (x 3 plane sets Re / Im / Re+Im for complex), int16 C_mid
        from gemmul8_b200 import tables as T
        nm = T.num_mat("FP8", N)
        out = {"sftA": w[L.sftA:L.sftA + 2 * m].view(np.int16).copy(), "sftB": w[L.sftB:L.sftB + 2 * n].view(np.int16).copy()}
        rawA = w[L.A_lo:L.A_lo + L.sizeA * nm * G].reshape(G, nm, L.m_pad, L.k_pad)[:, :, :m]
        rawB = w[L.B_lo:L.B_lo + L.sizeB * nm * G].reshape(G, nm, n, L.k_pad)
        out["A_raw"], out["B_raw"] = rawA[0].copy(), rawB[0].copy()
        out["A_raw_sets"], out["B_raw_sets"] = [rawA[g].copy() for g in range(G)], [rawB[g].copy() for g in range(G)]
        cm = w[L.C_mid:L.C_mid + L.mid_bytes * L.sizeC * N].view(np.int16)
        out["C_mid"] = cm.reshape(N, n, L.m_pad, 2).copy() if cplx else cm.reshape(N, n, L.m_pad).copy()
        out["layout"] = L
        return out
    nA = num_moduli + int(enable_skip_scalA)
    nB = num_moduli + int(enable_skip_scalB)
    out = {}
    out["sftA"] = w[L.sftA:L.sftA + 2 * m].view(np.int16).copy()
    out["sftB"] = w[L.sftB:L.sftB + 2 * n].view(np.int16).copy()
    A_lo = w[L.A_lo:L.A_lo + L.sizeA * nA * G].view(np.int8).reshape(G, nA, L.m_pad, L.k_pad)
    B_lo = w[L.B_lo:L.B_lo + L.sizeB * nB * G].view(np.int8).reshape(G, nB, n, L.k_pad)
    out["A_lo"] = [A_lo[g, :N, :m].copy() for g in range(G)]
    out["B_lo"] = [B_lo[g, :N].copy() for g in range(G)]
    cm = w[L.C_mid:L.C_mid + L.mid_bytes * L.sizeC * N].view(np.int8)
    out["C_mid"] = cm.reshape(N, n, L.m_pad, 2).copy() if cplx else cm.reshape(N, n, L.m_pad).copy()
    out["layout"] = L
    return out


def bits_equal(x: np.ndarray, y: np.ndarray) -> bool:
    if x.shape != y.shape or x.dtype != y.dtype:
        return False
    return np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8))


def first_diff(x, y, name=""):
    d = np.argwhere(x != y)
    if len(d) == 0:
        return f"{name}: equal"
    i = tuple(d[0])
    return f"{name}: {len(d)} of {x.size} differ; first at {i}: {x[i]!r} vs {y[i]!r}"
