"""Host-side mirror of the reference's public interface for the hot path
(reference include/gemmul8.hpp:17-94): `work_size`, `gemm`, plus the workspace layout helpers the
parity tests use.  Everything computes on the GPU through the C ABI (include/gemmul8_c.h); torch is
used only for device memory and streams."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from enum import IntEnum

import torch

from . import _lib
from . import tables as T


class Backend(IntEnum):  # gemmul8::Backend (include/gemmul8.hpp:19-20)
    INT8 = 0
    FP8 = 1


class Op(IntEnum):  # cublasOperation_t
    N = 0
    T = 1
    C = 2


_DTYPES = {torch.float32: 0, torch.float64: 1, torch.complex64: 2, torch.complex128: 3}


class Gemmul8Error(RuntimeError):
    pass


def _check(code: int, what: str):
    if code != 0:
        names = {10001: "INVALID_VALUE", 10002: "NOT_SUPPORTED", 10003: "NO_DEVICE_CODE (needs an sm_100a GPU; no fallback)"}
        raise Gemmul8Error(f"{what} failed: status {code} {names.get(code, '(cudaError_t)')}")


def _op(op) -> int:
    if isinstance(op, str):
        return int(Op[op.upper()])
    return int(op)


def pad256(x: int) -> int:
    return 256 * ((x + 255) // 256)


def work_size(m, n, k, num_moduli, is_complex=False, backend=Backend.INT8, enable_skip_scalA=False, enable_skip_scalB=False):
    """gemmul8::workSize<is_Complex, backend> (include/gemmul8.hpp:25-35).  Returns (total, workSizeA, workSizeB) bytes."""
    lib = _lib.load()
    wa, wb = ctypes.c_size_t(0), ctypes.c_size_t(0)
    tot = lib.g8_work_size(int(is_complex), int(backend), m, n, k, num_moduli, int(enable_skip_scalA), int(enable_skip_scalB),
                           ctypes.byref(wa), ctypes.byref(wb))
    return int(tot), int(wa.value), int(wb.value)


def _scalar_ptr(x, dtype, keep):
    """alpha / beta: python scalar -> host buffer; torch CUDA tensor -> device pointer (read on the device)."""
    if isinstance(x, torch.Tensor):
        assert x.dtype == dtype
        keep.append(x)
        return x.data_ptr()
    if dtype in (torch.complex64, torch.complex128):
        ct = ctypes.c_float if dtype == torch.complex64 else ctypes.c_double
        z = complex(x)
        buf = (ct * 2)(z.real, z.imag)
    else:
        ct = ctypes.c_float if dtype == torch.float32 else ctypes.c_double
        buf = (ct * 1)(float(x))
    keep.append(buf)
    return ctypes.addressof(buf)


def gemm(op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, num_moduli, fastmode, work,
         workA=None, workB=None, enable_skip_scalA=False, enable_skip_scalB=False, skip_scalA=False, skip_scalB=False,
         backend=Backend.INT8, stream=None, timing=False):
    """gemmul8::gemm / gemmul8::gemmLt (include/gemmul8.hpp:41-94): C = alpha*op(A)*op(B) + beta*C, column-major.

    A, B, C, work[A|B] are CUDA torch tensors used as flat device buffers (C's dtype selects S/D/C/ZGEMM).
    Returns the 4 phase times in ns (zeros unless timing=True, which synchronises like the reference does)."""
    lib = _lib.load()
    if C.dtype not in _DTYPES or A.dtype != C.dtype or B.dtype != C.dtype:
        raise Gemmul8Error("A, B, C must share one of float32/float64/complex64/complex128")
    keep = []
    d = _lib.GemmDesc()
    d.dtype, d.backend, d.op_A, d.op_B = _DTYPES[C.dtype], int(backend), _op(op_A), _op(op_B)
    d.m, d.n, d.k = m, n, k
    d.alpha, d.beta = _scalar_ptr(alpha, C.dtype, keep), _scalar_ptr(beta, C.dtype, keep)
    d.A, d.lda, d.B, d.ldb, d.C, d.ldc = A.data_ptr(), lda, B.data_ptr(), ldb, C.data_ptr(), ldc
    d.num_moduli, d.fastmode = num_moduli, int(bool(fastmode))
    d.work = work.data_ptr()
    d.workA = workA.data_ptr() if workA is not None else None
    d.workB = workB.data_ptr() if workB is not None else None
    d.enable_skip_scalA, d.enable_skip_scalB = int(enable_skip_scalA), int(enable_skip_scalB)
    d.skip_scalA, d.skip_scalB = int(skip_scalA), int(skip_scalB)
    s = stream if stream is not None else torch.cuda.current_stream(C.device)
    d.stream = s.cuda_stream
    phases = (ctypes.c_double * 4)()
    with torch.cuda.device(C.device):
        code = lib.g8_gemm(ctypes.byref(d), phases if timing else None)
    _check(code, "g8_gemm")
    return list(phases)


def matmul(A: torch.Tensor, B: torch.Tensor, num_moduli=14, fastmode=False, out=None, work=None):
    """Convenience: row-major torch matrices, C = A @ B through the emulator (C^T = B^T A^T in BLAS terms)."""
    assert A.is_cuda and A.dim() == 2 and B.dim() == 2 and A.shape[1] == B.shape[0]
    A, B = A.contiguous(), B.contiguous()
    m, k = A.shape
    n = B.shape[1]
    C = out if out is not None else torch.empty((m, n), dtype=A.dtype, device=A.device)
    if work is None:
        tot, _, _ = work_size(n, m, k, num_moduli, is_complex=A.is_complex())
        work = torch.empty(tot, dtype=torch.uint8, device=A.device)
    gemm(Op.N, Op.N, n, m, k, 1.0, B, n, A, k, 0.0, C, n, num_moduli, fastmode, work)
    return C


@dataclass
class Layout:
    """Byte offsets (from the 256-aligned base of `work`) of the reference-compatible workspace carve-up
    (gemmul8_real.hpp:95-107, gemmul8_complex.hpp:95-118) for the single-buffer case (workA = workB = None)."""
    k_pad: int
    m_pad: int
    n_pad: int
    sizeA: int
    sizeB: int
    sizeC: int
    groups: int
    A_lo: int
    sftA: int
    B_lo: int
    sftB: int
    C_mid: int
    mid_bytes: int


def layout(m, n, k, num_moduli, is_complex=False, enable_skip_scalA=False, enable_skip_scalB=False, backend=Backend.INT8) -> Layout:
    k_pad, m_pad, n_pad = pad256(k), pad256(m), pad256(n)
    sizeA, sizeB, sizeC = k_pad * m_pad, k_pad * n, m_pad * n
    g = 3 if is_complex else 1
    num_mat = T.num_mat("INT8" if int(backend) == 0 else "FP8", num_moduli)  # planes per operand group
    a_lo = 0
    sftA = a_lo + sizeA * (num_mat + int(enable_skip_scalA)) * g
    b_lo = sftA + 2 * m_pad
    sftB = b_lo + sizeB * (num_mat + int(enable_skip_scalB)) * g
    c_mid = sftB + 2 * n_pad
    mid = (1 if int(backend) == 0 else 2) * (2 if is_complex else 1)
    return Layout(k_pad, m_pad, n_pad, sizeA, sizeB, sizeC, g, a_lo, sftA, b_lo, sftB, c_mid, mid)


def aligned_view(work: torch.Tensor) -> torch.Tensor:
    """uint8 view of `work` starting at its first 256-byte aligned address (common.hpp:34-39)."""
    off = (-work.data_ptr()) % 256
    return work.view(torch.uint8)[off:]


def randmat(rows, cols, dtype, phi=-1.0, seed=12345, device="cuda"):
    """Column-major rows x cols matrix from the reference harness' generator (testing/make_matrix.hpp:33-82),
    returned as a flat tensor of rows*cols elements (ld = rows)."""
    lib = _lib.load()
    X = torch.empty(rows * cols, dtype=dtype, device=device)
    with torch.cuda.device(X.device):
        _check(lib.g8_randmat(_DTYPES[dtype], X.data_ptr(), rows, cols, float(phi), int(seed),
                              torch.cuda.current_stream(X.device).cuda_stream), "g8_randmat")
    return X
