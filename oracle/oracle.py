"""TEST INFRASTRUCTURE ONLY -- Python driver of the CPU restatement in oracle/g8_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It restates the reference's INT8 Ozaki-II pipeline stage by stage
(gemmul8_real.hpp:53-211, gemmul8_complex.hpp:53-226) on the CPU:

    shifts (accurate: scaling_accu_real.hpp:23-226 / fast: scaling_fast_real.hpp:6-49)
 -> split  (scaling.hpp + mod.hpp)          -> A_lo, B_lo   int8 planes, K-major, ld = pad256(k)
 -> exact integer GEMM + symmetric mod p    -> C_mid        int8 planes, column-major, ld = pad256(m)
 -> CRT FMA chain + unscale + alpha/beta    -> C

Parity pin: `tests/test_oracle.py` checks it against the reference's known-answer vector
(sample/dgemm_cuBLASLt_int8.cu:26-40 -> tests/golden/sample_kat.json) and a pure-Python big-integer
model; on the GPU box `tests/test_gpu_ref_parity.py`, `tests/test_gpu_configs.py` and `tests/test_gpu_skip_scaling.py` check the CUDA
path against the unmodified reference library itself (oracle/_ref/libgemmul8_ref.so).  The device's __log2f cannot be reproduced
on a CPU, so shift exponents may be passed in (from the device) and are otherwise computed with an
exact log2 plus an `ambiguous` mask for rows that sit on a floor() boundary.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

from gemmul8_b200 import tables as T

_HERE = Path(__file__).resolve().parent
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "_build" / "libg8oracle.so"
        src = _HERE / "g8_oracle.c"
        if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            subprocess.check_call(["make", "-C", str(_HERE), "oracle"], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(str(so))
        L.g8o_accu_shift.restype = ctypes.c_int32
        L.g8o_accu_shift.argtypes = [ctypes.c_int32, ctypes.c_float, ctypes.POINTER(ctypes.c_int)]
        L.g8o_fast_shift.restype = ctypes.c_int32
        L.g8o_fast_shift.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_float, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_int)]
        L.g8o_sumsq_ru_d.restype = ctypes.c_double
        L.g8o_sumsq_ru_d.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t]
        _LIB = L
    return _LIB


def pad256(x: int) -> int:
    return 256 * ((x + 255) // 256)


def _p(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _sz(x):
    return ctypes.c_size_t(int(x))


def _ilogb(x: float) -> int:
    return 0 if x == 0 else int(np.floor(np.log2(abs(float(x))))) if np.isfinite(x) else 0


def ilogb_exact(x: float) -> int:
    """ilogb with ilogb(0) := 0 (template_math.hpp:96-97)."""
    if x == 0:
        return 0
    m, e = np.frexp(np.float64(abs(x)))
    return int(e) - 1


def _components(X: np.ndarray):
    """Real view(s) of a (possibly complex) column-major matrix: list of (array, ld_in_elements, elem_stride)."""
    return X


class Operand:
    """op(X) addressed as rows x inner (rows = m for A, n for B)."""

    def __init__(self, X: np.ndarray, op: str, is_A: bool):
        # X is a 2-D numpy array holding the column-major matrix as X[row, col] (Fortran order preferred).
        self.X = np.asfortranarray(X)
        self.op = op.upper()
        # For A: op(A) is m x k; "row r, inner l" = op(A)[r, l].  For B: op(B) is k x n; "row c, inner l" = op(B)[l, c].
        if is_A:
            view = self.X if self.op == "N" else (self.X.T if self.op == "T" else self.X.conj().T)
        else:
            view = self.X.T if self.op == "N" else (self.X if self.op == "T" else self.X.conj())
        self.view = np.ascontiguousarray(view)  # rows x inner, C-order => row r contiguous along inner
        self.rows, self.inner = self.view.shape


def _real_parts(v: np.ndarray):
    if np.iscomplexobj(v):
        return np.ascontiguousarray(v.real), np.ascontiguousarray(v.imag)
    return (v,)


def accurate_shifts(A: Operand, B: Operand, num_moduli: int, backend="INT8"):
    """Accurate-mode shift exponents (scaling_accu_real.hpp / scaling_accu_complex.hpp).
    Returns sftA, sftB (int16, reference sign convention: stored NEGATED) and ambiguity masks."""
    L = lib()
    log2P = np.float32(T.log2P(backend, num_moduli))
    cplx = np.iscomplexobj(A.view)
    k_pad = pad256(A.inner)
    f32 = A.view.dtype in (np.float32, np.complex64)

    def s0_and_bar(O: Operand):
        parts = _real_parts(O.view)
        amax = np.zeros(O.rows, dtype=parts[0].dtype)
        for pz in parts:
            amax = np.maximum(amax, np.abs(pz).max(axis=1) if O.inner else 0)
        s0 = np.array([5 - ilogb_exact(float(a)) for a in amax], dtype=np.int16)
        bars = []
        for pz in parts:
            bar = np.zeros((O.rows, k_pad), dtype=np.int8)
            fn = L.g8o_extract_f if f32 else L.g8o_extract_d
            # trans=1: row r contiguous with ld = inner
            fn(_p(pz), _sz(O.inner), 1, _sz(O.rows), _sz(O.inner), _sz(k_pad), _p(s0), _p(bar))
            bars.append(bar.astype(np.int64))
        return s0, bars

    s0A, Abar = s0_and_bar(A)
    s0B, Bbar = s0_and_bar(B)
    if not cplx:
        Cbar = Abar[0] @ Bbar[0].T
        rowmax = Cbar.max(axis=1, initial=0)
        colmax = Cbar.max(axis=0, initial=0)
    else:
        # find_max.hpp:99-114: bound of |Re| and |Im| of the product
        ar, ai = Abar
        br, bi = Bbar
        c_re = ar @ br.T + ai @ bi.T
        c_im = ar @ bi.T + ai @ br.T
        cm = np.maximum(c_re, c_im)
        rowmax = cm.max(axis=1, initial=0)
        colmax = cm.max(axis=0, initial=0)
    rowmax = np.maximum(rowmax, 0)
    colmax = np.maximum(colmax, 0)

    def fin(s0, mx):
        out = np.zeros(len(s0), dtype=np.int16)
        amb = np.zeros(len(s0), dtype=bool)
        for i, (s, x) in enumerate(zip(s0, mx)):
            a = ctypes.c_int(0)
            g = L.g8o_accu_shift(int(x), log2P, ctypes.byref(a))
            out[i] = np.int16(-(int(s) + g)) if x > 0 else 0
            amb[i] = bool(a.value)
        return out, amb

    sftA, ambA = fin(s0A, rowmax)
    sftB, ambB = fin(s0B, colmax)
    return sftA, sftB, ambA, ambB


def e4m3_round_up(x: np.ndarray) -> np.ndarray:
    """smallest e4m3 value >= x (x >= 0, x < 448), as float64 (scaling.hpp:48-54)"""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    q = t.to(torch.float8_e4m3fn)
    back = q.to(torch.float32)
    bits = q.view(torch.uint8).to(torch.int32) + (back.double() < torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))).to(torch.int32)
    return bits.to(torch.uint8).view(torch.float8_e4m3fn).to(torch.float64).numpy()


def accurate_shifts_fp8(A: Operand, B: Operand, num_moduli: int):
    """FP8 accurate mode (scaling_accu_real.hpp with maxUFP = 7): the device accumulates the bound product in f32 and
    inflates it by (k+1)*2^-24 (find_max.hpp:82-96); we bracket its maximum and flag rows whose floor() depends on it."""
    L = lib()
    log2P = np.float32(T.log2P("FP8", num_moduli))
    k = A.inner

    cplx = np.iscomplexobj(A.view)

    def s0_and_bar(O):
        parts = [np.asarray(pz, dtype=np.float64) for pz in _real_parts(O.view)]
        amax = np.zeros(O.rows)
        for pz in parts:
            amax = np.maximum(amax, np.abs(pz).max(axis=1) if O.inner else 0.0)
        s0 = np.array([7 - ilogb_exact(float(a)) for a in amax], dtype=np.int16)
        return s0, [e4m3_round_up(np.abs(pz) * np.exp2(s0.astype(np.float64))[:, None]) for pz in parts]

    s0A, Abar = s0_and_bar(A)
    s0B, Bbar = s0_and_bar(B)
    if not cplx:
        Cbar = Abar[0] @ Bbar[0].T  # exact enough in f64 (values < 2^8, k <= 2^16)
    else:
        # complex: |Re| and |Im| of the product are bounded by |Ar||Br| + |Ai||Bi| and |Ar||Bi| + |Ai||Br| (2k-term sums on the device)
        (ar, ai), (br, bi) = Abar, Bbar
        Cbar = np.maximum(ar @ br.T + ai @ bi.T, ar @ bi.T + ai @ br.T)
        k = 2 * k
    ku = (k + 1) * 2.0 ** -24

    def fin(s0, mx):
        out = np.zeros(len(s0), dtype=np.int16)
        amb = np.zeros(len(s0), dtype=bool)
        for i, (s, x) in enumerate(zip(s0, mx)):
            if x <= 0:
                amb[i] = True
                continue
            lo = float(np.float32(x * (1 - ku)))
            hi = float(np.float32(x * (1 + ku) * (1 + 2 * ku)))
            g_lo = int(np.floor(float(log2P) - float.fromhex("0x1.000006p-1") * np.log2(hi)))
            g_hi = int(np.floor(float(log2P) - float.fromhex("0x1.000006p-1") * np.log2(lo)))
            t = float(log2P) - float.fromhex("0x1.000006p-1") * np.log2(hi)
            amb[i] = (g_lo != g_hi) or (t - np.floor(t) < 2 ** -10) or (np.floor(t) + 1 - t < 2 ** -10)
            out[i] = np.int16(-(int(s) + g_lo))
        return out, amb

    sftA, ambA = fin(s0A, Cbar.max(axis=1, initial=0))
    sftB, ambB = fin(s0B, Cbar.max(axis=0, initial=0))
    return sftA, sftB, ambA, ambB


def decode_fp8_planes(raw: np.ndarray, num_moduli: int):
    """Device planes of the FP8 backend (uint8 e4m3 codes, [num_mat, rows, k_pad], order of table.hpp:69-75) -> int16 residues
    [num_moduli, rows, k_pad]; also checks the Karatsuba plane hi + lo."""
    import torch

    val = torch.from_numpy(np.ascontiguousarray(raw).view(np.uint8)).view(torch.float8_e4m3fn).to(torch.float32).numpy().astype(np.int32)
    out = np.zeros((num_moduli,) + raw.shape[1:], dtype=np.int16)
    ok = True
    for idx in range(num_moduli):
        if idx < 6:
            base, s = 2 * idx, T.FP8_SQRT_MODULI[idx]
            out[idx] = s * val[base] + val[base + 1]
        else:
            base = 12 + 3 * (idx - 6)
            out[idx] = 16 * val[base] + val[base + 1]
            ok &= bool(np.array_equal(val[base + 2], val[base] + val[base + 1]))
    return out, ok


def fast_shifts(A: Operand, B: Operand, num_moduli: int, backend="INT8"):
    """Fast-mode shifts (scaling_fast_real.hpp:6-49) with a sequential round-up sum of squares."""
    L = lib()
    log2P = np.float32(T.log2P(backend, num_moduli))
    f32 = A.view.dtype in (np.float32, np.complex64)

    def one(O: Operand):
        parts = _real_parts(O.view)
        out = np.zeros(O.rows, dtype=np.int16)
        amb = np.zeros(O.rows, dtype=bool)
        for r in range(O.rows):
            amax = max(float(np.abs(pz[r]).max()) if O.inner else 0.0 for pz in parts)
            ss = 0.0
            for pz in parts:
                row = np.ascontiguousarray(pz[r].astype(np.float64))
                ss += L.g8o_sumsq_ru_d(_p(row), _sz(O.inner), 1, _sz(0), _sz(O.inner))
            if f32:
                ss = float(np.nextafter(np.float32(ss), np.float32(np.inf))) if np.float32(ss) < ss else float(np.float32(ss))
            a = ctypes.c_int(0)
            s = L.g8o_fast_shift(amax, ss, log2P, int(f32), ctypes.byref(a))
            out[r] = np.int16(-s) if ss > 0 else 0
            amb[r] = bool(a.value)
        return out, amb

    sftA, ambA = one(A)
    sftB, ambB = one(B)
    return sftA, sftB, ambA, ambB


def split(O: Operand, sft: np.ndarray, num_moduli: int, backend="INT8"):
    """int8 residue planes [component][num_moduli, rows, k_pad] of trunc(op(X) * 2^-sft)."""
    L = lib()
    k_pad = pad256(O.inner)
    mod = np.array(T.moduli(backend)[:num_moduli], dtype=np.int32)
    sft = np.ascontiguousarray(sft, dtype=np.int16)
    out = []
    if backend == "FP8":  # int16 residues (the device stores each as 2-3 e4m3 pieces, see decode_fp8_planes)
        for pz in _real_parts(O.view):
            planes = np.zeros((num_moduli, O.rows, k_pad), dtype=np.int16)
            fn = L.g8o_split16_f if pz.dtype == np.float32 else L.g8o_split16_d
            fn(_p(pz), _sz(O.inner), 1, _sz(O.rows), _sz(O.inner), _sz(k_pad), _p(sft), _p(mod), int(num_moduli),
               _p(planes), _sz(O.rows * k_pad))
            out.append(planes)
        if len(out) == 2:  # complex: third plane set (Re + Im) mod p, symmetric
            ssum = out[0].astype(np.int32) + out[1].astype(np.int32)
            ri = np.empty_like(out[0])
            for i, p in enumerate(mod):
                h = p // 2
                ri[i] = np.where(ssum[i] > h, ssum[i] - p, np.where(ssum[i] < -h, ssum[i] + p, ssum[i])).astype(np.int16)
            out.append(ri)
        return out
    for pz in _real_parts(O.view):
        planes = np.zeros((num_moduli, O.rows, k_pad), dtype=np.int8)
        fn = L.g8o_split_f if pz.dtype == np.float32 else L.g8o_split_d
        fn(_p(pz), _sz(O.inner), 1, _sz(O.rows), _sz(O.inner), _sz(k_pad), _p(sft), _p(mod), int(num_moduli),
           _p(planes), _sz(O.rows * k_pad))
        out.append(planes)
    if len(out) == 2:
        # third plane set: (Re + Im) mod p, symmetric (mod.hpp:315-355)
        s = out[0].astype(np.int32) + out[1].astype(np.int32)
        ri = np.empty_like(out[0])
        for i, p in enumerate(mod):
            w = np.where(s[i] > p // 2, s[i] - p, np.where(s[i] < -(p // 2), s[i] + p, s[i]))
            ri[i] = w.astype(np.int8)
        out.append(ri)
    return out


def gemm_mod(A_lo, B_lo, m: int, n: int, num_moduli: int, backend="INT8"):
    """C_mid planes: real -> int8 [N, n, m_pad] (column-major m_pad x n); complex -> int8 [N, n, m_pad, 2]."""
    L = lib()
    mod = np.array(T.moduli(backend)[:num_moduli], dtype=np.int32)
    m_pad = pad256(m)
    k_pad = A_lo[0].shape[2]
    if backend == "FP8" and len(A_lo) > 1:
        C = np.zeros((num_moduli, n, m_pad, 2), dtype=np.int16)
        L.g8o_gemm_mod_i16_cplx(_p(A_lo[0]), _p(A_lo[1]), _sz(A_lo[0].shape[1] * k_pad), _p(B_lo[0]), _p(B_lo[1]),
                                _sz(B_lo[0].shape[1] * k_pad), _sz(m), _sz(n), _sz(k_pad), _p(mod), int(num_moduli),
                                _p(C), _sz(m_pad), _sz(m_pad * n))
        return C
    if backend == "FP8":
        C = np.zeros((num_moduli, n, m_pad), dtype=np.int16)
        L.g8o_gemm_mod_i16(_p(A_lo[0]), _sz(A_lo[0].shape[1] * k_pad), _p(B_lo[0]), _sz(B_lo[0].shape[1] * k_pad),
                           _sz(m), _sz(n), _sz(k_pad), _p(mod), int(num_moduli), _p(C), _sz(m_pad), _sz(m_pad * n))
        return C
    if len(A_lo) == 1:
        C = np.zeros((num_moduli, n, m_pad), dtype=np.int8)
        L.g8o_gemm_mod_i8(_p(A_lo[0]), _sz(A_lo[0].shape[1] * k_pad), _p(B_lo[0]), _sz(B_lo[0].shape[1] * k_pad),
                          _sz(m), _sz(n), _sz(k_pad), _p(mod), int(num_moduli), _p(C), _sz(m_pad), _sz(m_pad * n))
        return C
    C = np.zeros((num_moduli, n, m_pad, 2), dtype=np.int8)
    L.g8o_gemm_mod_i8_cplx(_p(A_lo[0]), _p(A_lo[1]), _sz(A_lo[0].shape[1] * k_pad), _p(B_lo[0]), _p(B_lo[1]),
                           _sz(B_lo[0].shape[1] * k_pad), _sz(m), _sz(n), _sz(k_pad), _p(mod), int(num_moduli),
                           _p(C), _sz(m_pad), _sz(m_pad * n))
    return C


def _mode(alpha, beta):
    if alpha == 1 and beta == 0:
        return 0
    if alpha == 1 and beta == 1:
        return 1
    if alpha == -1 and beta == 0:
        return 2
    if alpha == -1 and beta == 1:
        return 3
    return 4


def crt(C_mid, m, n, num_moduli, sftA, sftB, dtype, alpha=1.0, beta=0.0, C0=None, backend="INT8", device_scalars=False):
    """CRT accumulate + unscale + alpha/beta (inverse_scaling_real.hpp / inverse_scaling_complex.hpp).
    Returns C as an (m, n) Fortran-ordered array of `dtype`. device_scalars forces the general path
    (the reference uses fma(beta,C,alpha*AB) whenever alpha is a device pointer)."""
    L = lib()
    dtype = np.dtype(dtype)
    cplx = dtype.kind == "c"
    f32 = dtype in (np.float32, np.complex64)
    pd = T.THRESHOLD[backend]["P_is_double"]
    use_dd = int((not f32) and num_moduli > pd)
    q1 = np.array(T.qPi_1(backend, num_moduli), dtype=np.float64)
    q2 = np.array(T.qPi_2(backend, num_moduli), dtype=np.float64).reshape(-1) if num_moduli > pd else np.zeros(2)
    P = np.array(T.P_dd(backend, num_moduli), dtype=np.float64)
    invP = float(T.invP(backend, num_moduli))
    m_pad = C_mid.shape[2]
    C = np.zeros((m, n), dtype=dtype, order="F") if C0 is None else np.array(C0, dtype=dtype, order="F", copy=True)
    sftA = np.ascontiguousarray(sftA, dtype=np.int16)
    sftB = np.ascontiguousarray(sftB, dtype=np.int16)
    mode = 4 if device_scalars else _mode(alpha, beta)
    ldc = C.strides[1] // C.itemsize
    if backend == "FP8" and cplx:
        a = np.array([complex(alpha).real, complex(alpha).imag], dtype=np.float32 if f32 else np.float64)
        b = np.array([complex(beta).real, complex(beta).imag], dtype=np.float32 if f32 else np.float64)
        if f32:
            L.g8o_crt16_c(_p(C_mid), _sz(m_pad), _sz(m_pad * n), _sz(m), _sz(n), int(num_moduli), _p(q1), _p(P),
                          ctypes.c_double(invP), _p(sftA), _p(sftB), mode, _p(a), _p(b), _p(C), _sz(ldc))
        else:
            L.g8o_crt16_z(_p(C_mid), _sz(m_pad), _sz(m_pad * n), _sz(m), _sz(n), int(num_moduli), use_dd, _p(q1), _p(q2),
                          _p(P), ctypes.c_double(invP), _p(sftA), _p(sftB), mode, _p(a), _p(b), _p(C), _sz(ldc))
        return C
    if backend == "FP8":
        if f32:
            L.g8o_crt16_f(_p(C_mid), _sz(m_pad), _sz(m_pad * n), _sz(m), _sz(n), int(num_moduli), _p(q1), _p(P),
                          ctypes.c_double(invP), _p(sftA), _p(sftB), mode, ctypes.c_float(alpha), ctypes.c_float(beta),
                          _p(C), _sz(ldc))
        else:
            L.g8o_crt16_d(_p(C_mid), _sz(m_pad), _sz(m_pad * n), _sz(m), _sz(n), int(num_moduli), use_dd, _p(q1), _p(q2),
                          _p(P), ctypes.c_double(invP), _p(sftA), _p(sftB), mode, ctypes.c_double(alpha),
                          ctypes.c_double(beta), _p(C), _sz(ldc))
        return C
    if not cplx:
        if f32:
            L.g8o_crt_f(_p(C_mid), _sz(m_pad), _sz(m_pad * n), _sz(m), _sz(n), int(num_moduli), _p(q1), _p(P),
                        ctypes.c_double(invP), _p(sftA), _p(sftB), mode, ctypes.c_float(alpha), ctypes.c_float(beta),
                        _p(C), _sz(ldc))
        else:
            L.g8o_crt_d(_p(C_mid), _sz(m_pad), _sz(m_pad * n), _sz(m), _sz(n), int(num_moduli), use_dd, _p(q1), _p(q2),
                        _p(P), ctypes.c_double(invP), _p(sftA), _p(sftB), mode, ctypes.c_double(alpha),
                        ctypes.c_double(beta), _p(C), _sz(ldc))
    else:
        a = np.array([complex(alpha).real, complex(alpha).imag], dtype=np.float32 if f32 else np.float64)
        b = np.array([complex(beta).real, complex(beta).imag], dtype=np.float32 if f32 else np.float64)
        if f32:
            L.g8o_crt_c(_p(C_mid), _sz(m_pad), _sz(m_pad * n), _sz(m), _sz(n), int(num_moduli), _p(q1), _p(P),
                        ctypes.c_double(invP), _p(sftA), _p(sftB), mode, _p(a), _p(b), _p(C), _sz(ldc))
        else:
            L.g8o_crt_z(_p(C_mid), _sz(m_pad), _sz(m_pad * n), _sz(m), _sz(n), int(num_moduli), use_dd, _p(q1), _p(q2),
                        _p(P), ctypes.c_double(invP), _p(sftA), _p(sftB), mode, _p(a), _p(b), _p(C), _sz(ldc))
    return C


def emulate(A, B, op_A="N", op_B="N", num_moduli=14, fastmode=False, alpha=1.0, beta=0.0, C0=None,
            sftA=None, sftB=None, backend="INT8", device_scalars=False):
    """Full restated pipeline.  A, B: 2-D numpy arrays (any order; treated as the BLAS matrices).
    Returns dict(C, sftA, sftB, ambA, ambB, A_lo, B_lo, C_mid)."""
    oa, ob = Operand(A, op_A, True), Operand(B, op_B, False)
    assert oa.inner == ob.inner
    m, n = oa.rows, ob.rows
    ambA = ambB = None
    if sftA is None or sftB is None:
        if backend == "FP8" and not fastmode:
            sA, sB, ambA, ambB = accurate_shifts_fp8(oa, ob, num_moduli)
        else:
            f = fast_shifts if fastmode else accurate_shifts
            sA, sB, ambA, ambB = f(oa, ob, num_moduli, backend)
        sftA = sA if sftA is None else sftA
        sftB = sB if sftB is None else sftB
    A_lo = split(oa, sftA, num_moduli, backend)
    B_lo = split(ob, sftB, num_moduli, backend)
    C_mid = gemm_mod(A_lo, B_lo, m, n, num_moduli, backend)
    dtype = np.result_type(np.asarray(A).dtype, np.asarray(B).dtype)
    C = crt(C_mid, m, n, num_moduli, sftA, sftB, dtype, alpha, beta, C0, backend, device_scalars)
    return dict(C=C, sftA=np.asarray(sftA, dtype=np.int16), sftB=np.asarray(sftB, dtype=np.int16), ambA=ambA,
                ambB=ambB, A_lo=A_lo, B_lo=B_lo, C_mid=C_mid)
