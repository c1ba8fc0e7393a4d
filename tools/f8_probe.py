#!/usr/bin/env python3
"""Is tcgen05.mma.kind::f8f6f4 (e4m3 x e4m3 -> f32 in TMEM) EXACT for small-integer operands while |sum| <= 2^24?
The reference's FP8 backend relies on this property of the library GEMM (mod.hpp:159-189); SURVEY flags it as unverified."""
import ctypes, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gemmul8_b200 import _lib, api
lib = _lib.load()
rng = np.random.default_rng(0)

def run(A, B):
    """A: (m, k_pad) ints, B: (n, k_pad) ints -> float32 (n, m_pad) raw accumulators"""
    m, k_pad = A.shape; n = B.shape[0]; m_pad = api.pad256(m)
    dA = torch.from_numpy(A.astype(np.float32)).cuda().to(torch.float8_e4m3fn).view(torch.int8).contiguous()
    dB = torch.from_numpy(B.astype(np.float32)).cuda().to(torch.float8_e4m3fn).view(torch.int8).contiguous()
    out = torch.zeros((n, m_pad), dtype=torch.float32, device="cuda")
    code = lib.g8_stage_gemm(7, 0, dA.data_ptr(), m * k_pad, dB.data_ptr(), n * k_pad, m, n, k_pad, 1, 0, None, None, out.data_ptr(), m_pad * n, m_pad,
                             None, None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize(); assert code == 0, code
    return out.cpu().numpy()[:, :m]

bad = 0
for k_pad in (256, 4096, 32768, 65536):
    m, n = 256, 128
    A = rng.integers(-16, 17, size=(m, k_pad)); B = rng.integers(-16, 17, size=(n, k_pad))
    # adversarial rows: huge running sum followed by +-1 products
    A[0, :] = 16; B[0, :] = 16; A[0, -64:] = 1
    A[1, :] = 16; B[1, :] = 16; A[1, ::2] = -16; A[1, -3:] = [1, 3, 5]
    A[2, :] = 15; B[2, :] = 15
    got = run(A, B)
    ref = (B.astype(np.int64) @ A.astype(np.int64).T)
    ok = np.array_equal(got.astype(np.int64), ref) and np.all(np.abs(ref) <= 2 ** 24)
    nbad = int(np.sum(got.astype(np.int64) != ref))
    print(f"k={k_pad}: exact={ok} mismatches={nbad} max|sum|={np.abs(ref).max()} (2^24={2**24})", flush=True)
    if not ok:
        i = np.argwhere(got.astype(np.int64) != ref)[:5]
        for c, r in i: print("   ", c, r, got[c, r], ref[c, r])
        bad += 1
print("F8 EXACT" if not bad else "F8 NOT EXACT")
