"""Skip-scaling API (workA / workB, enable_skip_scal*, skip_scal*; reference gemmul8_real.hpp:82-83,101-104,123-136 and the
hook's use of it, hook.cu:688-727) against the UNMODIFIED reference library, bit for bit.

Protocol per case: call 1 with separate workA / workB buffers and enable_skip_scal{A,B} = 1 (planes, shifts and -- accurate mode --
the extra bound plane are cached); call 2 with a NEW B, skip_scalA = 1 (A's cached planes re-used, incl. accurate mode's cached
bound plane of A, which the new B's shift depends on); call 3 with skip_scalA = skip_scalB = 1 and different alpha / beta.  The C of
every call must equal the reference's, and workA's cached planes must equal the reference's cached planes."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _ref():
    so = ROOT / "oracle/_ref/libgemmul8_ref.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libgemmul8_ref.so not built (needs /root/reference at build time)")
    sys.path.insert(0, str(ROOT))
    from bench import RefLib
    return RefLib().L


CASES = [(np.float64, 14, 0, (300, 200, 1000), "N", "N"), (np.float64, 14, 0, (257, 129, 515), "T", "T"), (np.float32, 6, 0, (130, 90, 400), "N", "T"),
         (np.complex128, 10, 0, (130, 90, 400), "C", "N"), (np.float64, 12, 1, (200, 100, 600), "N", "N")]


@pytest.mark.parametrize("fast", [False, True], ids=["accu", "fast"])
@pytest.mark.parametrize("case", CASES, ids=[f"{np.dtype(c[0]).name}-N{c[1]}-be{c[2]}-{c[4]}{c[5]}" for c in CASES])
def test_skip_scaling_matches_reference(cuda, case, fast):
    import torch
    import helpers as H
    import gemmul8_b200 as g8
    from gemmul8_b200 import api

    R = _ref()
    dtype, N, be, (m, n, k), opA, opB = case
    cplx = np.dtype(dtype).kind == "c"
    tdt = H.NP2T[np.dtype(dtype)]
    rng = np.random.default_rng(99)
    A = H.rand_matrix(rng, H.stored_shape(opA, m, k), dtype, phi=1.0)
    B1 = H.rand_matrix(rng, H.stored_shape(opB, k, n), dtype, phi=1.0)
    B2 = H.rand_matrix(rng, H.stored_shape(opB, k, n), dtype, phi=2.0)
    C0 = H.rand_matrix(rng, (m, n), dtype)
    dA, lda = H.to_dev_colmajor(A)
    dB1, ldb = H.to_dev_colmajor(B1)
    dB2, _ = H.to_dev_colmajor(B2)

    tot, wa, wb = g8.work_size(m, n, k, N, is_complex=cplx, backend=be, enable_skip_scalA=True, enable_skip_scalB=True)
    ra, rb = ctypes.c_size_t(0), ctypes.c_size_t(0)
    rtot = R.ref_work_size(int(cplx), be, m, n, k, N, 1, 1, ctypes.byref(ra), ctypes.byref(rb))
    assert (tot, wa, wb) == (rtot, ra.value, rb.value)

    def buffers():
        # work holds only the C part when workA / workB are given (gemmul8_real.hpp:101-105); poison everything
        return (torch.full((tot,), 0x5A, dtype=torch.uint8, device="cuda"), torch.full((wa,), 0x5A, dtype=torch.uint8, device="cuda"),
                torch.full((wb,), 0x5A, dtype=torch.uint8, device="cuda"))

    ours, refb = buffers(), buffers()
    keep = []
    st = torch.cuda.current_stream().cuda_stream
    plan = [  # (B, alpha, beta, skipA, skipB)
        (dB1, 1.0, 0.0, 0, 0),
        (dB2, 1.0, 0.0, 1, 0),
        (dB2, -0.75, 0.5, 1, 1),
        (dB1, 1.0, 1.0, 1, 0),
    ]
    for step, (dB, alpha, beta, skA, skB) in enumerate(plan):
        Cs = []
        for impl in ("ours", "ref"):
            dC, ldc = H.to_dev_colmajor(C0)
            w, wA, wB = ours if impl == "ours" else refb
            if impl == "ours":
                g8.gemm(opA, opB, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc, N, fast, w, workA=wA, workB=wB, enable_skip_scalA=True,
                        enable_skip_scalB=True, skip_scalA=bool(skA), skip_scalB=bool(skB), backend=be)
            else:
                pa, pb = api._scalar_ptr(alpha, tdt, keep), api._scalar_ptr(beta, tdt, keep)
                code = R.ref_gemm(api._DTYPES[tdt], be, 1, api._op(opA), api._op(opB), m, n, k, pa, dA.data_ptr(), lda, dB.data_ptr(), ldb, pb,
                                  dC.data_ptr(), ldc, N, int(fast), w.data_ptr(), wA.data_ptr(), wB.data_ptr(), 1, 1, skA, skB, ctypes.c_void_p(st), None)
                assert code == 0
            torch.cuda.synchronize()
            Cs.append(H.from_dev_colmajor(dC, m, n, ldc))
        assert H.bits_equal(Cs[0], Cs[1]), f"step {step}: C differs from the reference ({H.first_diff(Cs[0], Cs[1], 'C')})"
    # cached A side: the shift vector and (INT8) every plane row < m must match the reference's cache byte for byte
    k_pad, m_pad = api.pad256(k), api.pad256(m)
    G = 3 if cplx else 1
    from gemmul8_b200 import tables as T
    nm = T.num_mat("INT8" if be == 0 else "FP8", N)
    oa = api.aligned_view(ours[1]).cpu().numpy()
    rf = api.aligned_view(refb[1]).cpu().numpy()
    planes = (nm + 1) * G
    sft_off = k_pad * m_pad * planes
    assert np.array_equal(oa[sft_off:sft_off + 2 * m], rf[sft_off:sft_off + 2 * m]), "cached sftA differs"
    if be == 0:
        po = oa[:k_pad * m_pad * nm * G].reshape(nm * G, m_pad, k_pad)[:, :m, :k]
        pr = rf[:k_pad * m_pad * nm * G].reshape(nm * G, m_pad, k_pad)[:, :m, :k]
        assert np.array_equal(po, pr), "cached A planes differ"
