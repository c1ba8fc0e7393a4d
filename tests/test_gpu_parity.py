"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs.  Integer / byte results (shift exponents modulo the MUFU.LG2 caveat, residue planes, C_mid) and the final
floating-point C must match BIT FOR BIT -- the pipeline is exact-integer until the ordered FMA chain of the CRT."""
import ctypes
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def env(cuda):
    import torch
    import helpers as H
    from gemmul8_b200 import _lib, api, tables as T
    from oracle import oracle as O

    class E:
        pass

    e = E()
    e.torch, e.H, e.lib, e.api, e.T, e.O = torch, H, _lib.load(), api, T, O
    assert e.lib.g8_device_supported(0) == 1, "tests need an sm_100a device"
    return e


def _oracle_check(e, A, B, opA, opB, N, fast, **kw):
    H, O = e.H, e.O
    C, W = H.run_gemm(A, B, opA, opB, N, fast, return_work=True, **kw)
    m = C.shape[0]
    r = O.emulate(A, B, opA, opB, N, fast, sftA=W["sftA"], sftB=W["sftB"], alpha=kw.get("alpha", 1.0), beta=kw.get("beta", 0.0),
                  C0=kw.get("C0"), device_scalars=kw.get("device_scalars", False))
    for g in range(len(r["A_lo"])):
        assert np.array_equal(W["A_lo"][g], r["A_lo"][g]), f"A_lo group {g}"
        assert np.array_equal(W["B_lo"][g], r["B_lo"][g]), f"B_lo group {g}"
    assert np.array_equal(W["C_mid"][:, :, :m], r["C_mid"][:, :, :m]), "C_mid"
    assert H.bits_equal(C, r["C"]), H.first_diff(C, r["C"], "C")
    # the device's own shifts agree with the CPU formula except on floor() boundaries of the approximate log2
    r2 = O.emulate(A, B, opA, opB, N, fast)
    assert not np.any((r2["sftA"] != W["sftA"]) & ~r2["ambA"])
    assert not np.any((r2["sftB"] != W["sftB"]) & ~r2["ambB"])
    return C, W


def test_sample_known_answer(env):
    """reference sample/dgemm_cuBLASLt_int8.cu:26-40: N=15 accurate mode returns the exact product"""
    g = json.loads((ROOT / "tests/golden/sample_kat.json").read_text())
    A = np.array([float.fromhex(x) for x in g["A"]]).reshape(5, 4).T
    B = np.array([float.fromhex(x) for x in g["B"]]).reshape(3, 5).T
    Cx = np.array([float.fromhex(x) for x in g["C_exact"]]).reshape(3, 4).T
    C = env.H.run_gemm(A, B, num_moduli=15, fastmode=False)
    assert env.H.bits_equal(C, Cx)


@pytest.mark.parametrize("dtype,N", [(np.float64, 14), (np.float64, 7), (np.float64, 18), (np.float32, 6), (np.float32, 13),
                                     (np.complex128, 18), (np.complex128, 9), (np.complex64, 6)])
@pytest.mark.parametrize("fast", [False, True])
def test_end_to_end_vs_oracle(env, dtype, N, fast):
    rng = np.random.default_rng(hash((str(dtype), N, fast)) % 2 ** 32)
    for opA, opB in (("N", "N"), ("T", "N"), ("N", "T"), ("C", "C")):
        m, n, k = 150, 70, 333
        A = env.H.rand_matrix(rng, env.H.stored_shape(opA, m, k), dtype)
        B = env.H.rand_matrix(rng, env.H.stored_shape(opB, k, n), dtype)
        _oracle_check(env, A, B, opA, opB, N, fast)


@pytest.mark.parametrize("shape", [(1, 1, 1), (4, 3, 5), (32, 32, 32), (47, 33, 45), (257, 129, 255), (128, 128, 1024), (300, 1, 700), (1, 300, 513)])
def test_edge_shapes(env, shape):
    """debug/test.cu sweeps m=n=k in [32,47]; plus degenerate and ragged (non multiple of 256 / 128) shapes"""
    m, n, k = shape
    rng = np.random.default_rng(m * 1000003 + n * 1009 + k)
    for dtype, N in ((np.float64, 12), (np.complex128, 16)):
        A = env.H.rand_matrix(rng, (m, k), dtype)
        B = env.H.rand_matrix(rng, (k, n), dtype)
        for fast in (False, True):
            _oracle_check(env, A, B, "N", "N", N, fast)


def test_empty_dimensions(env):
    """BLAS quick-return semantics: k = 0 leaves C = beta * C (the product is empty), m = 0 or n = 0 touches nothing; status 0."""
    torch = env.torch
    import gemmul8_b200 as g8

    m, n = 70, 50
    dummy = torch.ones(64, dtype=torch.float64, device="cuda")
    C0 = torch.randn(m * n, dtype=torch.float64, device="cuda")
    work = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
    for fast in (False, True):
        for beta in (0.0, 0.5):
            C = C0.clone()
            g8.gemm("N", "N", m, n, 0, 1.0, dummy, m, dummy, 1, beta, C, m, 14, fast, work)
            torch.cuda.synchronize()
            assert torch.equal(C, beta * C0), (fast, beta)
        for mm, nn in ((0, n), (m, 0)):
            C = C0.clone()
            g8.gemm("N", "N", mm, nn, 33, 1.0, dummy, max(mm, 1), dummy, 33, 0.0, C, max(mm, 1), 14, fast, work)
            torch.cuda.synchronize()
            assert torch.equal(C, C0), (fast, mm, nn)


@pytest.mark.parametrize("alpha,beta,dev", [(1, 1, False), (-1, 0, False), (-1, 1, False), (0.75, -1.5, False), (0.75, -1.5, True), (1, 0, True), (0, 1, False)])
def test_alpha_beta_and_leading_dimensions(env, alpha, beta, dev):
    """the five (alpha,beta) classes of debug/test.cu:106-141, host and device scalars, lda/ldb/ldc > rows"""
    rng = np.random.default_rng(11)
    for dtype, N in ((np.float64, 14), (np.float32, 8), (np.complex128, 15)):
        m, n, k = 65, 40, 130
        A = env.H.rand_matrix(rng, (m, k), dtype)
        B = env.H.rand_matrix(rng, (k, n), dtype)
        C0 = env.H.rand_matrix(rng, (m, n), dtype)
        a, b = (alpha, beta)
        if np.dtype(dtype).kind == "c" and alpha == 0.75:
            a, b = 0.75 - 0.5j, -1.5 + 0.25j
        _oracle_check(env, A, B, "N", "N", N, False, alpha=a, beta=b, C0=C0, device_scalars=dev, lda=m + 3, ldb=k + 5, ldc=m + 7)


def test_zero_rows_columns_and_tiny_values(env):
    """empty rows/columns and badly scaled rows/columns (row 9 tiny, column 3 huge)"""
    rng = np.random.default_rng(3)
    m, n, k = 40, 30, 100
    A0 = env.H.rand_matrix(rng, (m, k), np.float64)
    B0 = env.H.rand_matrix(rng, (k, n), np.float64)
    A0[5, :] = 0.0
    B0[:, 7] = 0.0
    for fast in (False, True):
        A, B = A0.copy(), B0.copy()
        # accurate mode handles the whole binary64 range.  The reference's fast-mode shift (scaling_fast_real.hpp:6-14)
        # subtracts BOTH log2 of the row norm and ilogb(amax) and converts amax to float, so it loses 2^ilogb(amax) of
        # precision on badly scaled rows; we reproduce it bit for bit, hence no scaling in the fast-mode accuracy check.
        if not fast:
            A[9, :] *= 1e-200
            B[:, 3] *= 1e150
        C, W = env.H.run_gemm(A, B, "N", "N", 14, fast, return_work=True)
        assert np.all(C[5, :] == 0) and np.all(C[:, 7] == 0)
        ref = A @ B
        bound = np.abs(A) @ np.abs(B)  # componentwise error scale of a GEMM
        assert np.all(np.abs(C - ref) <= (1e-10 if fast else 1e-12) * bound)
        r = env.O.emulate(A, B, "N", "N", 14, fast, sftA=W["sftA"], sftB=W["sftB"])
        assert env.H.bits_equal(C, r["C"])


def test_tensor_core_gemm_vs_dp4a_and_numpy(env):
    """stage 2 alone: tcgen05 kernel == dp4a cross-check kernel == numpy int64, all epilogues, multi-tile shapes"""
    import sys
    sys.path.insert(0, str(ROOT / "tools"))
    import gpu_debug as D

    D.FAILS.clear()
    D.check_gemm(False, [(300, 200, 1000, 5), (700, 515, 512, 3)])
    assert not D.FAILS, D.FAILS


def test_crt_stage_all_modes(env):
    import sys
    sys.path.insert(0, str(ROOT / "tools"))
    import gpu_debug as D

    D.FAILS.clear()
    D.check_crt()
    assert not D.FAILS, D.FAILS


def test_fp8_backend_vs_oracle(env):
    """FP8 backend (gemmLt<T, FP8>, real AND complex types): e4m3 piece planes decode to the oracle's residues, C_mid and C are
    bit-exact given the device's shifts (tools/gpu_debug.py:check_fp8)."""
    import sys
    sys.path.insert(0, str(ROOT / "tools"))
    import gpu_debug as D

    D.FAILS.clear()
    D.check_fp8()
    assert not D.FAILS, D.FAILS


def test_full_size_properties(env):
    """BASELINE size (8192^3, N=14): size-independent checks -- linearity in alpha/beta is exact, permuting columns of B
    permutes columns of C bit for bit (column-wise shifts), and a 256x256 corner matches float64 numpy to emulated precision."""
    torch, g8 = env.torch, __import__("gemmul8_b200")
    S, N = 8192, 14
    A = g8.randmat(S, S, torch.float64, seed=12345)
    B = g8.randmat(S, S, torch.float64, seed=54321)
    C1 = torch.zeros(S * S, dtype=torch.float64, device="cuda")
    C2 = torch.zeros_like(C1)
    tot, _, _ = g8.work_size(S, S, S, N)
    work = torch.empty(tot, dtype=torch.uint8, device="cuda")
    for fast in (False, True):
        g8.gemm("N", "N", S, S, S, 1.0, A, S, B, S, 0.0, C1, S, N, fast, work)
        # run-to-run bit reproducibility
        g8.gemm("N", "N", S, S, S, 1.0, A, S, B, S, 0.0, C2, S, N, fast, work)
        assert torch.equal(C1, C2)
        # C2 = -AB + C1 must be exactly zero wherever finite
        g8.gemm("N", "N", S, S, S, -1.0, A, S, B, S, 1.0, C2, S, N, fast, work)
        assert float(C2.abs().max()) == 0.0
        # corner block against float64 matmul
        Am = A.view(S, S).t()[:256, :]   # rows 0..255 of the column-major matrix
        Bm = B.view(S, S).t()[:, :256]
        ref = (Am @ Bm)
        got = C1.view(S, S).t()[:256, :256]
        rel = float((got - ref).abs().max() / ref.abs().max())
        assert rel < (5e-11 if fast else 5e-13), rel
        # column permutation property
        perm = torch.randperm(S, device="cuda")
        Bp = B.view(S, S)[perm].contiguous().view(-1)   # columns of the column-major B permuted
        g8.gemm("N", "N", S, S, S, 1.0, A, S, Bp, S, 0.0, C2, S, N, fast, work)
        # shifts are row-/column-local (accurate mode: row max over ALL columns, which a permutation keeps) -> bitwise equal
        assert torch.equal(C2.view(S, S), C1.view(S, S)[perm])


@pytest.mark.parametrize("dtype,N", [(np.float64, 14), (np.complex128, 10), (np.float32, 7)])
@pytest.mark.parametrize("fast", [False, True])
def test_host_pipeline_bitwise(env, dtype, N, fast):
    """gemm_host (pinned host buffers, column-chunked, copies overlapped on three streams) == monolithic g8_gemm, bit for bit"""
    torch, H = env.torch, env.H
    import gemmul8_b200 as g8

    rng = np.random.default_rng(17)
    m, n, k = 190, 700, 300
    for opA, opB, beta, ldc in (("N", "N", 0.0, m), ("T", "N", 0.5, m + 5), ("N", "T", 0.0, m)):
        A = H.rand_matrix(rng, H.stored_shape(opA, m, k), dtype)
        B = H.rand_matrix(rng, H.stored_shape(opB, k, n), dtype)
        C0 = H.rand_matrix(rng, (m, n), dtype)
        want = H.run_gemm(A, B, opA, opB, N, fast, alpha=1.0, beta=beta, C0=C0, ldc=ldc)
        dA, lda = H.to_dev_colmajor(A)
        dB, ldb = H.to_dev_colmajor(B)
        dC, _ = H.to_dev_colmajor(C0, ldc)
        hA, hB, hC = dA.cpu().pin_memory(), dB.cpu().pin_memory(), dC.cpu().pin_memory()
        sentinel = hC.clone()
        for native in (True, False):
            hC.copy_(sentinel)
            g8.gemm_host(opA, opB, m, n, k, 1.0, hA, lda, hB, ldb, beta, hC, ldc, num_moduli=N, fastmode=fast, chunk=256, native=native)
            got = hC.numpy().reshape(n, ldc)[:, :m].T.copy()
            assert H.bits_equal(got, want), H.first_diff(got, want, f"C {opA}{opB} native={native}")
            if ldc > m:  # padding rows of the caller's buffer must be untouched
                assert np.array_equal(hC.numpy().reshape(n, ldc)[:, m:], sentinel.numpy().reshape(n, ldc)[:, m:])


@pytest.mark.parametrize("dtype,N,be", [(np.float64, 14, 0), (np.float32, 6, 0), (np.complex128, 11, 0), (np.complex64, 6, 0), (np.float64, 12, 1),
                                         (np.complex128, 9, 1)])
@pytest.mark.parametrize("fast", [False, True])
def test_native_host_pipeline_all_types_bitwise(env, dtype, N, be, fast):
    """g8_gemm_host (C ABI, csrc/g8_host.cu): every type, both backends, N/T/C operands, leading dimensions beyond the matrix, beta != 0,
    a chunk size that leaves a ragged last chunk, pageable AND pinned host memory, and plan re-use -- all bit-identical to g8_gemm"""
    torch, H = env.torch, env.H
    import gemmul8_b200 as g8

    rng = np.random.default_rng(23)
    m, n, k = 150, 600, 1100   # k_pad / 128 = 9 tiles: the cluster-fused accurate stage (i) runs with a ragged split of the tiles
    cplx = np.dtype(dtype).kind == "c"
    for opA, opB, alpha, beta, pad, pinned in (("N", "N", 1.0, 0.0, 0, True), ("T", "C" if cplx else "T", 0.75, -0.5, 3, True), ("C" if cplx else "T", "N", 1.0, 1.0, 0, False)):
        A = H.rand_matrix(rng, H.stored_shape(opA, m, k), dtype)
        B = H.rand_matrix(rng, H.stored_shape(opB, k, n), dtype)
        C0 = H.rand_matrix(rng, (m, n), dtype)
        lda, ldb, ldc = A.shape[0] + pad, B.shape[0] + pad, m + pad
        want = H.run_gemm(A, B, opA, opB, N, fast, alpha=alpha, beta=beta, C0=C0, lda=lda, ldb=ldb, ldc=ldc, backend=be)
        dA, _ = H.to_dev_colmajor(A, lda)
        dB, _ = H.to_dev_colmajor(B, ldb)
        dC, _ = H.to_dev_colmajor(C0, ldc)
        hA, hB, hC0 = dA.cpu(), dB.cpu(), dC.cpu()
        if pinned:
            hA, hB, hC0 = hA.pin_memory(), hB.pin_memory(), hC0.pin_memory()
        plan = g8.NativeHostGemm(m, n, k, H.NP2T[np.dtype(dtype)], N, fast, opA, opB, chunk=256, backend=be)
        for rep in range(2):   # the second call re-uses the plan's buffers, streams and events
            hC = hC0.clone()
            if pinned:
                hC = hC.pin_memory()
            plan.run(hA, hB, hC, alpha, beta, lda, ldb, ldc)
            torch.cuda.synchronize()
            got = hC.numpy().reshape(n, ldc)[:, :m].T.copy()
            assert H.bits_equal(got, want), H.first_diff(got, want, f"C {opA}{opB} rep={rep}")
            if pad:
                assert np.array_equal(hC.numpy().reshape(n, ldc)[:, m:], hC0.numpy().reshape(n, ldc)[:, m:])
        plan.close()


@pytest.mark.parametrize("nparts", [2, 8])
@pytest.mark.parametrize("dtype,N", [(np.float64, 14), (np.float32, 6)])
def test_kshard_stage_kernels_on_one_gpu(env, nparts, dtype, N):
    """K-sharded helpers: g8_stage_crt_parts (CRT that sums per-shard residues) == g8_stage_residue_sum + g8_stage_crt, bit for
    bit, and both == the oracle's CRT on the numpy sum; the fused GEMM -> scatter with every 'peer' mapped to buffers of THIS GPU
    == the plain GEMM epilogue."""
    torch, H, lib, api, T, O = env.torch, env.H, env.lib, env.api, env.T, env.O
    rng = np.random.default_rng(nparts * 100 + N)
    m, n = 300, 40
    mp = api.pad256(m)
    mods = T.moduli("INT8")[:N]
    parts = rng.integers(-127, 128, size=(nparts, N, n, mp)).astype(np.int8)
    for i, p in enumerate(mods):
        parts[:, i] = np.clip(parts[:, i], -(p // 2), p // 2)
    sA = rng.integers(-60, -20, size=m).astype(np.int16)
    sB = rng.integers(-60, -20, size=n).astype(np.int16)
    tot = parts.astype(np.int64).sum(axis=0)
    cm = np.empty((N, n, mp), dtype=np.int8)
    for i, p in enumerate(mods):
        r = np.mod(tot[i], p)
        cm[i] = np.where(r > p // 2, r - p, r).astype(np.int8)
    want = O.crt(cm, m, n, N, sA, sB, dtype, 1.0, 0.0)
    tdt = H.NP2T[np.dtype(dtype)]
    dparts = torch.from_numpy(parts).cuda()
    dsA = torch.zeros(mp, dtype=torch.int16, device="cuda"); dsA[:m] = torch.from_numpy(sA).cuda()
    dsB = torch.zeros(api.pad256(n), dtype=torch.int16, device="cuda"); dsB[:n] = torch.from_numpy(sB).cuda()
    keep = []
    pa, pb = api._scalar_ptr(1.0, tdt, keep), api._scalar_ptr(0.0, tdt, keep)
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for fusedsum in (True, False):
        dC = torch.zeros(n * m, dtype=tdt, device="cuda")
        if fusedsum:
            code = lib.g8_stage_crt_parts(api._DTYPES[tdt], dparts.data_ptr(), nparts, N * n * mp, mp, n * mp, m, n, N, dC.data_ptr(), m,
                                          dsA.data_ptr(), dsB.data_ptr(), pa, pb, st)
        else:
            dmid = torch.zeros(N * n * mp, dtype=torch.int8, device="cuda")
            code = lib.g8_stage_residue_sum(dparts.data_ptr(), nparts, N * n * mp, mp, n, mp, n * mp, N, 0, dmid.data_ptr(), mp, n * mp, st)
            assert code == 0
            assert np.array_equal(dmid.cpu().numpy().reshape(N, n, mp), cm)
            code = lib.g8_stage_crt(api._DTYPES[tdt], dmid.data_ptr(), mp, n * mp, m, n, N, dC.data_ptr(), m, dsA.data_ptr(), dsB.data_ptr(), pa, pb, st)
        assert code == 0
        torch.cuda.synchronize()
        outs.append(dC.cpu().numpy().reshape(n, m).T.copy())
    assert H.bits_equal(outs[0], outs[1])
    assert H.bits_equal(outs[0], want), H.first_diff(outs[0], want, "C")


def test_gemm_scatter_local_peers(env):
    """the scatter epilogues (shared-memory staging + cp.async.bulk) with world = 2 'ranks' whose buffers both live on this GPU:
    rank r's columns land in peer_out[r] exactly as the plain epilogues would have written them."""
    import ctypes
    torch, lib, api = env.torch, env.lib, env.api
    m, n, k_pad, N = 300, 512, 512, 3
    mp = api.pad256(m)
    g = torch.Generator(device="cuda").manual_seed(5)
    A_lo = torch.randint(-127, 128, (N, mp, k_pad), dtype=torch.int8, device="cuda", generator=g)
    B_lo = torch.randint(-127, 128, (N, n, k_pad), dtype=torch.int8, device="cuda", generator=g)
    st = torch.cuda.current_stream().cuda_stream
    nc = n // 2
    for epi, tdt in ((0, torch.int8), (1, torch.int32)):
        ref = torch.zeros(N, n, mp, dtype=tdt, device="cuda")
        assert lib.g8_stage_gemm(epi, 0, A_lo.data_ptr(), mp * k_pad, B_lo.data_ptr(), n * k_pad, m, n, k_pad, N, 0, None, None, ref.data_ptr(),
                                 n * mp, mp, None, None, st) == 0
        bufs = [torch.full((N, nc, mp), 77, dtype=tdt, device="cuda") for _ in range(2)]
        tbl = (ctypes.c_void_p * 2)(bufs[0].data_ptr(), bufs[1].data_ptr())
        for rank in (0, 1):  # the rank only changes the tile rotation
            for b in bufs:
                b.fill_(77)
            assert lib.g8_stage_gemm_scatter(epi, A_lo.data_ptr(), mp * k_pad, B_lo.data_ptr(), n * k_pad, m, n, k_pad, N, 0, tbl, 2, rank, nc * mp, mp, st) == 0
            torch.cuda.synchronize()
            for o in (0, 1):
                assert torch.equal(bufs[o][:, :, :m], ref[:, o * nc:(o + 1) * nc, :m]), (epi, rank, o)


@pytest.mark.parametrize("case", ["zgemm4096_int8_n18", "dgemm8192_fp8_n14"])
def test_full_size_properties_other_configs(env, case):
    """BASELINE.json configs 4 and 5 at their full sizes (the multi-product paths: per-product GEMM units in batches + combine pass):
    run-to-run bit reproducibility, exact cancellation C - AB == 0 with alpha = -1, beta = 1, and a corner block against a wide matmul."""
    torch, g8 = env.torch, __import__("gemmul8_b200")
    if case.startswith("zgemm"):
        S, N, dt, be, tol = 4096, 18, torch.complex128, 0, 1e-12
    else:
        S, N, dt, be, tol = 8192, 14, torch.float64, 1, 1e-9
    A = g8.randmat(S, S, dt, seed=12345)
    B = g8.randmat(S, S, dt, seed=54321)
    C1 = torch.zeros(S * S, dtype=dt, device="cuda")
    C2 = torch.zeros_like(C1)
    tot, _, _ = g8.work_size(S, S, S, N, is_complex=dt.is_complex, backend=be)
    work = torch.empty(tot, dtype=torch.uint8, device="cuda")
    for fast in (False, True):
        g8.gemm("N", "N", S, S, S, 1.0, A, S, B, S, 0.0, C1, S, N, fast, work, backend=be)
        g8.gemm("N", "N", S, S, S, 1.0, A, S, B, S, 0.0, C2, S, N, fast, work, backend=be)
        assert torch.equal(C1.view(torch.uint8), C2.view(torch.uint8))
        g8.gemm("N", "N", S, S, S, -1.0, A, S, B, S, 1.0, C2, S, N, fast, work, backend=be)
        assert float(C2.abs().max()) == 0.0
        Am = A.view(S, S).t()[:192, :]
        Bm = B.view(S, S).t()[:, :192]
        ref = Am @ Bm
        got = C1.view(S, S).t()[:192, :192]
        rel = float((got - ref).abs().max() / ref.abs().max())
        assert rel < tol * (50 if fast else 1), rel


@pytest.mark.parametrize("dtype,N,fast,be", [(np.float64, 14, False, 0), (np.float32, 6, True, 0), (np.complex128, 8, False, 0), (np.float64, 9, False, 1)])
def test_cuda_graph_capture_and_replay(env, dtype, N, fast, be):
    """g8_gemm never synchronises (no host round trip between the stages), so a whole emulated GEMM can be captured into a CUDA
    graph and replayed -- the way to amortise the launch cost of small problems.  Replays on NEW input data must equal eager calls
    bit for bit (tensor maps, scalars and workspace pointers are baked into the graph; the data is not)."""
    torch, H = env.torch, env.H
    import gemmul8_b200 as g8

    rng = np.random.default_rng(23)
    m, n, k = 256, 192, 384
    tdt = H.NP2T[np.dtype(dtype)]
    cplx = np.dtype(dtype).kind == "c"
    dA = torch.zeros(m * k, dtype=tdt, device="cuda")
    dB = torch.zeros(k * n, dtype=tdt, device="cuda")
    dC = torch.zeros(m * n, dtype=tdt, device="cuda")
    tot, _, _ = g8.work_size(m, n, k, N, is_complex=cplx, backend=be)
    work = torch.zeros(tot, dtype=torch.uint8, device="cuda")

    def load(seed):
        r = np.random.default_rng(seed)
        A = H.rand_matrix(r, (m, k), dtype); B = H.rand_matrix(r, (k, n), dtype)
        dA.copy_(H.to_dev_colmajor(A)[0]); dB.copy_(H.to_dev_colmajor(B)[0])
        return A, B

    load(1)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):       # warm-up outside capture (function attributes, lazy module loading)
        g8.gemm("N", "N", m, n, k, 1.0, dA, m, dB, k, 0.0, dC, m, N, fast, work, backend=be)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g8.gemm("N", "N", m, n, k, 1.0, dA, m, dB, k, 0.0, dC, m, N, fast, work, backend=be)
    for seed in (2, 3):
        A, B = load(seed)
        dC.zero_()
        graph.replay()
        torch.cuda.synchronize()
        got = H.from_dev_colmajor(dC, m, n, m)
        want = H.run_gemm(A, B, "N", "N", N, fast, backend=be)
        assert H.bits_equal(got, want), H.first_diff(got, want, f"graph replay seed {seed}")


def test_fp8_tensor_core_accumulation_is_exact(env):
    """The FP8 backend relies on exact f32 accumulation of small-integer e4m3 products while |sum| <= 2^24 (mod.hpp:159-189; SURVEY
    flags it as unverified for tcgen05 kind::f8f6f4).  tools/f8_probe.py drives the raw-accumulator epilogue with adversarial rows
    (huge running sums followed by +-1 products) up to k = 65536 and compares with exact integer arithmetic."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, str(ROOT / "tools" / "f8_probe.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "F8 EXACT" in r.stdout and "NOT EXACT" not in r.stdout, r.stdout[-1500:]


@pytest.mark.parametrize("dtype,N", [(np.float64, 14), (np.float32, 6), (np.complex128, 10), (np.complex64, 6)])
@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("backend", [0, 1])
def test_native_kshard_driver_world1_matches_gemm(env, dtype, N, fast, backend):
    """The native multi-GPU driver (g8_mg_comm_* / g8_mg_plan_* / g8_gemm_mg, csrc/g8_mg.cu) with a world of ONE rank: the whole path --
    mailbox all-reduce, flag barrier, bound-plane exchange + chained bound GEMM, GEMM -> scatter, owner-side shard sum (+ complex 3M
    recombination) + CRT -- runs on a single GPU and must reproduce g8_gemm: accurate mode bit for bit, fast mode to the documented
    tolerance (its statistics kernels differ).  backend 1 = FP8: local contraction into int16 residues, slab copies, shard sum mod p, FP8
    CRT.  (2 - 8 ranks: tests/test_gpu_multi.py and tests/native/mg_check.cu.)"""
    import ctypes
    torch, H = env.torch, env.H
    from gemmul8_b200 import _lib, api

    lib = _lib.load()
    rng = np.random.default_rng(31)
    m, n, k = 200, 512, 700
    cplx = np.dtype(dtype).kind == "c"
    tdt = H.NP2T[np.dtype(dtype)]
    for opA, opB in (("N", "N"), ("T", "C" if cplx else "T")):
        A = H.rand_matrix(rng, H.stored_shape(opA, m, k), dtype)
        B = H.rand_matrix(rng, H.stored_shape(opB, k, n), dtype)
        want = H.run_gemm(A, B, opA, opB, N, fast, backend=backend)
        dA, lda = H.to_dev_colmajor(A)
        dB, ldb = H.to_dev_colmajor(B)
        dC, ldc = H.to_dev_colmajor(np.zeros((m, n), dtype=dtype))
        comm, plan = ctypes.c_void_p(), ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        api._check(lib.g8_mg_comm_create(ctypes.byref(comm), 1, 0, 8 * (m + n) + 4096, handle), "comm_create")
        api._check(lib.g8_mg_comm_connect(comm, handle), "comm_connect")
        api._check(lib.g8_mg_plan_create_backend(ctypes.byref(plan), comm, api._DTYPES[tdt], backend, api._op(opA), api._op(opB), m, n, k, N, int(fast)), "plan_create")
        keep = []
        pa, pb = api._scalar_ptr(1.0, tdt, keep), api._scalar_ptr(0.0, tdt, keep)
        for _ in range(2):
            api._check(lib.g8_gemm_mg(plan, pa, dA.data_ptr(), lda, dB.data_ptr(), ldb, pb, dC.data_ptr(), ldc, torch.cuda.current_stream().cuda_stream), "gemm_mg")
        torch.cuda.synchronize()
        assert lib.g8_mg_comm_status(comm) == 0
        got = H.from_dev_colmajor(dC, m, n, ldc)
        lib.g8_mg_plan_destroy(plan)
        lib.g8_mg_comm_destroy(comm)
        if not fast:
            assert H.bits_equal(got, want), H.first_diff(got, want, f"C {opA}{opB}")
        else:
            assert np.abs(got - want).max() <= (1e-9 if np.dtype(dtype).itemsize >= 8 and np.dtype(dtype) != np.dtype(np.complex64) else 1e-3) * np.abs(want).max()


def test_bound_gemm_chained_over_kslabs(env):
    """g8_stage_gemm_bound_chain: one accumulator sums the products of `chain` gathered K-slabs; row / column maxima vs numpy"""
    import ctypes
    torch = env.torch
    from gemmul8_b200 import _lib, api

    lib = _lib.load()
    rng = np.random.default_rng(5)
    m, n, kp, chain = 300, 200, 512, 3
    mp = api.pad256(m)
    A = rng.integers(0, 65, size=(chain, mp, kp)).astype(np.int8)
    B = rng.integers(0, 65, size=(chain, n, kp)).astype(np.int8)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    rowmax = torch.zeros(mp, dtype=torch.int32, device="cuda")
    colmax = torch.zeros(api.pad256(n), dtype=torch.int32, device="cuda")
    api._check(lib.g8_stage_gemm_bound_chain(dA.data_ptr(), mp * kp, dB.data_ptr(), n * kp, m, n, kp, chain, rowmax.data_ptr(), colmax.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "bound_chain")
    torch.cuda.synchronize()
    Hs = sum(A[c, :m].astype(np.int64) @ B[c].astype(np.int64).T for c in range(chain))
    assert np.array_equal(rowmax.cpu().numpy()[:m], Hs.max(axis=1)) and np.array_equal(colmax.cpu().numpy()[:n], Hs.max(axis=0))
