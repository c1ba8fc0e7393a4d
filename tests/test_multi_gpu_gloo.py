"""K-sharded multi-GPU orchestration (gemmul8_b200/multi_gpu.py) exercised on CPU: world_size = 2, gloo backend,
with the CUDA stage kernels replaced by an oracle-backed `Stages` implementation (test infrastructure).  Checks that the
protocol -- statistics all-reduce, bound-product reduce-scatter, INT32 / residue exchange, slab-wise CRT -- reproduces the
single-process oracle result bit for bit."""
import ctypes
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from gemmul8_b200 import tables as T  # noqa: E402


def pad256(x):
    return 256 * ((x + 255) // 256)


class ShmPeerBuffer:
    """CPU stand-in for multi_gpu.PeerBuffer: one /dev/shm file per rank, mapped by every rank (what CUDA IPC does for HBM)."""

    def __init__(self, nbytes, group=None):
        self.rank, W = dist.get_rank(group), dist.get_world_size(group)
        tag = os.environ.get("MASTER_PORT", "0")
        self.paths = [f"/dev/shm/g8_gloo_test_{tag}_{id(self) % 1000}_{o}" for o in range(W)]
        name = [self.paths[self.rank]]
        np.memmap(name[0], dtype=np.uint8, mode="w+", shape=(nbytes,)).flush()
        names = [None] * W
        dist.all_gather_object(names, name[0], group=group)
        self.paths = names
        self.ptrs = [np.memmap(pth, dtype=np.uint8, mode="r+", shape=(nbytes,)) for pth in names]
        self.local = torch.from_numpy(self.ptrs[self.rank])
        self.group = group

    def close(self):
        dist.barrier(group=self.group)
        self.ptrs = []
        try:
            os.unlink(self.paths[self.rank])
        except OSError:
            pass


class OracleStages:
    """CPU stand-in for multi_gpu.CudaStages, built on oracle/ (numpy + g8_oracle.c)."""

    scatter_granularity = 1

    def side(self):      # the CUDA stages fork a helper stream here; on the CPU everything is sequential
        import contextlib
        return contextlib.nullcontext()

    def join(self):
        pass

    def peer_buffer(self, nbytes, group=None):
        return ShmPeerBuffer(nbytes, group)

    def gemm_scatter(self, epi, A_lo, strideA, B_lo, strideB, m, n, k_pad, units, peers, offset_bytes, out_stride, ldc, first=0):
        a, b = A_lo.numpy(), B_lo.numpy()
        W = len(peers.ptrs)
        oc = n // W
        dt = np.int8 if epi == 0 else np.int32
        isz = np.dtype(dt).itemsize
        for u in range(units):
            Au = a[u * strideA:u * strideA + m * k_pad].reshape(m, k_pad).astype(np.int64)
            Bu = b[u * strideB:u * strideB + n * k_pad].reshape(n, k_pad).astype(np.int64)
            H = Au @ Bu.T
            if epi == 0:
                H = self._sym(H, self.mods[first + u])
            for c in range(n):
                o, cl = divmod(c, oc)
                raw = peers.ptrs[o]
                start = offset_bytes + (u * out_stride + cl * ldc) * isz
                raw[start:start + m * isz] = np.ascontiguousarray(H[:, c].astype(dt)).view(np.uint8)

    def gemm_bound(self, A_lo, strideA, B_lo, strideB, m, n, k_pad, rowmax, colmax):
        a, b = A_lo.numpy(), B_lo.numpy()
        H = a[:m * k_pad].reshape(m, k_pad).astype(np.int64) @ b[:n * k_pad].reshape(n, k_pad).astype(np.int64).T
        rowmax[:m] = torch.from_numpy(np.maximum(rowmax.numpy()[:m], H.max(axis=1, initial=0)).astype(np.int32))
        colmax[:n] = torch.from_numpy(np.maximum(colmax.numpy()[:n], H.max(axis=0, initial=0)).astype(np.int32))

    def gemm_bound_chain(self, A_planes, strideA, B_planes, strideB, m, n, k_pad, chain, rowmax, colmax):
        a, b = A_planes.numpy(), B_planes.numpy()
        H = sum(a[c * strideA:c * strideA + m * k_pad].reshape(m, k_pad).astype(np.int64) @ b[c * strideB:c * strideB + n * k_pad].reshape(n, k_pad).astype(np.int64).T
                for c in range(chain))
        rowmax[:m] = torch.from_numpy(np.maximum(rowmax.numpy()[:m], H.max(axis=1, initial=0)).astype(np.int32))
        colmax[:n] = torch.from_numpy(np.maximum(colmax.numpy()[:n], H.max(axis=0, initial=0)).astype(np.int32))

    def maxabs_parts(self, parts, nparts, part_stride, rows, cols, ld, rowmax, colmax):
        p = parts.numpy().astype(np.int64)
        tot = sum(p[q * part_stride:q * part_stride + cols * ld] for q in range(nparts)).astype(np.int32)
        self.maxabs(torch.from_numpy(tot), rows, cols, ld, rowmax, colmax)

    def __init__(self, dtype, N):
        from oracle import oracle as O

        self.O, self.L, self.N, self.dtype = O, O.lib(), N, dtype
        self.mods = T.moduli("INT8")[:N]
        self.log2P = np.float32(T.log2P("INT8", N))

    def empty(self, n, dtype):
        return torch.zeros(n, dtype=dtype)

    zeros = empty

    def _rows_view(self, is_A, rows, k, X, ld):
        x = X.numpy()
        return np.ascontiguousarray(x.reshape(k, ld)[:, :rows].T) if is_A else np.ascontiguousarray(x.reshape(rows, ld)[:, :k])

    def _operand(self, view):
        o = self.O.Operand.__new__(self.O.Operand)
        o.view, (o.rows, o.inner) = view, view.shape
        return o

    def stats(self, is_A, op, rows, k, X, ld):
        v = self._rows_view(is_A, rows, k, X, ld).astype(np.float64)
        return torch.from_numpy(np.abs(v).max(axis=1)), torch.from_numpy((v * v).sum(axis=1))

    def shift_from_stats(self, amax, ss, kind, sft):
        for i in range(amax.numel()):
            a = float(amax[i])
            if kind == 1:
                sft[i] = 5 - self.O.ilogb_exact(a)
            else:
                amb = ctypes.c_int(0)
                sft[i] = -self.L.g8o_fast_shift(a, float(ss[i]) * (1 + 2.0 ** -48), self.log2P, 0, ctypes.byref(amb))

    def split(self, is_A, op, rows, k, X, ld, mode, sft, planes, plane_stride):
        view = self._rows_view(is_A, rows, k, X, ld)
        k_pad = pad256(k)
        pl = planes.numpy()
        if mode in (1, 2):  # the single-GPU modes: statistics -> fast shift (1) / s0 (2), then the split (1) / the bound plane (2)
            amax, ss = self.stats(is_A, op, rows, k, X, ld)
            self.shift_from_stats(amax, ss, 0 if mode == 1 else 1, sft)
            mode = 0 if mode == 1 else 3
        if mode == 3:
            bar = np.zeros((rows, k_pad), dtype=np.int8)
            fn = self.L.g8o_extract_f if view.dtype == np.float32 else self.L.g8o_extract_d
            s0 = np.ascontiguousarray(sft.numpy()[:rows])
            fn(ctypes.c_void_p(view.ctypes.data), ctypes.c_size_t(k), 1, ctypes.c_size_t(rows), ctypes.c_size_t(k), ctypes.c_size_t(k_pad),
               ctypes.c_void_p(s0.ctypes.data), ctypes.c_void_p(bar.ctypes.data))
            pl[:rows * k_pad] = bar.reshape(-1)
            return
        res = self.O.split(self._operand(view), sft.numpy()[:rows], self.N)[0]
        for i in range(self.N):
            pl[i * plane_stride:i * plane_stride + rows * k_pad] = res[i].reshape(-1)

    def gemm(self, epi, A_lo, strideA, B_lo, strideB, m, n, k_pad, units, out, out_stride, ldc, first=0):
        a, b, o = A_lo.numpy(), B_lo.numpy(), out.numpy()
        for u in range(units):
            Au = a[u * strideA:u * strideA + m * k_pad].reshape(m, k_pad).astype(np.int64)
            Bu = b[u * strideB:u * strideB + n * k_pad].reshape(n, k_pad).astype(np.int64)
            H = Au @ Bu.T
            if epi == 0:
                p = self.mods[first + u]
                r = np.mod(H, p)
                H = np.where(r > p // 2, r - p, r)
            for c in range(n):
                o[c * ldc + u * out_stride:c * ldc + u * out_stride + m] = H[:, c].astype(o.dtype)

    def maxabs(self, C, rows, cols, ld, rowmax, colmax):
        c = C.numpy().reshape(cols, ld)[:, :rows]
        rowmax[:rows] = torch.from_numpy(np.maximum(rowmax.numpy()[:rows], c.max(axis=0, initial=0)).astype(np.int32))
        colmax[:cols] = torch.from_numpy(np.maximum(colmax.numpy()[:cols], c.max(axis=1, initial=0)).astype(np.int32))

    def finalize_shift(self, sft, cmax, count):
        for i in range(count):
            amb = ctypes.c_int(0)
            g = self.L.g8o_accu_shift(int(cmax[i]), self.log2P, ctypes.byref(amb))
            sft[i] = -(int(sft[i]) + g) if int(cmax[i]) > 0 else 0

    def _sym(self, H, p):
        r = np.mod(H, p)
        return np.where(r > p // 2, r - p, r).astype(np.int8)

    def requant(self, C_hi, rows, cols, in_ld, in_us, units, C_mid, out_ld, out_us, first=0):
        src, dst = C_hi.numpy().astype(np.int64), C_mid.numpy()
        for u in range(units):
            for c in range(cols):
                dst[u * out_us + c * out_ld:u * out_us + c * out_ld + rows] = self._sym(src[u * in_us + c * in_ld:u * in_us + c * in_ld + rows], self.mods[first + u])

    def residue_sum(self, parts, nparts, part_stride, rows, cols, in_ld, in_us, units, C_mid, out_ld, out_us, first=0):
        src, dst = parts.numpy().astype(np.int64), C_mid.numpy()
        for u in range(units):
            for c in range(cols):
                acc = sum(src[q * part_stride + u * in_us + c * in_ld:q * part_stride + u * in_us + c * in_ld + rows] for q in range(nparts))
                dst[u * out_us + c * out_ld:u * out_us + c * out_ld + rows] = self._sym(acc, self.mods[first + u])

    def crt_parts(self, parts, nparts, part_stride, ldmid, plane_stride, m, n, C, ldc, sftA, sftB, alpha, beta):
        tmp = torch.zeros(self.N * n * ldmid, dtype=torch.int8)
        self.residue_sum(parts, nparts, part_stride, ldmid, n, ldmid, plane_stride, self.N, tmp, ldmid, n * ldmid)
        self.crt(tmp, ldmid, n * ldmid, m, n, C, ldc, sftA, sftB, alpha, beta)

    def crt(self, C_mid, ldmid, plane_stride, m, n, C, ldc, sftA, sftB, alpha, beta):
        cm = C_mid.numpy().reshape(self.N, n, ldmid)
        out = self.O.crt(np.ascontiguousarray(cm), m, n, self.N, sftA.numpy()[:m], sftB.numpy()[:n], self.dtype, alpha, beta)
        C.numpy().reshape(n, ldc)[:, :m] = out.T


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, variant, fast, dtype_name, q, sum_in_crt="1", bound="planes"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), G8_MG_SUM_IN_CRT=sum_in_crt, G8_MG_BOUND=bound)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gemmul8_b200 import multi_gpu
        from oracle import oracle as O

        np_dt = np.dtype(dtype_name)
        t_dt = {"float64": torch.float64, "float32": torch.float32}[dtype_name]
        N = 14 if dtype_name == "float64" else 6
        rng = np.random.default_rng(42)  # same stream on both ranks
        m, n, kl = (37, 26, 40) if world == 2 else (21, 4 * world, 24)
        A = ((rng.random((m, kl * world)) - 0.5) * np.exp(rng.standard_normal((m, kl * world)))).astype(np_dt)
        B = ((rng.random((kl * world, n)) - 0.5) * np.exp(rng.standard_normal((kl * world, n)))).astype(np_dt)
        Ar, Br = A[:, rank * kl:(rank + 1) * kl], B[rank * kl:(rank + 1) * kl, :]
        tA = torch.from_numpy(np.asfortranarray(Ar).T.copy().reshape(-1))  # column-major m x kl
        tB = torch.from_numpy(np.asfortranarray(Br).T.copy().reshape(-1))  # column-major kl x n
        plan = multi_gpu.KShardGemm(m, n, kl, N, fastmode=fast, dtype=t_dt, variant=variant, stages=OracleStages(np_dt, N))
        C = torch.zeros(plan.local_out_elems, dtype=t_dt)
        plan.run(tA, tB, C)
        if variant == "fused":  # a second step re-uses the peer-mapped receive areas: the inter-step ordering must hold
            C.zero_()
            plan.run(tA, tB, C)
        nc = n // world
        got = C.numpy().reshape(nc, m).T
        sA, sB = plan.sftA.numpy()[:m].copy(), plan.sftB.numpy()[:n].copy()
        # single-process reference on the concatenated operands (padded K is per-shard on the sharded side: zero columns only)
        ref = O.emulate(A, B, "N", "N", N, fast, sftA=None if not fast else sA, sftB=None if not fast else sB)
        ok_shift = np.array_equal(ref["sftA"], sA) and np.array_equal(ref["sftB"], sB)
        ok = np.array_equal(np.ascontiguousarray(got).view(np.uint8), np.ascontiguousarray(ref["C"][:, rank * nc:(rank + 1) * nc]).view(np.uint8))
        plan.close()
        q.put((rank, bool(ok), bool(ok_shift)))
    except Exception as e:  # report instead of letting the parent wait for the queue timeout
        q.put((rank, False, False))
        raise e
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant,bound", [("int32", "planes"), ("residue", "planes"), ("fused", "planes"), ("fused", "int32"), ("residue", "int32")])
@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("dtype_name", ["float64", "float32"])
def test_kshard_two_ranks_matches_single_process(variant, bound, fast, dtype_name):
    if fast and bound == "int32":
        pytest.skip("the bound exchange only exists in accurate mode")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, variant, fast, dtype_name, q, "1", bound)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ok_shift in res:
        assert ok_shift, f"rank {rank}: shifts differ from the single-process oracle"
        assert ok, f"rank {rank}: C slab differs from the single-process oracle"


def _nshard_worker(rank, world, port, fast, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gemmul8_b200 import multi_gpu
        from oracle import oracle as O

        N = 14
        rng = np.random.default_rng(43)  # same stream on both ranks
        m, nl, k = 37, 13, 80
        A = ((rng.random((m, k)) - 0.5) * np.exp(rng.standard_normal((m, k))))
        B = ((rng.random((k, nl * world)) - 0.5) * np.exp(rng.standard_normal((k, nl * world))))
        Br = B[:, rank * nl:(rank + 1) * nl]
        tA = torch.from_numpy(np.asfortranarray(A).T.copy().reshape(-1))
        tB = torch.from_numpy(np.asfortranarray(Br).T.copy().reshape(-1))
        plan = multi_gpu.NShardGemm(m, nl, k, N, fastmode=fast, dtype=torch.float64, stages=OracleStages(np.dtype(np.float64), N))
        C = torch.zeros(m * nl, dtype=torch.float64)
        plan.run(tA, tB, C)
        got = C.numpy().reshape(nl, m).T
        sA, sB = plan.sftA.numpy()[:m].copy(), plan.sftB.numpy()[:nl].copy()
        ref = O.emulate(A, B, "N", "N", N, fast, sftA=None if not fast else sA)   # single process, full B
        ok_shift = np.array_equal(ref["sftA"], sA) and (fast or np.array_equal(ref["sftB"][rank * nl:(rank + 1) * nl], sB))
        if fast:  # fast-mode B shifts: compare through the result computed with the oracle's own full-B shifts of these columns
            ref = O.emulate(A, B, "N", "N", N, fast, sftA=sA, sftB=np.concatenate([ref["sftB"][:rank * nl], sB, ref["sftB"][(rank + 1) * nl:]]))
        ok = np.array_equal(np.ascontiguousarray(got).view(np.uint8), np.ascontiguousarray(ref["C"][:, rank * nl:(rank + 1) * nl]).view(np.uint8))
        q.put((rank, bool(ok), bool(ok_shift)))
    except Exception as e:
        q.put((rank, False, False))
        raise e
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fast", [False, True])
def test_nshard_two_ranks_matches_single_process(fast):
    """column-sharded mode: accurate mode needs exactly one all_reduce(MAX) and then equals the single-process result on the full B"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nshard_worker, args=(r, world, port, fast, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ok_shift in res:
        assert ok_shift, f"rank {rank}: shifts differ from the single-process oracle"
        assert ok, f"rank {rank}: C slab differs from the single-process oracle"


@pytest.mark.parametrize("sum_in_crt", ["1", "0"])
def test_kshard_fused_eight_ranks(sum_in_crt):
    """world_size 8 (the box size): the fused variant with 8 shards, summed inside the CRT stage (default) or by the separate
    residue-sum pass (G8_MG_SUM_IN_CRT=0)"""
    world, port = 8, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, "fused", False, "float64", q, sum_in_crt)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ok_shift in res:
        assert ok_shift and ok, f"rank {rank}: differs from the single-process oracle"


def _modshard_worker(rank, world, port, fast, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gemmul8_b200 import multi_gpu
        from oracle import oracle as O

        rng = np.random.default_rng(44)  # same stream on all ranks: A and B are replicated
        m, n, k = 29, 4 * world, 70
        A = ((rng.random((m, k)) - 0.5) * np.exp(rng.standard_normal((m, k))))
        B = ((rng.random((k, n)) - 0.5) * np.exp(rng.standard_normal((k, n))))
        tA = torch.from_numpy(np.asfortranarray(A).T.copy().reshape(-1))
        tB = torch.from_numpy(np.asfortranarray(B).T.copy().reshape(-1))
        plan = multi_gpu.ModShardGemm(m, n, k, N, fastmode=fast, dtype=torch.float64, stages=OracleStages(np.dtype(np.float64), N))
        C = torch.zeros(plan.local_out_elems, dtype=torch.float64)
        for _ in range(2):  # the second step re-uses the peer-mapped C_mid
            plan.run(tA, tB, C)
        nc = n // world
        got = C.numpy().reshape(nc, m).T
        sA, sB = plan.sftA.numpy()[:m].copy(), plan.sftB.numpy()[:n].copy()
        ref = O.emulate(A, B, "N", "N", N, fast, sftA=None if not fast else sA, sftB=None if not fast else sB)
        ok_shift = np.array_equal(ref["sftA"], sA) and np.array_equal(ref["sftB"], sB)
        ok = np.array_equal(np.ascontiguousarray(got).view(np.uint8), np.ascontiguousarray(ref["C"][:, rank * nc:(rank + 1) * nc]).view(np.uint8))
        plan.close()
        q.put((rank, bool(ok), bool(ok_shift)))
    except Exception as e:
        q.put((rank, False, False))
        raise e
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N,fast", [(2, 14, False), (2, 14, True), (4, 3, False)])
def test_modshard_matches_single_process(world, N, fast):
    """modulus-set sharded mode: every rank contracts its subset of the moduli (world = 4, N = 3: one rank owns none) and scatters the
    residue tiles into the owners' C_mid; the reconstructed slab equals the single-process result"""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_modshard_worker, args=(r, world, port, fast, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ok_shift in res:
        assert ok_shift, f"rank {rank}: shifts differ from the single-process oracle"
        assert ok, f"rank {rank}: C slab differs from the single-process oracle"
