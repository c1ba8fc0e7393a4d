"""K-sharded multi-GPU emulated GEMM (BASELINE.json config 3).  One process per GPU,
`torch.distributed` (NCCL over NVLink 5 / NVSwitch) for the exchange steps, the C-ABI stage kernels for all numerics.

The reference is single-GPU; this entry point is new work.  Rank r owns a K-slab of both operands,
    A_r = op(A)[:, K_r]   (m x k_r)        B_r = op(B)[K_r, :]   (k_r x n)
and the exact integer structure of Ozaki-II makes the sharding exact:

  shifts    fast    : local row statistics (max |x|, round-up sum x^2)  -> all_reduce(MAX / SUM) -> reference shift formula
            accurate: all_reduce(MAX) of max |x| -> s0; local bound planes; INT32 bound partial -> reduce_scatter(SUM) ->
                      slab row/col maxima -> all_reduce(MAX) / all_gather -> final shifts   (identical on every rank, and
                      identical to a single-GPU run on the concatenated operands: max and integer sums are order-free)
  split     local, with the global shifts (`g8_stage_split` mode 0)
  contraction  local tcgen05 INT8 GEMMs over all moduli; then ONE collective:
            variant "int32"  : raw INT32 partials [col][modulus][row]  -> reduce_scatter(SUM, int32)   (the north-star wording)
            variant "residue": partials reduced mod p in the GEMM epilogue (int8) -> all_to_all (4x fewer bytes) -> sum + mod
            variant "fused"  : ONE kernel does the GEMM and the exchange -- the tcgen05 epilogue stores each residue tile straight into
                               the owning rank's buffer over NVLink (peer memory mapped with CUDA IPC), so the transfer overlaps the
                               math tile by tile; a 1-element all_reduce orders the ranks, then sum + mod.  Accurate mode scatters its
                               INT32 bound partial the same way.  What bench.py --gpus N uses.
  CRT       every rank reconstructs its column slab C[:, n_r] (`g8_stage_crt`)

Exactness: |sum| <= K_total * 2^14 < 2^31 for K_total <= 2^17, so the INT32 reduction cannot overflow; the residue variant
sums at most `world` int8 values.  The result of accurate mode is bit-identical to the single-GPU call on the full K.

The numerics live behind a small `Stages` interface: `CudaStages` (product) drives the C ABI; tests substitute a CPU
implementation built on the oracle to exercise THIS file's orchestration with the gloo backend.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import api
from . import tables as T


def pad256(x: int) -> int:
    return 256 * ((x + 255) // 256)


class CudaStages:
    """Stage kernels through include/gemmul8_c.h on CUDA tensors.  No fallback: fails if the library or an sm_100a device is missing."""

    scatter_granularity = 256  # columns: the fused GEMM -> scatter hands whole 256-column tiles to one owner

    def __init__(self, dtype, num_moduli, device):
        from . import _lib

        self.lib = _lib.load()
        self.dtype, self.N, self.dev = dtype, num_moduli, torch.device(device)
        self.dt = api._DTYPES[dtype]

    def _s(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    # The A side and the B side of the preprocessing are independent: `with st.side():` runs the enclosed launches on a helper stream
    # forked from the current one; st.join() makes the current stream wait for it.  (CPU test stages implement both as no-ops.)
    def side(self):
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream(self.dev)
        self._side.wait_stream(torch.cuda.current_stream(self.dev))
        return torch.cuda.stream(self._side)

    def join(self):
        if hasattr(self, "_side"):
            torch.cuda.current_stream(self.dev).wait_stream(self._side)

    def empty(self, n, dtype):
        return torch.empty(n, dtype=dtype, device=self.dev)

    def zeros(self, n, dtype):
        return torch.zeros(n, dtype=dtype, device=self.dev)

    def stats(self, is_A, op, rows, k, X, ld):
        amax, ss = self.empty(rows, torch.float64), self.empty(rows, torch.float64)
        api._check(self.lib.g8_stage_stats(self.dt, int(is_A), op, rows, k, X.data_ptr(), ld, amax.data_ptr(), ss.data_ptr(), self._s()), "stats")
        return amax, ss

    def shift_from_stats(self, amax, ss, kind, sft):
        api._check(self.lib.g8_stage_shift_from_stats(amax.data_ptr(), ss.data_ptr() if ss is not None else None, amax.numel(), self.N, kind,
                                                      sft.data_ptr(), self._s()), "shift_from_stats")

    def split(self, is_A, op, rows, k, X, ld, mode, sft, planes, plane_stride):
        api._check(self.lib.g8_stage_split(self.dt, int(is_A), op, rows, k, X.data_ptr(), ld, self.N, mode, sft.data_ptr(), planes.data_ptr(),
                                           plane_stride, self.N, self._s()), "split")

    def gemm(self, epi, A_lo, strideA, B_lo, strideB, m, n, k_pad, units, out, out_stride, ldc, first=0):
        api._check(self.lib.g8_stage_gemm(epi, 0, A_lo.data_ptr(), strideA, B_lo.data_ptr(), strideB, m, n, k_pad, units, first, None, None,
                                          out.data_ptr(), out_stride, ldc, None, None, self._s()), "gemm")

    # ---- peer-mapped buffers (CUDA IPC) and the fused GEMM -> scatter ----
    def peer_buffer(self, nbytes, group=None):
        """Allocate `nbytes` on this rank, map the same buffer of every other rank; returns PeerBuffer (local tensor view + pointer table)."""
        return PeerBuffer(self.lib, nbytes, self.dev, group)

    def gemm_scatter(self, epi, A_lo, strideA, B_lo, strideB, m, n, k_pad, units, peers, elem_offset_bytes, out_stride, ldc, first=0):
        W = len(peers.ptrs)
        tbl = (ctypes.c_void_p * W)(*[p + elem_offset_bytes for p in peers.ptrs])
        api._check(self.lib.g8_stage_gemm_scatter(epi, A_lo.data_ptr(), strideA, B_lo.data_ptr(), strideB, m, n, k_pad, units, first, tbl, W,
                                                  peers.rank, out_stride, ldc, self._s()), "gemm_scatter")

    def gemm_bound_chain(self, A_planes, strideA, B_planes, strideB, m, n, k_pad, chain, rowmax, colmax):
        """bound GEMM over `chain` gathered K-slabs (plane c = slab c) with the row / column maxima fused into the epilogue"""
        api._check(self.lib.g8_stage_gemm_bound_chain(A_planes.data_ptr(), strideA, B_planes.data_ptr(), strideB, m, n, k_pad, chain, rowmax.data_ptr(),
                                                      colmax.data_ptr(), self._s()), "gemm_bound_chain")

    def gemm_bound(self, A_lo, strideA, B_lo, strideB, m, n, k_pad, rowmax, colmax):
        """bound GEMM of accurate mode with the row / column maxima fused into the epilogue (the bound product is never written)"""
        api._check(self.lib.g8_stage_gemm(2, 0, A_lo.data_ptr(), strideA, B_lo.data_ptr(), strideB, m, n, k_pad, 1, 0, None, None, None, 0,
                                          pad256(m), rowmax.data_ptr(), colmax.data_ptr(), self._s()), "gemm_bound")

    def maxabs_parts(self, parts, nparts, part_stride, rows, cols, ld, rowmax, colmax):
        api._check(self.lib.g8_stage_maxabs_i32_parts(parts.data_ptr(), nparts, part_stride, rows, cols, ld, rowmax.data_ptr(), colmax.data_ptr(),
                                                      self._s()), "maxabs_parts")

    def maxabs(self, C, rows, cols, ld, rowmax, colmax):
        api._check(self.lib.g8_stage_maxabs_i32(C.data_ptr(), rows, cols, ld, rowmax.data_ptr(), colmax.data_ptr(), self._s()), "maxabs")

    def finalize_shift(self, sft, cmax, count):
        api._check(self.lib.g8_stage_finalize_shift(sft.data_ptr(), cmax.data_ptr(), count, self.N, self._s()), "finalize_shift")

    def requant(self, C_hi, rows, cols, in_ld, in_us, units, C_mid, out_ld, out_us, first=0):
        api._check(self.lib.g8_stage_requant_i32(C_hi.data_ptr(), rows, cols, in_ld, in_us, units, first, C_mid.data_ptr(), out_ld, out_us, self._s()), "requant")

    def residue_sum(self, parts, nparts, part_stride, rows, cols, in_ld, in_us, units, C_mid, out_ld, out_us, first=0):
        api._check(self.lib.g8_stage_residue_sum(parts.data_ptr(), nparts, part_stride, rows, cols, in_ld, in_us, units, first, C_mid.data_ptr(),
                                                 out_ld, out_us, self._s()), "residue_sum")

    def crt_parts(self, parts, nparts, part_stride, ldmid, plane_stride, m, n, C, ldc, sftA, sftB, alpha, beta):
        keep = []
        pa, pb = api._scalar_ptr(alpha, self.dtype, keep), api._scalar_ptr(beta, self.dtype, keep)
        api._check(self.lib.g8_stage_crt_parts(self.dt, parts.data_ptr(), nparts, part_stride, ldmid, plane_stride, m, n, self.N, C.data_ptr(), ldc,
                                               sftA.data_ptr(), sftB.data_ptr(), pa, pb, self._s()), "crt_parts")

    def crt(self, C_mid, ldmid, plane_stride, m, n, C, ldc, sftA, sftB, alpha, beta):
        keep = []
        pa, pb = api._scalar_ptr(alpha, self.dtype, keep), api._scalar_ptr(beta, self.dtype, keep)
        api._check(self.lib.g8_stage_crt(self.dt, C_mid.data_ptr(), ldmid, plane_stride, m, n, self.N, C.data_ptr(), ldc, sftA.data_ptr(),
                                         sftB.data_ptr(), pa, pb, self._s()), "crt")


class _RawCuda:
    """__cuda_array_interface__ view of a raw device allocation (so torch can wrap memory we cudaMalloc'ed for IPC)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class PeerBuffer:
    """One device buffer per rank, each mapped into every rank of the node with CUDA IPC (g8_peer_alloc / g8_peer_open).
    ptrs[o] is rank o's buffer as seen from THIS process; `local` is this rank's buffer as a uint8 torch tensor."""

    def __init__(self, lib, nbytes, device, group=None):
        self.lib, self.group = lib, group
        self.rank, W = dist.get_rank(group), dist.get_world_size(group)
        handle = ctypes.create_string_buffer(64)
        p = ctypes.c_void_p()
        api._check(lib.g8_peer_alloc(nbytes, ctypes.byref(p), handle), "peer_alloc")
        self.own = p.value
        handles = [None] * W
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs, self.opened = [], []
        for o in range(W):
            if o == self.rank:
                self.ptrs.append(self.own)
                continue
            q = ctypes.c_void_p()
            api._check(lib.g8_peer_open(ctypes.create_string_buffer(handles[o], 64), ctypes.byref(q)), "peer_open")
            self.ptrs.append(q.value)
            self.opened.append(q.value)
        self.local = torch.as_tensor(_RawCuda(self.own, nbytes), device=device)

    def close(self):
        for q in self.opened:
            self.lib.g8_peer_close(q)
        self.opened = []
        if self.own:
            dist.barrier(group=self.group)  # nobody may still have our buffer mapped when it is freed
            self.lib.g8_peer_free(self.own)
            self.own = None


# ---- collectives (NCCL on GPUs; the gloo branches exist for the CPU tests of this orchestration) ----
def _is_gloo(group=None):
    return dist.get_backend(group) == "gloo"


class _Done:
    def wait(self):
        return None


def reduce_scatter_sum(out, inp, group=None, async_op=False):
    """out <- sum over ranks of inp[rank * out.numel() : (rank + 1) * out.numel()].  Returns a handle with .wait()."""
    if _is_gloo(group):
        tmp = inp.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        r = dist.get_rank(group)
        out.copy_(tmp[r * out.numel():(r + 1) * out.numel()])
        return _Done()
    return dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group, async_op=async_op) or _Done()


def all_to_all(out, inp, group=None, async_op=False):
    if _is_gloo(group):
        W = dist.get_world_size(group)
        gathered = [torch.empty_like(inp) for _ in range(W)]
        dist.all_gather(gathered, inp, group=group)
        r, c = dist.get_rank(group), out.numel() // W
        for j in range(W):
            out[j * c:(j + 1) * c].copy_(gathered[j][r * c:(r + 1) * c])
        return _Done()
    return dist.all_to_all_single(out, inp, group=group, async_op=async_op) or _Done()


class KShardGemm:
    """C[:, slab(rank)] = alpha * sum_r A_r B_r (+ beta * C) for real S/D GEMM, op N/N layout of the local slabs
    (A_r: m x k_local column-major with ld = m; B_r: k_local x n column-major with ld = k_local).
    n must be a multiple of the world size (each rank reconstructs n / world columns).
    Accurate mode is bit-identical to the single-GPU call on the concatenated operands (order-free integer reductions).  Fast mode
    is NOT guaranteed to be: its round-up sum of squares is reduced per shard and then across ranks, a different order than the
    single-GPU kernel's, so a shift can differ by one on a floor() boundary (the result then agrees to the emulated precision)."""

    def __init__(self, m, n, k_local, num_moduli, fastmode=False, dtype=torch.float64, device=None, variant="residue", stages=None,
                 group=None, pipeline_groups=1):
        if dtype not in (torch.float32, torch.float64):
            raise NotImplementedError("K-sharded path: real S/D GEMM only")
        self.group = group
        self.W, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if n % self.W:
            raise ValueError("n must be divisible by the world size")
        if self.W * k_local > 2 ** 17:
            # |sum| <= K_total * 2^14 must stay below 2^31 for the INT32 partials / the bound product (include/gemmul8.hpp:29)
            raise ValueError(f"K-sharded path: total K = world * k_local = {self.W * k_local} exceeds 2^17 (INT32 accumulation bound)")
        self.m, self.n, self.k, self.N = m, n, k_local, num_moduli
        self.fast, self.dtype, self.variant = bool(fastmode), dtype, variant
        self.st = stages if stages is not None else CudaStages(dtype, num_moduli, device)
        self.m_pad, self.k_pad, self.n_pad = pad256(m), pad256(k_local), pad256(n)
        self.nc = n // self.W
        self.sizeA, self.sizeB = self.k_pad * self.m_pad, self.k_pad * n
        st, N = self.st, num_moduli
        self.A_lo = st.empty(self.sizeA * N, torch.int8)
        self.B_lo = st.empty(self.sizeB * N, torch.int8)
        self.sftA = st.zeros(self.m_pad, torch.int16)
        self.sftB = st.zeros(self.n_pad, torch.int16)
        # the moduli can be processed in `pipeline_groups` batches so that the exchange of batch g overlaps the GEMMs of batch g+1.
        # Measured on 2 x B200 (r01): no gain -- the persistent GEMM owns all 148 SMs, so NCCL's copy kernels only get SMs between
        # launches -- hence the default of ONE batch = one collective per call.
        G = max(1, min(pipeline_groups, N))
        self.batches = [(i * N // G, (i + 1) * N // G - i * N // G) for i in range(G)]  # (first modulus, count)
        per = n * N * self.m_pad
        # [modulus][col][row]; the fused variant lets the CRT kernel sum the (up to 8) shards itself and needs no C_mid
        # (G8_MG_SUM_IN_CRT=0 restores the separate residue_sum pass of round 1 for more than 4 shards)
        sum_in_crt = variant == "fused" and (self.W <= 4 or os.environ.get("G8_MG_SUM_IN_CRT", "1") != "0")
        self.C_mid = None if sum_in_crt else st.empty(N * self.nc * self.m_pad, torch.int8)
        self.peers = None
        if variant == "fused":
            if self.nc % st.scatter_granularity:
                raise ValueError(f"fused variant: n / world must be a multiple of {st.scatter_granularity} (the scatter is tile-granular)")
            # this rank's receive area, written by every rank's GEMM epilogue: [src rank][modulus][col in slab][row] int8, followed
            # (accurate mode) by [src rank][col in slab][row] int32 for the bound partial
            self.recv_bytes = per
            self.cbar_off = per
            total = per + (0 if (self.fast or os.environ.get("G8_MG_BOUND", "planes") == "planes") else 4 * self.W * self.nc * self.m_pad)
            self.peers = st.peer_buffer(total, group)
            self.recv = self.peers.local[:per].view(torch.int8)
            if total > per:
                self.cbar_parts = self.peers.local[per:].view(torch.int32)
            self.token = st.zeros(1, torch.int32)
        else:
            self.part = st.empty(per, torch.int32 if variant == "int32" else torch.int8)      # per batch: [col][modulus in batch][row]
            self.recv = st.empty(per // self.W if variant == "int32" else per, self.part.dtype)  # my column slab (x world for residue)
        # accurate mode, bound product over the full K: "planes" (default) exchanges the int8 bound PLANES (all-gather of A-bar, all-to-all
        # of the B-bar column slabs; 4-16x fewer bytes than INT32 partial products) and multiplies them locally with the maxima fused into
        # the GEMM epilogue; "int32" (G8_MG_BOUND=int32) is the round-1 exchange of INT32 partials.
        self.bound = os.environ.get("G8_MG_BOUND", "planes")
        if not self.fast and self.bound == "planes":
            self.abar_all = st.empty(self.W * self.sizeA, torch.int8)
            self.bbar_slabs = st.empty(self.W * self.nc * self.k_pad, torch.int8)
        if not self.fast and variant != "fused" and self.bound != "planes":
            self.cbar = st.empty(n * self.m_pad, torch.int32)
            self.cbar_slab = st.empty(self.nc * self.m_pad, torch.int32)
        self.local_out_elems = m * self.nc
        self._trace = None

    def local_out(self, C):
        return C[:self.local_out_elems]

    def _shifts(self, A, B):
        st, m, n, k, N = self.st, self.m, self.n, self.k, self.N
        amaxA, ssA = st.stats(True, 0, m, k, A, m)
        with st.side():
            amaxB, ssB = st.stats(False, 0, n, k, B, k)
        st.join()
        self._mark("stats")
        both_max = torch.cat([amaxA, amaxB])
        dist.all_reduce(both_max, op=dist.ReduceOp.MAX, group=self.group)
        amaxA, amaxB = both_max[:m], both_max[m:]
        if self.fast:
            both_ss = torch.cat([ssA, ssB])
            dist.all_reduce(both_ss, op=dist.ReduceOp.SUM, group=self.group)
            st.shift_from_stats(amaxA.contiguous(), both_ss[:m].contiguous(), 0, self.sftA)
            st.shift_from_stats(amaxB.contiguous(), both_ss[m:].contiguous(), 0, self.sftB)
            return
        # accurate: s0 from the global max, bound planes (aliasing plane 0), bound partial, reduce, maxima, final shifts
        st.shift_from_stats(amaxA.contiguous(), None, 1, self.sftA)
        st.shift_from_stats(amaxB.contiguous(), None, 1, self.sftB)
        st.split(True, 0, m, k, A, m, 3, self.sftA, self.A_lo, self.sizeA)
        with st.side():
            st.split(False, 0, n, k, B, k, 3, self.sftB, self.B_lo, self.sizeB)
        rowmax = st.zeros(self.m_pad, torch.int32)
        colmax_slab = st.zeros(self.nc, torch.int32)
        if self.bound == "planes":
            # A-bar travels (all-gather, the bulk of the exchange) while the B-bar kernel still runs on the helper stream
            if _is_gloo(self.group):
                dist.all_gather(list(self.abar_all.view(self.W, self.sizeA).unbind(0)), self.A_lo[:self.sizeA].contiguous(), group=self.group)
            else:
                dist.all_gather_into_tensor(self.abar_all, self.A_lo[:self.sizeA], group=self.group)
            st.join()
            self._mark("bound planes + A-bar gather")
            all_to_all(self.bbar_slabs, self.B_lo[:self.sizeB], self.group)
            self._mark("B-bar slab exchange")
            st.gemm_bound_chain(self.abar_all, self.sizeA, self.bbar_slabs, self.nc * self.k_pad, m, self.nc, self.k_pad, self.W, rowmax, colmax_slab)
        elif self.variant == "fused":
            # INT32 bound partial scattered by the GEMM epilogue into the owners' [src rank][col][row] areas, then summed + maxed there
            st.join()
            slab = self.nc * self.m_pad
            st.gemm_scatter(1, self.A_lo, self.sizeA, self.B_lo, self.sizeB, m, n, self.k_pad, 1, self.peers,
                            self.cbar_off + 4 * self.rank * slab, 0, self.m_pad)
            self._mark("bound gemm+scatter")
            self._rank_barrier()
            st.maxabs_parts(self.cbar_parts, self.W, slab, m, self.nc, self.m_pad, rowmax, colmax_slab)
        else:
            st.join()
            st.gemm(1, self.A_lo, self.sizeA, self.B_lo, self.sizeB, m, n, self.k_pad, 1, self.cbar, 0, self.m_pad)
            reduce_scatter_sum(self.cbar_slab, self.cbar, self.group)
            st.maxabs(self.cbar_slab, m, self.nc, self.m_pad, rowmax, colmax_slab)
        self._mark("bound max")
        dist.all_reduce(rowmax, op=dist.ReduceOp.MAX, group=self.group)
        colmax = st.zeros(self.n, torch.int32)
        dist.all_gather_into_tensor(colmax, colmax_slab, group=self.group) if not _is_gloo(self.group) else \
            dist.all_gather(list(colmax.view(self.W, self.nc).unbind(0)), colmax_slab, group=self.group)
        st.finalize_shift(self.sftA, rowmax, m)
        st.finalize_shift(self.sftB, colmax, n)

    def _rank_barrier(self):
        """Stream-ordered barrier across the ranks: a 1-element all_reduce.  It cannot complete on any rank before every rank has
        enqueued it, i.e. before every rank's preceding kernel (the scatter GEMM) has finished -- kernel completion makes its peer
        stores visible system-wide."""
        dist.all_reduce(self.token, op=dist.ReduceOp.MAX, group=self.group)

    def close(self):
        if self.peers is not None:
            self.peers.close()
            self.peers = None

    # ---- optional phase trace (G8_MG_TRACE=1): CUDA events on the compute stream, reported by trace_report() ----
    def _mark(self, name):
        if self._trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._trace.append((name, e))

    def trace_report(self):
        """[(phase, ms)] of the last run() when tracing is on (synchronises)."""
        if not self._trace:
            return []
        torch.cuda.synchronize()
        return [(self._trace[i + 1][0], self._trace[i][1].elapsed_time(self._trace[i + 1][1])) for i in range(len(self._trace) - 1)]

    def run(self, A, B, C, alpha=1.0, beta=0.0):
        """A: flat m*k_local, B: flat k_local*n (this rank's K-slab), C: flat buffer receiving the m x (n/world) slab (ld = m)."""
        st, m, n, k, N = self.st, self.m, self.n, self.k, self.N
        self._trace = [] if (os.environ.get("G8_MG_TRACE") == "1" and torch.cuda.is_available()) else None
        self._mark("start")
        self._shifts(A, B)
        self._mark("shifts")
        st.split(True, 0, m, k, A, m, 0, self.sftA, self.A_lo, self.sizeA)
        with st.side():
            st.split(False, 0, n, k, B, k, 0, self.sftB, self.B_lo, self.sizeB)
        st.join()
        self._mark("split")
        mp, nc, W = self.m_pad, self.nc, self.W
        if self.variant == "fused":
            # (the collectives of _shifts above ordered this step after every rank's reads of the previous step's receive area)
            st.gemm_scatter(0, self.A_lo, self.sizeA, self.B_lo, self.sizeB, m, n, self.k_pad, N, self.peers,
                            self.rank * N * nc * mp, nc * mp, mp)
            self._mark("gemm+scatter")
            self._rank_barrier()
            self._mark("barrier")
            r0 = self.rank * nc
            if self.C_mid is None:
                # the CRT sums the per-shard residues itself (no separate sum pass, no C_mid round trip): measured faster up to 4 shards
                st.crt_parts(self.recv, W, N * nc * mp, mp, nc * mp, m, nc, C, m, self.sftA, self.sftB[r0:r0 + nc], alpha, beta)
            else:
                st.residue_sum(self.recv, W, N * nc * mp, mp, nc, mp, nc * mp, N, self.C_mid, mp, nc * mp)
                st.crt(self.C_mid, mp, nc * mp, m, nc, C, m, self.sftA, self.sftB[r0:r0 + nc], alpha, beta)
            self._mark("sum+crt")
            return C
        pending = []  # (handle, batch) whose exchange is in flight

        def finish(h, u0, nu, recv):
            h.wait()  # the compute stream waits for the collective of this batch
            out = self.C_mid[u0 * nc * mp:]
            if self.variant == "int32":
                st.requant(recv, mp, nc, nu * mp, mp, nu, out, mp, nc * mp, first=u0)
            else:
                st.residue_sum(recv, W, nc * nu * mp, mp, nc, nu * mp, mp, nu, out, mp, nc * mp, first=u0)

        for (u0, nu) in self.batches:
            # batch buffers are carved from the big ones: send [n][nu][mp], receive slab [nc][nu][mp] (x W for the residue variant)
            part = self.part[u0 * n * mp:(u0 + nu) * n * mp]
            A_g, B_g = self.A_lo[u0 * self.sizeA:], self.B_lo[u0 * self.sizeB:]
            if self.variant == "int32":
                recv = self.recv[u0 * nc * mp:(u0 + nu) * nc * mp]
                st.gemm(1, A_g, self.sizeA, B_g, self.sizeB, m, n, self.k_pad, nu, part, mp, nu * mp, first=u0)
                h = reduce_scatter_sum(recv, part, self.group, async_op=True)   # ONE reduce of the INT32 partials per batch
            else:
                recv = self.recv[u0 * n * mp:(u0 + nu) * n * mp]
                st.gemm(0, A_g, self.sizeA, B_g, self.sizeB, m, n, self.k_pad, nu, part, mp, nu * mp, first=u0)
                h = all_to_all(recv, part, self.group, async_op=True)
            if pending:
                finish(*pending.pop(0))
            pending.append((h, u0, nu, recv))
        while pending:
            finish(*pending.pop(0))
        self._mark("gemm+exchange+requant")
        r0 = self.rank * nc
        st.crt(self.C_mid, mp, nc * mp, m, nc, C, m, self.sftA, self.sftB[r0:r0 + nc], alpha, beta)
        self._mark("crt")
        return C


class NativeKShardGemm:
    """K-sharded emulated GEMM driven ENTIRELY by the native library (C ABI g8_mg_comm_* / g8_mg_plan_* / g8_gemm_mg, csrc/g8_mg.cu):
    orchestration in C++, every inter-rank exchange through the library's own kernels over NVLink peer memory (fused GEMM -> scatter,
    mailbox all-reduce in fixed rank order, flag barrier) -- no NCCL on the path.  torch.distributed is used ONCE, at construction, to
    pass the 64-byte IPC handles around (any transport would do; tests/native/mg_check.cu uses a shared mapping and no Python at all).
    Same interface as KShardGemm.  backend=Backend.FP8: local contraction into int16 residues, peer copies of the owners' slabs, shard
    sum mod p on the owner (world * k_local <= 2^16)."""

    def __init__(self, m, n, k_local, num_moduli, fastmode=False, dtype=torch.float64, device=None, group=None, op_A="N", op_B="N", backend=0):
        from . import _lib
        if dtype not in api._DTYPES:
            raise NotImplementedError("K-sharded path: S/D/C/Z GEMM")
        self.lib, self.group = _lib.load(), group
        self.W, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.m, self.n, self.k, self.N, self.dtype = m, n, k_local, num_moduli, dtype
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.opA, self.opB = api._op(op_A), api._op(op_B)
        self.nc = n // self.W
        self.local_out_elems = m * self.nc
        self.comm, self.plan = ctypes.c_void_p(), ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        vec = max(8 * (m + n), 4 * (pad256(m) + pad256(n)))
        with torch.cuda.device(self.dev):
            api._check(self.lib.g8_mg_comm_create(ctypes.byref(self.comm), self.W, self.rank, vec, handle), "g8_mg_comm_create")
            handles = [None] * self.W
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            api._check(self.lib.g8_mg_comm_connect(self.comm, ctypes.create_string_buffer(b"".join(handles), 64 * self.W)), "g8_mg_comm_connect")
            dist.barrier(group=group)
            api._check(self.lib.g8_mg_plan_create_backend(ctypes.byref(self.plan), self.comm, api._DTYPES[dtype], int(backend), self.opA, self.opB, m, n,
                                                          k_local, num_moduli, int(bool(fastmode))), "g8_mg_plan_create")
        self.lda = m if self.opA == 0 else k_local
        self.ldb = k_local if self.opB == 0 else n

    def local_out(self, C):
        return C[:self.local_out_elems]

    def trace_report(self):
        return []

    def run(self, A, B, C, alpha=1.0, beta=0.0):
        keep = []
        pa, pb = api._scalar_ptr(alpha, self.dtype, keep), api._scalar_ptr(beta, self.dtype, keep)
        with torch.cuda.device(self.dev):
            api._check(self.lib.g8_gemm_mg(self.plan, pa, A.data_ptr(), self.lda, B.data_ptr(), self.ldb, pb, C.data_ptr(), self.m,
                                           torch.cuda.current_stream(self.dev).cuda_stream), "g8_gemm_mg")
        return C

    def close(self):
        if self.plan:
            with torch.cuda.device(self.dev):
                self.lib.g8_mg_plan_destroy(self.plan)
                dist.barrier(group=self.group)
                self.lib.g8_mg_comm_destroy(self.comm)
            self.plan, self.comm = ctypes.c_void_p(), ctypes.c_void_p()


class ModShardGemm:
    """Modulus-set sharded multi-GPU emulated GEMM (SURVEY section 8e "modulus-set shard"): every rank holds ALL of A and B, computes
    the shifts and the residue planes redundantly (no communication, bit-identical on every rank by construction), contracts only ITS
    subset of the moduli -- 14 moduli over 8 GPUs: 2,2,2,2,2,2,1,1 -- and scatters the residue tiles of column slab o straight into
    rank o's C_mid through the fused GEMM -> NVLink epilogue; rank o then reconstructs C[:, slab(o)].  No reduction is needed: every
    (modulus, column) pair has exactly one producer.  The split (and accurate mode's bound GEMM) is replicated, so the scaling is
    bounded by Amdahl: it is the mode for problems whose K cannot be sharded (K-shard needs the caller to hold K-slabs)."""

    def __init__(self, m, n, k, num_moduli, fastmode=False, dtype=torch.float64, device=None, stages=None, group=None):
        if dtype not in (torch.float32, torch.float64):
            raise NotImplementedError("modulus-sharded path: real S/D GEMM only")
        self.group = group
        self.W, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.m, self.n, self.k, self.N = m, n, k, num_moduli
        self.fast, self.dtype = bool(fastmode), dtype
        self.st = st = stages if stages is not None else CudaStages(dtype, num_moduli, device)
        if n % self.W or (n // self.W) % st.scatter_granularity:
            raise ValueError(f"n / world must be a multiple of {st.scatter_granularity} (the scatter is tile-granular)")
        self.m_pad, self.k_pad, self.n_pad = pad256(m), pad256(k), pad256(n)
        self.nc = n // self.W
        self.sizeA, self.sizeB = self.k_pad * self.m_pad, self.k_pad * n
        N, W = num_moduli, self.W
        base, extra = divmod(N, W)
        counts = [base + (1 if r < extra else 0) for r in range(W)]
        self.u0, self.nu = sum(counts[:self.rank]), counts[self.rank]
        self.A_lo = st.empty(self.sizeA * N, torch.int8)
        self.B_lo = st.empty(self.sizeB * N, torch.int8)
        self.sftA = st.zeros(self.m_pad, torch.int16)
        self.sftB = st.zeros(self.n_pad, torch.int16)
        self.peers = st.peer_buffer(N * self.nc * self.m_pad, group)   # my C_mid: [modulus][col in slab][row] int8, filled by all ranks
        self.C_mid = self.peers.local.view(torch.int8)
        self.token = st.zeros(1, torch.int32)
        self.local_out_elems = m * self.nc

    def local_out(self, C):
        return C[:self.local_out_elems]

    def close(self):
        if self.peers is not None:
            self.peers.close()
            self.peers = None

    def run(self, A, B, C, alpha=1.0, beta=0.0):
        """A: flat m*k (ld = m), B: flat k*n (ld = k), identical on every rank; C: flat m x (n/world) slab (ld = m)."""
        st, m, n, k, N, mp, nc = self.st, self.m, self.n, self.k, self.N, self.m_pad, self.nc
        # every rank has the full K: the shifts come from the very kernels of the single-GPU call (split modes 1 / 2), hence the same bits
        if self.fast:
            st.split(True, 0, m, k, A, m, 1, self.sftA, self.A_lo, self.sizeA)
            st.split(False, 0, n, k, B, k, 1, self.sftB, self.B_lo, self.sizeB)
        else:
            st.split(True, 0, m, k, A, m, 2, self.sftA, self.A_lo, self.sizeA)     # s0 + bound planes (aliasing plane 0)
            st.split(False, 0, n, k, B, k, 2, self.sftB, self.B_lo, self.sizeB)
            rowmax, colmax = st.zeros(self.m_pad, torch.int32), st.zeros(self.n_pad, torch.int32)
            st.gemm_bound(self.A_lo, self.sizeA, self.B_lo, self.sizeB, m, n, self.k_pad, rowmax, colmax)
            st.finalize_shift(self.sftA, rowmax, m)
            st.finalize_shift(self.sftB, colmax, n)
            st.split(True, 0, m, k, A, m, 0, self.sftA, self.A_lo, self.sizeA)
            st.split(False, 0, n, k, B, k, 0, self.sftB, self.B_lo, self.sizeB)
        # (the previous step's CRT on every rank is ordered before these remote writes by the barrier at the end of that step)
        if self.nu:
            st.gemm_scatter(0, self.A_lo[self.u0 * self.sizeA:], self.sizeA, self.B_lo[self.u0 * self.sizeB:], self.sizeB, m, n, self.k_pad,
                            self.nu, self.peers, self.u0 * nc * mp, nc * mp, mp, first=self.u0)
        dist.all_reduce(self.token, op=dist.ReduceOp.MAX, group=self.group)   # all producers done: my C_mid is complete
        r0 = self.rank * nc
        st.crt(self.C_mid, mp, nc * mp, m, nc, C, m, self.sftA, self.sftB[r0:r0 + nc], alpha, beta)
        dist.all_reduce(self.token, op=dist.ReduceOp.MAX, group=self.group)   # nobody overwrites a C_mid that is still being read
        return C


class NShardGemm:
    """Column-sharded multi-GPU emulated GEMM (SURVEY section 8e "N-shard"): rank r holds ALL of A (m x k) and the column slab
    B[:, n_r] (k x n_local) and produces C[:, n_r].  No bulk exchange at all: the only cross-rank dependency is the accurate-mode
    shift of A, which needs the row maxima of the bound product over ALL columns -> one all_reduce(MAX) of m int32 values.
    Fast mode needs no communication.  Both modes are bit-identical to the single-GPU call on the full B (tests/test_gpu_multi.py)."""

    def __init__(self, m, n_local, k, num_moduli, fastmode=False, dtype=torch.float64, device=None, stages=None, group=None):
        if dtype not in (torch.float32, torch.float64):
            raise NotImplementedError("N-sharded path: real S/D GEMM only")
        self.group = group
        self.m, self.n, self.k, self.N = m, n_local, k, num_moduli
        self.fast, self.dtype = bool(fastmode), dtype
        self.st = st = stages if stages is not None else CudaStages(dtype, num_moduli, device)
        self.m_pad, self.k_pad, self.n_pad = pad256(m), pad256(k), pad256(n_local)
        self.sizeA, self.sizeB = self.k_pad * self.m_pad, self.k_pad * n_local
        N = num_moduli
        self.A_lo = st.empty(self.sizeA * N, torch.int8)
        self.B_lo = st.empty(self.sizeB * N, torch.int8)
        self.sftA = st.zeros(self.m_pad, torch.int16)
        self.sftB = st.zeros(self.n_pad, torch.int16)
        self.C_mid = st.empty(N * n_local * self.m_pad, torch.int8)

    def run(self, A, B, C, alpha=1.0, beta=0.0):
        """A: flat m*k (ld = m), B: flat k*n_local (ld = k), C: flat m*n_local (ld = m); all column-major."""
        st, m, n, k, N, mp = self.st, self.m, self.n, self.k, self.N, self.m_pad
        # every rank has the full K: the shifts come from the very kernels of the single-GPU call (split modes 1 / 2), hence the same bits
        if self.fast:
            st.split(True, 0, m, k, A, m, 1, self.sftA, self.A_lo, self.sizeA)
            st.split(False, 0, n, k, B, k, 1, self.sftB, self.B_lo, self.sizeB)
        else:
            st.split(True, 0, m, k, A, m, 2, self.sftA, self.A_lo, self.sizeA)     # s0 + bound planes (aliasing plane 0)
            st.split(False, 0, n, k, B, k, 2, self.sftB, self.B_lo, self.sizeB)
            rowmax, colmax = st.zeros(self.m_pad, torch.int32), st.zeros(self.n_pad, torch.int32)
            st.gemm_bound(self.A_lo, self.sizeA, self.B_lo, self.sizeB, m, n, self.k_pad, rowmax, colmax)
            dist.all_reduce(rowmax, op=dist.ReduceOp.MAX, group=self.group)          # the ONLY collective of this mode
            st.finalize_shift(self.sftA, rowmax, m)
            st.finalize_shift(self.sftB, colmax, n)
            st.split(True, 0, m, k, A, m, 0, self.sftA, self.A_lo, self.sizeA)
            st.split(False, 0, n, k, B, k, 0, self.sftB, self.B_lo, self.sizeB)
        st.gemm(0, self.A_lo, self.sizeA, self.B_lo, self.sizeB, m, n, self.k_pad, N, self.C_mid, n * mp, mp)
        st.crt(self.C_mid, mp, n * mp, m, n, C, m, self.sftA, self.sftB, alpha, beta)
        return C


def verify_against_single_gpu(shard, world, rank, dev, num_moduli, fastmode, variant="fused", m=512, k_local=512, backend=0):
    """bench.py --gpus N: before anything is timed, run the sharded path on a reduced problem and compare this rank's column slab with
    the SINGLE-GPU g8.gemm on the assembled operands (every rank regenerates all slabs from the seeds, so no gather is needed).
    Accurate mode must match bit for bit; fast mode within 1e-9 (see KShardGemm).  Returns a dict for the bench line."""
    from . import api as g8api
    dt = torch.float64
    n = 256 * world
    nc = n // world if shard in ("k", "mod") else n
    seedA = lambda r: 777 + (1000 * r if shard == "k" else 0)
    seedB = lambda r: 999 + (1000 * r if shard in ("k", "n") else 0)
    if shard == "k":
        if backend and variant != "native":
            raise NotImplementedError("FP8 K-shard: the native driver only")
        g = (NativeKShardGemm(m, n, k_local, num_moduli, fastmode=fastmode, dtype=dt, device=dev, backend=backend) if variant == "native" else
             KShardGemm(m, n, k_local, num_moduli, fastmode=fastmode, dtype=dt, device=dev, variant=variant))
        k_tot = k_local * world
        A_full = torch.cat([g8api.randmat(m, k_local, dt, phi=0.5, seed=seedA(r), device=dev) for r in range(world)])
        B_full = torch.cat([g8api.randmat(k_local, n, dt, phi=0.5, seed=seedB(r), device=dev).view(n, k_local) for r in range(world)], dim=1).contiguous().view(-1)
        A_loc = A_full[rank * m * k_local:(rank + 1) * m * k_local].contiguous()
        B_loc = g8api.randmat(k_local, n, dt, phi=0.5, seed=seedB(rank), device=dev)
        n_full, cols = n, slice(rank * nc, (rank + 1) * nc)
    elif shard == "n":
        g = NShardGemm(m, n, k_local, num_moduli, fastmode=fastmode, dtype=dt, device=dev)
        k_tot = k_local
        A_full = g8api.randmat(m, k_tot, dt, phi=0.5, seed=seedA(0), device=dev)
        B_full = torch.cat([g8api.randmat(k_tot, n, dt, phi=0.5, seed=seedB(r), device=dev) for r in range(world)])
        A_loc, B_loc = A_full, g8api.randmat(k_tot, n, dt, phi=0.5, seed=seedB(rank), device=dev)
        n_full, cols = n * world, slice(rank * n, (rank + 1) * n)
    else:
        g = ModShardGemm(m, n, k_local, num_moduli, fastmode=fastmode, dtype=dt, device=dev)
        k_tot = k_local
        A_full = g8api.randmat(m, k_tot, dt, phi=0.5, seed=seedA(0), device=dev)
        B_full = g8api.randmat(k_tot, n, dt, phi=0.5, seed=seedB(0), device=dev)
        A_loc, B_loc = A_full, B_full
        n_full, cols = n, slice(rank * nc, (rank + 1) * nc)
    C_loc = torch.zeros(m * nc, dtype=dt, device=dev)
    g.run(A_loc, B_loc, C_loc)
    tot, _, _ = g8api.work_size(m, n_full, k_tot, num_moduli, backend=backend)
    work = torch.empty(tot, dtype=torch.uint8, device=dev)
    C_full = torch.zeros(m * n_full, dtype=dt, device=dev)
    g8api.gemm("N", "N", m, n_full, k_tot, 1.0, A_full, m, B_full, k_tot, 0.0, C_full, m, num_moduli, fastmode, work, backend=backend)
    torch.cuda.synchronize(dev)
    want = C_full.view(n_full, m)[cols].reshape(-1)
    same = bool(torch.equal(want.view(torch.int64), C_loc.view(torch.int64)))
    rel = float(((want - C_loc).abs().max() / want.abs().max().clamp_min(1e-300)).item())
    ok = same if not fastmode else rel <= 1e-9
    flags = torch.tensor([int(ok), int(same)], dtype=torch.int32, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    relt = torch.tensor([rel], dtype=torch.float64, device=dev)
    dist.all_reduce(relt, op=dist.ReduceOp.MAX)
    if hasattr(g, "close"):
        g.close()
    res = {"problem": f"{shard}-shard DGEMM {m}x{n_full}x{k_tot} num_moduli={num_moduli} fastmode={int(bool(fastmode))} over {world} GPUs vs single-GPU g8.gemm",
           "all_ranks_ok": bool(flags[0].item()), "bit_identical_all_ranks": bool(flags[1].item()), "max_rel_diff": float(relt.item())}
    if not res["all_ranks_ok"]:
        raise RuntimeError(f"multi-GPU verification FAILED: {res}")
    return res
