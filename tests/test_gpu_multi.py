"""Multi-GPU paths on real GPUs (need >= 2 devices; skipped otherwise): NCCL collectives + CUDA IPC peer memory + CUDA stage kernels,
on ALL visible GPUs (2, 4 or 8 ranks -- with more than 4 shards the owner side sums 8 per-shard residue arrays inside the CRT kernel).
K-shard, accurate mode: bit-identical to the single-GPU g8_gemm on the concatenated operands (all exchange variants).
N-shard and modulus-set shard: bit-identical in both modes."""
import os
import socket
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _world():
    import torch
    n = torch.cuda.device_count()
    return 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gemmul8_b200 as g8
        from gemmul8_b200 import multi_gpu

        ok_all = True
        for dtype, N in ((torch.float64, 14), (torch.float32, 6), (torch.complex128, 10), (torch.complex64, 6)):
            m, n, kl = 300, 256 * world, 384   # n / world = 256: one scatter tile per owner
            K = kl * world
            # full operands (same on every rank), column-major
            A = g8.randmat(m, K, dtype, phi=0.5, seed=11, device=f"cuda:{rank}")
            B = g8.randmat(K, n, dtype, phi=0.5, seed=22, device=f"cuda:{rank}")
            Ar = A.view(K, m)[rank * kl:(rank + 1) * kl].contiguous().view(-1)            # columns K_r of A
            Br = B.view(n, K)[:, rank * kl:(rank + 1) * kl].contiguous().view(-1)          # rows K_r of B
            for fast in (False, True):
                Cfull = torch.zeros(m * n, dtype=dtype, device=f"cuda:{rank}")
                tot, _, _ = g8.work_size(m, n, K, N, is_complex=dtype.is_complex)
                work = torch.empty(tot, dtype=torch.uint8, device=f"cuda:{rank}")
                g8.gemm("N", "N", m, n, K, 1.0, A, m, B, K, 0.0, Cfull, m, N, fast, work)
                for variant in ("int32", "residue", "fused", "fused-sumpass", "native"):
                    # "fused-sumpass": more than 4 shards with the separate residue_sum pass instead of the 8-part CRT kernel
                    os.environ["G8_MG_SUM_IN_CRT"] = "0" if variant == "fused-sumpass" else "1"
                    if variant == "fused-sumpass":
                        if world <= 4:
                            continue
                        variant = "fused"
                    if dtype.is_complex and variant != "native":
                        continue   # complex K-shards: the native driver (3M product units + owner-side recombination)
                    plan = (multi_gpu.NativeKShardGemm(m, n, kl, N, fastmode=fast, dtype=dtype, device=f"cuda:{rank}") if variant == "native" else
                            multi_gpu.KShardGemm(m, n, kl, N, fastmode=fast, dtype=dtype, device=f"cuda:{rank}", variant=variant))
                    C = torch.zeros(plan.local_out_elems, dtype=dtype, device=f"cuda:{rank}")
                    for _ in range(2 if variant in ("fused", "native") else 1):   # peer-mapped receive areas are re-used across steps
                        plan.run(Ar, Br, C)
                    torch.cuda.synchronize()
                    plan.close()
                    nc = n // world
                    want = Cfull.view(n, m)[rank * nc:(rank + 1) * nc].reshape(-1)
                    if not fast:
                        ok = torch.equal(C, want)          # accurate mode: bit-identical to the single-GPU call
                    else:
                        ok = bool(((C - want).abs().max() / want.abs().max()) < (1e-9 if dtype in (torch.float64, torch.complex128) else 1e-3))
                    ok_all &= ok
                    if not ok:
                        print(f"rank {rank} mismatch dtype={dtype} fast={fast} variant={variant}", flush=True)
        q.put((rank, ok_all))
    except Exception:
        q.put((rank, False))
        raise
    finally:
        dist.destroy_process_group()


def _f8_worker(rank, world, port, q):
    """FP8 backend, K-sharded through the native driver: accurate mode bit-identical to the single-GPU FP8 call (k_local is a multiple of
    32, so the chained bound GEMM performs the same binary32 MMA sequence), fast mode to tolerance"""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gemmul8_b200 as g8
        from gemmul8_b200 import multi_gpu

        ok_all = True
        for dtype, N in ((torch.float64, 12), (torch.float32, 6), (torch.complex128, 8)):
            m, n, kl = 300, 256 * world, 384
            K = kl * world
            A = g8.randmat(m, K, dtype, phi=0.5, seed=51, device=f"cuda:{rank}")
            B = g8.randmat(K, n, dtype, phi=0.5, seed=52, device=f"cuda:{rank}")
            Ar = A.view(K, m)[rank * kl:(rank + 1) * kl].contiguous().view(-1)
            Br = B.view(n, K)[:, rank * kl:(rank + 1) * kl].contiguous().view(-1)
            for fast in (False, True):
                Cfull = torch.zeros(m * n, dtype=dtype, device=f"cuda:{rank}")
                tot, _, _ = g8.work_size(m, n, K, N, is_complex=dtype.is_complex, backend=g8.Backend.FP8)
                work = torch.empty(tot, dtype=torch.uint8, device=f"cuda:{rank}")
                g8.gemm("N", "N", m, n, K, 1.0, A, m, B, K, 0.0, Cfull, m, N, fast, work, backend=g8.Backend.FP8)
                plan = multi_gpu.NativeKShardGemm(m, n, kl, N, fastmode=fast, dtype=dtype, device=f"cuda:{rank}", backend=int(g8.Backend.FP8))
                C = torch.zeros(plan.local_out_elems, dtype=dtype, device=f"cuda:{rank}")
                for _ in range(2):
                    plan.run(Ar, Br, C)
                torch.cuda.synchronize()
                plan.close()
                nc = n // world
                want = Cfull.view(n, m)[rank * nc:(rank + 1) * nc].reshape(-1)
                if not fast:
                    ok = torch.equal(C, want)
                else:
                    ok = bool(((C - want).abs().max() / want.abs().max()) < (1e-9 if dtype in (torch.float64, torch.complex128) else 1e-3))
                ok_all &= ok
                if not ok:
                    print(f"rank {rank} FP8 K-shard mismatch dtype={dtype} fast={fast} maxdiff={(C - want).abs().max().item()}", flush=True)
        q.put((rank, ok_all))
    except Exception:
        q.put((rank, False))
        raise
    finally:
        dist.destroy_process_group()


def test_kshard_fp8_native_matches_single_gpu(cuda):
    import torch.multiprocessing as mp

    world = _world()
    if world < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_f8_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok in res), res


def _nshard_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gemmul8_b200 as g8
        from gemmul8_b200 import multi_gpu

        ok_all = True
        for dtype, N in ((torch.float64, 14), (torch.float32, 6)):
            m, nl, k = 300, 160, 700
            n = nl * world
            A = g8.randmat(m, k, dtype, phi=1.0, seed=31, device=f"cuda:{rank}")
            B = g8.randmat(k, n, dtype, phi=1.0, seed=32, device=f"cuda:{rank}")
            Br = B.view(n, k)[rank * nl:(rank + 1) * nl].contiguous().view(-1)     # columns n_r of B
            for fast in (False, True):
                Cfull = torch.zeros(m * n, dtype=dtype, device=f"cuda:{rank}")
                tot, _, _ = g8.work_size(m, n, k, N)
                work = torch.empty(tot, dtype=torch.uint8, device=f"cuda:{rank}")
                g8.gemm("N", "N", m, n, k, 1.0, A, m, B, k, 0.0, Cfull, m, N, fast, work)
                plan = multi_gpu.NShardGemm(m, nl, k, N, fastmode=fast, dtype=dtype, device=f"cuda:{rank}")
                C = torch.zeros(m * nl, dtype=dtype, device=f"cuda:{rank}")
                plan.run(A, Br, C)
                torch.cuda.synchronize()
                want = Cfull.view(n, m)[rank * nl:(rank + 1) * nl].reshape(-1)
                ok = torch.equal(C, want)            # BOTH modes: bit-identical to the single-GPU call on the full B
                ok_all &= ok
                if not ok:
                    print(f"rank {rank} N-shard mismatch dtype={dtype} fast={fast}", flush=True)
        q.put((rank, ok_all))
    except Exception:
        q.put((rank, False))
        raise
    finally:
        dist.destroy_process_group()


def test_nshard_nccl_matches_single_gpu(cuda):
    import torch
    import torch.multiprocessing as mp

    world = _world()
    if world < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nshard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok in res), res


def test_kshard_nccl_matches_single_gpu(cuda):
    import torch
    import torch.multiprocessing as mp

    world = _world()
    if world < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok in res), res


def _modshard_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gemmul8_b200 as g8
        from gemmul8_b200 import multi_gpu

        ok_all = True
        for dtype, N in ((torch.float64, 14), (torch.float32, 6), (torch.float64, 3)):   # N = 3 < world (8 GPUs): some ranks own no modulus
            m, n, k = 300, 256 * world, 700
            A = g8.randmat(m, k, dtype, phi=1.0, seed=41, device=f"cuda:{rank}")
            B = g8.randmat(k, n, dtype, phi=1.0, seed=42, device=f"cuda:{rank}")
            for fast in (False, True):
                Cfull = torch.zeros(m * n, dtype=dtype, device=f"cuda:{rank}")
                tot, _, _ = g8.work_size(m, n, k, N)
                work = torch.empty(tot, dtype=torch.uint8, device=f"cuda:{rank}")
                g8.gemm("N", "N", m, n, k, 1.0, A, m, B, k, 0.0, Cfull, m, N, fast, work)
                plan = multi_gpu.ModShardGemm(m, n, k, N, fastmode=fast, dtype=dtype, device=f"cuda:{rank}")
                C = torch.zeros(plan.local_out_elems, dtype=dtype, device=f"cuda:{rank}")
                for _ in range(2):      # the peer-mapped C_mid is re-used across steps
                    plan.run(A, B, C)
                torch.cuda.synchronize()
                plan.close()
                nc = n // world
                want = Cfull.view(n, m)[rank * nc:(rank + 1) * nc].reshape(-1)
                ok = torch.equal(C, want)            # BOTH modes: shifts and planes are computed redundantly, identically, on every rank
                ok_all &= ok
                if not ok:
                    print(f"rank {rank} modulus-shard mismatch dtype={dtype} N={N} fast={fast}", flush=True)
        q.put((rank, ok_all))
    except Exception:
        q.put((rank, False))
        raise
    finally:
        dist.destroy_process_group()


def test_modshard_nccl_matches_single_gpu(cuda):
    import torch.multiprocessing as mp

    world = _world()
    if world < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_modshard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok in res), res


def test_native_multi_gpu_program_without_python(cuda):
    """tests/native/mg_check: one forked process per GPU drives g8_mg_comm_* / g8_mg_plan_* / g8_gemm_mg through the C ABI only (no
    Python, no NCCL, handles through a shared mapping) and compares every rank's slab with the single-GPU g8_gemm bit for bit"""
    import subprocess

    if _world() < 2:
        pytest.skip("needs 2 GPUs")
    exe = ROOT / "tests" / "native" / "_build" / "mg_check"
    if not exe.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "tests" / "native")])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mg_check OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
