#!/usr/bin/env python3
"""bench.py -- headline measurement of the Ozaki-II hot path (BASELINE.json: emulated DGEMM TFLOPS at
m=n=k=8192, num_moduli=14, INT8 backend) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode accu|fast]

One "step" = one emulated DGEMM (split -> 14 INT8 GEMMs -> CRT) on synthetic inputs generated like the
reference harness does (testing/make_matrix.hpp:33-82, phi=-1 => i.i.d. standard normal; seeds 12345 / 54321).
Prints ONE JSON line (see the contract in the task statement).  `value` is device-resident throughput
(inputs already in HBM), `e2e` goes through the public API with pinned HOST buffers (H2D of A,B and D2H
of C inside the timed region).

Workloads.  N = 1: BASELINE.json config 2 (DGEMM 8192^3).  N > 1: BASELINE.json config 3, DGEMM 16384^3 K-sharded over
the N GPUs (k = 16384 / N per GPU; N = 8 is the named configuration), fused GEMM -> NVLink scatter; the old weak-K
workload (8192 x 8192 x 8192 N) is measured too and reported under `extra.weak_k`.

`--impl reference` times the UNMODIFIED reference library (oracle/_ref/libgemmul8_ref.so, cuBLASLt-backed
gemmul8::gemmLt) on ONE GPU, same problem, same protocol: the reference has no CPU implementation of this path, so its
own GPU path is "the reference arm".  That arm imports NOTHING from gemmul8_b200: inputs come from the reference
harness' own generator (ref_randmat), device memory from torch.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="accu", choices=["accu", "fast"])
    ap.add_argument("--size", type=int, default=0, help="m = n (default: 8192 on one GPU, 16384 on several)")
    ap.add_argument("--moduli", type=int, default=14)
    ap.add_argument("--backend", default="int8", choices=["int8", "fp8"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the power / native-DGEMM / tensor-ceiling / weak-K legs")
    ap.add_argument("--mg-variant", default="native", choices=["int32", "residue", "fused", "native"],
                    help="K-shard exchange: int32 / residue = NCCL collectives after the GEMM; fused = GEMM -> NVLink scatter kernel, orchestrated "
                         "from Python with NCCL for the small vectors; native = the same kernels driven by the C ABI g8_gemm_mg (no NCCL, no Python)")
    ap.add_argument("--mg-shard", default="k", choices=["k", "n", "mod"],
                    help="multi-GPU sharding: k = K-sharded (the north-star path, default), n = column-sharded (every rank holds A and a "
                         "column slab of B / C; no bulk exchange), mod = modulus-set sharded (every rank holds A and B, contracts a subset "
                         "of the moduli, all-gathers the residue planes)")
    ap.add_argument("--k-local", type=int, default=0, help="K-slab per GPU in the K-sharded runs (default: size / gpus, i.e. the square "
                                                           "problem of BASELINE.json config 3; --k-local 8192 --size 8192 = the weak-K workload)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks / power
class ClockSampler:
    """SM clock / power / throttle reasons sampled through NVML every ~4 ms from a thread while the timed region runs
    (the nvidia-smi -lms loop of the profiling recipe gives only 1-2 samples for a 100 ms region)."""

    def __init__(self, index=0, period=0.004):
        self.index, self.rows, self._stop, self.th, self.h, self.period = index, [], threading.Event(), None, None, period
        self.errors, self.last_error = 0, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv, h = self.nv, self.h
        while not self._stop.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(h), time.perf_counter()))
            except Exception as e:  # keep sampling; report how many reads failed
                self.errors += 1
                self.last_error = repr(e)
            time.sleep(self.period)

    def start(self):
        if self.h is not None:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()

    def stop(self):
        if self.th is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self._stop.set()
        self.th.join()
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
                 "hw_power_brake": nv.nvmlClocksEventReasonHwPowerBrakeSlowdown}
        sm = [float(r[0]) for r in self.rows]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0}
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": self.max_sm,
                "power_w_median": round(statistics.median(r[1] for r in self.rows), 1),
                "power_w_max": round(max(r[1] for r in self.rows), 1), "reasons": reasons, "samples": len(sm),
                "read_errors": self.errors, "note": "NVML, 4 ms period, timed region only; NVML power is a ~1 s running average"}


def power_leg(fn, sync, flops_per_call, index=0, seconds=6.0, settle=2.0):
    """The reference's third axis (testing/getWatt.hpp:42-121, test_watt.hpp:5-263): run `fn` back to back for `seconds`, sample the
    board power through NVML every 100 ms on a thread, report the mean power of the samples taken after `settle` seconds (NVML's
    reading is a ~1 s running average) and GFLOPS per watt for the calls of the whole loop."""
    smp = ClockSampler(index, period=0.1)
    fn(); sync()
    smp.start()
    t0 = time.perf_counter()
    calls = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(8):
            fn()
        calls += 8
        sync()
    dt = time.perf_counter() - t0
    smp._stop.set()
    if smp.th is not None:
        smp.th.join()
    watts = [r[1] for r in smp.rows if r[3] - t0 >= settle]
    clk = [float(r[0]) for r in smp.rows if r[3] - t0 >= settle]
    if not watts:
        return None
    w = statistics.mean(watts)
    tf = flops_per_call * calls / dt * 1e-12
    return {"tflops": round(tf, 2), "watts": round(w, 1), "gflops_per_watt": round(tf * 1e3 / w, 2), "sm_mhz_median": statistics.median(clk),
            "seconds": round(dt, 2), "calls": calls, "samples": len(watts)}


# ------------------------------------------------------------------------------------------------ reference arm
class RefLib:
    """ctypes view of oracle/_ref/libgemmul8_ref.so (built from the unmodified reference sources by oracle/Makefile)."""

    def __init__(self):
        so = ROOT / "oracle" / "_ref" / "libgemmul8_ref.so"
        if not so.exists():
            raise FileNotFoundError(str(so))
        L = ctypes.CDLL(str(so))
        L.ref_work_size.restype = ctypes.c_size_t
        L.ref_work_size.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_size_t] * 3 + [ctypes.c_uint, ctypes.c_int, ctypes.c_int,
                                                                                            ctypes.c_void_p, ctypes.c_void_p]
        L.ref_gemm.restype = ctypes.c_int
        L.ref_gemm.argtypes = [ctypes.c_int] * 5 + [ctypes.c_size_t] * 3 + \
            [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
             ctypes.c_size_t, ctypes.c_uint, ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
        if hasattr(L, "ref_randmat"):
            L.ref_randmat.restype = ctypes.c_int
            L.ref_randmat.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_double, ctypes.c_ulonglong]
        if hasattr(L, "ref_lt_lowgemm"):
            L.ref_lt_lowgemm.restype = ctypes.c_int
            L.ref_lt_lowgemm.argtypes = [ctypes.c_int] + [ctypes.c_size_t] * 3 + [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_void_p]
        self.L = L


def event_time(torch, fn, reps, stream):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def lt_ceiling_reference(torch, ref, dev, stream, S=8192):
    """Plain cuBLASLt TN GEMM exactly as the reference's inner call sets it up (matmult.hpp:41-101,165-169), s8 x s8 -> s32 and
    e4m3 x e4m3 -> f32, 8192^3: best single call (burst) and 60 back-to-back calls (sustained).  The practical tensor ceilings of BASELINE.md 2a."""
    out = {}
    ws = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
    for be, name in ((0, "int8"), (1, "fp8_e4m3")):
        A = torch.randint(-127, 128, (S * S,), dtype=torch.int8, device=dev) if be == 0 else \
            torch.randint(-8, 9, (S * S,), device=dev).to(torch.float32).to(torch.float8_e4m3fn)
        B = A.clone()
        C = torch.empty(S * S, dtype=torch.int32 if be == 0 else torch.float32, device=dev)

        def call():
            rc = ref.L.ref_lt_lowgemm(be, S, S, S, A.data_ptr(), B.data_ptr(), C.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream.cuda_stream))
            if rc:
                raise RuntimeError(f"ref_lt_lowgemm({name}) -> {rc}")
        try:
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            burst = min(event_time(torch, call, 1, stream) for _ in range(10))
            sust = event_time(torch, call, 60, stream)
            ops = 2.0 * S ** 3
            out[name] = {"burst_tops": round(ops / burst * 1e-9, 1), "sustained_tops": round(ops / sust * 1e-9, 1)}
        except Exception as e:  # a missing heuristic must not kill the arm
            out[name] = {"error": str(e)}
        del A, B, C
    out["how"] = f"cublasLtMatmul TN {S}^3 through oracle/_ref (ref_lt_lowgemm): best of 10 single calls / 60 back-to-back calls, CUDA events"
    return out


def int8_ceiling_torch(torch, dev, stream, S=8192):
    """cuBLASLt s8 x s8 -> s32 8192^3 through torch._int_mm (library call, NOT on the product path): the measured INT8 tensor
    ceiling of this GPU at the clocks the power cap grants, next to the 2 x bf16 proxy of MEASURED_PEAKS.json."""
    try:
        a = torch.randint(-127, 128, (S, S), dtype=torch.int8, device=dev)
        b = torch.randint(-127, 128, (S, S), dtype=torch.int8, device=dev).t()  # column-major B: the TN form cuBLASLt prefers
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        burst = min(event_time(torch, lambda: torch._int_mm(a, b), 1, stream) for _ in range(10))
        sust = event_time(torch, lambda: torch._int_mm(a, b), 60, stream)
        ops = 2.0 * S ** 3
        return {"burst_tops": round(ops / burst * 1e-9, 1), "sustained_tops": round(ops / sust * 1e-9, 1),
                "how": f"torch._int_mm (cuBLASLt s8*s8->s32) {S}^3: best of 10 single calls / 60 back-to-back calls, CUDA events"}
    except Exception as e:
        return {"error": str(e)}


# ------------------------------------------------------------------------------------------------ cpu baseline
def cpu_host_openblas(target_s=8.0):
    """Host OpenBLAS DGEMM (NumPy's bundled scipy-openblas) on this box's cores: the native-FP64 CPU figure BASELINE.json asks to
    report next to the GPU numbers.  Bounded sample: the largest power-of-two cube that fits ~target_s."""
    import numpy as np
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    n = 1024
    a = rng.standard_normal((n, n)); b = rng.standard_normal((n, n))
    a @ b
    t0 = time.perf_counter(); a @ b; t1 = time.perf_counter() - t0
    gf = 2 * n ** 3 / t1
    size = n
    while size < 8192 and 3 * 2 * (2 * size) ** 3 / gf < target_s:
        size *= 2
    a = rng.standard_normal((size, size)); b = rng.standard_normal((size, size))
    a @ b
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); a @ b; ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return {"value": round(2 * size ** 3 / t * 1e-12, 4), "unit": "TFLOPS", "cores": cores,
            "sample": f"host OpenBLAS DGEMM {size}^3 via numpy (scipy-openblas), {cores} threads, median of 3 (native FP64, not the emulation)"}


def cpu_baseline_oracle_port(N=14, fast=False):
    """The CPU baseline of this tier: the oracle's restatement of the reference algorithm (oracle/g8_oracle.c, plain C, one thread) on
    a bounded sample of the same workload (emulated DGEMM, same num_moduli / mode)."""
    import numpy as np
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    m = n = 320; k = 1024
    A = rng.standard_normal((m, k)); B = rng.standard_normal((k, n))
    t0 = time.perf_counter()
    O.emulate(A, B, num_moduli=N, fastmode=fast)
    t = time.perf_counter() - t0
    return {"value": round(2 * m * n * k / t * 1e-12, 9), "unit": "TFLOPS", "cores": 1, "kind": "port",
            "sample": f"oracle/g8_oracle.c (CPU restatement of split -> {N} exact int GEMMs -> CRT, scalar C, 1 thread) emulated DGEMM "
                      f"{m}x{n}x{k} num_moduli={N} fastmode={int(fast)}: {t:.1f} s"}


def ncu_traffic_bytes(kernel_substr):
    """DRAM bytes per launch of a kernel from the newest committed ncu summary (captured once per kernel change with `ncu --set full`)."""
    import csv
    for name in ("r02_ncu_full_accu_summary.csv", "r02_ncu_full_fast_summary.csv", "r01f_ncu_full_fast_summary.csv"):
        f = ROOT / "profiles" / name
        if not f.exists():
            continue
        rows = list(csv.reader(f.open()))
        hdr, units = rows[0], rows[1]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        for r in rows[2:]:
            if kernel_substr in r[0]:
                tot = 0.0
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    i = hdr.index(m)
                    tot += float(r[i]) * scale.get(units[i], 1.0)
                return int(tot), name
    return None, None


# ------------------------------------------------------------------------------------------------ reference arm main
def main_reference(args):
    """rank 0, one GPU, nothing of gemmul8_b200 imported."""
    import torch
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "reference", "unavailable": "no CUDA device (the reference library is GPU-only)"}))
        return 0
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    try:
        ref = RefLib()
    except (FileNotFoundError, OSError) as e:
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/libgemmul8_ref.so not loadable: {e}"}))
        return 0
    if not hasattr(ref.L, "ref_randmat"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgemmul8_ref.so predates ref_randmat: rebuild with make -C oracle ref"}))
        return 0
    S = args.size or (8192 if args.gpus == 1 else 16384)
    N, fast, be = args.moduli, args.mode == "fast", (0 if args.backend == "int8" else 1)
    m = n = k = S
    dt = torch.float64
    stream = torch.cuda.current_stream(dev)
    A = torch.empty(m * k, dtype=dt, device=dev); B = torch.empty(k * n, dtype=dt, device=dev); C = torch.zeros(m * n, dtype=dt, device=dev)
    assert ref.L.ref_randmat(1, A.data_ptr(), m, k, -1.0, 12345) == 0
    assert ref.L.ref_randmat(1, B.data_ptr(), k, n, -1.0, 54321) == 0
    tot = ref.L.ref_work_size(0, be, m, n, k, N, 0, 0, None, None)
    work = torch.empty(tot, dtype=torch.uint8, device=dev)
    one, zero = (ctypes.c_double * 1)(1.0), (ctypes.c_double * 1)(0.0)

    def step(timing=None):
        code = ref.L.ref_gemm(1, be, 1, 0, 0, m, n, k, ctypes.addressof(one), A.data_ptr(), m, B.data_ptr(), k, ctypes.addressof(zero),
                              C.data_ptr(), m, N, int(fast), work.data_ptr(), None, None, 0, 0, 0, 0, ctypes.c_void_p(stream.cuda_stream), timing)
        assert code == 0, code

    hA = torch.empty(m * k, dtype=dt).pin_memory(); hA.copy_(A)
    hB = torch.empty(k * n, dtype=dt).pin_memory(); hB.copy_(B)
    hC = torch.empty(m * n, dtype=dt).pin_memory()

    def step_e2e():
        A.copy_(hA, non_blocking=True)
        B.copy_(hB, non_blocking=True)
        step()
        hC.copy_(C, non_blocking=True)

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        if sampler is not None:
            sampler.start()
        ms = event_time(torch, fn, steps, stream)
        return ms

    warmup = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    ms_dev = timed(step, args.steps, warmup, sampler)
    clocks = sampler.stop()
    ms_e2e = timed(step_e2e, max(2, min(args.steps, 5)), 1)
    tm = (ctypes.c_double * 4)()
    ph = []
    for _ in range(3):
        step(tm)
        ph.append(list(tm))
    phases = [statistics.mean(p[i] for p in ph) for i in range(4)]
    flops = 2.0 * m * n * k
    value = flops / (ms_dev * 1e-3) * 1e-12
    out = {
        "metric": "emulated DGEMM TFLOPS @ N=8192 num_moduli=14; INT8 TC-pipe % of peak",
        "value": round(value, 2), "unit": "TFLOPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup,
        "ms_per_step": round(ms_dev, 4), "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None,
        "dtype": "int8 tensor-core residues (s8 x s8 -> s32) + f64 CRT; emulates f64", "data": "synthetic",
        "config": {"workload": f"DGEMM {m}x{n}x{k} {args.backend.upper()} num_moduli={N} fastmode={int(fast)} opN/opN alpha=1 beta=0"
                               + ("" if args.gpus == 1 else f" (the {args.gpus}-GPU arm's whole problem on ONE GPU: the reference library is single-GPU)"),
                   "inputs": "curand normal (phi=-1), seeds 12345/54321, generated by the reference's testing/make_matrix.hpp (ref_randmat)",
                   "l2": "inputs and residue planes are larger than the 126 MB L2; no explicit flush",
                   "timing": "CUDA events on the launch stream"},
        "e2e": {"value": round(flops / (ms_e2e * 1e-3) * 1e-12, 2), "unit": "TFLOPS", "ms_per_step": round(ms_e2e, 3),
                "h2d_bytes_per_step": int(hA.numel() * 8 + hB.numel() * 8), "d2h_bytes_per_step": int(hC.numel() * 8)},
        "gpu_launches": 0, "clocks": clocks, "impl": "reference",
        "phase_ms": {"split": round(phases[0] * 1e-6, 4), "gemm": round(phases[1] * 1e-6, 4), "requant": round(phases[2] * 1e-6, 4),
                     "crt": round(phases[3] * 1e-6, 4)},
    }
    out["cpu_baseline"] = {"value": out["value"], "unit": "TFLOPS", "cores": os.cpu_count(), "kind": "reference",
                           "sample": "the reference has NO CPU implementation of this path; this arm is the unmodified reference "
                                     "library (oracle/_ref, cuBLASLt-backed gemmul8::gemmLt) on one GPU, full workload per step"}
    if not args.no_extras:
        extra = {}
        del work
        torch.cuda.empty_cache()
        if hasattr(ref.L, "ref_lt_lowgemm"):
            extra["lt_ceiling"] = lt_ceiling_reference(torch, ref, dev, stream)
        work = torch.empty(tot, dtype=torch.uint8, device=dev)
        extra["power"] = {"reference_emulated_dgemm": power_leg(step, lambda: torch.cuda.synchronize(dev), flops, local_rank)}
        out["extra"] = extra
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------ ours
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        return 0 if rank != 0 else main_reference(args)  # rank 0 alone runs the reference arm

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    if not torch.cuda.is_available():
        print(json.dumps({"impl": args.impl, "error": "no CUDA device: gemmul8_b200 has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        try:  # pin this rank to the CPUs next to its GPU before any pinned host buffer is allocated (first touch decides the NUMA node)
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import gemmul8_b200 as g8
    from gemmul8_b200 import api

    N = args.moduli
    fast = args.mode == "fast"
    be = 0 if args.backend == "int8" else 1
    S = args.size or (16384 if distributed else 8192)
    m = n = S
    shard = args.mg_shard if distributed else None
    if shard == "k":
        k_local = args.k_local or S // world   # the square problem, K split over the ranks (BASELINE.json config 3 at 8 GPUs)
        k_total, n_total = k_local * world, n
    elif shard == "n":
        k_local, k_total, n_total = S, S, n * world   # column-sharded: every rank has the full K and n_local = S columns (weak scaling in n)
    else:
        k_local, k_total, n_total = (args.k_local or S), (args.k_local or S), n
    dt = torch.float64
    stream = torch.cuda.current_stream(dev)

    def make_inputs(k_loc, shard_kind):
        # synthetic inputs, generated on the device with the reference harness' generator; K-shards draw their slab with a rank-specific seed
        sa = 12345 + (0 if shard_kind in ("n", "mod", None) else 1000 * rank)
        sb = 54321 + (0 if shard_kind in ("mod", None) else 1000 * rank)
        return (g8.randmat(m, k_loc, dt, phi=-1.0, seed=sa, device=dev), g8.randmat(k_loc, n, dt, phi=-1.0, seed=sb, device=dev))

    def make_mg(k_loc, shard_kind):
        from gemmul8_b200 import multi_gpu
        if shard_kind == "n":
            g = multi_gpu.NShardGemm(m, n, k_loc, N, fastmode=fast, dtype=dt, device=dev)
            g.local_out_elems, g.local_out, g.trace_report = m * n, (lambda C_: C_), (lambda: [])
            return g
        if shard_kind == "mod":
            g = multi_gpu.ModShardGemm(m, n, k_loc, N, fastmode=fast, dtype=dt, device=dev)
            g.trace_report = lambda: []
            return g
        if args.mg_variant == "native":
            return multi_gpu.NativeKShardGemm(m, n, k_loc, N, fastmode=fast, dtype=dt, device=dev, backend=be)
        if be:
            raise SystemExit("--backend fp8 with --gpus N > 1: K-shard through the native driver only (--mg-variant native --mg-shard k)")
        return multi_gpu.KShardGemm(m, n, k_loc, N, fastmode=fast, dtype=dt, device=dev, variant=args.mg_variant)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, sampler=None, repeats=1):
        """`repeats` timed regions of `steps` calls each (CUDA events, max over ranks per region); returns the median region."""
        for _ in range(warmup):
            fn()
        res = []
        for r in range(repeats):
            barrier()
            if sampler is not None and r == 0:
                sampler.start()  # clocks are sampled DURING the timed region(s) only
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            if distributed:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            res.append(ms / steps)
        return statistics.median(res), res

    # ---- multi-GPU: verify the sharded path against the single-GPU call on a reduced problem BEFORE timing ----
    verify = None
    if distributed:
        from gemmul8_b200 import multi_gpu
        verify = multi_gpu.verify_against_single_gpu(shard, world, rank, dev, N, fast, variant=args.mg_variant, backend=be if shard == "k" else 0)

    A, B = make_inputs(k_local, shard)
    mg = make_mg(k_local, shard) if distributed else None
    out_elems = m * n if mg is None else mg.local_out_elems
    C = torch.zeros(m * n if mg is None else max(out_elems, 1), dtype=dt, device=dev)
    work = None
    if mg is None:
        tot, _, _ = g8.work_size(m, n, k_local, N, backend=be)
        work = torch.empty(tot, dtype=torch.uint8, device=dev)

    def step_device():
        if mg is not None:
            mg.run(A, B, C)
        else:
            g8.gemm("N", "N", m, n, k_local, 1.0, A, m, B, k_local, 0.0, C, m, N, fast, work, backend=be)

    # pinned host copies for the end-to-end leg
    hA = torch.empty(A.numel(), dtype=dt).pin_memory(); hA.copy_(A)
    hB = torch.empty(B.numel(), dtype=dt).pin_memory(); hB.copy_(B)
    hC = torch.empty(out_elems, dtype=dt).pin_memory()
    h2d_bytes, d2h_bytes = int(hA.numel() * 8 + hB.numel() * 8), int(hC.numel() * 8)
    # the repo's public host-buffer API (C ABI g8_gemm_host; csrc/g8_host.cu): both backends
    host_plan = g8.NativeHostGemm(m, n, k_local, dt, N, fast, "N", "N", chunk=1024, device=dev, backend=be) if mg is None else None

    def step_e2e():
        if host_plan is not None:
            # the repo's public host-buffer API: H2D of A and B, compute and D2H of C, pipelined over column chunks
            host_plan.run(hA, hB, hC, 1.0, 0.0, m, k_local, m)
            return
        A.copy_(hA, non_blocking=True)
        B.copy_(hB, non_blocking=True)
        step_device()
        hC.copy_(C if mg is None else mg.local_out(C), non_blocking=True)

    warmup = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    ms_dev, regions = timed(step_device, args.steps, warmup, sampler, repeats=3 if distributed else 1)
    clocks = sampler.stop()
    rank_clocks = None
    if distributed:
        allc = [None] * world
        dist.all_gather_object(allc, {"rank": rank, "sm_mhz": clocks.get("sm_mhz"), "power_w_median": clocks.get("power_w_median")})
        rank_clocks = allc
        if os.environ.get("G8_MG_TRACE") == "1":
            mine = [(n_, round(t_, 3)) for n_, t_ in mg.trace_report()]
            allr = [None] * world
            dist.all_gather_object(allr, mine)
            if rank == 0:
                print("[mg trace]", mine, file=sys.stderr)
                for key in ("gemm+scatter", "barrier", "sum+crt", "split"):
                    print(f"[mg trace all ranks] {key}:", [dict(t).get(key) for t in allr], file=sys.stderr)
    ms_e2e, _ = timed(step_e2e, max(2, min(args.steps, 5)), 1)
    h2d_rank_gbps = None
    if distributed:
        # what bounds the multi-GPU e2e leg: all ranks pull their slabs from host memory at the same time (per-rank H2D rate, concurrent)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            A.copy_(hA, non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        mine = 3 * hA.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) * 1e-9
        allg = [None] * world
        dist.all_gather_object(allg, round(mine, 1))
        h2d_rank_gbps = allg

    flops = 2.0 * m * n_total * k_total
    value = flops / (ms_dev * 1e-3) * 1e-12
    e2e_val = flops / (ms_e2e * 1e-3) * 1e-12

    # ---- roofline of the dominant kernel (the INT8 tcgen05 GEMM), measured with CUDA events inside g8_gemm ----
    roofline, phases, extra = None, None, {}
    if mg is None:
        ph = []
        for _ in range(max(3, min(args.steps, 10))):
            ph.append(g8.gemm("N", "N", m, n, k_local, 1.0, A, m, B, k_local, 0.0, C, m, N, fast, work, timing=True, backend=be))
        phases = [statistics.mean(p[i] for p in ph) for i in range(4)]
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
        which = "2 x measured bf16_tflops_sustained (MEASURED_PEAKS.json)" if peaks else "2 x fallback 1.4 PFLOP/s bf16 sustained"
        kp, mp = api.pad256(k_local), api.pad256(m)
        # the fused launch covers all moduli (FP8: three products per modulus); the bound GEMM of accurate mode is a separate launch in phase 0
        units = N if be == 0 else 3 * N
        ops = 2.0 * mp * n * kp * units
        t_gemm = phases[1] * 1e-9
        ach = ops / t_gemm * 1e-12
        traffic, traffic_src = ncu_traffic_bytes("gemm_i8_tc_kernel<0, 2>") if (be == 0 and S == 8192 and N == 14) else (None, None)
        roofline = {"bound": "tensor", "kernel": "gemm_i8_tc_kernel<EPI_MOD_I8> (tcgen05.mma kind::i8, all moduli in one launch)",
                    "achieved": round(ach, 1), "peak": round(2 * bf16, 1), "unit": "TFLOP/s",
                    "unit_note": "dense int8 (or fp8) tensor operations per second, counted like FLOPs (2 per multiply-add)",
                    "frac": round(ach / (2 * bf16), 4), "peak_source": which, "peak_nominal": 4500.0,
                    "frac_of_nominal": round(ach / 4500.0, 4), "traffic": traffic,
                    "traffic_source": f"dram__bytes_read.sum + dram__bytes_write.sum of this kernel, ncu --set full capture profiles/{traffic_src}",
                    "kernel_ms": round(t_gemm * 1e3, 4), "ops_per_launch": ops}
        if not args.no_extras and be == 0:
            # release the emulator's buffers while the library ceilings are measured (no interference with the numbers above)
            ceil8 = int8_ceiling_torch(torch, dev, stream)
            roofline["peak_int8_measured"] = ceil8
            if "sustained_tops" in ceil8:
                roofline["frac_of_int8_measured"] = round(ach / ceil8["sustained_tops"], 4)
                roofline["peak_int8_source"] = "cuBLASLt s8 8192^3 (torch._int_mm), sustained; the reference arm reports the same call through its own shim (extra.lt_ceiling)"
            # the reference's third axis: board power over a sustained loop, ours and native cublasDgemm (test_watt.hpp, test_flops.hpp:352-386)
            sync = lambda: torch.cuda.synchronize(dev)
            pw = {"ours_emulated_dgemm": power_leg(step_device, sync, flops, local_rank)}
            Am, Bm = A.view(k_local, m).t(), B.view(n, k_local).t()   # column-major m x k / k x n as torch views
            Cn = torch.empty((m, n), dtype=dt, device=dev)
            pw["native_cublas_dgemm"] = power_leg(lambda: torch.matmul(Am, Bm, out=Cn), sync, flops, local_rank)
            pw["how"] = "6 s back-to-back loop each, NVML board power every 100 ms, mean of the samples after 2 s (testing/getWatt.hpp, test_watt.hpp protocol, shortened from 10 s)"
            extra["power"] = pw
            del Cn
    elif shard == "k" and not args.no_extras and (args.k_local == 0 and args.size == 0):
        # the round-1 weak-K workload (8192 x 8192 x 8192 N) for continuity
        if hasattr(mg, "close"):
            mg.close()
        del mg, A, B, C, hA, hB, hC
        torch.cuda.empty_cache()
        m = n = S2 = 8192
        A, B = make_inputs(S2, "k")
        mg = make_mg(S2, "k")
        C = torch.zeros(mg.local_out_elems, dtype=dt, device=dev)
        ms_w, _ = timed(lambda: mg.run(A, B, C), max(5, args.steps // 2), 3)
        extra["weak_k"] = {"workload": f"DGEMM {S2}x{S2}x{S2 * world} K-sharded (k = {S2} per GPU)", "ms_per_step": round(ms_w, 4),
                           "value": round(2.0 * S2 * S2 * S2 * world / (ms_w * 1e-3) * 1e-12, 2), "unit": "TFLOPS"}
        m = n = S

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # our kernels per call -- single GPU: fast = stats, splitA, splitB(+stats), GEMM, CRT; accurate = stage (i) x 2, bound GEMM, 2 x finalize,
    # split x 2, GEMM, CRT.  K-sharded (fused): stats x2, shift x2, [bound planes x2, bound GEMM+scatter, maxabs, finalize x2], split x2,
    # GEMM+scatter, CRT (the 8-shard sum is inside the CRT kernel)
    if mg is None:
        launches_per_call = 5 if fast else 9
    else:
        launches_per_call = 8 if fast else 14
    out = {
        "metric": "emulated DGEMM TFLOPS @ N=8192 num_moduli=14; INT8 TC-pipe % of peak",
        "value": round(value, 2), "unit": "TFLOPS", "n_gpus": n_gpus, "steps": args.steps, "warmup": warmup,
        "ms_per_step": round(ms_dev, 4), "higher_is_better": True, "scaling": "weak" if not distributed or shard == "n" else "strong",
        "vs_baseline": None,
        "dtype": "int8 tensor-core residues (s8 x s8 -> s32) + f64 CRT; emulates f64", "data": "synthetic",
        "config": {"workload": f"DGEMM {m}x{n_total}x{k_total} {args.backend.upper()} num_moduli={N} fastmode={int(fast)} opN/opN alpha=1 beta=0"
                               + {"n": f", column-sharded over {world} GPUs (n={n} per GPU, A replicated, one all_reduce(MAX) in accurate mode)",
                                  "mod": f", modulus-set sharded over {world} GPUs (A, B replicated, all_gather of the int8 residue planes)",
                                  "k": f", K-sharded over {world} GPUs (k={k_local} per GPU; BASELINE.json config 3 at 8 GPUs), variant={args.mg_variant}",
                                  None: ""}[shard],
                   "inputs": "curand normal (phi=-1), seeds 12345/54321 as testing/make_matrix.hpp",
                   "l2": "inputs and residue planes are larger than the 126 MB L2; no explicit flush",
                   "timing": "CUDA events on the launch stream, max over ranks" + ("; median of 3 timed regions" if distributed else "")},
        "e2e": {"value": round(e2e_val, 2), "unit": "TFLOPS", "ms_per_step": round(ms_e2e, 3),
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                **({"h2d_gbps_per_rank_concurrent": h2d_rank_gbps, "note": "per-rank bytes; every rank copies its K-slabs from NUMA-local pinned memory at "
                    "the same time: the host side (PCIe root complexes / memory) bounds the aggregate"} if h2d_rank_gbps else {})},
        "gpu_launches": launches_per_call * args.steps * (3 if distributed else 1),
        "clocks": clocks,
        "impl": args.impl,
    }
    if distributed:
        out["timed_regions_ms_per_step"] = [round(x, 4) for x in regions]
        out["rank_clocks"] = rank_clocks
        out["verify"] = verify
    if phases is not None:
        out["phase_ms"] = {"split": round(phases[0] * 1e-6, 4), "gemm": round(phases[1] * 1e-6, 4),
                           "requant": round(phases[2] * 1e-6, 4), "crt": round(phases[3] * 1e-6, 4)}
    if roofline is not None:
        out["roofline"] = roofline
    if not args.no_cpu_baseline and world == 1:
        try:
            cb = cpu_baseline_oracle_port(N, fast)
        except Exception as e:  # the oracle is optional for the number itself
            cb = {"value": None, "unit": "TFLOPS", "cores": 1, "kind": "port", "sample": f"oracle port unavailable: {e}"}
        cb["host_openblas_dgemm"] = cpu_host_openblas()
        out["cpu_baseline"] = cb
    if extra:
        out["extra"] = extra
    print(json.dumps(out))
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
