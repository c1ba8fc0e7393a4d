#!/usr/bin/env python3
"""The all-moduli INT8 GEMM of DGEMM 8192^3 N=14 alone, back to back for 3 s: ms per launch (CUDA events), median SM clock and
mean board power (NVML).  The rasterisation band width comes from G8_GEMM_GROUP (read once by the library)."""
import json, os, sys, time, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gemmul8_b200 as g8
from gemmul8_b200 import _lib, api
from bench import ClockSampler
lib = _lib.load()
S, N = 8192, 14
kp = mp = S
A_lo = torch.randint(-127, 128, (N * kp * mp,), dtype=torch.int8, device="cuda")
B_lo = torch.randint(-127, 128, (N * kp * S,), dtype=torch.int8, device="cuda")
C_mid = torch.empty(N * mp * S, dtype=torch.int8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
call = lambda: api._check(lib.g8_stage_gemm(0, 0, A_lo.data_ptr(), kp * mp, B_lo.data_ptr(), kp * S, S, S, kp, N, 0, None, None, C_mid.data_ptr(), mp * S, mp, None, None, st), "gemm")
for _ in range(5): call()
torch.cuda.synchronize()
smp = ClockSampler(0, period=0.05); smp.start()
t0 = time.perf_counter(); calls = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.perf_counter() - t0 < 3.0:
    for _ in range(20): call()
    calls += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
smp._stop.set(); smp.th.join()
rows = [r for r in smp.rows if r[3] - t0 > 1.0]
print(json.dumps({"group": os.environ.get("G8_GEMM_GROUP", "16"), "ms_per_launch": round(e0.elapsed_time(e1) / calls, 4), "sm_mhz_median": statistics.median(r[0] for r in rows),
                  "power_w_mean": round(statistics.mean(r[1] for r in rows), 1), "calls": calls}))
