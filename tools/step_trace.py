#!/usr/bin/env python3
"""Per-call latency trace of the headline DGEMM: back-to-back (sustained) and with idle gaps (burst).  Diagnostic for the
power-cap behaviour: usage  step_trace.py [fast|accu] [steps]"""
import sys, time, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gemmul8_b200 as g8

fast = (sys.argv[1] == "fast") if len(sys.argv) > 1 else True
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 120
S, N = 8192, 14
dt = torch.float64
A = g8.randmat(S, S, dt, seed=12345); B = g8.randmat(S, S, dt, seed=54321)
C = torch.zeros(S * S, dtype=dt, device="cuda")
tot, _, _ = g8.work_size(S, S, S, N)
work = torch.empty(tot, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
def call():
    g8.gemm("N", "N", S, S, S, 1.0, A, S, B, S, 0.0, C, S, N, fast, work)
for _ in range(3): call()
torch.cuda.synchronize(); time.sleep(1.0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record(st)
for i in range(steps):
    call(); ev[i + 1].record(st)
torch.cuda.synchronize()
t = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
print("sustained ms per call, groups of 10:", [round(statistics.mean(t[i:i + 10]), 3) for i in range(0, steps, 10)])
time.sleep(1.0)
tb = []
for i in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); call(); b.record(st); torch.cuda.synchronize()
    tb.append(a.elapsed_time(b)); time.sleep(0.05)
print("burst (50 ms idle between calls) ms:", round(statistics.median(tb), 3), "min", round(min(tb), 3))
