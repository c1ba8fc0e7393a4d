#!/usr/bin/env python3
"""One warm-up + a few emulated DGEMM calls, for ncu (tools/ncu recipes in profiles/README.md)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gemmul8_b200 as g8

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14
fast = (sys.argv[3] == "fast") if len(sys.argv) > 3 else False
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
be = 1 if (len(sys.argv) > 5 and sys.argv[5] == "fp8") else 0
dt = torch.float64
A = g8.randmat(S, S, dt, seed=12345)
B = g8.randmat(S, S, dt, seed=54321)
C = torch.zeros(S * S, dtype=dt, device="cuda")
tot, _, _ = g8.work_size(S, S, S, N, backend=be)
work = torch.empty(tot, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    g8.gemm("N", "N", S, S, S, 1.0, A, S, B, S, 0.0, C, S, N, fast, work, backend=be)
torch.cuda.synchronize()
print("done")
