#!/usr/bin/env python3
"""Latency of small emulated GEMMs: eager calls vs CUDA-graph replay (the launch-bound regime, e.g. BASELINE config 1, SGEMM 1024^3 N=6)."""
import sys, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gemmul8_b200 as g8

for (S, dt, N, fast) in ((1024, torch.float32, 6, True), (1024, torch.float32, 6, False), (512, torch.float64, 14, False), (2048, torch.float64, 14, False)):
    A = g8.randmat(S, S, dt, seed=1); B = g8.randmat(S, S, dt, seed=2)
    C = torch.zeros(S * S, dtype=dt, device="cuda")
    work = torch.empty(g8.work_size(S, S, S, N)[0], dtype=torch.uint8, device="cuda")
    call = lambda: g8.gemm("N", "N", S, S, S, 1.0, A, S, B, S, 0.0, C, S, N, fast, work)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        call()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        call()
    def timeit(fn, reps=200):
        for _ in range(20): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    te, tg = timeit(call), timeit(graph.replay)
    fl = 2.0 * S ** 3
    print(f"S={S} {str(dt).split('.')[-1]} N={N} {'fast' if fast else 'accu'}: eager {te:.1f} us ({fl / te * 1e-6:.1f} TFLOPS)  graph replay {tg:.1f} us ({fl / tg * 1e-6:.1f} TFLOPS)")
