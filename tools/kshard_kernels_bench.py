#!/usr/bin/env python3
"""Single-GPU micro-benchmark of the owner-side K-shard kernels at the headline size: residue_sum + crt  vs  crt_parts,
for world sizes 2 and 8 (parts live on this GPU; the arithmetic and the traffic are what a rank sees)."""
import sys, ctypes
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gemmul8_b200 import _lib, api

lib = _lib.load()
m = n_full = 8192; N = 14; mp = 8192
st = torch.cuda.current_stream().cuda_stream
keep = []
pa, pb = api._scalar_ptr(1.0, torch.float64, keep), api._scalar_ptr(0.0, torch.float64, keep)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10

def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for W in (2, 8):
    nc = n_full // W
    parts = torch.randint(-90, 91, (W, N, nc, mp), dtype=torch.int8, device="cuda")
    cmid = torch.zeros(N * nc * mp, dtype=torch.int8, device="cuda")
    C = torch.zeros(nc * m, dtype=torch.float64, device="cuda")
    sA = torch.full((mp,), -40, dtype=torch.int16, device="cuda"); sB = torch.full((8192,), -40, dtype=torch.int16, device="cuda")
    t_sum = timeit(lambda: lib.g8_stage_residue_sum(parts.data_ptr(), W, N * nc * mp, mp, nc, mp, nc * mp, N, 0, cmid.data_ptr(), mp, nc * mp, st))
    t_crt = timeit(lambda: lib.g8_stage_crt(1, cmid.data_ptr(), mp, nc * mp, m, nc, N, C.data_ptr(), m, sA.data_ptr(), sB.data_ptr(), pa, pb, st))
    t_fused = timeit(lambda: lib.g8_stage_crt_parts(1, parts.data_ptr(), W, N * nc * mp, mp, nc * mp, m, nc, N, C.data_ptr(), m, sA.data_ptr(), sB.data_ptr(), pa, pb, st))
    gb_sum = (W + 1) * N * nc * mp / 1e9
    print(f"W={W}: residue_sum {t_sum:.3f} ms ({gb_sum / t_sum:.0f} GB/s)  crt {t_crt:.3f} ms  sum+crt {t_sum + t_crt:.3f} ms | crt_parts {t_fused:.3f} ms")
