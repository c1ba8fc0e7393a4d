#!/usr/bin/env python3
"""Single-GPU probe: is the fused GEMM -> scatter epilogue limited by its own issue rate (one 256-byte bulk copy per thread per tile) or by
NVLink?  Runs the scatter kernel of one rank of a W-way K-shard of the 16384^3 case with ALL receive areas on this GPU (no NVLink in the
picture) beside the plain mod-p GEMM of the same shape.  If local scatter == NVLink scatter time, the epilogue itself is the limit."""
import sys, ctypes
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gemmul8_b200 import _lib

lib = _lib.load()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
N = 14
st = torch.cuda.current_stream().cuda_stream


def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for W in (2, 4, 8):
    kp = S // W; m = n = S; nc = n // W
    A = torch.randint(-127, 128, (N, m, kp), dtype=torch.int8, device="cuda")
    B = torch.randint(-127, 128, (N, n, kp), dtype=torch.int8, device="cuda")
    out = torch.empty(N * n * m, dtype=torch.int8, device="cuda")           # plain: N x (n x m)
    recv = torch.empty(W, W * N * nc * m, dtype=torch.int8, device="cuda")   # every "peer" holds W parts of N x (nc x m)
    ptrs = (ctypes.c_void_p * W)(*[recv[o].data_ptr() for o in range(W)])

    def plain():
        rc = lib.g8_stage_gemm(0, 0, A.data_ptr(), m * kp, B.data_ptr(), n * kp, m, n, kp, N, 0, None, None, out.data_ptr(), n * m, m, None, None, st)
        assert rc == 0, rc

    def scatter():
        rc = lib.g8_stage_gemm_scatter(0, A.data_ptr(), m * kp, B.data_ptr(), n * kp, m, n, kp, N, 0, ptrs, W, 0, nc * m, m, st)
        assert rc == 0, rc

    # the GPU is power-capped on dense random planes: interleave the two kernels call by call and take medians, otherwise the order decides
    for _ in range(3): plain(); scatter()
    torch.cuda.synchronize()
    tps, tss = [], []
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(); plain(); ev[1].record(); scatter(); ev[2].record()
        torch.cuda.synchronize()
        tps.append(ev[0].elapsed_time(ev[1])); tss.append(ev[1].elapsed_time(ev[2]))
    tp, ts = sorted(tps)[len(tps) // 2], sorted(tss)[len(tss) // 2]
    same = all(torch.equal(out.view(N, n, m)[:, o * nc:(o + 1) * nc], recv[o].view(W, N, nc, m)[0]) for o in range(W))
    ops = 2.0 * m * n * kp * N
    print(f"W={W} k_local={kp}: plain {tp:.3f} ms ({ops / tp / 1e12:.2f} POP/s)  local-scatter {ts:.3f} ms ({ops / ts / 1e12:.2f} POP/s)  ratio {ts / tp:.3f}  identical={same}", flush=True)
    del A, B, out, recv
