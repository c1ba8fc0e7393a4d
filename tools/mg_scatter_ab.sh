#!/bin/bash
# 8-GPU A/B of the GEMM -> scatter epilogue: TMA tensor store (default library) vs the previous per-thread bulk copies (variant library)
# The variant is NOT kept in the tree: build it from the commit before the change with
#   git show 7adef48:gemmul8_b200/csrc/g8_gemm_i8.cu > /tmp/old.cu   (compile like csrc/Makefile, link the other objects of build/obj)
# into gemmul8_b200/lib/variants/libg8core_old.so; GEMMUL8_B200_LIB selects the library the Python layer loads.
N=${1:-8}
mkdir -p gpurun_out
run() { # name lib args...
  name=$1; lib=$2; shift 2
  echo "== N=$N $name: $*"
  GEMMUL8_B200_LIB=$lib G8_MG_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-extras "$@" 2> gpurun_out/ab_${N}_${name}.err | tail -1 > gpurun_out/ab_${N}_${name}.json
  grep "mg trace" gpurun_out/ab_${N}_${name}.err | tail -3
  python - <<PY || tail -8 gpurun_out/ab_${N}_${name}.err
import json
j = json.loads(open("gpurun_out/ab_${N}_${name}.json").read())
print("   ", j["value"], "TFLOPS", j["ms_per_step"], "ms  regions", j.get("timed_regions_ms_per_step"), " verify", (j.get("verify") or {}).get("bit_identical_all_ranks"))
PY
}
NEW=$PWD/gemmul8_b200/lib/libg8core.so; OLD=$PWD/gemmul8_b200/lib/variants/libg8core_old.so
run fast_fused_tma  $NEW --mode fast --mg-variant fused
run fast_fused_bulk $OLD --mode fast --mg-variant fused
run accu_native_tma  $NEW --mode accu --mg-variant native
run accu_native_bulk $OLD --mode accu --mg-variant native
