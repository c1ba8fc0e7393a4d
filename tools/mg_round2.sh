#!/bin/bash
# multi-GPU session (N = 2, 4 or 8 GPUs of one box): hardware parity tests on all GPUs, then the K-sharded bench
# (BASELINE.json config 3 sizes: DGEMM 16384^3, k = 16384 / N per GPU) for the exchange variants, plus the N-shard and modulus-shard modes.
#   usage: tools/mg_round2.sh N [tag] [steps...]    steps: test bench alt
N=${1:-2}; tag=${2:-r02mg}; shift 2
steps=${@:-test bench alt}
has() { [[ " $steps " == *" $1 "* ]]; }
mkdir -p gpurun_out
run() { # name, extra bench args...
  name=$1; shift
  echo "== N=$N $name: $*"
  G8_MG_TRACE=${TRACE:-0} timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 "$@" 2> gpurun_out/${tag}_${N}_${name}.err | tail -1 > gpurun_out/${tag}_${N}_${name}.json
  grep "mg trace" gpurun_out/${tag}_${N}_${name}.err | tail -6
  python - <<PY || tail -8 gpurun_out/${tag}_${N}_${name}.err
import json
j = json.loads(open("gpurun_out/${tag}_${N}_${name}.json").read())
print("   ", j["value"], "TFLOPS", j["ms_per_step"], "ms  regions", j.get("timed_regions_ms_per_step"), " e2e", j["e2e"]["value"], " verify", (j.get("verify") or {}).get("bit_identical_all_ranks"), (j.get("extra") or {}).get("weak_k"))
PY
}
if has test; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_${N}_pytest.log
fi
if has bench; then
  TRACE=1 run accu_fused --mode accu --mg-variant fused
  TRACE=1 run fast_fused --mode fast --mg-variant fused --no-extras
  run accu_native --mode accu --mg-variant native --no-extras
  run fast_native --mode fast --mg-variant native --no-extras
fi
if has confirm; then
  run default
  TRACE=1 run accu_fused --mode accu --mg-variant fused --no-extras
  run fast_native --mode fast --no-extras
fi
if has alt; then
  G8_MG_BOUND=int32 run accu_fused_boundint32 --mode accu --mg-variant fused --no-extras
  G8_MG_SUM_IN_CRT=0 run accu_fused_sumpass --mode accu --mg-variant fused --no-extras
  run accu_nshard --mode accu --mg-shard n --size 8192 --no-extras
  run accu_modshard --mode accu --mg-shard mod --size 8192 --no-extras
fi
ls gpurun_out | grep ${tag}_${N} | head -30
