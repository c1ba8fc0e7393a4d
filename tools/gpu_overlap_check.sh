#!/bin/bash
# tests + stage sweep + bench with / without the two-stream overlap of the A / B side preprocessing
tag=${1:-r02e}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
bash tools/gpu_sweep.sh $tag
for ov in 1 0; do
  G8_OVERLAP_SIDES=$ov timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | tail -1 > gpurun_out/${tag}_bench_ours_accu_ov$ov.json
  G8_OVERLAP_SIDES=$ov timeout 300 python bench.py --steps 20 --warmup 5 --mode fast --no-cpu-baseline --no-extras 2>&1 | tail -1 > gpurun_out/${tag}_bench_ours_fast_ov$ov.json
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        j = json.load(open(f)); print(f, j["value"], j["ms_per_step"], j["phase_ms"], j["e2e"]["value"], j["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", open(f).read()[-600:])
PY
