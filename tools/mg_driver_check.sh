#!/bin/bash
# exactly what the driver runs for the scaling bench at N GPUs: both arms under torchrun, default flags
N=${1:-4}; tag=${2:-r02drv}
mkdir -p gpurun_out
for impl in reference ours; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl $impl --gpus $N --steps 10 --warmup 3 \
     2> gpurun_out/${tag}_${N}_${impl}.err | tail -1 > gpurun_out/${tag}_${N}_${impl}.json
  python - <<PY || tail -12 gpurun_out/${tag}_${N}_${impl}.err
import json
j = json.loads(open("gpurun_out/${tag}_${N}_${impl}.json").read())
print("$impl", j.get("value"), j.get("unit"), j.get("ms_per_step"), "ms | e2e", (j.get("e2e") or {}).get("value"), "|", j["config"]["workload"][:110], "| verify", (j.get("verify") or {}).get("bit_identical_all_ranks"), "| extra", list((j.get("extra") or {}).keys()))
PY
done
