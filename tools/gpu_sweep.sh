#!/bin/bash
# kernel-variant sweep on one GPU (each variant = one short process); output: gpurun_out/<tag>_sweep.jsonl
tag=${1:-r02b}
mkdir -p gpurun_out
run() { echo "== $*" >&2; env "$@" timeout 300 python tools/stage_bench.py 8192 14 20 2>&1 | tail -1 >> gpurun_out/${tag}_sweep.jsonl; }
run G8_X=default
run G8_OVERLAP_SIDES=0
run G8_SPLIT_TILES_PER_BLOCK=1
python - <<PY
import json
for ln in open("gpurun_out/${tag}_sweep.jsonl"):
    try:
        j = json.loads(ln)
        print(j["env"], {k: v for k, v in j.items() if k not in ("env", "S", "N")})
    except Exception as e:
        print("ERR", ln[:300])
PY
