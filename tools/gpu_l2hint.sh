#!/bin/bash
# GEMM L2 cache-hint experiment: G8_GEMM_L2HINT bits (1 band evict_last, 2 panels evict_first, 4 streaming C_mid stores) x band width
tag=${1:-r02h}
mkdir -p gpurun_out
for cfg in "0 16" "4 16" "1 16" "5 16" "7 16" "5 32" "7 32" "5 8"; do
  set -- $cfg
  echo "== G8_GEMM_L2HINT=$1 G8_GEMM_GROUP=$2"
  G8_GEMM_L2HINT=$1 G8_GEMM_GROUP=$2 timeout 200 python tools/gemm_band_probe.py 2>&1 | tail -1 | sed "s/^{/{\"l2hint\": $1, /" | tee -a gpurun_out/${tag}_gemm_l2hint.jsonl
  G8_GEMM_L2HINT=$1 G8_GEMM_GROUP=$2 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum \
      --clock-control none -k regex:gemm_i8_tc -s 1 -c 1 --csv python tools/profile_one.py 8192 14 fast 2 2>/dev/null | grep -E "gemm_i8_tc" | awk -F'","' -v g="hint=$1 group=$2" '{print g, $(NF-2), $(NF-1), $NF}' | tee -a gpurun_out/${tag}_gemm_l2hint_ncu.txt
done
# end to end with the best-looking settings
for h in 0 5 7; do
  G8_GEMM_L2HINT=$h timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('bench accu l2hint=$h', j['value'], j['ms_per_step'], j['phase_ms'])"
done
