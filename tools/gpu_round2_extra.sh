#!/bin/bash
# smoke(), compute-sanitizer on the round-2 paths, and the GEMM rasterisation-band sweep with DRAM bytes / clocks (VERDICT task 4 evidence)
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
{
  for tool in memcheck synccheck; do
    echo "== compute-sanitizer --tool $tool tools/sanitizer_run.py"
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitizer_run.py 2>&1 | grep -E "ERROR SUMMARY|done|Invalid|Error|relerr" | tail -14
    echo "== compute-sanitizer --tool $tool tools/sanitizer_extra.py"
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitizer_extra.py 2>&1 | grep -E "ERROR SUMMARY|done|Invalid|Error|bit-identical|maxdiff" | tail -20
  done
} > gpurun_out/${tag}_sanitizer.txt 2>&1
tail -12 gpurun_out/${tag}_sanitizer.txt
# band sweep: time (CUDA events, inside stage_bench), DRAM bytes (ncu), SM clock / power (NVML during a 3 s loop)
for grp in 4 8 16 32; do
  echo "== G8_GEMM_GROUP=$grp"
  G8_GEMM_GROUP=$grp timeout 200 python tools/gemm_band_probe.py 2>&1 | tail -1 | tee -a gpurun_out/${tag}_gemm_band.jsonl
  G8_GEMM_GROUP=$grp timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum \
      --clock-control none -k regex:gemm_i8_tc -s 1 -c 1 --csv python tools/profile_one.py 8192 14 fast 2 2>/dev/null | grep -E "gemm_i8_tc" | awk -F'","' -v g=$grp '{print "group="g, $(NF-2), $(NF-1), $NF}' | tee -a gpurun_out/${tag}_gemm_band_ncu.txt
done
