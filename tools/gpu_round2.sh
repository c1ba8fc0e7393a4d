#!/bin/bash
# Round-2 single-GPU session: parity tests, both bench arms, ncu launch lists and one ncu --set full capture of the hot kernels.
# Outputs under gpurun_out/ (scratch); summaries are copied into profiles/ by tools/ncu_summary.py afterwards.
#   usage: tools/gpu_round2.sh <tag> [steps...]   steps: test bench ncu harness   (default: all but harness)
mkdir -p gpurun_out
tag=${1:-r02a}; shift
steps=${@:-test bench ncu}
has() { [[ " $steps " == *" $1 "* ]]; }
if has test; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.log
fi
if has bench; then
  timeout 600 python bench.py --steps 20 --warmup 5 --mode accu --impl reference 2> gpurun_out/${tag}_bench_ref_accu.err | tail -1 > gpurun_out/${tag}_bench_ref_accu.json
  timeout 600 python bench.py --steps 20 --warmup 5 --mode accu 2> gpurun_out/${tag}_bench_ours_accu.err | tail -1 > gpurun_out/${tag}_bench_ours_accu.json
  timeout 300 python bench.py --steps 20 --warmup 5 --mode fast --impl reference --no-extras 2> /dev/null | tail -1 > gpurun_out/${tag}_bench_ref_fast.json
  timeout 300 python bench.py --steps 20 --warmup 5 --mode fast --no-cpu-baseline --no-extras 2> gpurun_out/${tag}_bench_ours_fast.err | tail -1 > gpurun_out/${tag}_bench_ours_fast.json
  python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        j = json.loads(open(f).read())
        print(f.split("/")[-1], j.get("value"), "TFLOPS", j.get("ms_per_step"), "ms e2e", j.get("e2e", {}).get("value"), j.get("phase_ms"), (j.get("roofline") or {}).get("frac"), (j.get("clocks") or {}).get("sm_mhz"))
        if j.get("extra"): print("   extra:", json.dumps(j["extra"])[:900])
        if (j.get("roofline") or {}).get("peak_int8_measured"): print("   int8 ceiling:", j["roofline"]["peak_int8_measured"])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[:300]); print(open(f.replace(".json", ".err")).read()[-1500:] if __import__("os").path.exists(f.replace(".json", ".err")) else "")
PY
fi
if has ncu; then
  for mode in accu fast; do
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_${mode}_8192_N14.csv \
        python tools/profile_one.py 8192 14 $mode 2 > /dev/null 2>&1
  done
  # accurate mode: 9 launches per call; skip the first call, capture the second (all kernels of one call)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_i8_tc|split_row|crt_kernel|stats_row|accu_stage1|finalize" -s 9 -c 9 -f -o gpurun_out/${tag}_full_accu \
      python tools/profile_one.py 8192 14 accu 2 > gpurun_out/${tag}_ncu_full_accu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_i8_tc|split_row|crt_kernel|stats_row" -s 5 -c 5 -f -o gpurun_out/${tag}_full_fast \
      python tools/profile_one.py 8192 14 fast 2 > gpurun_out/${tag}_ncu_full_fast.log 2>&1
  # gpurun_out/ travels back only while it stays below 64 MiB: keep the raw metric pages (small CSV), drop the reports (50+ MB each)
  for r in accu fast; do
    ncu -i gpurun_out/${tag}_full_${r}.ncu-rep --page raw --csv > gpurun_out/${tag}_full_${r}_raw.csv 2>/dev/null
    rm -f gpurun_out/${tag}_full_${r}.ncu-rep
  done
fi
if has harness; then
  # the reference's own benchmark harness (testing/test.cu, unmodified), linked against the reference library and against this repo's libgemmul8.a
  ( cd gpurun_out && mkdir -p harness_ref harness_ours
    ( cd harness_ref && timeout 480 ../../oracle/_ref/harness_ref flops DGEMM INT8 > run.log 2>&1 )
    ( cd harness_ours && GEMMUL8_PHASE_TIMING=1 timeout 480 ../../oracle/_ref/harness_ours flops DGEMM INT8 > run.log 2>&1 )
    tail -3 harness_ref/run.log harness_ours/run.log )
fi
if has fp8stats; then
  timeout 600 python tools/fp8_shift_stats.py 8192 -1 1 2>&1 | tail -16
fi
if has sanitize; then
  timeout 900 python tools/sanitizer_run.py 2>&1 | tail -12
fi
ls -la gpurun_out | tail -15
