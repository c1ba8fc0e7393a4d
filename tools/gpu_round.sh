#!/bin/bash
# One GPU session: parity tests, both bench arms, ncu launch lists and one ncu --set full capture of the hot kernels.
# Outputs under gpurun_out/ (scratch); summaries are copied into profiles/ by tools/ncu_summary.py afterwards.
mkdir -p gpurun_out
tag=${1:-r01e}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest.log
for mode in accu fast; do
  python bench.py --steps 30 --warmup 5 --mode $mode --impl reference 2>&1 | tail -1 > gpurun_out/${tag}_bench_ref_${mode}.json
  python bench.py --steps 30 --warmup 5 --mode $mode $( [ $mode = fast ] && echo --no-cpu-baseline ) 2>&1 | tail -1 > gpurun_out/${tag}_bench_ours_${mode}.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/*_bench_*.json")):
    try:
        j = json.loads(open(f).read())
        print(f.split("/")[-1], j.get("value"), "TFLOPS", j.get("ms_per_step"), "ms e2e", j.get("e2e", {}).get("value"), j.get("phase_ms"), (j.get("roofline") or {}).get("frac"), j.get("clocks"))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[:300])
PY
for mode in accu fast; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_${mode}_8192_N14.csv \
      python tools/profile_one.py 8192 14 $mode 2 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"gemm_i8_tc|split_row|crt_kernel|stats_row" -s 5 -c 5 -f -o gpurun_out/${tag}_full_fast \
    python tools/profile_one.py 8192 14 fast 2 > gpurun_out/${tag}_ncu_full.log 2>&1
# FP8 backend (config 5): the per-product GEMM, the combine pass and the FP8 split
ncu --set full --clock-control none --import-source on -k regex:"gemm_i8_tc|f8_combine|split_row" -s 4 -c 4 -f -o gpurun_out/${tag}_full_fp8 \
    python tools/profile_one.py 8192 14 fast 2 fp8 > gpurun_out/${tag}_ncu_full_fp8.log 2>&1
ls -la gpurun_out | tail -12
