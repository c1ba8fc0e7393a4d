#!/bin/bash
# multi-GPU session: NCCL parity test, then the K-sharded bench for the exchange variants (phase trace on rank 0)
N=${1:-2}
mkdir -p gpurun_out
[ -z "$SKIP_TEST" ] && timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
for mode in ${MODES:-accu fast}; do
  for v in ${VARIANTS:-residue fused}; do
    echo "== N=$N $mode $v"
    G8_MG_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 --mode $mode --mg-variant $v ${BENCH_EXTRA} 2> gpurun_out/mg_${N}_${mode}_${v}.err | tail -1 > gpurun_out/mg_${N}_${mode}_${v}.json
    grep "mg trace" gpurun_out/mg_${N}_${mode}_${v}.err | tail -5
    python -c "
import json,sys
j=json.loads(open('gpurun_out/mg_${N}_${mode}_${v}.json').read()); print(j['value'],'TFLOPS',j['ms_per_step'],'ms e2e',j['e2e']['value'])" || tail -5 gpurun_out/mg_${N}_${mode}_${v}.err
  done
done
