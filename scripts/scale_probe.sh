#!/bin/bash
# multi-GPU diagnosis: step time with / without the gradient all-reduce, and with SMs left to NCCL.  usage: scale_probe.sh <tag> <N>
TAG=${1:-scale}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() {  # name, env, extra args
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline $3 > $OUT/$1.json 2> $OUT/$1.err
  echo "$1 rc=$?"; python -c "
import json,sys
d=json.loads([l for l in open('$OUT/$1.json') if l.startswith('{')][-1]); print('$1', 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2))" 2>/dev/null || tail -3 $OUT/$1.err
}
for m in ${MARGINS:-8 16}; do run n${N}_margin$m "ODWSCL_SM_MARGIN=$m" ""; done
# VARIANTS: "name:ENV=v,ENV=v ..." (e.g. late:ODWSCL_SM_MARGIN=0,ODWSCL_BUCKET_MB=4096)
for v in ${VARIANTS:-}; do run n${N}_${v%%:*} "$(echo ${v#*:} | tr ',' ' ')" ""; done
for f in $OUT/n${N}_*.err; do echo $f; grep "resident per-step" $f | cut -c1-160; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/n1.json 2> $OUT/n1.err; python -c "
import json; d=json.loads([l for l in open('$OUT/n1.json') if l.startswith('{')][-1]); print('n1 ms/step', round(d['ms_per_step'],2), 'value', round(d['value']))"
