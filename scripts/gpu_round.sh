#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the top kernel.
# Usage (under gpurun): bash scripts/gpu_round.sh [tag] [stages]   stages: any of t,s,b,l,n (default all)
TAG=${1:-r01}
ST=${2:-tsbln}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" > $OUT/env.txt 2>&1
if [[ $ST == *t* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q -rA > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -40 $OUT/pytest_gpu.log
fi
if [[ $ST == *s* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
  tail -5 $OUT/smoke.log
fi
if [[ $ST == *b* ]]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
  cat $OUT/bench.json; tail -5 $OUT/bench.err
fi
if [[ $ST == *l* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-range > $OUT/bench_under_ncu.log 2>&1
  echo "launch list rc=$?"; wc -l $OUT/launches.csv
fi
if [[ $ST == *n* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"roi_pool_fwd_nhwc7|roi_pool_bwd" -s 6 -c 2 \
      -o $OUT/prof_roipool -f python scripts/run_roipool_once.py > $OUT/ncu_full.log 2>&1
  echo "ncu full rc=$?"; ls -la $OUT
fi
if [[ $ST == *v* ]]; then
  timeout 600 python scripts/bench_conv.py 2 > $OUT/bench_conv.jsonl 2>&1; echo "bench_conv rc=$?"; cat $OUT/bench_conv.jsonl
fi
if [[ $ST == *c* ]]; then
  # full capture of the tcgen05 conv kernel: launches 4.. = timed conv1_2, then skip to conv4_2-shaped ones
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tf32_kernel" -s 3 -c 1 \
      -o $OUT/prof_conv_a -f python scripts/bench_conv.py 2 > $OUT/ncu_conv.log 2>&1
  # every bench_conv layer runs the persistent CTA-pair kernel (13 launches each): launch 84 = a timed conv4_2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tf32_2cta_sk_kernel" -s 84 -c 1 \
      -o $OUT/prof_conv_b -f python scripts/bench_conv.py 2 >> $OUT/ncu_conv.log 2>&1
  echo "ncu conv rc=$?"; ls -la $OUT
fi
if [[ $ST == *w* ]]; then
  # the CTA-pair WGRAD at the conv4_2 / conv5_x shape
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_wgrad_tf32_2cta_kernel" -s 3 -c 1 \
      -o $OUT/prof_wgrad -f python scripts/run_wgrad_once.py > $OUT/ncu_wgrad.log 2>&1
  echo "ncu wgrad rc=$?"
fi
