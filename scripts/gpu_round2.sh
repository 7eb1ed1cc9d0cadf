#!/bin/bash
# Round-2 GPU visit: launch list of the step (ncu), stock-GPU arm, other BASELINE configs.   usage: gpu_round2.sh <tag> <stages>
TAG=${1:-r02x}; ST=${2:-lsc}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONUNBUFFERED=1
if [[ $ST == *l* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-range > $OUT/bench_under_ncu.log 2>&1
  echo "launch list rc=$?"; wc -l $OUT/launches.csv; python scripts/launch_summary.py $OUT/launches.csv 2 > $OUT/launches.txt; head -50 $OUT/launches.txt
fi
if [[ $ST == *s* ]]; then
  timeout 600 python bench.py --impl stock-gpu --steps 5 --warmup 3 > $OUT/stock_gpu.json 2> $OUT/stock_gpu.err; echo "stock rc=$?"; cat $OUT/stock_gpu.json; tail -3 $OUT/stock_gpu.err
fi
if [[ $ST == *c* ]]; then
  for c in 2 3 4; do
    timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err; echo "cfg$c rc=$?"
    head -c 400 $OUT/bench_cfg$c.json; echo; tail -3 $OUT/bench_cfg$c.err
  done
fi
if [[ $ST == *b* ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 400 $OUT/bench.json; tail -4 $OUT/bench.err
fi
if [[ $ST == *r* ]]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
fi
