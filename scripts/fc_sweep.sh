#!/bin/bash
# GPU box: schedule knobs of the fc GEMM on the big shapes + one full ncu capture (ours vs cuBLAS at the fc7 shape)
OUT=gpurun_out/${1:-fcsweep}
mkdir -p $OUT
SH="fc6.fwd,fc7.fwd,sim0.fwd,fc6.dgrad,fc6.wgrad,sim0.wgrad"
for st in 6 7; do for pn in 4 8 16; do
  echo "== stages $st panel $pn" >> $OUT/sweep.txt
  ODWSCL_FC_STAGES=$st ODWSCL_FC_PANEL=$pn timeout 200 python scripts/bench_fc.py 4000 384 $SH 2>&1 | grep gemm | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-10s ours %.3f ms %4.0f TF/s | cublas %.3f ms %4.0f TF/s'%(d['gemm'],d['ms'],d['tflops'],d['cublas_ms'],d['cublas_tflops']))" >> $OUT/sweep.txt
done; done
cat $OUT/sweep.txt
FC_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fc_gemm_tf32_2cta|cutlass" -s 2 -c 2 \
   -o $OUT/prof_fc7 -f python scripts/bench_fc.py 4000 384 fc7.fwd > $OUT/ncu_fc7.log 2>&1
ls -la $OUT
