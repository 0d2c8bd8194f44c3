#!/bin/bash
# A/B a tuning environment variable on the GPU box: tools/ab_env.sh VAR "v1 v2 ..." -> gpurun_out/ab_<VAR>.log
VAR=$1; VALS=$2
mkdir -p gpurun_out
for v in $VALS; do
  echo "=== $VAR=$v  C2 (5 Mbp x 50, 0.1% err, erate 0.01)"
  env $VAR=$v python tools/stage_timing.py 5e6 50 0.001 0.01 2>&1 | tail -4
  echo "=== $VAR=$v  noisy (2 Mbp x 40, 3% err, erate 0.06, 10-20 kb)"
  env $VAR=$v python tools/stage_timing.py 2e6 40 0.03 0.06 uniform 10000 20000 2>&1 | tail -4
done 2>&1 | tee gpurun_out/ab_$VAR.log
