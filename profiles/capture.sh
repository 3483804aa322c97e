#!/bin/bash
# Reproduces the ncu evidence under profiles/ (run on a B200 box through gpurun; never a bench number).
#   bash profiles/capture.sh <tag>      -> gpurun_out/<tag>_*.{csv,ncu-rep,log}
# Every ncu pass runs only after the same command has exited 0 without ncu.
set -u
# NOTE: consider `export BZ_FIXED_WINDOW=14` for the --set full passes: ncu saves / restores all allocated device memory between
# replay passes, and the default c = 16 tables are 138 GB (r1h took 21 minutes instead of 7).
TAG=${1:-r2}
export BZ_FIXED_WINDOW=${BZ_FIXED_WINDOW:-14}      # 35 GB of tables instead of 138: ncu saves / restores device memory between replay passes
OUT=gpurun_out
CMD="python bench.py --no-extras --batch 64 --inflight 1 --steps 2 --warmup 3"
mkdir -p $OUT
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
# 1. launch list: every kernel with its device time (cold cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
# 2. --set full captures of the dominant kernels (steady state: skip the first launches)
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $OUT/${TAG}_$1 $CMD > $OUT/${TAG}_ncu_$1.log 2>&1
  ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>/dev/null
}
cap fb_accumulate fb_accumulate_kernel 25 6
cap quotient "q_shot_t0|eval_program_kernel" 5 3          # h(X): generated straight-line kernel (round 2) / interpreter
cap ntt ntt_pass 40 6
cap fb_fold fb_fold_kernel 25 3
cap gp_finish grand_product_finish_batch_kernel 1 1
ls -la $OUT | grep ${TAG}_
