#!/bin/bash
# BASELINE config 4: raw MSM (Pallas and Vesta) and Fp NTT size sweeps on one B200 -> one bench.py JSON line per size.
#   bash profiles/sweep.sh <tag>   ->  gpurun_out/<tag>_sweep.jsonl
TAG=${1:-r1}
OUT=gpurun_out/${TAG}_sweep.jsonl
: > $OUT
for L in 12 14 16 18 20 22 24; do
  python bench.py --workload msm --curve 1 --log2n $L --steps 3 --warmup 3 --cpu-sample-log 16 2>/dev/null | tail -1 >> $OUT
  python bench.py --workload ntt --log2n $L --steps 3 --warmup 3 --cpu-sample-log 16 2>/dev/null | tail -1 >> $OUT
done
for L in 16 20; do python bench.py --workload msm --curve 0 --log2n $L --steps 3 --warmup 3 --cpu-sample-log 16 2>/dev/null | tail -1 >> $OUT; done
wc -l $OUT
