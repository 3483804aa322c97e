# launch list of one scaled proof with every commitment on the bucket-MSM path (what a k = 20 proof runs), at k = 18 to keep it short
mkdir -p gpurun_out
export BZ_NO_CPU_BASELINE=1 BZ_FORCE_GENERAL_MSM=1
CMD="python bench.py --workload board_scaled --k 18 --steps 1 --warmup 1"
timeout 600 $CMD > gpurun_out/k18g_plain.log 2>&1 || { tail -5 gpurun_out/k18g_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/k18g_launches.csv $CMD > gpurun_out/k18g_ncu.log 2>&1
python profiles/summarize.py --launches gpurun_out/k18g_launches.csv | head -30
python - <<'PY'
import json
for l in open('gpurun_out/k18g_plain.log'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],1), d['roofline']['kernel_ms'], d['verified'])
PY
