run() { # name cores lanes blocking
  BZ_BLOCKING_SYNC=$4 taskset -c 0-$(($2-1)) python bench.py --no-extras --inflight $3 > gpurun_out/sync_$1.log 2>&1
}
run c4_l4_spin 4 4 0
run c4_l6_spin 4 6 0
run c4_l6_block 4 6 1
run c4_l8_block 4 8 1
run c16_l6_block 16 6 1
run c16_l6_spin 16 6 0
taskset -c 0-3 python bench.py --no-extras > gpurun_out/sync_c4_auto.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sync_*.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            ok=True; d=json.loads(l); print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['config'].get('host_wait'), d['config']['workload'][40:70], d.get('single_proof_ms'))
    if not ok: print(f, open(f).read()[-400:])
PY
