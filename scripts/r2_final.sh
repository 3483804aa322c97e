# round-2 final single-GPU evidence: GPU suite, smoke, both bench arms, the other headline configs, the size sweeps
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_ref_shot.log 2>&1
timeout 900 python bench.py --impl reference --workload board --steps 3 --warmup 1 > gpurun_out/final_ref_board.log 2>&1
timeout 1500 python bench.py > gpurun_out/final_bench.log 2>&1
timeout 900 python bench.py --workload board --no-extras > gpurun_out/final_board.log 2>&1
timeout 1500 bash profiles/sweep.sh r2 2>&1 | tail -2
python - <<'PY'
import json
def lines(f):
    return [json.loads(l) for l in open(f) if l.startswith('{')]
for f in ['gpurun_out/final_ref_shot.log','gpurun_out/final_ref_board.log']:
    for d in lines(f): print(f, round(d['value'],3), d['unit'], 'runs', d.get('runs'), d.get('value_min_max'), 'cores', d['cpu_baseline']['cores'], 'one_thread', d.get('one_thread',{}).get('value'), d.get('cpu_native'))
for f in ['gpurun_out/final_bench.log','gpurun_out/final_board.log']:
    for d in lines(f):
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single', d.get('single_proof_ms'), d['roofline']['kernel_ms'], d['int_pipe']['frac_of_imad_peak'], d['roofline']['frac'], d['cpu_baseline'] and d['cpu_baseline']['value'], d.get('setup'), d['clocks'])
        for k,v in (d.get('extras') or {}).items(): print('  ', k, v.get('metric'), round(v.get('value',0),2), 'e2e', round(v['e2e']['value'],2) if v.get('e2e') else None, v.get('verified'), v.get('single_proof_ms'))
for d in lines('gpurun_out/r2_sweep.jsonl'):
    print(d['config']['workload'][:40], round(d['value']/ (1e6 if 'msm' in d['metric'] else 1),1), 'e2e', round(d['e2e']['value']/(1e6 if 'msm' in d['metric'] else 1),1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']/(1e6 if 'msm' in d['metric'] else 1),3))
PY
