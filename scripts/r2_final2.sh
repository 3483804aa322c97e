# round-2 final single-GPU evidence after the last prover changes: GPU suite, smoke, scaled proofs, both headline benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
export BZ_NO_CPU_BASELINE=1
for k in 16 18 20; do
timeout 600 python bench.py --workload board_scaled --k $k --steps 2 --warmup 3 > gpurun_out/scaled${k}_n1.log 2>&1
done
unset BZ_NO_CPU_BASELINE
timeout 1500 python bench.py > gpurun_out/final_bench.log 2>&1
timeout 900 python bench.py --workload board --no-extras > gpurun_out/final_board.log 2>&1
python - <<'PY'
import json
def lines(f):
    return [json.loads(l) for l in open(f) if l.startswith('{')]
for f in ['gpurun_out/scaled16_n1.log', 'gpurun_out/scaled18_n1.log', 'gpurun_out/scaled20_n1.log']:
    for d in lines(f): print(f, round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), d['roofline']['kernel_ms'] if d['roofline'] else None, d['verified'], d['single_proof_ms'])
for f in ['gpurun_out/final_bench.log','gpurun_out/final_board.log']:
    for d in lines(f):
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single', d.get('single_proof_ms'), d['roofline']['kernel_ms'], d['int_pipe']['frac_of_imad_peak'], d['roofline']['frac'], d['cpu_baseline'] and d['cpu_baseline']['value'], d['clocks'])
        for k,v in (d.get('extras') or {}).items(): print('  ', k, v.get('metric'), round(v.get('value',0),2), 'e2e', round(v['e2e']['value'],2) if v.get('e2e') else None, v.get('verified'), v.get('single_proof_ms'))
PY
