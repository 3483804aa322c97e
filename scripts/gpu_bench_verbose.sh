BZ_BENCH_VERBOSE=1 python bench.py > gpurun_out/v_bench.log 2> gpurun_out/v_bench.err; grep -a "free / total" gpurun_out/v_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/v_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['int_pipe']['mixed_adds_per_step'], {k:round(v/d['steps'],2) for k,v in d['roofline']['kernel_ms'].items()}, d.get('verified'), d.get('single_proof_ms'))
        print({k:(v.get('value'),v.get('unit'),v.get('error'), (v.get('int_pipe') or {}).get('mixed_adds_per_step')) for k,v in d.get('extras',{}).items()})
PY
