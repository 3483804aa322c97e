mkdir -p gpurun_out
export BZ_NO_CPU_BASELINE=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-extras --steps 3 --warmup 3 > gpurun_out/last_shot_n2.log 2>&1
python - <<'PY'
import json
ok=False
for l in open('gpurun_out/last_shot_n2.log'):
    if l.startswith('{'):
        ok=True
        d=json.loads(l); print('n2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['n_gpus'], d['verified'], d['verify'] and d['verify']['accepted'])
if not ok: print(open('gpurun_out/last_shot_n2.log').read()[-1500:])
PY
