# 2-GPU run: sharded-proof parity test, then the scaled proof benched at world 1 and 2
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
for k in 16 18; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload board_scaled --k $k --steps 2 --warmup 3 > gpurun_out/scaled${k}_n2.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scaled*_n?.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],2), round(d['ms_per_step'],1), d['n_gpus'], d['scaling'], d['roofline']['kernel_ms'] if d['roofline'] else None, d['verified'])
PY
grep -v "^\[W\|^W1018\|^\*\*\*" gpurun_out/scaled16_n2.log | grep -i "error" | tail -3
