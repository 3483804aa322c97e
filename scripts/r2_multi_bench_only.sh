# N-GPU run (N = $1): the scaled proof at world N (sizes in $KS)
N=${1:-4}
mkdir -p gpurun_out
export BZ_NO_CPU_BASELINE=1
for k in ${KS:-18 20}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload board_scaled --k $k --steps 2 --warmup 3 > gpurun_out/scaled${k}_n$N.log 2>&1
done
python - $N <<'PY'
import json,glob,sys
N=sys.argv[1]
for f in sorted(glob.glob(f'gpurun_out/scaled*_n{N}.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],3), round(d['ms_per_step'],1), d['n_gpus'], d['roofline']['kernel_ms'] if d['roofline'] else None, d['verified'], d['single_proof_ms'])
PY
