# launch list of a k = 20 proof (bucket MSM on every commitment, IPA on materialised generators) -> profiles/r2_k20_launches_by_kernel.csv
mkdir -p gpurun_out
export BZ_NO_CPU_BASELINE=1
CMD="python bench.py --workload board_scaled --k 20 --steps 1 --warmup 1"
timeout 600 $CMD > gpurun_out/k20_plain.log 2>&1 || { tail -5 gpurun_out/k20_plain.log; exit 1; }
# skip the set-up launches (Params::new, keygen: ~600), then two proofs' worth
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 2200 --csv --log-file gpurun_out/k20_launches.csv $CMD > gpurun_out/k20_ncu.log 2>&1
python profiles/summarize.py --launches gpurun_out/k20_launches.csv > gpurun_out/k20_launches_by_kernel.csv
head -32 gpurun_out/k20_launches_by_kernel.csv
