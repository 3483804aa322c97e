# --set full of the bucket-accumulation kernel of a raw 2^22 Pallas MSM (round-2 counting-sort pipeline)
mkdir -p gpurun_out
export BZ_NO_CPU_BASELINE=1
CMD="python bench.py --workload msm --log2n 22 --steps 2 --warmup 1"
timeout 300 $CMD > gpurun_out/r2msm_plain.log 2>&1 || { tail -5 gpurun_out/r2msm_plain.log; exit 1; }
timeout 280 ncu --set full --clock-control none --import-source on -k regex:"msm_segment_kernel|msm_reduce_kernel|msm_scatter_kernel" -s 3 -c 3 -f -o gpurun_out/r2msm_bucket $CMD > gpurun_out/r2msm_ncu.log 2>&1
ncu -i gpurun_out/r2msm_bucket.ncu-rep --page raw --csv > gpurun_out/r2msm_bucket_raw.csv 2>/dev/null
python profiles/summarize.py gpurun_out/r2msm_bucket_raw.csv | cut -c1-400
