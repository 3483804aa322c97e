export BZ_FB_PAIRS=1
CMD="python bench.py --no-extras --batch 64 --inflight 1 --steps 1 --warmup 3"
$CMD > gpurun_out/pp_plain.log 2>&1 || { tail -5 gpurun_out/pp_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fb_|batch_invert" -c 800 --csv --log-file gpurun_out/pp_launches.csv $CMD > gpurun_out/pp_ncu.log 2>&1
python profiles/summarize.py --launches gpurun_out/pp_launches.csv | head -12
ncu --set full --clock-control none --import-source on -k regex:fb_accumulate_pairs_kernel -s 30 -c 2 -f -o gpurun_out/pp_pairs $CMD > gpurun_out/pp_ncu2.log 2>&1
ncu -i gpurun_out/pp_pairs.ncu-rep --page raw --csv > gpurun_out/pp_pairs_raw.csv 2>/dev/null
python profiles/summarize.py gpurun_out/pp_pairs_raw.csv
