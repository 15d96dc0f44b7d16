#!/bin/bash
# full ncu capture of the bucket-reduction kernel of an IPP-round MSM + refreshed per-config measurements
mkdir -p gpurun_out
python tools/bench_configs.py --inflight 32 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
python tools/prof_step.py 1024 2 > gpurun_out/r_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_reduce_chunks --launch-skip 6 -c 1 -f -o gpurun_out/r_reduce python tools/prof_step.py 1024 2 > gpurun_out/r_ncu.log 2>&1
ncu -i gpurun_out/r_reduce.ncu-rep --page details --csv > gpurun_out/r_reduce_details.csv 2>/dev/null
ncu -i gpurun_out/r_reduce.ncu-rep --page raw --csv > gpurun_out/r_reduce_raw.csv 2>/dev/null
tail -2 gpurun_out/r_ncu.log
