#!/bin/bash
mkdir -p gpurun_out
python tools/prof_step.py 1024 2 > gpurun_out/acc_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate --launch-skip 6 -c 1 -o gpurun_out/acc2 -f python tools/prof_step.py 1024 2 > gpurun_out/acc2_ncu.log 2>&1
ncu -i gpurun_out/acc2.ncu-rep --page raw --csv > gpurun_out/acc2_raw.csv 2>/dev/null
ncu -i gpurun_out/acc2.ncu-rep --page source --csv > gpurun_out/acc2_source.csv 2>/dev/null
ls -la gpurun_out | grep acc2
