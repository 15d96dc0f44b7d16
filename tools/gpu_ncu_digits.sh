#!/bin/bash
# full ncu capture of the sort-stage kernel (digit decomposition + scatter) of an IPP-round MSM
mkdir -p gpurun_out
python tools/prof_step.py 1024 2 > gpurun_out/d_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_digits --launch-skip 13 -c 1 -f -o gpurun_out/d_digits python tools/prof_step.py 1024 2 > gpurun_out/d_ncu.log 2>&1
ncu -i gpurun_out/d_digits.ncu-rep --page details --csv > gpurun_out/d_digits_details.csv 2>/dev/null
ncu -i gpurun_out/d_digits.ncu-rep --page raw --csv > gpurun_out/d_digits_raw.csv 2>/dev/null
tail -3 gpurun_out/d_ncu.log
