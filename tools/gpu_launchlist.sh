#!/bin/bash
mkdir -p gpurun_out
python tools/prof_step.py 1024 2 > gpurun_out/ll_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ll_launches.csv python tools/prof_step.py 1024 2 > gpurun_out/ll_ncu.log 2>&1
tail -2 gpurun_out/ll_ncu.log
