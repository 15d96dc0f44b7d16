#!/bin/bash
# evidence for the small-statement path: parity test of the one-kernel MSM, launch list of four config-4 statements,
# ncu --set full of the three new kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k "small_msm or edge or matches_c_oracle" 2>&1 | tail -3
python tools/prof_small.py 2 > gpurun_out/r02_small_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02_launches_cfg4.csv python tools/prof_small.py 2 > gpurun_out/r02_ncu_list_small.log 2>&1
tail -1 gpurun_out/r02_small_plain.log; wc -l gpurun_out/r02_launches_cfg4.csv
ncu --set full --clock-control none --import-source on -k regex:"k_msm_small|k_ipp_round_small|k_csc_small" --launch-skip 6 -c 6 -f -o gpurun_out/r02_small_kernels python tools/prof_small.py 2 > gpurun_out/r02_ncu_small_full.log 2>&1
tail -2 gpurun_out/r02_ncu_small_full.log
ncu -i gpurun_out/r02_small_kernels.ncu-rep --page details --csv > gpurun_out/r02_small_kernels_details.csv 2>/dev/null
ncu -i gpurun_out/r02_small_kernels.ncu-rep --page raw --csv > gpurun_out/r02_small_kernels_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
