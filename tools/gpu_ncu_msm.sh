#!/bin/bash
# ncu captures of the MSM kernels at 2^18 points (raw MSM, uniform scalars): full set for accumulate / bucket_reduce / digits
mkdir -p gpurun_out
python tools/prof_msm.py 18 3 > gpurun_out/r02_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_accumulate|k_bucket_reduce|k_digits|k_scan_meta" --launch-skip 10 -c 5 -f -o gpurun_out/r02_msm_kernels python tools/prof_msm.py 18 3 > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_prof_plain.log; tail -5 gpurun_out/r02_ncu_full.log
ncu -i gpurun_out/r02_msm_kernels.ncu-rep --page details --csv > gpurun_out/r02_msm_kernels_details.csv 2>/dev/null
ls -la gpurun_out/r02_msm_kernels*
