#!/bin/bash
# Round-2 evidence: ncu full captures of the MSM kernels (raw MSM 2^18), launch list of one config-2 step, other configs.
mkdir -p gpurun_out
python tools/prof_msm.py 18 3 > gpurun_out/r02_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_accumulate|k_bucket_reduce|k_sort_|k_digits|k_scan_meta" --launch-skip 12 -c 6 -f -o gpurun_out/r02_msm_kernels python tools/prof_msm.py 18 3 > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log
ncu -i gpurun_out/r02_msm_kernels.ncu-rep --page details --csv > gpurun_out/r02_msm_kernels_details.csv 2>/dev/null
python tools/prof_step.py 1024 2 > gpurun_out/r02_plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_cfg2.csv python tools/prof_step.py 1024 2 > gpurun_out/r02_ncu_list_step.log 2>&1
tail -2 gpurun_out/r02_ncu_list_step.log; wc -l gpurun_out/r02_launches_cfg2.csv
python tools/bench_configs.py --inflight 48 --only 1,3 > gpurun_out/r02_configs_1gpu.jsonl 2> gpurun_out/r02_configs_1gpu.err; cat gpurun_out/r02_configs_1gpu.jsonl | cut -c1-400; tail -2 gpurun_out/r02_configs_1gpu.err
