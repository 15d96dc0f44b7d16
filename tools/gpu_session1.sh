#!/bin/bash
# One GPU session: parity tests, phase trace, bench, ncu launch list, ncu full capture of the accumulate kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
BPG_TRACE=1 python tools/gpu_cfg2.py 1024 64 > gpurun_out/s1_cfg2.log 2> gpurun_out/s1_cfg2_trace.log
python bench.py --steps 5 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench_err.log
python tools/prof_step.py 1024 2 > gpurun_out/s1_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s1_launches.csv \
    python tools/prof_step.py 1024 2 > gpurun_out/s1_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate --launch-skip 6 -c 2 -o gpurun_out/s1_accumulate \
    python tools/prof_step.py 1024 2 > gpurun_out/s1_ncu_full.log 2>&1
ls -la gpurun_out
