#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden_sizes.py tests/test_gpu_r1cs.py -m gpu -x -q > gpurun_out/r02_pytest_4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_4.log
tail -5 gpurun_out/r02_pytest_4.log
BPG_REDUCE_TRACE=1 python tools/gpu_msm_stages.py 12 14 16 18 20 22 > gpurun_out/r02_msm_stages_4.jsonl 2> gpurun_out/r02_msm_stages_4.err
cat gpurun_out/r02_msm_stages_4.jsonl; grep "bpg reduce" gpurun_out/r02_msm_stages_4.err | awk 'NR%5==0' | head -14
python tools/gpu_acc_sweep.py 18 > gpurun_out/r02_acc_sweep4.jsonl 2> gpurun_out/r02_acc_sweep4.err
cat gpurun_out/r02_acc_sweep4.jsonl; tail -3 gpurun_out/r02_acc_sweep4.err
