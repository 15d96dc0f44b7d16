#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden_sizes.py tests/test_gpu_r1cs.py tests/test_gpu_random_circuits.py -m gpu -x -q > gpurun_out/r02_pytest_9.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_9.log
tail -8 gpurun_out/r02_pytest_9.log
python tools/gpu_msm_stages.py 14 16 18 20 22 > gpurun_out/r02_msm_stages_9.jsonl 2> gpurun_out/r02_msm_stages_9.err
cat gpurun_out/r02_msm_stages_9.jsonl; tail -3 gpurun_out/r02_msm_stages_9.err
BPG_SMEM_SORT=0 python tools/gpu_msm_stages.py 18 > gpurun_out/r02_msm_stages_9_old.jsonl 2>&1; cat gpurun_out/r02_msm_stages_9_old.jsonl
