#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden_sizes.py tests/test_gpu_batch_api.py tests/test_gpu_r1cs.py -m gpu -x -q --durations=8 > gpurun_out/r02_pytest_second.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_second.log
tail -15 gpurun_out/r02_pytest_second.log
python tools/gpu_msm_stages.py 12 14 16 18 20 22 > gpurun_out/r02_msm_stages_second.jsonl 2> gpurun_out/r02_msm_stages_second.err
cat gpurun_out/r02_msm_stages_second.jsonl; tail -3 gpurun_out/r02_msm_stages_second.err
python tools/gpu_acc_sweep.py 18 20 > gpurun_out/r02_acc_sweep.jsonl 2> gpurun_out/r02_acc_sweep.err
cat gpurun_out/r02_acc_sweep.jsonl; tail -3 gpurun_out/r02_acc_sweep.err
