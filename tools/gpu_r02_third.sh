#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden_sizes.py tests/test_gpu_r1cs.py -m gpu -x -q > gpurun_out/r02_pytest_third.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_third.log
tail -5 gpurun_out/r02_pytest_third.log
./tools/imad_peak > gpurun_out/r02_imad_peak.jsonl 2>&1; cat gpurun_out/r02_imad_peak.jsonl
python tools/gpu_msm_stages.py 12 14 16 18 20 22 > gpurun_out/r02_msm_stages_third.jsonl 2> gpurun_out/r02_msm_stages_third.err
cat gpurun_out/r02_msm_stages_third.jsonl; tail -3 gpurun_out/r02_msm_stages_third.err
python tools/gpu_acc_sweep.py 18 20 > gpurun_out/r02_acc_sweep2.jsonl 2> gpurun_out/r02_acc_sweep2.err
cat gpurun_out/r02_acc_sweep2.jsonl; tail -3 gpurun_out/r02_acc_sweep2.err
