#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_r1cs.py tests/test_gpu_batch_api.py -m gpu -x -q > gpurun_out/r02_pytest_7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_7.log
tail -8 gpurun_out/r02_pytest_7.log
python tools/gpu_varbase_sweep.py > gpurun_out/r02_varbase.jsonl 2> gpurun_out/r02_varbase.err; cat gpurun_out/r02_varbase.jsonl; tail -3 gpurun_out/r02_varbase.err
BPG_VARBASE_MIN=999999999 python tools/gpu_varbase_sweep.py 10 12 14 16 18 > gpurun_out/r02_varbase_dyn.jsonl 2>&1; cat gpurun_out/r02_varbase_dyn.jsonl
