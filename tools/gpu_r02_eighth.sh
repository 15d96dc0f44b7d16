#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k "variable_base or arbitrary" > gpurun_out/r02_pytest_8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_8.log
tail -4 gpurun_out/r02_pytest_8.log
BPG_ACC_TRACE=1 python tools/gpu_varbase_sweep.py 20 > gpurun_out/r02_varbase20.jsonl 2> gpurun_out/r02_varbase20.err; tail -2 gpurun_out/r02_varbase20.err
python tools/gpu_varbase_sweep.py > gpurun_out/r02_varbase.jsonl 2> gpurun_out/r02_varbase.err; cat gpurun_out/r02_varbase.jsonl; tail -3 gpurun_out/r02_varbase.err
