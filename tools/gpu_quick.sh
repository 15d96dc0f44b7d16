#!/bin/bash
# quick GPU check: parity tests + config-2 timing/trace
mkdir -p gpurun_out
nproc > gpurun_out/q_host.log; lscpu | grep -E "Model name|Socket|Thread|Core" >> gpurun_out/q_host.log; nvidia-smi -L >> gpurun_out/q_host.log
python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
BPG_TRACE=1 python tools/gpu_cfg2.py 1024 64 > gpurun_out/q_cfg2.log 2> gpurun_out/q_cfg2_trace.log; echo "cfg2 rc=$?" >> gpurun_out/q_cfg2.log
tail -3 gpurun_out/q_pytest.log; tail -12 gpurun_out/q_cfg2.log; cat gpurun_out/q_host.log
