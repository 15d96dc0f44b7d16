#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden_sizes.py tests/test_gpu_r1cs.py tests/test_gpu_batch_api.py -m gpu -x -q > gpurun_out/r02_pytest_5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_5.log
tail -5 gpurun_out/r02_pytest_5.log
BPG_REDUCE_TRACE=1 python tools/gpu_msm_stages.py 12 14 16 18 20 22 > gpurun_out/r02_msm_stages_5.jsonl 2> gpurun_out/r02_msm_stages_5.err
cat gpurun_out/r02_msm_stages_5.jsonl; grep "bpg reduce" gpurun_out/r02_msm_stages_5.err | awk 'NR%5==0' | head -14
python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench_5.json 2> gpurun_out/r02_bench_5.err
tail -c 5000 gpurun_out/r02_bench_5.json; tail -5 gpurun_out/r02_bench_5.err
