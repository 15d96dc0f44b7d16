#!/bin/bash
# first GPU call of round 2: MSM parity (incl. golden full sizes), per-stage MSM timing, a short bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden_sizes.py tests/test_gpu_r1cs.py -m gpu -x -q --durations=8 > gpurun_out/r02_pytest_first.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_first.log
tail -15 gpurun_out/r02_pytest_first.log
python tools/gpu_msm_stages.py 12 14 16 18 20 22 > gpurun_out/r02_msm_stages_first.jsonl 2> gpurun_out/r02_msm_stages_first.err
cat gpurun_out/r02_msm_stages_first.jsonl; tail -3 gpurun_out/r02_msm_stages_first.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_first.json 2> gpurun_out/r02_bench_first.err
tail -c 3000 gpurun_out/r02_bench_first.json; tail -5 gpurun_out/r02_bench_first.err
