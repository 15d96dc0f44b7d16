#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_golden_sizes.py -m gpu -x -q -k "fold or config2" > gpurun_out/r02_pytest_fold.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_fold.log; tail -12 gpurun_out/r02_pytest_fold.log
for f in 0 512 2048; do BPG_IPP_FOLD_N=$f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r02_bench_fold$f.json 2> gpurun_out/r02_bench_fold$f.err; python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_fold$f.json'))
print('fold_n=$f', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'lat', round(d['latency']['ms_per_proof'],1), 'peak', d['roofline']['peak'], d['roofline']['frac'])
PY
done
