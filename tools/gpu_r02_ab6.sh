#!/bin/bash
mkdir -p gpurun_out
for f in 128 160; do
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/gpu_timeline.py 384 $f 2>> gpurun_out/r02_ab5.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'connections': 32, 'inflight': $f, 'per_statement_ms': round(d['per_statement_ms'], 3), 'acc_frac': round(d['accumulate_running_frac'], 3), 'busy': round(d['union_busy_frac'], 3), 'cpu': d['host_cpu_ms_per_statement']['process']}))" | tee -a gpurun_out/r02_ab5.jsonl
done
CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --inflight 96 --steps 16 --no-cpu-baseline --no-extra-legs > gpurun_out/r02_bench_c32_96.json 2> gpurun_out/r02_bench_c32_96.err; python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_c32_96.json'))
print('conn32 inflight96 bench', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'lat', round(d['latency']['ms_per_proof'],1), d['host']['cpu_s_per_proof_rank0'])
PY
