#!/bin/bash
# config 4 with 32 hardware queues: statements in flight
mkdir -p gpurun_out
: > gpurun_out/r02_c4c.jsonl
for f in 32 48 64 96 128; do
CUDA_DEVICE_MAX_CONNECTIONS=32 TIMELINE_MODE=c4 timeout 300 python tools/gpu_timeline.py 4096 $f 2>> gpurun_out/r02_c4.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'connections': 32, 'inflight': $f, 'per_statement_ms': round(d['per_statement_ms'], 3), 'per_s': round(1e3 / d['per_statement_ms']), 'busy': round(d['union_busy_frac'], 3), 'conc8': d['concurrency_time_frac'].get('8'), 'cpu_ms': d['host_cpu_ms_per_statement']['process']}))" | tee -a gpurun_out/r02_c4c.jsonl
done
