#!/bin/bash
# hardware queues x statements in flight (config 2, tools/gpu_timeline.py, 256 statements)
mkdir -p gpurun_out
: > gpurun_out/r02_ab5.jsonl
for c in 8 32; do for f in 48 64 96; do
CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 200 python tools/gpu_timeline.py 256 $f 2>> gpurun_out/r02_ab5.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'connections': $c, 'inflight': $f, 'per_statement_ms': round(d['per_statement_ms'], 3), 'acc_frac': round(d['accumulate_running_frac'], 3), 'busy': round(d['union_busy_frac'], 3)}))" | tee -a gpurun_out/r02_ab5.jsonl
done; done
