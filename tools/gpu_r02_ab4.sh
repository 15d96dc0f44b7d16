#!/bin/bash
# generator-fold threshold under the final code: statements/s (tools/gpu_timeline.py, 256 statements, 48 in flight)
mkdir -p gpurun_out
: > gpurun_out/r02_ab4.jsonl
for f in 512 1024 256; do
BPG_IPP_FOLD_N=$f timeout 200 python tools/gpu_timeline.py 256 48 2>> gpurun_out/r02_ab4.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'fold_n': $f, 'per_statement_ms': round(d['per_statement_ms'], 3), 'acc_frac': round(d['accumulate_running_frac'], 3), 'busy': round(d['union_busy_frac'], 3), 'top3': [(t['kernel'], t['per_statement'], t['mean_us']) for t in d['top'][:3]]}))" | tee -a gpurun_out/r02_ab4.jsonl
done
