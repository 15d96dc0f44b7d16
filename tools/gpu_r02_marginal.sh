#!/bin/bash
# marginal cost of the stages in the concurrent regime: statements/s with a stage's kernels not launched (BPG_X_SKIP)
mkdir -p gpurun_out
: > gpurun_out/r02_marginal.jsonl
for m in 0 1 2 3 7 8 16 24; do
  BPG_X_SKIP=$m timeout 120 python tools/gpu_timeline.py 128 48 2>> gpurun_out/r02_marginal.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'skip': $m, 'per_statement_ms': round(d['per_statement_ms'], 3), 'acc_frac': round(d['accumulate_running_frac'], 3), 'busy': round(d['union_busy_frac'], 3), 'phases': d['wall_ms_per_statement_by_phase']}))" | tee -a gpurun_out/r02_marginal.jsonl
done
