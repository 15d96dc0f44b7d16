#!/bin/bash
# config 4: statements/s against the number of statements in flight; config 2 sanity
mkdir -p gpurun_out
for f in 48 96 144; do
TIMELINE_MODE=c4 timeout 300 python tools/gpu_timeline.py 2048 $f 2>> gpurun_out/r02_c4.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('inflight=$f', json.dumps({'per_statement_ms': round(d['per_statement_ms'], 3), 'busy': round(d['union_busy_frac'], 3), 'conc8': d['concurrency_time_frac'].get('8'), 'phase_sum': d['phase_sum_ms']}))"
done
timeout 300 python tools/gpu_timeline.py 256 48 2>> gpurun_out/r02_c4.err | cut -c1-420
