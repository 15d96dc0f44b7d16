#!/bin/bash
# config 4 (small statements): full GPU tests, then statements/s with and without the one-kernel small MSM
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_c4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_c4.log
for k in 1 0; do
BPG_SMALL_KERNEL=$k TIMELINE_MODE=c4 timeout 300 python tools/gpu_timeline.py 1024 48 2>> gpurun_out/r02_c4.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('small_kernel=$k', json.dumps({'per_statement_ms': round(d['per_statement_ms'], 3), 'kernels_per_statement': d['kernels'] / d['statements'], 'busy': round(d['union_busy_frac'], 3), 'phases': d['wall_ms_per_statement_by_phase'], 'top': [(t['kernel'], t['per_statement'], t['mean_us']) for t in d['top'][:6]]}))"
done
