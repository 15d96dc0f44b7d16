#!/bin/bash
# occupancy cap of k_accumulate (unused dynamic shared memory) so that other statements' sort kernels run beside it
mkdir -p gpurun_out
: > gpurun_out/r02_ab3.jsonl
run() {
  local label=$1; shift
  env "$@" timeout 200 python tools/gpu_timeline.py 192 48 2>> gpurun_out/r02_ab3.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'case': '$label', 'per_statement_ms': round(d['per_statement_ms'], 3), 'acc_frac': round(d['accumulate_running_frac'], 3), 'busy': round(d['union_busy_frac'], 3), 'top3': [(t['kernel'], t['mean_us']) for t in d['top'][:3]]}))" | tee -a gpurun_out/r02_ab3.jsonl
}
run base X=1
run pad57k_3ctas BPG_ACC_SMEM_PAD=58000 BPG_TARGET_CHUNKS=85248
run pad57k_3ctas_2waves BPG_ACC_SMEM_PAD=58000 BPG_TARGET_CHUNKS=113664
run pad57k_conn32 BPG_ACC_SMEM_PAD=58000 BPG_TARGET_CHUNKS=85248 CUDA_DEVICE_MAX_CONNECTIONS=32
run base_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
