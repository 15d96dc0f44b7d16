#!/bin/bash
# throughput-regime A/B of scheduling knobs (statements/s through tools/gpu_timeline.py, 128 statements, 48 in flight)
mkdir -p gpurun_out
: > gpurun_out/r02_ab2.jsonl
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 150 python tools/gpu_timeline.py 128 48 2>> gpurun_out/r02_ab2.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'case': '$label', 'per_statement_ms': round(d['per_statement_ms'], 3), 'acc_frac': round(d['accumulate_running_frac'], 3), 'busy': round(d['union_busy_frac'], 3), 'top3': [(t['kernel'], t['mean_us']) for t in d['top'][:3]]}))" | tee -a gpurun_out/r02_ab2.jsonl
}
run base TIMELINE_DUMP=gpurun_out/r02_timeline_events.txt.gz
run smem_sort0 BPG_SMEM_SORT=0
run acc_variant1 BPG_ACC_VARIANT=1
run chunks_1wave BPG_TARGET_CHUNKS=75776
run chunks_2wave BPG_TARGET_CHUNKS=151552
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
