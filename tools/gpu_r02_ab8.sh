#!/bin/bash
# one rank of the 2-GPU box (12 host cores) emulated on one GPU: statements in flight 48 / 64 / 96; and 8 cores (4-GPU box)
mkdir -p gpurun_out
: > gpurun_out/r02_ab8.jsonl
for spec in "0-11 48" "0-11 64" "0-11 96" "0-7 64"; do
set -- $spec
taskset -c $1 python bench.py --inflight $2 --steps 8 --no-cpu-baseline --no-extra-legs > gpurun_out/r02_bench_ncore.json 2> gpurun_out/r02_bench_ncore.err
python - <<PY | tee -a gpurun_out/r02_ab8.jsonl
import json
d=json.load(open('gpurun_out/r02_bench_ncore.json'))
print(json.dumps({'host_cores': '$1', 'inflight': $2, 'value': round(d['value'],1), 'e2e': round(d['e2e']['value'],1), 'cpu_s_per_proof': round(d['host']['cpu_s_per_proof_rank0'],4)}))
PY
done
