#!/bin/bash
# one rank of the 8-GPU box emulated on one GPU: 4 host cores (taskset), statements in flight 16 / 32 / 48 / 64
mkdir -p gpurun_out
: > gpurun_out/r02_ab7.jsonl
for f in 16 32 48 64; do
taskset -c 0-3 python bench.py --inflight $f --steps 8 --no-cpu-baseline --no-extra-legs > gpurun_out/r02_bench_4core.json 2> gpurun_out/r02_bench_4core.err
python - <<PY | tee -a gpurun_out/r02_ab7.jsonl
import json
d=json.load(open('gpurun_out/r02_bench_4core.json'))
print(json.dumps({'host_cores': 4, 'inflight': $f, 'value': round(d['value'],1), 'e2e': round(d['e2e']['value'],1), 'cpu_s_per_proof': round(d['host']['cpu_s_per_proof_rank0'],4)}))
PY
done
