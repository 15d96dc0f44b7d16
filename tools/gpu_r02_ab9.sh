#!/bin/bash
# text leg (front end inside the timer): statements in flight 48 / 64 / 96 with the one-operation range proofs
mkdir -p gpurun_out
: > gpurun_out/r02_ab9.jsonl
for f in 64 96; do
BPG_BENCH_TEXT_INFLIGHT=$f python bench.py --steps 8 --no-cpu-baseline --config4-count 256 --sharded-lg 18 > gpurun_out/r02_bench_text.json 2> gpurun_out/r02_bench_text.err
python - <<PY | tee -a gpurun_out/r02_ab9.jsonl
import json
d=json.load(open('gpurun_out/r02_bench_text.json'))
print(json.dumps({'text_inflight': $f, 'value': round(d['value'],1), 'e2e_statement': round(d['e2e_statement']['value'],1)}))
PY
done
