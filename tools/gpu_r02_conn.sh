#!/bin/bash
# A/B of the number of hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS) for 48 statements in flight, and the launch
# list of one prove+verify with the generator fold forced on (the schedule the throughput legs run).
mkdir -p gpurun_out
for c in 8 32; do CUDA_DEVICE_MAX_CONNECTIONS=$c python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r02_bench_conn$c.json 2> gpurun_out/r02_bench_conn$c.err; python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_conn$c.json'))
print('connections=$c', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'lat', round(d['latency']['ms_per_proof'],1))
PY
done
BPG_IPP_FOLD_N=512 python tools/prof_step.py 1024 2 > gpurun_out/r02_plain_step_fold.log 2>&1 && \
BPG_IPP_FOLD_N=512 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_cfg2_fold.csv python tools/prof_step.py 1024 2 > gpurun_out/r02_ncu_list_step_fold.log 2>&1
tail -2 gpurun_out/r02_ncu_list_step_fold.log; wc -l gpurun_out/r02_launches_cfg2_fold.csv
