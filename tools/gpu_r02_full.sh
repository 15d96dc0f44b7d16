#!/bin/bash
# full GPU suite + bench (both arms) + smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r02_pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_full.log; tail -16 gpurun_out/r02_pytest_full.log
python tools/gpu_msm_stages.py 12 14 16 18 20 22 > gpurun_out/r02_msm_stages_full.jsonl 2> gpurun_out/r02_msm_stages_full.err; grep '"lg": 18\|"lg": 22' gpurun_out/r02_msm_stages_full.jsonl | cut -c1-330
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_full.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e', d['e2e']['value'], 'stmt', d.get('e2e_statement',{}).get('value'), 'lat', d['latency']['ms_per_proof'])
print('msm', d['msm']['ms'], d['msm']['frac_of_imad_peak_whole_msm'], d['msm']['stage_us'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','peak','peak_sustained','frac','share_of_step')})
print('config4', d.get('config4',{}).get('value')); print('sharded', d.get('msm_sharded',{}).get('mpoints_per_s'))
print('cpu', d.get('cpu_baseline')); print(d['clocks'])
r=json.load(open('gpurun_out/r02_bench_ref.json')); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
