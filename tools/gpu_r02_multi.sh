#!/bin/bash
# multi-GPU bench (weak scaling of config 2 + config4 + sharded MSM legs); N from the first argument
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "rc=$?"; tail -3 gpurun_out/r02_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_${N}gpu.json'))
print({k:d[k] for k in ('n_gpus','value','ms_per_step')}, 'e2e', d['e2e']['value'], 'stmt', d.get('e2e_statement',{}).get('value'))
print('config4', d.get('config4',{}).get('value'), 'sharded', d.get('msm_sharded',{}).get('mpoints_per_s'), d.get('msm_sharded',{}).get('result'))
print('roofline', d['roofline']['frac'], d['roofline']['peak'], d['roofline']['peak_sustained'], 'msm', d['msm']['frac_of_imad_peak_whole_msm'], d['host']['cores'])
PY
