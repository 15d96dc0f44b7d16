#!/bin/bash
mkdir -p gpurun_out
for d in ${SWEEP:-8 12 16 24}; do
  python bench.py --steps ${STEPS:-8} --warmup 3 --inflight $d --no-cpu-baseline > gpurun_out/sweep_$d.json 2> gpurun_out/sweep_$d.err
  python - <<PY
import json
try:
    l = json.loads(open('gpurun_out/sweep_$d.json').read().strip().splitlines()[-1])
    h = l.get('host', {})
    print('inflight', $d, 'value %.2f' % l['value'], 'e2e %.2f' % l['e2e']['value'], 'lat %.1f ms' % l['latency']['ms_per_proof'], 'roof %.3f' % l['roofline']['frac'],
          'cpu/proof %.1f ms' % (1e3 * h.get('cpu_s_per_proof_rank0', 0)), 'rng', h.get('rng_streams'), h.get('rng_vector_batches'), h.get('rng_streams_alone'))
except Exception as e:
    print('inflight', $d, 'FAILED', e); print(open('gpurun_out/sweep_$d.err').read()[-2000:])
PY
done
