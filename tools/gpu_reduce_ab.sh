#!/bin/bash
# A/B of the k_reduce_chunks geometry (BPG_REDUCE=threads,blocks): MSM parity tests under the alternative geometry, then
# the headline bench for each setting.
mkdir -p gpurun_out
BPG_REDUCE=32,256 python -m pytest tests/test_gpu_msm.py tests/test_gpu_r1cs.py -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_pytest.log
tail -3 gpurun_out/ab_pytest.log
for cfg in 64,128 32,256 32,512 64,128 32,256; do
  BPG_REDUCE=$cfg python bench.py --no-cpu-baseline --steps 6 > gpurun_out/ab_$cfg.json 2> gpurun_out/ab_$cfg.err
  python - "$cfg" <<'PY'
import json, sys
cfg = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % cfg).read().strip().splitlines()[-1])
    print(cfg, "value %.1f e2e %.1f stmt %.1f latency %.1f ms msm %.1f us" % (d["value"], d["e2e"]["value"], d["e2e_statement"]["value"], d["latency"]["ms_per_proof"], d["msm"]["ms"] * 1e3))
except Exception as e:
    print(cfg, "failed", e)
PY
done
