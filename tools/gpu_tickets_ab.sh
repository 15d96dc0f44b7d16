#!/bin/bash
# A/B of the ticket scatter (BPG_TICKETS=1: the scatter pass reuses the ranks returned by the histogram pass's atomics)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_r1cs.py tests/test_gpu_baseline_sizes.py -m gpu -x -q > gpurun_out/tk_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tk_pytest.log
tail -3 gpurun_out/tk_pytest.log
for t in 1 0 1 0; do
  BPG_TICKETS=$t python bench.py --no-cpu-baseline --steps 6 > gpurun_out/tk_$t.json 2> gpurun_out/tk_$t.err
  python - "$t" <<'PY'
import json, sys
t = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/tk_%s.json" % t).read().strip().splitlines()[-1])
    print("tickets", t, "value %.1f e2e %.1f latency %.1f ms msm %.1f us sort %.0f GB/s" % (d["value"], d["e2e"]["value"], d["latency"]["ms_per_proof"], d["msm"]["ms"] * 1e3, d["roofline_sort"]["achieved"]))
except Exception as e:
    print(t, "failed", e)
PY
done
