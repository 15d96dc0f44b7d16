#!/bin/bash
mkdir -p gpurun_out
nproc
for n in 32 48 64 24; do
  python bench.py --no-cpu-baseline --steps 12 --inflight $n > gpurun_out/if_$n.json 2> gpurun_out/if_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/if_%s.json" % n).read().strip().splitlines()[-1])
    print("inflight", n, "value %.1f e2e %.1f stmt %.1f cpu_s/proof %.3f" % (d["value"], d["e2e"]["value"], d["e2e_statement"]["value"], d["host"]["cpu_s_per_proof_rank0"]))
except Exception as e:
    print(n, "failed", e)
PY
done
