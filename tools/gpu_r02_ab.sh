#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do python tools/gpu_msm_stages.py 18 22 2>/dev/null | cut -c1-300; done > gpurun_out/r02_ab_stages.log; cat gpurun_out/r02_ab_stages.log
(BPG_F3=1 python tools/gpu_ab_stmt.py; BPG_F3=0 python tools/gpu_ab_stmt.py; BPG_F3=1 python tools/gpu_ab_stmt.py; BPG_F3=0 python tools/gpu_ab_stmt.py) > gpurun_out/r02_ab_f3.log 2>&1; cat gpurun_out/r02_ab_f3.log
