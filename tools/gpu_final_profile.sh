#!/bin/bash
# Round-end evidence: tests, bench (both arms), launch list, full ncu capture of k_accumulate.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log; tail -2 gpurun_out/f_pytest.log
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"
python tools/prof_step.py 1024 2 > gpurun_out/f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/f_launches.csv python tools/prof_step.py 1024 2 > gpurun_out/f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate --launch-skip 6 -c 1 -f -o gpurun_out/f_accumulate python tools/prof_step.py 1024 2 > gpurun_out/f_ncu_full.log 2>&1
ncu -i gpurun_out/f_accumulate.ncu-rep --page details --csv > gpurun_out/f_accumulate_details.csv 2>/dev/null
ncu -i gpurun_out/f_accumulate.ncu-rep --page raw --csv > gpurun_out/f_accumulate_raw.csv 2>/dev/null
cat gpurun_out/f_bench.json | cut -c1-600; cat gpurun_out/f_bench_ref.json | cut -c1-400
python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f_smoke.log
