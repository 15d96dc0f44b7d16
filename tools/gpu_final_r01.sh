#!/bin/bash
# round-1 closing run: GPU parity suite, smoke(), the default bench line, the other configs with the final code
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python tools/bench_configs.py --only 1,3,4 --inflight 48 > gpurun_out/final_configs.jsonl 2> gpurun_out/final_configs.err; echo "configs rc=$?"
tail -c 300 gpurun_out/final_configs.err
