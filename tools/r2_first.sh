#!/bin/bash
# round-2 first GPU batch: regenerate golden (new full-shape conv cases), GPU suite, smoke, short bench
mkdir -p gpurun_out
timeout 600 python tests/golden/make_golden.py > gpurun_out/r2_golden.log 2>&1; tail -2 gpurun_out/r2_golden.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -15 gpurun_out/r2a_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log)
(timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo bench rc=$?; cut -c1-600 gpurun_out/r2a_bench.json)
