#!/bin/bash
# Round-2 measurement batch on the GPU box (run through gpurun from the repo root): GPU tests, smoke, both bench arms, the ncu launch list
# of the bench command and one `--set full` capture of each dominant kernel.  Everything lands in gpurun_out/.
P=${1:-r2z}
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest.log 2>&1; tail -2 gpurun_out/${P}_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${P}_smoke.log 2>&1; tail -1 gpurun_out/${P}_smoke.log | cut -c1-400)
(timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo bench rc=$?)
(LIC360_BENCH_DECODE_MODE=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${P}_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-ext > gpurun_out/${P}_ncu_bench.log 2>&1; echo ncu rc=$?)
for spec in "wf_old4_kernel 150 old4" "wf_chain4_kernel 150 chain4" "wf_chain1_kernel 50 chain1" "cconv_ec_kernel 7 ec"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 --launch-count 1 -o gpurun_out/${P}_prof_$3 -f python tools/decode_once.py 1 1 > gpurun_out/${P}_ncu_$3.log 2>&1
done
ls -la gpurun_out/${P}_prof_*.ncu-rep
cut -c1-300 gpurun_out/${P}_bench.json
(timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err; echo benchref rc=$?; cut -c1-500 gpurun_out/${P}_bench_ref.json)
