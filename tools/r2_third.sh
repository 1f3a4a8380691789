#!/bin/bash
# round-2 third GPU batch: full suite with the tiled R/Q kernel, bench, streaming-op roofline, ncu captures of the streaming kernels
P=${1:-r2d}
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/${P}_pytest.log 2>&1; tail -6 gpurun_out/${P}_pytest.log | cut -c1-400
(timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo bench rc=$?; cut -c1-300 gpurun_out/${P}_bench.json)
(timeout 600 python tools/bench_ops.py 32 > gpurun_out/${P}_bench_ops.log 2>&1; cp gpurun_out/ops_roofline.json gpurun_out/${P}_ops_roofline.json; cat gpurun_out/${P}_bench_ops.log)
for spec in "gmm_table_kernel gmm_table" "entropy_table_kernel entropy_table" "quant_fwd8_kernel quant_fwd8" "entropy_gmm_fwd_kernel entropy_gmm_fwd"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip 3 --launch-count 1 -o gpurun_out/${P}_prof_$2 -f python tools/bench_ops.py 32 > gpurun_out/${P}_ncu_$2.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cconv_ec_rq_tile_kernel --launch-skip 8 --launch-count 1 -o gpurun_out/${P}_prof_rq_tile -f python tools/decode_once.py 1 1 > gpurun_out/${P}_ncu_rq_tile.log 2>&1
ls -la gpurun_out/${P}_prof_*.ncu-rep
