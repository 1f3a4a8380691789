#!/bin/bash
P=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_ops.py tests/test_gpu_modules.py -m gpu -q -x > gpurun_out/${P}_pytest.log 2>&1; tail -5 gpurun_out/${P}_pytest.log | cut -c1-400
(timeout 300 python tools/bench_ec.py 20 > gpurun_out/${P}_bench_ec.log 2>&1; cat gpurun_out/${P}_bench_ec.log | cut -c1-700)
(timeout 300 python tools/encode_timing.py > gpurun_out/${P}_enc.log 2>&1; tail -12 gpurun_out/${P}_enc.log)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cconv_ec_mma_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/${P}_prof_ec_mma -f python tools/bench_ec.py 2 > gpurun_out/${P}_ncu_ec_mma.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cconv_ec_rq_tile_kernel --launch-skip 8 --launch-count 1 -o gpurun_out/${P}_prof_rq_tile -f python tools/decode_once.py 1 1 > gpurun_out/${P}_ncu_rq_tile.log 2>&1
bash tools/r2_sanitize.sh
ls -la gpurun_out/${P}_prof_*.ncu-rep
