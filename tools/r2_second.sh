#!/bin/bash
# round-2 second GPU batch: golden with the projects cases, new tests first, then the full suite
mkdir -p gpurun_out
timeout 600 python tests/golden/make_golden.py > gpurun_out/r2_golden.log 2>&1; tail -2 gpurun_out/r2_golden.log
cp gpurun_out/golden/ops_golden.npz tests/golden/ops_golden.npz
timeout 1700 python -m pytest tests/test_gpu_reference_scripts.py tests/test_gpu_ddp_train.py tests/test_gpu_parity_baseline.py -q -x > gpurun_out/r2b_new.log 2>&1; tail -30 gpurun_out/r2b_new.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_reference_scripts.py --deselect tests/test_gpu_ddp_train.py --deselect tests/test_gpu_parity_baseline.py > gpurun_out/r2b_pytest.log 2>&1; tail -8 gpurun_out/r2b_pytest.log
