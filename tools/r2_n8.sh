#!/bin/bash
# 8-GPU measurements: configs[4] DDP training step (batch 32 = 4 per GPU at 512x1024) and the codec bench at N = 8
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/ddp_train_step.py --batch 4 --height 512 --width 1024 --steps 5 --out gpurun_out/r2_ddp_n8.json > gpurun_out/r2_ddp_n8.log 2>&1; echo "ddp rc=$?"; tail -3 gpurun_out/r2_ddp_n8.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n8.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "in flight", d.get("images_in_flight"), "config3", d.get("config3_1024x2048_batch16"))
PY
