"""GPU: BASELINE.json configs[4] at test size -- the reference's train/ scripts (model_zoo.CMP, EntropyNet2, the loss and train() of
trainDDP_IMP_ENT.py) run unchanged on this repo's operator layer under DistributedDataParallel / NCCL, and the DDP gradients equal
the single-process gradients (tools/ddp_train_step.py).  Uses every visible GPU up to 2; the 8-GPU measurement of the same script
is recorded under profiles/."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_ddp_training_step_matches_single_process(tmp_path):
    have = any(os.path.exists(os.path.join(r, "train", "trainDDP_IMP_ENT.py")) for r in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")))
    if not have:
        pytest.skip("the reference's train/ scripts are not staged (make -f oracle/Makefile.ref pyref)")
    n = min(2, torch.cuda.device_count())
    out = str(tmp_path / "ddp.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "ddp_train_step.py"), "--batch", "1", "--height", "512",
           "--width", "512", "--steps", "2", "--viewport", "64", "--out", out]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert p.returncode == 0, "ddp_train_step failed:\n%s\n%s" % (p.stdout[-3000:], p.stderr[-6000:])
    rep = json.loads(open(out).read())
    assert rep["ok"] and rep["world_size"] == n
    assert rep["grad_check"]["tensors"] > 100 and rep["grad_check"]["max_rel_diff_ddp_vs_single_process"] <= 1e-3, rep["grad_check"]
    assert rep["native_launches_fwd_bwd_pair"] > 0 and rep["ms_per_step"] > 0
    dst = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(dst):
        json.dump(rep, open(os.path.join(dst, "r2_ddp_test_n%d.json" % n), "w"), indent=1)
