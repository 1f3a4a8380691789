#!/usr/bin/env python
"""BASELINE.json configs[4]: the DDP training step of the reference -- train/model_zoo.py::CMP (analysis / synthesis transforms,
ImpMap, QUANT, Dtow, EntropyNet2 = 3 x [MaskConv2 x 12 + PReLU] -> ContextReshape -> EntropyGmm) with the viewport loss of
train/trainDDP_IMP_ENT.py:20-48 (MultiProject x 2, SSIM, MSE, rate) -- run UNCHANGED from the reference's train/ directory on top of
this repo's operator layer (`sphere_operator` alias -> lic360_operator -> lic360 mirror -> C-ABI kernels), under
torch.nn.parallel.DistributedDataParallel over NCCL exactly as Job() does (:105-108,142).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/ddp_train_step.py \
         [--batch 4] [--height 512] [--width 1024] [--steps 5] [--out gpurun_out/r2_ddp_nN.json]

What it does on every rank (same seeded data on all ranks, so that the DDP-averaged gradients must equal a single-process run):
  1. builds CMP twice from the same seed: one wrapped in DDP, one plain replica;
  2. one forward/backward of the reference loss on both -> max relative gradient difference (the "equal to the single-process
     run" check; fp32 atomics in the QUANT / MultiProject backward make the last bits order-dependent);
  3. calls the reference's own train() (:20-48) on a 2-batch synthetic loader (proves the trainer code runs unchanged);
  4. times K steps with the gradient all-reduce and K steps under no_sync() -> step time and all-reduce share.
Rank 0 prints one JSON line.  Test / measurement infrastructure; nothing in the product imports it.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "360-image-compression_b200")


def reference_root():
    for cand in (os.environ.get("LIC360_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.exists(os.path.join(cand, "train", "trainDDP_IMP_ENT.py")):
            return cand
    raise SystemExit("reference train/ scripts not found (run `make -f oracle/Makefile.ref pyref` in the build container)")


class _Sampler(object):
    def set_epoch(self, e):
        self.epoch = e


class _Loader(object):
    """the three things train() touches: .sampler.set_epoch, iteration, len(.dataset) / len()"""

    def __init__(self, batches):
        self.batches, self.sampler, self.dataset = batches, _Sampler(), [0] * (len(batches) * batches[0].shape[0])

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


class _Log(object):
    def __init__(self):
        self.lines = []

    def log(self, s):
        self.lines.append(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--viewport", type=int, default=171)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")

    import torch
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    tk, mb = types.ModuleType("tkinter"), types.ModuleType("tkinter.messagebox")
    mb.NO, tk.messagebox = "no", mb
    sys.modules.setdefault("tkinter", tk)
    sys.modules.setdefault("tkinter.messagebox", mb)
    ref = reference_root()
    os.environ.setdefault("LIC360_REFERENCE_ROOT", ref)  # GDN / SSIM / Logger / ModuleSaver pass-through of lic360_operator
    sys.path.insert(0, os.path.join(ref, "train"))
    sys.path.insert(0, PKG)
    import lic360
    import lic360_operator  # this repo's module layer (the train/ scripts reach it through the `sphere_operator` alias)
    assert "360-image-compression_b200" in lic360_operator.__file__
    with contextlib.redirect_stdout(io.StringIO()):
        import model_zoo                      # reference train/model_zoo.py, unchanged
        import trainDDP_IMP_ENT as T          # reference trainer, unchanged (imported, not spawned)
    assert os.path.abspath(model_zoo.__file__).startswith(os.path.abspath(ref)) and os.path.abspath(T.__file__).startswith(os.path.abspath(ref))
    from sphere_operator import MultiProject, SSIM

    args = argparse.Namespace(batch_size=a.batch, alpha=1.0, beta=3000.0, gamma=30.0, clip=0.006, log_interval=1, gpu_id=local_rank,
                              channels=192, code_channels=192, quant_levels=8, rt=0.15, scale_const=0.7, scale_weight=0.7, la=0.0018,
                              lb=0.0001, init=False)

    def build():
        torch.manual_seed(1234)
        with contextlib.redirect_stdout(io.StringIO()):
            return model_zoo.CMP(args).to(dev)

    model, replica = build(), build()
    ddp = DDP(model, [local_rank])
    pr1 = MultiProject(a.viewport, int(a.viewport * 1.5), 0.5, False, local_rank).to(dev)
    pr2 = MultiProject(a.viewport, int(a.viewport * 1.5), 0.5, False, local_rank).to(dev)
    sloss = SSIM(11, 3).to(dev)
    g = torch.Generator().manual_seed(77)  # the SAME data on every rank
    low = torch.rand((a.batch, 3, a.height // 16, a.width // 16), generator=g)
    data = torch.nn.functional.interpolate(low, size=(a.height, a.width), mode="bilinear", align_corners=False).clamp(0, 1).to(dev).contiguous()

    def loss_of(m, x):  # trainDDP_IMP_ENT.py:32-39
        y, ent_vec, rt, _, mask = m(x)
        py, px = pr1(y), pr2(x)
        ssim_loss = 1 - sloss(px, py)
        mse_loss = torch.mean((px - py) * (px - py))
        ent_loss = torch.sum(ent_vec) / torch.sum(mask).item()
        return args.beta * mse_loss + args.alpha * ssim_loss + args.gamma * ent_loss, (mse_loss.item(), ssim_loss.item(), ent_loss.item(), rt.item())

    # ---- 2. gradients: DDP (averaged over ranks that all hold the same batch) vs the plain replica
    l0 = lic360.launch_count()
    ddp.train(); replica.train()
    ddp.zero_grad(); replica.zero_grad()
    la, parts = loss_of(ddp, data)
    la.backward()
    lb, _ = loss_of(replica, data)
    lb.backward()
    torch.cuda.synchronize()
    worst, worst_name, checked, gnorm = 0.0, "", 0, 0.0
    for (n1, p1), (_, p2) in zip(model.named_parameters(), replica.named_parameters()):
        if p1.grad is None or p2.grad is None:
            assert p1.grad is None and p2.grad is None, n1
            continue
        scale = max(p2.grad.abs().max().item(), 1e-12)
        d = (p1.grad - p2.grad).abs().max().item() / scale
        gnorm += float(p2.grad.double().pow(2).sum())
        checked += 1
        if d > worst:
            worst, worst_name = d, n1
    launches_fb = lic360.launch_count() - l0

    # ---- 3. the reference's own train() on a 2-batch loader
    opt_ent = torch.optim.Adam(ddp.module.ent.parameters(), lr=1e-4)
    opt_quant = torch.optim.SGD([ddp.module.quant.count], lr=0.001)
    log = _Log()
    T.train(args, ddp, dev, _Loader([data, data]), opt_ent, opt_quant, 1, log, pr1, pr2, True)
    torch.cuda.synchronize()

    # ---- 4. step time with and without the gradient all-reduce
    opt_all = torch.optim.Adam([{"params": ddp.module.encoder.parameters()}, {"params": ddp.module.decoder.parameters()},
                                {"params": [ddp.module.quant.weight]}, {"params": ddp.module.ent.parameters()}], lr=1e-5)

    def step(sync):
        opt_all.zero_grad(); opt_quant.zero_grad()
        ctx = contextlib.nullcontext() if sync else ddp.no_sync()
        with ctx:
            l, _ = loss_of(ddp, data)
            l.backward()
        torch.nn.utils.clip_grad_norm_(ddp.module.ent.parameters(), args.clip)
        opt_all.step(); opt_quant.step()

    def timed(sync):
        for _ in range(2):
            step(sync)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step(sync)
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_sync, ms_nosync = timed(True), timed(False)
    nparam = sum(p.numel() for p in model.parameters())
    rep = {"config": "configs[4]: DDP training step, train/model_zoo.py::CMP + trainDDP_IMP_ENT.py loss, reference scripts unchanged",
           "world_size": world, "batch_per_gpu": a.batch, "global_batch": a.batch * world, "image": [a.height, a.width], "viewport": a.viewport,
           "parameters": nparam, "gradient_bytes_fp32": 4 * nparam,
           "loss": float(la.item()), "loss_parts_mse_ssim_ent_rt": parts, "loss_replica": float(lb.item()),
           "grad_check": {"tensors": checked, "max_rel_diff_ddp_vs_single_process": worst, "worst_tensor": worst_name, "grad_l2": gnorm ** 0.5},
           "native_launches_fwd_bwd_pair": int(launches_fb),
           "reference_train_fn_log": log.lines[-2:],
           "ms_per_step": ms_sync, "ms_per_step_no_allreduce": ms_nosync, "allreduce_share": max(0.0, 1 - ms_nosync / ms_sync),
           "images_per_s": a.batch * world / (ms_sync / 1e3), "timing": "CUDA events, max over ranks"}
    ok = worst <= 1e-3 and abs(la.item() - lb.item()) <= 1e-4 * abs(lb.item())
    rep["ok"] = bool(ok)
    if rank == 0:
        line = json.dumps(rep)
        print(line)
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            open(a.out, "w").write(line + "\n")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
