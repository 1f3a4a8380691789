"""Encoder context-conv layer (CconvEcBatch.forward_act_batch through the C-ABI) timed in its two forms of the old-term pass:
SIMT fp32 FMA (default) and tensor cores (LIC360_EC_MMA=1: mma.sync TF32, 3-way split), CUDA events, L2 flushed between iterations.
Also prints the relative difference between the two and against the reference extension when it is present.
usage: bench_ec.py [iters]   -> gpurun_out/ec_mma_bench.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "360-image-compression_b200"), os.path.join(ROOT, "oracle", "_ref")):
    sys.path.insert(0, p)
import numpy as np
import torch
import lic360
from util import conv_weights, rng, rel_err

dev = "cuda:0"
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
try:
    import lic360_ref
except Exception:
    lic360_ref = None
NNZ = {"code_hidden": 470080, "code_first": 112880, "code_last": 352560, "imp_hidden": 311040}  # non-zero MACs per pixel per net
out = []
for name, (N, G, cin, cout, H, W, con, act, nsets) in {
        "code_hidden": (3, 48, 4, 4, 64, 128, 6, True, 3), "code_first": (3, 48, 1, 4, 64, 128, 5, True, 3),
        "code_last": (3, 48, 4, 3, 64, 128, 6, False, 3), "imp_hidden": (1, 1, 144, 144, 32, 64, 6, True, 0)}.items():
    r = rng(5)
    x = torch.from_numpy(np.abs(r.standard_normal((N, G * cin, H, W))).astype(np.float32)).to(dev)
    w, b, a = [torch.from_numpy(v).to(dev) for v in conv_weights(r, nsets, G * cout, G * cin)]

    def call(backend):
        op = backend.CconvEcOp(G * cin, G, G * cout, 5, con, 0, False)
        sfx = "_batch" if nsets else ""
        f = (lambda: getattr(op, "forward_act" + sfx)(x, w, b, a)[0]) if act else (lambda: getattr(op, "forward" + sfx)(x, w, b)[0])
        return op, f

    res = {"layer": name, "shape": [N, G * cin, H, W], "cout": G * cout}
    ys = {}
    for mode in ("simt", "mma"):
        os.environ["LIC360_EC_MMA"] = "1" if mode == "mma" else "0"
        op, f = call(lic360)
        for _ in range(3):
            y = f()
        torch.cuda.synchronize()
        ms = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); y = f(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        ys[mode] = y.clone()
        flops = 2.0 * NNZ[name] * H * W * (nsets or 1)
        res[mode + "_us"] = ms[len(ms) // 2] * 1e3
        res[mode + "_tflops_algorithmic"] = flops / (ms[len(ms) // 2] * 1e-3) / 1e12
    os.environ["LIC360_EC_MMA"] = "0"
    res["rel_err_mma_vs_simt"] = rel_err(ys["mma"].cpu().numpy(), ys["simt"].cpu().numpy())
    if lic360_ref is not None:
        _, f = call(lic360_ref)
        yr = f().clone().cpu().numpy()
        res["rel_err_simt_vs_reference"] = rel_err(ys["simt"].cpu().numpy(), yr)
        res["rel_err_mma_vs_reference"] = rel_err(ys["mma"].cpu().numpy(), yr)
    print(res)
    out.append(res)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ec_mma_bench.json"), "w"), indent=1)
