"""Ad-hoc: time the per-op codec path (this repo vs the reference CUDA extension) on one 512x1024 latent."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "360-image-compression_b200"), os.path.join(ROOT, "oracle", "_ref")):
    sys.path.insert(0, p)
import numpy as np, torch
import lic360, lic360_pipeline as pl
from util import synthetic_latent, t, n
dev = "cuda:0"
q, mask, lv = synthetic_latent(2024, H=64, W=128)
params = pl.make_codec_params(dev)
tq, tm, tl = t(q, dev), t(mask, dev), t(lv, dev)
fused = pl.FusedCodec(params)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    bi, bc = fused.encode(tq, tm, tl); te = fused.last_timing()
    t1 = time.time()
    code, mup = fused.decode(bi, bc); td = fused.last_timing()
    torch.cuda.synchronize(); t2 = time.time()
    ok = bool(torch.equal(code, tq * tm)) and bool(torch.equal(mup, tm))
    print("fused it%d: enc %.1f ms (coder %.1f, gpu wait %.1f) dec %.1f ms (coder %.1f, gpu wait %.1f) bytes imp %d code %d roundtrip %s -> %.2f Mpx/s" % (
        it, (t1 - t0) * 1e3, te['host_coder_ms'], te['gpu_wait_ms'], (t2 - t1) * 1e3, td['host_coder_ms'], td['gpu_wait_ms'], len(bi), len(bc), ok, 0.524288 / (t2 - t0)), flush=True)
backends = [("b200_per_op", lic360)] if "--perop" in sys.argv else []
if "--ref" in sys.argv:
    import lic360_ref
    backends.append(("reference_cuda_ext", lic360_ref))
for name, be in backends:
    codec = pl.PerOpCodec(be, params)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        bi, bc = codec.encode(tq, tm, tl)
        torch.cuda.synchronize(); t1 = time.time()
        code, mup = codec.decode(bi, bc, 32, 64)
        torch.cuda.synchronize(); t2 = time.time()
        ok = bool(torch.equal(code, tq * tm)) and bool(torch.equal(mup, tm))
        print("%s it%d: enc %.1f ms dec %.1f ms  bytes imp %d code %d  bpp %.4f  roundtrip %s  -> %.2f Mpx/s" % (
            name, it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(bi), len(bc), (len(bi) + len(bc)) * 8 / 512 / 1024, ok, 0.524288 / (t2 - t0)), flush=True)
