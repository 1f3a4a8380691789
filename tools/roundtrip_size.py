"""Ad-hoc: fused codec round trip at another latent size. usage: roundtrip_size.py H W [n decodes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "360-image-compression_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import lic360, lic360_pipeline as pl
from util import synthetic_latent, t, n
dev = "cuda:0"
H, W = int(sys.argv[1]), int(sys.argv[2])
q, mask, lv = synthetic_latent(77, H=H, W=W)
params = pl.make_codec_params(dev)
fused = pl.FusedCodec(params, H=H, W=W, mode=int(os.environ.get("LIC360_MODE", "0")))
tq, tm, tl = t(q, dev), t(mask, dev), t(lv, dev)
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 2):
    t0 = time.time(); bi, bc = fused.encode(tq, tm, tl); t1 = time.time()
    code, mup = fused.decode(bi, bc); t2 = time.time()
    ok = bool(np.array_equal(n(code), q * mask) and np.array_equal(n(mup), mask))
    mpx = (8 * H) * (8 * W) / 1e6
    print("latent %dx%d (%dx%d px): encode %.1f ms decode %.1f ms exact=%s bytes %d+%d -> %.2f Mpx/s" % (H, W, 8 * H, 8 * W, (t1 - t0) * 1e3, (t2 - t1) * 1e3, ok, len(bi), len(bc), mpx / (t2 - t0)))
