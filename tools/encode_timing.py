"""Ad-hoc: N encodes of a 512x1024 latent through the fused codec with the host-side timing split."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "360-image-compression_b200")):
    sys.path.insert(0, p)
import torch
import lic360, lic360_pipeline as pl
from util import synthetic_latent, t
dev = "cuda:0"
q, mask, lv = synthetic_latent(2024, H=64, W=128)
params = pl.make_codec_params(dev)
tq, tm, tl = t(q, dev), t(mask, dev), t(lv, dev)
fused = pl.FusedCodec(params)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t0 = time.time(); bi, bc = fused.encode(tq, tm, tl); t1 = time.time()
    print("encode %.1f ms" % ((t1 - t0) * 1e3), fused.last_timing())
