"""Ad-hoc profiling target: one encode, then N decodes of a 512x1024 latent through the fused codec.
usage: decode_once.py [N decodes] [mode: 0 pipelined | 1 serialized per-kernel timing]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "360-image-compression_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import lic360, lic360_pipeline as pl
from util import synthetic_latent, t, n
dev = "cuda:0"
q, mask, lv = synthetic_latent(2024, H=64, W=128)
params = pl.make_codec_params(dev)
tq, tm, tl = t(q, dev), t(mask, dev), t(lv, dev)
fused = pl.FusedCodec(params)
t0 = time.time(); bi, bc = fused.encode(tq, tm, tl); t1 = time.time()
print("encode %.1f ms" % ((t1 - t0) * 1e3), fused.last_timing(), "bytes", len(bi), len(bc), "launches", lic360.launch_count())
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fused.set_mode(mode)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    t0 = time.time(); code, mup = fused.decode(bi, bc); t1 = time.time()
    ok = bool(np.array_equal(n(code), q * mask) and np.array_equal(n(mup), mask))
    print("decode %.1f ms exact=%s" % ((t1 - t0) * 1e3, ok), fused.last_timing(), "launches", lic360.launch_count())
    if mode == 1:
        print("  code stream kernels:", fused.kernel_times(0))
        print("  imp  stream kernels:", fused.kernel_times(1))
