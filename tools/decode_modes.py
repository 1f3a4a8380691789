"""Ad-hoc: decode time of a latent in the graph-replay mode (0) and the low-latency persistent mode (2); reports where the first
wrong symbol is when a mode does not decode exactly."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "360-image-compression_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import lic360, lic360_pipeline as pl
from util import synthetic_latent, t, n
dev = "cuda:0"
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 128)
q, mask, lv = synthetic_latent(2024, H=H, W=W)
params = pl.make_codec_params(dev)
tq, tm, tl = t(q, dev), t(mask, dev), t(lv, dev)
for mode in (0, 2, 0, 2):
    cd = pl.FusedCodec(params, H=H, W=W, mode=mode)
    bi, bc = cd.encode(tq, tm, tl)
    ts = []
    for it in range(5):
        t0 = time.time(); code, mup = cd.decode(bi, bc); torch.cuda.synchronize(); ts.append((time.time() - t0) * 1e3)
    got, exp = n(code), q * mask
    ok = bool(np.array_equal(got, exp) and np.array_equal(n(mup), mask))
    msg = ""
    if not ok:
        bad = np.argwhere(got != exp)
        steps = bad[:, 1] + bad[:, 2] + bad[:, 3]
        msg = " wrong symbols %d, first wrong step %d (g,h,w)=%s mask_ok=%s" % (len(bad), steps.min(), bad[np.argmin(steps)][1:], np.array_equal(n(mup), mask))
    print("mode %d: decode ms %s exact=%s%s %s" % (mode, ["%.2f" % x for x in ts], ok, msg, {k: round(v, 2) for k, v in cd.last_timing().items()}), flush=True)
    del cd
