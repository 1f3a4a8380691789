#!/bin/bash
# compute-sanitizer passes over one small encode + decode of the fused codec (memcheck: out-of-bounds / misaligned accesses of every
# kernel incl. the TMA-fed ones; racecheck: shared-memory hazards inside the chain / old-term / R-Q kernels; synccheck: barrier misuse)
mkdir -p gpurun_out
cat > /tmp/san_target.py <<'PY'
import os, sys
ROOT = os.environ["GRAFT_REPO_ROOT"] if "GRAFT_REPO_ROOT" in os.environ else os.getcwd()
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "360-image-compression_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import lic360_pipeline as pl
from util import synthetic_latent, t, n
H, W = int(sys.argv[1]), int(sys.argv[2])
q, mask, lv = synthetic_latent(11, H=H, W=W)
cd = pl.FusedCodec(pl.make_codec_params("cuda:0", seed=3), H=H, W=W)
bi, bc = cd.encode(t(q), t(mask), t(lv))
code, mup = cd.decode(bi, bc)
print("round trip exact:", bool(np.array_equal(n(code), q * mask) and np.array_equal(n(mup), mask)), len(bi), len(bc))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_target.py 12 20 > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|round trip" gpurun_out/r2_sanitizer_$tool.log | tail -3
done
