"""Host arithmetic coder microbenchmark (no GPU): ns per symbol of the packed-row slab paths the fused codec uses
(lic360_coder_encode_rows / lic360_coder_decode_rows) on peaky 8-bin tables (~1.6 bits/symbol, like the code stream).
usage: python tools/bench_coder.py [symbols]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "360-image-compression_b200")):
    sys.path.insert(0, p)
import numpy as np
from lic360 import _lib
from test_oracle_coder import _pack_rows

L = _lib.LIB
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
r = np.random.default_rng(1)
base = 20000
w = r.random((base, 8)) ** 6 + 1e-3
cdf = np.cumsum(w / w.sum(1, keepdims=True), 1)
tab = np.zeros((base, 9), np.int64)
tab[:, 1:] = np.round(cdf * 65536)
for i in range(8):
    tab[:, i + 1] = np.maximum(tab[:, i + 1], tab[:, i] + 1)
tab[:, -1] = 65536
for i in range(7, 0, -1):
    tab[:, i] = np.minimum(tab[:, i], tab[:, i + 1] - 1)
p = np.diff(tab, axis=1) / 65536.
lab = np.array([r.choice(8, p=pp) for pp in p])
rep = (rows + base - 1) // base
tab = np.concatenate([tab] * rep)[:rows].astype(np.int32)
lab = np.concatenate([lab] * rep)[:rows].astype(np.int32)
mask = np.ones(rows, np.float32)
packed, blank = _pack_rows(tab, lab, mask, 0), _pack_rows(tab, np.zeros_like(lab), mask, 0)
h = ctypes.c_void_p(L.lic360_coder_create(b"x", 3.5))
for it in range(3):
    L.lic360_coder_start_encoder_mem(h)
    t = time.perf_counter()
    assert L.lic360_coder_encode_rows(h, packed.ctypes.data, rows, 0) == 0
    te = time.perf_counter() - t
    n = L.lic360_coder_finish_mem(h)
    buf = np.zeros(n, np.uint8)
    L.lic360_coder_get_bytes(h, buf.ctypes.data, n)
    out = np.zeros(rows, np.float32)
    L.lic360_coder_start_decoder_mem(h, buf.ctypes.data, n)
    t = time.perf_counter()
    assert L.lic360_coder_decode_rows(h, blank.ctypes.data, rows, 0, out.ctypes.data) == 0
    td = time.perf_counter() - t
    print("%.2f bits/symbol  encode %.1f ns/symbol  decode %.1f ns/symbol  exact=%s" % (8 * n / rows, te / rows * 1e9, td / rows * 1e9, np.array_equal(out, lab)))
