"""Generates tests/golden/coder_golden.json from the REFERENCE's own coder classes
(oracle/_ref/libref_coder.so = ArithmeticCoder.cpp + BitIoStream.cpp compiled in place). Run in the build container:
    python tests/golden/make_coder_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402
from test_oracle_coder import CASES, _case  # noqa: E402

assert O.have_ref_coder(), "build oracle/_ref first: make -f oracle/Makefile.ref coder"
gold = {}
for seed, rows, ncode, masked in CASES:
    tab, lab, mask = _case(seed, rows, ncode, masked)
    c = O.RefCoder()
    c.start_encoder()
    if rows:
        c.encode_rows(tab, lab, mask)
    data = c.end_encoder()
    gold["%d,%d,%d,%d" % (seed, rows, ncode, int(masked))] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest()}
json.dump(gold, open(os.path.join(HERE, "coder_golden.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(gold, indent=1))
