"""CPU: the oracle coder and the product coder against the reference's own coder classes and golden bitstreams."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from util import random_tables, rng

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coder_golden.json")


def _case(seed, rows, ncode, masked):
    r = rng(seed)
    tab = random_tables(r, rows, ncode)
    lab = r.integers(0, ncode, rows).astype(np.int32)
    mask = (r.random(rows) > 0.35).astype(np.float32) if masked else None
    return tab, lab, mask


CASES = [(1, 5000, 8, True), (2, 5000, 8, False), (3, 700, 49, False), (4, 1, 8, False), (5, 64, 2, True),
         (6, 0, 8, False), (7, 20000, 8, True)]


def _product_coder():
    from lic360 import _lib
    return _lib.LIB


def _product_encode(tab, lab, mask):
    L = _product_coder()
    h = ctypes.c_void_p(L.lic360_coder_create(b"unused", 3.5))
    assert L.lic360_coder_start_encoder_mem(h) == 0
    if tab.shape[0]:
        rc = L.lic360_coder_encodes(h, tab.ctypes.data, tab.shape[1] - 1, lab.ctypes.data,
                                    None if mask is None else mask.ctypes.data, tab.shape[0])
        assert rc == 0, L.lic360_last_error()
    n = L.lic360_coder_finish_mem(h)
    buf = np.zeros(max(n, 1), np.uint8)
    L.lic360_coder_get_bytes(h, buf.ctypes.data, n)
    L.lic360_coder_destroy(h)
    return buf[:n].tobytes()


def _product_decode(data, tab, mask):
    L = _product_coder()
    h = ctypes.c_void_p(L.lic360_coder_create(b"unused", 3.5))
    arr = np.frombuffer(data, np.uint8).copy()
    assert L.lic360_coder_start_decoder_mem(h, arr.ctypes.data, len(arr)) == 0
    out = np.zeros(tab.shape[0], np.float32)
    if tab.shape[0]:
        rc = L.lic360_coder_decodes(h, tab.ctypes.data, tab.shape[1] - 1, None if mask is None else mask.ctypes.data,
                                    tab.shape[0], out.ctypes.data)
        assert rc == 0, L.lic360_last_error()
    L.lic360_coder_destroy(h)
    return out


def _oracle_encode(tab, lab, mask, cls=O.OracleCoder):
    c = cls()
    c.start_encoder()
    if tab.shape[0]:
        c.encode_rows(tab, lab, mask)
    return c.end_encoder()


@pytest.mark.parametrize("seed,rows,ncode,masked", CASES)
def test_roundtrip_and_cross_equality(seed, rows, ncode, masked, lib_built):
    tab, lab, mask = _case(seed, rows, ncode, masked)
    ob = _oracle_encode(tab, lab, mask)
    pb = _product_encode(tab, lab, mask)
    assert ob == pb, "product coder and oracle coder disagree"
    if O.have_ref_coder():
        rb = _oracle_encode(tab, lab, mask, O.RefCoder)
        assert rb == ob, "oracle coder differs from the reference's own ArithmeticEncoder"
    expect = lab.astype(np.float32) if mask is None else np.where(mask > 0.5, lab, 3.5).astype(np.float32)
    assert np.array_equal(_product_decode(pb, tab, mask), expect)
    d = O.OracleCoder()
    d.start_decoder(ob)
    if rows:
        assert np.array_equal(d.decode_rows(tab, mask), expect)
    if O.have_ref_coder() and rows:
        rd = O.RefCoder()
        rd.start_decoder(ob)
        assert np.array_equal(rd.decode_rows(tab, mask), expect)


def test_golden_bitstreams(lib_built):
    """SHA-256 of streams produced by the reference's own coder classes (tests/golden/make_coder_golden.py)."""
    gold = json.load(open(GOLDEN))
    for key, ref in gold.items():
        seed, rows, ncode, masked = [int(v) for v in key.split(",")]
        tab, lab, mask = _case(seed, rows, ncode, bool(masked))
        for enc in (_oracle_encode, _product_encode):
            data = enc(tab, lab, mask)
            assert len(data) == ref["bytes"] and hashlib.sha256(data).hexdigest() == ref["sha256"], (key, enc.__name__)


def test_step_wise_equals_one_shot(lib_built):
    """the codec calls the coder once per wavefront step; chunking must not change the stream"""
    tab, lab, mask = _case(11, 3000, 8, True)
    one = _product_encode(tab, lab, mask)
    L = _product_coder()
    h = ctypes.c_void_p(L.lic360_coder_create(b"unused", 3.5))
    L.lic360_coder_start_encoder_mem(h)
    for a in range(0, 3000, 250):
        t, l, m = tab[a:a + 250], lab[a:a + 250], mask[a:a + 250]
        assert L.lic360_coder_encodes(h, t.ctypes.data, 8, l.ctypes.data, m.ctypes.data, 250) == 0
    n = L.lic360_coder_finish_mem(h)
    buf = np.zeros(n, np.uint8)
    L.lic360_coder_get_bytes(h, buf.ctypes.data, n)
    L.lic360_coder_destroy(h)
    assert buf.tobytes() == one


def test_error_paths(lib_built, tmp_path):
    L = _product_coder()
    h = ctypes.c_void_p(L.lic360_coder_create(str(tmp_path / "nope" / "x").encode(), 3.5))
    assert L.lic360_coder_start_encoder(h) != 0 and b"cannot open" in L.lic360_last_error()
    assert L.lic360_coder_start_decoder(h) != 0
    tab = np.array([[0, 10, 10, 65536]], np.int32)  # symbol 1 has zero frequency (ArithmeticCoder.cpp:46-47)
    lab = np.array([1], np.int32)
    L.lic360_coder_start_encoder_mem(h)
    assert L.lic360_coder_encodes(h, tab.ctypes.data, 3, lab.ctypes.data, None, 1) != 0
    assert b"zero frequency" in L.lic360_last_error()
    lab[0] = 7
    assert L.lic360_coder_encodes(h, tab.ctypes.data, 3, lab.ctypes.data, None, 1) != 0
    L.lic360_coder_destroy(h)


def test_file_api_matches_reference_layout(lib_built, tmp_path):
    """lic360.Coder (pybind-level mirror): file written at end_encoder, read back by start_decoder."""
    import torch
    import lic360
    tab, lab, mask = _case(21, 1200, 8, True)
    fn = str(tmp_path / "stream")
    c = lic360.Coder("tmp", 3.5)
    c.reset_fname(fn)
    c.start_encoder()
    c.encodes_mask(torch.from_numpy(tab), 8, torch.from_numpy(lab), torch.from_numpy(mask), 1200)
    c.end_encoder()
    assert open(fn, "rb").read() == _oracle_encode(tab, lab, mask)
    c.start_decoder()
    out = c.decodes_mask(torch.from_numpy(tab), 8, torch.from_numpy(mask), 1200)
    assert out.shape == (1200,) and np.array_equal(out.numpy(), np.where(mask > 0.5, lab, 3.5).astype(np.float32))


# ---- packed CDF rows (the format the fused codec's kernels hand to the host coder) -----------------------------------
def _pack_rows(tab, lab, mask, kind):
    """numpy restatement of pack_gmm_row (tables_dev.cuh) / the importance-row packing (codec.cu), include/lic360_b200.h."""
    rows = tab.shape[0]
    if kind == 0:
        out = np.zeros((rows, 8), np.uint16)
        inner = tab[:, 1:8].astype(np.int64)
        out[:, :7] = (inner & 0xFFFF).astype(np.uint16)
        ovf = (((inner >> 16) & 1) << np.arange(7)).sum(1)
        m = np.ones(rows, np.int64) if mask is None else (mask >= 0.5).astype(np.int64)
        out[:, 7] = ((lab.astype(np.int64) & 7) | (m << 8) | (ovf << 9)).astype(np.uint16)
    else:
        out = np.zeros((rows, 64), np.uint16)
        inner = tab[:, 1:49].astype(np.int64)
        out[:, :48] = (inner & 0xFFFF).astype(np.uint16)
        out[:, 48] = lab.astype(np.uint16)
        bits = (inner >> 16) & 1
        for w in range(3):
            out[:, 49 + w] = (bits[:, 16 * w:16 * w + 16] << np.arange(16)).sum(1).astype(np.uint16)
    return np.ascontiguousarray(out)


def _overflow_tables(r, rows, ncode):
    """rows whose last interior bin is 65536 itself (bit 16 set, low word 0: an empty top symbol), the one value of a valid table that
    needs the 17th bit of the packed format"""
    tab = random_tables(r, rows, ncode).astype(np.int64)
    k = rows // 3
    tab[:k, -2] = 65536
    return tab.astype(np.int32)


@pytest.mark.parametrize("kind,ncode,rows,masked", [(0, 8, 6000, True), (0, 8, 1500, False), (1, 49, 900, False), (0, 8, 0, True)])
def test_packed_rows_equal_table_path(kind, ncode, rows, masked, lib_built):
    L = _product_coder()
    r = rng(100 + kind + rows)
    tab = random_tables(r, rows, ncode)
    lab = r.integers(0, ncode, rows).astype(np.int32)
    if rows:  # make sure every symbol value and the extreme bins occur
        lab[:ncode] = np.arange(ncode)
    mask = (r.random(rows) > 0.35).astype(np.float32) if masked else None
    want = _product_encode(tab, lab, mask)   # pinned against the reference coder by test_roundtrip / test_golden
    packed = _pack_rows(tab, lab, mask, kind)
    h = ctypes.c_void_p(L.lic360_coder_create(b"unused", 3.5))
    assert L.lic360_coder_start_encoder_mem(h) == 0
    for a in range(0, rows, 257):  # step-wise, as the codec calls it
        chunk = np.ascontiguousarray(packed[a:a + 257])
        assert L.lic360_coder_encode_rows(h, chunk.ctypes.data, chunk.shape[0], kind) == 0, L.lic360_last_error()
    n = L.lic360_coder_finish_mem(h)
    buf = np.zeros(max(n, 1), np.uint8)
    L.lic360_coder_get_bytes(h, buf.ctypes.data, n)
    assert buf[:n].tobytes() == want
    # decode: the rows of a decoder carry no symbol
    blank = _pack_rows(tab, np.zeros_like(lab), mask, kind)
    arr = np.frombuffer(want, np.uint8).copy()
    assert L.lic360_coder_start_decoder_mem(h, arr.ctypes.data, len(arr)) == 0
    out = np.full(rows, -1, np.float32)
    for a in range(0, rows, 257):
        chunk = np.ascontiguousarray(blank[a:a + 257])
        o = out[a:a + 257]
        assert L.lic360_coder_decode_rows(h, chunk.ctypes.data, chunk.shape[0], kind, o.ctypes.data) == 0, L.lic360_last_error()
    expect = lab.astype(np.float32) if mask is None else np.where(mask > 0.5, lab, 3.5).astype(np.float32)
    assert np.array_equal(out, expect)
    L.lic360_coder_destroy(h)


def test_packed_rows_overflow_bins_and_corrupt_stream(lib_built):
    L = _product_coder()
    r = rng(77)
    rows = 3000
    tab = _overflow_tables(r, rows, 8)
    assert (np.diff(tab.astype(np.int64), axis=1)[:, :-1] > 0).all() and (tab[:rows // 3, 7] == 65536).all()
    lab = r.integers(0, 7, rows).astype(np.int32)   # symbol 7 is empty in the overflow rows: never coded there
    mask = np.ones(rows, np.float32)
    packed, blank = _pack_rows(tab, lab, mask, 0), _pack_rows(tab, np.zeros_like(lab), mask, 0)
    h = ctypes.c_void_p(L.lic360_coder_create(b"unused", 3.5))
    L.lic360_coder_start_encoder_mem(h)
    assert L.lic360_coder_encode_rows(h, packed.ctypes.data, rows, 0) == 0, L.lic360_last_error()
    n = L.lic360_coder_finish_mem(h)
    buf = np.zeros(n, np.uint8)
    L.lic360_coder_get_bytes(h, buf.ctypes.data, n)
    L.lic360_coder_start_decoder_mem(h, buf.ctypes.data, n)
    out = np.zeros(rows, np.float32)
    assert L.lic360_coder_decode_rows(h, blank.ctypes.data, rows, 0, out.ctypes.data) == 0, L.lic360_last_error()
    assert np.array_equal(out, lab.astype(np.float32))
    # calling without a started decoder / with a bad kind is an error, not a crash
    h2 = ctypes.c_void_p(L.lic360_coder_create(b"unused", 3.5))
    assert L.lic360_coder_decode_rows(h2, blank.ctypes.data, rows, 0, out.ctypes.data) != 0
    assert L.lic360_coder_decode_rows(h, blank.ctypes.data, rows, 2, out.ctypes.data) != 0
    L.lic360_coder_destroy(h)
    L.lic360_coder_destroy(h2)
