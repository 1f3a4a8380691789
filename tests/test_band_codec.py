"""Latitude-band partition of one image (lic360_shard.BandCodec, BASELINE.json configs[3] / SURVEY.md s8e option 3) -- host logic on CPU.
The per-band codec here is the CPU rendition (oracle/cpu_codec.py: test infrastructure standing in for the GPU codec); the GPU
version of these checks is tests/test_gpu_codec.py::test_band_codec_round_trip."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import synthetic_latent

H, W, NB = 8, 8, 2


class _CpuBand(object):
    """CpuCodec behind the (imp_bytes, code_bytes) -> (code, mask) decode signature of FusedCodec"""

    def __init__(self, h, w):
        import lic360_codec_ops as ops
        from oracle import cpu_codec
        params = {'code': ops.make_entropy_params(48, 4, 3, 3, 7, 'cpu'), 'imp': ops.make_entropy_params(1, 144, 49, None, 8, 'cpu')}
        self.h, self.w = h, w
        self.cd = cpu_codec.CpuCodec(cpu_codec.params_to_numpy(params))

    def encode(self, code, mask, imap):
        return self.cd.encode(np.ascontiguousarray(code), np.ascontiguousarray(mask), np.ascontiguousarray(imap))

    def decode(self, bi, bc):
        return self.cd.decode(bi, bc, self.h // 2, self.w // 2)


class _Np(object):
    """numpy arrays with the two tensor methods split_bands uses"""

    def __init__(self, a):
        self.a = a
        self.shape = a.shape

    def __getitem__(self, k):
        return _Np(self.a[k])

    def contiguous(self):
        return np.ascontiguousarray(self.a)


def test_band_geometry_and_container():
    import lic360_shard as sh
    assert sh.band_rows(64, 4) == [(0, 16), (16, 32), (32, 48), (48, 64)]
    for bad in ((64, 0), (64, 3), (12, 4), (10, 2)):  # not a divisor / odd band height
        with pytest.raises(ValueError):
            sh.band_rows(*bad)
    assert sh.rank_bands(8, 3, 1) == [1, 4, 7] and sh.rank_bands(2, 4, 3) == []
    assert sorted(sum((sh.rank_bands(8, 3, r) for r in range(3)), [])) == list(range(8))
    streams = [(b"ab", b"cde"), (b"", b"x"), (b"\x00" * 5, b"")]
    blob = sh.pack_band_streams(64, 128, streams)
    assert sh.unpack_band_streams(blob) == (64, 128, streams)
    for broken in (blob[:-1], blob + b"z", b"XXXX" + blob[4:], blob[:10]):
        with pytest.raises(ValueError):
            sh.unpack_band_streams(broken)
    with pytest.raises(RuntimeError):
        sh.exchange_band_streams({0: (b"", b"")}, 2)  # a band is missing


def _inputs():
    q, mask, lv = synthetic_latent(321, H=H, W=W)
    return q, mask, lv


def test_band_codec_single_process():
    """Each band's streams are exactly what the codec emits for that band as an image of its own; the container decodes to the image."""
    import lic360_shard as sh
    q, mask, lv = _inputs()
    bc = sh.BandCodec(_CpuBand, H, W, NB)
    blob = bc.encode(_Np(q), _Np(mask), _Np(lv))
    h, w, streams = sh.unpack_band_streams(blob)
    assert (h, w, len(streams)) == (H, W, NB)
    for b, (r0, r1) in enumerate(sh.band_rows(H, NB)):
        alone = _CpuBand(r1 - r0, W).encode(q[:, :, r0:r1], mask[:, :, r0:r1], lv[:, :, r0 // 2:r1 // 2])
        assert streams[b] == alone
    code, m = bc.decode(blob)
    assert np.array_equal(code, q * mask) and np.array_equal(m, mask)
    with pytest.raises(ValueError):
        sh.BandCodec(_CpuBand, H, W, 4).decode(blob)  # 2-row bands: a different partition


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lic360_shard as sh
        q, mask, lv = _inputs()
        bc = sh.BandCodec(_CpuBand, H, W, NB, world=world, rank=rank)
        blob = bc.encode(_Np(q), _Np(mask), _Np(lv))
        code, m = bc.decode(blob)
        out.put((rank, bc.mine, blob, bool(np.array_equal(code, q * mask) and np.array_equal(m, mask))))
    finally:
        dist.destroy_process_group()


def test_band_codec_two_ranks_gloo():
    """world_size 2: one band per rank, no collective while coding; both ranks end up with the single-process container and image."""
    import lic360_shard as sh
    q, mask, lv = _inputs()
    ref_blob = sh.BandCodec(_CpuBand, H, W, NB).encode(_Np(q), _Np(mask), _Np(lv))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [[0], [1]]
    assert all(r[2] == ref_blob and r[3] for r in res)
