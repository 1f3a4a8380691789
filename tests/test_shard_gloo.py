"""N>1 host logic on CPU: world_size-2 gloo run of the image-sharded codec path (no GPU).  Each rank codes its shard of
tiny synthetic latents with the CPU rendition (oracle/cpu_codec.py -- test infrastructure standing in for the GPU
codec), the ranks exchange only bookkeeping (max step time, stream sizes), and the result equals a single-process run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import synthetic_latent

N_IMAGES, H, W = 3, 4, 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _codec():
    import lic360_codec_ops as ops
    from oracle import cpu_codec
    params = {'code': ops.make_entropy_params(48, 4, 3, 3, 7, 'cpu'), 'imp': ops.make_entropy_params(1, 144, 49, None, 8, 'cpu')}
    return cpu_codec.CpuCodec(cpu_codec.params_to_numpy(params))


def _code_images(indices):
    codec = _codec()
    out = {}
    for i in indices:
        q, mask, lv = synthetic_latent(100 + i, H=H, W=W)
        bi, bc = codec.encode(q, mask, lv)
        code, m = codec.decode(bi, bc, H // 2, W // 2)
        assert np.array_equal(code, q * mask) and np.array_equal(m, mask)
        out[i] = (len(bi), len(bc))
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lic360_shard as sh
        mine = sh.rank_images(N_IMAGES, world, rank)
        local = _code_images(mine)
        slowest = sh.max_over_ranks(10.0 + rank)
        table = sh.gather_stream_sizes(local, N_IMAGES)
        q.put((rank, mine, slowest, table))
    finally:
        dist.destroy_process_group()


def test_rank_images_partition():
    import lic360_shard as sh
    for n in (0, 1, 3, 16):
        for world in (1, 2, 4, 8):
            parts = [sh.rank_images(n, world, r) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert [len(sh.rank_images(16, w, 0)) for w in (1, 2, 4, 8)] == [16, 8, 4, 2]
    with pytest.raises(ValueError):
        sh.rank_images(4, 2, 2)


def test_world2_gloo_sharded_codec():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _code_images(range(N_IMAGES))
    assert sorted(res[0][1] + res[1][1]) == list(range(N_IMAGES))         # every image coded by exactly one rank
    for rank, mine, slowest, table in res:
        assert slowest == 11.0                                            # max over ranks, on every rank
        assert table == [single[i] for i in range(N_IMAGES)]              # same bytes as the unsharded run


def test_gather_rejects_double_coding():
    import lic360_shard as sh
    with pytest.raises(RuntimeError):
        sh.gather_stream_sizes({0: (1, 2)}, 2)
