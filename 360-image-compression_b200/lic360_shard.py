"""Image-level sharding of the codec path across the GPUs of one box (every image has its own two
bitstreams, coder state and wavefront, so the path partitions by image with NO data-path collective).

One process per GPU (torchrun / mp.spawn, like the reference's trainers, train/trainDDP_IMP_ENT.py:105-108); the only
cross-rank traffic is bookkeeping: the max-over-ranks step time and the gathered stream sizes.  Used by bench.py
(NCCL on the GPU box) and covered by a world_size-2 gloo test on CPU (tests/test_shard_gloo.py).
"""
import torch
import torch.distributed as dist


def rank_images(n_images, world, rank):
    """Indices of the images rank `rank` codes: contiguous blocks, remainder to the first ranks (16 images over
    1/2/4/8 GPUs -> 16/8/4/2 per GPU, BASELINE.json configs[2])."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank_images: bad world/rank %r/%r" % (world, rank))
    base, rem = divmod(int(n_images), world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def max_over_ranks(value, device="cpu"):
    """Step time as the driver wants it: the slowest rank's."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_stream_sizes(local, n_images, device="cpu"):
    """local: {image index: (imp_bytes, code_bytes)} of this rank -> the full table on every rank.
    Each image must be reported by exactly one rank (checked)."""
    tab = torch.zeros((n_images, 3), dtype=torch.int64, device=device)
    for i, (a, b) in local.items():
        tab[i, 0] = int(a)
        tab[i, 1] = int(b)
        tab[i, 2] = 1
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tab, op=dist.ReduceOp.SUM)
    owners = tab[:, 2].tolist()
    if any(o != 1 for o in owners):
        raise RuntimeError("gather_stream_sizes: images coded by %s ranks (expected exactly one each)" % owners)
    return [(int(a), int(b)) for a, b, _ in tab.tolist()]


# ---------------------------------------------------------------------------------------------------------------------
# Latitude bands of ONE image (BASELINE.json configs[3], SURVEY.md s8e option 3).
#
# Inside one image the wavefront is strictly sequential and the reference's bitstream is ONE arithmetic-coded stream per
# network, so a single image does not shard in the reference's format ("replicas only").  What does shard is a FORMAT
# EXTENSION: the latent is cut into `nbands` horizontal (latitude) bands and every band is coded as an independent
# sub-image with its own two bitstreams -- the context of a band's first rows is the zero padding an image border has.
# Every band is exactly what the reference produces when the band is handed to it as an image of its own (that is the
# parity statement, per band); the total size differs from the single-stream size by the context lost at the band edges.
# Bands are dealt to the ranks round-robin; the codec path has no collective, the ranks only exchange the finished byte
# strings (encode) and the decoded rows (decode) -- assembly, not coding.
# ---------------------------------------------------------------------------------------------------------------------
import struct

BAND_MAGIC = b"L3B1"


def band_rows(H, nbands):
    """Row ranges [r0, r1) of the bands of an H-row latent.  Rows come in pairs (one importance row covers two latent
    rows), so H / nbands must be an even integer."""
    if nbands < 1 or H % nbands or (H // nbands) % 2:
        raise ValueError("band_rows: %d rows do not split into %d bands of an even number of rows" % (H, nbands))
    hb = H // nbands
    return [(b * hb, (b + 1) * hb) for b in range(nbands)]


def rank_bands(nbands, world, rank):
    """Bands of rank `rank`: round-robin, so that neighbouring bands (similar content, similar cost) land on different GPUs."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank_bands: bad world/rank %r/%r" % (world, rank))
    return list(range(rank, nbands, world))


def split_bands(code, mask, imap, nbands):
    """(1,C,H,W) code and mask, (1,1,H/2,W/2) importance levels -> per band (code_b, mask_b, imap_b), contiguous copies."""
    H = code.shape[2]
    out = []
    for r0, r1 in band_rows(H, nbands):
        out.append((code[:, :, r0:r1].contiguous(), mask[:, :, r0:r1].contiguous(), imap[:, :, r0 // 2:r1 // 2].contiguous()))
    return out


def pack_band_streams(H, W, streams):
    """streams: [(imp_bytes, code_bytes)] in band order -> one container: magic, nbands, H, W, the 2*nbands lengths, the payloads."""
    head = BAND_MAGIC + struct.pack("<III", len(streams), H, W)
    head += b"".join(struct.pack("<II", len(a), len(b)) for a, b in streams)
    return head + b"".join(a + b for a, b in streams)


def unpack_band_streams(blob):
    """-> (H, W, [(imp_bytes, code_bytes)]); raises ValueError on a malformed container."""
    if len(blob) < 16 or blob[:4] != BAND_MAGIC:
        raise ValueError("unpack_band_streams: not a band container")
    nb, H, W = struct.unpack_from("<III", blob, 4)
    off = 16 + 8 * nb
    if nb < 1 or len(blob) < off:
        raise ValueError("unpack_band_streams: truncated header")
    lens = [struct.unpack_from("<II", blob, 16 + 8 * b) for b in range(nb)]
    if off + sum(a + b for a, b in lens) != len(blob):
        raise ValueError("unpack_band_streams: payload size does not match the header")
    out = []
    for a, b in lens:
        out.append((bytes(blob[off:off + a]), bytes(blob[off + a:off + a + b])))
        off += a + b
    return H, W, out


def exchange_band_streams(local, nbands):
    """local: {band: (imp_bytes, code_bytes)} of this rank -> the full list on every rank (bookkeeping traffic only)."""
    table = dict(local)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, local)
        table = {}
        for p in parts:
            for k, v in p.items():
                if k in table:
                    raise RuntimeError("exchange_band_streams: band %d coded by two ranks" % k)
                table[k] = v
    if sorted(table) != list(range(nbands)):
        raise RuntimeError("exchange_band_streams: bands %s of %d present" % (sorted(table), nbands))
    return [table[b] for b in range(nbands)]


class BandCodec(object):
    """One image as `nbands` independently coded latitude bands, dealt round-robin to the ranks of the process group (or all on
    this process without one).  `make_codec(h, w)` builds the per-band codec (FusedCodec on a GPU box; the tests pass the CPU
    rendition): anything with encode(code, mask, imap) -> (imp_bytes, code_bytes) and decode(imp_bytes, code_bytes) -> (code, mask)."""

    def __init__(self, make_codec, H, W, nbands, world=1, rank=0, in_flight=1):
        self.H, self.W, self.nbands, self.world, self.rank = H, W, nbands, world, rank
        self.rows = band_rows(H, nbands)
        self.mine = rank_bands(nbands, world, rank)
        hb = H // nbands
        self.codecs = [make_codec(hb, W) for _ in range(max(1, min(in_flight, len(self.mine))))]

    def _map(self, fn, items):
        """fn(codec, item) over this rank's items, one worker thread per codec (each codec = one band in flight)."""
        if len(self.codecs) == 1 or len(items) <= 1:
            return [fn(self.codecs[0], it) for it in items]
        import threading
        out, errs, lock, nxt = [None] * len(items), [], threading.Lock(), [0]

        def worker(cd):
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= len(items):
                    return
                try:
                    out[i] = fn(cd, items[i])
                except Exception as e:  # noqa: BLE001 -- re-raised below
                    errs.append(e)
                    return

        th = [threading.Thread(target=worker, args=(cd,)) for cd in self.codecs]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return out

    def encode_local(self, code, mask, imap):
        """Codes this rank's bands of the full-size inputs -> {band: (imp_bytes, code_bytes)}."""
        bands = split_bands(code, mask, imap, self.nbands)
        res = self._map(lambda cd, b: cd.encode(*bands[b]), self.mine)
        return dict(zip(self.mine, res))

    def encode(self, code, mask, imap):
        """-> the container (identical on every rank)."""
        return pack_band_streams(self.H, self.W, exchange_band_streams(self.encode_local(code, mask, imap), self.nbands))

    def decode_local(self, blob):
        """Decodes this rank's bands -> {band: (code_b, mask_b)}."""
        H, W, streams = unpack_band_streams(blob)
        if (H, W, len(streams)) != (self.H, self.W, self.nbands):
            raise ValueError("BandCodec.decode: container is %dx%d in %d bands, codec is %dx%d in %d" % (H, W, len(streams), self.H, self.W, self.nbands))
        res = self._map(lambda cd, b: cd.decode(*streams[b]), self.mine)
        return dict(zip(self.mine, res))

    def decode(self, blob, like=None):
        """-> (code, mask) of the whole image on every rank (the decoded bands are exchanged: assembly, not coding)."""
        local = self.decode_local(blob)
        if self.world > 1 and dist.is_available() and dist.is_initialized():
            parts = [None] * self.world
            dist.all_gather_object(parts, {b: (c.cpu(), m.cpu()) if hasattr(c, "cpu") else (c, m) for b, (c, m) in local.items()})
            table = {}
            for p in parts:
                table.update(p)
        else:
            table = local
        cat = torch.cat if isinstance(next(iter(table.values()))[0], torch.Tensor) else None
        if cat is None:
            import numpy as np
            return (np.concatenate([table[b][0] for b in range(self.nbands)], axis=2),
                    np.concatenate([table[b][1] for b in range(self.nbands)], axis=2))
        dev = like.device if like is not None else next(iter(local.values()))[0].device
        return (torch.cat([table[b][0].to(dev) for b in range(self.nbands)], dim=2),
                torch.cat([table[b][1].to(dev) for b in range(self.nbands)], dim=2))
