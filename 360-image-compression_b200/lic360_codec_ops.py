"""Per-op entropy codec loops at the `lic360` op level.

Restates the four codec-form drivers of the reference demo -- EntEncoderFast (lic360_demo.py:95-141),
ImpEntEncoderFast (:143-189), EntDecoder (:191-238), ImpEntDecoder (:241-290) -- directly on the pybind-level op
API (`XOp.forward(...)`), which this repo's `lic360` mirror and the reference extension expose identically.  The
`backend` argument selects which one is driven, so the same loops produce (a) this repo's bitstreams and (b) the
reference CUDA extension's bitstreams/timings for the baseline (bench.py, tests/test_gpu_ref.py).

Network parameters are passed as a plain dict in codec form (lic360_demo.py:296-322): for the code stream
  net.0.{weight,bias,relu}, net.{1..5}.conv{1,2}.{weight,bias,relu}, net.6.{weight,bias}
with a leading dim of 3 = [weight_net, delta_net, mean_net]; for the importance stream the same keys unstacked.
"""
import torch


def make_entropy_params(ngroup, cpg, nlast, batch, seed, device, delta_bias=2.0):
    """Seeded random-init parameters in codec form ("random-init model-idx 3", SURVEY.md s7 hard part 7):
    He-normal weights (MaskConstrain.py:31), zero bias, PReLU slope 0.25 (nn.PReLU default), delta-net last bias 2
    (test/model_zoo.py:263)."""
    g = torch.Generator().manual_seed(seed)
    lead = (batch,) if batch else ()

    def conv(cin, cout, act):
        fan_in = ngroup * cin * 25
        w = torch.randn(lead + (ngroup * cout, ngroup * cin, 5, 5), generator=g) * (2.0 / fan_in) ** 0.5
        p = {'weight': w, 'bias': torch.zeros(lead + (ngroup * cout,))}
        if act:
            p['relu'] = torch.full(lead + (ngroup * cout,), 0.25)
        return p

    params = {}
    for k, v in conv(1, cpg, True).items():
        params['net.0.' + k] = v
    for b in range(1, 6):
        for cname in ('conv1', 'conv2'):
            for k, v in conv(cpg, cpg, True).items():
                params['net.%d.%s.%s' % (b, cname, k)] = v
    for k, v in conv(cpg, nlast, False).items():
        params['net.6.' + k] = v
    if batch:
        params['net.6.bias'][1] = delta_bias  # [weight_net, delta_net, mean_net]
    return {k: v.to(device).contiguous() for k, v in params.items()}


class _Net(object):
    """12 context convs (+5 residual adds) in EC (whole frame) or DC (wavefront) form."""

    def __init__(self, backend, params, ngroup, cpg, nlast, batch, dc, gid):
        self.p, self.batch, self.dc = params, batch, dc
        Op = backend.CconvDcOp if dc else backend.CconvEcOp
        mk = lambda cin, cout, constrain: Op(ngroup * cin, ngroup, ngroup * cout, 5, constrain, gid, False)
        self.first = mk(1, cpg, 5)
        self.blocks = [(mk(cpg, cpg, 6), mk(cpg, cpg, 6)) for _ in range(5)]
        self.last = mk(cpg, nlast, 6)
        self.adds = [backend.TileAddOp(ngroup, gid, False) for _ in range(5)] if dc else None

    def stateful(self):
        ops = []
        if self.dc:
            ops = [self.first, self.last] + [c for b in self.blocks for c in b] + list(self.adds)
        return ops

    def _conv(self, op, x, key, act):
        w, b = self.p[key + '.weight'], self.p[key + '.bias']
        sfx = '_batch' if self.batch else ''
        if act:
            return getattr(op, 'forward_act' + sfx)(x, w, b, self.p[key + '.relu'])[0]
        return getattr(op, 'forward' + sfx)(x, w, b)[0]

    def __call__(self, x):
        y = self._conv(self.first, x, 'net.0', True)
        for i, (c1, c2) in enumerate(self.blocks):
            t = self._conv(c2, self._conv(c1, y, 'net.%d.conv1' % (i + 1), True), 'net.%d.conv2' % (i + 1), True)
            y = self.adds[i].forward(t, y)[0] if self.dc else t + y
        return self._conv(self.last, y, 'net.6', False)


def _plan(backend, ref_tensor, gid, ops):
    ctx = backend.CodeContexOp(gid, False)
    p1, p2 = ctx.forward(ref_tensor)
    for op in ops:
        op.set_param(p1, p2)
        op.restart()
    return ctx


class EntEncoder(object):
    """Code stream encoder, lic360_demo.py:95-141."""

    def __init__(self, backend, params, ngroup=48, bin_num=8, gid=0):
        self.b, self.ngroup, self.gid = backend, ngroup, gid
        self.bias = (bin_num - 1) / 2.
        self.net = _Net(backend, params, ngroup, 4, 3, 3, False, gid)
        self.ext = backend.TileExtractOp(ngroup, True, gid, False)
        self.ext_label = backend.TileExtractOp(ngroup, True, gid, False)
        self.ext_mask = backend.TileExtractOp(ngroup, True, gid, False)
        self.gmm = backend.EntropyGmmTableOp(bin_num, self.bias, 3, 65536, 1e-6, gid, False)
        self.mcoder = backend.Coder('tmp', 3.5)

    @torch.no_grad()
    def encode(self, data, mask, fname):
        self._ctx = _plan(self.b, data, self.gid, [self.ext, self.ext_label, self.ext_mask])
        self.mcoder.reset_fname(fname)
        self.mcoder.start_encoder()
        h, w = data.shape[2:]
        tdata = ((data - self.bias) * mask).contiguous()
        y = self.net(torch.cat([tdata, tdata, tdata], dim=0).contiguous())
        for _ in range(h + w + self.ngroup - 2):
            z, le = self.ext.forward_batch(y)
            vec = self.gmm.forward_batch(z, le)[0]
            ln = int(le[0].item())
            label = self.ext_label.forward(data)[0]
            tm = self.ext_mask.forward(mask)[0]
            pred, tlabel, tm = vec.type(torch.int32).to('cpu'), label.type(torch.int32).to('cpu'), tm.type(torch.float32).to('cpu').contiguous()
            self.mcoder.encodes_mask(pred, 8, tlabel, tm, ln)
        self.mcoder.end_encoder()


class ImpEntEncoder(object):
    """Importance stream encoder, lic360_demo.py:143-189."""

    def __init__(self, backend, params, bin_num=48, gid=0):
        self.b, self.gid, self.bin_num = backend, gid, bin_num
        self.net = _Net(backend, params, 1, bin_num * 3, bin_num + 1, None, False, gid)
        self.ext = backend.TileExtractOp(1, True, gid, False)
        self.ext_label = backend.TileExtractOp(1, True, gid, False)
        self.table = backend.EntropyTableOp(bin_num + 1, 65536, gid, False)
        self.scale = backend.ScaleOp(-1, float(2. / (bin_num - 1.)), gid, False)
        self.mcoder = backend.Coder('tmp', 3.5)

    @torch.no_grad()
    def encode(self, data, fname):
        data = data.contiguous()
        self._ctx = _plan(self.b, data, self.gid, [self.ext, self.ext_label])
        self.mcoder.reset_fname(fname)
        self.mcoder.start_encoder()
        h, w = data.shape[2:]
        n1 = self.bin_num + 2
        y = self.net(self.scale.forward(data)[0])
        for _ in range(h + w + 1 - 2):
            z, le = self.ext.forward(y)
            vec = self.table.forward(z, le)[0]
            ln = int(le[0].item())
            label = self.ext_label.forward(data)[0]
            pred, tlabel = vec.view(-1, n1).type(torch.int32).to('cpu'), label.view(-1).type(torch.int32).to('cpu')
            self.mcoder.encodes(pred, n1 - 1, tlabel, ln)
        self.mcoder.end_encoder()


class EntDecoder(object):
    """Code stream decoder, lic360_demo.py:191-238."""

    def __init__(self, backend, params, ngroup=48, bin_num=8, gid=0):
        self.b, self.ngroup, self.gid = backend, ngroup, gid
        self.bias = (bin_num - 1) / 2.
        self.ipt = backend.TileInputOp(ngroup, -3.5, 1, 3, gid, False)
        self.net = _Net(backend, params, ngroup, 4, 3, 3, True, gid)
        self.ext = backend.TileExtractOp(ngroup, True, gid, False)
        self.ext_mask = backend.TileExtractOp(ngroup, True, gid, False)
        self.gmm = backend.EntropyGmmTableOp(bin_num, self.bias, 3, 65536, 1e-6, gid, False)
        self.mcoder = backend.Coder('tmp', 3.5)

    @torch.no_grad()
    def decode(self, mask, fname):
        h, w = mask.shape[2:]
        dev = mask.device
        pout = torch.zeros((1, 1, h, w), dtype=torch.float32, device=dev)
        self._ctx = _plan(self.b, pout, self.gid, [self.ipt, self.ext, self.ext_mask] + self.net.stateful())
        self.mcoder.reset_fname(fname)
        self.mcoder.start_decoder()
        for _ in range(h + w + self.ngroup - 2):
            b = self.ipt.forward(pout)[0]
            y = self.net(b)
            z, le = self.ext.forward_batch(y)
            vec = self.gmm.forward_batch(z, le)[0]
            ln = int(le[0].item())
            mt = self.ext_mask.forward(mask)[0]
            pred, mt = vec.type(torch.int32).to('cpu').view(-1, 9), mt.to('cpu').contiguous()
            pout = self.mcoder.decodes_mask(pred, 8, mt, ln).to(dev).view(1, 1, h, w).contiguous()
        b = self.ipt.forward(pout)[0]
        return (b[0:1] + self.bias * mask).contiguous()


class ImpEntDecoder(object):
    """Importance stream decoder, lic360_demo.py:241-290 (returns the importance levels, before Imp2mask/Dtow)."""

    def __init__(self, backend, params, bin_num=48, gid=0):
        self.b, self.gid, self.bin_num = backend, gid, bin_num
        self.scale = float(2. / (bin_num - 1))
        self.ipt = backend.TileInputOp(1, -1, self.scale, 1, gid, False)
        self.net = _Net(backend, params, 1, bin_num * 3, bin_num + 1, None, True, gid)
        self.ext = backend.TileExtractOp(1, True, gid, False)
        self.table = backend.EntropyTableOp(bin_num + 1, 65536, gid, False)
        self.mcoder = backend.Coder('tmp', 3.5)

    @torch.no_grad()
    def decode(self, fname, h=32, w=64, device='cuda:0'):
        n1 = self.bin_num + 2
        pout = torch.zeros((1, 1, h, w), dtype=torch.float32, device=device)
        self._ctx = _plan(self.b, pout, self.gid, [self.ipt, self.ext] + self.net.stateful())
        self.mcoder.reset_fname(fname)
        self.mcoder.start_decoder()
        for _ in range(h + w + 1 - 2):
            b = self.ipt.forward(pout)[0]
            y = self.net(b)
            z, le = self.ext.forward(y)
            vec = self.table.forward(z, le)[0]
            ln = int(le[0].item())
            pred = vec.type(torch.int32).to('cpu').view(-1, n1)
            pout = self.mcoder.decodes(pred, n1 - 1, ln).to(device).view(1, 1, h, w).contiguous()
        b = self.ipt.forward(pout)[0]
        code = ((b + 1) / self.scale).contiguous()
        return torch.floor(code + 1e-5).type(torch.float32).contiguous()
