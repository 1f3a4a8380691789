"""Module layer (`lic360_operator`) on the GPU: the nn.Module classes a reference user instantiates, driven the way
train/model_zoo.py (training form) and test/lic360_demo.py (wavefront form) drive them.
 * state dicts with the reference's parameter names load into the module nets; the module nets reproduce the per-op nets bit for bit
   and the wavefront (CconvDcBatch + TileAdd) form reproduces the whole-frame (CconvEcBatch) form bit for bit;
 * the training form MaskConv2 -> ContextReshape -> EntropyGmm back-propagates the same gradients as a float64 torch autograd
   formulation of the same loss (tolerance 1e-5 relative, the float tier of BASELINE.json's north_star);
 * SphereLatScaleNet / SpherePad / SphereCutEdge / Dtow / ContextShift autograd plumbing against torch formulations;
 * a CPU tensor is an error (there is no CPU path)."""
import numpy as np
import pytest
import torch
from torch import nn

from util import n, t, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


class _ResEc(nn.Module):
    def __init__(self, mk):
        super().__init__()
        self.conv1, self.conv2 = mk(), mk()

    def forward(self, x):
        return self.conv2(self.conv1(x)) + x


class _ResDc(nn.Module):
    def __init__(self, mk, add):
        super().__init__()
        self.conv1, self.conv2, self.add = mk(), mk(), add

    def forward(self, x):
        return self.add(self.conv2(self.conv1(x)), x)


def _module_net(M, ngroup, cpg, nlast, batch, dc):
    """The 12-conv context net of lic360_demo.py:95-141 (EC) / :191-238 (DC) built from the module classes; keys net.0, net.1.conv1 ..."""
    if batch:
        Conv = M.CconvDcBatch if dc else M.CconvEcBatch
        mk = lambda cin, cout, hidden, act: Conv(ngroup, cin, cout, 5, batch, hidden, act)
    else:
        Conv = M.CconvDc if dc else M.CconvEc
        mk = lambda cin, cout, hidden, act: Conv(ngroup, cin, cout, 5, hidden, act)
    hid = lambda: mk(cpg, cpg, True, True)
    blocks = [(_ResDc(hid, M.TileAdd(ngroup)) if dc else _ResEc(hid)) for _ in range(5)]
    net = nn.Sequential(mk(1, cpg, False, True), *blocks, mk(cpg, nlast, True, False))
    holder = nn.Module()
    holder.net = net
    return holder


def _stateful(holder):
    import lic360_operator as M
    return [m for m in holder.modules() if isinstance(m, (M.CconvDc, M.CconvDcBatch, M.TileAdd))]


@pytest.mark.parametrize("ngroup,cpg,nlast,batch,H,W", [(12, 4, 3, 3, 7, 9), (1, 24, 9, None, 6, 10)])
def test_module_nets_load_state_dict_and_match_per_op_nets(ngroup, cpg, nlast, batch, H, W):
    import lic360
    import lic360_operator as M
    import lic360_codec_ops as ops
    params = ops.make_entropy_params(ngroup, cpg, nlast, batch, 11, DEV)
    ec, dc = _module_net(M, ngroup, cpg, nlast, batch, False), _module_net(M, ngroup, cpg, nlast, batch, True)
    for holder in (ec, dc):
        holder.to(DEV)
        missing, unexpected = holder.load_state_dict(params, strict=True)
        assert not missing and not unexpected
    # the last conv has no PReLU and therefore no `relu` parameter (CconvEc.py:70, lic360_demo.py:296-322)
    assert "net.6.relu" not in ec.state_dict() and "net.0.relu" in ec.state_dict()
    r = np.random.default_rng(3)
    x1 = (r.integers(0, 8, (1, ngroup, H, W)) - 3.5).astype(np.float32)
    x = t(np.concatenate([x1] * (batch or 1)), DEV)
    with torch.no_grad():
        y_mod = n(ec.net(x)).copy()
        y_op = n(ops._Net(lic360, params, ngroup, cpg, nlast, batch, False, 0)(x)).copy()
        assert np.array_equal(y_mod.view(np.int32), y_op.view(np.int32))
        ctx = M.CodeContex()
        p1, p2 = ctx(torch.zeros((1, 1, H, W), device=DEV))
        for m in _stateful(dc):
            m.set_param(p1, p2)
            m.restart()
        for _ in range(H + W + ngroup - 2):
            y_dc = dc.net(x)
        assert np.array_equal(n(y_dc).view(np.int32), y_mod.view(np.int32))
        # restart() rewinds the wavefront: a second pass over the same input gives the same frame
        for m in _stateful(dc):
            m.restart()
        for _ in range(H + W + ngroup - 2):
            y_dc = dc.net(x)
        assert np.array_equal(n(y_dc).view(np.int32), y_mod.view(np.int32))


def _phi(z):
    return 0.5 + 0.5 * torch.erf(z / 2 ** 0.5)


def _gmm_loss64(w, d, m, l):
    """EntropyGmm (entropy_gmm_cuda.cu:36-68) as a differentiable float64 expression."""
    p = (w * (_phi((l + 0.5 - m) / d) - _phi((l - 0.5 - m) / d))).sum(1)
    return -torch.log(p + 1e-7), p


def _rows(x, G):
    """ContextReshape (context_reshape_cuda.cu:30-41): (N, G*cpg, H, W) -> (N*G*H*W, cpg)."""
    N, C, H, W = x.shape
    return x.view(N, G, C // G, H, W).permute(0, 1, 3, 4, 2).reshape(-1, C // G)


def test_training_form_gradients_match_float64_autograd():
    """train/model_zoo.py entropy model in miniature: three masked-conv stacks (weight / delta / mean) -> ContextReshape -> EntropyGmm."""
    import lic360_operator as M
    torch.backends.cudnn.allow_tf32 = False      # the convolutions are torch's on both sides; keep them fp32 so 1e-5 is about OUR ops
    from oracle import oracle as O
    G, cpg, ng, N, H, W = 4, 3, 3, 2, 9, 12
    torch.manual_seed(5)
    r = np.random.default_rng(5)

    def stack():
        return nn.ModuleList([M.MaskConv2(G, 1, cpg, 5, False), M.MaskConv2(G, cpg, ng, 5, True)]).to(DEV)

    nets = [stack() for _ in range(3)]
    for s in nets:
        for conv in s:
            conv.bias.data.normal_(0, 0.1)
    raw = {k: [(c.weight.detach().clone(), c.bias.detach().clone()) for c in s] for k, s in zip("wdm", nets)}
    label = t(r.integers(0, 8, (N, G, H, W)).astype(np.float32), DEV)
    xin = ((label - 3.5) / 4).requires_grad_()

    def head(y, k):
        if k == "w":
            return torch.softmax(y, 1)
        if k == "d":
            return torch.nn.functional.softplus(y) + 0.6
        return 3.5 + 2 * torch.tanh(y)

    # -- reference formulation: the SAME fp32 cuDNN convolutions on explicitly masked weights (MaskConstrain semantics from the
    #    oracle), then everything this repository implements (reshape, mixture likelihood) as float64 torch autograd
    x64 = xin.detach().clone().requires_grad_()
    lab64 = label.double()
    leaves, heads64 = {}, {}
    for k in "wdm":
        y = x64
        leaves[k] = []
        for li, (w, b) in enumerate(raw[k]):
            mask = t((O.mask_constrain(np.ones(tuple(w.shape), np.float32), G, 5 if li == 0 else 6) != 0).astype(np.float32), DEV)
            w64, b64 = (w * mask).requires_grad_(), b.clone().requires_grad_()   # leaf = the masked weight
            leaves[k].append((w64, b64, mask))
            y = nn.functional.conv2d(y, w64, b64, padding=2)
            if li == 0:
                y = nn.functional.leaky_relu(y, 0.2)
        heads64[k] = _rows(y.double(), G)
    w_r, d_r, m_r = (head(heads64[k], k) for k in "wdm")
    loss64, p64 = _gmm_loss64(w_r, d_r, m_r, _rows(lab64, G))
    keep = (p64 > 0.02).detach()          # symbols whose probability float32 resolves (fb - fa cancels below that)
    assert keep.double().mean() > 0.5
    top = keep.double() * t(r.uniform(0.5, 1.5, keep.numel()), DEV)
    (loss64 * top).sum().backward()

    # -- module formulation
    # one ContextReshape per use, as in the reference models: an op instance remembers the shape of its LAST forward for its
    # backward (base_opt.hpp:43-72), so sharing one between tensors of different shapes is a caller error there too
    reshape, ent = {k: M.ContextReshape(G) for k in "wdml"}, M.EntropyGmm(ng, 0)
    outs = {}
    for k, s in zip("wdm", nets):
        y = nn.functional.leaky_relu(s[0](xin), 0.2)
        outs[k] = head(reshape[k](s[1](y)), k)
    lab_rows = reshape["l"](label)
    loss = ent(outs["w"], outs["d"], outs["m"], lab_rows)
    assert loss.shape[0] == N * G * H * W
    (loss.view(-1) * top.float()).sum().backward()

    kk = keep.cpu().numpy()
    assert rel_err(n(loss).reshape(-1)[kk], n(loss64)[kk]) < TOL
    assert rel_err(n(xin.grad), n(x64.grad)) < 5 * TOL
    for k, s in zip("wdm", nets):
        for conv, (w64, b64, mask) in zip(s, leaves[k]):
            # MaskConv2 masks the stored weights in place on every forward (MaskConstrain.py:36); the masking runs on .data, outside
            # autograd, so the weight gradient is the dense conv2d gradient (masked taps are re-zeroed by the next forward)
            assert np.all(n(conv.weight.data)[n(mask) == 0] == 0) and np.any(n(conv.weight.data)[n(mask) != 0] != 0)
            assert rel_err(n(conv.weight.grad), n(w64.grad)) < 5 * TOL, k
            assert rel_err(n(conv.bias.grad), n(b64.grad)) < 5 * TOL, k


def test_sphere_lat_scale_net_gradients():
    import lic360_operator as M
    npart, N, C, H, W = 8, 2, 5, 32, 24
    torch.manual_seed(2)
    mod = M.SphereLatScaleNet(npart)
    x = torch.randn(N, C, H, W, device=DEV, requires_grad=True)
    g = torch.randn(N, C, H, W, device=DEV)
    out = mod(x)
    out.backward(g)
    got = [x.grad.clone()] + [p.grad.clone() for p in mod.net.parameters()]
    x.grad = None
    mod.zero_grad()
    wrow = mod.net(mod.data.data).view(npart).repeat_interleave(H // npart)
    ref = x * wrow.view(1, 1, H, 1)
    assert rel_err(n(out), n(ref)) < TOL
    ref.backward(g)
    exp = [x.grad.clone()] + [p.grad.clone() for p in mod.net.parameters()]
    for a, b in zip(got, exp):
        assert rel_err(n(a), n(b)) < 2e-5
    assert mod.data.requires_grad is False and tuple(mod.data.shape) == (1, 1, npart)


def test_layout_modules_backward_is_the_adjoint():
    """SpherePad / SphereCutEdge / Dtow / ContextShift / ContextReshape are linear maps: <A x, g> == <x, A^T g> with A^T the module's
    backward (SpherePad.py:7-22, SphereCutEdge.py:7-22, Dtow.py:6-19, ContextShift.py:6-20, ContextReshape.py:6-20)."""
    import lic360_operator as M
    torch.manual_seed(9)
    cases = [(M.SpherePad(2), (2, 3, 8, 16)), (M.SphereCutEdge(2), (2, 3, 12, 20)), (M.Dtow(2, False), (1, 8, 6, 10)),
             (M.Dtow(2, True), (1, 8, 6, 10)), (M.ContextShift(False, 1), (1, 6, 5, 7)), (M.ContextReshape(4), (2, 12, 5, 7))]
    for mod, shape in cases:
        x = torch.randn(*shape, device=DEV, dtype=torch.float32, requires_grad=True)
        y = mod(x)
        g = torch.randn_like(y)
        lhs = float((y.double() * g.double()).sum())
        y.backward(g)
        rhs = float((x.detach().double() * x.grad.double()).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0), (type(mod).__name__, lhs, rhs)


def test_quant_module_forward_backward_matches_op():
    """QUANT.py:31-44: the module's parameters (log-width weights, running count) and its Function give what the bound op gives."""
    import lic360
    import lic360_operator as M
    C, bins = 6, 8
    mod = M.QUANT(C, bins, check_iters=100, weight_decay=0.9, ntop=1, top_alpha=0.1)
    assert tuple(mod.weight.shape) == (C, bins) and tuple(mod.count.shape) == (C, bins)
    assert abs(float(mod.weight[0, 0]) - 1 / 9) < 1e-7 and abs(float(mod.weight[0, 1]) - np.log(1 / 9)) < 1e-6
    torch.manual_seed(4)
    x = (torch.rand(2, C, 8, 12, device=DEV) * 1.2 - 0.1).requires_grad_()
    g = torch.randn(2, C, 8, 12, device=DEV)
    mod.train()
    y = mod(x)
    y.backward(g)
    op = lic360.QuantOp(C, bins, 0.9, 100, 1, 0.1, 0, False)
    w = mod.weight.detach().clone()
    cnt = torch.zeros(C, bins, device=DEV)
    y2 = op.forward(x.detach(), w, cnt, True)[0]
    o = op.backward([g], x.detach(), y2)
    assert torch.equal(y, y2) and torch.equal(x.grad, o[0])
    assert rel_err(n(mod.weight.grad), n(o[1])) < TOL   # per-level sums are float atomics: order-dependent in the last bits
    # the running histogram travels as the "gradient" of `count` (QUANT.py:27-29), negated so that SGD accumulates it (quant_cuda.cu:149)
    assert torch.equal(mod.count.grad, o[2]) and float(o[2].sum()) == -x.numel()
    lv = torch.unique(y.detach())
    assert lv.numel() <= bins * C


def test_modules_reject_cpu_tensors():
    import lic360_operator as M
    with pytest.raises(RuntimeError, match="no CPU path"):
        M.ContextReshape(4)(torch.zeros(1, 4, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        M.CconvEc(2, 1, 4, 5)(torch.zeros(1, 2, 4, 4))
