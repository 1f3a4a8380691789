"""autograd shims between torch and the native ops (one per reference `*_AF` class)."""
import torch


def _op_of(op, t):
    gid = t.device.index
    if gid is None:
        raise RuntimeError("lic360_operator: CUDA tensor expected (there is no CPU path)")
    if gid not in op:
        raise RuntimeError("lic360_operator: module is bound to GPU %s but got a tensor on cuda:%d -- "
                           "move the module with .to()" % (sorted(op), gid))
    return op[gid]


class ForwardOnly(torch.autograd.Function):
    """forward = op.<method>(*tensors); no gradient (CconvEc.py:6-58, CconvDc.py:6-56, TileInput.py:6-19,
    Imp2mask.py:6-17, EntropyTable.py:6-18, EntropyGmmTable.py:6-21,34-47, Dquant.py:7-18)."""

    @staticmethod
    def forward(ctx, op, method, nout, *tensors):
        outs = getattr(_op_of(op, tensors[0]), method)(*tensors)
        ctx.nin = len(tensors)
        for o in outs[:nout]:
            if isinstance(o, torch.Tensor):
                ctx.mark_non_differentiable(o)
        return outs[0] if nout == 1 else tuple(outs[:nout])

    @staticmethod
    def backward(ctx, *grads):
        return (None,) * (3 + ctx.nin)


class FwdBwd(torch.autograd.Function):
    """forward = op.forward(x, *extra), backward = op.backward(grad, *extra) -> grad wrt x
    (SpherePad.py:7-22, SphereCutEdge.py:7-22, ContextReshape.py:6-20, ContextShift.py:6-20, Dtow.py:6-19,
    Scale.py:6-19 -- whose bound op has no backward, so a Scale gradient raises exactly like the reference)."""

    @staticmethod
    def forward(ctx, op, inplace, x):
        # In-place ops (SpherePad(inplace), SphereTrim) hand back x itself.  Like the reference's *_AF classes they do NOT tell
        # autograd (no mark_dirty): the reference's models pad / trim tensors in place that a conv has already saved for its
        # backward (train/model_zoo.py:15-30 and the like), which a version bump would turn into "modified by an inplace
        # operation" errors -- the train/ scripts have to run unchanged, so the module layer keeps the reference's semantics.
        ctx.op = op
        return _op_of(op, x).forward(x)[0]

    @staticmethod
    def backward(ctx, grad):
        grad = grad if grad.is_contiguous() else grad.contiguous()
        return None, None, _op_of(ctx.op, grad).backward(grad)[0]


class MaskConstrainFn(torch.autograd.Function):
    """MaskConstrain.py:7-22: in-place masking of the weights (forward) and of the incoming gradient (backward)."""

    @staticmethod
    def forward(ctx, x, op):
        _op_of(op, x).forward(x)
        ctx.op = op
        ctx.mark_dirty(x)
        return x

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous()
        _op_of(ctx.op, grad).backward(grad)
        return grad, None


class TileAddFn(torch.autograd.Function):
    """TileAdd.py:6-17."""

    @staticmethod
    def forward(ctx, x, y, op):
        out = _op_of(op, x).forward(x, y)[0]
        ctx.mark_dirty(x)
        return out

    @staticmethod
    def backward(ctx, grad):
        return None, None, None


class EntropyGmmFn(torch.autograd.Function):
    """EntropyGmm.py:6-20: the four gradients are produced by the op (cached in its forward)."""

    @staticmethod
    def forward(ctx, weight, delta, mean, label, op):
        ctx.op = op
        return _op_of(op, weight).forward(weight, delta, mean, label)[0]

    @staticmethod
    def backward(ctx, grad):
        o = _op_of(ctx.op, grad).backward(grad.contiguous())
        return o[0], o[1], o[2], o[3], None


class QuantFn(torch.autograd.Function):
    """QUANT.py:7-29."""

    @staticmethod
    def forward(ctx, x, weight, count, op, training):
        x = x if x.is_contiguous() else x.contiguous()
        outs = _op_of(op, x).forward(x, weight, count, training)
        ctx.save_for_backward(x, outs[0])
        ctx.op = op
        return outs[0] if len(outs) == 1 else (outs[0], outs[1])

    @staticmethod
    def backward(ctx, *grads):
        grads = [g if g.is_contiguous() else g.contiguous() for g in grads]
        x, out = ctx.saved_tensors
        o = _op_of(ctx.op, x).backward(grads, x, out)
        return o[0], o[1], o[2].clone().detach(), None, None


class ImpMapFn(torch.autograd.Function):
    """ImpMap.py:8-57 (IMP_MAP_AF for ntop == 1, IMP_MAP_AF2 otherwise)."""

    @staticmethod
    def forward(ctx, x, imp, level, op, ntop):
        imp = torch.floor(imp * level) / level
        outs = _op_of(op, x).forward(x, imp)
        ctx.op = op
        ctx.save_for_backward(imp, outs[1])
        rt = torch.mean(imp)
        return (outs[0], outs[2], rt) if ntop > 1 else (outs[0], rt)

    @staticmethod
    def backward(ctx, *grads):
        imp, constrain = ctx.saved_tensors
        new_constrain = torch.mean(imp, dim=3) - constrain
        o = _op_of(ctx.op, grads[0]).backward(grads[0].contiguous(), imp, new_constrain.contiguous())
        return o[0], o[1], None, None, None


class SphereLatScaleFn(torch.autograd.Function):
    """SphereLatScaleNet.py:7-23: data gradient from the op, band-weight gradient reduced here."""

    @staticmethod
    def forward(ctx, x, weight, op):
        ctx.op = op
        ctx.save_for_backward(weight, x)
        return _op_of(op, x).forward(x, weight)[0]

    @staticmethod
    def backward(ctx, grad):
        weight, data = ctx.saved_tensors
        gx = _op_of(ctx.op, grad).backward(grad.contiguous(), weight)[0]
        per_row = torch.sum(grad * data, (0, 1, 3))
        gw = torch.sum(per_row.view(weight.size(-1), -1), 1).view_as(weight)
        return gx, gw, None
