"""`lic360` -- host-side mirror of the reference's pybind11 module (extension/main.cpp:4-178).

Same class names, constructor signatures, methods and ownership rules as the reference extension, so that
`lic360_operator/*.py`, `test/lic360_demo.py` and the `train/` scripts run unchanged.  Every method forwards raw
device pointers to the C-ABI of `liblic360_b200.so` (include/lic360_b200.h); there is no torch type below this file
and no CPU fallback: importing this module without the built library raises ImportError, calling an op with a
CPU tensor raises RuntimeError.

Ownership (base_opt.hpp:43-72): each op instance owns its output tensors and re-allocates them only when the
input shape changes; returned tensors ALIAS them and are overwritten by the next call of the same op.
"""
import ctypes
import os

import torch

from ._lib import LIB, check, cstream, ptr, ptr_or_null

__all__ = [
    "CconvDcOp", "CconvEcOp", "CodeContexOp", "Coder", "ContexShiftOp", "ContextReshapeOp", "DquantOp", "DtowOp", "ProjectsOp",
    "EntropyGmmOp", "EntropyGmmTableOp", "EntropyTableOp", "Imp2maskOp", "ImpMapOp", "MaskConstrainOp", "QuantOp",
    "ScaleOp", "SphereCutEdgeOp", "SphereLatScaleOp", "SpherePadOp", "SphereTrimOp", "TileAddOp", "TileExtractOp",
    "TileInputOp", "launch_count",
]


def launch_count():
    """Number of kernels launched by the native library in this process."""
    return int(LIB.lic360_launch_count())


def _f32(t, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("lic360: %s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("lic360: %s must be float32 (got %s)" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _on_tensor_device(fn):
    """Run an op method with the CUDA device of its first tensor argument current (and give the caller's device back):
    the kernels are launched on `cstream(x)`, a stream of x's device, which is only legal while that device is current.
    The reference binds the device once per op (base_opt.hpp:20-23); here several GPUs per process are a supported mode
    (the module layer keeps {gid: op} dicts), so every call is guarded."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        dev = None
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapped


class _BaseOp(object):
    """base_opt (base_opt.hpp:4-80): device binding + cached output buffers."""

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        for name, fn in list(vars(cls).items()):
            if callable(fn) and (name.startswith("forward") or name.startswith("backward")):
                setattr(cls, name, _on_tensor_device(fn))

    def __init__(self, device=0, timeit=False):
        self.device_ = -1
        self.timeit_ = bool(timeit)
        self._top = None
        self._bottom = None
        self._shape = None
        self.to(device)

    def to(self, device):
        device = int(device)
        if device == self.device_:
            return
        self.device_ = device
        self._top = None
        self._bottom = None
        self._shape = None
        self._on_init()

    def _on_init(self):
        pass

    def _reshape(self, shape):
        shape = tuple(int(s) for s in shape)
        if shape == self._shape:
            return False
        self._shape = shape
        return True

    def _tops(self, ref, shapes):
        if self._top is None or tuple(self._top[0].shape) != tuple(shapes[0]):
            self._top = [torch.empty(s, dtype=torch.float32, device=ref.device) for s in shapes]
        return self._top

    def _bottoms(self, ref, shapes):
        if self._bottom is None or tuple(self._bottom[0].shape) != tuple(shapes[0]):
            self._bottom = [torch.empty(s, dtype=torch.float32, device=ref.device) for s in shapes]
        return self._bottom


class _PlanMixin(object):
    """set_param / restart / per-op step counter of the stateful wavefront ops (cconv_dc.hpp:21-27)."""

    def _plan_init(self):
        self.param_set_ = False
        self.plan_sum_ = 0
        self.index_mat_ = None
        self.plan_idx_mat_ = None

    def set_param(self, idx, pidx):
        if not idx.is_cuda or idx.dtype != torch.int32:
            raise RuntimeError("lic360: the index plan must be an int32 CUDA tensor")
        if pidx.is_cuda or pidx.dtype != torch.int32:
            raise RuntimeError("lic360: plane_idx must be an int32 CPU tensor")
        self.index_mat_ = idx.contiguous()
        self.plan_idx_mat_ = pidx.contiguous()
        self.param_set_ = True

    def restart(self):
        self.plan_sum_ = 0

    def _check_plan(self, H, W):
        if not self.param_set_:
            raise RuntimeError("lic360: Slice Index has not been initialized (call set_param first)")
        if self.plan_idx_mat_.numel() != H + W or self.index_mat_.numel() != 2 * H * W:
            raise RuntimeError("lic360: index plan does not match the tensor size %dx%d" % (H, W))

    def _next_psum(self):
        p = self.plan_sum_
        self.plan_sum_ += 1
        return p


# ------------------------------------------------------------------------------------------------ context conv
class _CconvBase(_BaseOp):
    def __init__(self, channel, ngroup, nout, kernel_size, constrain, device=0, timeit=False):
        self.channel_, self.ngroup_, self.nout_ = int(channel), int(ngroup), int(nout)
        self.kernel_size_, self.constrain_ = int(kernel_size), int(constrain)
        self._pack_key = None
        self._wp = self._wq = None
        _BaseOp.__init__(self, device, timeit)

    def _pack(self, weight, nsets):
        weight = _f32(weight, "weight")
        key = (weight.data_ptr(), weight._version, tuple(weight.shape), weight.device.index)
        if key != self._pack_key:
            if weight.numel() != nsets * self.nout_ * self.channel_ * self.kernel_size_ ** 2:
                raise RuntimeError("lic360: weight shape %s does not match the op" % (tuple(weight.shape),))
            npf = LIB.lic360_cconv_wp_floats(nsets, self.channel_, self.nout_, self.ngroup_)
            nqf = LIB.lic360_cconv_wq_floats(nsets, self.channel_, self.nout_, self.ngroup_)
            self._wp = torch.empty(npf, dtype=torch.float32, device=weight.device)
            self._wq = torch.empty(max(nqf, 1), dtype=torch.float32, device=weight.device)
            check(LIB.lic360_cconv_pack(ptr(weight), ptr(self._wp), ptr(self._wq), nsets, self.channel_, self.nout_,
                                        self.ngroup_, self.kernel_size_, self.constrain_, cstream(weight)))
            self._pack_key = key
        return self._wp, self._wq

    @staticmethod
    def _nsets(weight, batch):
        return int(weight.size(0)) if batch else 1


class CconvEcOp(_CconvBase):
    """cconv_ec_opt (cconv_ec.hpp:5-33; main.cpp:96-102)."""

    def _run(self, x, weight, bias, act, batch):
        x = _f32(x, "input")
        n, c, h, w = x.shape
        if c != self.channel_:
            raise RuntimeError("lic360: CconvEc expects %d input channels, got %d" % (self.channel_, c))
        nsets = self._nsets(weight, batch)
        wp, wq = self._pack(weight, nsets)
        top = self._tops(x, [(n, self.nout_, h, w)])
        check(LIB.lic360_cconv_ec_forward(ptr(x), ptr(wp), ptr(wq), ptr(_f32(bias, "bias")),
                                          ptr_or_null(None if act is None else _f32(act, "act")), None, ptr(top[0]),
                                          n, c, h, w, self.nout_, self.ngroup_, self.constrain_, nsets, cstream(x)))
        return top

    def forward(self, x, weight, bias):
        return self._run(x, weight, bias, None, False)

    def forward_act(self, x, weight, bias, act):
        return self._run(x, weight, bias, act, False)

    def forward_batch(self, x, weight, bias):
        return self._run(x, weight, bias, None, True)

    def forward_act_batch(self, x, weight, bias, act):
        return self._run(x, weight, bias, act, True)


class CconvDcOp(_CconvBase, _PlanMixin):
    """cconv_dc_opt (cconv_dc.hpp:5-41; main.cpp:86-94): one wavefront step per call into a persistent frame."""

    def _on_init(self):
        self._plan_init()

    def _run(self, x, weight, bias, act, batch):
        x = _f32(x, "input")
        n, c, h, w = x.shape
        if c != self.channel_:
            raise RuntimeError("lic360: CconvDc expects %d input channels, got %d" % (self.channel_, c))
        if self._reshape((n, c, h, w)):
            self.plan_sum_ = 0  # cconv_dc_cuda.cu:12-17
        self._check_plan(h, w)
        nsets = self._nsets(weight, batch)
        wp, wq = self._pack(weight, nsets)
        top = self._tops(x, [(n, self.nout_, h, w)])
        psum = self._next_psum()
        check(LIB.lic360_cconv_dc_forward(ptr(x), ptr(wp), ptr(wq), ptr(_f32(bias, "bias")),
                                          ptr_or_null(None if act is None else _f32(act, "act")), None, ptr(top[0]),
                                          n, c, h, w, self.nout_, self.ngroup_, self.constrain_, nsets,
                                          ptr(self.index_mat_), ptr(self.plan_idx_mat_), psum, cstream(x)))
        return top

    def forward(self, x, weight, bias):
        return self._run(x, weight, bias, None, False)

    def forward_act(self, x, weight, bias, act):
        return self._run(x, weight, bias, act, False)

    def forward_batch(self, x, weight, bias):
        return self._run(x, weight, bias, None, True)

    def forward_act_batch(self, x, weight, bias, act):
        return self._run(x, weight, bias, act, True)


# ------------------------------------------------------------------------------------------------ wavefront plumbing
class CodeContexOp(_BaseOp):
    """code_contex_opt (code_contex.hpp, code_contex_cuda.cu:11-38): index plan, built on the host once per shape."""

    def forward(self, x):
        h, w = int(x.shape[2]), int(x.shape[3])
        if self._reshape((h, w)) or self._top is None:
            idx = torch.zeros((h, w, 2), dtype=torch.int32)
            plan = torch.zeros((h + w,), dtype=torch.int32)
            check(LIB.lic360_code_contex(h, w, ptr(idx), ptr(plan)))
            self._top = [idx.to(torch.device("cuda", self.device_)), plan]
        return self._top

    def backward(self, top_diff):
        return []


class TileExtractOp(_BaseOp, _PlanMixin):
    """tile_extract_opt (tile_extract.hpp; main.cpp:104-110)."""

    def __init__(self, ngroup, label, device=0, timeit=False):
        self.ngroup_, self.label_ = int(ngroup), bool(label)
        _BaseOp.__init__(self, device, timeit)

    def _on_init(self):
        self._plan_init()
        self._count = ctypes.c_int(0)

    def _run(self, x, batch):
        x = _f32(x, "input")
        n, c, h, w = x.shape
        if self._reshape((n, c, h, w)):
            self.plan_sum_ = 0
            self.top_num_ = torch.zeros((1,), dtype=torch.int32)
        self._check_plan(h, w)
        cpn = c // self.ngroup_
        top = self._tops(x, [(n, cpn, h, w)])
        psum = self._next_psum()
        if batch:
            check(LIB.lic360_tile_extract_batch(ptr(x), ptr(top[0]), n, c, h, w, self.ngroup_, ptr(self.index_mat_),
                                                ptr(self.plan_idx_mat_), psum, ctypes.byref(self._count), cstream(x)))
        else:
            check(LIB.lic360_tile_extract(ptr(x), ptr(top[0]), n, c, h, w, self.ngroup_, int(self.label_),
                                          ptr(self.index_mat_), ptr(self.plan_idx_mat_), psum,
                                          ctypes.byref(self._count), cstream(x)))
        self.top_num_[0] = self._count.value
        return [top[0], self.top_num_]

    def forward(self, x):
        return self._run(x, False)

    def forward_batch(self, x):
        return self._run(x, True)


class TileInputOp(_BaseOp, _PlanMixin):
    """tile_input_opt (tile_input.hpp; main.cpp:112-117)."""

    def __init__(self, ngroup, bias, scale, replicate, device=0, timeit=False):
        self.ngroup_, self.bias_, self.scale_, self.rep_ = int(ngroup), float(bias), float(scale), int(replicate)
        _BaseOp.__init__(self, device, timeit)

    def _on_init(self):
        self._plan_init()

    def forward(self, x):
        x = _f32(x, "input")
        n, h, w = int(x.shape[0]), int(x.shape[2]), int(x.shape[3])
        if self._reshape((n, self.ngroup_, h, w)):
            self.plan_sum_ = 0
        self._check_plan(h, w)
        top = self._tops(x, [(self.rep_ * n, self.ngroup_, h, w)])
        psum = self._next_psum()
        check(LIB.lic360_tile_input(ptr(x), ptr(top[0]), n, self.ngroup_, h, w, self.bias_, self.scale_, self.rep_,
                                    ptr(self.index_mat_), ptr(self.plan_idx_mat_), psum, cstream(x)))
        return top


class TileAddOp(_BaseOp, _PlanMixin):
    """tile_add_opt (tile_add.hpp; main.cpp:119-124): y[slab] += x[slab], in place on the first argument."""

    def __init__(self, ngroup, device=0, timeit=False):
        self.ngroup_ = int(ngroup)
        _BaseOp.__init__(self, device, timeit)

    def _on_init(self):
        self._plan_init()

    def forward(self, y, x):
        if not y.is_contiguous():
            raise RuntimeError("lic360: TileAdd works in place and needs a contiguous first argument")
        _f32(y, "input")
        x = _f32(x, "input2")
        n, c, h, w = y.shape
        if self._reshape((n, c, h, w)):
            self.plan_sum_ = 0
        self._check_plan(h, w)
        psum = self._next_psum()
        check(LIB.lic360_tile_add(ptr(y), ptr(x), n, c, h, w, self.ngroup_, ptr(self.index_mat_),
                                  ptr(self.plan_idx_mat_), psum, cstream(y)))
        return [y]


# ------------------------------------------------------------------------------------------------ CDF tables
class EntropyGmmTableOp(_BaseOp):
    """entropy_gmm_table_opt (entropy_gmm_table.hpp; main.cpp:126-130)."""

    def __init__(self, nstep, bias, num_gaussian, total_region, beta=1e-6, device=0, timeit=False):
        self.nstep_, self.bias_, self.ng_ = int(nstep), float(bias), int(num_gaussian)
        self.total_, self.beta_ = int(total_region), float(beta)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, weight, delta, mean, tnum):
        for t, nm in ((weight, "weight"), (delta, "delta"), (mean, "mean")):
            _f32(t, nm)
            if not t.is_contiguous():
                raise RuntimeError("lic360: EntropyGmmTable rewrites %s in place and needs it contiguous" % nm)
        n, _, h, w = weight.shape
        top = self._tops(weight, [(n * h * w, self.nstep_ + 1)])
        tn = int(tnum[0])
        check(LIB.lic360_gmm_table(ptr(weight), ptr(delta), ptr(mean), ptr(top[0]), tn, self.ng_, self.nstep_,
                                   self.bias_, self.total_, self.beta_, cstream(weight)))
        return top

    def forward_batch(self, data, tnum):
        _f32(data, "input")
        if not data.is_contiguous():
            raise RuntimeError("lic360: EntropyGmmTable rewrites its input in place and needs it contiguous")
        n, c, h, w = data.shape
        top = self._tops(data, [(n * h * w // 3, self.nstep_ + 1)])
        tn = int(tnum[0])
        stride = n * c * h * w // 3  # entropy_gmm_table_cuda.cu:167
        if tn > 0:
            flat = data.view(-1)
            check(LIB.lic360_gmm_table(ptr(flat), ptr(flat[stride:]), ptr(flat[2 * stride:]), ptr(top[0]), tn, self.ng_,
                                       self.nstep_, self.bias_, self.total_, self.beta_, cstream(data)))
        return top


class EntropyTableOp(_BaseOp):
    """entropy_table_opt (entropy_table.hpp; main.cpp:150-153)."""

    def __init__(self, nstep, total_region, device=0, timeit=False):
        self.nstep_, self.total_ = int(nstep), int(total_region)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, x, count):
        x = _f32(x, "input")
        n, _, h, w = x.shape
        top = self._tops(x, [(n * h * w, self.nstep_ + 1)])
        check(LIB.lic360_entropy_table(ptr(x), ptr(top[0]), int(count[0]), self.nstep_, self.total_, cstream(x)))
        return top


class EntropyGmmOp(_BaseOp):
    """entropy_gmm_opt (entropy_gmm.hpp; main.cpp:67-71): NLL + gradients cached in the forward."""

    def __init__(self, num_gaussian=3, ignore_label=-1, device=0, timeit=False):
        self.ng_, self.ignore_ = int(num_gaussian), int(ignore_label)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, weight, delta, mean, label):
        weight, delta, mean, label = _f32(weight), _f32(delta), _f32(mean), _f32(label)
        s, ng = int(weight.shape[0]), int(weight.shape[1])
        if ng != self.ng_:
            raise RuntimeError("lic360: the last dim of weight must equal the number of gaussians")
        top = self._tops(weight, [(s,)])
        bd = self._bottoms(weight, [(s, ng), (s, ng), (s, ng), (s, 1)])
        self._s = s
        check(LIB.lic360_entropy_gmm_forward(ptr(weight), ptr(delta), ptr(mean), ptr(label), ptr(bd[0]), ptr(bd[1]),
                                             ptr(bd[2]), ptr(bd[3]), ptr(top[0]), s, ng, cstream(weight)))
        return top

    def backward(self, top_diff):
        top_diff = _f32(top_diff)
        bd = self._bottom
        check(LIB.lic360_entropy_gmm_backward(ptr(bd[0]), ptr(bd[1]), ptr(bd[2]), ptr(bd[3]), ptr(top_diff), self._s,
                                              self.ng_, cstream(top_diff)))
        return bd


# ------------------------------------------------------------------------------------------------ layout ops
class ContextReshapeOp(_BaseOp):
    """context_reshape_opt (main.cpp:61-65)."""

    def __init__(self, ngroup, device=0, timeit=False):
        self.ngroup_ = int(ngroup)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, x):
        x = _f32(x)
        n, c, h, w = x.shape
        self._reshape((n, c, h, w))
        top = self._tops(x, [(n * h * w * self.ngroup_, c // self.ngroup_)])
        check(LIB.lic360_context_reshape(ptr(x), ptr(top[0]), n, c, h, w, self.ngroup_, 0, cstream(x)))
        return top

    def backward(self, top_diff):
        top_diff = _f32(top_diff)
        n, c, h, w = self._shape
        bd = self._bottoms(top_diff, [(n, c, h, w)])
        check(LIB.lic360_context_reshape(ptr(top_diff), ptr(bd[0]), n, c, h, w, self.ngroup_, 1, cstream(top_diff)))
        return bd


class ContexShiftOp(_BaseOp):
    """contex_shift_opt (main.cpp:55-59). The skewed tensor is zero-filled where the reference leaves it
    uninitialised (contex_shift_cuda.cu:66-90 writes into an at::empty buffer)."""

    def __init__(self, inv, cpn=1, device=0, timeit=False):
        self.inv_, self.cpn_ = bool(inv), int(cpn)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, x):
        x = _f32(x)
        n, c, h, w = x.shape
        self._reshape((n, c, h, w))
        g = c // self.cpn_
        if self.inv_:
            ho = h - w - g + 2
            top = self._tops(x, [(n, c, ho, w)])
            check(LIB.lic360_contex_shift(ptr(x), ptr(top[0]), n, c, ho, w, self.cpn_, 1, 0, cstream(x)))
        else:
            ho = h + w + g - 2
            top = self._tops(x, [(n, c, ho, w)])
            check(LIB.lic360_contex_shift(ptr(x), ptr(top[0]), n, c, h, w, self.cpn_, 0, 1, cstream(x)))
        return top

    def backward(self, top_diff):
        top_diff = _f32(top_diff)
        n, c, h, w = self._shape
        g = c // self.cpn_
        bd = self._bottoms(top_diff, [(n, c, h, w)])
        if self.inv_:  # zero-fill then scatter (contex_shift_cuda.cu:128-131)
            check(LIB.lic360_contex_shift(ptr(top_diff), ptr(bd[0]), n, c, h - w - g + 2, w, self.cpn_, 0, 1,
                                          cstream(top_diff)))
        else:
            check(LIB.lic360_contex_shift(ptr(top_diff), ptr(bd[0]), n, c, h, w, self.cpn_, 1, 0, cstream(top_diff)))
        return bd


class MaskConstrainOp(_BaseOp):
    """mask_constrain_opt (main.cpp:73-78): zero the masked weights / weight grads in place; returns None."""

    def __init__(self, constrain=5, ngroup=1, device=0, timeit=False):
        self.constrain_, self.ngroup_ = int(constrain), int(ngroup)
        _BaseOp.__init__(self, device, timeit)

    def _apply(self, t):
        _f32(t)
        if not t.is_contiguous():
            raise RuntimeError("lic360: MaskConstrain works in place and needs a contiguous tensor")
        co, ci, k, k2 = t.shape
        check(LIB.lic360_mask_constrain(ptr(t), co, ci, k, self.ngroup_, self.constrain_, cstream(t)))

    def forward(self, x):
        self._apply(x)

    def backward(self, top_diff):
        self._apply(top_diff)


# ------------------------------------------------------------------------------------------------ quantisation / importance
class QuantOp(_BaseOp):
    """quant_opt (quant.hpp; main.cpp:42-46)."""

    def __init__(self, channel, bin_num, weight_decay=0.9, check_iters=100, ntop=1, top_alpha=0.1, device=0,
                 timeit=False):
        self.channel_, self.bin_num_ = int(channel), int(bin_num)
        self.weight_decay_, self.mod_, self.ntop_, self.top_alpha_ = float(weight_decay), int(check_iters), int(ntop), float(top_alpha)
        _BaseOp.__init__(self, device, timeit)

    def _on_init(self):
        dev = torch.device("cuda", self.device_)
        self.weight_ = torch.zeros((self.channel_, self.bin_num_), dtype=torch.float32, device=dev)
        self.count_data_ = torch.zeros((self.channel_, self.bin_num_), dtype=torch.float32, device=dev)
        self.quant_ = None
        self.iter_ = 0

    def forward(self, x, weight, ncount, train):
        x = _f32(x)
        n, c, h, w = x.shape
        if self._reshape((n, c, h, w)) or self.quant_ is None:
            self.quant_ = torch.zeros((n, c, h, w), dtype=torch.int32, device=x.device)
        top = self._tops(x, [(n, c, h, w)] * (2 if self.ntop_ > 1 else 1))
        if train and self.iter_ % self.mod_ == 0 and self.iter_ != 0:  # quant_cuda.cu:119-120
            check(LIB.lic360_quant_update_weight(ptr(weight), ptr(ncount), self.channel_, self.bin_num_,
                                                 self.weight_decay_, cstream(x)))
        check(LIB.lic360_quant_forward(ptr(x), ptr(_f32(weight)), ptr(self.weight_), ptr(top[0]),
                                       ptr_or_null(top[1] if self.ntop_ > 1 else None), ptr(self.quant_),
                                       ptr(self.count_data_), n, c, h, w, self.bin_num_, cstream(x)))
        if train:
            self.iter_ += 1
        return top

    def backward(self, top_diff, bottom_data, top_data):
        td0 = _f32(top_diff[0])
        td1 = _f32(top_diff[1]) if self.ntop_ > 1 else None
        n, c, h, w = self._shape
        bd = self._bottoms(td0, [(n, c, h, w), (self.channel_, self.bin_num_)])
        check(LIB.lic360_quant_backward(ptr(td0), ptr_or_null(td1), ptr(_f32(bottom_data)), ptr(_f32(top_data)),
                                        ptr(self.quant_), ptr(self.weight_), ptr(bd[0]), ptr(bd[1]), n, c, h, w,
                                        self.bin_num_, self.top_alpha_, cstream(td0)))
        return [bd[0], bd[1], self.count_data_]


class DquantOp(_BaseOp):
    """dquant_opt (dquant.hpp; main.cpp:145-148)."""

    def __init__(self, channel, bin_num, device=0, timeit=False):
        self.nchannel_, self.bin_num_ = int(channel), int(bin_num)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, x, mask, weight_old):
        x, mask = _f32(x), _f32(mask)
        n, c, h, w = x.shape
        if self._reshape((n, c, h, w)) or getattr(self, "weight_", None) is None:
            self.weight_ = torch.zeros((self.nchannel_, self.bin_num_), dtype=torch.float32, device=x.device)
        top = self._tops(x, [(n, c, h, w)])
        check(LIB.lic360_dquant_forward(ptr(x), ptr(mask), ptr(_f32(weight_old)), ptr(self.weight_), ptr(top[0]), n, c,
                                        h, w, self.bin_num_, cstream(x)))
        return top


class ImpMapOp(_BaseOp):
    """imp_map_opt (imp_map.hpp; main.cpp:30-34)."""

    def __init__(self, levels, alpha, gamma, rt, scale_constrain, scale_weight, imp_kernel=0, ntop=1, device=0,
                 timeit=False):
        self.levels_, self.alpha_, self.gamma_, self.rt_ = int(levels), float(alpha), float(gamma), float(rt)
        self.scale_constrain_, self.scale_weight_ = float(scale_constrain), float(scale_weight)
        self.imp_kernel_, self.ntop_ = int(imp_kernel), int(ntop)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, x, imp):
        x, imp = _f32(x), _f32(imp)
        n, c, h, w = x.shape
        fresh = self._reshape((n, c, h, w)) or self._top is None
        shapes = [(n, c, h, w), (n, 1, h)] + ([(n, c, h, w)] if self.ntop_ > 1 else [])
        top = self._tops(x, shapes)
        if fresh:  # reshape_init_alpha_constrain (imp_map_cuda.cu:41-69)
            self.alpha_t_ = torch.empty((h,), dtype=torch.float32, device=x.device)
            check(LIB.lic360_imp_map_init(ptr(top[1]), ptr(self.alpha_t_), n, h, self.alpha_, self.rt_,
                                          self.scale_constrain_, self.scale_weight_, cstream(x)))
        check(LIB.lic360_imp_map_forward(ptr(x), ptr(imp), ptr(top[0]), ptr_or_null(top[2] if self.ntop_ > 1 else None),
                                         n, c, h, w, self.levels_, cstream(x)))
        return top

    def backward(self, top_diff, bottom_imp, sphere_constrain):
        top_diff, bottom_imp, sphere_constrain = _f32(top_diff), _f32(bottom_imp), _f32(sphere_constrain)
        n, c, h, w = self._shape
        bd = self._bottoms(top_diff, [(n, c, h, w), (n, 1, h, w)])
        check(LIB.lic360_imp_map_backward(ptr(top_diff), ptr(bottom_imp), ptr(sphere_constrain), ptr(self.alpha_t_),
                                          ptr(bd[0]), ptr(bd[1]), n, c, h, w, self.levels_, self.gamma_,
                                          self.imp_kernel_, cstream(top_diff)))
        return bd


class Imp2maskOp(_BaseOp):
    """imp2mask_opt (imp2mask.hpp; main.cpp:160-163)."""

    def __init__(self, levels, channels, device=0, timeit=False):
        self.levels_, self.channel_ = int(levels), int(channels)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, x):
        x = _f32(x)
        n, _, h, w = x.shape
        top = self._tops(x, [(n, self.channel_, h, w)])
        check(LIB.lic360_imp2mask(ptr(x), ptr(top[0]), n, self.channel_, h, w, self.levels_, cstream(x)))
        return top


class ScaleOp(_BaseOp):
    """scale_opt (scale.hpp; main.cpp:155-158). No backward is bound in the reference either."""

    def __init__(self, bias, scale, device=0, timeit=False):
        self.bias_, self.scale_ = float(bias), float(scale)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, x):
        x = _f32(x)
        top = self._tops(x, [tuple(x.shape)])
        check(LIB.lic360_scale(ptr(x), ptr(top[0]), x.numel(), self.bias_, self.scale_, cstream(x)))
        return top


# ------------------------------------------------------------------------------------------------ sphere geometry
class SpherePadOp(_BaseOp):
    """sphere_pad_opt (sphere_pad.hpp; main.cpp:12-16)."""

    def __init__(self, pad, inplace=False, device=0, timeit=False):
        self.pad_, self.inplace_ = int(pad), bool(inplace)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, x):
        _f32(x)
        n, c, h, w = x.shape
        self._reshape((n, c, h, w))
        if self.inplace_:
            if not x.is_contiguous():
                raise RuntimeError("lic360: in-place SpherePad needs a contiguous tensor")
            check(LIB.lic360_sphere_pad_inplace(ptr(x), n * c, h, w, self.pad_, cstream(x)))
            return [x]
        x = x.contiguous()
        top = self._tops(x, [(n, c, h + 2 * self.pad_, w + 2 * self.pad_)])
        check(LIB.lic360_sphere_pad(ptr(x), ptr(top[0]), n * c, h, w, self.pad_, cstream(x)))
        return top

    def backward(self, top_diff):
        _f32(top_diff)
        n, c, h, w = self._shape
        if self.inplace_:
            if not top_diff.is_contiguous():
                raise RuntimeError("lic360: in-place SpherePad backward needs a contiguous tensor")
            check(LIB.lic360_sphere_pad_backward(None, ptr(top_diff), n * c, h - 2 * self.pad_, w - 2 * self.pad_,
                                                 self.pad_, 1, cstream(top_diff)))
            return [top_diff]
        top_diff = top_diff.contiguous()
        bd = self._bottoms(top_diff, [(n, c, h, w)])
        check(LIB.lic360_sphere_pad_backward(ptr(bd[0]), ptr(top_diff), n * c, h, w, self.pad_, 0, cstream(top_diff)))
        return bd


class SphereTrimOp(_BaseOp):
    """sphere_trim_opt (sphere_trim.hpp; main.cpp:18-22): zero the border in place, forward and backward."""

    def __init__(self, pad=1, device=0, timeit=False):
        self.pad_ = int(pad)
        _BaseOp.__init__(self, device, timeit)

    def _trim(self, t):
        _f32(t)
        if not t.is_contiguous():
            raise RuntimeError("lic360: SphereTrim works in place and needs a contiguous tensor")
        n, c, h, w = t.shape
        check(LIB.lic360_sphere_trim(ptr(t), n * c, h, w, self.pad_, cstream(t)))
        return [t]

    def forward(self, x):
        self._reshape(tuple(x.shape))
        return self._trim(x)

    def backward(self, top_diff):
        return self._trim(top_diff)


class SphereCutEdgeOp(_BaseOp):
    """sphere_cut_edge_opt (sphere_cut_edge.hpp; main.cpp:24-28)."""

    def __init__(self, pad=1, device=0, timeit=False):
        self.pad_ = int(pad)
        _BaseOp.__init__(self, device, timeit)

    def forward(self, x):
        x = _f32(x)
        n, c, h, w = x.shape
        self._reshape((n, c, h, w))
        top = self._tops(x, [(n, c, h - 2 * self.pad_, w - 2 * self.pad_)])
        check(LIB.lic360_sphere_cut_edge(ptr(x), ptr(top[0]), n * c, h, w, self.pad_, 0, cstream(x)))
        return top

    def backward(self, top_diff):
        top_diff = _f32(top_diff)
        n, c, h, w = self._shape
        bd = self._bottoms(top_diff, [(n, c, h, w)])
        check(LIB.lic360_sphere_cut_edge(ptr(top_diff), ptr(bd[0]), n * c, h, w, self.pad_, 1, cstream(top_diff)))
        return bd


class SphereLatScaleOp(_BaseOp):
    """sphere_lat_scale_opt (sphere_lat_scale.hpp; main.cpp:48-53)."""

    def __init__(self, npart, device=0, timeit=False):
        self.npart_ = int(npart)
        _BaseOp.__init__(self, device, timeit)

    def set_npart(self, npart):
        self.npart_ = int(npart)
        self._shape = None

    def forward(self, x, weight):
        x, weight = _f32(x), _f32(weight)
        n, c, h, w = x.shape
        self._reshape((n, c, h, w))
        top = self._tops(x, [(n, c, h, w)])
        check(LIB.lic360_sphere_lat_scale(ptr(x), ptr(weight), ptr(top[0]), n * c, h, w, self.npart_, cstream(x)))
        return top

    def backward(self, top_diff, weight):
        top_diff, weight = _f32(top_diff), _f32(weight)
        n, c, h, w = self._shape
        bd = self._bottoms(top_diff, [(n, c, h, w)])
        check(LIB.lic360_sphere_lat_scale(ptr(top_diff), ptr(weight), ptr(bd[0]), n * c, h, w, self.npart_,
                                          cstream(top_diff)))
        return bd


class DtowOp(_BaseOp):
    """dtow_opt (dtow.hpp; main.cpp:36-40) -- SURVEY s8(f)-1."""

    def __init__(self, stride, d2w, device=0, timeit=False):
        self.stride_, self.d2w_ = int(stride), bool(d2w)
        _BaseOp.__init__(self, device, timeit)

    def _out_shape(self, n, c, h, w):
        s = self.stride_
        return (n, c // (s * s), h * s, w * s) if self.d2w_ else (n, c * s * s, h // s, w // s)

    def forward(self, x):
        x = _f32(x)
        n, c, h, w = x.shape
        self._reshape((n, c, h, w))
        top = self._tops(x, [self._out_shape(n, c, h, w)])
        check(LIB.lic360_dtow(ptr(x), ptr(top[0]), n, c, h, w, self.stride_, int(self.d2w_), cstream(x)))
        return top

    def backward(self, top_diff):
        top_diff = _f32(top_diff)
        n, c, h, w = self._shape
        on, oc, oh, ow = self._out_shape(n, c, h, w)
        bd = self._bottoms(top_diff, [(n, c, h, w)])
        check(LIB.lic360_dtow(ptr(top_diff), ptr(bd[0]), on, oc, oh, ow, self.stride_, int(not self.d2w_),
                              cstream(top_diff)))
        return bd


class ProjectsOp(_BaseOp):
    """projects_opt (projects.hpp:6-34; main.cpp:6-10) -- SURVEY s8(f)-2: the 14 viewports of MultiProject.
    theta / phi: 14 angles each in units of pi, fov in units of pi (projects.hpp:8-19)."""

    def __init__(self, h_out, w_out, theta, phi, fov=0.33333, near=False, device=0, timeit=False):
        self.h_out_, self.w_out_, self.fov_, self.near_ = int(h_out), int(w_out), float(fov), bool(near)
        if len(theta) < 14 or len(phi) < 14:
            raise RuntimeError("lic360: ProjectsOp needs 14 theta and 14 phi values")
        self._theta = (ctypes.c_float * 14)(*[float(v) for v in theta[:14]])
        self._phi = (ctypes.c_float * 14)(*[float(v) for v in phi[:14]])
        self._xyz = self._tf = None
        self._hw = None
        _BaseOp.__init__(self, device, timeit)

    def _on_init(self):  # projects_opt::init: the rays are rebuilt when the op moves to another device
        self._xyz = self._tf = None
        self._hw = None

    def _geometry(self, x):
        n, c, h, w = x.shape
        inner = self.h_out_ * self.w_out_
        if self._xyz is None or self._xyz.device != x.device:
            self._xyz = torch.empty((14, inner, 3), dtype=torch.float32, device=x.device)
            self._tf = torch.empty((14, inner, 2), dtype=torch.float32, device=x.device)
            check(LIB.lic360_projects_init(ptr(self._xyz), self.h_out_, self.w_out_, self._theta, self._phi, self.fov_, cstream(x)))
            self._hw = None
        if self._hw != (h, w):  # projects_opt::reshape -> update()
            check(LIB.lic360_projects_update(ptr(self._xyz), ptr(self._tf), self.h_out_, self.w_out_, h, w, cstream(x)))
            self._hw = (h, w)

    def forward(self, x):
        x = _f32(x)
        n, c, h, w = x.shape
        self._geometry(x)
        self._reshape((n, c, h, w))
        top = self._tops(x, [(n * 14, c, self.h_out_, self.w_out_)])
        check(LIB.lic360_projects_forward(ptr(x), ptr(self._tf), ptr(top[0]), n * c, h, w, self.h_out_, self.w_out_,
                                          int(self.near_), cstream(x)))
        return top

    def backward(self, top_diff):
        top_diff = _f32(top_diff)
        if self._shape is None or self._tf is None:
            raise RuntimeError("lic360: ProjectsOp.backward called before forward")
        n, c, h, w = self._shape
        bd = self._bottoms(top_diff, [(n, c, h, w), (n, c, h, w)])
        check(LIB.lic360_projects_backward(ptr(top_diff), ptr(self._tf), ptr(bd[0]), ptr(bd[1]), n * c, h, w, self.h_out_,
                                           self.w_out_, int(self.near_), cstream(top_diff)))
        return bd


# ------------------------------------------------------------------------------------------------ host coder
class Coder(object):
    """Coder (coder.h:10-63; main.cpp:132-143): host arithmetic coder, same bitstream format."""

    def __init__(self, fname, file_value):
        self._h = ctypes.c_void_p(LIB.lic360_coder_create(os.fsencode(fname), float(file_value)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            LIB.lic360_coder_destroy(h)
            self._h = None

    def reset_fname(self, fname):
        check(LIB.lic360_coder_reset_fname(self._h, os.fsencode(fname)))

    def start_encoder(self):
        check(LIB.lic360_coder_start_encoder(self._h))

    def end_encoder(self):
        check(LIB.lic360_coder_end_encoder(self._h))

    def start_decoder(self):
        check(LIB.lic360_coder_start_decoder(self._h))

    @staticmethod
    def _i32(t):
        t = t.to("cpu").to(torch.int32)
        return t if t.is_contiguous() else t.contiguous()

    def encode(self, table, ncode, total, symbol):
        table = self._i32(table)
        check(LIB.lic360_coder_encode_one(self._h, ptr(table), int(ncode), int(total), int(symbol)))

    def decode(self, table, ncode, total):
        table = self._i32(table)
        out = ctypes.c_int(0)
        check(LIB.lic360_coder_decode_one(self._h, ptr(table), int(ncode), int(total), ctypes.byref(out)))
        return out.value

    def encodes(self, table, ncode, labels, num):
        table, labels = self._i32(table), self._i32(labels)
        check(LIB.lic360_coder_encodes(self._h, ptr(table), int(ncode), ptr(labels), None, int(num)))

    def encodes_mask(self, table, ncode, labels, mask, num):
        table, labels = self._i32(table), self._i32(labels)
        mask = mask.to("cpu").to(torch.float32).contiguous()
        check(LIB.lic360_coder_encodes(self._h, ptr(table), int(ncode), ptr(labels), ptr(mask), int(num)))

    def decodes(self, table, ncode, num):
        table = self._i32(table)
        out = torch.empty((table.size(0),), dtype=torch.float32)
        check(LIB.lic360_coder_decodes(self._h, ptr(table), int(ncode), None, int(num), ptr(out)))
        return out

    def decodes_mask(self, table, ncode, mask, num):
        table = self._i32(table)
        mask = mask.to("cpu").to(torch.float32).contiguous()
        out = torch.empty((table.size(0),), dtype=torch.float32)
        check(LIB.lic360_coder_decodes(self._h, ptr(table), int(ncode), ptr(mask), int(num), ptr(out)))
        return out
