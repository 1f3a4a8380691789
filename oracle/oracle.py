"""TEST INFRASTRUCTURE ONLY -- numpy front-end of oracle/liboracle.so (the CPU restatement, lic360_oracle.c) and of
oracle/_ref/libref_coder.so (the reference's own host coder classes).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module; the
product (360-image-compression_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "lic360_oracle.c")
_REF_CODER = os.path.join(_HERE, "_ref", "libref_coder.so")

if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])

_L = ctypes.CDLL(_SO)
_P, _I, _F, _Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


def _sig(name, args, res=None):
    fn = getattr(_L, name)
    fn.argtypes = args
    fn.restype = res
    return fn


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data


# ------------------------------------------------------------------------------------------------ plan
_code_contex = _sig("orc_code_contex", [_I, _I, _P, _P])
_slab = _sig("orc_slab", [_P, _I, _I, _I, _I, _P, _P])


def code_contex(H, W):
    idx = np.zeros(2 * H * W, np.int32)
    plan = np.zeros(H + W, np.int32)
    _code_contex(H, W, _p(idx), _p(plan))
    return idx, plan


def slab(plan, H, W, G, psum):
    s, l = ctypes.c_int(0), ctypes.c_int(0)
    _slab(_p(plan), H, W, G, psum, ctypes.byref(s), ctypes.byref(l))
    return s.value, l.value


# ------------------------------------------------------------------------------------------------ conv
_ec = _sig("orc_cconv_ec", [_P] * 5 + [_I] * 9)
_dc = _sig("orc_cconv_dc_step", [_P] * 5 + [_I] * 9 + [_P, _P, _I])


def cconv_ec(x, w, b, slope, G, constrain, nsets=1):
    x, w, b = _f(x), _f(w), _f(b)
    slope = None if slope is None else _f(slope)
    N, Cin, H, W = x.shape
    Cout, k = w.shape[-4], w.shape[-1]
    out = np.zeros((N, Cout, H, W), np.float32)
    _ec(_p(x), _p(w), _p(b), _p(slope), _p(out), N, Cin, H, W, Cout, G, k, constrain, nsets)
    return out


def cconv_dc_step(x, w, b, slope, out, G, constrain, nsets, idx, plan, psum):
    x, w, b = _f(x), _f(w), _f(b)
    slope = None if slope is None else _f(slope)
    N, Cin, H, W = x.shape
    Cout, k = w.shape[-4], w.shape[-1]
    assert out.dtype == np.float32 and out.flags.c_contiguous
    _dc(_p(x), _p(w), _p(b), _p(slope), _p(out), N, Cin, H, W, Cout, G, k, constrain, nsets, _p(idx), _p(plan), psum)
    return out


# ---- CPU-baseline form (channel-last, contiguous masked dot products; bench.py's cpu arm only)
_pwc = _sig("orc_pack_weights_cl", [_P, _P, _I, _I, _I, _I])
_ecc = _sig("orc_cconv_ec_cl", [_P] * 6 + [_I] * 9)
_dcc = _sig("orc_cconv_dc_step_cl", [_P] * 6 + [_I] * 9 + [_P, _P, _I])
_tec = _sig("orc_tile_extract_cl", [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I], _I)


def pack_weights_cl(w, nsets):
    """(nsets, Cout, Cin, k, k) or (Cout, Cin, k, k) -> (nsets, Cout, k*k, Cin)"""
    w = _f(w)
    Cout, Cin, k = w.shape[-4], w.shape[-3], w.shape[-1]
    wt = np.zeros((nsets, Cout, k * k, Cin), np.float32)
    _pwc(_p(w), _p(wt), nsets, Cout, Cin, k)
    return wt


def cconv_ec_cl(x, wt, b, slope, resid, G, constrain):
    """x (N,H,W,Cin) channel-last -> (N,H,W,Cout)"""
    x, b = _f(x), _f(b)
    slope = None if slope is None else _f(slope)
    resid = None if resid is None else _f(resid)
    N, H, W, Cin = x.shape
    nsets, Cout, kk, _ = wt.shape
    out = np.zeros((N, H, W, Cout), np.float32)
    _ecc(_p(x), _p(wt), _p(b), _p(slope), _p(resid), _p(out), N, Cin, H, W, Cout, G, int(round(kk ** 0.5)), constrain, nsets)
    return out


def cconv_dc_step_cl(x, wt, b, slope, resid, out, G, constrain, idx, plan, psum):
    b = _f(b)
    slope = None if slope is None else _f(slope)
    N, H, W, Cin = x.shape
    nsets, Cout, kk, _ = wt.shape
    assert x.dtype == np.float32 and x.flags.c_contiguous and out.dtype == np.float32 and out.flags.c_contiguous
    _dcc(_p(x), _p(wt), _p(b), _p(slope), _p(resid), _p(out), N, Cin, H, W, Cout, G, int(round(kk ** 0.5)), constrain, nsets,
         _p(idx), _p(plan), psum)
    return out


def tile_extract_cl(x, out, G, idx, plan, psum):
    N, H, W, C = x.shape
    return _tec(_p(x), _p(out), N, C, H, W, G, _p(idx), _p(plan), psum)


# ------------------------------------------------------------------------------------------------ tile ops
_te = _sig("orc_tile_extract", [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I], _I)
_teb = _sig("orc_tile_extract_batch", [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I], _I)
_ti = _sig("orc_tile_input", [_P, _P, _I, _I, _I, _I, _F, _F, _I, _P, _P, _I])
_ta = _sig("orc_tile_add", [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I])


def tile_extract(x, out, G, label, idx, plan, psum):
    x = _f(x)
    N, C, H, W = x.shape
    return _te(_p(x), _p(out), N, C, H, W, G, int(label), _p(idx), _p(plan), psum)


def tile_extract_batch(x, out, G, idx, plan, psum):
    x = _f(x)
    N, C, H, W = x.shape
    return _teb(_p(x), _p(out), N, C, H, W, G, _p(idx), _p(plan), psum)


def tile_input(sym, frame, N, G, H, W, bias, scale, rep, idx, plan, psum):
    sym = _f(sym)
    _ti(_p(sym), _p(frame), N, G, H, W, bias, scale, rep, _p(idx), _p(plan), psum)
    return frame


def tile_add(y, x, G, idx, plan, psum):
    x = _f(x)
    N, C, H, W = y.shape
    _ta(_p(y), _p(x), N, C, H, W, G, _p(idx), _p(plan), psum)
    return y


# ------------------------------------------------------------------------------------------------ tables
_gt = _sig("orc_gmm_table", [_P, _P, _P, _P, _I, _I, _I, _F, _I, _F])
_et = _sig("orc_entropy_table", [_P, _P, _I, _I, _I])
_egf = _sig("orc_entropy_gmm_fwd", [_P] * 9 + [_I, _I])
_egb = _sig("orc_entropy_gmm_bwd", [_P] * 5 + [_I, _I])


def gmm_table(weight, delta, mean, nstep=8, bias=3.5, total=65536, beta=1e-6):
    """Returns (table, softmaxed weight, clamped delta); inputs are not modified."""
    w, d, m = _f(weight).copy(), _f(delta).copy(), _f(mean)
    rows, ng = w.shape
    out = np.zeros((rows, nstep + 1), np.float32)
    _gt(_p(w), _p(d), _p(m), _p(out), rows, ng, nstep, bias, total, beta)
    return out, w, d


def entropy_table(logits, total=65536):
    x = _f(logits)
    rows, w = x.shape
    out = np.zeros((rows, w + 1), np.float32)
    _et(_p(x), _p(out), rows, w, total)
    return out


def entropy_gmm_fwd(weight, delta, mean, label):
    w, d, m, l = _f(weight), _f(delta), _f(mean), _f(label)
    S, ng = w.shape
    wd, dd, md = np.zeros_like(w), np.zeros_like(w), np.zeros_like(w)
    ld, loss = np.zeros((S, 1), np.float32), np.zeros(S, np.float32)
    _egf(_p(w), _p(d), _p(m), _p(l), _p(wd), _p(dd), _p(md), _p(ld), _p(loss), S, ng)
    return loss, wd, dd, md, ld


def entropy_gmm_bwd(wd, dd, md, ld, top):
    wd, dd, md, ld, top = wd.copy(), dd.copy(), md.copy(), ld.copy(), _f(top)
    S, ng = wd.shape
    _egb(_p(wd), _p(dd), _p(md), _p(ld), _p(top), S, ng)
    return wd, dd, md, ld


# ------------------------------------------------------------------------------------------------ layout / mask
_cr = _sig("orc_context_reshape", [_P, _P, _I, _I, _I, _I, _I, _I])
_cs = _sig("orc_contex_shift", [_P, _P, _I, _I, _I, _I, _I, _I])
_mc = _sig("orc_mask_constrain", [_P, _I, _I, _I, _I, _I])


def context_reshape(x, G):
    x = _f(x)
    N, C, H, W = x.shape
    out = np.zeros((N * G * H * W, C // G), np.float32)
    _cr(_p(x), _p(out), N, C, H, W, G, 0)
    return out


def context_reshape_bwd(rows, N, C, H, W, G):
    rows = _f(rows)
    out = np.zeros((N, C, H, W), np.float32)
    _cr(_p(rows), _p(out), N, C, H, W, G, 1)
    return out


def contex_shift(x, cpn):
    x = _f(x)
    N, C, H, W = x.shape
    out = np.zeros((N, C, H + W + C // cpn - 2, W), np.float32)
    _cs(_p(x), _p(out), N, C, H, W, cpn, 0)
    return out


def contex_shift_inv(xs, cpn):
    xs = _f(xs)
    N, C, Hs, W = xs.shape
    H = Hs - W - C // cpn + 2
    out = np.zeros((N, C, H, W), np.float32)
    _cs(_p(xs), _p(out), N, C, H, W, cpn, 1)
    return out


def mask_constrain(w, G, constrain):
    w = _f(w).copy()
    Cout, Cin, k, _ = w.shape
    _mc(_p(w), Cout, Cin, k, G, constrain)
    return w


# ------------------------------------------------------------------------------------------------ quant / imp
_ql = _sig("orc_quant_levels", [_P, _P, _I, _I])
_qf = _sig("orc_quant_fwd", [_P] * 5 + [_I] * 4)
_qb = _sig("orc_quant_bwd", [_P] * 8 + [_I] * 4 + [_F, _I])
_dl = _sig("orc_dquant_levels", [_P, _P, _I, _I])
_df = _sig("orc_dquant_fwd", [_P] * 4 + [_I] * 4)
_imf = _sig("orc_imp_map_fwd", [_P] * 4 + [_I] * 5)
_imc = _sig("orc_imp_map_constrain", [_P, _P, _I, _I, _F, _F, _F, _F])
_imbd = _sig("orc_imp_map_bwd_data", [_P] * 3 + [_I] * 5)
_imbi = _sig("orc_imp_map_bwd_imp", [_P] * 5 + [_I] * 5 + [_F, _I])
_i2m = _sig("orc_imp2mask", [_P, _P, _I, _I, _I, _I])
_sc = _sig("orc_scale", [_P, _P, _Z, _F, _F])


def quant_levels(wb):
    wb = _f(wb)
    C, L = wb.shape
    lv = np.zeros_like(wb)
    _ql(_p(wb), _p(lv), C, L)
    return lv


def quant_fwd(x, levels):
    x, levels = _f(x), _f(levels)
    N, C, H, W = x.shape
    L = levels.shape[1]
    y, q = np.zeros_like(x), np.zeros_like(x)
    count = np.zeros((C, L), np.float32)
    _qf(_p(x), _p(levels), _p(y), _p(q), _p(count), N, C, H * W, L)
    return y, q, count


def quant_bwd(td0, td1, x, y, q, levels, alpha):
    td0, x, y, q, levels = _f(td0), _f(x), _f(y), _f(q), _f(levels)
    td1 = None if td1 is None else _f(td1)
    N, C, H, W = x.shape
    L = levels.shape[1]
    bd, wd = np.zeros_like(x), np.zeros((C, L), np.float32)
    _qb(_p(td0), _p(td1), _p(x), _p(y), _p(q), _p(levels), _p(bd), _p(wd), N, C, H * W, L, alpha, 2 if td1 is not None else 1)
    return bd, wd


def dquant_levels(wb):
    wb = _f(wb)
    C, L = wb.shape
    cum = np.zeros_like(wb)
    _dl(_p(wb), _p(cum), C, L)
    return cum


def dquant_fwd(q, mask, cum):
    q, mask, cum = _f(q), _f(mask), _f(cum)
    N, C, H, W = q.shape
    y = np.zeros_like(q)
    _df(_p(q), _p(mask), _p(cum), _p(y), N, C, H * W, cum.shape[1])
    return y


def imp_map_fwd(x, imp, levels):
    x, imp = _f(x), _f(imp)
    N, C, H, W = x.shape
    out, mask = np.zeros_like(x), np.zeros_like(x)
    _imf(_p(x), _p(imp), _p(out), _p(mask), N, C, H, W, levels)
    return out, mask


def imp_map_constrain(N, H, alpha, rt, sc, sw):
    c, a = np.zeros((N, 1, H), np.float32), np.zeros(H, np.float32)
    _imc(_p(c), _p(a), N, H, alpha, rt, sc, sw)
    return c, a


def imp_map_bwd(top, imp, sphere, alpha_t, levels, gamma, imp_kernel):
    top, imp, sphere, alpha_t = _f(top), _f(imp), _f(sphere), _f(alpha_t)
    N, C, H, W = top.shape
    dd, di = np.zeros_like(top), np.zeros((N, 1, H, W), np.float32)
    _imbd(_p(top), _p(imp), _p(dd), N, C, H, W, levels)
    version = imp_kernel if imp_kernel in (1, 2, 3) else 0
    _imbi(_p(top), _p(imp), _p(sphere), _p(alpha_t), _p(di), N, C, H, W, levels, gamma, version)
    return dd, di


def imp2mask(lv, C, levels):
    lv = _f(lv)
    N, _, H, W = lv.shape
    out = np.zeros((N, C, H, W), np.float32)
    _i2m(_p(lv), _p(out), N, C, H * W, levels)
    return out


def scale(x, bias, scale_):
    x = _f(x)
    out = np.zeros_like(x)
    _sc(_p(x), _p(out), x.size, bias, scale_)
    return out


# ------------------------------------------------------------------------------------------------ sphere / dtow
_sp = _sig("orc_sphere_pad", [_P, _P, _I, _I, _I, _I])
_spi = _sig("orc_sphere_pad_inplace", [_P, _I, _I, _I, _I])
_spb = _sig("orc_sphere_pad_bwd", [_P, _P, _I, _I, _I, _I, _I])
_st = _sig("orc_sphere_trim", [_P, _I, _I, _I, _I])
_sce = _sig("orc_sphere_cut_edge", [_P, _P, _I, _I, _I, _I, _I])
_sls = _sig("orc_sphere_lat_scale", [_P, _P, _P, _I, _I, _I, _I])
_dt = _sig("orc_dtow", [_P, _P, _I, _I, _I, _I, _I, _I])


def sphere_pad(x, pad):
    x = _f(x)
    N, C, H, W = x.shape
    out = np.zeros((N, C, H + 2 * pad, W + 2 * pad), np.float32)
    _sp(_p(x), _p(out), N * C, H, W, pad)
    return out


def sphere_pad_inplace(x, pad):
    x = _f(x).copy()
    N, C, H, W = x.shape
    _spi(_p(x), N * C, H, W, pad)
    return x


def sphere_pad_bwd(top, pad, inplace):
    top = _f(top).copy()
    N, C, Ho, Wo = top.shape
    H, W = Ho - 2 * pad, Wo - 2 * pad
    if inplace:
        _spb(None, _p(top), N * C, H, W, pad, 1)
        return top
    bottom = np.zeros((N, C, H, W), np.float32)
    _spb(_p(bottom), _p(top), N * C, H, W, pad, 0)
    return bottom


def sphere_trim(x, pad):
    x = _f(x).copy()
    N, C, H, W = x.shape
    _st(_p(x), N * C, H, W, pad)
    return x


def sphere_cut_edge(x, pad):
    x = _f(x)
    N, C, H, W = x.shape
    out = np.zeros((N, C, H - 2 * pad, W - 2 * pad), np.float32)
    _sce(_p(x), _p(out), N * C, H, W, pad, 0)
    return out


def sphere_cut_edge_bwd(g, pad):
    g = _f(g)
    N, C, Ho, Wo = g.shape
    out = np.zeros((N, C, Ho + 2 * pad, Wo + 2 * pad), np.float32)
    _sce(_p(g), _p(out), N * C, Ho + 2 * pad, Wo + 2 * pad, pad, 1)
    return out


def sphere_lat_scale(x, weight, npart):
    x, weight = _f(x), _f(weight)
    N, C, H, W = x.shape
    out = np.zeros_like(x)
    _sls(_p(x), _p(weight), _p(out), N * C, H, W, npart)
    return out


def dtow(x, stride, d2w):
    x = _f(x)
    N, C, H, W = x.shape
    s = stride
    out = np.zeros((N, C // (s * s), H * s, W * s) if d2w else (N, C * s * s, H // s, W // s), np.float32)
    _dt(_p(x), _p(out), N, C, H, W, s, int(d2w))
    return out


# ------------------------------------------------------------------------------------------------ MultiProject
_pj_c = _sig("orc_projects_coords", [_P, _I, _I, _P, _P, _F, _I, _I])
_pj_f = _sig("orc_projects_forward", [_P, _P, _P, _I, _I, _I, _I, _I])
_pj_b = _sig("orc_projects_backward", [_P, _P, _P, _P, _I, _I, _I, _I, _I])
PROJECT_THETAS = [-0.5, 0, 0.5, 1, -0.5, 0, 0.5, 1, -0.5, 0, 0.5, 1, 0, 0]   # MultiProject.py:28-29
PROJECT_PHIS = [0, 0, 0, 0, 0.25, 0.25, 0.25, 0.25, -0.25, -0.25, -0.25, -0.25, 0.5, -0.5]


def projects_coords(h_out, w_out, fov, H, W, thetas=PROJECT_THETAS, phis=PROJECT_PHIS):
    tf = np.zeros((14, h_out * w_out, 2), np.float32)
    th, ph = np.asarray(thetas, np.float32), np.asarray(phis, np.float32)
    _pj_c(_p(tf), h_out, w_out, _p(th), _p(ph), float(fov), H, W)
    return tf


def projects_forward(x, h_out, w_out, fov, near=False):
    x = _f(x)
    N, C, H, W = x.shape
    tf = projects_coords(h_out, w_out, fov, H, W)
    out = np.zeros((14 * N, C, h_out, w_out), np.float32)
    _pj_f(_p(x), _p(tf), _p(out), N * C, H, W, h_out * w_out, int(near))
    return out


def projects_backward(top, shape, h_out, w_out, fov, near=False):
    top = _f(top)
    N, C, H, W = shape
    tf = projects_coords(h_out, w_out, fov, H, W)
    grad, count = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
    _pj_b(_p(top), _p(tf), _p(grad), _p(count), N * C, H, W, h_out * w_out, int(near))
    return grad, count


# ------------------------------------------------------------------------------------------------ coders
for _n, _a, _r in (("orc_ac_new", [], _P), ("orc_ac_free", [_P], None), ("orc_ac_error", [_P], _I),
                   ("orc_ac_start_encoder", [_P], None), ("orc_ac_encode_rows", [_P, _P, _I, _P, _P, _I], None),
                   ("orc_ac_end_encoder", [_P], ctypes.c_long), ("orc_ac_get_bytes", [_P, _P, ctypes.c_long], ctypes.c_long),
                   ("orc_ac_start_decoder", [_P, _P, ctypes.c_long], None),
                   ("orc_ac_decode_rows", [_P, _P, _I, _P, _I, _F, _P], None)):
    _sig(_n, _a, _r)


class OracleCoder(object):
    """Restated coder (lic360_oracle.c). encode/decode whole streams in memory."""

    def __init__(self, fill=3.5):
        self.h = ctypes.c_void_p(_L.orc_ac_new())
        self.fill = fill

    def __del__(self):
        if getattr(self, "h", None):
            _L.orc_ac_free(self.h)
            self.h = None

    def start_encoder(self):
        _L.orc_ac_start_encoder(self.h)

    def encode_rows(self, table, label, mask=None):
        table = np.ascontiguousarray(table, np.int32)
        label = np.ascontiguousarray(label, np.int32)
        mask = None if mask is None else _f(mask)
        _L.orc_ac_encode_rows(self.h, _p(table), table.shape[1] - 1, _p(label), _p(mask), table.shape[0])
        if _L.orc_ac_error(self.h):
            raise RuntimeError("oracle coder: symbol has zero frequency")

    def end_encoder(self):
        n = _L.orc_ac_end_encoder(self.h)
        buf = np.zeros(n, np.uint8)
        _L.orc_ac_get_bytes(self.h, _p(buf), n)
        return buf.tobytes()

    def start_decoder(self, data):
        self._in = np.frombuffer(data, np.uint8).copy()
        _L.orc_ac_start_decoder(self.h, _p(self._in), len(self._in))

    def decode_rows(self, table, mask=None):
        table = np.ascontiguousarray(table, np.int32)
        mask = None if mask is None else _f(mask)
        out = np.zeros(table.shape[0], np.float32)
        _L.orc_ac_decode_rows(self.h, _p(table), table.shape[1] - 1, _p(mask), table.shape[0], self.fill, _p(out))
        return out


def have_ref_coder():
    return os.path.exists(_REF_CODER)


class RefCoder(object):
    """The reference's own ArithmeticEncoder/Decoder + BitIoStream classes (oracle/_ref/libref_coder.so)."""

    _lib = None

    def __init__(self, fill=3.5):
        if RefCoder._lib is None:
            R = ctypes.CDLL(_REF_CODER)
            R.refcoder_create.restype = _P
            R.refcoder_create.argtypes = [_F]
            R.refcoder_destroy.argtypes = [_P]
            R.refcoder_error.restype = ctypes.c_char_p
            R.refcoder_error.argtypes = [_P]
            R.refcoder_start_encoder.argtypes = [_P]
            R.refcoder_encode_rows.argtypes = [_P, _P, _I, _P, _P, _I]
            R.refcoder_end_encoder.argtypes = [_P]
            R.refcoder_end_encoder.restype = ctypes.c_long
            R.refcoder_get_bytes.argtypes = [_P, _P, ctypes.c_long]
            R.refcoder_get_bytes.restype = ctypes.c_long
            R.refcoder_start_decoder.argtypes = [_P, _P, ctypes.c_long]
            R.refcoder_decode_rows.argtypes = [_P, _P, _I, _P, _I, _P]
            RefCoder._lib = R
        self.R = RefCoder._lib
        self.h = ctypes.c_void_p(self.R.refcoder_create(fill))

    def __del__(self):
        if getattr(self, "h", None):
            self.R.refcoder_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc:
            raise RuntimeError("reference coder: " + self.R.refcoder_error(self.h).decode())

    def start_encoder(self):
        self._chk(self.R.refcoder_start_encoder(self.h))

    def encode_rows(self, table, label, mask=None):
        table = np.ascontiguousarray(table, np.int32)
        label = np.ascontiguousarray(label, np.int32)
        mask = None if mask is None else _f(mask)
        self._chk(self.R.refcoder_encode_rows(self.h, _p(table), table.shape[1] - 1, _p(label), _p(mask), table.shape[0]))

    def end_encoder(self):
        n = self.R.refcoder_end_encoder(self.h)
        if n < 0:
            self._chk(1)
        buf = np.zeros(n, np.uint8)
        self.R.refcoder_get_bytes(self.h, _p(buf), n)
        return buf.tobytes()

    def start_decoder(self, data):
        self._in = np.frombuffer(data, np.uint8).copy()
        self._chk(self.R.refcoder_start_decoder(self.h, _p(self._in), len(self._in)))

    def decode_rows(self, table, mask=None):
        table = np.ascontiguousarray(table, np.int32)
        mask = None if mask is None else _f(mask)
        out = np.zeros(table.shape[0], np.float32)
        self._chk(self.R.refcoder_decode_rows(self.h, _p(table), table.shape[1] - 1, _p(mask), table.shape[0], _p(out)))
        return out
