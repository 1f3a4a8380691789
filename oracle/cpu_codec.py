"""TEST INFRASTRUCTURE ONLY -- CPU rendition of the whole entropy codec (both streams, encode and decode) on top of the
oracle restatement (oracle/lic360_oracle.c, OpenMP) and the reference's own host coder when oracle/_ref is built.

Used by tests (config 1 of BASELINE.json) and by bench.py's `cpu_baseline` / `--impl reference` legs. It follows the
same drivers as the product (test/lic360_demo.py:124-141,173-189,220-238,272-290) so it produces the same two streams.
"""
import numpy as np

from oracle import oracle as O

LAYER_KEYS = ['net.0'] + ['net.%d.conv%d' % (b, c) for b in range(1, 6) for c in (1, 2)] + ['net.6']


def params_to_numpy(params):
    return {k: {kk: vv.detach().cpu().numpy() for kk, vv in v.items()} for k, v in params.items()}


def _coder():
    return O.RefCoder() if O.have_ref_coder() else O.OracleCoder()


class CpuCodec(object):
    def __init__(self, params_np):
        self.p = params_np

    # ---- networks -----------------------------------------------------------------------------------------------
    def _layer(self, p, i):
        k = LAYER_KEYS[i]
        return p[k + '.weight'], p[k + '.bias'], p.get(k + '.relu')

    def _net_ec(self, p, x, G, nsets):
        w, b, a = self._layer(p, 0)
        y = O.cconv_ec(x, w, b, a, G, 5, nsets)
        for blk in range(5):
            w, b, a = self._layer(p, 1 + 2 * blk)
            t = O.cconv_ec(y, w, b, a, G, 6, nsets)
            w, b, a = self._layer(p, 2 + 2 * blk)
            y = O.cconv_ec(t, w, b, a, G, 6, nsets) + y
        w, b, a = self._layer(p, 11)
        return O.cconv_ec(y, w, b, None, G, 6, nsets)

    def _net_dc_step(self, p, frames, G, nsets, idx, plan, psum):
        w, b, a = self._layer(p, 0)
        O.cconv_dc_step(frames[0], w, b, a, frames[1], G, 5, nsets, idx, plan, psum)
        for blk in range(5):
            l1, l2 = 1 + 2 * blk, 2 + 2 * blk
            w, b, a = self._layer(p, l1)
            O.cconv_dc_step(frames[l1], w, b, a, frames[l1 + 1], G, 6, nsets, idx, plan, psum)
            w, b, a = self._layer(p, l2)
            O.cconv_dc_step(frames[l2], w, b, a, frames[l2 + 1], G, 6, nsets, idx, plan, psum)
            O.tile_add(frames[l2 + 1], frames[l1], G, idx, plan, psum)
        w, b, a = self._layer(p, 11)
        O.cconv_dc_step(frames[11], w, b, None, frames[12], G, 6, nsets, idx, plan, psum)

    @staticmethod
    def _frames(p, nsets, G, H, W):
        fr = [np.zeros((nsets, G, H, W), np.float32)]
        for k in LAYER_KEYS:
            fr.append(np.zeros((nsets, p[k + '.weight'].shape[-4], H, W), np.float32))
        return fr

    # ---- code stream --------------------------------------------------------------------------------------------
    def encode_code(self, q, mask):
        G, (H, W) = 48, q.shape[2:]
        idx, plan = O.code_contex(H, W)
        x = np.concatenate([(q - 3.5) * mask] * 3).astype(np.float32)
        y = self._net_ec(self.p['code'], x, G, 3)
        c = _coder()
        c.start_encoder()
        buf = np.zeros(9 * H * W, np.float32)
        lab, mk = np.zeros(H * W, np.float32), np.zeros(H * W, np.float32)
        stride = 3 * H * W
        for psum in range(H + W + G - 2):
            n = O.tile_extract_batch(y, buf, G, idx, plan, psum)
            tab, _, _ = O.gmm_table(buf[:n * 3].reshape(n, 3), buf[stride:stride + n * 3].reshape(n, 3), buf[2 * stride:2 * stride + n * 3].reshape(n, 3))
            O.tile_extract(q, lab, G, True, idx, plan, psum)
            O.tile_extract(mask, mk, G, True, idx, plan, psum)
            c.encode_rows(tab.astype(np.int32), lab[:n].astype(np.int32), mk[:n])
        return c.end_encoder()

    def decode_code(self, data, mask):
        G, (H, W) = 48, mask.shape[2:]
        idx, plan = O.code_contex(H, W)
        c = _coder()
        c.start_decoder(data)
        fr = self._frames(self.p['code'], 3, G, H, W)
        buf = np.zeros(9 * H * W, np.float32)
        mk = np.zeros(H * W, np.float32)
        stride = 3 * H * W
        pout = np.zeros(0, np.float32)
        for psum in range(H + W + G - 2):
            O.tile_input(pout, fr[0], 1, G, H, W, -3.5, 1.0, 3, idx, plan, psum)
            self._net_dc_step(self.p['code'], fr, G, 3, idx, plan, psum)
            n = O.tile_extract_batch(fr[12], buf, G, idx, plan, psum)
            tab, _, _ = O.gmm_table(buf[:n * 3].reshape(n, 3), buf[stride:stride + n * 3].reshape(n, 3), buf[2 * stride:2 * stride + n * 3].reshape(n, 3))
            O.tile_extract(mask, mk, G, True, idx, plan, psum)
            pout = c.decode_rows(tab.astype(np.int32), mk[:n])
        O.tile_input(pout, fr[0], 1, G, H, W, -3.5, 1.0, 3, idx, plan, H + W + G - 2)
        return fr[0][0:1] + 3.5 * mask

    # ---- importance stream --------------------------------------------------------------------------------------
    def encode_imp(self, lv):
        H, W = lv.shape[2:]
        idx, plan = O.code_contex(H, W)
        y = self._net_ec(self.p['imp'], O.scale(lv, -1.0, 2.0 / 47), 1, 1)
        c = _coder()
        c.start_encoder()
        buf, lab = np.zeros(49 * H * W, np.float32), np.zeros(H * W, np.float32)
        for psum in range(H + W - 1):
            n = O.tile_extract(y, buf, 1, True, idx, plan, psum)
            tab = O.entropy_table(buf[:n * 49].reshape(n, 49))
            O.tile_extract(lv, lab, 1, True, idx, plan, psum)
            c.encode_rows(tab.astype(np.int32), lab[:n].astype(np.int32))
        return c.end_encoder()

    def decode_imp(self, data, H, W):
        idx, plan = O.code_contex(H, W)
        c = _coder()
        c.start_decoder(data)
        fr = self._frames(self.p['imp'], 1, 1, H, W)
        buf = np.zeros(49 * H * W, np.float32)
        levels = np.zeros((1, 1, H, W), np.float32)
        pout = np.zeros(0, np.float32)
        for psum in range(H + W - 1):
            O.tile_input(pout, fr[0], 1, 1, H, W, -1.0, 2.0 / 47, 1, idx, plan, psum)
            O.tile_input(pout, levels, 1, 1, H, W, 0.0, 1.0, 1, idx, plan, psum) if psum else None
            self._net_dc_step(self.p['imp'], fr, 1, 1, idx, plan, psum)
            n = O.tile_extract(fr[12], buf, 1, True, idx, plan, psum)
            pout = c.decode_rows(O.entropy_table(buf[:n * 49].reshape(n, 49)).astype(np.int32))
        O.tile_input(pout, levels, 1, 1, H, W, 0.0, 1.0, 1, idx, plan, H + W - 1)
        return levels

    # ---- whole image --------------------------------------------------------------------------------------------
    def encode(self, q, mask, lv):
        return self.encode_imp(lv), self.encode_code(q, mask)

    def decode(self, imp_bytes, code_bytes, h, w):
        lv = self.decode_imp(imp_bytes, h, w)
        mask = O.dtow(O.imp2mask(lv, 192, 48), 2, True)
        return self.decode_code(code_bytes, mask), mask


class CpuCodecFast(CpuCodec):
    """The CPU arm that bench.py times: the same drivers, streams and tables as CpuCodec, with the context conv in its CPU-friendly
    form (oracle.c "CPU BASELINE form": channel-last activations, one contiguous fp32 dot product per tap over the masked channel
    range, OpenMP over positions).  Encoder (whole frame) and decoder (wavefront) call the same per-output routine, so the codec
    round-trips exactly; against CpuCodec its conv differs in the last bits (fp32 tree vs double), like any two float-tier forms."""

    def __init__(self, params_np):
        CpuCodec.__init__(self, params_np)
        self.wt = {}
        for key, nsets in (('code', 3), ('imp', 1)):
            self.wt[key] = [O.pack_weights_cl(self.p[key][k + '.weight'], nsets) for k in LAYER_KEYS]

    def _lay(self, key, i):
        p, k = self.p[key], LAYER_KEYS[i]
        return self.wt[key][i], p[k + '.bias'], p.get(k + '.relu')

    def _net_ec_cl(self, key, x, G):
        w, b, a = self._lay(key, 0)
        y = O.cconv_ec_cl(x, w, b, a, None, G, 5)
        for blk in range(5):
            w, b, a = self._lay(key, 1 + 2 * blk)
            t = O.cconv_ec_cl(y, w, b, a, None, G, 6)
            w, b, a = self._lay(key, 2 + 2 * blk)
            y = O.cconv_ec_cl(t, w, b, a, y, G, 6)   # conv2(conv1(y)) + y, lic360_demo.py:39-41
        w, b, a = self._lay(key, 11)
        return O.cconv_ec_cl(y, w, b, None, None, G, 6)

    def _net_dc_step_cl(self, key, fr, G, idx, plan, psum):
        for l in range(12):
            w, b, a = self._lay(key, l)
            resid = fr[l - 1] if (l >= 2 and l <= 10 and l % 2 == 0) else None   # TileAdd of the residual blocks, fused
            O.cconv_dc_step_cl(fr[l], w, b, a if l < 11 else None, resid, fr[l + 1], G, 5 if l == 0 else 6,
                               idx, plan, psum)

    @staticmethod
    def _frames_cl(p, nsets, G, H, W):
        fr = [np.zeros((nsets, H, W, G), np.float32)]
        for k in LAYER_KEYS:
            fr.append(np.zeros((nsets, H, W, p[k + '.weight'].shape[-4]), np.float32))
        return fr

    @staticmethod
    def _slab(plan, H, W, G, p):
        la, lb = max(0, p - G + 1), min(p, H + W - 2)
        return int(plan[la]), int(plan[lb + 1] - plan[la])

    def encode_code(self, q, mask):
        G, (H, W) = 48, q.shape[2:]
        idx, plan = O.code_contex(H, W)
        x1 = np.ascontiguousarray(((q - 3.5) * mask).astype(np.float32).transpose(0, 2, 3, 1))
        y = self._net_ec_cl('code', np.ascontiguousarray(np.concatenate([x1] * 3)), G)
        c = _coder()
        c.start_encoder()
        buf = np.zeros(9 * H * W, np.float32)
        lab, mk = np.zeros(H * W, np.float32), np.zeros(H * W, np.float32)
        for psum in range(H + W + G - 2):
            n = O.tile_extract_cl(y, buf, G, idx, plan, psum) // 3
            rows = buf[:9 * n].reshape(3, n, 3)
            tab, _, _ = O.gmm_table(rows[0], rows[1], rows[2])
            O.tile_extract(q, lab, G, True, idx, plan, psum)
            O.tile_extract(mask, mk, G, True, idx, plan, psum)
            c.encode_rows(tab.astype(np.int32), lab[:n].astype(np.int32), mk[:n])
        return c.end_encoder()

    def decode_code(self, data, mask):
        G, (H, W) = 48, mask.shape[2:]
        idx, plan = O.code_contex(H, W)
        c = _coder()
        c.start_decoder(data)
        fr = self._frames_cl(self.p['code'], 3, G, H, W)
        buf = np.zeros(9 * H * W, np.float32)
        mk = np.zeros(H * W, np.float32)
        pout = np.zeros(0, np.float32)

        def scatter(psum):  # TileInput (tile_input_cuda.cu:27-43): symbols of step psum - 1, value = s - 3.5 (exact), 3 replicas
            if psum == 0:
                return
            st, ln = self._slab(plan, H, W, G, psum - 1)
            th, tw = idx[st:st + ln], idx[H * W + st:H * W + st + ln]
            fr[0][:, th, tw, psum - 1 - th - tw] = (pout[:ln] - np.float32(3.5))[None, :]

        for psum in range(H + W + G - 2):
            scatter(psum)
            self._net_dc_step_cl('code', fr, G, idx, plan, psum)
            n = O.tile_extract_cl(fr[12], buf, G, idx, plan, psum) // 3
            rows = buf[:9 * n].reshape(3, n, 3)
            tab, _, _ = O.gmm_table(rows[0], rows[1], rows[2])
            O.tile_extract(mask, mk, G, True, idx, plan, psum)
            pout = c.decode_rows(tab.astype(np.int32), mk[:n])
        scatter(H + W + G - 2)
        return np.ascontiguousarray(fr[0][0:1].transpose(0, 3, 1, 2)) + 3.5 * mask

    def encode_imp(self, lv):
        H, W = lv.shape[2:]
        idx, plan = O.code_contex(H, W)
        y = self._net_ec_cl('imp', O.scale(lv, -1.0, 2.0 / 47).reshape(1, H, W, 1), 1)   # one channel: NCHW == NHWC
        c = _coder()
        c.start_encoder()
        buf, lab = np.zeros(49 * H * W, np.float32), np.zeros(H * W, np.float32)
        for psum in range(H + W - 1):
            n = O.tile_extract_cl(y, buf, 1, idx, plan, psum)
            tab = O.entropy_table(buf[:n * 49].reshape(n, 49))
            O.tile_extract(lv, lab, 1, True, idx, plan, psum)
            c.encode_rows(tab.astype(np.int32), lab[:n].astype(np.int32))
        return c.end_encoder()

    def decode_imp(self, data, H, W):
        idx, plan = O.code_contex(H, W)
        c = _coder()
        c.start_decoder(data)
        fr = self._frames_cl(self.p['imp'], 1, 1, H, W)
        buf = np.zeros(49 * H * W, np.float32)
        levels = np.zeros((1, 1, H, W), np.float32)
        pout = np.zeros(0, np.float32)
        f0 = fr[0].reshape(1, 1, H, W)   # one channel: the same memory as an NCHW frame
        for psum in range(H + W - 1):
            O.tile_input(pout, f0, 1, 1, H, W, -1.0, 2.0 / 47, 1, idx, plan, psum)
            O.tile_input(pout, levels, 1, 1, H, W, 0.0, 1.0, 1, idx, plan, psum) if psum else None
            self._net_dc_step_cl('imp', fr, 1, idx, plan, psum)
            n = O.tile_extract_cl(fr[12], buf, 1, idx, plan, psum)
            pout = c.decode_rows(O.entropy_table(buf[:n * 49].reshape(n, 49)).astype(np.int32))
        O.tile_input(pout, levels, 1, 1, H, W, 0.0, 1.0, 1, idx, plan, H + W - 1)
        return levels
