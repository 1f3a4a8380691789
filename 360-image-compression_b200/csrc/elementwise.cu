// Streaming (HBM-bound) ops of the entropy path: quantisation, importance masking, context reshape/shift,
// weight masking, spherical pad/trim/crop/latitude scale, depth<->width shuffle.
// Replace the kernels of /root/reference/extension/{quant,dquant,imp_map,imp2mask,scale,context_reshape,
// contex_shift,mask_constrain,sphere_pad,sphere_trim,sphere_cut_edge,sphere_lat_scale,dtow}_cuda.cu.
// Style: 128-bit accesses along the contiguous (w) axis when the plane size and pointers allow it, grid sized in
// multiples of the SM count, per-plane parameters hoisted out of the element loop, border-only enumeration for the
// in-place sphere ops (the reference launches over every element to touch only the border).
#include "common.cuh"

namespace lic360 {

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ----------------------------------------------------------------------------------------- Scale
// scale_cuda.cu:24-29: x*scale + bias (one FFMA)
template <int V>
__global__ void scale_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, float bias, float scale) {
    const size_t nv = n / V;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
        if (V == 4) {
            float4 v = reinterpret_cast<const float4*>(in)[i];
            v.x = v.x * scale + bias; v.y = v.y * scale + bias; v.z = v.z * scale + bias; v.w = v.w * scale + bias;
            reinterpret_cast<float4*>(out)[i] = v;
        } else {
            out[i] = in[i] * scale + bias;
        }
    }
}

// ----------------------------------------------------------------------------------------- ImpMap / Imp2mask
// imp_map_cuda.cu:79-110: keep channel c iff c < (int)(imp*levels + 1e-5) * cpl  (float product, +1e-5 in double)
__device__ __forceinline__ int imp_channels(float imp, int levels, int cpl) { return static_cast<int>(imp * levels + 0.00001) * cpl; }
// imp2mask_cuda.cu:25-38: c < (int)(level + 1e-5) * cpn
__device__ __forceinline__ int imp2mask_channels(float lv, int cpn) { return static_cast<int>(lv + 1e-5) * cpn; }

// one CTA row = one (n,c) plane chunk; mode 0: ImpMap (x may be masked, mask optional), mode 1: Imp2mask
template <int V, int MODE>
__global__ void imp_mask_kernel(const float* __restrict__ x, const float* __restrict__ imp, float* __restrict__ out,
                                float* __restrict__ mask, int N, int C, int HW, int levels, int cpl) {
    const int per_plane = HW / V;
    const size_t total = (size_t)N * C * per_plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(i % per_plane);
        const size_t plane = i / per_plane;
        const int c = (int)(plane % C), n = (int)(plane / C);
        if (V == 4) {
            const float4 im = reinterpret_cast<const float4*>(imp + (size_t)n * HW)[s];
            const float iv[4] = {im.x, im.y, im.z, im.w};
            float ov[4], mv[4];
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == 0) xv = reinterpret_cast<const float4*>(x + plane * HW)[s];
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int ch = MODE == 0 ? imp_channels(iv[k], levels, cpl) : imp2mask_channels(iv[k], cpl);
                const bool keep = c < ch;
                mv[k] = keep ? 1.f : 0.f;
                ov[k] = keep ? xa[k] : 0.f;
            }
            if (MODE == 0) {
                reinterpret_cast<float4*>(out + plane * HW)[s] = make_float4(ov[0], ov[1], ov[2], ov[3]);
                if (mask) reinterpret_cast<float4*>(mask + plane * HW)[s] = make_float4(mv[0], mv[1], mv[2], mv[3]);
            } else {
                reinterpret_cast<float4*>(out + plane * HW)[s] = make_float4(mv[0], mv[1], mv[2], mv[3]);
            }
        } else {
            const float iv = imp[(size_t)n * HW + s];
            const int ch = MODE == 0 ? imp_channels(iv, levels, cpl) : imp2mask_channels(iv, cpl);
            const bool keep = c < ch;
            if (MODE == 0) {
                out[plane * HW + s] = keep ? x[plane * HW + s] : 0.f;
                if (mask) mask[plane * HW + s] = keep ? 1.f : 0.f;
            } else {
                out[plane * HW + s] = keep ? 1.f : 0.f;
            }
        }
    }
}

// imp_map_cuda.cu:27-69: |cos| latitude profile, normalised by its max, -> constrain (N,1,H) and alpha_t (H)
__global__ void imp_map_init_kernel(float* __restrict__ constrain, float* __restrict__ alpha_t, int N, int H, float alpha,
                                    float rt, float sc, float sw) {
    extern __shared__ float prof[];  // [H]
    __shared__ float smax;
    const float pi = acos(-1.0);
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float a = cos((0.5 - (h + 0.5) / H) * pi);
        prof[h] = a < 0 ? -a : a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = prof[0];
        for (int h = 1; h < H; h++) m = fmaxf(m, prof[h]);
        smax = m;
    }
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        const float t = prof[h] / smax;
        alpha_t[h] = alpha / (t * sw + 1 - sw);
        const float cst = rt * (t * sc + 1 - sc);
        for (int n = 0; n < N; n++) constrain[n * H + h] = cst;
    }
}

// imp_map_cuda.cu:139-153
__global__ void imp_map_bwd_data_kernel(const float* __restrict__ top, const float* __restrict__ imp,
                                        float* __restrict__ bottom, int N, int C, int HW, int levels, int cpl) {
    const size_t total = (size_t)N * C * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(i % HW);
        const int c = (int)((i / HW) % C), n = (int)(i / ((size_t)HW * C));
        const int ch = static_cast<int>(floor(imp[(size_t)n * HW + s] * levels)) * cpl;
        bottom[i] = c < ch ? top[i] : 0.f;
    }
}

// imp_map_cuda.cu:156-238; version 0:v1 1:v2 2:v3 3:v4 (switch at :264-293)
__global__ void imp_map_bwd_imp_kernel(const float* __restrict__ top, const float* __restrict__ imp,
                                       const float* __restrict__ sphere, const float* __restrict__ alpha_t,
                                       float* __restrict__ imp_diff, int N, int C, int H, int W, int levels, int cpl,
                                       float gamma, int version) {
    const int HW = H * W;
    for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < N * HW; index += gridDim.x * blockDim.x) {
        const int ps = index % HW, pn = index / HW, ph = ps / W;
        const float sc = sphere[index / W];
        const int ch = static_cast<int>(imp[index] * levels + 0.00001) * cpl;
        const float* t = top + (size_t)pn * C * HW + ps;
        if (version == 3) {
            const float decay = sc < 0 ? 0.1 : 1;
            const float cost = alpha_t[ph];
            float tmp = 0, tmax = -10000;
            int target = 0;
            for (int i = 0; i < C; i++) {
                tmp = tmp + fabs(t[(size_t)i * HW]) - cost * decay;
                if (tmp > tmax) { tmax = tmp; target = i; }
            }
            imp_diff[index] = target < ch ? gamma : (target > ch ? -gamma : 0.f);
        } else {
            float diff = 0;
            if (sc > 0) diff = version == 0 ? alpha_t[ph] * (C - ch) : alpha_t[ph];
            for (int i = (version == 2 ? 0 : ch); i < C; i++) diff -= fabs(t[(size_t)i * HW]);
            imp_diff[index] = diff;
        }
    }
}

// ----------------------------------------------------------------------------------------- Quant / Dquant
// quant_cuda.cu:35-43
__global__ void quant_levels_kernel(const float* __restrict__ wb, float* __restrict__ lv, int n, int L) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        lv[i] = (i % L == 0) ? wb[i] : expf(wb[i]);
}

// quant_cuda.cu:46-76. One CTA walks chunks of single (n,c) planes: the channel's levels sit in registers/smem and
// the histogram is accumulated in shared memory (integers in float: exact), one atomic per bin per chunk instead
// of one per element.
constexpr int QL_MAX = 32;
template <int V>
__global__ void __launch_bounds__(256) quant_fwd_kernel(const float* __restrict__ x, const float* __restrict__ lv,
                                                        float* __restrict__ y, float* __restrict__ q,
                                                        int32_t* __restrict__ qi, float* __restrict__ count, int NC, int C,
                                                        int HW, int L, int chunks_per_plane, int chunk_elems) {
    __shared__ float slv[QL_MAX];
    __shared__ int shist[QL_MAX];
    for (int job = blockIdx.x; job < NC * chunks_per_plane; job += gridDim.x) {
        const int plane = job / chunks_per_plane, chunk = job % chunks_per_plane;
        const int c = plane % C;
        __syncthreads();
        if ((int)threadIdx.x < L) { slv[threadIdx.x] = lv[c * L + threadIdx.x]; shist[threadIdx.x] = 0; }
        __syncthreads();
        const int e0 = chunk * chunk_elems, e1 = min(HW, e0 + chunk_elems);
        const size_t base = (size_t)plane * HW;
        for (int e = e0 + threadIdx.x * V; e < e1; e += blockDim.x * V) {
            float xv[4], yv[4], qv[4];
            if (V == 4) {
                const float4 t = *reinterpret_cast<const float4*>(x + base + e);
                xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
            } else xv[0] = x[base + e];
#pragma unroll
            for (int k = 0; k < V; k++) {
                float tmp = xv[k] - slv[0];
                int j;
                if (tmp < 0) { j = 0; yv[k] = slv[0]; }
                else {
                    j = 1;
                    for (; j < L; j++) { tmp -= slv[j]; if (tmp < 0) break; }
                    if (j == L) j--;
                    if (tmp + tmp + slv[j] < 0) { tmp = tmp + slv[j]; j--; }
                    yv[k] = xv[k] - tmp;
                }
                qv[k] = (float)j;
                // warp-aggregated histogram update: one shared-memory atomic per distinct level in the warp
                const unsigned act = __activemask();
                const unsigned same = __match_any_sync(act, j);
                if ((int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&shist[j], __popc(same));
            }
            if (V == 4) {
                *reinterpret_cast<float4*>(y + base + e) = make_float4(yv[0], yv[1], yv[2], yv[3]);
                if (q) *reinterpret_cast<float4*>(q + base + e) = make_float4(qv[0], qv[1], qv[2], qv[3]);
                *reinterpret_cast<int4*>(qi + base + e) = make_int4((int)qv[0], (int)qv[1], (int)qv[2], (int)qv[3]);
            } else {
                y[base + e] = yv[0];
                if (q) q[base + e] = qv[0];
                qi[base + e] = (int)qv[0];
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < L && shist[threadIdx.x] != 0) atomicAdd(count + c * L + threadIdx.x, -(float)shist[threadIdx.x]);
    }
}

// The configured quantiser (8 levels, test/model_zoo.py:342) on vectorisable planes: one WARP per 1024-element piece of an (n,c) plane,
// no CTA barriers; the channel's 8 levels sit in registers, every lane keeps a private histogram in two 64-bit words (4 x 16-bit
// counters each: a lane sees at most 32 elements of a piece, a warp 1024), one shuffle reduction and at most 8 global atomics per
// piece.  Same arithmetic per element as quant_fwd_kernel (quant_cuda.cu:46-76).
constexpr int Q8_PIECE = 1024;
__global__ void __launch_bounds__(256) quant_fwd8_kernel(const float* __restrict__ x, const float* __restrict__ lv, float* __restrict__ y,
                                                         float* __restrict__ q, int32_t* __restrict__ qi, float* __restrict__ count,
                                                         int NC, int C, int HW, int pieces_per_plane) {
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    for (int job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); job < NC * pieces_per_plane; job += warps_total) {
        const int plane = job / pieces_per_plane, piece = job % pieces_per_plane;
        const int c = plane % C;
        float s[8];
#pragma unroll
        for (int j = 0; j < 8; j++) s[j] = __ldg(lv + c * 8 + j);
        const int e0 = piece * Q8_PIECE, e1 = min(HW, e0 + Q8_PIECE);
        const size_t base = (size_t)plane * HW;
        unsigned long long h0 = 0, h1 = 0;
        for (int e = e0 + lane * 4; e < e1; e += 128) {
            const float4 t = *reinterpret_cast<const float4*>(x + base + e);
            const float xv[4] = {t.x, t.y, t.z, t.w};
            float yv[4], qv[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float tmp = xv[k] - s[0];
                int j;
                if (tmp < 0) { j = 0; yv[k] = s[0]; }
                else {
                    j = 1;
#pragma unroll
                    for (int jj = 1; jj < 8; jj++) {   // first j with a negative remainder; lanes that found it stop subtracting
                        if (j == jj) { tmp -= s[jj]; if (!(tmp < 0)) j = jj + 1; }
                    }
                    if (j == 8) j--;
                    float sj = s[1];
#pragma unroll
                    for (int jj = 2; jj < 8; jj++) sj = j == jj ? s[jj] : sj;
                    if (tmp + tmp + sj < 0) { tmp = tmp + sj; j--; }
                    yv[k] = xv[k] - tmp;
                }
                qv[k] = (float)j;
                if (j < 4) h0 += 1ull << (16 * j); else h1 += 1ull << (16 * (j - 4));
            }
            *reinterpret_cast<float4*>(y + base + e) = make_float4(yv[0], yv[1], yv[2], yv[3]);
            if (q) *reinterpret_cast<float4*>(q + base + e) = make_float4(qv[0], qv[1], qv[2], qv[3]);
            *reinterpret_cast<int4*>(qi + base + e) = make_int4((int)qv[0], (int)qv[1], (int)qv[2], (int)qv[3]);
        }
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) { h0 += __shfl_xor_sync(0xffffffffu, h0, k); h1 += __shfl_xor_sync(0xffffffffu, h1, k); }
        if (lane < 8) {
            const unsigned n = (unsigned)(((lane < 4 ? h0 : h1) >> (16 * (lane & 3))) & 0xFFFFull);
            if (n) atomicAdd(count + c * 8 + lane, -(float)n);
        }
    }
}

// quant_cuda.cu:88-117 (dead-level repair + histogram decay), one thread per channel / element
__global__ void quant_check_weight_kernel(float* __restrict__ weight, const float* __restrict__ count, int C, int levels) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C; i += gridDim.x * blockDim.x) {
        int j = levels - 1;
        for (; j > 1; j--) if (count[i * levels + j] >= 1e-3) break;
        float tmp = weight[i * levels + j] - log(static_cast<float>(levels - j));
        for (; j < levels; j++) weight[i * levels + j] = tmp;
        if (count[i * levels] < 1e-3) {
            weight[i * levels] = weight[i * levels] + exp(weight[i * levels + 1]);
            tmp = log((exp(weight[i * levels + 1]) + exp(weight[i * levels + 2])) / 2);
            weight[i * levels + 1] = tmp;
            weight[i * levels + 2] = tmp;
        }
    }
}

// quant_cuda.cu:181-204: weight_diff[c][j] = sum over elements with index >= j of (y - x), then * level (j > 0).
// One CTA per channel; per-index partial sums in shared memory, suffix-summed at the end.
__global__ void __launch_bounds__(256) quant_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                               const int32_t* __restrict__ qi, const float* __restrict__ lv,
                                                               float* __restrict__ wdiff, int N, int C, int HW, int L) {
    __shared__ float ssum[QL_MAX];
    const int c = blockIdx.x;
    if ((int)threadIdx.x < L) ssum[threadIdx.x] = 0.f;
    __syncthreads();
    float loc[QL_MAX];
    for (int j = 0; j < L; j++) loc[j] = 0.f;
    for (int n = 0; n < N; n++) {
        const size_t base = ((size_t)n * C + c) * HW;
        for (int e = threadIdx.x; e < HW; e += blockDim.x) {
            const float d = y[base + e] - x[base + e];
            const int j = qi[base + e];
#pragma unroll
            for (int k = 0; k < QL_MAX; k++) if (k == j) loc[k] += d;
        }
    }
    for (int j = 0; j < L; j++) {
        float v = loc[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&ssum[j], v);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float run = 0.f;
        for (int j = L - 1; j >= 0; j--) {
            run += ssum[j];
            wdiff[c * L + j] = j == 0 ? run : run * lv[c * L + j];
        }
    }
}

// quant_cuda.cu:207-235 + the copy at :256: bottom = top_diff0 (+ alpha*top_diff1/beta when ntop > 1)
__global__ void quant_bwd_data_kernel(const float* __restrict__ td0, const float* __restrict__ td1,
                                      const float* __restrict__ x, const float* __restrict__ y,
                                      const int32_t* __restrict__ qi, const float* __restrict__ lv,
                                      float* __restrict__ bottom, size_t total, int C, int HW, int level, float alpha) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float g = td0[i];
        if (td1) {
            const int tc = (int)((i / HW) % C);
            const int qq = qi[i];
            const float* w = lv + tc * level;
            float beta = 1.0;
            if (y[i] < x[i]) beta = qq < level - 1 ? w[qq + 1] : 10000;
            else if (y[i] > x[i]) beta = qq > 0 ? w[qq] : 10000;
            else {
                if (qq == 0) beta = w[qq + 1];
                else if (qq < level - 1) beta = (w[qq] + w[qq + 1]) / 2.0;
                else beta = w[qq];
            }
            if (beta < 0.001) beta = 0.001;
            g = g + alpha * td1[i] / beta;
        }
        bottom[i] = g;
    }
}

// dquant_cuda.cu:24-31
__global__ void dquant_levels_kernel(const float* __restrict__ wb, float* __restrict__ cum, int C, int level) {
    for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < C; index += gridDim.x * blockDim.x) {
        cum[index * level] = wb[index * level];
        for (int i = 1; i < level; i++) cum[index * level + i] = cum[index * level + i - 1] + exp(wb[index * level + i]);
    }
}

// dquant_cuda.cu:34-47: y = mask > 0 ? cum[c][(int)(q + 1e-5)] : cum[c][0]
template <int V>
__global__ void dquant_fwd_kernel(const float* __restrict__ qv, const float* __restrict__ mask, const float* __restrict__ cum,
                                  float* __restrict__ y, int NC, int C, int HW, int L) {
    const int per_plane = HW / V;
    const size_t total = (size_t)NC * per_plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t plane = i / per_plane;
        const int s = (int)(i % per_plane);
        const float* w = cum + (plane % C) * L;
        if (V == 4) {
            const float4 a = reinterpret_cast<const float4*>(qv + plane * HW)[s];
            const float4 m = reinterpret_cast<const float4*>(mask + plane * HW)[s];
            float4 o;
            o.x = m.x > 0 ? __ldg(w + static_cast<int>(a.x + 0.00001)) : __ldg(w);
            o.y = m.y > 0 ? __ldg(w + static_cast<int>(a.y + 0.00001)) : __ldg(w);
            o.z = m.z > 0 ? __ldg(w + static_cast<int>(a.z + 0.00001)) : __ldg(w);
            o.w = m.w > 0 ? __ldg(w + static_cast<int>(a.w + 0.00001)) : __ldg(w);
            reinterpret_cast<float4*>(y + plane * HW)[s] = o;
        } else {
            const size_t e = plane * HW + s;
            y[e] = mask[e] > 0 ? __ldg(w + static_cast<int>(qv[e] + 0.00001)) : __ldg(w);
        }
    }
}

// ----------------------------------------------------------------------------------------- context reshape / shift / mask
// context_reshape_cuda.cu:30-41,63-74: (N, G*cpg, H, W) <-> (N*G*H*W, cpg). One thread per (n,g,s): cpg strided
// (warp-coalesced) accesses on the NCHW side, cpg consecutive floats on the row side.
__global__ void context_reshape_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int G, int cpg, int HW,
                                       int backward) {
    const size_t total = (size_t)N * G * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(i % HW);
        const size_t ng = i / HW;
        for (int c = 0; c < cpg; c++) {
            const size_t a = (ng * cpg + c) * HW + s;  // NCHW side
            const size_t b = i * cpg + c;              // row side
            if (backward) out[a] = in[b]; else out[b] = in[a];
        }
    }
}

// contex_shift_cuda.cu:36-64: row h of channel c goes to row h + w + c/cpn of the skewed tensor
__global__ void contex_shift_kernel(const float* __restrict__ in, float* __restrict__ out, int NC, int C, int H, int W,
                                    int Hs, int cpn, int mode) {
    const size_t total = (size_t)NC * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W), h = (int)((i / W) % H);
        const size_t nc = i / ((size_t)W * H);
        const int c = (int)(nc % C);
        const size_t b = (nc * Hs + h + w + c / cpn) * W + w;
        if (mode == 0) out[b] = in[i]; else out[i] = in[b];
    }
}

// mask_constrain_cuda.cu:17-41
__global__ void mask_constrain_kernel(float* __restrict__ w, int total, int Cin, int k, int gi, int go, int constrain) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kw = i % k, kh = (i / k) % k;
        const int tc = (i / (k * k)) % Cin / gi;
        const int tn = i / (k * k) / Cin / go;
        const int lhs = kw + kh + tc, rhs = tn + k - 1;
        if (constrain == 5 ? lhs >= rhs : lhs > rhs) w[i] = 0.f;
    }
}

// ----------------------------------------------------------------------------------------- sphere ops
// source coordinate of a padded position: longitude wrap; beyond the poles reflect the row AND mirror the
// column (sphere_pad_cuda.cu:33-41)
__device__ __forceinline__ void sphere_src(int ph, int pw, int H, int W, int pad, int& th, int& tw) {
    th = ph - pad;
    tw = (pw - pad + W) % W;
    if (th < 0 || th >= H) { th = (2 * H - 1 - th) % H; tw = (2 * W - 1 - tw) % W; }
}

// one warp per OUTPUT row (8 rows per CTA): the row's source row / pole flip is decided once, the lanes walk the
// columns with 32-bit arithmetic only, loads and stores are contiguous
constexpr int ROWS_PER_CTA = 8;

__global__ void __launch_bounds__(32 * ROWS_PER_CTA) sphere_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int NC, int H,
                                                                     int W, int pad) {
    const int Ho = H + 2 * pad, Wo = W + 2 * pad;
    const int row = blockIdx.x * ROWS_PER_CTA + threadIdx.y;
    if (row >= NC * Ho) return;
    const int n = row / Ho, ph = row % Ho;
    int th = ph - pad;
    const bool flip = th < 0 || th >= H;
    if (flip) th = (2 * H - 1 - th) % H;
    const float* src = in + ((size_t)n * H + th) * W;
    float* dst = out + (size_t)row * Wo;
    for (int pw = threadIdx.x; pw < Wo; pw += 32) {
        int tw = (pw - pad + W) % W;
        if (flip) tw = (2 * W - 1 - tw) % W;
        dst[pw] = __ldg(src + tw);
    }
}

// enumerate only the border of a padded (Hp, Wp) plane: b in [0, Hp*Wp - H*W)
__device__ __forceinline__ void border_coord(int b, int Hp, int Wp, int pad, int& ph, int& pw) {
    const int top = pad * Wp;
    if (b < top) { ph = b / Wp; pw = b % Wp; return; }
    b -= top;
    if (b < top) { ph = Hp - pad + b / Wp; pw = b % Wp; return; }
    b -= top;
    ph = pad + b / (2 * pad);
    const int r = b % (2 * pad);
    pw = r < pad ? r : Wp - 2 * pad + r;
}

// sphere_pad_cuda.cu:48-65 (in place) and sphere_trim_cuda.cu:17-26 (zero): border elements only
template <int MODE>  // 0: sphere pad in place, 1: trim
__global__ void sphere_border_kernel(float* __restrict__ data, int NC, int Hp, int Wp, int pad) {
    const int H = Hp - 2 * pad, W = Wp - 2 * pad;
    const int nb = Hp * Wp - H * W;
    const size_t total = (size_t)NC * nb;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / nb;
        int ph, pw;
        border_coord((int)(i % nb), Hp, Wp, pad, ph, pw);
        float* plane = data + n * Hp * Wp;
        if (MODE == 1) plane[ph * Wp + pw] = 0.f;
        else {
            int th, tw;
            sphere_src(ph, pw, H, W, pad, th, tw);
            plane[ph * Wp + pw] = plane[(th + pad) * Wp + tw + pad];
        }
    }
}

// sphere_pad_cuda.cu:108-170
__global__ void sphere_pad_bwd_kernel(float* __restrict__ bottom, float* __restrict__ top, int NC, int H, int W, int pad,
                                      int inplace) {
    const int Ho = H + 2 * pad, Wo = W + 2 * pad;
    const size_t total = (size_t)NC * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int pw = (int)(i % W), ph = (int)((i / W) % H);
        const size_t n = i / ((size_t)W * H);
        const float* t = top + n * Ho * Wo;
        int th = ph + pad, tw = pw + pad;
        const bool wb = pw < pad || pw >= W - pad, hb = ph < pad || ph >= H - pad;
        if (inplace && !wb && !hb) continue;
        float acc = t[th * Wo + tw];
        if (wb) {
            tw = pw < pad ? pw + W + pad : pw - W + pad;
            acc += t[th * Wo + tw];
        }
        if (hb) {
            th = ph < pad ? pad - ph - 1 : (2 * H - 1 - ph) + pad;
            tw = W - 1 - pw + pad;
            acc += t[th * Wo + tw];
            if (wb) {
                tw = pw < pad ? pad - pw - 1 : 2 * W - pw - 1 + pad;
                acc += t[th * Wo + tw];
            }
        }
        if (inplace) top[n * Ho * Wo + (ph + pad) * Wo + pw + pad] = acc; else bottom[i] = acc;
    }
}

// sphere_cut_edge_cuda.cu:31-41 (crop) / :64-78 (zero-pad back)
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) sphere_cut_edge_kernel(const float* __restrict__ in, float* __restrict__ out, int NC,
                                                                          int H, int W, int pad, int backward) {
    const int Ho = H - 2 * pad, Wo = W - 2 * pad;
    const int row = blockIdx.x * ROWS_PER_CTA + threadIdx.y;  // output row
    if (!backward) {
        if (row >= NC * Ho) return;
        const int n = row / Ho, h = row % Ho;
        const float* src = in + ((size_t)n * H + h + pad) * W + pad;
        float* dst = out + (size_t)row * Wo;
        for (int w = threadIdx.x; w < Wo; w += 32) dst[w] = __ldg(src + w);
    } else {
        if (row >= NC * H) return;
        const int n = row / H, h = row % H;
        const bool hin = h >= pad && h < Ho + pad;
        const float* src = in + ((size_t)n * Ho + h - pad) * Wo - pad;
        float* dst = out + (size_t)row * W;
        for (int w = threadIdx.x; w < W; w += 32) dst[w] = (hin && w >= pad && w < Wo + pad) ? __ldg(src + w) : 0.f;
    }
}

// sphere_lat_scale_cuda.cu:31-38,61-68
template <int V>
__global__ void sphere_lat_scale_kernel(const float* __restrict__ in, const float* __restrict__ weight, float* __restrict__ out,
                                        size_t total, int H, int W, int hp) {
    const size_t nv = total / V;
    const int wv = W / V;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
        const int band = (int)((i / wv) % H) / hp;
        const float s = __ldg(weight + band);
        if (V == 4) {
            float4 v = reinterpret_cast<const float4*>(in)[i];
            v.x *= s; v.y *= s; v.z *= s; v.w *= s;
            reinterpret_cast<float4*>(out)[i] = v;
        } else out[i] = in[i] * s;
    }
}

// dtow_cuda.cu:38-75 (and the backward kernels :105-142, which are the opposite direction's forward)
// one warp per row of the WIDTH-side tensor (C/s^2, H*s, W*s); depth side (C, H, W).  d2w: contiguous stores, the
// loads interleave s channel rows (each contiguous); w2d: contiguous loads, s interleaved store rows.
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) dtow_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int C, int H,
                                                               int W, int stride, int d2w) {
    const int p2 = stride * stride;
    const int Co = C / p2, Ho = H * stride, Wo = W * stride;
    const int row = blockIdx.x * ROWS_PER_CTA + threadIdx.y;
    if (row >= N * Co * Ho) return;
    const int ho = row % Ho, pc = (row / Ho) % Co, n = row / (Ho * Co);
    const int c0 = pc * p2 + (ho % stride) * stride;                // depth channel of column phase 0
    const size_t drow = (((size_t)n * C + c0) * H + ho / stride) * W;  // + phase * H * W + wo / stride
    const size_t wrow = (size_t)row * Wo;
    const size_t cs = (size_t)H * W;
    if (stride == 2) {
        for (int wo = threadIdx.x; wo < Wo; wo += 32) {
            const size_t dj = drow + (wo & 1) * cs + (wo >> 1);
            if (d2w) out[wrow + wo] = __ldg(in + dj); else out[dj] = __ldg(in + wrow + wo);
        }
    } else {
        for (int wo = threadIdx.x; wo < Wo; wo += 32) {
            const size_t dj = drow + (wo % stride) * cs + wo / stride;
            if (d2w) out[wrow + wo] = __ldg(in + dj); else out[dj] = __ldg(in + wrow + wo);
        }
    }
}

}  // namespace lic360
using namespace lic360;

#define S_(s) as_stream(s)

extern "C" int lic360_scale(const float* in_dev, float* out_dev, size_t n, float bias, float scale, void* stream) {
    if (n == 0) return LIC360_OK;
    if (n % 4 == 0 && aligned16(in_dev) && aligned16(out_dev))
        scale_kernel<4><<<stream_grid(n / 4, 256), 256, 0, S_(stream)>>>(in_dev, out_dev, n, bias, scale);
    else
        scale_kernel<1><<<stream_grid(n, 256), 256, 0, S_(stream)>>>(in_dev, out_dev, n, bias, scale);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_imp_map_forward(const float* x_dev, const float* imp_dev, float* out_dev, float* mask_dev, int N,
                                      int C, int H, int W, int levels, void* stream) {
    LIC360_CHECK_ARG(levels > 0 && C % levels == 0, "channels must be a multiple of levels (imp_map_cuda.cu:12)");
    const int HW = H * W, cpl = C / levels;
    const size_t total = (size_t)N * C * HW;
    if (total == 0) return LIC360_OK;
    const bool v4 = HW % 4 == 0 && aligned16(x_dev) && aligned16(imp_dev) && aligned16(out_dev) && (!mask_dev || aligned16(mask_dev));
    if (v4) imp_mask_kernel<4, 0><<<stream_grid(total / 4, 256), 256, 0, S_(stream)>>>(x_dev, imp_dev, out_dev, mask_dev, N, C, HW, levels, cpl);
    else imp_mask_kernel<1, 0><<<stream_grid(total, 256), 256, 0, S_(stream)>>>(x_dev, imp_dev, out_dev, mask_dev, N, C, HW, levels, cpl);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_imp2mask(const float* in_dev, float* out_dev, int N, int C, int H, int W, int levels, void* stream) {
    LIC360_CHECK_ARG(levels > 0 && C % levels == 0, "channels must be a multiple of levels");
    const int HW = H * W, cpn = C / levels;
    const size_t total = (size_t)N * C * HW;
    if (total == 0) return LIC360_OK;
    const bool v4 = HW % 4 == 0 && aligned16(in_dev) && aligned16(out_dev);
    if (v4) imp_mask_kernel<4, 1><<<stream_grid(total / 4, 256), 256, 0, S_(stream)>>>(nullptr, in_dev, out_dev, nullptr, N, C, HW, levels, cpn);
    else imp_mask_kernel<1, 1><<<stream_grid(total, 256), 256, 0, S_(stream)>>>(nullptr, in_dev, out_dev, nullptr, N, C, HW, levels, cpn);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_imp_map_init(float* constrain_dev, float* alpha_t_dev, int N, int H, float alpha, float rt, float sc,
                                   float sw, void* stream) {
    LIC360_CHECK_ARG(H > 0 && H <= 8192 && N > 0, "bad shape");
    imp_map_init_kernel<<<1, 256, H * sizeof(float), S_(stream)>>>(constrain_dev, alpha_t_dev, N, H, alpha, rt, sc, sw);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_imp_map_backward(const float* top_diff_dev, const float* imp_dev, const float* sphere_constrain_dev,
                                       const float* alpha_t_dev, float* data_diff_dev, float* imp_diff_dev, int N, int C,
                                       int H, int W, int levels, float gamma, int imp_kernel, void* stream) {
    LIC360_CHECK_ARG(levels > 0 && C % levels == 0, "channels must be a multiple of levels");
    const int HW = H * W, cpl = C / levels;
    const size_t total = (size_t)N * C * HW;
    if (total == 0) return LIC360_OK;
    imp_map_bwd_data_kernel<<<stream_grid(total, 256), 256, 0, S_(stream)>>>(top_diff_dev, imp_dev, data_diff_dev, N, C, HW, levels, cpl);
    LAUNCH_CHECK();
    const int version = imp_kernel == 1 ? 1 : imp_kernel == 2 ? 2 : imp_kernel == 3 ? 3 : 0;  // imp_map_cuda.cu:264-293
    imp_map_bwd_imp_kernel<<<stream_grid((size_t)N * HW, 128), 128, 0, S_(stream)>>>(top_diff_dev, imp_dev, sphere_constrain_dev, alpha_t_dev,
                                                                                   imp_diff_dev, N, C, H, W, levels, cpl, gamma, version);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_quant_forward(const float* x_dev, const float* wb_dev, float* levels_dev, float* y_dev, float* q_dev,
                                    int32_t* qint_dev, float* count_dev, int N, int C, int H, int W, int L, void* stream) {
    LIC360_CHECK_ARG(L >= 2 && L <= QL_MAX, "bin_num must be in [2,32]");
    const int HW = H * W;
    LIC360_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(float) * C * L, S_(stream)));  // quant_cuda.cu:149
    quant_levels_kernel<<<stream_grid((size_t)C * L, 256), 256, 0, S_(stream)>>>(wb_dev, levels_dev, C * L, L);
    LAUNCH_CHECK();
    if ((size_t)N * C * HW == 0) return LIC360_OK;
    const bool v4 = HW % 4 == 0 && aligned16(x_dev) && aligned16(y_dev) && aligned16(qint_dev) && (!q_dev || aligned16(q_dev));
    const int chunk_elems = 256 * 4 * 4;  // 4 float4 per thread per chunk
    const int cpp = (HW + chunk_elems - 1) / chunk_elems;
    const int grid = stream_grid((size_t)N * C * cpp, 1);
    if (v4 && L == 8) {
        const int ppp = (HW + Q8_PIECE - 1) / Q8_PIECE;
        quant_fwd8_kernel<<<stream_grid((size_t)N * C * ppp, 8), 256, 0, S_(stream)>>>(x_dev, levels_dev, y_dev, q_dev, qint_dev, count_dev, N * C, C, HW, ppp);
    } else if (v4) quant_fwd_kernel<4><<<grid, 256, 0, S_(stream)>>>(x_dev, levels_dev, y_dev, q_dev, qint_dev, count_dev, N * C, C, HW, L, cpp, chunk_elems);
    else quant_fwd_kernel<1><<<grid, 256, 0, S_(stream)>>>(x_dev, levels_dev, y_dev, q_dev, qint_dev, count_dev, N * C, C, HW, L, cpp, chunk_elems);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_quant_update_weight(float* wb_dev, float* ncount_dev, int C, int L, float decay, void* stream) {
    LIC360_CHECK_ARG(L >= 3, "bin_num must be >= 3");
    quant_check_weight_kernel<<<stream_grid(C, 128), 128, 0, S_(stream)>>>(wb_dev, ncount_dev, C, L);
    LAUNCH_CHECK();
    return lic360_scale(ncount_dev, ncount_dev, (size_t)C * L, 0.f, decay, stream);  // ml_quant_scale, quant_cuda.cu:110-115
}

extern "C" int lic360_quant_backward(const float* top_diff0_dev, const float* top_diff1_dev, const float* x_dev,
                                     const float* y_dev, const int32_t* qint_dev, const float* levels_dev,
                                     float* bottom_diff_dev, float* weight_diff_dev, int N, int C, int H, int W, int L,
                                     float top_alpha, void* stream) {
    LIC360_CHECK_ARG(L >= 2 && L <= QL_MAX, "bin_num must be in [2,32]");
    const int HW = H * W;
    const size_t total = (size_t)N * C * HW;
    if (total == 0) return LIC360_OK;
    quant_bwd_weight_kernel<<<C, 256, 0, S_(stream)>>>(x_dev, y_dev, qint_dev, levels_dev, weight_diff_dev, N, C, HW, L);
    LAUNCH_CHECK();
    quant_bwd_data_kernel<<<stream_grid(total, 256), 256, 0, S_(stream)>>>(top_diff0_dev, top_diff1_dev, x_dev, y_dev, qint_dev,
                                                                         levels_dev, bottom_diff_dev, total, C, HW, L, top_alpha);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_dquant_forward(const float* q_dev, const float* mask_dev, const float* wb_dev, float* cum_dev,
                                     float* y_dev, int N, int C, int H, int W, int L, void* stream) {
    const int HW = H * W;
    dquant_levels_kernel<<<stream_grid(C, 128), 128, 0, S_(stream)>>>(wb_dev, cum_dev, C, L);
    LAUNCH_CHECK();
    const size_t total = (size_t)N * C * HW;
    if (total == 0) return LIC360_OK;
    const bool v4 = HW % 4 == 0 && aligned16(q_dev) && aligned16(mask_dev) && aligned16(y_dev);
    if (v4) dquant_fwd_kernel<4><<<stream_grid(total / 4, 256), 256, 0, S_(stream)>>>(q_dev, mask_dev, cum_dev, y_dev, N * C, C, HW, L);
    else dquant_fwd_kernel<1><<<stream_grid(total, 256), 256, 0, S_(stream)>>>(q_dev, mask_dev, cum_dev, y_dev, N * C, C, HW, L);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_context_reshape(const float* in_dev, float* out_dev, int N, int C, int H, int W, int G, int backward,
                                      void* stream) {
    LIC360_CHECK_ARG(G > 0 && C % G == 0, "channels must be a multiple of ngroup");
    const size_t total = (size_t)N * G * H * W;
    if (total == 0) return LIC360_OK;
    context_reshape_kernel<<<stream_grid(total, 256), 256, 0, S_(stream)>>>(in_dev, out_dev, N, G, C / G, H * W, backward);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_contex_shift(const float* in_dev, float* out_dev, int N, int C, int H, int W, int cpn, int mode,
                                   int zero_fill, void* stream) {
    LIC360_CHECK_ARG(cpn > 0 && C % cpn == 0 && H > 0, "bad shape");
    const int Hs = H + W + C / cpn - 2;
    if (zero_fill && mode == 0)
        LIC360_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(float) * (size_t)N * C * Hs * W, S_(stream)));
    const size_t total = (size_t)N * C * H * W;
    if (total == 0) return LIC360_OK;
    contex_shift_kernel<<<stream_grid(total, 256), 256, 0, S_(stream)>>>(in_dev, out_dev, N * C, C, H, W, Hs, cpn, mode);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_mask_constrain(float* w_dev, int Cout, int Cin, int ksize, int G, int constrain, void* stream) {
    LIC360_CHECK_ARG(G > 0 && Cin % G == 0 && Cout % G == 0, "channels must be multiples of ngroup");
    const int total = Cout * Cin * ksize * ksize;
    if (total == 0) return LIC360_OK;
    mask_constrain_kernel<<<stream_grid(total, 256), 256, 0, S_(stream)>>>(w_dev, total, Cin, ksize, Cin / G, Cout / G, constrain);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_sphere_pad(const float* in_dev, float* out_dev, int NC, int H, int W, int pad, void* stream) {
    const size_t total = (size_t)NC * (H + 2 * pad) * (W + 2 * pad);
    if (total == 0) return LIC360_OK;
    LIC360_CHECK_ARG((size_t)NC * (H + 2 * pad) < 2000000000u, "tensor too large");
    const int rows = NC * (H + 2 * pad);
    sphere_pad_kernel<<<(rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA, dim3(32, ROWS_PER_CTA), 0, S_(stream)>>>(in_dev, out_dev, NC, H, W, pad);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_sphere_pad_inplace(float* data_dev, int NC, int Hp, int Wp, int pad, void* stream) {
    LIC360_CHECK_ARG(Hp > 2 * pad && Wp > 2 * pad && pad > 0, "padded tensor smaller than its border");
    const size_t total = (size_t)NC * ((size_t)Hp * Wp - (size_t)(Hp - 2 * pad) * (Wp - 2 * pad));
    sphere_border_kernel<0><<<stream_grid(total, 256), 256, 0, S_(stream)>>>(data_dev, NC, Hp, Wp, pad);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_sphere_pad_backward(float* bottom_dev, float* top_dev, int NC, int H, int W, int pad, int inplace,
                                          void* stream) {
    const size_t total = (size_t)NC * H * W;
    if (total == 0) return LIC360_OK;
    sphere_pad_bwd_kernel<<<stream_grid(total, 256), 256, 0, S_(stream)>>>(bottom_dev, top_dev, NC, H, W, pad, inplace);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_sphere_trim(float* data_dev, int NC, int H, int W, int pad, void* stream) {
    LIC360_CHECK_ARG(H > 2 * pad && W > 2 * pad && pad > 0, "tensor smaller than its border");
    const size_t total = (size_t)NC * ((size_t)H * W - (size_t)(H - 2 * pad) * (W - 2 * pad));
    sphere_border_kernel<1><<<stream_grid(total, 256), 256, 0, S_(stream)>>>(data_dev, NC, H, W, pad);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_sphere_cut_edge(const float* in_dev, float* out_dev, int NC, int H, int W, int pad, int backward,
                                      void* stream) {
    LIC360_CHECK_ARG(H > 2 * pad && W > 2 * pad, "tensor smaller than its border");
    const size_t total = backward ? (size_t)NC * H * W : (size_t)NC * (H - 2 * pad) * (W - 2 * pad);
    if (total == 0) return LIC360_OK;
    LIC360_CHECK_ARG((size_t)NC * H < 2000000000u, "tensor too large");
    const int rows = backward ? NC * H : NC * (H - 2 * pad);
    sphere_cut_edge_kernel<<<(rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA, dim3(32, ROWS_PER_CTA), 0, S_(stream)>>>(in_dev, out_dev, NC, H, W, pad, backward);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_sphere_lat_scale(const float* in_dev, const float* weight_dev, float* out_dev, int NC, int H, int W,
                                       int npart, void* stream) {
    LIC360_CHECK_ARG(npart > 0 && H % npart == 0, "height must be a multiple of npart (sphere_lat_scale_cuda.cu:13)");
    const size_t total = (size_t)NC * H * W;
    if (total == 0) return LIC360_OK;
    if (W % 4 == 0 && aligned16(in_dev) && aligned16(out_dev))
        sphere_lat_scale_kernel<4><<<stream_grid(total / 4, 256), 256, 0, S_(stream)>>>(in_dev, weight_dev, out_dev, total, H, W, H / npart);
    else
        sphere_lat_scale_kernel<1><<<stream_grid(total, 256), 256, 0, S_(stream)>>>(in_dev, weight_dev, out_dev, total, H, W, H / npart);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_dtow(const float* in_dev, float* out_dev, int N, int C, int H, int W, int stride, int d2w, void* stream) {
    // d2w: in (N,C,H,W) -> out (N,C/s^2,H*s,W*s); else in (N,C,H,W) -> out (N,C*s^2,H/s,W/s)
    LIC360_CHECK_ARG(stride > 0, "bad stride");
    int Cd, Hd, Wd;  // depth-side dims
    if (d2w) { LIC360_CHECK_ARG(C % (stride * stride) == 0, "channels not divisible by stride^2"); Cd = C; Hd = H; Wd = W; }
    else { LIC360_CHECK_ARG(H % stride == 0 && W % stride == 0, "size not divisible by stride"); Cd = C * stride * stride; Hd = H / stride; Wd = W / stride; }
    const size_t total = (size_t)N * Cd * Hd * Wd;
    if (total == 0) return LIC360_OK;
    LIC360_CHECK_ARG(total / ((size_t)Wd * stride) < 2000000000u, "tensor too large");
    const int rows = N * (Cd / (stride * stride)) * Hd * stride;
    dtow_kernel<<<(rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA, dim3(32, ROWS_PER_CTA), 0, S_(stream)>>>(in_dev, out_dev, N, Cd, Hd, Wd, stride, d2w);
    LAUNCH_CHECK();
    return LIC360_OK;
}
