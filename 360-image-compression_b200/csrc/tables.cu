// GMM parameters -> per-symbol integer CDF tables (code stream), softmax logits -> CDF tables (importance
// stream), and the training-form GMM negative log-likelihood.
// Replace /root/reference/extension/entropy_gmm_table_cuda.cu:29-191, entropy_table_cuda.cu:24-96,
// entropy_gmm_cuda.cu:36-124.
//
// The CDF bins are in the bit-exact tier: every float/double promotion of the reference expressions is kept
// literally (see the comments at each line) and the same libdevice expf/erff/IEEE division are used
// (this file must NOT be compiled with --use_fast_math; the reference is not, setup.py:9-16).
// Differences in structure: the reference runs 4 launches per table (softmax, clamp, bins, fix-up) over global
// memory; here one thread owns one symbol row, keeps everything in registers, and the rows of a CTA are
// staged in shared memory so that global stores are fully coalesced.
#include "common.cuh"
#include "tables_dev.cuh"

namespace lic360 {

constexpr int TBL_THREADS = 128;

// One thread per symbol. weight/delta are rewritten in place (reference behaviour, :29-57).
__global__ void __launch_bounds__(TBL_THREADS) gmm_table_kernel(float* __restrict__ weight, float* __restrict__ delta,
                                                                const float* __restrict__ mean, float* __restrict__ out,
                                                                int rows, int ng, int nstep, float bias, float total,
                                                                float beta, float s2) {
    extern __shared__ float srow[];  // [TBL_THREADS][nt]
    const int nt = nstep + 1;
    const int row0 = blockIdx.x * TBL_THREADS;
    const int r = row0 + threadIdx.x;
    float* o = srow + threadIdx.x * nt;
    if (r < rows) {
        float wv[16], dv[16], mv[16];
        for (int i = 0; i < ng; i++) {
            wv[i] = weight[(size_t)r * ng + i];
            dv[i] = delta[(size_t)r * ng + i];
            mv[i] = mean[(size_t)r * ng + i];
        }
        gmm_row(wv, dv, mv, o, ng, nstep, bias, total, beta, s2);
        for (int i = 0; i < ng; i++) {  // softmax / clamp written back in place, entropy_gmm_table_cuda.cu:29-57
            weight[(size_t)r * ng + i] = wv[i];
            delta[(size_t)r * ng + i] = dv[i];
        }
    }
    __syncthreads();
    const int nrow = min(TBL_THREADS, rows - row0);
    float* dst = out + (size_t)row0 * nt;
    for (int e = threadIdx.x; e < nrow * nt; e += TBL_THREADS) dst[e] = srow[e];
}

// softmax of w logits -> cumulative integer table, entropy_table_cuda.cu:24-50, then fix-up :53-76
__global__ void __launch_bounds__(TBL_THREADS) entropy_table_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                    int rows, int w, float total) {
    extern __shared__ float srow[];  // [TBL_THREADS][w+1] (an odd row stride was measured slower: the index arithmetic of the copy-out)
    const int nt = w + 1, ns = nt;
    const int row0 = blockIdx.x * TBL_THREADS;
    const int nrow = min(TBL_THREADS, rows - row0);
    // coalesced load of the logits of this CTA's rows into the (wider) shared rows
    for (int e = threadIdx.x; e < nrow * w; e += TBL_THREADS) srow[(e / w) * ns + 1 + e % w] = in[(size_t)row0 * w + e];
    __syncthreads();
    if ((int)threadIdx.x < nrow) {
        entropy_row(srow + threadIdx.x * ns, w, total);
    }
    __syncthreads();
    float* dst = out + (size_t)row0 * nt;
    for (int e = threadIdx.x; e < nrow * nt; e += TBL_THREADS) dst[e] = srow[e];
}

// Training NLL and its four cached gradients, entropy_gmm_cuda.cu:36-68 (promotion order kept).
__global__ void entropy_gmm_fwd_kernel(const float* __restrict__ bw, const float* __restrict__ bd,
                                       const float* __restrict__ bm, const float* __restrict__ label,
                                       float* __restrict__ wd, float* __restrict__ dd, float* __restrict__ md,
                                       float* __restrict__ ld, float* __restrict__ loss, int S, int ng) {
    for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < S; index += gridDim.x * blockDim.x) {
        float s2 = 1. / sqrt(float(2.0));
        float sp2 = 1. / sqrt(2. * acos(-1.0));
        float sum_p = 0;
        float lds = 0;
        const float lab = label[index];
        for (int i = 0; i < ng; i++) {
            const size_t q = (size_t)index * ng + i;
            const float w = bw[q];
            float xa = lab - 0.5 - bm[q];
            float xb = lab + 0.5 - bm[q];
            float id = 1. / bd[q];
            float fa = 0.5 + 0.5 * erff(xa * id * s2);
            float fb = 0.5 + 0.5 * erff(xb * id * s2);
            float p = fb - fa;
            sum_p = sum_p + w * p;
            float ga = sp2 * id * exp(-0.5 * xa * xa * id * id);
            float gb = sp2 * id * exp(-0.5 * xb * xb * id * id);
            lds += (gb - ga) * w;
            dd[q] = id * (-xb * gb + xa * ga) * w;
            md[q] = (ga - gb) * w;
            wd[q] = p;
        }
        loss[index] = -log(sum_p + 0.0000001);
        float ip = -1. / (sum_p + 0.0000001);
        ld[index] = lds * ip;
        for (int i = 0; i < ng; i++) {
            const size_t q = (size_t)index * ng + i;
            dd[q] *= ip; md[q] *= ip; wd[q] *= ip;
        }
    }
}

// entropy_gmm_cuda.cu:94-106
__global__ void entropy_gmm_bwd_kernel(float* __restrict__ wd, float* __restrict__ dd, float* __restrict__ md,
                                       float* __restrict__ ld, const float* __restrict__ top, int S, int ng) {
    const int total = S * ng;
    for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < total; index += gridDim.x * blockDim.x) {
        const int pn = index / ng, pg = index % ng;
        const float t = top[pn];
        if (pg == 0) ld[pn] *= t;
        wd[index] *= t; dd[index] *= t; md[index] *= t;
    }
}

}  // namespace lic360
using namespace lic360;

extern "C" int lic360_gmm_table(float* weight_dev, float* delta_dev, const float* mean_dev, float* out_dev, int rows,
                                int ng, int nstep, float bias, int total_region, float beta, void* stream) {
    LIC360_CHECK_ARG(ng >= 1 && ng <= 16, "the number of Gaussians must be in [1,16] (entropy_gmm_table_cuda.cu:13)");
    LIC360_CHECK_ARG(nstep >= 1 && nstep <= 95, "nstep out of range");
    if (rows <= 0) return LIC360_OK;
    const float s2 = 1. / sqrt(2.0);
    const int grid = (rows + TBL_THREADS - 1) / TBL_THREADS;
    gmm_table_kernel<<<grid, TBL_THREADS, TBL_THREADS * (nstep + 1) * sizeof(float), as_stream(stream)>>>(
        weight_dev, delta_dev, mean_dev, out_dev, rows, ng, nstep, bias, (float)total_region, beta, s2);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_entropy_table(const float* in_dev, float* out_dev, int rows, int nstep, int total_region,
                                    void* stream) {
    LIC360_CHECK_ARG(nstep >= 1 && nstep <= 64, "nstep must be <= 64 (entropy_table_cuda.cu:13)");
    if (rows <= 0) return LIC360_OK;
    const int grid = (rows + TBL_THREADS - 1) / TBL_THREADS;
    entropy_table_kernel<<<grid, TBL_THREADS, TBL_THREADS * (nstep + 1) * sizeof(float), as_stream(stream)>>>(
        in_dev, out_dev, rows, nstep, (float)total_region);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_entropy_gmm_forward(const float* weight_dev, const float* delta_dev, const float* mean_dev,
                                          const float* label_dev, float* wdiff_dev, float* ddiff_dev, float* mdiff_dev,
                                          float* ldiff_dev, float* loss_dev, int S, int ng, void* stream) {
    if (S <= 0) return LIC360_OK;
    entropy_gmm_fwd_kernel<<<stream_grid(S, 256), 256, 0, as_stream(stream)>>>(weight_dev, delta_dev, mean_dev, label_dev,
                                                                             wdiff_dev, ddiff_dev, mdiff_dev, ldiff_dev,
                                                                             loss_dev, S, ng);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_entropy_gmm_backward(float* wdiff_dev, float* ddiff_dev, float* mdiff_dev, float* ldiff_dev,
                                           const float* top_diff_dev, int S, int ng, void* stream) {
    if (S <= 0) return LIC360_OK;
    entropy_gmm_bwd_kernel<<<stream_grid((size_t)S * ng, 256), 256, 0, as_stream(stream)>>>(wdiff_dev, ddiff_dev, mdiff_dev,
                                                                                          ldiff_dev, top_diff_dev, S, ng);
    LAUNCH_CHECK();
    return LIC360_OK;
}
