// Wavefront-slab gather/scatter ops: TileExtract(.Batch), TileInput, TileAdd.
// Replace /root/reference/extension/tile_extract_cuda.cu:31-151, tile_input_cuda.cu:27-76, tile_add_cuda.cu:22-61.
// All are indexing ops (bit-exact tier). Threads are mapped with the slab position fastest so that the plan
// lookups are coalesced; the diagonal gathers themselves are stride-(W-1) by construction of the layout.
#include "common.cuh"

namespace lic360 {

// out[(n*L + l)*cpn + ci] = x[n, tc*cpn+ci, th, tw]   (tile_extract_cuda.cu:31-45)
// batch: out[(n/B)*plane + ((n%B)*L + l)*cpn + ci]     (tile_extract_cuda.cu:101-118)
__global__ void tile_extract_kernel(const float* __restrict__ x, float* __restrict__ out, const int32_t* __restrict__ idx,
                                    int start, int L, int N, int C, int H, int W, int cpn, int psum, int B,
                                    size_t plane) {
    const int HW = H * W;
    const int total = N * L * cpn;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int l = i % L, ci = (i / L) % cpn, n = i / (L * cpn);
        const int th = __ldg(idx + start + l), tw = __ldg(idx + start + l + HW);
        const int tc = psum - th - tw;
        const float v = x[(((size_t)n * C + tc * cpn + ci) * H + th) * W + tw];
        const size_t dst = B > 0 ? (size_t)(n / B) * plane + ((size_t)(n % B) * L + l) * cpn + ci
                                 : ((size_t)n * L + l) * cpn + ci;
        out[dst] = v;
    }
}

// frame[r, n, tc, th, tw] = scale*s + bias for r < rep   (tile_input_cuda.cu:27-43)
__global__ void tile_input_kernel(const float* __restrict__ in, float* __restrict__ frame, const int32_t* __restrict__ idx,
                                  int start, int L, int N, int G, int H, int W, int psum, float bias, float scale,
                                  int rep, size_t stride_out) {
    const int HW = H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * L; i += gridDim.x * blockDim.x) {
        const int l = i % L, n = i / L;
        const int th = __ldg(idx + start + l), tw = __ldg(idx + start + l + HW);
        const int tc = psum - th - tw;
        const float v = fmaf(scale, in[i], bias);
        const size_t p = (((size_t)n * G + tc) * H + th) * W + tw;
        for (int r = 0; r < rep; r++) frame[p + r * stride_out] = v;
    }
}

// y[n, tc*cpg+og, th, tw] += x[same]   (tile_add_cuda.cu:22-38)
__global__ void tile_add_kernel(float* __restrict__ y, const float* __restrict__ x, const int32_t* __restrict__ idx,
                                int start, int L, int N, int C, int H, int W, int cpg, int G, int psum) {
    const int HW = H * W;
    const int total = N * L * cpg;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int l = i % L, og = (i / L) % cpg, n = i / (L * cpg);
        const int th = __ldg(idx + start + l), tw = __ldg(idx + start + l + HW);
        const int tc = psum - th - tw;
        if (tc < 0 || tc >= G) continue;
        const size_t p = (((size_t)n * C + tc * cpg + og) * H + th) * W + tw;
        y[p] = y[p] + x[p];
    }
}

}  // namespace lic360
using namespace lic360;

static int extract_common(const float* x_dev, float* out_dev, int N, int C, int H, int W, int G, int label, int batch,
                          const int32_t* idx_dev, const int32_t* plan_host, int psum, int* count_host, void* stream) {
    const int cpn = C / G, mod = H + W + G - 2;
    *count_host = 0;
    if (label) {
        if (psum >= mod) return LIC360_OK;  // tile_extract_cuda.cu:63,135
    } else {
        if (psum == 0) {  // tile_extract_cuda.cu:78-80
            LIC360_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(float) * (size_t)N * cpn * H * W, as_stream(stream)));
            return LIC360_OK;
        }
        if (psum > mod) return LIC360_OK;
        psum -= 1;  // tile_extract_cuda.cu:82
    }
    int start, len;
    slab_of(plan_host, H, W, G, psum, &start, &len);
    const int B = batch ? N / 3 : 0;
    *count_host = batch ? B * len : N * len;
    const size_t total = (size_t)N * len * cpn;
    if (total == 0) return LIC360_OK;
    tile_extract_kernel<<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(
        x_dev, out_dev, idx_dev, start, len, N, C, H, W, cpn, psum, B, (size_t)cpn * H * W * (batch ? B : 0));
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_tile_extract(const float* x_dev, float* out_dev, int N, int C, int H, int W, int G, int label,
                                   const int32_t* idx_dev, const int32_t* plan_host, int psum, int* count_host,
                                   void* stream) {
    LIC360_CHECK_ARG(G > 0 && C % G == 0 && count_host && plan_host, "bad arguments");
    return extract_common(x_dev, out_dev, N, C, H, W, G, label, 0, idx_dev, plan_host, psum, count_host, stream);
}

extern "C" int lic360_tile_extract_batch(const float* x_dev, float* out_dev, int N, int C, int H, int W, int G,
                                         const int32_t* idx_dev, const int32_t* plan_host, int psum, int* count_host,
                                         void* stream) {
    LIC360_CHECK_ARG(G > 0 && C % G == 0 && N % 3 == 0 && count_host && plan_host, "batch extract needs N = 3*B");
    return extract_common(x_dev, out_dev, N, C, H, W, G, 1, 1, idx_dev, plan_host, psum, count_host, stream);
}

extern "C" int lic360_tile_input(const float* in_dev, float* frame_dev, int N, int G, int H, int W, float bias,
                                 float scale, int rep, const int32_t* idx_dev, const int32_t* plan_host, int psum,
                                 void* stream) {
    LIC360_CHECK_ARG(G > 0 && rep > 0 && plan_host, "bad arguments");
    const int mod = H + W + G - 2;
    const size_t stride_out = (size_t)N * G * H * W;
    if (psum == 0) {  // tile_input_cuda.cu:57-59
        LIC360_CUDA(cudaMemsetAsync(frame_dev, 0, sizeof(float) * rep * stride_out, as_stream(stream)));
        return LIC360_OK;
    }
    if (psum > mod) return LIC360_OK;
    psum -= 1;
    int start, len;
    slab_of(plan_host, H, W, G, psum, &start, &len);
    if (len == 0) return LIC360_OK;
    tile_input_kernel<<<stream_grid((size_t)N * len, 256), 256, 0, as_stream(stream)>>>(
        in_dev, frame_dev, idx_dev, start, len, N, G, H, W, psum, bias, scale, rep, stride_out);
    LAUNCH_CHECK();
    return LIC360_OK;
}

extern "C" int lic360_tile_add(float* y_dev, const float* x_dev, int N, int C, int H, int W, int G,
                               const int32_t* idx_dev, const int32_t* plan_host, int psum, void* stream) {
    LIC360_CHECK_ARG(G > 0 && C % G == 0 && plan_host, "bad arguments");
    int start, len;
    if (psum > H + W + G - 3) return LIC360_OK;  // beyond the last wavefront the reference reads plan out of range
    slab_of(plan_host, H, W, G, psum, &start, &len);
    const int cpg = C / G;
    const size_t total = (size_t)N * len * cpg;
    if (total == 0) return LIC360_OK;
    tile_add_kernel<<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(y_dev, x_dev, idx_dev, start, len, N, C, H, W,
                                                                          cpg, G, psum);
    LAUNCH_CHECK();
    return LIC360_OK;
}
