// Probe: which TMA box-load variants work on this GPU/driver. Each variant runs in its own process (a fault is sticky).
// usage: tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Maps { CUtensorMap tm[12]; };
struct Pad { char bytes[1608]; };

__device__ __forceinline__ bool try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

__device__ void body(const CUtensorMap* tm, const float4* wsrc, float* out, int c0, int c1, int c2, int bulk) {
    extern __shared__ unsigned char raw[];
    const unsigned raw_s = (unsigned)__cvta_generic_to_shared(raw);
    unsigned char* base = raw + ((128u - (raw_s & 127u)) & 127u);
    float* band = reinterpret_cast<float*>(base);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + 6912);
    const unsigned band_s = (unsigned)__cvta_generic_to_shared(band);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(bars);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int bytes = 36 * 9 * 4 * 4 + (bulk ? 1600 : 0);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                     ::"r"(band_s), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
        if (bulk)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(band_s + 5184), "l"(wsrc), "r"(1600), "r"(bar) : "memory");
    }
    while (!try_wait(bar, 0)) {}
    for (int i = threadIdx.x; i < 36 * 9 * 4 + (bulk ? 400 : 0); i += blockDim.x) out[i] = band[i];
}

__global__ void k_direct(const __grid_constant__ CUtensorMap tm, const float4* wsrc, float* out, int c0, int c1, int c2, int bulk) {
    body(&tm, wsrc, out, c0, c1, c2, bulk);
}
__global__ void k_struct(const __grid_constant__ Pad pad, const __grid_constant__ Maps maps, const float4* wsrc, float* out, int l, int c0, int c1, int c2, int bulk) {
    body(&maps.tm[l], wsrc, out, c0, c1, c2, bulk);
}
__global__ void k_global(const CUtensorMap* maps, const float4* wsrc, float* out, int l, int c0, int c1, int c2, int bulk) {
    body(&maps[l], wsrc, out, c0, c1, c2, bulk);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    int HS = 64, D = 191, C = 144, c0 = 0, c1 = 0, c2 = 0, mode = 0, bulk = 0, l = 0;
    switch (v) {
        case 0: break;                                   // direct param, in-bounds
        case 1: c0 = -2; c1 = -4; break;                 // negative start
        case 2: HS = 32; D = 95; C = 1; break;           // tensor smaller than the box
        case 3: HS = 36; D = 23; C = 4; c0 = -2; c1 = -4; break;
        case 4: mode = 1; l = 5; break;                  // struct array, dynamic index
        case 5: mode = 2; l = 5; break;                  // descriptors in global memory
        case 6: bulk = 1; break;                         // + bulk copy on the same barrier
        case 7: mode = 1; l = 3; bulk = 1; c0 = 30; c1 = 185; c2 = 140; break;  // everything, box crossing the far corner
        case 100: c0 = atoi(argv[2]); c1 = atoi(argv[3]); c2 = atoi(argv[4]); break;  // explicit coordinates
    }
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("v%d: no encode fn\n", v); return 2; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const size_t nfl = (size_t)HS * D * (C < 4 ? 4 : C);
    std::vector<float> h(nfl);
    for (size_t i = 0; i < nfl; i++) h[i] = (float)(i % 1000);
    float *x, *out; float4* w;
    cudaMalloc(&x, nfl * 4); cudaMemcpy(x, h.data(), nfl * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 8192 * 4); cudaMemset(out, 0, 8192 * 4);
    cudaMalloc(&w, 1600); cudaMemset(w, 0, 1600);
    Maps maps; memset(&maps, 0, sizeof(maps));
    const cuuint64_t dims[3] = {(cuuint64_t)HS, (cuuint64_t)D, (cuuint64_t)C};
    const cuuint64_t strides[2] = {(cuuint64_t)HS * 4, (cuuint64_t)D * HS * 4};
    const cuuint32_t box[3] = {36, 9, 4}, estr[3] = {1, 1, 1};
    for (int i = 0; i < 12; i++) {
        CUresult r = enc(&maps.tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("v%d: encode failed %d\n", v, (int)r); return 3; }
    }
    const size_t smem = 128 + 6912 + 64;
    if (mode == 0) k_direct<<<1, 32, smem>>>(maps.tm[0], w, out, c0, c1, c2, bulk);
    else if (mode == 1) { Pad pad; memset(&pad, 0, sizeof(pad)); k_struct<<<1, 32, smem>>>(pad, maps, w, out, l, c0, c1, c2, bulk); }
    else { CUtensorMap* md; cudaMalloc(&md, sizeof(maps)); cudaMemcpy(md, &maps, sizeof(maps), cudaMemcpyHostToDevice); k_global<<<1, 32, smem>>>(md, w, out, l, c0, c1, c2, bulk); }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(1296);
    cudaMemcpy(o.data(), out, 1296 * 4, cudaMemcpyDeviceToHost);
    // expected value of band[c][s][i] = x[(c2+c)][c1+s][c0+i] or 0 outside
    int bad = 0;
    for (int c = 0; c < 4; c++) for (int s = 0; s < 9; s++) for (int i = 0; i < 36; i++) {
        const int cc = c2 + c, dd = c1 + s, hh = c0 + i;
        float exp = 0.f;
        if (cc >= 0 && cc < C && dd >= 0 && dd < D && hh >= 0 && hh < HS) exp = h[((size_t)cc * D + dd) * HS + hh];
        if (o[(c * 9 + s) * 36 + i] != exp) bad++;
    }
    printf("v%d: sync=%s mismatches=%d\n", v, cudaGetErrorString(e), bad);
    return e == cudaSuccess && bad == 0 ? 0 : 1;
}
