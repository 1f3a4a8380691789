// Probe: shared-memory load cost (cycles per warp instruction, one warp per SM and 4 warps per SM) of the access patterns the old-term
// kernels choose between: LDS.128 with one address per warp (broadcast), one per half-warp, one per quarter-warp, one per lane
// (contiguous); LDS.64 / LDS.32 broadcast and contiguous.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lds_probe tools/lds_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int BYTES, int GROUP>  // GROUP = lanes sharing one address (32 = broadcast, 1 = contiguous per lane)
__global__ void k(float* out, long long* cyc, int iters) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = lane / GROUP;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int stride = BYTES / 4;                 // floats per access
    int base = g * stride * (GROUP == 1 ? 1 : 25);  // distinct rows for distinct groups
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int idx = (base + u * stride * 32 * 0 + u * 4 * 40) & 8191 & ~(stride - 1);
            if (BYTES == 16) { float4 v = *reinterpret_cast<const float4*>(sm + idx); acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }
            if (BYTES == 8)  { float2 v = *reinterpret_cast<const float2*>(sm + idx); acc0 += v.x; acc1 += v.y; }
            if (BYTES == 4)  { acc0 += sm[idx]; }
        }
        base = (base + (int)acc0 * 0 + 4) & 4095;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

template <int BYTES, int GROUP>
void run(const char* name, float* out, long long* cyc) {
    for (int warps : {1, 4, 8}) {
        const int iters = 2000;
        k<BYTES, GROUP><<<1, 32 * warps>>>(out, cyc, iters);
        k<BYTES, GROUP><<<1, 32 * warps>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-40s warps %d: %.2f cycles per warp-load (SM total %.2f per load)\n", name, warps, (double)c / (iters * 16), (double)c / (iters * 16 * warps));
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    run<16, 32>("LDS.128 broadcast (1 address)", out, cyc);
    run<16, 16>("LDS.128 one address per half-warp", out, cyc);
    run<16, 8>("LDS.128 one address per quarter-warp", out, cyc);
    run<16, 1>("LDS.128 contiguous per lane", out, cyc);
    run<8, 32>("LDS.64 broadcast", out, cyc);
    run<8, 1>("LDS.64 contiguous per lane", out, cyc);
    run<4, 32>("LDS.32 broadcast", out, cyc);
    run<4, 1>("LDS.32 contiguous per lane", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
