// Probe: cost of a cluster barrier round on this GPU, alone and with a producer/consumer exchange through L2.
// usage: cluster_probe   (prints microseconds per round for several variants)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(384, 1) k_sync_only(int rounds, float* buf) {
    for (int r = 0; r < rounds; r++) cg::this_cluster().sync();
    if (buf && threadIdx.x == 0 && rounds < 0) buf[0] = 1.f;
}

// every round: 25 ldcg float4 of the neighbours' data written in the previous round, 100 FMAs, one float4 store, barrier
__global__ void __launch_bounds__(384, 1) k_exchange(int rounds, float4* buf, int per_cta, int nloads, int do_sync) {
    const int nthr = gridDim.x * blockDim.x;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < rounds; r++) {
        const float4* src = buf + (size_t)(r & 1) * nthr;
        float4* dst = buf + (size_t)((r + 1) & 1) * nthr;
        float4 x[25];
#pragma unroll
        for (int i = 0; i < 25; i++) {
            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nloads) x[i] = __ldcg(src + (gid + i * 37) % nthr);
        }
#pragma unroll
        for (int i = 0; i < 25; i++) { acc.x = fmaf(x[i].x, 1.0001f, acc.x); acc.y = fmaf(x[i].y, 1.0001f, acc.y); acc.z = fmaf(x[i].z, 0.5f, acc.z); acc.w = fmaf(x[i].w, 0.5f, acc.w); }
        dst[gid] = acc;
        if (do_sync) cg::this_cluster().sync();
    }
    (void)per_cta;
}

static float run(int cluster, int ctas, int threads, int which, int rounds, float4* buf, int nloads, int do_sync) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(threads);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(e0);
        cudaError_t r = which == 0 ? cudaLaunchKernelEx(&cfg, k_sync_only, rounds, (float*)buf)
                                   : cudaLaunchKernelEx(&cfg, k_exchange, rounds, buf, 0, nloads, do_sync);
        cudaEventRecord(e1);
        if (r != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1.f; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best * 1e3f;
}

int main() {
    cudaFuncSetAttribute(k_sync_only, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k_exchange, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    float4* buf; cudaMalloc(&buf, 2 * 48 * 384 * sizeof(float4)); cudaMemset(buf, 0, 2 * 48 * 384 * sizeof(float4));
    const int R = 120;
    for (int cl : {1, 2, 4, 8, 16}) {
        const float t0 = run(cl, 3 * cl, 192, 0, 0, buf, 0, 1);
        const float t1 = run(cl, 3 * cl, 192, 0, R, buf, 0, 1);
        const float t2 = run(cl, 3 * cl, 192, 1, R, buf, 25, 1);
        const float t3 = run(cl, 3 * cl, 192, 1, R, buf, 25, 0);
        const float t4 = run(cl, 3 * cl, 192, 1, R, buf, 1, 1);
        printf("cluster %2d x3, 192 thr: empty launch %.1f us | barrier round %.3f us | 25 ldcg+fma+st+barrier %.3f us | same without barrier %.3f us | 1 ldcg+st+barrier %.3f us\n",
               cl, t0, (t1 - t0) / R, (t2 - t0) / R, (t3 - t0) / R, (t4 - t0) / R);
    }
    return 0;
}
