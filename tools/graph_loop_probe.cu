// Probe: how fast can a host <-> GPU ping-pong step be driven?  (the decoder's step: kernels -> rows to the host -> host decodes ->
// symbols back -> next step).  Three ways to start step p + 1 once the host has answered step p:
//   A. cudaGraphLaunch of a 3-kernel graph per step (what codec.cu does)
//   B. a device-side WHILE loop (CUDA 12.4 conditional graph node): the body's first kernel spins on a mapped host flag, the host only
//      stores the flag -- no launch on the critical path
//   C. one persistent kernel that spins on the flag itself (lower bound: PCIe round trip + nothing)
// Each step's last kernel writes `ready = p` into mapped host memory; the host polls it and immediately answers `go = p + 1`.
// Prints microseconds per step.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/graph_loop_probe tools/graph_loop_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_work(int* scratch) {  // stands for scatter / chain: a little dependent work
    if (threadIdx.x == 0) atomicAdd(scratch, 1);
}
__global__ void k_ready(volatile int* ready, int* step) {  // last kernel of a step: publish the step number to the host
    *ready = *step;
    __threadfence_system();
}
__global__ void k_advance(int* step) { *step += 1; }
__global__ void k_wait_go(const volatile int* go, const int* step) {  // first kernel of a step (B): spin until the host has answered
    const int p = *step;
    unsigned long long t0 = 0, t1;
    for (unsigned spins = 1; *go < p; spins++)
        if ((spins & 0xFF) == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 2000000000ull) break;
        }
}
__global__ void k_set_cond(cudaGraphConditionalHandle h, const int* step, int nsteps) { cudaGraphSetConditional(h, *step < nsteps ? 1 : 0); }
__global__ void k_persistent(const volatile int* go, volatile int* ready, int nsteps) {
    for (int p = 0; p < nsteps; p++) {
        unsigned long long t0 = 0, t1;
        for (unsigned spins = 1; *go < p; spins++)
            if ((spins & 0xFF) == 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t0 == 0) t0 = t1;
                else if (t1 - t0 > 2000000000ull) return;
            }
        *ready = p;
        __threadfence_system();
    }
}

using clk = std::chrono::steady_clock;
static double us_since(clk::time_point t) { return std::chrono::duration<double, std::micro>(clk::now() - t).count(); }

int main() {
    const int N = 2000;
    int *go, *ready, *step, *scratch;
    CK(cudaHostAlloc(&go, 64, cudaHostAllocMapped));
    CK(cudaHostAlloc(&ready, 64, cudaHostAllocMapped));
    CK(cudaMalloc(&step, 4));
    CK(cudaMalloc(&scratch, 4));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    auto host_loop = [&](bool launch, cudaGraphExec_t g) -> double {  // answers every step; returns us per step
        auto t0 = clk::now();
        for (int p = 0; p < N; p++) {
            if (launch) cudaGraphLaunch(g, s);
            else __atomic_store_n(go, p, __ATOMIC_RELEASE);
            while (__atomic_load_n(ready, __ATOMIC_ACQUIRE) < p) {}
        }
        return us_since(t0) / N;
    };
    // ---- A: one graph launch per step: [work] -> [work] -> [ready; advance]
    {
        CK(cudaMemset(step, 0, 4));
        *ready = -1;
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        k_work<<<24, 128, 0, s>>>(scratch);
        k_work<<<24, 384, 0, s>>>(scratch);
        k_ready<<<1, 1, 0, s>>>(ready, step);
        k_advance<<<1, 1, 0, s>>>(step);
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemset(step, 0, 4));
            *ready = -1;
            const double us = host_loop(true, ge);
            CK(cudaStreamSynchronize(s));
            printf("A. cudaGraphLaunch per step (4 kernel nodes): %.2f us per step\n", us);
        }
    }
    // ---- B: device-side while loop
    {
        cudaGraph_t g;
        CK(cudaGraphCreate(&g, 0));
        cudaGraphConditionalHandle h;
        CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = h;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t cond;
        cudaError_t e = cudaGraphAddNode(&cond, g, nullptr, 0, &np);
        if (e != cudaSuccess) { printf("B. conditional node not available: %s\n", cudaGetErrorString(e)); }
        else {
            cudaGraph_t body = np.conditional.phGraph_out[0];
            CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            k_wait_go<<<1, 1, 0, s>>>(go, step);
            k_work<<<24, 128, 0, s>>>(scratch);
            k_work<<<24, 384, 0, s>>>(scratch);
            k_ready<<<1, 1, 0, s>>>(ready, step);
            k_advance<<<1, 1, 0, s>>>(step);
            k_set_cond<<<1, 1, 0, s>>>(h, step, N);
            CK(cudaStreamEndCapture(s, nullptr));
            cudaGraphExec_t ge;
            CK(cudaGraphInstantiate(&ge, g, 0));
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaMemset(step, 0, 4));
                *ready = -1;
                __atomic_store_n(go, -1, __ATOMIC_RELEASE);
                CK(cudaGraphLaunch(ge, s));
                const double us = host_loop(false, nullptr);
                CK(cudaStreamSynchronize(s));
                printf("B. device-side WHILE node, host only stores a flag (6 kernel nodes per iteration): %.2f us per step\n", us);
            }
        }
    }
    // ---- C: persistent kernel
    for (int rep = 0; rep < 2; rep++) {
        *ready = -1;
        __atomic_store_n(go, -1, __ATOMIC_RELEASE);
        k_persistent<<<1, 1, 0, s>>>(go, ready, N);
        const double us = host_loop(false, nullptr);
        CK(cudaStreamSynchronize(s));
        printf("C. persistent kernel, flag ping-pong only: %.2f us per step\n", us);
    }
    printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
