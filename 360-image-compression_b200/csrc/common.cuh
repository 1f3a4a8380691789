// Shared helpers for the lic360_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/lic360_b200.h"

namespace lic360 {

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs; grids of streaming kernels are sized in multiples of this

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;  // counted by LAUNCH_CHECK (and by graph replays)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// grid for a grid-stride streaming kernel: enough CTAs to cover `work` items with `per_cta` items each,
// rounded up to a multiple of the SM count and capped at `waves` resident waves.
inline int stream_grid(size_t work, int per_cta, int ctas_per_sm = 8) {
    size_t need = (work + per_cta - 1) / per_cta;
    size_t cap = (size_t)kNumSM * ctas_per_sm;
    if (need >= cap) return (int)cap;
    size_t g = ((need + kNumSM - 1) / kNumSM) * kNumSM;
    if (g == 0) g = kNumSM;
    return (int)g;
}

#define LIC360_CHECK_ARG(cond, msg)                        \
    do {                                                   \
        if (!(cond)) {                                     \
            lic360::set_error("%s: %s", __func__, msg);    \
            return LIC360_ERR_ARG;                         \
        }                                                  \
    } while (0)

#define LIC360_CUDA(call)                                                               \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            lic360::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e_)); \
            return LIC360_ERR_CUDA;                                                     \
        }                                                                               \
    } while (0)

#define LAUNCH_CHECK()                                                                        \
    do {                                                                                      \
        lic360::g_launches++;                                                                 \
        cudaError_t e_ = cudaGetLastError();                                                  \
        if (e_ != cudaSuccess) {                                                              \
            lic360::set_error("%s: kernel launch failed -> %s", __func__, cudaGetErrorString(e_)); \
            return LIC360_ERR_CUDA;                                                           \
        }                                                                                     \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: remember the largest value raised so far for
// every device, so that an op or codec created on cuda:1 after cuda:0 in the same process raises it there as well.
constexpr int kMaxDevices = 64;
struct SmemAttr {
    size_t raised[kMaxDevices];
    SmemAttr() { for (int i = 0; i < kMaxDevices; i++) raised[i] = 48 * 1024; }
    template <typename F>
    cudaError_t ensure(F func, size_t bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
        if (bytes <= raised[dev]) return cudaSuccess;
        e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess) raised[dev] = bytes;
        return e;
    }
};

// the codec entry points bind their device for the duration of the call and give the caller's current device back
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// wavefront slab of step psum (reference: cconv_dc_cuda.cu:113-117)
inline void slab_of(const int32_t* plan, int H, int W, int G, int psum, int* start, int* len) {
    int la = psum >= G ? psum - G + 1 : 0;
    int lb = psum > H + W - 2 ? H + W - 2 : psum;
    *start = plan[la];
    int l = plan[lb + 1] - plan[la];
    *len = l > 0 ? l : 0;
}

}  // namespace lic360
