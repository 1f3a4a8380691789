// error string, launch counter, version; host-only index plan (CodeContexOp).
#include <stdarg.h>
#include "common.cuh"

namespace lic360 {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace lic360

extern "C" const char* lic360_last_error(void) { return lic360::g_err; }
extern "C" int lic360_version(void) { return 100; }
extern "C" long long lic360_launch_count(void) { return lic360::g_launches.load(); }

// Diagonal-major index plan. Replaces code_contex_opt::reshape (code_contex_cuda.cu:11-32).
extern "C" int lic360_code_contex(int H, int W, int32_t* idx_host, int32_t* plan_host) {
    LIC360_CHECK_ARG(H > 0 && W > 0 && idx_host && plan_host, "bad arguments");
    int32_t* rows = idx_host;
    int32_t* cols = idx_host + (size_t)H * W;
    int n = 0;
    for (int d = 0; d < H + W - 1; d++) {
        plan_host[d] = n;
        const int h_first = d < W ? 0 : d - W + 1;
        const int h_last = d < H ? d : H - 1;
        for (int h = h_first; h <= h_last; h++, n++) {
            rows[n] = h;
            cols[n] = d - h;
        }
    }
    plan_host[H + W - 1] = n;
    return LIC360_OK;
}

extern "C" int lic360_slab(const int32_t* plan_host, int H, int W, int G, int psum, int* start, int* len) {
    LIC360_CHECK_ARG(plan_host && start && len && psum >= 0, "bad arguments");
    lic360::slab_of(plan_host, H, W, G, psum, start, len);
    return LIC360_OK;
}
