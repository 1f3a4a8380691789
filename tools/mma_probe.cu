// Probe of the legacy tensor path (mma.sync.m16n8k8 TF32, fp32 accumulate) on this GPU -- the facts the tensor-core form of
// the context conv's old terms (csrc/conv_mma.cuh) relies on:
//   1. throughput per SM (how many warps it takes to saturate, MAC/clk/SM) -> is it worth it against 128 FFMA/clk/SM
//   2. a result element depends only on its A row, its B column and its C element (not on its position in the tile,
//      not on the other rows / columns)              -> encoder tiles and decoder slab tiles can agree bit for bit
//   3. an all-zero B column (or A row) leaves C unchanged -> masked / skipped k-steps are interchangeable
//   4. rounding of the accumulation (RN or RZ) and whether fp32 operand bits below TF32 precision are ignored
//   5. ldmatrix (b16 x4) delivers the TF32 A fragment of a row-major [16][8] fp32 tile
//   6. accuracy of the split product (hi*hi + lo*hi + hi*lo + lo*lo in two accumulators) against fp64 over K = 4800
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ---------------------------------------------------------------- 1. throughput
template <int KIND, int NACC>
__global__ void k_rate(int iters, float* out) {
    float acc[NACC][4];
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; i++) a[i] = __float_as_uint(1.0f + threadIdx.x * 0.001f) & 0xffffe000u;
    for (int i = 0; i < 2; i++) b[i] = __float_as_uint(0.5f) & 0xffffe000u;
    for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) acc[j][i] = 0.f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < NACC; j++) {
            if (KIND == 0) mma_tf32(acc[j], a, b);
            else mma_bf16(acc[j], a, b);
        }
    }
    float s = 0.f;
    for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) s += acc[j][i];
    if (s == 123.456f) out[0] = s;
}

template <int KIND, int NACC>
static void rate(const char* name, int warps, int ctas_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_rate<KIND, NACC><<<sms * ctas_per_sm, 32 * warps>>>(iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double macs_per_mma = KIND == 0 ? 16. * 8 * 8 : 16. * 8 * 16;
    const double total = (double)sms * ctas_per_sm * warps * iters * NACC * macs_per_mma;
    const double tflops = 2 * total / (best * 1e-3) / 1e12;
    printf("rate %-5s warps/CTA %2d CTAs/SM %d acc/warp %d: %.3f ms, %.1f TFLOP/s dense, %.0f MAC/clk/SM at %d MHz nominal\n", name, warps,
           ctas_per_sm, NACC, best, tflops, total / sms / (best * 1e-3 * clk_khz * 1e3), clk_khz / 1000);
    cudaFree(out);
}

// ---------------------------------------------------------------- 2..5 semantics
// one warp: D = A(16x8, row-major fp32 in global) * B(8x8, [k][n]) + C(16x8), operands passed as raw bits (no masking here)
__global__ void k_one(const float* A, const float* B, const float* C, float* D, int use_ldsm) {
    __shared__ __align__(16) float sA[16 * 8];
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    for (int i = lane; i < 128; i += 32) sA[i] = A[i];
    __syncwarp();
    uint32_t a[4], b[2];
    if (use_ldsm) {
        // four 8x8 b16 matrices = four [8 rows][4 tf32]: (rows 0-7, k 0-3), (rows 8-15, k 0-3), (rows 0-7, k 4-7), (rows 8-15, k 4-7)
        // lane l supplies the row address of matrix l/8, row l%8
        const int m = lane >> 3, r = lane & 7;
        const float* p = sA + ((m & 1) * 8 + r) * 8 + (m >> 1) * 4;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
    } else {
        a[0] = __float_as_uint(sA[g * 8 + t]);
        a[1] = __float_as_uint(sA[(g + 8) * 8 + t]);
        a[2] = __float_as_uint(sA[g * 8 + t + 4]);
        a[3] = __float_as_uint(sA[(g + 8) * 8 + t + 4]);
    }
    b[0] = __float_as_uint(B[t * 8 + g]);
    b[1] = __float_as_uint(B[(t + 4) * 8 + g]);
    float d[4] = {C[g * 8 + 2 * t], C[g * 8 + 2 * t + 1], C[(g + 8) * 8 + 2 * t], C[(g + 8) * 8 + 2 * t + 1]};
    mma_tf32(d, a, b);
    D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1]; D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float frand() { return (float)((rand() / (double)RAND_MAX) * 2.0 - 1.0); }

struct Dev {
    float *A, *B, *C, *D;
    Dev() { cudaMalloc(&A, 512); cudaMalloc(&B, 256); cudaMalloc(&C, 512); cudaMalloc(&D, 512); }
    void run(const float* a, const float* b, const float* c, float* d, int ldsm = 0) {
        cudaMemcpy(A, a, 512, cudaMemcpyHostToDevice); cudaMemcpy(B, b, 256, cudaMemcpyHostToDevice); cudaMemcpy(C, c, 512, cudaMemcpyHostToDevice);
        k_one<<<1, 32>>>(A, B, C, D, ldsm);
        cudaMemcpy(d, D, 512, cudaMemcpyDeviceToHost);
    }
};

static uint32_t bits(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }

static void semantics() {
    Dev dev;
    float A[128], B[64], C[128], D[128], D2[128];
    srand(7);
    // 2a. same row vector in every A row, same column vector in every B column, same C: all 128 results identical?
    int bad = 0;
    for (int trial = 0; trial < 200; trial++) {
        float ar[8], bc[8], c0 = frand() * 4.f;
        for (int k = 0; k < 8; k++) { ar[k] = tf32_trunc(frand() * exp2f((float)(rand() % 20 - 10))); bc[k] = tf32_trunc(frand() * exp2f((float)(rand() % 20 - 10))); }
        for (int i = 0; i < 16; i++) for (int k = 0; k < 8; k++) A[i * 8 + k] = ar[k];
        for (int k = 0; k < 8; k++) for (int n = 0; n < 8; n++) B[k * 8 + n] = bc[k];
        for (int i = 0; i < 128; i++) C[i] = c0;
        dev.run(A, B, C, D);
        for (int i = 1; i < 128; i++) bad += bits(D[i]) != bits(D[0]);
        // 2b. now randomise every other row / column: element (5, 3) keeps its value?
        const float keep = D[0];
        for (int i = 0; i < 16; i++) if (i != 5) for (int k = 0; k < 8; k++) A[i * 8 + k] = tf32_trunc(frand() * 100.f);
        for (int n = 0; n < 8; n++) if (n != 3) for (int k = 0; k < 8; k++) B[k * 8 + n] = tf32_trunc(frand() * 100.f);
        for (int i = 0; i < 128; i++) if (i != 5 * 8 + 3) C[i] = frand();
        dev.run(A, B, C, D2);
        bad += bits(D2[5 * 8 + 3]) != bits(keep);
    }
    printf("position independence: %d mismatches over 200 trials (0 = a result depends only on its A row, B column and C)\n", bad);
    // 3. zero B column / zero A row leaves C unchanged
    bad = 0;
    for (int trial = 0; trial < 200; trial++) {
        for (int i = 0; i < 128; i++) { A[i] = tf32_trunc(frand() * 1000.f); C[i] = frand() * exp2f((float)(rand() % 60 - 30)); }
        for (int i = 0; i < 64; i++) B[i] = tf32_trunc(frand());
        for (int k = 0; k < 8; k++) B[k * 8 + 2] = 0.f;            // column 2
        for (int k = 0; k < 8; k++) A[7 * 8 + k] = 0.f;            // row 7
        C[3 * 8 + 2] = -0.0f;
        dev.run(A, B, C, D);
        for (int i = 0; i < 16; i++) bad += bits(D[i * 8 + 2]) != bits(C[i * 8 + 2]);
        for (int n = 0; n < 8; n++) bad += bits(D[7 * 8 + n]) != bits(C[7 * 8 + n]);
    }
    printf("zero column / zero row is a no-op on C: %d mismatches (incl. C = -0: %s)\n", bad, bad ? "see count" : "kept");
    // 4. rounding of the accumulation: C = 1, one product = 1.5 * 2^-24 (exact sum 1 + 1.5 ulp/2): RN -> 1 + 2^-23, RZ -> 1
    for (int i = 0; i < 128; i++) { A[i] = 0.f; C[i] = 1.0f; }
    for (int i = 0; i < 64; i++) B[i] = 0.f;
    for (int i = 0; i < 16; i++) A[i * 8] = exp2f(-12.f);
    for (int n = 0; n < 8; n++) B[n] = 1.5f * exp2f(-12.f);
    dev.run(A, B, C, D);
    printf("accumulate rounding: 1 + 1.5*2^-24 -> %a (%s)\n", D[0], D[0] == 1.0f ? "truncated: RZ" : "rounded up: RN-like");
    for (int i = 0; i < 128; i++) C[i] = -1.0f;
    for (int n = 0; n < 8; n++) B[n] = -1.5f * exp2f(-12.f);
    dev.run(A, B, C, D);
    printf("accumulate rounding: -1 - 1.5*2^-24 -> %a\n", D[0]);
    // eight products of 2^-26 each against C = 1: are they summed before they meet C (-> 1 + 2^-23) or dropped one by one (-> 1)?
    for (int i = 0; i < 128; i++) { A[i] = exp2f(-13.f); C[i] = 1.0f; }
    for (int i = 0; i < 64; i++) B[i] = exp2f(-13.f);
    dev.run(A, B, C, D);
    printf("8 x 2^-26 + 1 -> %a (1 + 2^-23 = %a)\n", D[0], 1.0f + exp2f(-23.f));
    // operand bits below TF32: ignored (truncated) or rounded?
    for (int i = 0; i < 128; i++) { A[i] = 0.f; C[i] = 0.f; }
    for (int i = 0; i < 64; i++) B[i] = 0.f;
    const float xa = 1.0f + exp2f(-11.f) + exp2f(-12.f) + exp2f(-13.f);  // low 13 bits: 0x1C00 > half of 2^-10
    for (int i = 0; i < 16; i++) A[i * 8] = xa;
    for (int n = 0; n < 8; n++) B[n] = 1.0f;
    dev.run(A, B, C, D);
    printf("fp32 operand bits below tf32: 1+2^-11+2^-12+2^-13 times 1 -> %a (%s)\n", D[0], D[0] == 1.0f ? "truncated" : D[0] == 1.0f + exp2f(-10.f) ? "rounded to nearest tf32" : "kept beyond tf32!");
    // does the order of the k slots matter? same multiset of products, rotated k
    bad = 0;
    for (int trial = 0; trial < 200; trial++) {
        float ar[8], bc[8];
        for (int k = 0; k < 8; k++) { ar[k] = tf32_trunc(frand() * exp2f((float)(rand() % 16 - 8))); bc[k] = tf32_trunc(frand()); }
        for (int i = 0; i < 16; i++) for (int k = 0; k < 8; k++) { A[i * 8 + k] = ar[(k + i) % 8]; }
        for (int i = 0; i < 128; i++) C[i] = 0.37f;
        // row i uses rotation i: B column n must use the same rotation to pair the same products -> use one column set per row: compare rows via separate runs
        float ref = 0.f;
        for (int rot = 0; rot < 8; rot++) {
            for (int k = 0; k < 8; k++) for (int n = 0; n < 8; n++) B[k * 8 + n] = bc[(k + rot) % 8];
            dev.run(A, B, C, D);
            if (rot == 0) ref = D[0 * 8];
            else bad += bits(D[rot * 8]) != bits(ref);
        }
    }
    printf("k-slot order: %d of 1400 rotated sums differ from the unrotated one (0 = the 8 products are summed order-independently)\n", bad);
    // 5. ldmatrix A fragment
    for (int i = 0; i < 128; i++) { A[i] = tf32_trunc(frand()); C[i] = frand(); }
    for (int i = 0; i < 64; i++) B[i] = tf32_trunc(frand());
    dev.run(A, B, C, D, 0);
    dev.run(A, B, C, D2, 1);
    bad = 0;
    for (int i = 0; i < 128; i++) bad += bits(D[i]) != bits(D2[i]);
    double maxerr = 0;
    for (int i = 0; i < 16; i++) for (int n = 0; n < 8; n++) {
        double s = C[i * 8 + n];
        for (int k = 0; k < 8; k++) s += (double)A[i * 8 + k] * B[k * 8 + n];
        maxerr = fmax(maxerr, fabs(s - D[i * 8 + n]));
    }
    printf("ldmatrix.x4 A fragment == scalar-load A fragment: %d mismatches; max |D - fp64| = %.3g\n", bad, maxerr);
}

// ---------------------------------------------------------------- 6. accuracy of the split product
// one warp, K = nk8 * 8: row i of A, column n of B random; per 8-k step four MMAs into two accumulators, every `chunk` steps
// the accumulators are folded into an fp32 sum with IEEE adds (the canonical order of conv_mma.cuh)
__global__ void k_split(const float* A, const float* B, float* D, int nk8, int chunk, int nterms) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    float P[4] = {0.f, 0.f, 0.f, 0.f};
    float ua[4] = {0.f, 0.f, 0.f, 0.f}, ub[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < nk8; s++) {
        const float* a8 = A + (size_t)s * 128;
        const float* b8 = B + (size_t)s * 64;
        const float af[4] = {a8[g * 8 + t], a8[(g + 8) * 8 + t], a8[g * 8 + t + 4], a8[(g + 8) * 8 + t + 4]};
        const float bf[2] = {b8[t * 8 + g], b8[(t + 4) * 8 + g]};
        uint32_t ah[4], al[4], bh[2], bl[2];
        for (int i = 0; i < 4; i++) { ah[i] = __float_as_uint(af[i]) & 0xffffe000u; al[i] = __float_as_uint(af[i] - __uint_as_float(ah[i])) & 0xffffe000u; }
        for (int i = 0; i < 2; i++) { bh[i] = __float_as_uint(bf[i]) & 0xffffe000u; bl[i] = __float_as_uint(bf[i] - __uint_as_float(bh[i])) & 0xffffe000u; }
        mma_tf32(ua, ah, bh);
        mma_tf32(ua, al, bh);
        mma_tf32(ub, ah, bl);
        if (nterms == 4) mma_tf32(ub, al, bl);
        if ((s + 1) % chunk == 0 || s + 1 == nk8) {
            for (int i = 0; i < 4; i++) { P[i] = P[i] + (ua[i] + ub[i]); ua[i] = 0.f; ub[i] = 0.f; }
        }
    }
    D[g * 8 + 2 * t] = P[0]; D[g * 8 + 2 * t + 1] = P[1]; D[(g + 8) * 8 + 2 * t] = P[2]; D[(g + 8) * 8 + 2 * t + 1] = P[3];
}

static void accuracy() {
    const int nk8 = 600;  // K = 4800 = 192 channels x 25 taps
    std::vector<float> A((size_t)nk8 * 128), B((size_t)nk8 * 64);
    srand(11);
    // activations like PReLU outputs (mostly positive), weights zero-mean
    for (auto& v : A) v = fabsf(frand()) * 2.f + (rand() % 4 == 0 ? -0.3f * fabsf(frand()) : 0.f);
    for (auto& v : B) v = frand() * 0.03f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 512);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    double ref[128], fp32ref[128], scale = 0;
    for (int i = 0; i < 16; i++) for (int n = 0; n < 8; n++) {
        double s = 0; float f = 0.f;
        for (int st = 0; st < nk8; st++) for (int k = 0; k < 8; k++) {
            s += (double)A[(size_t)st * 128 + i * 8 + k] * B[(size_t)st * 64 + k * 8 + n];
            f = fmaf(A[(size_t)st * 128 + i * 8 + k], B[(size_t)st * 64 + k * 8 + n], f);
        }
        ref[i * 8 + n] = s; fp32ref[i * 8 + n] = f; scale = fmax(scale, fabs(s));
    }
    double e32 = 0;
    for (int i = 0; i < 128; i++) e32 = fmax(e32, fabs(fp32ref[i] - ref[i]));
    printf("accuracy over K = 4800 (max |x - fp64| / max |fp64|): sequential fp32 fmaf %.3g\n", e32 / scale);
    for (int nterms = 3; nterms <= 4; nterms++)
        for (int chunk : {25, 50, 100, 600}) {
            float D[128];
            k_split<<<1, 32>>>(dA, dB, dD, nk8, chunk, nterms);
            cudaMemcpy(D, dD, 512, cudaMemcpyDeviceToHost);
            double e = 0, bias = 0;
            for (int i = 0; i < 128; i++) { e = fmax(e, fabs(D[i] - ref[i])); bias += (D[i] - ref[i]); }
            printf("  split tf32 %d terms, fold every %3d k-steps: max rel err %.3g, mean signed err %.3g\n", nterms, chunk, e / scale, bias / 128 / scale);
        }
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    semantics();
    accuracy();
    for (int warps : {4, 8, 16}) rate<0, 4>("tf32", warps, 1);
    rate<0, 8>("tf32", 8, 1);
    rate<0, 8>("tf32", 8, 2);
    rate<0, 2>("tf32", 16, 2);
    rate<1, 4>("bf16", 8, 1);
    rate<1, 8>("bf16", 8, 2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
