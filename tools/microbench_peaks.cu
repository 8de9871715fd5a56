// FP32 / issue-rate peaks of the SM as this library's kernels can reach them (sm_100a), measured with plain loops.
// Written for the roofline of the ALU-bound kernels (optimize_kernel, medoid): bench.py reads the JSON this prints.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_peaks tools/microbench_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;             // independent chains per thread

#define LOOP_BODY(BODY) \
    float a[CH]; unsigned long long p[CH]; int q[CH]; \
    for (int c = 0; c < CH; ++c) { a[c] = seed[c] + threadIdx.x * 1e-6f; p[c] = ((unsigned long long)__float_as_uint(a[c]) << 32) | __float_as_uint(a[c] * 0.5f); q[c] = threadIdx.x + c; } \
    const float b = seed[8], d = seed[9]; const unsigned long long b2 = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b); \
    const unsigned long long d2 = ((unsigned long long)__float_as_uint(d) << 32) | __float_as_uint(d); (void)b2; (void)d2; (void)q; (void)p; \
    for (int it = 0; it < ITERS; ++it) { _Pragma("unroll") for (int c = 0; c < CH; ++c) { BODY } } \
    float s = 0; for (int c = 0; c < CH; ++c) s += a[c] + __uint_as_float((unsigned)p[c]) + __uint_as_float((unsigned)(p[c] >> 32)) + (float)q[c]; \
    if (s == 12345.678f) out[0] = s;

__global__ void k_ffma(const float* seed, float* out) { LOOP_BODY(asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(b), "f"(d));) }
__global__ void k_fmul_fadd(const float* seed, float* out) { LOOP_BODY(asm volatile("mul.rn.f32 %0, %0, %1;\n\tadd.rn.f32 %0, %0, %2;" : "+f"(a[c]) : "f"(b), "f"(d));) }
__global__ void k_ffma2(const float* seed, float* out) { LOOP_BODY(asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(b2), "l"(d2));) }
__global__ void k_fmul2_fadd2(const float* seed, float* out) { LOOP_BODY(asm volatile("mul.rn.f32x2 %0, %0, %1;\n\tadd.rn.f32x2 %0, %0, %2;" : "+l"(p[c]) : "l"(b2), "l"(d2));) }
__global__ void k_fadd(const float* seed, float* out) { LOOP_BODY(asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(d));) }
__global__ void k_iadd_lop(const float* seed, float* out) { LOOP_BODY(asm volatile("add.s32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(q[c]) : "r"(it), "r"(c + 77));) }
__global__ void k_mix_ffma_iadd(const float* seed, float* out) { LOOP_BODY(asm volatile("fma.rn.f32 %0, %0, %2, %3;\n\tadd.s32 %1, %1, %4;" : "+f"(a[c]), "+r"(q[c]) : "f"(b), "f"(d), "r"(it));) }
__global__ void k_mix_ffma_fadd_iadd(const float* seed, float* out) { LOOP_BODY(asm volatile("fma.rn.f32 %0, %0, %2, %3;\n\tadd.s32 %1, %1, %4;\n\txor.b32 %1, %1, %5;" : "+f"(a[c]), "+r"(q[c]) : "f"(b), "f"(d), "r"(it), "r"(c + 3));) }
__global__ void k_div(const float* seed, float* out) { LOOP_BODY(asm volatile("div.rn.f32 %0, %1, %0;" : "+f"(a[c]) : "f"(b));) }
__global__ void k_sqrt(const float* seed, float* out) { LOOP_BODY(asm volatile("sqrt.rn.f32 %0, %0;\n\tadd.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b));) }
__global__ void k_lds(const float* seed, float* out) {
    __shared__ float4 sm[256];
    sm[threadIdx.x & 255] = make_float4(seed[0], seed[1], seed[2], seed[3]);
    __syncthreads();
    LOOP_BODY(float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(it + c) & 255]))); a[c] += v.x;)
}

template <typename K>
double run(K kern, const char* name, double ops_per_body, double flops_per_body, const float* seed, float* out, int sms, FILE* js, bool last = false) {
    const int blocks = sms * 2, threads = 1024;
    kern<<<blocks, threads>>>(seed, out);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        kern<<<blocks, threads>>>(seed, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double bodies = (double)blocks * threads * ITERS * CH;      // thread-level bodies
    const double winstr = bodies * ops_per_body / 32.0;               // warp instructions
    const double gwips = winstr / (best * 1e-3) / 1e9;                // G warp-instr / s, whole chip
    const double tflops = bodies * flops_per_body / (best * 1e-3) / 1e12;
    fprintf(stderr, "%-22s %8.3f ms  %8.1f G warp-instr/s  %7.2f TFLOP/s\n", name, best, gwips, tflops);
    fprintf(js, "  \"%s\": {\"ms\": %.4f, \"g_warp_instr_per_s\": %.2f, \"tflops\": %.3f}%s\n", name, best, gwips, tflops, last ? "" : ",");
    return gwips;
}

int main() {
    int dev = 0, sms = 148, clk = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    float h[16]; for (int i = 0; i < 16; ++i) h[i] = 1.0f + 1e-3f * i;
    h[8] = 0.999f; h[9] = 1e-3f;
    float *seed, *out;
    cudaMalloc(&seed, sizeof(h)); cudaMalloc(&out, 64);
    cudaMemcpy(seed, h, sizeof(h), cudaMemcpyHostToDevice);
    FILE* js = stdout;
    fprintf(js, "{\n  \"sms\": %d, \"clock_khz\": %d,\n", sms, clk);
    run(k_ffma, "ffma", 1, 2, seed, out, sms, js);
    run(k_ffma2, "ffma2", 1, 4, seed, out, sms, js);
    run(k_fmul_fadd, "fmul_fadd", 2, 2, seed, out, sms, js);
    run(k_fmul2_fadd2, "fmul2_fadd2", 2, 4, seed, out, sms, js);
    run(k_fadd, "fadd", 1, 1, seed, out, sms, js);
    run(k_iadd_lop, "iadd_xor", 2, 0, seed, out, sms, js);
    run(k_mix_ffma_iadd, "mix_ffma_iadd", 2, 2, seed, out, sms, js);
    run(k_mix_ffma_fadd_iadd, "mix_ffma_iadd_xor", 3, 2, seed, out, sms, js);
    run(k_div, "div_rn", 1, 1, seed, out, sms, js);
    run(k_sqrt, "sqrt_rn_fadd", 2, 2, seed, out, sms, js);
    run(k_lds, "lds128_fadd", 2, 1, seed, out, sms, js, true);
    fprintf(js, "}\n");
    return 0;
}
