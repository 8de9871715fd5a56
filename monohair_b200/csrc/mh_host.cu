// Host-side pieces of the C ABI: error string, camera record packing, host test hooks.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include "mh_common.cuh"
#include "mh_topk.cuh"

static thread_local char g_err[512] = "";

void mh_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

long long g_mh_launches = 0;
extern "C" int64_t mh_launch_count(void) { return (int64_t)g_mh_launches; }
extern "C" const char* mh_last_error(void) { return g_err; }
extern "C" int mh_version(void) { return 100; }

extern "C" int mh_views_pack_camera_host(const float* pose, const float* ndc_prj, const float* rinv, float* cam) {
    MH_CHECK_ARG(pose && ndc_prj && rinv && cam, "null pointer");
    MhCam c;
    memset(&c, 0, sizeof(c));
    for (int i = 0; i < 12; ++i) c.p[i] = pose[i];
    c.fx = ndc_prj[0]; c.fy = ndc_prj[1]; c.cx = ndc_prj[2]; c.cy = ndc_prj[3];
    for (int i = 0; i < 9; ++i) c.rinv[i] = rinv[i];
    c.t[0] = pose[3]; c.t[1] = pose[7]; c.t[2] = pose[11];
    memcpy(cam, &c, sizeof(c));
    return 0;
}

extern "C" int mh_debug_topk_host(const float* values, int32_t V, int32_t k, int32_t* idx, float* val) {
    MH_CHECK_ARG(values && idx && val && V > 0 && k > 0 && k <= V, "bad arguments");
    std::vector<MhKV> q(V);
    for (int j = 0; j < V; ++j) { q[j].v = values[j]; q[j].i = j; }
    mh_topk_torch_cpu(q.data(), V, k);
    for (int j = 0; j < k; ++j) { idx[j] = q[j].i; val[j] = q[j].v; }
    return 0;
}


// ---- GPU self-check of mh_div2 against the IEEE operator (tests/test_gpu_pmvo.py) ----
namespace {
__global__ void div_check_kernel(unsigned long long seed, long long n, unsigned long long* mism) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // counter-based generator (splitmix64)
    auto mix = [](unsigned long long z) { z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
                                          z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); };
    const unsigned long long r0 = mix(seed + 3 * i), r1 = mix(seed + 3 * i + 1), r2 = mix(seed + 3 * i + 2);
    // mantissas uniform, exponents in the ranges the kernels see (and beyond): 2^-40 .. 2^40
    auto mk = [](unsigned long long r) { const int e = 127 - 40 + (int)((r >> 40) % 81); const unsigned m = (unsigned)(r & 0x7fffff);
                                         return __uint_as_float((unsigned)((r >> 63) << 31) | ((unsigned)e << 23) | m); };
    const float a0 = mk(r0), a1 = (i & 7) == 0 ? 0.0f : mk(r1), b = mk(r2);
    float q0, q1;
    mh_div2(a0, a1, b, q0, q1);
    const float e0 = __fdiv_rn(a0, b), e1 = __fdiv_rn(a1, b);
    if (__float_as_uint(q0) != __float_as_uint(e0) || __float_as_uint(q1) != __float_as_uint(e1)) atomicAdd(mism, 1ull);
}
}  // namespace

extern "C" int mh_debug_div_check(void* stream, int64_t n, uint64_t seed, uint64_t* mismatches_dev) {
    MH_CHECK_ARG(mismatches_dev && n > 0, "bad arguments");
    div_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, n, (unsigned long long*)mismatches_dev);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
