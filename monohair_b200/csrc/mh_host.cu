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

// test hook: count operand triples for which mh_div2 differs from (a0 / b, a1 / b) in any bit
__global__ void div2_check_kernel(const float* __restrict__ a0, const float* __restrict__ a1, const float* __restrict__ b,
                                  int64_t n, unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float q0, q1;
        mh_div2(a0[i], a1[i], b[i], q0, q1);
        const float r0 = __fdiv_rn(a0[i], b[i]), r1 = __fdiv_rn(a1[i], b[i]);
        bad += (__float_as_uint(q0) != __float_as_uint(r0)) + (__float_as_uint(q1) != __float_as_uint(r1));
    }
    if (bad) atomicAdd(mismatches, bad);
}
extern "C" int mh_debug_div2_check(void* stream, const float* a0, const float* a1, const float* b, int64_t n,
                                   unsigned long long* mismatches) {
    MH_CHECK_ARG(a0 && a1 && b && mismatches && n >= 0, "bad arguments");
    if (n == 0) return 0;
    div2_check_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(a0, a1, b, n, mismatches);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
