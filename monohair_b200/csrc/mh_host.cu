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
