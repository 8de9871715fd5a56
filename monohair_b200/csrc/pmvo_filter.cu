// Per-view projection + centre-pixel gathers + multi-view counters:
//   mh_filter_count / mh_filter_decide   PMVO.filter_points            (PMVO.py:402-459)
//   mh_visible_count                     PMVO.compute_unvisible_points (PMVO.py:461-480)
//   mh_head_count                        PMVO.filter_head_points       (PMVO.py:110-136)
// One thread per point, views in the outer loop so that all resident CTAs walk the views together and one
// view's mapC plane (16 B/px) stays L2-resident while it is being gathered.  Everything these kernels need of a (point, view)
// pair -- depth, mask', PxP maximum of the confidence -- sits in ONE mapC texel: one 32 B sector per pair.
// Bound: L2/HBM gather bandwidth; algorithmic bytes per (point, view) = 12 B ({depth, mask', max conf}).
#include "mh_common.cuh"

namespace {

constexpr int CAM_CHUNK = 96;      // views staged in shared memory at a time (12 KB)

enum Mode { FILTER = 0, VISCOUNT = 1, HEAD = 2 };

template <int MODE>
__global__ void __launch_bounds__(256)
count_kernel(mh_views vw, const float* __restrict__ pts, int64_t N, float thr_v, float thr_c,
             float* __restrict__ out) {
    __shared__ MhCam cams[CAM_CHUNK];
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = n < N;
    float x = 0, y = 0, z = 0;
    if (live) { x = pts[3 * n]; y = pts[3 * n + 1]; z = pts[3 * n + 2]; }
    constexpr int NC = (MODE == FILTER) ? 5 : (MODE == HEAD ? 2 : 1);
    MhCascade<NC> acc;
    acc.init(vw.V);
    const float4* __restrict__ mapC = reinterpret_cast<const float4*>(vw.mapC);
    const size_t plane = (size_t)vw.H * vw.W;
    const float Wf = (float)vw.W, Hf = (float)vw.H;
    for (int vb = 0; vb < vw.V; vb += CAM_CHUNK) {
        const int nv = min(CAM_CHUNK, vw.V - vb);
        __syncthreads();
        for (int i = threadIdx.x; i < nv * MH_CAM_STRIDE; i += blockDim.x)
            reinterpret_cast<float*>(cams)[i] = vw.cam[(size_t)vb * MH_CAM_STRIDE + i];
        __syncthreads();
        if (!live) continue;
        for (int j = 0; j < nv; ++j) {
            const int v = vb + j;
            const MhCam& cm = cams[j];
            float cx, cy, cz, xp, yp;
            mh_world_to_cam(cm.p, x, y, z, cx, cy, cz);
            mh_cam_to_xy(cm.fx, cm.fy, cm.cx, cm.cy, Wf, Hf, cx, cy, cz, xp, yp);
            int row, col; bool oob;
            mh_round_clamp(xp, yp, vw.W, vw.H, row, col, oob);
            const size_t pix = (size_t)v * plane + (size_t)row * vw.W + col;
            const float4 dm = __ldg(mapC + pix);
            const float delta = (-cz / 2.0f) * 255.0f - dm.x;
            if (MODE == FILTER) {
                float cmax = dm.z;                                   // PxP maximum, same texel (one sector per pair)
                if (oob) cmax = 0.0f;
                const float vis = (delta > 0.1f || oob) ? 0.0f : 1.0f;
                const float vis1 = (delta > thr_v || oob) ? 0.0f : 1.0f;
                const float lowc = (cmax < thr_c) ? 1.0f : 0.0f;
                acc.begin_row(v);
                acc.add(0, vis);
                acc.add(1, vis * dm.y);
                acc.add(2, vis * lowc);
                acc.add(3, vis1);
                acc.add(4, vis1 * dm.y);
            } else if (MODE == VISCOUNT) {
                acc.begin_row(v);
                acc.add(0, (delta > thr_v || oob) ? 0.0f : 1.0f);
            } else {
                const float vis = (delta >= thr_v) ? 0.0f : 1.0f;
                acc.begin_row(v);
                acc.add(0, vis);
                acc.add(1, vis * dm.y);
            }
        }
    }
    if (!live) return;
    float res[NC];
    acc.finish(vw.V, res);
#pragma unroll
    for (int k = 0; k < NC; ++k) out[(int64_t)k * N + n] = res[k];
}

__global__ void decide_kernel(const float* __restrict__ c, int64_t N, uint8_t* __restrict__ surface,
                              uint8_t* __restrict__ filter) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float sv = c[n], svm = c[N + n], si = c[2 * N + n], sv1 = c[3 * N + n], sv1m = c[4 * N + n];
    const bool low = si > 4.0f;
    const bool hair = (sv - svm) < (sv * 1.0f / 2.0f);
    const bool hair1 = (sv1 - sv1m) < (sv1 * 1.0f / 2.0f);
    const bool surf = sv > 1.0f;
    surface[n] = (surf && !low && hair) ? 1 : 0;
    filter[n] = ((sv1 > 1.0f) && !surf && !low && hair1) ? 1 : 0;
}

int check_views(const mh_views* vw) {
    MH_CHECK_ARG(vw && vw->mapC && vw->mapP && vw->cam, "null views");
    MH_CHECK_ARG(vw->V > 0 && vw->H > 0 && vw->W > 0, "bad view sizes");
    return 0;
}

}  // namespace

extern "C" int mh_filter_count(void* stream, const mh_views* views, const float* points, int64_t N,
                               float visible_threshold, float conf_threshold, float* counters) {
    if (check_views(views)) return 1;
    if (N == 0) return 0;
    MH_CHECK_ARG(points && counters && N > 0, "bad arguments");
    count_kernel<FILTER><<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        *views, points, N, visible_threshold, conf_threshold, counters);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_filter_decide(void* stream, const float* counters, int64_t N, uint8_t* surface, uint8_t* filter) {
    if (N == 0) return 0;
    MH_CHECK_ARG(counters && surface && filter && N > 0, "bad arguments");
    decide_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(counters, N, surface, filter);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_visible_count(void* stream, const mh_views* views, const float* points, int64_t N,
                                float dz_threshold, float* count) {
    if (check_views(views)) return 1;
    if (N == 0) return 0;
    MH_CHECK_ARG(points && count && N > 0, "bad arguments");
    count_kernel<VISCOUNT><<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        *views, points, N, dz_threshold, 0.0f, count);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_head_count(void* stream, const mh_views* views, const float* points, int64_t N,
                             float visible_threshold, float* counters) {
    if (check_views(views)) return 1;
    if (N == 0) return 0;
    MH_CHECK_ARG(points && counters && N > 0, "bad arguments");
    count_kernel<HEAD><<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        *views, points, N, visible_threshold, 0.0f, counters);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

// ---- Compute_Visible_and_Ori centre values (PMVO.py:346-376), one thread per (view, point) ----------------
namespace {
__global__ void centre_kernel(mh_views vw, const float* __restrict__ pts, int64_t N, float* __restrict__ visible,
                              float* __restrict__ ori, float* __restrict__ conf, float* __restrict__ mask,
                              int32_t* __restrict__ rowcol) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (n >= N) return;
    const MhCam cm = *reinterpret_cast<const MhCam*>(vw.cam + (size_t)v * MH_CAM_STRIDE);
    float cx, cy, cz, xp, yp;
    mh_world_to_cam(cm.p, pts[3 * n], pts[3 * n + 1], pts[3 * n + 2], cx, cy, cz);
    mh_cam_to_xy(cm.fx, cm.fy, cm.cx, cm.cy, (float)vw.W, (float)vw.H, cx, cy, cz, xp, yp);
    int row, col; bool oob;
    mh_round_clamp(xp, yp, vw.W, vw.H, row, col, oob);
    const size_t pix = (size_t)v * vw.H * vw.W + (size_t)row * vw.W + col;
    const float4 dm = __ldg(reinterpret_cast<const float4*>(vw.mapC) + pix);
    const float4 oc = __ldg(reinterpret_cast<const float4*>(vw.mapP) + pix);
    float vis = mh_visible((-cz / 2.0f) * 255.0f, dm.x);
    if (oob) vis = -1.0f;
    const size_t o = (size_t)v * N + n;
    visible[o] = vis;
    ori[2 * o] = dm.w; ori[2 * o + 1] = oc.w;
    conf[o] = fminf(fmaxf(oc.z, 1e-6f), 1.0f);
    if (mask) mask[o] = dm.y;
    if (rowcol) { rowcol[2 * o] = row; rowcol[2 * o + 1] = oob ? -col - 1 : col; }
}

__global__ void head_decide_kernel(const float* __restrict__ c, const double* __restrict__ dist, const float* __restrict__ pts,
                                   int64_t N, double dist_thr, double z_thr, uint8_t* __restrict__ filt) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float sv = c[n], si = c[N + n];
    const bool keep = (sv - si) < (sv * 1.0f / 2.0f);
    const bool head_top = (dist[n] < dist_thr) && ((double)pts[3 * n + 2] < z_thr);
    filt[n] = (!keep && !head_top) ? 1 : 0;
}
}  // namespace

extern "C" int mh_centre_gather(void* stream, const mh_views* views, const float* points, int64_t N, float* visible,
                                float* ori, float* conf, float* mask, int32_t* rowcol) {
    if (check_views(views)) return 1;
    if (N == 0) return 0;
    MH_CHECK_ARG(points && visible && ori && conf && N > 0, "bad arguments");
    dim3 grid((unsigned)((N + 255) / 256), views->V);
    centre_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*views, points, N, visible, ori, conf, mask, rowcol);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_head_decide(void* stream, const float* counters, const double* scalp_dist, const float* points,
                              int64_t N, double dist_threshold, double z_threshold, uint8_t* filter) {
    MH_CHECK_ARG(counters && scalp_dist && points && filter && N >= 0, "bad arguments");
    if (N == 0) return 0;
    head_decide_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(counters, scalp_dist, points, N,
                                                                                   dist_threshold, z_threshold, filter);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
