// PMVO.forward as one fused kernel (PMVO.py:39-78):
//   Compute_Visible_and_Ori (:346-376) -> Find_max_conf_from_visible_view (:339-343) -> for 10 base views:
//   sample_next_3d_pos (:263-335) -> compute_reproject_ori (:219-241) -> compute_prj_loss (:151-209)
//   -> per-point best over base views -> normalised direction.
//
// One CTA (OPT_WARPS warps) per point at a time, persistent over a work counter.
//   A   thread v      : project the point into view v, centre gathers, visibility, Conf'
//   B   warp0/lane0   : torch.topk order over the views (mh_topk.cuh)          } concurrently
//   A2  other warps   : stage each VISIBLE view's PxP patch in shared memory   }
//                       as {unit ori, conf}, dropping entries that can never win the reference's scan:
//                       ineligible ones (conf<=thr in a patch that has some conf>thr, except entry 0) and exact
//                       duplicates of an earlier kept direction (strict '<' keeps the first) -- both exact.
//   C   warp w        : base views w, w+OPT_WARPS, ...; each lane owns SPL consecutive depth samples and walks the
//                       visible views x staged entries with all of them in flight (ILP); the [V,N,S] tensors of
//                       the reference never exist.  View sums follow torch.sum's cascade order.
//   D   thread 0      : best base view, direction.
// Invisible views have weight exactly 0 in the reference (compute_weight, :211-215) so skipping them is exact.
//
// Bound: FP32 ALU / issue: ~7 instr per (sample, view, staged entry) + ~70 per (sample, view).
// HBM/L2 traffic is only the gathers of phase A/A2: (8 + 16) B per (point, view) + 16 B per (point, visible view,
// patch entry).
#include "mh_common.cuh"
#include "mh_topk.cuh"

namespace {

#ifndef OPT_WARPS_N
#define OPT_WARPS_N 5
#endif
constexpr int OPT_WARPS = OPT_WARPS_N;
constexpr int OPT_THREADS = 32 * OPT_WARPS;
#ifndef OPT_MIN_BLOCKS
#define OPT_MIN_BLOCKS 4
#endif
constexpr int MAX_SPL = 4;                     // samples per lane: S <= 128

struct Smem {
    MhCam* cams;        // [V]
    float* camz;        // [V] camera z of the point
    float* xp;          // [V] float pixel x (col)
    float* yp;          // [V] float pixel y (row)
    float* vis;         // [V]
    float* orr;         // [V] raw ori (d_row) at the centre pixel
    float* orc;         // [V]
    int* pix;           // [V] row*W+col (clamped)
    int* ecnt;          // [V] staged entries of view v (0 = invisible)
    int* vlist;         // [V] visible views ascending
    MhKV* q;            // [V]
    float* off;         // [S]
    float2* exy;        // [V][PP] unit patch directions (d_row, d_col)
    float* ec;          // [V][PP] patch confidences (clamped)
    float* res_loss;    // [NUM_BASE]
    int* res_arg;       // [NUM_BASE]
    int* res_flags;     // [NUM_BASE]  bit0 = high_conf, bit1 = valid
    int* misc;          // [4]: 0 = nvis, 1/2 = current point index (lo/hi)
};

__host__ __device__ inline size_t smem_layout(Smem* s, unsigned char* base, int V, int S, int PP) {
    size_t o = 0;
    auto take = [&](size_t bytes, size_t align) { o = (o + align - 1) / align * align; size_t r = o; o += bytes; return r; };
    size_t o_cams = take(sizeof(MhCam) * V, 16);
    size_t o_camz = take(4 * V, 4), o_xp = take(4 * V, 4), o_yp = take(4 * V, 4), o_vis = take(4 * V, 4);
    size_t o_orr = take(4 * V, 4), o_orc = take(4 * V, 4), o_pix = take(4 * V, 4), o_ecnt = take(4 * V, 4);
    size_t o_vl = take(4 * V, 4), o_q = take(sizeof(MhKV) * V, 8), o_off = take(4 * S, 4);
    size_t o_exy = take(sizeof(float2) * (size_t)V * PP, 8), o_ec = take(4 * (size_t)V * PP, 4);
    size_t o_rl = take(4 * MH_NUM_BASE, 4), o_ra = take(4 * MH_NUM_BASE, 4), o_rf = take(4 * MH_NUM_BASE, 4);
    size_t o_misc = take(16, 4);
    if (s) {
        s->cams = (MhCam*)(base + o_cams); s->camz = (float*)(base + o_camz); s->xp = (float*)(base + o_xp);
        s->yp = (float*)(base + o_yp); s->vis = (float*)(base + o_vis); s->orr = (float*)(base + o_orr);
        s->orc = (float*)(base + o_orc); s->pix = (int*)(base + o_pix); s->ecnt = (int*)(base + o_ecnt);
        s->vlist = (int*)(base + o_vl); s->q = (MhKV*)(base + o_q); s->off = (float*)(base + o_off);
        s->exy = (float2*)(base + o_exy); s->ec = (float*)(base + o_ec);
        s->res_loss = (float*)(base + o_rl); s->res_arg = (int*)(base + o_ra);
        s->res_flags = (int*)(base + o_rf); s->misc = (int*)(base + o_misc);
    }
    return o;
}

// sample s of the ray through the point's pixel in base view i stepped 2 px along the 2-D orientation
// (sample_next_3d_pos :290-328 + Camera.reprojection, Camera_utils.py:95-104).
MH_D void sample_point(const Smem& sm, int i, float off, float Wf, float Hf, float& wx, float& wy, float& wz) {
    const MhCam& cm = sm.cams[i];
    float nx = sm.xp[i] + sm.orc[i] * 2.0f;
    float ny = sm.yp[i] + sm.orr[i] * 2.0f;
    nx = nx / Wf; ny = ny / Hf;
    nx = nx * 2.0f - 1.0f; ny = ny * 2.0f - 1.0f;
    nx = -nx;
    const float zs = sm.camz[i] + off;
    const float d0 = (nx - cm.cx) / cm.fx * zs - cm.t[0];
    const float d1 = (ny - cm.cy) / cm.fy * zs - cm.t[1];
    const float d2 = zs - cm.t[2];
    // MKL's order for inv(3x3, column-major) @ [3,M] (row stride 1) at the reference's sizes: (r0*d0 + r2*d2) + r1*d1,
    // products rounded separately (probe: DESIGN.md §4).
    wx = (cm.rinv[0] * d0 + cm.rinv[2] * d2) + cm.rinv[1] * d1;
    wy = (cm.rinv[3] * d0 + cm.rinv[5] * d2) + cm.rinv[4] * d1;
    wz = (cm.rinv[6] * d0 + cm.rinv[8] * d2) + cm.rinv[7] * d1;
}

// torch.min(dim) semantics: first minimum, first NaN wins.
MH_D bool arg_better(float av, int ai, float bv, int bi) {
    const bool an = av != av, bn = bv != bv;
    if (an || bn) return (an && bn) ? (ai < bi) : an;
    return av < bv || (av == bv && ai < bi);
}

// One base view handled by one warp: compute_prj_loss (:151-209) for its S samples.
template <int SPL, typename Cascade>
MH_D void process_base(const Smem& sm, const mh_views& vw, int b, int S, int PP, float thr_c, int lane) {
    const int V = vw.V;
    const float Wf = (float)vw.W, Hf = (float)vw.H;
    const int bview = sm.q[2 * b].i;
    const float bval = sm.q[2 * b].v;
    const bool valid = (b == 0) || (bval > 0.0f);                                     // :57-64
    if (!valid) {                                                                      // warp-uniform
        if (lane == 0) { sm.res_loss[b] = 0.0f; sm.res_arg[b] = 0; sm.res_flags[b] = 0; }
        return;
    }
    const int nvis = sm.misc[0];
    float wx[SPL], wy[SPL], wz[SPL];
    int cnt[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = min(lane * SPL + j, S - 1);
        sample_point(sm, bview, sm.off[s], Wf, Hf, wx[j], wy[j], wz[j]);
        cnt[j] = 0;
    }
    Cascade acc;                                                                       // channels: 2j = sum l*w, 2j+1 = sum w
    acc.init(V);
    for (int jv = 0; jv < nvis; ++jv) {
        const int v = sm.vlist[jv];
        const MhCam& cm = sm.cams[v];
        const float ypv = sm.yp[v], xpv = sm.xp[v];
        float y0[SPL], y1[SPL];
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            float cx, cy, cz, xs, ys;
            mh_world_to_cam(cm.p, wx[j], wy[j], wz[j], cx, cy, cz);
            mh_cam_to_xy(cm.fx, cm.fy, cm.cx, cm.cy, Wf, Hf, cx, cy, cz, xs, ys);
            mh_normalize2(ys - ypv, xs - xpv, y0[j], y1[j]);                           // (d_row, d_col) :237
        }
        // scan of the staged entries (:164-182): entry 0 seeds, later entries replace on strict '<'
        const float2* __restrict__ exy = sm.exy + (size_t)v * PP;
        const int ne = sm.ecnt[v];
        float bl[SPL];
        int bk[SPL];
        {
            const float2 x = exy[0];
#pragma unroll
            for (int j = 0; j < SPL; ++j) { bl[j] = 1.0f - fabsf(x.x * y0[j] + x.y * y1[j]); bk[j] = 0; }
        }
#pragma unroll 4
        for (int k = 1; k < ne; ++k) {
            const float2 x = exy[k];
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                const float l = 1.0f - fabsf(x.x * y0[j] + x.y * y1[j]);
                if (l < bl[j]) { bl[j] = l; bk[j] = k; }
            }
        }
        const float* __restrict__ ec = sm.ec + (size_t)v * PP;
        acc.begin_row(v);
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const float c = ec[bk[j]];
            acc.add(2 * j, bl[j] * c);                                                 // min_loss * weight :195
            acc.add(2 * j + 1, c);
            cnt[j] += (c > 0.0f) ? 1 : 0;
        }
    }
    float sums[2 * SPL];
    acc.finish(V, sums);
    float Lraw[SPL];
    bool pos[SPL];
    int npos = 0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const bool act = lane * SPL + j < S;
        pos[j] = act && ((sums[2 * j + 1] / (float)cnt[j]) > thr_c);                   // :198
        Lraw[j] = sums[2 * j] / sums[2 * j + 1];                                       // :201
        npos += __popc(__ballot_sync(0xffffffffu, pos[j]));
    }
    const bool low = npos < 5;                                                          // :199
    float bestv = 0.0f; int besti = 0x7fffffff; bool bestpos = false;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s < S) {
            const float L = (low || pos[j]) ? Lraw[j] : 1.0f;                          // :203-204
            if (besti == 0x7fffffff || arg_better(L, s, bestv, besti)) { bestv = L; besti = s; bestpos = pos[j]; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        const int op = __shfl_xor_sync(0xffffffffu, (int)bestpos, o);
        if (oi != 0x7fffffff && (besti == 0x7fffffff || arg_better(ov, oi, bestv, besti))) { bestv = ov; besti = oi; bestpos = op != 0; }
    }
    if (lane == 0) {
        sm.res_loss[b] = bestv;
        sm.res_arg[b] = besti;
        sm.res_flags[b] = (bestpos ? 1 : 0) | 2;
    }
}

template <int SPL, bool BIGV>
__global__ void __launch_bounds__(OPT_THREADS, OPT_MIN_BLOCKS)
optimize_kernel(mh_views vw, const float* __restrict__ pts, int64_t N, const float* __restrict__ offsets, int S,
                float thr_c, float* __restrict__ out_ori, float* __restrict__ out_loss,
                uint8_t* __restrict__ out_hc, int32_t* __restrict__ dbg_bidx, float* __restrict__ dbg_bval,
                float* __restrict__ dbg_best, float* __restrict__ dbg_loss_b, int32_t* __restrict__ dbg_arg_b,
                unsigned long long* __restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem sm;
    const int V = vw.V, P = vw.P, PP = P * P, half = P / 2;
    smem_layout(&sm, smem_raw, V, S, PP);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float Wf = (float)vw.W, Hf = (float)vw.H;
    const float4* __restrict__ mapC = reinterpret_cast<const float4*>(vw.mapC);
    const float4* __restrict__ mapP = reinterpret_cast<const float4*>(vw.mapP);
    const size_t plane = (size_t)vw.H * vw.W;

    for (int i = tid; i < V * MH_CAM_STRIDE; i += OPT_THREADS) reinterpret_cast<float*>(sm.cams)[i] = vw.cam[i];
    for (int i = tid; i < S; i += OPT_THREADS) sm.off[i] = offsets[i];
    if (tid == 0) { sm.misc[1] = (int)blockIdx.x; sm.misc[2] = 0; }
    __syncthreads();

    for (;;) {
        const int64_t n = (int64_t)(unsigned)sm.misc[1] | ((int64_t)(unsigned)sm.misc[2] << 32);
        if (n >= N) break;
        const float px = pts[3 * n], py = pts[3 * n + 1], pz = pts[3 * n + 2];

        // ---- A: per view projection + centre gathers (Compute_Visible_and_Ori :353-362) ----
        for (int v = tid; v < V; v += OPT_THREADS) {
            const MhCam& cm = sm.cams[v];
            float cx, cy, cz, xp, yp;
            mh_world_to_cam(cm.p, px, py, pz, cx, cy, cz);
            mh_cam_to_xy(cm.fx, cm.fy, cm.cx, cm.cy, Wf, Hf, cx, cy, cz, xp, yp);
            int row, col; bool oob;
            mh_round_clamp(xp, yp, vw.W, vw.H, row, col, oob);
            const int pix = row * vw.W + col;
            const float4 dm = __ldg(mapC + (size_t)v * plane + pix);
            const float4 oc = __ldg(mapP + (size_t)v * plane + pix);
            float vis = mh_visible((-cz / 2.0f) * 255.0f, dm.x);
            if (oob) vis = -1.0f;
            const float conf = fminf(fmaxf(oc.z, 1e-6f), 1.0f);
            const float confp = (vis < 1.0f) ? conf * fmaxf(vis, 0.0f) : conf;      // :340
            sm.camz[v] = cz; sm.xp[v] = xp; sm.yp[v] = yp; sm.vis[v] = vis;
            sm.orr[v] = dm.w; sm.orc[v] = oc.w; sm.pix[v] = pix;
            sm.q[v].v = confp; sm.q[v].i = v;
        }
        __syncthreads();

        if (warp == 0) {
            // ---- B: base views ----
            if (lane == 0) mh_topk_torch_cpu(sm.q, V, MH_TOPK);
        } else {
            // ---- A2: stage patches of visible views ----
            if (warp == 1) {
                int cnt = 0;
                for (int v0 = 0; v0 < V; v0 += 32) {
                    const int v = v0 + lane;
                    const bool isv = v < V && sm.vis[v] != -1.0f;
                    const unsigned m = __ballot_sync(0xffffffffu, isv);
                    if (isv) sm.vlist[cnt + __popc(m & ((1u << lane) - 1))] = v;
                    cnt += __popc(m);
                }
                if (lane == 0) sm.misc[0] = cnt;
            }
            for (int v = warp - 1; v < V; v += OPT_WARPS - 1) {
                if (sm.vis[v] == -1.0f) { if (lane == 0) sm.ecnt[v] = 0; continue; }
                const int pix = sm.pix[v];
                const int row = pix / vw.W, col = pix - row * vw.W;
                const float4* __restrict__ mp = mapP + (size_t)v * plane;
                const float cmax = fminf(fmaxf(__ldg(reinterpret_cast<const float*>(mapC + (size_t)v * plane + pix) + 2), 1e-6f), 1.0f);
                const bool hi = cmax > thr_c;                                        // :162
                float2* sxy = sm.exy + (size_t)v * PP;
                float* sc = sm.ec + (size_t)v * PP;
                int kept = 0;
                for (int p0 = 0; p0 < PP; p0 += 32) {
                    const int p = p0 + lane;
                    float ex0 = 0.0f, ex1 = 0.0f, ecf = 0.0f;
                    bool elig = false;
                    if (p < PP) {
                        const int di = p / P - half, dj = p % P - half;                 // row offset outer (:494-500)
                        const int r = min(max(row + di, 0), vw.H - 1), c = min(max(col + dj, 0), vw.W - 1);
                        const float4 t = __ldg(mp + (size_t)r * vw.W + c);
                        ex0 = t.x; ex1 = t.y;                                            // stored normalised (pmvo_views.cu)
                        ecf = fminf(fmaxf(t.z, 1e-6f), 1.0f);
                        elig = (p == 0) || !hi || (ecf > thr_c);
                        // duplicate of an entry already kept in an earlier 32-chunk?
                        if (elig) for (int k = 0; k < kept; ++k) if (sxy[k].x == ex0 && sxy[k].y == ex1) { elig = false; break; }
                    }
                    // duplicate of an earlier eligible lane in this chunk?  (bitwise match: -0/+0 are kept apart,
                    // which only keeps a harmless extra entry)
                    const unsigned long long key = ((unsigned long long)__float_as_uint(ex0) << 32) | __float_as_uint(ex1);
                    const unsigned peers = __match_any_sync(0xffffffffu, key);
                    const unsigned em = __ballot_sync(0xffffffffu, elig);
                    const bool dup = (peers & em & ((1u << lane) - 1)) != 0;
                    const bool keep = elig && !dup;
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    __syncwarp();
                    if (keep) {
                        const int slot = kept + __popc(m & ((1u << lane) - 1));
                        sxy[slot] = make_float2(ex0, ex1);
                        sc[slot] = ecf;
                    }
                    kept += __popc(m);
                    __syncwarp();
                }
                if (lane == 0) sm.ecnt[v] = kept;
            }
        }
        __syncthreads();

        // ---- C: base views, one warp each ----
        for (int b = warp; b < MH_NUM_BASE; b += OPT_WARPS) {
            if (BIGV) process_base<SPL, MhCascade<2 * SPL>>(sm, vw, b, S, PP, thr_c, lane);
            else process_base<SPL, MhCascadeSmall<2 * SPL>>(sm, vw, b, S, PP, thr_c, lane);
        }
        __syncthreads();

        // ---- D: best base view (forward :57-74) ----
        if (tid == 0) {
            float ml = sm.res_loss[0];
            int bb = 0;
            for (int b = 1; b < MH_NUM_BASE; ++b)
                if ((sm.res_flags[b] & 2) && sm.res_loss[b] < ml) { ml = sm.res_loss[b]; bb = b; }
            float wx, wy, wz;
            sample_point(sm, sm.q[2 * bb].i, sm.off[sm.res_arg[bb]], Wf, Hf, wx, wy, wz);
            const float dx = wx - px, dy = wy - py, dz = wz - pz;
            const float nn = mh_norm3(dx, dy, dz);
            out_ori[3 * n] = dx / nn; out_ori[3 * n + 1] = dy / nn; out_ori[3 * n + 2] = dz / nn;
            out_loss[n] = ml;
            out_hc[n] = (uint8_t)(sm.res_flags[bb] & 1);
            if (dbg_best) { dbg_best[3 * n] = wx; dbg_best[3 * n + 1] = wy; dbg_best[3 * n + 2] = wz; }
            const unsigned long long nxt = atomicAdd(work_counter, 1ull) + gridDim.x;
            if (dbg_bidx) for (int k = 0; k < MH_TOPK; ++k) { dbg_bidx[(size_t)k * N + n] = sm.q[k].i; dbg_bval[(size_t)k * N + n] = sm.q[k].v; }
            if (dbg_loss_b) for (int b = 0; b < MH_NUM_BASE; ++b) {
                dbg_loss_b[(size_t)b * N + n] = (sm.res_flags[b] & 2) ? sm.res_loss[b] : __int_as_float(0x7fc00000);
                dbg_arg_b[(size_t)b * N + n] = (sm.res_flags[b] & 2) ? sm.res_arg[b] : -1;
            }
            sm.misc[1] = (int)(unsigned)(nxt & 0xffffffffull);
            sm.misc[2] = (int)(unsigned)(nxt >> 32);
        }
        __syncthreads();
    }
}

typedef void (*OptKernel)(mh_views, const float*, int64_t, const float*, int, float, float*, float*, uint8_t*, int32_t*,
                          float*, float*, float*, int32_t*, unsigned long long*);

OptKernel pick_kernel(int S, int V) {
    const int spl = (S + 31) / 32;
    const bool big = V >= 256;
    switch (spl) {
        case 1: return big ? optimize_kernel<1, true> : optimize_kernel<1, false>;
        case 2: return big ? optimize_kernel<2, true> : optimize_kernel<2, false>;
        case 3: return big ? optimize_kernel<3, true> : optimize_kernel<3, false>;
        default: return big ? optimize_kernel<4, true> : optimize_kernel<4, false>;
    }
}

}  // namespace

extern "C" int64_t mh_pmvo_optimize_workspace_bytes(const mh_views* views, int64_t N) {
    (void)views; (void)N;
    return 256;          // the work counter
}

extern "C" int mh_pmvo_optimize(void* stream, const mh_views* vw, const float* points, int64_t N,
                                const float* offsets, int32_t S, float conf_threshold,
                                float* ori, float* loss, uint8_t* high_conf,
                                int32_t* dbg_base_idx, float* dbg_base_val, float* dbg_best_sample,
                                float* dbg_loss_b, int32_t* dbg_arg_b, void* workspace, int64_t workspace_bytes) {
    MH_CHECK_ARG(vw && vw->mapC && vw->mapP && vw->cam, "null views");
    if (N == 0) return 0;
    MH_CHECK_ARG(points && offsets && ori && loss && high_conf && workspace && N > 0, "null pointer");
    MH_CHECK_ARG(workspace_bytes >= 256, "workspace too small");
    MH_CHECK_ARG(vw->V >= MH_TOPK, "PMVO.forward needs at least 20 views (torch.topk(...,20), PMVO.py:341)");
    MH_CHECK_ARG(S >= 1 && S <= 32 * MAX_SPL, "num_sample must be in [1,128]");
    MH_CHECK_ARG((vw->P & 1) && vw->P >= 1, "patch size must be odd");
    MH_CHECK_ARG((dbg_base_idx == nullptr) == (dbg_base_val == nullptr), "dbg_base_idx/val must come together");
    MH_CHECK_ARG((dbg_loss_b == nullptr) == (dbg_arg_b == nullptr), "dbg_loss_b/arg_b must come together");
    const size_t smem = smem_layout(nullptr, nullptr, vw->V, S, vw->P * vw->P);
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if ((int64_t)smem > (int64_t)max_smem) {
        mh_set_error("mh_pmvo_optimize: V*P*P = %d patch entries need %zu B of shared memory (> %d B available)",
                     vw->V * vw->P * vw->P, smem, max_smem);
        return 1;
    }
    OptKernel kern = pick_kernel(S, vw->V);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mh_set_error("mh_pmvo_optimize: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 2; }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, OPT_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)mh_sm_count() * per_sm;
    if (grid > N) grid = N;
    cudaMemsetAsync(workspace, 0, 8, (cudaStream_t)stream);
    kern<<<(unsigned)grid, OPT_THREADS, smem, (cudaStream_t)stream>>>(
        *vw, points, N, offsets, S, conf_threshold, ori, loss, high_conf, dbg_base_idx, dbg_base_val,
        dbg_best_sample, dbg_loss_b, dbg_arg_b, reinterpret_cast<unsigned long long*>(workspace));
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
