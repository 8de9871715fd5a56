// Refine stage kernels (PMVO.refine module function, PMVO.py:602-686):
//   mh_pmvo_refine_loss   PMVO.refine's single-sample reprojection loss (PMVO.py:81-93)
//   mh_knn                scipy.spatial.KDTree.query(k) replacement: exact kNN on a uniform grid
//   mh_nn_dist            scalp_tree.query(points, k=1) distance (PMVO.py:104)
//   mh_medoid_gather      compute_points_similarity on gathered neighbours (PMVO_utils.py:366-382)
#include "mh_common.cuh"
#include "mh_torch_sum.cuh"

namespace {

// ------------------------------------------------------------------------------------------ refine loss
// RL_GROUP (64) threads per point, one view per thread; each thread scans its view's PxP patch straight from the
// resident map with one patch row of loads in flight at a time.  With a single sample "low_conf_index" is always
// true (sum over 1 sample < 5, PMVO.py:199) so the result is sum(l*w)/sum(w) over the views in torch.sum's cascade
// order (done by the group's first thread from shared memory).  Launches are small (5000-point chunks of the
// sequential refine loop), so the mapping favours latency: 64-way parallelism per point.
constexpr int RL_GROUP = 64, RL_PTS = 4;

template <int PT>   // PT = patch size known at compile time (0 = generic, <= 17)
__global__ void __launch_bounds__(RL_GROUP * RL_PTS)
refine_loss_kernel(mh_views vw, const float* __restrict__ pts, const float* __restrict__ dir, int64_t N,
                   float thr_c, float* __restrict__ out_loss) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MhCam* cams = reinterpret_cast<MhCam*>(smem_raw);
    float* lw = reinterpret_cast<float*>(cams + vw.V);          // [RL_PTS][V][2]
    const int V = vw.V, P = PT ? PT : vw.P, half = P / 2;
    const int grp = threadIdx.x / RL_GROUP, gl = threadIdx.x % RL_GROUP;
    for (int i = threadIdx.x; i < V * MH_CAM_STRIDE; i += blockDim.x) reinterpret_cast<float*>(cams)[i] = vw.cam[i];
    __syncthreads();
    float* my = lw + (size_t)grp * V * 2;
    const float Wf = (float)vw.W, Hf = (float)vw.H;
    const float4* __restrict__ mapC = reinterpret_cast<const float4*>(vw.mapC);
    const float4* __restrict__ mapP = reinterpret_cast<const float4*>(vw.mapP);
    const size_t plane = (size_t)vw.H * vw.W;
    constexpr int MAXP = PT ? PT : 17;
    for (int64_t base_n = (int64_t)blockIdx.x * RL_PTS; base_n < N; base_n += (int64_t)gridDim.x * RL_PTS) {
        const int64_t n = base_n + grp;
        const bool live = n < N;
        float px = 0, py = 0, pz = 0, qx = 0, qy = 0, qz = 0;
        if (live) {
            px = pts[3 * n]; py = pts[3 * n + 1]; pz = pts[3 * n + 2];
            // next = p + ori * 0.005 / 4   (PMVO.py:86)
            qx = px + dir[3 * n] * 0.005f / 4.0f; qy = py + dir[3 * n + 1] * 0.005f / 4.0f; qz = pz + dir[3 * n + 2] * 0.005f / 4.0f;
        }
        for (int v = gl; v < V && live; v += RL_GROUP) {
            const MhCam& cm = cams[v];
            float cx, cy, cz, xp, yp, xs, ys;
            mh_world_to_cam(cm.p, px, py, pz, cx, cy, cz);
            mh_cam_to_xy(cm.fx, cm.fy, cm.cx, cm.cy, Wf, Hf, cx, cy, cz, xp, yp);
            int row, col; bool oob;
            mh_round_clamp(xp, yp, vw.W, vw.H, row, col, oob);
            const float4* __restrict__ mp = mapP + (size_t)v * plane;
            const float4 dm = __ldg(mapC + (size_t)v * plane + (size_t)row * vw.W + col);
            float vis = mh_visible((-cz / 2.0f) * 255.0f, dm.x);
            if (oob) vis = -1.0f;
            float l_w = 0.0f, w = 0.0f;
            if (vis != -1.0f) {
                float c2x, c2y, c2z;
                mh_world_to_cam(cm.p, qx, qy, qz, c2x, c2y, c2z);
                mh_cam_to_xy(cm.fx, cm.fy, cm.cx, cm.cy, Wf, Hf, c2x, c2y, c2z, xs, ys);
                float y0, y1;
                mh_normalize2(ys - yp, xs - xp, y0, y1);
                const float cmax = fminf(fmaxf(dm.z, 1e-6f), 1.0f);     // PxP maximum: same texel as depth / mask
                const bool hi = cmax > thr_c;
                float bl = 0.0f, bc = 0.0f;
                for (int di = 0; di < P; ++di) {
                    const int r = min(max(row + di - half, 0), vw.H - 1);
                    float4 t[MAXP];
#pragma unroll
                    for (int dj = 0; dj < MAXP; ++dj)
                        if (dj < P) t[dj] = __ldg(mp + (size_t)r * vw.W + min(max(col + dj - half, 0), vw.W - 1));
#pragma unroll
                    for (int dj = 0; dj < MAXP; ++dj) {
                        if (dj < P) {
                            const float x0 = t[dj].x, x1 = t[dj].y;               // stored normalised (pmvo_views.cu)
                            const float cf = fminf(fmaxf(t[dj].z, 1e-6f), 1.0f);
                            const float l = 1.0f - fabsf(x0 * y0 + x1 * y1);
                            if (di == 0 && dj == 0) { bl = l; bc = cf; }
                            else if (l < bl && (!hi || cf > thr_c)) { bl = l; bc = cf; }
                        }
                    }
                }
                w = bc;
                l_w = bl * bc;
            }
            my[2 * v] = l_w;
            my[2 * v + 1] = w;
        }
        __syncthreads();
        if (gl == 0 && live) {
            MhCascade<2> acc;
            acc.init(V);
            for (int v = 0; v < V; ++v) {
                if (my[2 * v + 1] == 0.0f && my[2 * v] == 0.0f) continue;
                acc.begin_row(v);
                acc.add(0, my[2 * v]);
                acc.add(1, my[2 * v + 1]);
            }
            float sres[2];
            acc.finish(V, sres);
            out_loss[n] = sres[0] / sres[1];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------ kNN grid
struct Grid {
    double ox, oy, oz, h;       // origin and cell size
    int nx, ny, nz;
};

__device__ __forceinline__ int cell_of(double p, double o, double h, int n) {
    int c = (int)floor((p - o) / h);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void knn_count_kernel(Grid g, const float* __restrict__ ref, int64_t n, int* __restrict__ counts,
                                 int* __restrict__ cell_of_pt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = cell_of(ref[3 * i], g.ox, g.h, g.nx), cy = cell_of(ref[3 * i + 1], g.oy, g.h, g.ny),
        cz = cell_of(ref[3 * i + 2], g.oz, g.h, g.nz);
    int c = (cz * g.ny + cy) * g.nx + cx;
    cell_of_pt[i] = c;
    atomicAdd(counts + c, 1);
}

__global__ void knn_fill_kernel(const int* __restrict__ cell_of_pt, int64_t n, const int* __restrict__ starts,
                                int* __restrict__ cursor, int* __restrict__ items) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of_pt[i];
    items[starts[c] + atomicAdd(cursor + c, 1)] = (int)i;
}

// ---- single-CTA-per-tile exclusive scan (3 kernels) ----
constexpr int SCAN_BLOCK = 1024, SCAN_ITEMS = 4;

__global__ void scan_local_kernel(const int* __restrict__ in, int* __restrict__ out, int64_t n, int* __restrict__ block_sums) {
    __shared__ int sh[SCAN_BLOCK];
    const int64_t base = ((int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
    int v[SCAN_ITEMS], tot = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; tot += v[k]; }
    sh[threadIdx.x] = tot;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
        int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    int run = sh[threadIdx.x] - tot;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = sh[threadIdx.x];
}
__global__ void scan_sums_kernel(int* __restrict__ block_sums, int nb, int* __restrict__ total) {
    // one CTA; sequential over chunks of blockDim
    __shared__ int sh[SCAN_BLOCK];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += SCAN_BLOCK) {
        int i = b0 + threadIdx.x;
        int x = (i < nb) ? block_sums[i] : 0;
        sh[threadIdx.x] = x;
        __syncthreads();
        for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
            int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) block_sums[i] = carry + sh[threadIdx.x] - x;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry += sh[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
__global__ void scan_add_kernel(int* __restrict__ out, int64_t n, const int* __restrict__ block_sums) {
    const int64_t base = ((int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
    const int add = block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) out[base + k] += add;
}

}  // namespace

// exclusive scan of n ints: out[i] = sum_{j<i} in[j]; out[n] = total.  scratch >= ceil(n/4096) ints.
int mh_exclusive_scan(cudaStream_t st, const int* in, int* out, int64_t n, int* scratch) {
    const int64_t per = (int64_t)SCAN_BLOCK * SCAN_ITEMS;
    const int nb = (int)((n + per - 1) / per);
    scan_local_kernel<<<nb, SCAN_BLOCK, 0, st>>>(in, out, n, scratch);
    MH_COUNT_LAUNCH();
    scan_sums_kernel<<<1, SCAN_BLOCK, 0, st>>>(scratch, nb, out + n);
    MH_COUNT_LAUNCH();
    scan_add_kernel<<<nb, SCAN_BLOCK, 0, st>>>(out, n, scratch);
    MH_COUNT_LAUNCH();
    return 0;
}

namespace {

// ---- query: one warp per query, candidate buffer of 256 (dist2, idx) kept in shared memory ----
constexpr int KNN_WARPS = 4, KNN_BUF = 256, KNN_MAXK = 128;

struct Cand { double d; int i; int pad; };

__device__ __forceinline__ bool cand_less(const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }

// bitonic sort of 256 candidates by one warp (8 per lane)
__device__ void warp_sort256(Cand* buf, int lane) {
    for (int k = 2; k <= KNN_BUF; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < KNN_BUF / 2; t += 32) {
                // element pair (i, i^j) with i having bit j clear
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int p = i | j;
                bool up = (i & k) == 0;
                Cand a = buf[i], b = buf[p];
                if (cand_less(b, a) == up) { buf[i] = b; buf[p] = a; }
            }
            __syncwarp();
        }
    }
}

template <typename QT>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_query_kernel(Grid g, const float* __restrict__ ref, const int* __restrict__ starts, const int* __restrict__ items,
                 const QT* __restrict__ query, int64_t nq, int K, int* __restrict__ out_idx,
                 const unsigned char* __restrict__ only_flagged) {
    __shared__ Cand bufs[KNN_WARPS][KNN_BUF];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Cand* buf = bufs[warp];
    const double INF = 1e300;
    for (int64_t q = (int64_t)blockIdx.x * KNN_WARPS + warp; q < nq; q += (int64_t)gridDim.x * KNN_WARPS) {
        if (only_flagged && !only_flagged[q]) continue;
        const double qx = query[3 * q], qy = query[3 * q + 1], qz = query[3 * q + 2];
        const int cx = cell_of(qx, g.ox, g.h, g.nx), cy = cell_of(qy, g.oy, g.h, g.ny), cz = cell_of(qz, g.oz, g.h, g.nz);
        for (int t = lane; t < KNN_BUF; t += 32) { buf[t].d = INF; buf[t].i = 0x7fffffff; }
        __syncwarp();
        int npend = 0;                 // pending candidates live in buf[K .. K+npend)
        double bound = INF;            // current k-th best distance
        const int cap = KNN_BUF - K;
        const int maxR = max(max(max(cx, g.nx - 1 - cx), max(cy, g.ny - 1 - cy)), max(cz, g.nz - 1 - cz));
        for (int R = 0; R <= maxR; ++R) {
            // cells on the Chebyshev shell of radius R, clipped to the grid
            const int z0 = max(cz - R, 0), z1 = min(cz + R, g.nz - 1);
            const int y0 = max(cy - R, 0), y1 = min(cy + R, g.ny - 1);
            const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    const bool edge_zy = (abs(z - cz) == R) || (abs(y - cy) == R);
                    // on an edge row every x is on the shell (contiguous cells -> contiguous items);
                    // otherwise only x = cx-R and cx+R.
                    for (int seg = 0; seg < (edge_zy ? 1 : 2); ++seg) {
                        int xa, xb;
                        if (edge_zy) { xa = x0; xb = x1; }
                        else {
                            int xx = seg == 0 ? cx - R : cx + R;
                            if (xx < 0 || xx >= g.nx || (seg == 1 && R == 0)) continue;
                            xa = xb = xx;
                        }
                        const int rowc = (z * g.ny + y) * g.nx;
                        const int s = starts[rowc + xa], e = starts[rowc + xb + 1];
                        for (int it0 = s; it0 < e; it0 += 32) {
                            const int it = it0 + lane;
                            bool ok = false;
                            Cand c; c.d = INF; c.i = 0; c.pad = 0;
                            if (it < e) {
                                const int ri = items[it];
                                const double dx = qx - (double)ref[3 * ri], dy = qy - (double)ref[3 * ri + 1], dz = qz - (double)ref[3 * ri + 2];
                                c.d = dx * dx + dy * dy + dz * dz;
                                c.i = ri;
                                ok = c.d <= bound;
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, ok);
                            const int add = __popc(m);
                            if (npend + add > cap) {
                                warp_sort256(buf, lane);
                                for (int t = K + lane; t < KNN_BUF; t += 32) { buf[t].d = INF; buf[t].i = 0x7fffffff; }
                                __syncwarp();
                                bound = buf[K - 1].d;
                                npend = 0;
                                ok = ok && c.d <= bound;
                            }
                            const unsigned m2 = __ballot_sync(0xffffffffu, ok);
                            if (ok) buf[K + npend + __popc(m2 & ((1u << lane) - 1))] = c;
                            npend += __popc(m2);
                            __syncwarp();
                        }
                    }
                }
            // termination: every unvisited point lies outside the (2R+1)^3 block around the query cell
            if (npend > 0) {
                warp_sort256(buf, lane);
                for (int t = K + lane; t < KNN_BUF; t += 32) { buf[t].d = INF; buf[t].i = 0x7fffffff; }
                __syncwarp();
                bound = buf[K - 1].d;
                npend = 0;
            }
            double margin = INF;
            if (cx - R > 0) margin = fmin(margin, qx - (g.ox + (double)(cx - R) * g.h));
            if (cx + R < g.nx - 1) margin = fmin(margin, (g.ox + (double)(cx + R + 1) * g.h) - qx);
            if (cy - R > 0) margin = fmin(margin, qy - (g.oy + (double)(cy - R) * g.h));
            if (cy + R < g.ny - 1) margin = fmin(margin, (g.oy + (double)(cy + R + 1) * g.h) - qy);
            if (cz - R > 0) margin = fmin(margin, qz - (g.oz + (double)(cz - R) * g.h));
            if (cz + R < g.nz - 1) margin = fmin(margin, (g.oz + (double)(cz + R + 1) * g.h) - qz);
            if (margin > 0 && bound < margin * margin) break;
        }
        for (int t = lane; t < K; t += 32) out_idx[q * K + t] = buf[t].i;
        __syncwarp();
    }
}


// ---- fast path: gather ring candidates once, pick a radius by bisection, sort <= 256 survivors once ----------
// The cell size is ~ the k-NN radius, so rings 0..1 usually hold all k neighbours.  Candidates (squared float64
// distance, index) of the visited rings are kept in shared memory (KF_CAP per warp).  A bisection on the squared
// radius finds U >= d_k with at most KF_SORT candidates inside; survivors are compacted, and the search ends when the
// visited block contains the ball of radius sqrt(U).  One bitonic sort orders the survivors by (distance, index).
// Queries that overflow the buffer or cannot be separated (hundreds of equidistant points) are flagged and redone
// by the general kernel above.
constexpr int KF_WARPS = 4, KF_CAP = 768, KF_SORT = 128;   // survivors sorted once (128: one 28-stage bitonic network)

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename QT>
__global__ void __launch_bounds__(KF_WARPS * 32)
knn_query_fast_kernel(Grid g, const float* __restrict__ ref, const int* __restrict__ starts, const int* __restrict__ items,
                      const QT* __restrict__ query, int64_t nq, int K, int* __restrict__ out_idx,
                      unsigned char* __restrict__ flags) {
    __shared__ double sd_all[KF_WARPS][KF_CAP];
    __shared__ int si_all[KF_WARPS][KF_CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sd = sd_all[warp];
    int* si = si_all[warp];
    const double INF = 1e300;
    for (int64_t q = (int64_t)blockIdx.x * KF_WARPS + warp; q < nq; q += (int64_t)gridDim.x * KF_WARPS) {
        const double qx = query[3 * q], qy = query[3 * q + 1], qz = query[3 * q + 2];
        const int cx = cell_of(qx, g.ox, g.h, g.nx), cy = cell_of(qy, g.oy, g.h, g.ny), cz = cell_of(qz, g.oz, g.h, g.nz);
        const int maxR = max(max(max(cx, g.nx - 1 - cx), max(cy, g.ny - 1 - cy)), max(cz, g.nz - 1 - cz));
        int nc = 0;
        double U = INF;
        bool fail = false, done = false;
        for (int R = 0; R <= maxR && !fail && !done; ++R) {
            // The shell of radius R as a flat list of item ranges: every (z,y) row gets two slots -- an edge row
            // (|dz| == R or |dy| == R) is one contiguous x-range, an interior row contributes its two end cells.
            // Lanes fetch the ranges of 32 slots at once, a warp scan lays them out, and the candidates are then
            // gathered four per lane per round, so the dependent loads (range -> item -> coordinates) are paid per
            // batch of 128 candidates instead of per row.
            const int side = 2 * R + 1, nslots = 2 * side * side;
            for (int slot0 = 0; slot0 < nslots && !fail; slot0 += 32) {
                const int slot = slot0 + lane;
                int s_beg = 0, s_len = 0;
                if (slot < nslots) {
                    const int row = slot >> 1, half2 = slot & 1;
                    const int z = cz - R + row / side, y = cy - R + row % side;
                    if (z >= 0 && z < g.nz && y >= 0 && y < g.ny) {
                        const bool edge_zy = (abs(z - cz) == R) || (abs(y - cy) == R);
                        int xa = 0, xb = -1;
                        if (edge_zy) { if (half2 == 0) { xa = max(cx - R, 0); xb = min(cx + R, g.nx - 1); } }
                        else {
                            const int xx = half2 == 0 ? cx - R : cx + R;
                            if (xx >= 0 && xx < g.nx && !(half2 == 1 && R == 0)) xa = xb = xx;
                        }
                        if (xb >= xa) {
                            const int rowc = (z * g.ny + y) * g.nx;
                            s_beg = starts[rowc + xa];
                            s_len = starts[rowc + xb + 1] - s_beg;
                        }
                    }
                }
                int incl = s_len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                const int excl = incl - s_len;
                for (int c0 = 0; c0 < total && !fail; c0 += 128) {
                    double d4[4]; int r4[4]; bool ok4[4];
                    int it4[4];
#pragma unroll
                    for (int u2 = 0; u2 < 4; ++u2) {
                        const int c = c0 + lane + 32 * u2;
                        // segment of candidate c: the last slot whose exclusive offset is <= c (5-step search via shuffles)
                        int lo = 0;
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1) {
                            const int probe = lo + step;
                            const int pe = __shfl_sync(0xffffffffu, excl, probe & 31);
                            if (probe < 32 && pe <= c) lo = probe;
                        }
                        // skip empty slots that share the same offset: take the slot whose range really contains c
                        const int sb = __shfl_sync(0xffffffffu, s_beg, lo), se = __shfl_sync(0xffffffffu, excl, lo);
                        it4[u2] = (c < total) ? sb + (c - se) : -1;
                    }
#pragma unroll
                    for (int u2 = 0; u2 < 4; ++u2) r4[u2] = it4[u2] >= 0 ? items[it4[u2]] : -1;
#pragma unroll
                    for (int u2 = 0; u2 < 4; ++u2) {
                        d4[u2] = INF; ok4[u2] = false;
                        if (r4[u2] >= 0) {
                            const int ri = r4[u2];
                            const double dx = qx - (double)ref[3 * ri], dy = qy - (double)ref[3 * ri + 1], dz = qz - (double)ref[3 * ri + 2];
                            d4[u2] = dx * dx + dy * dy + dz * dz;
                            ok4[u2] = d4[u2] <= U;
                        }
                    }
#pragma unroll
                    for (int u2 = 0; u2 < 4; ++u2) {
                        const unsigned m = __ballot_sync(0xffffffffu, ok4[u2]);
                        if (nc + __popc(m) > KF_CAP) { fail = true; break; }
                        if (ok4[u2]) { const int sl = nc + __popc(m & ((1u << lane) - 1)); sd[sl] = d4[u2]; si[sl] = r4[u2]; }
                        nc += __popc(m);
                    }
                }
            }
            if (fail) break;
            if ((R == 0 && maxR > 0) || nc < K) continue;
            __syncwarp();
            // bisection for U: count(d <= U) in [K, KF_SORT]
            double hi = 0.0;
            for (int t = lane; t < nc; t += 32) hi = fmax(hi, sd[t]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            int chi = nc;
            double lo = 0.0;
            for (int iter = 0; iter < 60 && chi > KF_SORT; ++iter) {
                const double mid = 0.5 * (lo + hi);
                int c = 0;
                for (int t = lane; t < nc; t += 32) c += (sd[t] <= mid) ? 1 : 0;
                c = warp_sum(c);
                if (c >= K) { hi = mid; chi = c; } else lo = mid;
            }
            if (chi > KF_SORT) { fail = true; break; }
            U = hi;
            // compact survivors in place
            if (chi < nc) {
                int base = 0;
                for (int t0 = 0; t0 < nc; t0 += 32) {
                    const int t = t0 + lane;
                    double d = INF; int ri = 0;
                    if (t < nc) { d = sd[t]; ri = si[t]; }
                    const bool keep = t < nc && d <= U;
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    __syncwarp();
                    if (keep) { const int slot = base + __popc(m & ((1u << lane) - 1)); sd[slot] = d; si[slot] = ri; }
                    base += __popc(m);
                    __syncwarp();
                }
                nc = base;
            }
            double margin = INF;
            if (cx - R > 0) margin = fmin(margin, qx - (g.ox + (double)(cx - R) * g.h));
            if (cx + R < g.nx - 1) margin = fmin(margin, (g.ox + (double)(cx + R + 1) * g.h) - qx);
            if (cy - R > 0) margin = fmin(margin, qy - (g.oy + (double)(cy - R) * g.h));
            if (cy + R < g.ny - 1) margin = fmin(margin, (g.oy + (double)(cy + R + 1) * g.h) - qy);
            if (cz - R > 0) margin = fmin(margin, qz - (g.oz + (double)(cz - R) * g.h));
            if (cz + R < g.nz - 1) margin = fmin(margin, (g.oz + (double)(cz + R + 1) * g.h) - qz);
            if (margin > 0 && U < margin * margin) done = true;
        }
        if (fail || !done || nc < K || nc > KF_SORT) {
            if (lane == 0) flags[q] = 1;
            __syncwarp();
            continue;
        }
        // one bitonic sort of the survivors by (distance, index), padded to a power of two
        const int n2 = nc <= 128 ? 128 : 256;
        for (int t = nc + lane; t < n2; t += 32) { sd[t] = INF; si[t] = 0x7fffffff; }
        __syncwarp();
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < n2 / 2; t += 32) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int p = i | j;
                    const bool up = (i & k) == 0;
                    const double da = sd[i], db = sd[p];
                    const int ia = si[i], ib = si[p];
                    const bool b_lt_a = db < da || (db == da && ib < ia);
                    if (b_lt_a == up) { sd[i] = db; sd[p] = da; si[i] = ib; si[p] = ia; }
                }
                __syncwarp();
            }
        for (int t = lane; t < K; t += 32) out_idx[q * K + t] = si[t];
        if (lane == 0) flags[q] = 0;
        __syncwarp();
    }
}

__global__ void nn_dist_kernel(const double* __restrict__ ref, int64_t nref, const float* __restrict__ query,
                               int64_t nq, double* __restrict__ dist) {
    // brute force, one thread per query, reference tiles through shared memory
    __shared__ double tile[256 * 3];
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double qx = 0, qy = 0, qz = 0, best = 1e300;
    if (q < nq) { qx = query[3 * q]; qy = query[3 * q + 1]; qz = query[3 * q + 2]; }
    for (int64_t r0 = 0; r0 < nref; r0 += 256) {
        const int nt = (int)min((int64_t)256, nref - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < nt * 3; i += blockDim.x) tile[i] = ref[3 * r0 + i];
        __syncthreads();
        for (int i = 0; i < nt; ++i) {
            const double dx = qx - tile[3 * i], dy = qy - tile[3 * i + 1], dz = qz - tile[3 * i + 2];
            best = fmin(best, dx * dx + dy * dy + dz * dz);
        }
    }
    if (q < nq) dist[q] = sqrt(best);
}

// ------------------------------------------------------------------------------------------ medoid
constexpr int MED_WARPS = 4;

__global__ void __launch_bounds__(MED_WARPS * 32)
medoid_gather_kernel(const float* __restrict__ ori, const int* __restrict__ nbr, int64_t n, int K,
                     float* __restrict__ out, int* __restrict__ out_k) {
    extern __shared__ float4 med_u[];                   // [MED_WARPS][K] unit directions (one 16 B load per term)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* u = med_u + (size_t)warp * K;
    for (int64_t i = (int64_t)blockIdx.x * MED_WARPS + warp; i < n; i += (int64_t)gridDim.x * MED_WARPS) {
        for (int k = lane; k < K; k += 32) {
            const int r = nbr[i * K + k];
            const float a = ori[3 * r], b = ori[3 * r + 1], c = ori[3 * r + 2];
            const float nn = fmaxf(mh_norm3(a, b, c), 1e-8f);
            u[k] = make_float4(a / nn, b / nn, c / nn, 0.0f);
        }
        __syncwarp();
        float best = -1e30f; int bk = 0x7fffffff;
        const int K32 = K & ~31, left = K - K32;
        const bool split = left >= 1 && left <= 4 && K >= 8;            // the last rows: 8 lanes each (mh_torch_sum.cuh)
        for (int k = lane; k < (split ? K32 : K); k += 32) {
            const float4 w = u[k];
            float s = mh_torch_inner_sum(K, [&](int j) { const float4 v = u[j]; return fabsf((w.x * v.x + w.y * v.y) + w.z * v.z); });
            s = s / (float)K;
            if (s > best) { best = s; bk = k; }        // ascending k: first maximum kept
        }
        if (split) {
            const int k = min(K32 + (lane >> 3), K - 1);
            const float4 w = u[k];
            float s = mh_torch_inner_sum_split8(K, lane, [&](int j) { const float4 v = u[j]; return fabsf((w.x * v.x + w.y * v.y) + w.z * v.z); });
            s = s / (float)K;
            if ((lane >> 3) < left && s > best) { best = s; bk = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
        }
        if (lane == 0) {
            const int r = nbr[i * K + bk];
            out[3 * i] = ori[3 * r]; out[3 * i + 1] = ori[3 * r + 1]; out[3 * i + 2] = ori[3 * r + 2];
            if (out_k) out_k[i] = bk;
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int mh_pmvo_refine_loss(void* stream, const mh_views* vw, const float* points, const float* dir,
                                   int64_t N, float conf_threshold, float* loss) {
    MH_CHECK_ARG(vw && vw->mapC && vw->mapP && vw->cam, "null views");
    if (N == 0) return 0;
    MH_CHECK_ARG(points && dir && loss && N > 0, "bad arguments");
    const size_t smem = sizeof(MhCam) * vw->V + sizeof(float) * 2 * vw->V * RL_PTS;
    MH_CHECK_ARG(smem <= 200 * 1024, "too many views");
    MH_CHECK_ARG(vw->P >= 1 && vw->P <= 17 && (vw->P & 1), "patch size must be odd and <= 17");
    auto kern = vw->P == 7 ? refine_loss_kernel<7> : vw->P == 5 ? refine_loss_kernel<5> : vw->P == 9 ? refine_loss_kernel<9>
                                                                                                    : refine_loss_kernel<0>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // One block per RL_PTS consecutive points, no grid-stride loop: blocks are dispatched in index order, so the points
    // in flight at any time are a contiguous (= spatially coherent) window whose patches stay L2-resident.  A capped,
    // strided grid lets the resident blocks drift apart over the whole point set and sent ~11 kB per point to DRAM.
    const int64_t blocks = (N + RL_PTS - 1) / RL_PTS;
    MH_CHECK_ARG(blocks < (1ll << 31), "too many points for one launch");
    kern<<<(unsigned)blocks, RL_GROUP * RL_PTS, smem, (cudaStream_t)stream>>>(*vw, points, dir, N, conf_threshold, loss);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

// workspace layout for kNN: [grid header 64 B][counts ncell+1][starts ncell+1][cursor ncell][cell_of_pt n][items n][scan scratch]
static const int64_t KNN_MAX_CELLS = 1ll << 24;

extern "C" int64_t mh_knn_workspace_bytes(int64_t n_ref, int64_t n_query, int32_t k) {
    (void)n_query; (void)k;
    return 256 + 4 * (3 * (KNN_MAX_CELLS + 2) + 2 * n_ref + KNN_MAX_CELLS / 4096 + 16) + ((n_query + 15) / 16) * 16;
}

namespace {
template <typename QT>
int knn_run(void* stream, const float* ref, int64_t n_ref, const QT* query, int64_t n_query, int32_t k,
            const double* bbox_host, double cell_size, int32_t* idx, void* workspace, int64_t workspace_bytes) {
    MH_CHECK_ARG(ref && query && idx && workspace && bbox_host, "null pointer");
    MH_CHECK_ARG(k >= 1 && k <= KNN_MAXK && k <= n_ref, "k must be in [1,128] and <= n_ref");
    MH_CHECK_ARG(workspace_bytes >= mh_knn_workspace_bytes(n_ref, n_query, k), "workspace too small");
    MH_CHECK_ARG(cell_size > 0, "cell size must be > 0");
    if (n_query == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    double hdr[7] = {bbox_host[0], bbox_host[1], bbox_host[2], bbox_host[3], bbox_host[4], bbox_host[5], cell_size};
    Grid g;
    g.h = hdr[6];
    g.ox = hdr[0]; g.oy = hdr[1]; g.oz = hdr[2];
    auto dim = [&](double lo, double hi) { double d = floor((hi - lo) / g.h) + 1; return (int)fmin(fmax(d, 1.0), 1024.0); };
    g.nx = dim(hdr[0], hdr[3]); g.ny = dim(hdr[1], hdr[4]); g.nz = dim(hdr[2], hdr[5]);
    while ((int64_t)g.nx * g.ny * g.nz > KNN_MAX_CELLS) {      // coarsen
        g.h *= 1.26;
        g.nx = dim(hdr[0], hdr[3]); g.ny = dim(hdr[1], hdr[4]); g.nz = dim(hdr[2], hdr[5]);
    }
    const int64_t ncell = (int64_t)g.nx * g.ny * g.nz;
    int* base = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + 256);
    int* counts = base;
    int* starts = counts + (KNN_MAX_CELLS + 2);
    int* cursor = starts + (KNN_MAX_CELLS + 2);
    int* cell_of_pt = cursor + (KNN_MAX_CELLS + 2);
    int* items = cell_of_pt + n_ref;
    int* scratch = items + n_ref;
    unsigned char* flags = reinterpret_cast<unsigned char*>(scratch + (KNN_MAX_CELLS / 4096 + 16));
    cudaMemsetAsync(counts, 0, sizeof(int) * (ncell + 1), st);
    cudaMemsetAsync(cursor, 0, sizeof(int) * ncell, st);
    knn_count_kernel<<<(unsigned)((n_ref + 255) / 256), 256, 0, st>>>(g, ref, n_ref, counts, cell_of_pt);
    MH_COUNT_LAUNCH();
    mh_exclusive_scan(st, counts, starts, ncell, scratch);
    knn_fill_kernel<<<(unsigned)((n_ref + 255) / 256), 256, 0, st>>>(cell_of_pt, n_ref, starts, cursor, items);
    MH_COUNT_LAUNCH();
    int64_t blocks = (n_query + KNN_WARPS - 1) / KNN_WARPS;
    const int64_t cap = (int64_t)mh_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    knn_query_fast_kernel<QT><<<(unsigned)blocks, KF_WARPS * 32, 0, st>>>(g, ref, starts, items, query, n_query, k, idx, flags);
    MH_COUNT_LAUNCH();
    knn_query_kernel<QT><<<(unsigned)blocks, KNN_WARPS * 32, 0, st>>>(g, ref, starts, items, query, n_query, k, idx, flags);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
}  // namespace

extern "C" int mh_knn(void* stream, const float* ref, int64_t n_ref, const float* query, int64_t n_query, int32_t k,
                      const double* bbox_host, double cell_size, int32_t* idx, void* workspace, int64_t workspace_bytes) {
    return knn_run<float>(stream, ref, n_ref, query, n_query, k, bbox_host, cell_size, idx, workspace, workspace_bytes);
}

// float64 queries against float32 references: what scipy's KDTree(select_points).query(filter_unvisible_points) sees in
// PMVO.refine step (iii) (PMVO.py:660-671: the .npy candidates are float64 and are cast to float32 only AFTER the query)
extern "C" int mh_knn_q64(void* stream, const float* ref, int64_t n_ref, const double* query, int64_t n_query, int32_t k,
                          const double* bbox_host, double cell_size, int32_t* idx, void* workspace, int64_t workspace_bytes) {
    return knn_run<double>(stream, ref, n_ref, query, n_query, k, bbox_host, cell_size, idx, workspace, workspace_bytes);
}

extern "C" int mh_nn_dist(void* stream, const double* ref, int64_t n_ref, const float* query, int64_t n_query, double* dist) {
    MH_CHECK_ARG(ref && query && dist && n_ref >= 1, "bad arguments");
    if (n_query == 0) return 0;
    nn_dist_kernel<<<(unsigned)((n_query + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ref, n_ref, query, n_query, dist);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_medoid_gather(void* stream, const float* ori, const int32_t* nbr, int64_t n, int32_t K,
                                float* out, int32_t* out_k) {
    MH_CHECK_ARG(ori && nbr && out && K >= 1 && K <= 1024, "bad arguments");
    if (n == 0) return 0;
    const size_t smem = sizeof(float4) * K * MED_WARPS;
    int64_t blocks = (n + MED_WARPS - 1) / MED_WARPS;
    const int64_t cap = (int64_t)mh_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    medoid_gather_kernel<<<(unsigned)blocks, MED_WARPS * 32, smem, (cudaStream_t)stream>>>(ori, nbr, n, K, out, out_k);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

// ---- per-chunk in-place update of PMVO.refine step (i) (PMVO.py:629-641) ---------------------------------
namespace {
__global__ void refine_update_kernel(const float* __restrict__ center, const float* __restrict__ upd_loss,
                                     const uint8_t* __restrict__ head_filter, int64_t n, float* __restrict__ ori,
                                     float* __restrict__ loss) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float l = head_filter[i] ? -1.0f : upd_loss[i];              // PMVO.py:92
    if (l == -1.0f) l = 0.5f;                                    // :639
    loss[i] = l;
    const float c0 = center[3 * i], c1 = center[3 * i + 1], c2 = center[3 * i + 2];
    const float o0 = ori[3 * i], o1 = ori[3 * i + 1], o2 = ori[3 * i + 2];
    const float nc = fmaxf(mh_norm3(c0, c1, c2), 1e-8f), no = fmaxf(mh_norm3(o0, o1, o2), 1e-8f);
    const float sim = fabsf(((c0 / nc) * (o0 / no) + (c1 / nc) * (o1 / no)) + (c2 / nc) * (o2 / no));   // :631-633
    if (sim < 0.95f) { ori[3 * i] = c0; ori[3 * i + 1] = c1; ori[3 * i + 2] = c2; }                       // :634-636
}
}  // namespace

extern "C" int mh_refine_update(void* stream, const float* center, const float* upd_loss, const uint8_t* head_filter,
                                int64_t n, float* ori, float* loss) {
    MH_CHECK_ARG(center && upd_loss && head_filter && ori && loss && n >= 0, "bad arguments");
    if (n == 0) return 0;
    refine_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(center, upd_loss, head_filter, n, ori, loss);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}


// ---- the whole chunk-sequential pass of PMVO.refine step (i) in one call (PMVO.py:608-641) ------------------
// Reference semantics: for each chunk of `sub_num` points, IN ORDER: centre = medoid of the neighbours' CURRENT
// orientations; loss = single-sample re-score of (point, centre); ori <- centre where |cos(centre, ori)| < 0.95.
// Later chunks gather orientations already updated by earlier chunks (Gauss-Seidel across chunks, Jacobi within).
//
// Dependency analysis: the re-score of chunk c reads only (points, centre_c) and feeds nothing back into later
// chunks; only the orientation update does.  So the sequential part is the medoid + update chain alone:
//   1. refine_sweep_kernel: ONE persistent launch.  CTAs draw point tickets in index order.  A point of chunk c gathers
//      neighbour j from `ori_new` if j's chunk is earlier than c and from the untouched input otherwise -- exactly the
//      values the reference's in-place array holds when it processes chunk c.  `ori_new` starts filled with a PENDING
//      bit pattern; a gather that meets it spins on that word until the owner has stored the final value (every word
//      goes pending -> final exactly once, so three non-pending words ARE the final triple: no flag, no fence, no extra
//      round trip).  The dependency is per point, not per chunk: nothing drains at a chunk boundary, only the rare
//      neighbour that is still in flight is waited for.  Tickets are drawn in order and a CTA only ever waits on earlier
//      tickets, which are finished or held by running CTAs: no deadlock for any grid size.
//      Four warps share one point so that a wait costs a quarter of a warp-per-point medoid.
//   2. one mh_pmvo_refine_loss launch over all n points with the stored centres;
//   3. refine_finish_kernel: loss rules (head filter / -1 -> 0.5, PMVO.py:92, :639).
namespace {
constexpr int SW_THREADS = 128;                       // one CTA per point: the chunk hand-over costs one point's latency

// PENDING: an all-ones NaN.  Arithmetic produces the canonical 0x7fffffff, never this pattern; a stored value that
// carries it anyway (it can only arrive as an input NaN payload) is written as 0xfffffffe, still a NaN.
constexpr unsigned SW_PENDING = 0xffffffffu;

MH_D unsigned sw_ld(const float* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
MH_D void sw_st(float* p, float x) {
    unsigned v = __float_as_uint(x);
    if (v == SW_PENDING) v = 0xfffffffeu;
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// final value of ori_new[3*rr .. 3*rr+2] (waits while its owner is still in flight)
MH_D void sw_wait3(const float* p, float& a, float& b, float& c) {
    unsigned ua = sw_ld(p), ub = sw_ld(p + 1), uc = sw_ld(p + 2);
    while (ua == SW_PENDING || ub == SW_PENDING || uc == SW_PENDING) {
        __nanosleep(64);
        ua = sw_ld(p); ub = sw_ld(p + 1); uc = sw_ld(p + 2);
    }
    a = __uint_as_float(ua); b = __uint_as_float(ub); c = __uint_as_float(uc);
}

// ctl: [0] next ticket
__global__ void __launch_bounds__(SW_THREADS)
refine_sweep_kernel(const float* __restrict__ ori_old, float* ori_new, const int* __restrict__ nbr, int64_t n, int K,
                    int sub_num, float* __restrict__ center, int* ctl) {
    extern __shared__ float4 sw_u[];                     // [K] unit directions (w unused)
    __shared__ int s_ticket;
    __shared__ float s_best[SW_THREADS / 32];
    __shared__ int s_bk[SW_THREADS / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (;;) {
        if (tid == 0) s_ticket = atomicAdd(ctl, 1);
        __syncthreads();
        const int t = s_ticket;
        if (t >= n) break;
        const int64_t i = t;
        const int first = (t / sub_num) * sub_num;       // neighbours below `first` belong to earlier chunks
        for (int k = tid; k < K; k += SW_THREADS) {
            const int rr = nbr[i * K + k];
            float a, b, cc;
            if (rr < first) {
                sw_wait3(ori_new + 3 * (int64_t)rr, a, b, cc);
            } else {
                const float* src = ori_old + 3 * (int64_t)rr;
                a = __ldcg(src); b = __ldcg(src + 1); cc = __ldcg(src + 2);
            }
            const float nn = fmaxf(mh_norm3(a, b, cc), 1e-8f);
            sw_u[k] = make_float4(a / nn, b / nn, cc / nn, 0.0f);
        }
        __syncthreads();
        float best = -1e30f; int bk = 0x7fffffff;
        const int K32 = K & ~31, left = K - K32;
        const bool split = left >= 1 && left <= 4 && K >= 8;            // the last rows: 8 lanes each (mh_torch_sum.cuh)
        for (int k = tid; k < (split ? K32 : K); k += SW_THREADS) {
            const float4 w = sw_u[k];
            float sm = mh_torch_inner_sum(K, [&](int j) { const float4 v = sw_u[j]; return fabsf((w.x * v.x + w.y * v.y) + w.z * v.z); });
            sm = sm / (float)K;
            if (sm > best) { best = sm; bk = k; }        // ascending k: first maximum kept
        }
        if (split && warp == (K32 >> 5) % (SW_THREADS / 32)) {          // the warp that would have had those rows
            const int k = min(K32 + (lane >> 3), K - 1);
            const float4 w = sw_u[k];
            float sm = mh_torch_inner_sum_split8(K, lane, [&](int j) { const float4 v = sw_u[j]; return fabsf((w.x * v.x + w.y * v.y) + w.z * v.z); });
            sm = sm / (float)K;
            if ((lane >> 3) < left && sm > best) { best = sm; bk = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
        }
        if (lane == 0) { s_best[warp] = best; s_bk[warp] = bk; }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int w = 1; w < SW_THREADS / 32; ++w) {
                const float ob = s_best[w]; const int ok = s_bk[w];
                if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
            }
            const int rr = nbr[i * K + bk];
            float c0, c1, c2;
            if (rr < first) {
                sw_wait3(ori_new + 3 * (int64_t)rr, c0, c1, c2);           // already final: the gather above saw it
            } else {
                const float* src = ori_old + 3 * (int64_t)rr;
                c0 = __ldcg(src); c1 = __ldcg(src + 1); c2 = __ldcg(src + 2);
            }
            center[3 * i] = c0; center[3 * i + 1] = c1; center[3 * i + 2] = c2;
            const float o0 = ori_old[3 * i], o1 = ori_old[3 * i + 1], o2 = ori_old[3 * i + 2];
            const float nc = fmaxf(mh_norm3(c0, c1, c2), 1e-8f), no = fmaxf(mh_norm3(o0, o1, o2), 1e-8f);
            const float sim = fabsf(((c0 / nc) * (o0 / no) + (c1 / nc) * (o1 / no)) + (c2 / nc) * (o2 / no));   // PMVO.py:631-633
            const bool upd = sim < 0.95f;                                                                      // :634-636
            sw_st(ori_new + 3 * i, upd ? c0 : o0); sw_st(ori_new + 3 * i + 1, upd ? c1 : o1); sw_st(ori_new + 3 * i + 2, upd ? c2 : o2);
        }
        // the next iteration's first barrier (after the ticket draw) orders the reuse of sw_u / s_best
    }
}

// ---- the same sweep spread over the GPUs of one node -------------------------------------------------------
// Rank r owns the points of every world-th block of SWD_BLOCK consecutive indices and handles them in increasing order.
// Every rank keeps a FULL copy of ori_new / center in symmetric memory (each rank can store into every peer's copy over
// NVLink).  A finished point is stored into ALL copies; a gather that meets the PENDING pattern spins on its OWN copy
// until the owner's store has landed -- the per-point hand-over of the single-GPU kernel, with the store fanned out over
// peer memory, so the exchange overlaps the medoid work point by point and no collective follows the kernel.
// No deadlock: every wait is on a smaller point index, and the globally smallest unfinished point is always the one its
// owner is working on.  A bounded spin turns a lost peer into an error flag instead of a hung GPU.
constexpr int SWD_BLOCK = 64;
constexpr int SWD_MAX_WORLD = 16;
struct SwdPeers {
    float* ori_new[SWD_MAX_WORLD];
    float* center[SWD_MAX_WORLD];
};

MH_D unsigned swd_ld(const float* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
MH_D void swd_st(float* p, float x) {
    unsigned v = __float_as_uint(x);
    if (v == SW_PENDING) v = 0xfffffffeu;
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
MH_D bool swd_wait3(const float* p, float& a, float& b, float& c, long long limit) {
    unsigned ua = swd_ld(p), ub = swd_ld(p + 1), uc = swd_ld(p + 2);
    long long spins = 0;
    while (ua == SW_PENDING || ub == SW_PENDING || uc == SW_PENDING) {
        if (++spins > limit) return false;
        __nanosleep(100);
        ua = swd_ld(p); ub = swd_ld(p + 1); uc = swd_ld(p + 2);
    }
    a = __uint_as_float(ua); b = __uint_as_float(ub); c = __uint_as_float(uc);
    return true;
}

// ctl: [0] next local ticket, [1] error flag (a wait ran out)
__global__ void __launch_bounds__(SW_THREADS)
refine_sweep_dist_kernel(const float* __restrict__ ori_old, SwdPeers peers, const int* __restrict__ nbr_local, int64_t n,
                         int64_t n_local, int K, int sub_num, int rank, int world, long long spin_limit, int* ctl) {
    extern __shared__ float4 sw_u[];
    __shared__ int s_ticket;
    __shared__ float s_best[SW_THREADS / 32];
    __shared__ int s_bk[SW_THREADS / 32];
    __shared__ float s_out[6];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* mine = peers.ori_new[rank];
    for (;;) {
        if (tid == 0) s_ticket = atomicAdd(ctl, 1);
        __syncthreads();
        const int q = s_ticket;
        if (q >= n_local) break;
        const int64_t i = ((int64_t)(q / SWD_BLOCK) * world + rank) * SWD_BLOCK + (q % SWD_BLOCK);
        const int first = (int)(i / sub_num) * sub_num;
        bool lost = false;
        for (int k = tid; k < K; k += SW_THREADS) {
            const int rr = nbr_local[(int64_t)q * K + k];
            float a = 0.0f, b = 0.0f, cc = 0.0f;
            if (rr < first) {
                lost = !swd_wait3(mine + 3 * (int64_t)rr, a, b, cc, spin_limit) || lost;
            } else {
                const float* src = ori_old + 3 * (int64_t)rr;
                a = __ldcg(src); b = __ldcg(src + 1); cc = __ldcg(src + 2);
            }
            const float nn = fmaxf(mh_norm3(a, b, cc), 1e-8f);
            sw_u[k] = make_float4(a / nn, b / nn, cc / nn, 0.0f);
        }
        if (lost) atomicExch(ctl + 1, 1);
        __syncthreads();
        float best = -1e30f; int bk = 0x7fffffff;
        const int K32 = K & ~31, left = K - K32;
        const bool split = left >= 1 && left <= 4 && K >= 8;            // the last rows: 8 lanes each (mh_torch_sum.cuh)
        for (int k = tid; k < (split ? K32 : K); k += SW_THREADS) {
            const float4 w = sw_u[k];
            float sm = mh_torch_inner_sum(K, [&](int j) { const float4 v = sw_u[j]; return fabsf((w.x * v.x + w.y * v.y) + w.z * v.z); });
            sm = sm / (float)K;
            if (sm > best) { best = sm; bk = k; }
        }
        if (split && warp == (K32 >> 5) % (SW_THREADS / 32)) {          // the warp that would have had those rows
            const int k = min(K32 + (lane >> 3), K - 1);
            const float4 w = sw_u[k];
            float sm = mh_torch_inner_sum_split8(K, lane, [&](int j) { const float4 v = sw_u[j]; return fabsf((w.x * v.x + w.y * v.y) + w.z * v.z); });
            sm = sm / (float)K;
            if ((lane >> 3) < left && sm > best) { best = sm; bk = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
        }
        if (lane == 0) { s_best[warp] = best; s_bk[warp] = bk; }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int w = 1; w < SW_THREADS / 32; ++w) {
                const float ob = s_best[w]; const int ok = s_bk[w];
                if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
            }
            const int rr = nbr_local[(int64_t)q * K + bk];
            float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
            if (rr < first) {
                if (!swd_wait3(mine + 3 * (int64_t)rr, c0, c1, c2, spin_limit)) atomicExch(ctl + 1, 1);
            } else {
                const float* src = ori_old + 3 * (int64_t)rr;
                c0 = __ldcg(src); c1 = __ldcg(src + 1); c2 = __ldcg(src + 2);
            }
            const float o0 = ori_old[3 * i], o1 = ori_old[3 * i + 1], o2 = ori_old[3 * i + 2];
            const float nc = fmaxf(mh_norm3(c0, c1, c2), 1e-8f), no = fmaxf(mh_norm3(o0, o1, o2), 1e-8f);
            const float sim = fabsf(((c0 / nc) * (o0 / no) + (c1 / nc) * (o1 / no)) + (c2 / nc) * (o2 / no));   // PMVO.py:631-633
            const bool upd = sim < 0.95f;                                                                      // :634-636
            s_out[0] = upd ? c0 : o0; s_out[1] = upd ? c1 : o1; s_out[2] = upd ? c2 : o2;
            s_out[3] = c0; s_out[4] = c1; s_out[5] = c2;
        }
        __syncthreads();
        // fan the result out: thread (w, c) stores component c into rank w's copies
        if (tid < 3 * world) {
            const int w = tid / 3, c = tid - 3 * w;
            swd_st(peers.ori_new[w] + 3 * i + c, s_out[c]);
            peers.center[w][3 * i + c] = s_out[3 + c];
        }
        // the next iteration's first barrier orders the reuse of sw_u / s_best / s_out
    }
}

__global__ void refine_finish_kernel(const float* __restrict__ upd_loss, const uint8_t* __restrict__ head_filter, int64_t n,
                                     float* __restrict__ loss) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float l = head_filter[i] ? -1.0f : upd_loss[i];              // PMVO.py:92
    if (l == -1.0f) l = 0.5f;                                    // :639
    loss[i] = l;
}
}  // namespace

// ---- the three steps as separate entry points (the multi-GPU host shards step 2 over ranks) ----------------------
extern "C" int64_t mh_refine_sweep_workspace_bytes(int64_t n, int64_t sub_num) {
    const int64_t chunks = sub_num > 0 ? (n + sub_num - 1) / sub_num : 0;
    return 4 * (16 + chunks + 16);
}

extern "C" int mh_refine_sweep(void* stream, const float* ori, const int32_t* nbr, int32_t K, int64_t n, int64_t sub_num,
                               float* ori_new, float* center, void* scratch, int64_t scratch_bytes) {
    MH_CHECK_ARG(ori && nbr && ori_new && center && scratch, "null pointer");
    MH_CHECK_ARG(sub_num > 0 && sub_num < (1ll << 30) && K >= 1 && K <= 1024 && n >= 0 && n < (1ll << 31) - 1, "bad arguments");
    MH_CHECK_ARG(scratch_bytes >= mh_refine_sweep_workspace_bytes(n, sub_num), "scratch too small");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t chunks = (n + sub_num - 1) / sub_num;
    int* ctl = reinterpret_cast<int*>(scratch);
    cudaMemsetAsync(ctl, 0, sizeof(int) * (16 + chunks), st);
    cudaMemsetAsync(ori_new, 0xff, sizeof(float) * 3 * (size_t)n, st);      // every word PENDING
    const size_t smem = sizeof(float4) * (size_t)K;
    cudaFuncSetAttribute(refine_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, refine_sweep_kernel, SW_THREADS, smem);
    MH_CHECK_ARG(per_sm >= 1, "K too large for the sweep kernel's shared memory");
    int64_t blocks = (int64_t)mh_sm_count() * per_sm;
    if (blocks > n) blocks = n;
    refine_sweep_kernel<<<(unsigned)blocks, SW_THREADS, smem, st>>>(ori, ori_new, nbr, n, K, (int)sub_num, center, ctl);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}


// Distributed sweep (one node, symmetric memory).  nbr_local: neighbour lists of THIS rank's points in its own order
// (mh_refine_sweep_dist_index gives the global index of local row q); peer_ori_new / peer_center: `world` device
// pointers (host array) to every rank's [n][3] copies, peer_*[rank] being the local one.  The caller fills every word of
// its ori_new copy with 0xffffffff and passes a cross-rank barrier BEFORE the launch, and another one after it (peers
// keep storing into the local copies until their own kernels end).  error_flag (device int32): set to 1 if a wait ran out.
// max_blocks > 0 caps the grid (all ranks' kernels must be resident at the same time: one kernel per GPU in production).
extern "C" int64_t mh_refine_sweep_dist_block(void) { return SWD_BLOCK; }

extern "C" int64_t mh_refine_sweep_dist_local_count(int64_t n, int32_t rank, int32_t world) {
    const int64_t stride = (int64_t)SWD_BLOCK * world;
    const int64_t full = n / stride, rem = n % stride;
    int64_t c = full * SWD_BLOCK;
    const int64_t lo = (int64_t)rank * SWD_BLOCK;
    if (rem > lo) c += (rem - lo < SWD_BLOCK) ? rem - lo : SWD_BLOCK;
    return c;
}

extern "C" int mh_refine_sweep_dist(void* stream, const float* ori, const int32_t* nbr_local, int32_t K, int64_t n,
                                    int64_t sub_num, int32_t rank, int32_t world, const uint64_t* peer_ori_new,
                                    const uint64_t* peer_center, double spin_seconds, int32_t max_blocks, void* scratch,
                                    int64_t scratch_bytes, int32_t* error_flag) {
    MH_CHECK_ARG(ori && peer_ori_new && peer_center && scratch && error_flag, "null pointer");
    MH_CHECK_ARG(world >= 1 && world <= SWD_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world (<= 16 ranks)");
    MH_CHECK_ARG(sub_num > 0 && sub_num < (1ll << 30) && K >= 1 && K <= 1024 && n >= 0 && n < (1ll << 31) - 1, "bad arguments");
    MH_CHECK_ARG(scratch_bytes >= 64, "scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_local = mh_refine_sweep_dist_local_count(n, rank, world);
    cudaMemsetAsync(error_flag, 0, sizeof(int32_t), st);
    if (n_local == 0) return 0;
    MH_CHECK_ARG(nbr_local, "null neighbour table");
    SwdPeers peers;
    for (int w = 0; w < SWD_MAX_WORLD; ++w) {
        peers.ori_new[w] = reinterpret_cast<float*>(w < world ? peer_ori_new[w] : 0);
        peers.center[w] = reinterpret_cast<float*>(w < world ? peer_center[w] : 0);
        MH_CHECK_ARG(w >= world || (peers.ori_new[w] && peers.center[w]), "null peer buffer");
    }
    int* ctl = reinterpret_cast<int*>(scratch);
    cudaMemsetAsync(ctl, 0, 64, st);
    const size_t smem = sizeof(float4) * (size_t)K;
    cudaFuncSetAttribute(refine_sweep_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, refine_sweep_dist_kernel, SW_THREADS, smem);
    MH_CHECK_ARG(per_sm >= 1, "K too large for the sweep kernel's shared memory");
    int64_t blocks = (int64_t)mh_sm_count() * per_sm;
    if (blocks > n_local) blocks = n_local;
    if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;      // tests: several "ranks" co-resident on one GPU
    const long long limit = (long long)((spin_seconds > 0 ? spin_seconds : 2.0) * 4.0e6);      // ~0.25 us per poll
    refine_sweep_dist_kernel<<<(unsigned)blocks, SW_THREADS, smem, st>>>(ori, peers, nbr_local, n, n_local, K, (int)sub_num,
                                                                         rank, world, limit, ctl);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    cudaMemcpyAsync(error_flag, ctl + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
    return 0;
}

extern "C" int mh_refine_finish(void* stream, const float* upd_loss, const uint8_t* head_filter, int64_t n, float* loss) {
    MH_CHECK_ARG(upd_loss && head_filter && loss && n >= 0, "bad arguments");
    if (n == 0) return 0;
    refine_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(upd_loss, head_filter, n, loss);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

// all three in one call.  scratch: [center n*3][ori_new n*3][upd n] floats, [sweep scratch]
extern "C" int64_t mh_refine_chunks_workspace_bytes(int64_t n, int64_t sub_num) {
    return 4 * 7 * n + mh_refine_sweep_workspace_bytes(n, sub_num);
}

extern "C" int mh_refine_chunks(void* stream, const mh_views* vw, const float* points, const int32_t* nbr, int32_t K,
                                const uint8_t* head_filter, int64_t n, int64_t sub_num, float conf_threshold,
                                float* ori, float* loss, void* scratch, int64_t scratch_bytes) {
    MH_CHECK_ARG(vw && points && nbr && head_filter && ori && loss && scratch, "null pointer");
    MH_CHECK_ARG(sub_num > 0 && n >= 0, "bad arguments");
    MH_CHECK_ARG(scratch_bytes >= mh_refine_chunks_workspace_bytes(n, sub_num), "scratch too small");
    if (n == 0) return 0;
    float* center = reinterpret_cast<float*>(scratch);
    float* ori_new = center + 3 * n;
    float* upd = ori_new + 3 * n;
    int rc = mh_refine_sweep(stream, ori, nbr, K, n, sub_num, ori_new, center, upd + n, mh_refine_sweep_workspace_bytes(n, sub_num));
    if (rc == 0) rc = mh_pmvo_refine_loss(stream, vw, points, center, n, conf_threshold, upd);
    if (rc == 0) rc = mh_refine_finish(stream, upd, head_filter, n, loss);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(ori, ori_new, sizeof(float) * 3 * n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { mh_set_error("mh_refine_chunks: %s", cudaGetErrorString(e)); return 2; }
    return 0;
}
