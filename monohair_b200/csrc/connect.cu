// HairGrow connect stage, device parts (HairGrowing.find_connect_info, HairGrow.py:436-505 and :548-584; the occupancy
// test of :510-534 and Utils/PMVO_utils.py:random_move_strands :618-658).
//
// The reference builds one scipy KDTree per strand plus two over the strand end points and walks the strands in a Python
// loop: for each end of each strand, the 50 nearest end points within connect_threshold (roots first, tips if that gives
// nothing), and per candidate the distance of every point of the strand to the candidate strand.  Everything is float64.
// Here: one warp per strand.  The end-point search is a float64 scan over all end points (S^2 x 4 distances: 4e10 for
// 100 k strands, ~10 ms of FP64 on a B200 -- no tree, no radius heuristics, exact), candidates kept in shared memory and
// ordered by (distance, index) like KDTree.query's output; the per-candidate strand-to-strand distances are lanes over the
// strand's points.  Output per strand: {root: (partner, its end), tip: (partner, its end)}.
#include "mh_common.cuh"

namespace {

constexpr int CN_WARPS = 4;
constexpr int CN_CAP = 256;           // end points within the radius kept per query (the 50 nearest are used, KDTree k = 50)
constexpr int CN_K = 50;

struct Cand { double d; int idx; };

__device__ __forceinline__ double dist3(const double* a, const double* b) {
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrt(x * x + y * y + z * z);                 // scipy: sum of squares in coordinate order, sqrt on output
}
// np.sum(a*b, -1) over 3 elements and np.linalg.norm(.., axis=-1): ((a0b0 + a1b1) + a2b2)
__device__ __forceinline__ double dot3d(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// One end of strand i against one end-point set.  Returns the partner strand or -1 (find_best_connect_strands).
// same_type: root2root / tip2tip (orientation test sim < -thr), else sim > thr.
__device__ int best_partner(const double* __restrict__ pts, const int64_t* __restrict__ off, const int* __restrict__ len, int S,
                            int i, const double* e_pos, const double* e_ori, bool other_is_root, bool same_type, double thr,
                            double dot_thr, Cand* cand, int* s_cnt, int lane, int* overflow) {
    // ---- end points within thr (KDTree.query(point, k=50, distance_upper_bound=thr): strictly closer than the bound)
    if (lane == 0) *s_cnt = 0;
    __syncwarp();
    for (int j = lane; j < S; j += 32) {
        const double* q = pts + 3 * (off[j] + (other_is_root ? 0 : len[j] - 1));
        const double d = dist3(e_pos, q);
        if (d < thr) {
            const int slot = atomicAdd(s_cnt, 1);
            if (slot < CN_CAP) { cand[slot].d = d; cand[slot].idx = j; }
        }
    }
    __syncwarp();
    if (lane == 0 && *s_cnt > CN_CAP) atomicExch(overflow, 1);        // more than CN_CAP end points in the radius: reported, not truncated silently
    int n = min(*s_cnt, CN_CAP);
    // order by (distance, index): rank sort, lanes over candidates
    __shared__ Cand sorted_all[CN_WARPS][CN_CAP];
    Cand* srt = sorted_all[threadIdx.x >> 5];
    for (int a = lane; a < n; a += 32) {
        const Cand c = cand[a];
        int rk = 0;
        for (int b = 0; b < n; ++b) rk += (cand[b].d < c.d || (cand[b].d == c.d && cand[b].idx < c.idx)) ? 1 : 0;
        srt[rk] = c;
    }
    __syncwarp();
    n = min(n, CN_K);
    const int Li = len[i];
    const double* si = pts + 3 * off[i];
    const double sl = dist3(si, si + 3 * (Li - 1)) * 2 / 3;           // np.linalg.norm(strand[0] - strand[-1]) * 2 / 3
    const double eo_n = sqrt(dot3d(e_ori, e_ori));
    double best_loss = 0.0;
    int best = -1;
    for (int c = 0; c < n; ++c) {
        const int j = srt[c].idx;
        if (j == i) continue;                                           // query(): the strand itself is dropped
        const double* sj = pts + 3 * off[j];
        const int Lj = len[j];
        double o[3];
        if (other_is_root) { o[0] = sj[3] - sj[0]; o[1] = sj[4] - sj[1]; o[2] = sj[5] - sj[2]; }
        else { const double* t = sj + 3 * (Lj - 1); o[0] = t[0] - t[-3]; o[1] = t[1] - t[-2]; o[2] = t[2] - t[-1]; }
        const double sim = dot3d(e_ori, o) / (eo_n * sqrt(dot3d(o, o)));
        const bool ok_ori = same_type ? (sim < -dot_thr) : (sim > dot_thr);
        // distance of every point of strand i to strand j (strands_tree[j].query(strand, 1))
        int near5 = 0, near10 = 0;
        double d_first = 0.0, d_last = 0.0;
        for (int p = lane; p < Li; p += 32) {
            double m = 1e300;
            for (int q = 0; q < Lj; ++q) {
                const double x = si[3 * p] - sj[3 * q], y = si[3 * p + 1] - sj[3 * q + 1], z = si[3 * p + 2] - sj[3 * q + 2];
                m = fmin(m, x * x + y * y + z * z);
            }
            m = sqrt(m);
            near5 += m < 0.005;
            near10 += m < 0.01;
            if (p == 0) d_first = m;
            if (p == Li - 1) d_last = m;
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            near5 += __shfl_xor_sync(0xffffffffu, near5, o2);
            near10 += __shfl_xor_sync(0xffffffffu, near10, o2);
        }
        d_first = __shfl_sync(0xffffffffu, d_first, 0);
        d_last = __shfl_sync(0xffffffffu, d_last, (Li - 1) & 31);
        bool ok_dist = (Li < 6) ? (near5 < 4) : (near10 <= 6);        // (the >= 80 points rule of :567 is overwritten here)
        if (d_first < sl && d_last < sl && Li > 20) ok_dist = false;
        if (ok_ori && ok_dist) {
            const double loss = srt[c].d * (1 - fabs(sim));
            if (best < 0 || loss < best_loss) { best_loss = loss; best = j; }   // np.argmin: first minimum
        }
    }
    return best;
}

__global__ void __launch_bounds__(CN_WARPS * 32)
connect_find_kernel(const double* __restrict__ pts, const int64_t* __restrict__ off, const int* __restrict__ len, int S,
                    double thr, double dot_thr, int* __restrict__ info /*[S][4]: root partner, its end (1 root, 2 tip), tip ...*/,
                    int* __restrict__ overflow) {
    __shared__ Cand s_cand[CN_WARPS][CN_CAP];
    __shared__ int s_cnt[CN_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * CN_WARPS + warp;
    if (i >= S) return;
    const int L = len[i];
    const double* s = pts + 3 * off[i];
    double root[3], tip[3], root_ori[3], tip_ori[3];
    for (int k = 0; k < 3; ++k) {
        root[k] = s[k];
        tip[k] = s[3 * (L - 1) + k];
        root_ori[k] = s[3 + k] - s[k];
        tip_ori[k] = s[3 * (L - 1) + k] - s[3 * (L - 2) + k];
    }
    int r_best, r_type = 1, t_best, t_type = 1;
    r_best = best_partner(pts, off, len, S, i, root, root_ori, true, true, thr, dot_thr, s_cand[warp], &s_cnt[warp], lane, overflow);      // root2root
    if (r_best < 0) { r_type = 2; r_best = best_partner(pts, off, len, S, i, root, root_ori, false, false, thr, dot_thr, s_cand[warp], &s_cnt[warp], lane, overflow); }   // root2tip
    t_best = best_partner(pts, off, len, S, i, tip, tip_ori, true, false, thr, dot_thr, s_cand[warp], &s_cnt[warp], lane, overflow);       // tip2root
    if (t_best < 0) { t_type = 2; t_best = best_partner(pts, off, len, S, i, tip, tip_ori, false, true, thr, dot_thr, s_cand[warp], &s_cnt[warp], lane, overflow); }     // tip2tip
    if (lane == 0) {
        info[4 * i] = r_best; info[4 * i + 1] = r_best < 0 ? 0 : r_type;
        info[4 * i + 2] = t_best; info[4 * i + 3] = t_best < 0 ? 0 : t_type;
    }
}

// Share of a strand's points that fall on occupied voxels, as the reference evaluates it (HairGrow.py:514-523):
// world point (float64) -> points_to_voxel (flip y, z; minus the FLOAT32 voxel_min; / 0.0025) -> torch.round (half to even) ->
// occ[0, z, y, x] with torch's negative-index wrap.  frac[i] = sum(occ) / n, or -1 when an index leaves the grid upwards
// (the reference's `check = False; break`).  shift: the random perturbation added to every point (:531), float64 [3].
// voxel_space != 0: the points are voxel coordinates already (random_move_strands, PMVO_utils.py:629).
__global__ void strand_occupancy_kernel(const double* __restrict__ pts, const int64_t* __restrict__ off, const int* __restrict__ len,
                                        int S, const double* __restrict__ shift, const float4* __restrict__ vol, int gx, int gy,
                                        int gz, int voxel_space, double* __restrict__ frac) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= S) return;
    const double* s = pts + 3 * off[warp];
    const int L = len[warp];
    const double vmx = (double)(-0.32f), vmy = (double)(-0.32f), vmz = (double)(-0.24f);
    double sum = 0.0;
    int bad = 0;
    for (int p = lane; p < L; p += 32) {
        double x = s[3 * p], y = s[3 * p + 1], z = s[3 * p + 2];
        if (shift) { x += shift[3 * warp]; y += shift[3 * warp + 1]; z += shift[3 * warp + 2]; }
        long long ix, iy, iz;
        if (voxel_space) { ix = (long long)rint(x); iy = (long long)rint(y); iz = (long long)rint(z); }
        else {
            y *= -1; z *= -1;
            ix = (long long)rint((x - vmx) / 0.0025); iy = (long long)rint((y - vmy) / 0.0025); iz = (long long)rint((z - vmz) / 0.0025);
        }
        if (iz >= gz || iy >= gy || ix >= gx) { bad = 1; continue; }
        if (ix < 0) ix += gx;
        if (iy < 0) iy += gy;
        if (iz < 0) iz += gz;
        if (ix < 0 || iy < 0 || iz < 0) { bad = 1; continue; }           // torch would raise here
        sum += (double)vol[((size_t)iz * gy + iy) * gx + ix].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); bad |= __shfl_xor_sync(0xffffffffu, bad, o); }
    if (lane == 0) frac[warp] = bad ? -1.0 : sum / (double)L;
}

}  // namespace

extern "C" int mh_connect_find(void* stream, const double* points, const int64_t* offsets, const int32_t* lengths, int64_t n_strands,
                               double connect_threshold, double dot_threshold, int32_t* info, int32_t* overflow) {
    MH_CHECK_ARG(info && overflow && (n_strands == 0 || (points && offsets && lengths)) && n_strands >= 0 && n_strands < (1ll << 31), "bad arguments");
    if (n_strands == 0) return 0;
    connect_find_kernel<<<(unsigned)((n_strands + CN_WARPS - 1) / CN_WARPS), CN_WARPS * 32, 0, (cudaStream_t)stream>>>(
        points, offsets, lengths, (int)n_strands, connect_threshold, dot_threshold, info, overflow);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_strand_occupancy(void* stream, const double* points, const int64_t* offsets, const int32_t* lengths,
                                   int64_t n_strands, const double* shift, const void* volume, int32_t gx, int32_t gy, int32_t gz,
                                   int32_t voxel_space, double* frac) {
    MH_CHECK_ARG(frac && volume && (n_strands == 0 || (points && offsets && lengths)) && n_strands >= 0, "bad arguments");
    if (n_strands == 0) return 0;
    strand_occupancy_kernel<<<(unsigned)((n_strands * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        points, offsets, lengths, (int)n_strands, shift, reinterpret_cast<const float4*>(volume), gx, gy, gz, voxel_space, frac);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
