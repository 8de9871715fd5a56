// Multi-view orientation / occupancy fusion into the voxel volume (PMVO.refine, PMVO.py:695-764).
//
// The reference builds a Python dict keyed by "x_y_z" and calls compute_points_similarity per voxel.  Here (details at
// "fusion passes" below): a zero fill of the volume on an auxiliary stream, concurrent with a binning kernel that groups
// the points into per-voxel records through one 64-bit atomic per (warp, voxel) group on a persistent, self-cleaning
// 8 B/voxel plane; then a medoid kernel (warp per record, lane = candidate, torch.mean's summation order) that writes the
// winners straight into the volume -- or into a compact winner list, the multi-GPU exchange format.
// Volume layout: float4 [gz][gy][gx] = {ori.x, -ori.y, -ori.z, occ}: the frame HairGrowing works in
// (HairGrow.py:45-55), one 16 B fetch per trace step.
// Bound: HBM streaming.  Algorithmic bytes = n*(12+12+4) point/direction/key + n*16 record traffic + 16 B/voxel volume
// write; 256x256x192: 12.58 M voxels -> 201 MB of volume writes dominate.
#include <algorithm>
#include <cstdlib>
#include "mh_common.cuh"
#include "mh_torch_sum.cuh"


namespace {

struct VGrid { double mx, my, mz, vs; int gx, gy, gz; };

// p2v (PMVO_utils.py:386-404): points[:,1:] *= -1 ; round((p - min)/vsize) in float64, half to even; clip.
__device__ __forceinline__ void p2v(const VGrid& g, float px, float py, float pz, int& x, int& y, int& z) {
    const double fx = rint(((double)px - g.mx) / g.vs);
    const double fy = rint((-(double)py - g.my) / g.vs);
    const double fz = rint((-(double)pz - g.mz) / g.vs);
    // astype(int32) then clip; values are far inside int32 range for any sane input, clamp in double first
    x = (int)fmin(fmax(fx, 0.0), (double)(g.gx - 1));
    y = (int)fmin(fmax(fy, 0.0), (double)(g.gy - 1));
    z = (int)fmin(fmax(fz, 0.0), (double)(g.gz - 1));
}

// ---- fusion passes ---------------------------------------------------------------------------------------
// Shape of the problem: n points (1.7 M) fall into M occupied voxels (81 k, ~20 points each) of a 12.6 M-voxel grid
// whose 16 B/voxel zero fill (201 MB) is the only bandwidth-sized term; grouping points by voxel is latency bound
// (atomics, dependent accesses) and the medoid is instruction bound.
//   fill       memset of the volume on an auxiliary stream, concurrent with the binning kernel
//   K1 bin     per point : p2v key -> ONE 64-bit atomic on the dense `plane` {count | slot+1} gives the arrival rank;
//                          the voxel's first arrival allocates the voxel's record from its BLOCK's slot range (shared-
//                          memory counter: no global allocation counter) and publishes the slot in the high word.  The
//                          point's {unit direction, id} goes into the record: one contiguous run of (1 + CAP) float4 =
//                          header {key} + CAP entries; arrivals beyond CAP are pushed on a per-voxel chain.
//   K2 medoid  per voxel : (after the fill) one warp per record, lane = candidate; entries rank-sorted back into point
//                          order through shared memory; medoid under |cos| in torch.mean's summation order; the
//                          winner's raw direction goes into the volume; the voxel's plane entry is reset.
//   K3 big     records of more than CAP points (rare), through global scratch.
// The plane is persistent workspace state: all-zero on entry, all-zero again on exit (only M entries are touched), so
// no per-call clear of a dense array and no per-point key/rank arrays exist.  HBM traffic ~= points + directions +
// volume (+ records, mostly L2-resident between the kernels).
struct FuseHdr { int n_big, max_cnt, n_over, scratch_cursor, n_winners, pad[11]; };
static_assert(sizeof(FuseHdr) == 64, "header size");
typedef unsigned long long u64;
constexpr int FUSE_CAP = 32;                 // entries per record
constexpr int64_t FUSE_STRIDE = FUSE_CAP + 1;
constexpr int FUSE_BLOCK = 256;              // points per K1 block = record slots owned by the block

// torch.argmax semantics: NaN counts as the maximum, ties go to the lowest index
__device__ __forceinline__ bool arg_better_max(float s, int k, float best, int bk) {
    const bool sn = s != s, bn = best != best;
    if (sn || bn) return sn && (!bn || k < bk);
    return s > best || (s == best && k < bk);
}

// spin read of a plane word: relaxed is enough (nothing else is read on the strength of it in this kernel; an acquire
// would also invalidate L1 on every poll)
__device__ __forceinline__ u64 vf_ld_relaxed(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// one coordinate of p2v: rint(d / vs) with d = p - min.  d * (1/vs) differs from the correctly rounded quotient by a
// couple of ulps, so both round to the same integer unless the product sits within 1e-9 of a half-integer; only then is
// the division evaluated.
__device__ __forceinline__ double p2v_coord(double d, double vs, double inv) {
    const double q = d * inv;
    const double f = q - floor(q);
    if (fabs(f - 0.5) < 1e-9 || !(fabs(q) < 1e9)) return rint(d / vs);
    return rint(q);
}
__device__ __forceinline__ void p2v_fast(const VGrid& g, double inv, float px, float py, float pz, int& x, int& y, int& z) {
    const double fx = p2v_coord((double)px - g.mx, g.vs, inv);
    const double fy = p2v_coord(-(double)py - g.my, g.vs, inv);
    const double fz = p2v_coord(-(double)pz - g.mz, g.vs, inv);
    x = (int)fmin(fmax(fx, 0.0), (double)(g.gx - 1));
    y = (int)fmin(fmax(fy, 0.0), (double)(g.gy - 1));
    z = (int)fmin(fmax(fz, 0.0), (double)(g.gz - 1));
}

__global__ void __launch_bounds__(FUSE_BLOCK, 8)
bin_kernel(VGrid g, double inv_vs, const float* __restrict__ pts, const float* __restrict__ dirs,
           const uint8_t* __restrict__ valid, int64_t n, u64* plane,
           float4* records, int* over_head, float4* __restrict__ over_ent, int* __restrict__ over_next,
           int* block_cnt, FuseHdr* hdr, int* __restrict__ vox_index) {
    __shared__ int s_alloc;
    const int64_t i = (int64_t)blockIdx.x * FUSE_BLOCK + threadIdx.x;
    {
        if (threadIdx.x == 0) s_alloc = 0;
        __syncthreads();
        const int lane = threadIdx.x & 31;
        const unsigned lt = (1u << lane) - 1u;
        const bool active = i < n && (valid == nullptr || valid[i] != 0);
        int k = -1 - lane;                               // inactive lanes get distinct dummy keys
        float4 e = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (active) {
            // both gathers in flight before any arithmetic
            const float px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
            float a = dirs[3 * i], b = dirs[3 * i + 1], c = dirs[3 * i + 2];
            asm volatile("" : "+f"(a), "+f"(b), "+f"(c));              // keep the loads above the p2v branches
            int x, y, z;
            p2v_fast(g, inv_vs, px, py, pz, x, y, z);
            k = (z * g.gy + y) * g.gx + x;
            if (vox_index) vox_index[i] = (x * g.gy + y) * g.gz + z;
            if (b > 0.0f) { a = a * -1.0f; b = b * -1.0f; c = c * -1.0f; }   // PMVO.py:702-703
            const float na = fmaxf(mh_norm3(a, b, c), 1e-8f);        // cosine_similarity's per-operand normalisation
            e = make_float4(a / na, b / na, c / na, __int_as_float((int)i));
        }
        // one leader per distinct key in the warp (neighbouring points usually share a voxel)
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        const int leader = __ffs(peers) - 1, npeers = __popc(peers);
        const bool lead = active && lane == leader;
        int64_t slot1 = 0;
        int base = 0;
        bool alloc = false;
        if (lead) {
            const u64 old = atomicAdd(plane + k, (u64)npeers);
            base = (int)(old & 0xffffffffu);
            slot1 = (int64_t)(old >> 32);
            alloc = (base == 0);                         // the voxel's first arrival
            if (base + npeers > FUSE_CAP) atomicMax(&hdr->max_cnt, base + npeers);   // crowded voxels only (rare)
        }
        // first arrivals take a record from the block's slot range and publish slot+1 in the plane's high word
        const unsigned am = __ballot_sync(0xffffffffu, alloc);
        if (am) {
            const int fa = __ffs(am) - 1;
            int s0 = 0;
            if (lane == fa) { s0 = atomicAdd(&s_alloc, __popc(am)); atomicAdd(block_cnt + blockIdx.x, __popc(am)); }
            s0 = __shfl_sync(0xffffffffu, s0, fa);
            if (alloc) {
                slot1 = (int64_t)blockIdx.x * FUSE_BLOCK + s0 + __popc(am & lt) + 1;
                reinterpret_cast<int*>(records + (slot1 - 1) * FUSE_STRIDE)[0] = k;      // read by K2 only
                atomicAdd(plane + k, (u64)slot1 << 32);
            }
        }
        __syncwarp();
        if (lead && slot1 == 0) {                        // the allocating warp publishes without waiting on anyone
            do { slot1 = (int64_t)(vf_ld_relaxed(plane + k) >> 32); } while (slot1 == 0);
        }
        slot1 = __shfl_sync(0xffffffffu, slot1, leader);
        base = __shfl_sync(0xffffffffu, base, leader);
        if (active) {
            const int r = base + __popc(peers & lt);
            if (r < FUSE_CAP) {
                records[(slot1 - 1) * FUSE_STRIDE + 1 + r] = e;
            } else {                                     // crowded voxel: push on its overflow chain
                const int o = atomicAdd(&hdr->n_over, 1);
                over_ent[o] = e;
                over_next[o] = atomicExch(over_head + (slot1 - 1), o + 1);
            }
        }
    }
}

// |cos| of two unit vectors as torch.cosine_similarity's dot evaluates it
__device__ __forceinline__ float vf_absdot(const float4& w, const float4& v) { return fabsf((w.x * v.x + w.y * v.y) + w.z * v.z); }

// compute_points_similarity (PMVO_utils.py:366-382): argmax_k mean_j |cos(o_k, o_j)| with the points in original order
// (first-index tie-break).  One warp per record, lane k = candidate k; the row sum follows torch.mean over K contiguous
// floats (mh_torch_sum.cuh): 8 vector-lane accumulators acc[t & 7] over t < 8*(K/8) (for K < 40 the 4 row accumulators
// collapse to this sequential form), the scalar tail summed first, then the 8 accumulators added in order; K < 8 takes
// torch's scalar path.  Records of more than CAP points go through global scratch with the generic summation.
constexpr int MEDOID_WARPS = 8;                          // one CTA per binning block: its warps share that block's records

// warp arg-max of non-negative (or NaN) means with torch.argmax semantics: positive floats and NaN order like their bit
// patterns (NaN above everything = counts as the maximum), ties go to the lowest candidate index.
__device__ __forceinline__ int vf_warp_argmax(float mean, int k, bool valid) {
    const unsigned bits = valid ? (__float_as_uint(mean) & 0x7fffffffu) + 1u : 0u;      // -0 -> +0; 0 = "no candidate"
    const unsigned mx = __reduce_max_sync(0xffffffffu, bits);
    return (int)__reduce_min_sync(0xffffffffu, (bits == mx && valid) ? (unsigned)k : 0x7fffffffu);
}

// winner's direction: flip to dir.y <= 0 (PMVO.py:702-703), then the frame HairGrowing works in.  LIST: into the compact
// winner list {dir, voxel key} at index w (the multi-GPU exchange format), else straight into the zero-filled volume
template <bool LIST>
__device__ __forceinline__ void store_winner(float4* __restrict__ out, int w, int key, float a, float b, float c) {
    if (b > 0.0f) { a = a * -1.0f; b = b * -1.0f; c = c * -1.0f; }
    if (LIST) out[w] = make_float4(a, -b, -c, __int_as_float(key));
    else out[key] = make_float4(a, -b, -c, 1.0f);
}

// CTA b walks the records allocated by binning block b (contiguous slots, so the hardware block scheduler does the load
// balancing); warp w takes records w, w+8, ...  The next record's header / count / entries are loaded one record
// ahead, and the winner's raw direction is fetched at the end of a record and stored into the (already zero-filled)
// volume at the end of the next one, so no load is waited for where it is issued.
template <bool LIST>
__global__ void __launch_bounds__(MEDOID_WARPS * 32, 5)
fuse_medoid_kernel(const int* __restrict__ block_cnt, const float4* __restrict__ records, FuseHdr* hdr,
                   int4* __restrict__ big_list, u64* __restrict__ plane, const float* __restrict__ dirs,
                   float4* __restrict__ out) {
    __shared__ float4 s_u[MEDOID_WARPS][FUSE_CAP];
    __shared__ int4 s_ids[MEDOID_WARPS][FUSE_CAP / 4];
    __shared__ int s_wbase;
    const int cnt = block_cnt[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int wbase = 0;
    if (LIST) {                                          // the block's winners take consecutive list entries
        if (cnt == 0) return;
        if (threadIdx.x == 0) s_wbase = atomicAdd(&hdr->n_winners, cnt);
        __syncthreads();
        wbase = s_wbase;
    }
    if (warp >= cnt) return;
    float4* u = s_u[warp];
    const int4* ids4 = s_ids[warp];
    const float4* rec = records + ((int64_t)blockIdx.x * FUSE_BLOCK + warp) * FUSE_STRIDE;
    constexpr int64_t STEP = MEDOID_WARPS * FUSE_STRIDE;
    // software pipeline: {key, K, e} of the current record were loaded during the previous one
    int key_n = reinterpret_cast<const int*>(rec)[0];
    float4 e_n = rec[1 + lane];
    int K_n = (int)(__ldcg(plane + key_n) & 0xffffffffu);
    bool pend = false;                                   // lane 0: a winner whose direction load is in flight
    int pend_key = 0, pend_w = 0;
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    for (int ri = warp; ri < cnt; ri += MEDOID_WARPS, rec += STEP) {
        const int key = key_n, K = K_n;
        const float4 e = e_n;
        const bool more = ri + MEDOID_WARPS < cnt;
        if (more) {
            key_n = reinterpret_cast<const int*>(rec + STEP)[0];
            e_n = rec[STEP + 1 + lane];
        }
        int best_id = -1;
        if (K <= FUSE_CAP) {
            const int id = (lane < K) ? __float_as_int(e.w) : 0x7fffffff;
            reinterpret_cast<int*>(s_ids[warp])[lane] = id;
            __syncwarp();
            int rk = 0;                                  // rank = number of smaller ids in the record
            for (int t4 = 0; 4 * t4 < K; ++t4) {
                const int4 v = ids4[t4];
                rk += (v.x < id) + (v.y < id) + (v.z < id) + (v.w < id);
            }
            if (lane < K) u[rk] = e;
            __syncwarp();
            const float4 me = u[min(lane, K - 1)];       // candidate = sorted position `lane`
            float total;
            if (K < 8) {
                // torch's scalar path: 4 accumulators over rows of 4, leftovers into the first, then p0+p1+p2+p3
                float x[7];
#pragma unroll
                for (int t = 0; t < 7; ++t) x[t] = (t < K) ? vf_absdot(me, u[t]) : 0.0f;
                if (K >= 4) {
                    float q0 = 0.0f + x[0];
                    const float q1 = 0.0f + x[1], q2 = 0.0f + x[2], q3 = 0.0f + x[3];
#pragma unroll
                    for (int t = 4; t < 7; ++t) if (t < K) q0 += x[t];
                    q0 += q1; q0 += q2; q0 += q3;
                    total = q0;
                } else {
                    total = 0.0f;
#pragma unroll
                    for (int t = 0; t < 3; ++t) if (t < K) total += x[t];
                }
            } else {
                const int nv8 = K & ~7;
                float acc[8];
#pragma unroll
                for (int l = 0; l < 8; ++l) acc[l] = 0.0f;
                for (int t0 = 0; t0 < nv8; t0 += 8) {
#pragma unroll
                    for (int l = 0; l < 8; ++l) acc[l] += vf_absdot(me, u[t0 + l]);
                }
                total = 0.0f;
                for (int t = nv8; t < K; ++t) total += vf_absdot(me, u[t]);
#pragma unroll
                for (int l = 0; l < 8; ++l) total += acc[l];
            }
            const int bk = vf_warp_argmax(total / (float)K, lane, lane < K);
            best_id = __float_as_int(u[bk].w);
            __syncwarp();
        } else if (lane == 0) {
            // crowded voxel: left to fuse_medoid_big_kernel (keeps this kernel's register count low)
            big_list[atomicAdd(&hdr->n_big, 1)] = make_int4((int)((int64_t)blockIdx.x * FUSE_BLOCK + ri), K, wbase + ri, key);
        }
        if (more) K_n = (int)(__ldcg(plane + key_n) & 0xffffffffu);      // key_n has had a record's time to arrive
        if (lane == 0) {
            if (pend) store_winner<LIST>(out, pend_w, pend_key, p0, p1, p2);
            pend = best_id >= 0;
            if (pend) {                                  // raw direction: consumed (flipped, stored) one record later
                p0 = dirs[3 * best_id]; p1 = dirs[3 * best_id + 1]; p2 = dirs[3 * best_id + 2];
                pend_key = key; pend_w = wbase + ri;
            }
            plane[key] = 0ull;                           // leave the plane clean for the next call
        }
    }
    if (lane == 0 && pend) store_winner<LIST>(out, pend_w, pend_key, p0, p1, p2);
}

// Records of more than CAP points (rare): one warp each; record + overflow chain gathered into global scratch, ranked
// into point order there, generic torch.mean row sums.
template <bool LIST>
__global__ void __launch_bounds__(256)
fuse_medoid_big_kernel(const float4* __restrict__ records, const int* __restrict__ over_head, const float4* __restrict__ over_ent,
                       const int* __restrict__ over_next, FuseHdr* hdr, const int4* __restrict__ big_list, float4* scratch,
                       const float* __restrict__ dirs, float4* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbig = hdr->n_big;
    for (int bi = blockIdx.x * 8 + warp; bi < nbig; bi += gridDim.x * 8) {
        const int slot = big_list[bi].x, K = big_list[bi].y;
        const float4* rec = records + (int64_t)slot * FUSE_STRIDE;
        int sb = 0;
        if (lane == 0) sb = atomicAdd(&hdr->scratch_cursor, 2 * K);
        sb = __shfl_sync(0xffffffffu, sb, 0);
        float4* raw = scratch + sb;
        float4* srt = raw + K;
        raw[lane] = rec[1 + lane];
        if (lane == 0) {
            int t = FUSE_CAP;
            for (int o = over_head[slot]; o != 0 && t < K; o = over_next[o - 1]) raw[t++] = over_ent[o - 1];
        }
        __syncwarp();
        for (int a = lane; a < K; a += 32) {
            const float4 ea = __ldcg(raw + a);
            const int v = __float_as_int(ea.w);
            int rk = 0;
            for (int t = 0; t < K; ++t) rk += (__float_as_int(__ldcg(raw + t).w) < v) ? 1 : 0;
            srt[rk] = ea;
        }
        __syncwarp();
        float best = -1e30f; int bk = 0x7fffffff;
        for (int k = lane; k < K; k += 32) {
            const float4 me = __ldcg(srt + k);
            const float sum = mh_torch_inner_sum(K, [&](int t) { return vf_absdot(me, __ldcg(srt + t)); }) / (float)K;
            if (arg_better_max(sum, k, best, bk)) { best = sum; bk = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (arg_better_max(ob, ok, best, bk)) { best = ob; bk = ok; }
        }
        if (lane == 0) {
            const int bid = __float_as_int(__ldcg(srt + bk).w);
            store_winner<LIST>(out, big_list[bi].z, big_list[bi].w, dirs[3 * (int64_t)bid], dirs[3 * (int64_t)bid + 1], dirs[3 * (int64_t)bid + 2]);
        }
        __syncwarp();
    }
}

constexpr int FILL_CHUNK = 32768;            // bytes per TMA bulk store of the zero fill
// Zero fill by TMA bulk stores: a zeroed shared-memory tile is the source of every store, so nothing has to be waited
// for until the end; one issuing thread per CTA, one CTA per SM.
__global__ void __launch_bounds__(32, 1)
fill_bulk_kernel(unsigned char* __restrict__ dst, u64 bytes) {
    extern __shared__ __align__(128) unsigned char zbuf[];
    for (int t = threadIdx.x; t < FILL_CHUNK / 16; t += 32) reinterpret_cast<int4*>(zbuf)[t] = make_int4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the TMA unit
    __syncwarp();
    if (threadIdx.x == 0) {
        const unsigned saddr = (unsigned)__cvta_generic_to_shared(zbuf);
        const u64 nchunks = (bytes + FILL_CHUNK - 1) / FILL_CHUNK;
        // evict-first: the zero lines stream through L2 without displacing the records / winners the concurrent
        // grouping and medoid kernels keep there
        u64 pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        for (u64 c = blockIdx.x; c < nchunks; c += gridDim.x) {
            const u64 off = c * FILL_CHUNK;
            const unsigned sz = (unsigned)((bytes - off < (u64)FILL_CHUNK) ? (bytes - off) : (u64)FILL_CHUNK);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                         :: "l"(dst + off), "r"(saddr), "r"(sz), "l"(pol) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// winners -> volume (already zero filled): 16 B per occupied voxel; entries with key < 0 are padding
__global__ void __launch_bounds__(256)
scatter_kernel(const float4* __restrict__ winners, int64_t count, int64_t nvox, float4* __restrict__ volume) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const float4 w = winners[t];
        const int key = __float_as_int(w.w);
        if (key >= 0 && key < nvox) volume[key] = make_float4(w.x, w.y, w.z, 1.0f);
    }
}

// winner export: count out, unused capacity marked with key -1
__global__ void __launch_bounds__(256)
winners_finish_kernel(float4* __restrict__ winners, int64_t capacity, const FuseHdr* __restrict__ hdr, int* __restrict__ count_out) {
    const int nw = hdr->n_winners;
    for (int64_t t = nw + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < capacity; t += (int64_t)gridDim.x * blockDim.x)
        winners[t] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
    if (blockIdx.x == 0 && threadIdx.x == 0 && count_out) *count_out = nw;
}

__global__ void overwrite_kernel(VGrid g, const float* __restrict__ pts, const float* __restrict__ dirs, int64_t n,
                                 int* __restrict__ winner) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x, y, z;
    p2v(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], x, y, z);
    atomicMax(winner + (int64_t)(z * g.gy + y) * g.gx + x, (int)i + 1);       // numpy scatter: last writer wins
}
__global__ void overwrite_apply_kernel(int64_t nvox, const int* __restrict__ winner, const float* __restrict__ dirs,
                                       float4* __restrict__ volume) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nvox) return;
    const int w = winner[g];
    if (w > 0) {
        const int i = w - 1;
        volume[g] = make_float4(dirs[3 * i], -dirs[3 * i + 1], -dirs[3 * i + 2], 1.0f);
    }
}

// float4 [gz][gy][gx] -> Occ [gy][gx][gz], Ori [gy][gx][3*gz] float64 (PMVO.py:753-756)
__global__ void to_mat_kernel(const float4* __restrict__ vol, int gx, int gy, int gz, double* __restrict__ occ,
                              double* __restrict__ ori) {
    __shared__ float4 tile[32][33];
    // transpose (z, x) within a fixed y: read x-fastest, write z-fastest
    const int y = blockIdx.z;
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
    for (int dz = threadIdx.y; dz < 32; dz += blockDim.y) {
        const int x = x0 + threadIdx.x, z = z0 + dz;
        if (x < gx && z < gz) tile[dz][threadIdx.x] = vol[((size_t)z * gy + y) * gx + x];
    }
    __syncthreads();
    for (int dx = threadIdx.y; dx < 32; dx += blockDim.y) {
        const int x = x0 + dx, z = z0 + threadIdx.x;
        if (x < gx && z < gz) {
            const float4 v = tile[threadIdx.x][dx];
            const size_t o = ((size_t)y * gx + x);
            occ[o * gz + z] = (double)v.w;
            ori[o * 3 * gz + z] = (double)v.x;
            ori[o * 3 * gz + gz + z] = (double)(-v.y);
            ori[o * 3 * gz + 2 * gz + z] = (double)(-v.z);
        }
    }
}
__global__ void from_mat_kernel(const double* __restrict__ occ, const double* __restrict__ ori, int gx, int gy, int gz,
                                float4* __restrict__ vol) {
    __shared__ float4 tile[32][33];
    const int y = blockIdx.z;
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
    for (int dx = threadIdx.y; dx < 32; dx += blockDim.y) {
        const int x = x0 + dx, z = z0 + threadIdx.x;
        if (x < gx && z < gz) {
            const size_t o = ((size_t)y * gx + x);
            // get_ground_truth_3D_ori/occ cast to float32; HairGrowing.__init__ negates channels 1,2
            const float a = (float)ori[o * 3 * gz + z], b = (float)ori[o * 3 * gz + gz + z], c = (float)ori[o * 3 * gz + 2 * gz + z];
            tile[threadIdx.x][dx] = make_float4(a, b * -1.0f, c * -1.0f, (float)occ[o * gz + z]);
        }
    }
    __syncthreads();
    for (int dz = threadIdx.y; dz < 32; dz += blockDim.y) {
        const int x = x0 + threadIdx.x, z = z0 + dz;
        if (x < gx && z < gz) vol[((size_t)z * gy + y) * gx + x] = tile[dz][threadIdx.x];
    }
}

VGrid make_grid(const double* vmin, double vs, int gx, int gy, int gz) {
    VGrid g; g.mx = vmin[0]; g.my = vmin[1]; g.mz = vmin[2]; g.vs = vs; g.gx = gx; g.gy = gy; g.gz = gz; return g;
}

}  // namespace

// per-call workspace: [hdr 64 B][over_head S int][block_cnt nb int][records S*(CAP+1) float4][over_ent n float4]
// [scratch 2n float4][big_list n/CAP int4][over_next n int] with S = nb * 256 record slots, nb = ceil(n / 256)
namespace {
// auxiliary stream for the volume zero fill (one per device and host thread)
struct AuxStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
AuxStream& aux_stream() {
    static thread_local AuxStream a[16];
    int dev = 0;
    cudaGetDevice(&dev);
    AuxStream& x = a[dev & 15];
    if (!x.s) {
        cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming);
    }
    return x;
}
int64_t fuse_nb(int64_t n) { return (n + FUSE_BLOCK - 1) / FUSE_BLOCK; }
}  // namespace
extern "C" int64_t mh_voxel_fuse_workspace_bytes(int64_t n, int32_t gx, int32_t gy, int32_t gz) {
    (void)gx; (void)gy; (void)gz;
    const int64_t nb = fuse_nb(n), S = nb * FUSE_BLOCK;
    return 64 + 4 * (S + nb + 16) + 16 * (S * FUSE_STRIDE + 3 * n + 8) + 16 * (n / FUSE_CAP + 2) + 4 * (n + 4);
}
// persistent plane: 8 B per voxel {count | slot+1}, all-zero between calls
extern "C" int64_t mh_voxel_fuse_plane_bytes(int32_t gx, int32_t gy, int32_t gz) { return 8 * (int64_t)gx * gy * gz; }
extern "C" int mh_voxel_fuse_plane_init(void* stream, void* plane, int32_t gx, int32_t gy, int32_t gz) {
    MH_CHECK_ARG(plane && gx > 0 && gy > 0 && gz > 0, "bad arguments");
    cudaError_t e = cudaMemsetAsync(plane, 0, (size_t)mh_voxel_fuse_plane_bytes(gx, gy, gz), (cudaStream_t)stream);
    if (e != cudaSuccess) { mh_set_error("mh_voxel_fuse_plane_init: %s", cudaGetErrorString(e)); return 2; }
    return 0;
}

namespace {
int fuse_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
// zero fill of `bytes` on stream s by TMA bulk stores, one CTA per SM (used where the fill has the GPU to itself: the
// winner scatter of the multi-GPU path, empty inputs; the single-GPU fusion overlaps a memset with its binning kernel)
void launch_fill(void* dst, size_t bytes, cudaStream_t s) {
    if ((bytes & 15) || (reinterpret_cast<uintptr_t>(dst) & 15) || fuse_env("MH_FUSE_FILL", 1) == 0) { cudaMemsetAsync(dst, 0, bytes, s); return; }
    static thread_local bool attr[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 15]) {
        cudaFuncSetAttribute(fill_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_CHUNK);
        attr[dev & 15] = true;
    }
    const size_t nchunks = (bytes + FILL_CHUNK - 1) / FILL_CHUNK;
    fill_bulk_kernel<<<(unsigned)std::min<size_t>((size_t)mh_sm_count(), nchunks), 32, FILL_CHUNK, s>>>(reinterpret_cast<unsigned char*>(dst), (u64)bytes);
    MH_COUNT_LAUNCH();
}

// LIST = false: fused volume into `out` (float4 [gz][gy][gx]); LIST = true: winners into `out` (float4 [capacity]), their
// number into *count, unused capacity marked with key -1
template <bool LIST>
int fuse_run(cudaStream_t st, const float* points, const float* dirs, const uint8_t* valid, int64_t n, const double* voxel_min_host,
             double voxel_size, int32_t gx, int32_t gy, int32_t gz, float4* out, int64_t capacity, int32_t* count,
             int32_t* vox_index, void* plane, void* workspace) {
    const int64_t nvox = (int64_t)gx * gy * gz;
    const int64_t nb = fuse_nb(n), S = nb * FUSE_BLOCK;
    const VGrid g = make_grid(voxel_min_host, voxel_size, gx, gy, gz);
    FuseHdr* hdr = reinterpret_cast<FuseHdr*>(workspace);
    int* over_head = reinterpret_cast<int*>(hdr + 1);
    int* block_cnt = over_head + S;
    float4* records = reinterpret_cast<float4*>(block_cnt + ((nb + 3) / 4) * 4);
    float4* over_ent = records + S * FUSE_STRIDE;
    float4* scratch = over_ent + (n + 1);
    int4* big_list = reinterpret_cast<int4*>(scratch + (2 * n + 2));       // at most n / CAP records overflow
    int* over_next = reinterpret_cast<int*>(big_list + (n / FUSE_CAP + 2));
    u64* pl = reinterpret_cast<u64*>(plane);
    // The zero fill of the volume (201 MB at 256x256x192: the only bandwidth-sized term) is a memset on an auxiliary
    // stream, concurrent with the latency-bound binning kernel; the medoid kernel, which writes the results into the
    // volume, joins it first.  Measured alternatives (profiles/r2_fuse.md): streaming the fill from the binning / medoid
    // kernels' own threads is additive; a fill that keeps running underneath the medoid kernel (TMA bulk stores from one
    // CTA per SM, winners deferred to a scatter kernel) slows both sides by more than the overlap gains.
    AuxStream& ax = aux_stream();
    if (!LIST) {
        cudaEventRecord(ax.fork, st);
        cudaStreamWaitEvent(ax.s, ax.fork, 0);
        cudaMemsetAsync(out, 0, sizeof(float4) * nvox, ax.s);
        cudaEventRecord(ax.join, ax.s);
    }
    cudaMemsetAsync(hdr, 0, sizeof(FuseHdr) + sizeof(int) * (S + nb), st);  // header, overflow chain heads, per-block record counts
    bin_kernel<<<(unsigned)nb, FUSE_BLOCK, 0, st>>>(g, 1.0 / voxel_size, points, dirs, valid, n, pl, records, over_head, over_ent,
                                                    over_next, block_cnt, hdr, vox_index);
    MH_COUNT_LAUNCH();
    if (!LIST) cudaStreamWaitEvent(st, ax.join, 0);
    fuse_medoid_kernel<LIST><<<(unsigned)nb, MEDOID_WARPS * 32, 0, st>>>(block_cnt, records, hdr, big_list, pl, dirs, out);
    MH_COUNT_LAUNCH();
    fuse_medoid_big_kernel<LIST><<<(unsigned)std::min<int64_t>((int64_t)mh_sm_count() * 4, n / (8 * FUSE_CAP) + 1), 256, 0, st>>>(
        records, over_head, over_ent, over_next, hdr, big_list, scratch, dirs, out);
    MH_COUNT_LAUNCH();
    if (LIST) {
        winners_finish_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)mh_sm_count() * 2, (capacity + 255) / 256)), 256, 0, st>>>(
            out, capacity, hdr, count);
        MH_COUNT_LAUNCH();
    }
    return 0;
}
}  // namespace

extern "C" int mh_voxel_fuse(void* stream, const float* points, const float* dirs, const uint8_t* valid, int64_t n,
                             const double* voxel_min_host, double voxel_size, int32_t gx, int32_t gy, int32_t gz,
                             void* volume, int32_t* vox_index, void* plane, void* workspace, int64_t workspace_bytes) {
    MH_CHECK_ARG(volume && workspace && plane && voxel_min_host && (n == 0 || (points && dirs)), "null pointer");
    MH_CHECK_ARG(gx > 0 && gy > 0 && gz > 0 && voxel_size > 0, "bad grid");
    MH_CHECK_ARG((int64_t)gx * gy * gz < (1ll << 31) && n < (1ll << 31) - 512, "grid or point count too large for int32 keys");
    MH_CHECK_ARG(workspace_bytes >= mh_voxel_fuse_workspace_bytes(n, gx, gy, gz), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) launch_fill(volume, sizeof(float4) * (size_t)gx * gy * gz, st);
    else fuse_run<false>(st, points, dirs, valid, n, voxel_min_host, voxel_size, gx, gy, gz, reinterpret_cast<float4*>(volume), 0, nullptr,
                         vox_index, plane, workspace);
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_voxel_fuse_winners(void* stream, const float* points, const float* dirs, const uint8_t* valid, int64_t n,
                                     const double* voxel_min_host, double voxel_size, int32_t gx, int32_t gy, int32_t gz,
                                     void* winners /*float4 [capacity]*/, int64_t capacity, int32_t* count, int32_t* vox_index,
                                     void* plane, void* workspace, int64_t workspace_bytes) {
    MH_CHECK_ARG(winners && count && workspace && plane && voxel_min_host && (n == 0 || (points && dirs)), "null pointer");
    MH_CHECK_ARG(gx > 0 && gy > 0 && gz > 0 && voxel_size > 0, "bad grid");
    MH_CHECK_ARG((int64_t)gx * gy * gz < (1ll << 31) && n < (1ll << 31) - 512, "grid or point count too large for int32 keys");
    MH_CHECK_ARG(capacity >= std::min<int64_t>(n, (int64_t)gx * gy * gz), "winner capacity below min(n, voxels)");
    MH_CHECK_ARG(workspace_bytes >= mh_voxel_fuse_workspace_bytes(n, gx, gy, gz), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(workspace, 0, sizeof(FuseHdr), st);
        winners_finish_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)mh_sm_count() * 2, (capacity + 255) / 256)), 256, 0, st>>>(
            reinterpret_cast<float4*>(winners), capacity, reinterpret_cast<const FuseHdr*>(workspace), count);
        MH_COUNT_LAUNCH();
    } else {
        fuse_run<true>(st, points, dirs, valid, n, voxel_min_host, voxel_size, gx, gy, gz, reinterpret_cast<float4*>(winners), capacity, count,
                       vox_index, plane, workspace);
    }
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_voxel_scatter(void* stream, const void* winners, int64_t m, int32_t gx, int32_t gy, int32_t gz, void* volume,
                                int32_t zero_fill) {
    MH_CHECK_ARG(volume && (m == 0 || winners) && gx > 0 && gy > 0 && gz > 0 && m >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nvox = (int64_t)gx * gy * gz;
    if (zero_fill) launch_fill(volume, sizeof(float4) * nvox, st);
    if (m > 0) {
        scatter_kernel<<<(unsigned)std::min<int64_t>((int64_t)mh_sm_count() * 2, (m + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(winners), m, nvox, reinterpret_cast<float4*>(volume));
        MH_COUNT_LAUNCH();
    }
    MH_CHECK_LAUNCH();
    return 0;
}

/* after synchronising the stream: the largest number of points that fell into one voxel in the last mh_voxel_fuse
 * call on this workspace if some voxel held more than 32 (such voxels take the overflow chain and
 * fuse_medoid_big_kernel), else 0.  Informational. */
extern "C" int mh_voxel_fuse_max_points(const void* workspace, int32_t* max_k_host) {
    MH_CHECK_ARG(workspace && max_k_host, "null pointer");
    cudaError_t e = cudaMemcpy(max_k_host, &reinterpret_cast<const FuseHdr*>(workspace)->max_cnt, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { mh_set_error("mh_voxel_fuse_max_points: %s", cudaGetErrorString(e)); return 2; }
    return 0;
}

extern "C" int mh_voxel_overwrite(void* stream, const float* points, const float* dirs, int64_t n,
                                     const double* voxel_min_host, double voxel_size, int32_t gx, int32_t gy,
                                     int32_t gz, void* volume, void* winner_ws /* int32 [gz*gy*gx] */) {
    MH_CHECK_ARG(volume && winner_ws && voxel_min_host && (n == 0 || (points && dirs)), "null pointer");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nvox = (int64_t)gx * gy * gz;
    const VGrid g = make_grid(voxel_min_host, voxel_size, gx, gy, gz);
    cudaMemsetAsync(winner_ws, 0, sizeof(int) * nvox, st);
    overwrite_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, points, dirs, n, reinterpret_cast<int*>(winner_ws));
    MH_COUNT_LAUNCH();
    overwrite_apply_kernel<<<(unsigned)((nvox + 255) / 256), 256, 0, st>>>(nvox, reinterpret_cast<int*>(winner_ws), dirs,
                                                                          reinterpret_cast<float4*>(volume));
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_volume_to_mat(void* stream, const void* volume, int32_t gx, int32_t gy, int32_t gz, double* occ, double* ori) {
    MH_CHECK_ARG(volume && occ && ori && gx > 0 && gy > 0 && gz > 0, "bad arguments");
    dim3 grid((gx + 31) / 32, (gz + 31) / 32, gy), block(32, 8);
    to_mat_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(volume), gx, gy, gz, occ, ori);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_volume_from_mat(void* stream, const double* occ, const double* ori, int32_t gx, int32_t gy, int32_t gz, void* volume) {
    MH_CHECK_ARG(volume && occ && ori && gx > 0 && gy > 0 && gz > 0, "bad arguments");
    dim3 grid((gx + 31) / 32, (gz + 31) / 32, gy), block(32, 8);
    from_mat_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(occ, ori, gx, gy, gz, reinterpret_cast<float4*>(volume));
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
