// Multi-view orientation / occupancy fusion into the voxel volume (PMVO.refine, PMVO.py:695-764).
//
// The reference builds a Python dict keyed by "x_y_z" and calls compute_points_similarity per voxel.  Here:
//   1 key      p2v in float64 with round-half-even (PMVO_utils.py:386-404); per-voxel count (atomics)
//   2 scan     exclusive scan of the dense count volume -> bucket starts
//   3 fill     point ids into their voxel's bucket
//   4 fuse     one thread per voxel, x-fastest: empty voxels stream zeros, occupied voxels sort their bucket
//              back into original point order (first-index tie-break of argmax) and write the medoid.
// Volume layout: float4 [gz][gy][gx] = {ori.x, -ori.y, -ori.z, occ}: the frame HairGrowing works in
// (HairGrow.py:45-55), one 16 B fetch per trace step.  Bucket keys use the same z,y,x order so pass 4 reads
// its 8 B of bucket bounds and writes its 16 B fully coalesced.
// Bound: HBM streaming.  Algorithmic bytes = n*(12+12) point reads + 4 B/voxel count write+read (x2 for the
// scan) + 16 B/voxel volume write; 256x256x192: 12.58 M voxels -> 201 MB of volume writes dominate.
#include <algorithm>
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

// flipped direction of point i (PMVO.py:702-703: ori[ori.y>0] *= -1)
__device__ __forceinline__ void load_dir(const float* __restrict__ dirs, int i, float& a, float& b, float& c) {
    a = dirs[3 * i]; b = dirs[3 * i + 1]; c = dirs[3 * i + 2];
    if (b > 0.0f) { a = a * -1.0f; b = b * -1.0f; c = c * -1.0f; }
}

// Pass 1 (per point): voxel key, arrival rank inside the voxel (atomic count in a dense int32 plane), and the first
// arrival of each voxel registers it in the compact list of occupied voxels.
__global__ void __launch_bounds__(256)
count_kernel_vf(VGrid g, const float* __restrict__ pts, int64_t n, int* __restrict__ cnt, int* __restrict__ key,
                int* __restrict__ rank, int* __restrict__ occ_list, int* __restrict__ hdr, int* __restrict__ vox_index) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int k = -1 - lane;                                   // inactive lanes get distinct dummy keys
    if (i < n) {
        int x, y, z;
        p2v(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], x, y, z);
        k = (z * g.gy + y) * g.gx + x;
        key[i] = k;
        if (vox_index) vox_index[i] = (x * g.gy + y) * g.gz + z;
    }
    // warp-aggregated arrival ranks: neighbouring points usually share a voxel, so one atomic serves the group
    const unsigned peers = __match_any_sync(0xffffffffu, k);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader && i < n) base = atomicAdd(cnt + k, __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n) {
        const int r = base + __popc(peers & ((1u << lane) - 1));
        rank[i] = r;
        if (r == 0) occ_list[atomicAdd(hdr, 1)] = k;
    }
}

// Pass 2 (per occupied voxel): reserve a bucket [base, base+count) (any disjoint placement will do, so an atomic
// cursor -- one add per warp -- replaces a scan); the dense plane now holds the bucket base.
__global__ void __launch_bounds__(256)
bucket_kernel(const int* __restrict__ occ_list, int* __restrict__ cnt, int* __restrict__ cnt_list, int* __restrict__ hdr) {
    const int M = hdr[0];
    const int lane = threadIdx.x & 31;
    for (int j0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; j0 < M; j0 += gridDim.x * blockDim.x) {
        const int j = j0 + lane;
        int c = 0, k = 0;
        if (j < M) { k = occ_list[j]; c = cnt[k]; cnt_list[j] = c; }
        int incl = c, mx = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        int base = 0;
        if (lane == 31) { base = atomicAdd(hdr + 2, incl); atomicMax(hdr + 1, mx); }
        base = __shfl_sync(0xffffffffu, base, 31);
        if (j < M) cnt[k] = base + incl - c;
    }
}

// Pass 3 (per point): drop the point id into its voxel's bucket.
__global__ void __launch_bounds__(256)
scatter_kernel(int64_t n, const int* __restrict__ key, const int* __restrict__ rank, const int* __restrict__ cnt,
               int* __restrict__ bucket) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bucket[cnt[key[i]] + rank[i]] = (int)i;
}

// Pass 4: one warp per occupied voxel.  Bucket ids are rank-sorted back into original point order (argmax ties go
// to the first point) and the medoid under |cos| is taken with torch.mean's summation order
// (compute_points_similarity, PMVO_utils.py:366-382).  K <= FUSE_FASTK stays in shared memory; larger voxels use the
// global scratch `sorted` (rare).
constexpr int FUSE_WARPS = 8, FUSE_FASTK = 64;

__global__ void __launch_bounds__(FUSE_WARPS * 32)
fuse_medoid_kernel(const int* __restrict__ occ_list, const int* __restrict__ cnt_list, const int* __restrict__ cnt,
                   const int* __restrict__ hdr, const int* __restrict__ bucket, int* __restrict__ sorted,
                   const float* __restrict__ dirs, float4* __restrict__ volume) {
    __shared__ int s_idx_all[FUSE_WARPS][FUSE_FASTK];
    __shared__ float s_u_all[FUSE_WARPS][FUSE_FASTK * 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = hdr[0];
    for (int j = blockIdx.x * FUSE_WARPS + warp; j < M; j += gridDim.x * FUSE_WARPS) {
        const int g = occ_list[j];
        const int K = cnt_list[j], base = cnt[g];
        int bk = 0, best_id;
        if (K == 1) {
            best_id = bucket[base];
        } else if (K <= FUSE_FASTK) {
            int* s_idx = s_idx_all[warp];
            float* u = s_u_all[warp];
            for (int a = lane; a < K; a += 32) {
                const int v = bucket[base + a];
                int rank = 0;
                for (int b = 0; b < K; ++b) rank += (bucket[base + b] < v) ? 1 : 0;
                s_idx[rank] = v;
            }
            __syncwarp();
            for (int a = lane; a < K; a += 32) {
                float a0, a1, a2;
                load_dir(dirs, s_idx[a], a0, a1, a2);
                const float na = fmaxf(mh_norm3(a0, a1, a2), 1e-8f);
                u[3 * a] = a0 / na; u[3 * a + 1] = a1 / na; u[3 * a + 2] = a2 / na;
            }
            __syncwarp();
            float best = -1e30f; bk = 0x7fffffff;
            for (int k = lane; k < K; k += 32) {
                const float a0 = u[3 * k], a1 = u[3 * k + 1], a2 = u[3 * k + 2];
                float sum = mh_torch_inner_sum(K, [&](int t) { return fabsf((a0 * u[3 * t] + a1 * u[3 * t + 1]) + a2 * u[3 * t + 2]); });
                sum = sum / (float)K;
                if (sum > best) { best = sum; bk = k; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
            }
            best_id = s_idx[bk];
            __syncwarp();
        } else {
            // crowded voxel: same algorithm through global memory
            for (int a = lane; a < K; a += 32) {
                const int v = bucket[base + a];
                int rank = 0;
                for (int b = 0; b < K; ++b) rank += (bucket[base + b] < v) ? 1 : 0;
                sorted[base + rank] = v;
            }
            __syncwarp();
            float best = -1e30f; bk = 0x7fffffff;
            for (int k = lane; k < K; k += 32) {
                float a0, a1, a2;
                load_dir(dirs, sorted[base + k], a0, a1, a2);
                const float na = fmaxf(mh_norm3(a0, a1, a2), 1e-8f);
                a0 = a0 / na; a1 = a1 / na; a2 = a2 / na;
                float sum = mh_torch_inner_sum(K, [&](int t) {
                    float b0, b1, b2;
                    load_dir(dirs, sorted[base + t], b0, b1, b2);
                    const float nb = fmaxf(mh_norm3(b0, b1, b2), 1e-8f);
                    return fabsf((a0 * (b0 / nb) + a1 * (b1 / nb)) + a2 * (b2 / nb)); });
                sum = sum / (float)K;
                if (sum > best) { best = sum; bk = k; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
            }
            best_id = sorted[base + bk];
            __syncwarp();
        }
        if (lane == 0) {
            float o0, o1, o2;
            load_dir(dirs, best_id, o0, o1, o2);
            volume[g] = make_float4(o0, -o1, -o2, 1.0f);
        }
    }
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

// workspace: [hdr 64 B: #occupied, max count, bucket cursor][cnt nvox][key n][rank n][occ_list n][cnt_list n][bucket n][sorted n]
extern "C" int64_t mh_voxel_fuse_workspace_bytes(int64_t n, int32_t gx, int32_t gy, int32_t gz) {
    const int64_t nvox = (int64_t)gx * gy * gz;
    return 64 + 4 * (nvox + 6 * n + 16);
}

namespace {
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
}  // namespace

extern "C" int mh_voxel_fuse(void* stream, const float* points, const float* dirs, int64_t n,
                             const double* voxel_min_host, double voxel_size, int32_t gx, int32_t gy, int32_t gz,
                             void* volume, int32_t* vox_index, void* workspace, int64_t workspace_bytes) {
    MH_CHECK_ARG(volume && workspace && voxel_min_host && (n == 0 || (points && dirs)), "null pointer");
    MH_CHECK_ARG(gx > 0 && gy > 0 && gz > 0 && voxel_size > 0, "bad grid");
    MH_CHECK_ARG((int64_t)gx * gy * gz < (1ll << 31) && n < (1ll << 31) - 1, "grid or point count too large for int32 keys");
    MH_CHECK_ARG(workspace_bytes >= mh_voxel_fuse_workspace_bytes(n, gx, gy, gz), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nvox = (int64_t)gx * gy * gz;
    const VGrid g = make_grid(voxel_min_host, voxel_size, gx, gy, gz);
    int* hdr = reinterpret_cast<int*>(workspace);
    int* cnt = hdr + 16;
    int* key = cnt + nvox;
    int* rank = key + n;
    int* occ_list = rank + n;
    int* cnt_list = occ_list + n;
    int* bucket = cnt_list + n;
    int* sorted = bucket + n;
    // The 16 B/voxel zero fill of the volume (the bandwidth-bound part) runs on an auxiliary stream, concurrently
    // with the latency-bound bucket construction; the two join before the per-voxel results are written.
    AuxStream& ax = aux_stream();
    cudaEventRecord(ax.fork, st);
    cudaStreamWaitEvent(ax.s, ax.fork, 0);
    cudaMemsetAsync(volume, 0, sizeof(float4) * nvox, ax.s);
    cudaEventRecord(ax.join, ax.s);
    cudaMemsetAsync(hdr, 0, 64, st);
    if (n > 0) {
        cudaMemsetAsync(cnt, 0, sizeof(int) * nvox, st);
        const unsigned nb = (unsigned)((n + 255) / 256);
        count_kernel_vf<<<nb, 256, 0, st>>>(g, points, n, cnt, key, rank, occ_list, hdr, vox_index);
        MH_COUNT_LAUNCH();
        bucket_kernel<<<(unsigned)std::min<int64_t>(nb, (int64_t)mh_sm_count() * 8), 256, 0, st>>>(occ_list, cnt, cnt_list, hdr);
        MH_COUNT_LAUNCH();
        scatter_kernel<<<nb, 256, 0, st>>>(n, key, rank, cnt, bucket);
        MH_COUNT_LAUNCH();
    }
    cudaStreamWaitEvent(st, ax.join, 0);
    if (n > 0) {
        int64_t blocks = (n + FUSE_WARPS - 1) / FUSE_WARPS;
        const int64_t cap = (int64_t)mh_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        fuse_medoid_kernel<<<(unsigned)blocks, FUSE_WARPS * 32, 0, st>>>(occ_list, cnt_list, cnt, hdr, bucket, sorted, dirs,
                                                                        reinterpret_cast<float4*>(volume));
        MH_COUNT_LAUNCH();
    }
    MH_CHECK_LAUNCH();
    return 0;
}

/* after synchronising the stream: the largest number of points that fell into one voxel in the last mh_voxel_fuse
 * call on this workspace (informational: crowded voxels take the global-memory path of fuse_medoid_kernel). */
extern "C" int mh_voxel_fuse_max_points(const void* workspace, int32_t* max_k_host) {
    MH_CHECK_ARG(workspace && max_k_host, "null pointer");
    cudaError_t e = cudaMemcpy(max_k_host, reinterpret_cast<const int*>(workspace) + 1, 4, cudaMemcpyDeviceToHost);
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
