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

// Pass 1: voxel key of every point and a per-voxel linked list threaded through next[]; the list head lives in the
// (zeroed) volume itself, in the .w slot of the voxel's float4, as the integer id+1 of the last point inserted.
__global__ void __launch_bounds__(256)
link_kernel(VGrid g, const float* __restrict__ pts, int64_t n, float4* __restrict__ volume, int* __restrict__ next,
            int* __restrict__ key, int* __restrict__ vox_index) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x, y, z;
    p2v(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], x, y, z);
    const int k = (z * g.gy + y) * g.gx + x;
    key[i] = k;
    if (vox_index) vox_index[i] = (x * g.gy + y) * g.gz + z;
    next[i] = atomicExch(reinterpret_cast<int*>(volume + k) + 3, (int)i + 1);
}

// Pass 2: the point that is its voxel's list head resolves the voxel.  Up to 4 points are handled in registers
// (ids sorted back into original order, medoid under |cos| -- for K <= 4 torch.mean's order is the plain sequential
// sum); larger voxels go to a work list for the warp-per-voxel kernel.
__global__ void __launch_bounds__(256)
head_kernel(int64_t n, const int* __restrict__ key, const int* __restrict__ next, const float* __restrict__ dirs,
            float4* __restrict__ volume, int* __restrict__ worklist, int* __restrict__ wl_count, int* __restrict__ max_k) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = key[i];
    if (__float_as_int(volume[k].w) != (int)i + 1) return;
    int id[4] = {0, 0, 0, 0};
    int cnt = 0;
    for (int j = (int)i + 1; j != 0; j = next[j - 1]) { if (cnt < 4) id[cnt] = j - 1; ++cnt; }
    if (cnt > 4) {
        worklist[atomicAdd(wl_count, 1)] = k;
        atomicMax(max_k, cnt);
        return;
    }
    // sort ids ascending (original point order decides argmax ties)
    for (int a = 1; a < cnt; ++a) for (int b = a; b > 0 && id[b - 1] > id[b]; --b) { const int t_ = id[b]; id[b] = id[b - 1]; id[b - 1] = t_; }
    float u[4][3], raw[4][3];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        if (a < cnt) {
            load_dir(dirs, id[a], raw[a][0], raw[a][1], raw[a][2]);
            const float nn = fmaxf(mh_norm3(raw[a][0], raw[a][1], raw[a][2]), 1e-8f);
            u[a][0] = raw[a][0] / nn; u[a][1] = raw[a][1] / nn; u[a][2] = raw[a][2] / nn;
        } else { u[a][0] = u[a][1] = u[a][2] = raw[a][0] = raw[a][1] = raw[a][2] = 0.0f; }
    }
    int bk = 0;
    if (cnt > 1) {
        float best = -1e30f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (a < cnt) {
                float sum = 0.0f;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (b < cnt) sum += fabsf((u[a][0] * u[b][0] + u[a][1] * u[b][1]) + u[a][2] * u[b][2]);
                sum = sum / (float)cnt;
                if (sum > best) { best = sum; bk = a; }
            }
        }
    }
    float o0 = raw[0][0], o1 = raw[0][1], o2 = raw[0][2];
#pragma unroll
    for (int a = 1; a < 4; ++a) if (bk == a) { o0 = raw[a][0]; o1 = raw[a][1]; o2 = raw[a][2]; }
    volume[k] = make_float4(o0, -o1, -o2, 1.0f);
}

// Pass 3: one warp per voxel with more than 4 points: ids gathered from the list, rank-sorted into original order,
// medoid under |cos| with torch.mean's summation order (compute_points_similarity, PMVO_utils.py:366-382).
constexpr int FUSE_WARPS = 4, FUSE_MAXK = 1024;

__global__ void __launch_bounds__(FUSE_WARPS * 32)
fuse_medoid_kernel(const int* __restrict__ worklist, const int* __restrict__ wl_count, const int* __restrict__ next,
                   const float* __restrict__ dirs, float4* __restrict__ volume) {
    extern __shared__ __align__(16) unsigned char fm_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* s_raw = reinterpret_cast<int*>(fm_smem) + (size_t)warp * FUSE_MAXK * 5;      // [MAXK] unsorted ids
    int* s_idx = s_raw + FUSE_MAXK;                                                  // [MAXK] sorted ids
    float* u = reinterpret_cast<float*>(s_idx + FUSE_MAXK);                           // [MAXK][3]
    const int nwl = *wl_count;
    for (int w = blockIdx.x * FUSE_WARPS + warp; w < nwl; w += gridDim.x * FUSE_WARPS) {
        const int g = worklist[w];
        int K = 0;
        if (lane == 0) {
            for (int j = __float_as_int(volume[g].w); j != 0 && K < FUSE_MAXK; j = next[j - 1]) s_raw[K++] = j - 1;
        }
        K = __shfl_sync(0xffffffffu, K, 0);
        __syncwarp();
        for (int a = lane; a < K; a += 32) {
            const int v = s_raw[a];
            int rank = 0;
            for (int b = 0; b < K; ++b) rank += (s_raw[b] < v) ? 1 : 0;
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
        float best = -1e30f; int bk = 0x7fffffff;
        for (int k = lane; k < K; k += 32) {
            const float a0 = u[3 * k], a1 = u[3 * k + 1], a2 = u[3 * k + 2];
            float sum = mh_torch_inner_sum(K, [&](int j) { return fabsf((a0 * u[3 * j] + a1 * u[3 * j + 1]) + a2 * u[3 * j + 2]); });
            sum = sum / (float)K;
            if (sum > best) { best = sum; bk = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
        }
        if (lane == 0) {
            float o0, o1, o2;
            load_dir(dirs, s_idx[bk], o0, o1, o2);
            volume[g] = make_float4(o0, -o1, -o2, 1.0f);
        }
        __syncwarp();
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

// workspace: [hdr 64 B: wl_count, max_k][next n][key n][worklist n/5+4]
extern "C" int64_t mh_voxel_fuse_workspace_bytes(int64_t n, int32_t gx, int32_t gy, int32_t gz) {
    (void)gx; (void)gy; (void)gz;
    return 64 + 4 * (2 * n + n / 5 + 8);
}

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
    int* next = hdr + 16;
    int* key = next + n;
    int* worklist = key + n;
    cudaMemsetAsync(hdr, 0, 64, st);
    cudaMemsetAsync(volume, 0, sizeof(float4) * nvox, st);            // empty voxels: the streaming 16 B/voxel write
    if (n == 0) return 0;
    link_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, points, n, reinterpret_cast<float4*>(volume), next, key, vox_index);
    MH_COUNT_LAUNCH();
    head_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, key, next, dirs, reinterpret_cast<float4*>(volume), worklist, hdr, hdr + 1);
    MH_COUNT_LAUNCH();
    const size_t smem = (size_t)FUSE_WARPS * FUSE_MAXK * 5 * sizeof(int);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(fuse_medoid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    int64_t blocks = (n / 5 + FUSE_WARPS) / FUSE_WARPS;
    const int64_t cap = (int64_t)mh_sm_count() * 2;
    if (blocks > cap) blocks = cap;
    fuse_medoid_kernel<<<(unsigned)blocks, FUSE_WARPS * 32, smem, st>>>(worklist, hdr, next, dirs, reinterpret_cast<float4*>(volume));
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

/* after synchronising the stream: the largest number of points that fell into one voxel in the last mh_voxel_fuse
 * call on this workspace; voxels with more than 1024 points are fused from their first 1024 (in list order). */
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
