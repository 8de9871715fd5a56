// HairGrow strand tracing through the fused volume (HairGrow.py:59-299).
//
// Volume: float4 [gz][gy][gx] = {ori.x, ori.y, ori.z, occ} in HairGrowing's frame (one 16 B fetch per step).
// The reference trace is a NEAREST-voxel walk with unit steps (position += orientation), not a trilinear one
// (SURVEY.md §9-R15).  Each walk is a chain of dependent fetches, so the kernels run one thread per
// (seed, direction) and rely on many resident walks to cover the L2/HBM latency; strand geometry does not depend
// on the `flag` volume (§9-R9), which is handled afterwards by the ordered acceptance pass mh_accept_strands.
// Bound: memory latency / random 32 B sector rate, not streaming bandwidth.  Algorithmic bytes per step = 16.
#include "mh_common.cuh"

namespace {

struct Vol { const float4* v; int gx, gy, gz; };

// .type(torch.long) truncates toward zero, then clamp (HairGrow.py:66-69, §9-R10)
__device__ __forceinline__ int vox_index(const Vol& g, float x, float y, float z) {
    const int ix = min(max((int)x, 0), g.gx - 1), iy = min(max((int)y, 0), g.gy - 1), iz = min(max((int)z, 0), g.gz - 1);
    return (iz * g.gy + iy) * g.gx + ix;
}
// torch.dot of 3-vectors on CPU: (a0*b0 + a1*b1) + a2*b2, products rounded separately (probe)
__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return (a0 * b0 + a1 * b1) + a2 * b2;
}

// One direction of HairGrowing.trace (HairGrow.py:78-105 forward, :116-143 backward with sign=-1).
// Emit(k, x,y,z) is called for the k-th accepted step (k = 0,1,...).  Returns the number of accepted steps.
template <typename Emit>
__device__ __forceinline__ int walk(const Vol& g, float px, float py, float pz, float sign, float thr, int max_steps, Emit emit) {
    float4 cur = __ldg(g.v + vox_index(g, px, py, pz));
    float tx = cur.x, ty = cur.y, tz = cur.z;
    int count = 0;
    for (;;) {
        if (cur.w == 0.0f) break;                                   // occ of the current voxel
        const float nx = px + sign * tx, ny = py + sign * ty, nz = pz + sign * tz;
        const float4 nxt = __ldg(g.v + vox_index(g, nx, ny, nz));
        if (dot3(nxt.x, nxt.y, nxt.z, tx, ty, tz) < thr) break;
        px = nx; py = ny; pz = nz;
        tx = nxt.x; ty = nxt.y; tz = nxt.z;
        cur = nxt;
        emit(count, px, py, pz);
        if (++count >= max_steps) break;
    }
    return count;
}

__global__ void trace_count_kernel(Vol g, const float* __restrict__ seeds, int64_t n, float thr, int max_steps,
                                   int* __restrict__ n_fwd, int* __restrict__ n_bwd) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    const int64_t i = t >> 1;
    const bool bwd = t & 1;
    const int c = walk(g, seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2], bwd ? -1.0f : 1.0f, thr, max_steps,
                       [](int, float, float, float) {});
    (bwd ? n_bwd : n_fwd)[i] = c;
}

__global__ void trace_write_kernel(Vol g, const float* __restrict__ seeds, int64_t n, float thr, int max_steps,
                                   const int* __restrict__ n_fwd, const int* __restrict__ n_bwd,
                                   const int64_t* __restrict__ offsets, int min_len, float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    const int64_t i = t >> 1;
    const bool bwd = t & 1;
    const int nb = n_bwd[i], nf = n_fwd[i];
    if (nb + nf + 1 < min_len) return;
    float* base = out + 3 * offsets[i];
    const float sx = seeds[3 * i], sy = seeds[3 * i + 1], sz = seeds[3 * i + 2];
    if (bwd) {
        // strand.insert(0, .): backward step k lands at position nb-1-k
        walk(g, sx, sy, sz, -1.0f, thr, max_steps, [=](int k, float x, float y, float z) {
            float* p = base + 3 * (nb - 1 - k); p[0] = x; p[1] = y; p[2] = z; });
    } else {
        float* p0 = base + 3 * nb; p0[0] = sx; p0[1] = sy; p0[2] = sz;
        walk(g, sx, sy, sz, 1.0f, thr, max_steps, [=](int k, float x, float y, float z) {
            float* p = base + 3 * (nb + 1 + k); p[0] = x; p[1] = y; p[2] = z; });
    }
}

// HairGrowing.traceFromScalp (HairGrow.py:154-223)
__global__ void trace_scalp_kernel(Vol g, const float* __restrict__ roots, const float* __restrict__ normals, int64_t n,
                                   float thr, int max_steps, int max_inner, float* __restrict__ out, int* __restrict__ length) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float px = roots[3 * i], py = roots[3 * i + 1], pz = roots[3 * i + 2];
    const float n0 = normals[3 * i], n1 = normals[3 * i + 1], n2 = normals[3 * i + 2];
    // d = (0,1,0); m = min(dot(normal,d)+1, 1); dot = (n0*0 + n1*1) + n2*0 = n1 exactly for finite inputs
    const float m = fminf(dot3(n0, n1, n2, 0.0f, 1.0f, 0.0f) + 1.0f, 1.0f);
    float tx = n0 + 0.0f * m, ty = n1 + 1.0f * m, tz = n2 + 0.0f * m;
    { const float nn = mh_norm3(tx, ty, tz); tx = tx / nn; ty = ty / nn; tz = tz / nn; }
    float* o = out + (size_t)i * 3 * (max_steps + 1);
    o[0] = px; o[1] = py; o[2] = pz;
    int count = 0;
    bool inner = true;
    float occ = __ldg(g.v + vox_index(g, px, py, pz)).w;
    for (;;) {
        if (occ == 0.0f && !inner) break;
        const float nx = px + tx, ny = py + ty, nz = pz + tz;
        const float4 nxt = __ldg(g.v + vox_index(g, nx, ny, nz));
        float ax = nxt.x, ay = nxt.y, az = nxt.z;
        if (mh_norm3(ax, ay, az) < 0.1f && inner) {
            if (dot3(tx, ty, tz, n0, n1, n2) < 0.85f) { ax = tx; ay = ty; az = tz; }
            else {
                ax = tx + 0.0f * m; ay = ty + 1.0f * m; az = tz + 0.0f * m;
                const float nn = mh_norm3(ax, ay, az);
                ax = ax / nn; ay = ay / nn; az = az / nn;
            }
        } else {
            if (dot3(ax, ay, az, tx, ty, tz) < thr && !inner) {
                if (dot3(-ax, -ay, -az, tx, ty, tz) < thr) break;
                ax = -ax; ay = -ay; az = -az;
            }
            if (dot3(ax, ay, az, tx, ty, tz) < 0.0f && inner) { ax = -ax; ay = -ay; az = -az; }
            inner = false;
        }
        px = nx; py = ny; pz = nz;
        tx = ax; ty = ay; tz = az;
        occ = nxt.w;
        ++count;
        o[3 * count] = px; o[3 * count + 1] = py; o[3 * count + 2] = pz;
        if (count >= max_steps) break;
        if (count >= max_inner && inner) break;
    }
    length[i] = inner ? 0 : count + 1;
}

// Ordered acceptance: a single warp walks the strands in order (the reference's sequential flag logic).
// mode 0: gate on flag[seed voxel] >= 3, then flag[voxels] += 1 once per unique voxel (gather-all then scatter-all,
//         which is what `flag[idx] += 1` does with duplicate indices);  mode 1: no gate, flag[voxels] = 1.
// ---- acceptance pass (GenerateGuideStrandFromScalp / randomlyGenerateSegments flag logic, HairGrow.py:235-260, 280-293)
// Reference: strands in seed order; strand i is kept iff it exists (>= 5 points) and flag[seed voxel_i] < 3 at that
// moment; a kept strand bumps flag once per unique voxel it visits.  The decision of i therefore depends on the kept
// strands j < i that run through i's seed voxel.
//
// mode 1 (scalp roots): no gate, flag = 1 on visited voxels -> order-free, fully parallel (accept_all_kernel).
// mode 0: batches of AB consecutive strands, exact within and across batches (accept_batch_kernel, one CTA):
//   1. base_i = flag[seed voxel_i] as left by the previous batches; candidates = existing strands with base_i < 3;
//      their seed voxels go into a shared-memory hash;
//   2. every candidate j probes the hash with the voxels it visits: a hit on the seed voxel of a LATER candidate i
//      sets bit j of dep_i (AB-bit rows in shared memory);
//   3. candidates with an empty dep row are kept outright; the others are resolved in index order by one warp:
//      kept_i = base_i + popc(dep_i & kept) < 3  (dep_i only holds earlier strands, whose fate is known by then);
//   4. kept strands add 1 to flag at each unique voxel they visit (first occurrence within the strand, atomics).
// This is the sequential loop's result bit for bit; the sequential part shrinks to step 3 (a few shared-memory
// words per contested strand).
constexpr int AB = 512, AB_WORDS = AB / 32, AHT = 2048, A_THREADS = 1024, A_MAXLEN = 520;

__global__ void accept_all_kernel(const float* __restrict__ pts, const int64_t* __restrict__ offsets,
                                  const int* __restrict__ lengths, int64_t n, int gx, int gy, int gz,
                                  float* __restrict__ flag, uint8_t* __restrict__ accepted) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    Vol g; g.v = nullptr; g.gx = gx; g.gy = gy; g.gz = gz;
    const int len = lengths[i];
    if (lane == 0) accepted[i] = len > 0 ? 1 : 0;
    const float* p = pts + 3 * offsets[i];
    for (int k = lane; k < len; k += 32) flag[vox_index(g, p[3 * k], p[3 * k + 1], p[3 * k + 2])] = 1.0f;
}

__device__ __forceinline__ unsigned accept_hash(int v) { return ((unsigned)v * 2654435761u) >> 21; }      // 11 bits = AHT

__global__ void __launch_bounds__(A_THREADS)
accept_batch_kernel(const float* __restrict__ pts, const int64_t* __restrict__ offsets, const int* __restrict__ lengths,
                    const float* __restrict__ seeds, int64_t n, int gx, int gy, int gz, float* flag,
                    uint8_t* __restrict__ accepted) {
    extern __shared__ __align__(16) unsigned char a_smem[];
    int* hkey = reinterpret_cast<int*>(a_smem);                   // [AHT] seed voxel or -1
    int* hhead = hkey + AHT;                                      // [AHT] first candidate with that seed voxel
    int* snext = hhead + AHT;                                     // [AB]  chain of candidates sharing a seed voxel
    int* slen = snext + AB;                                       // [AB]
    float* sbase = reinterpret_cast<float*>(slen + AB);           // [AB]  flag at the seed voxel when the batch starts
    unsigned* dep = reinterpret_cast<unsigned*>(sbase + AB);      // [AB][AB_WORDS]
    unsigned* keptm = dep + AB * AB_WORDS;                        // [AB_WORDS] kept strands of the batch
    unsigned* contm = keptm + AB_WORDS;                           // [AB_WORDS] contested candidates (non-empty dep row)
    int* vbuf = reinterpret_cast<int*>(contm + AB_WORDS);         // [32][A_MAXLEN] voxel ids of the strand a warp works on
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Vol g; g.v = nullptr; g.gx = gx; g.gy = gy; g.gz = gz;
    for (int64_t b0 = 0; b0 < n; b0 += AB) {
        const int nb = (int)min((int64_t)AB, n - b0);
        for (int k = tid; k < AHT; k += A_THREADS) { hkey[k] = -1; hhead[k] = -1; }
        for (int k = tid; k < AB * AB_WORDS; k += A_THREADS) dep[k] = 0u;
        if (tid < AB_WORDS) { keptm[tid] = 0u; contm[tid] = 0u; }
        __syncthreads();
        // 1. candidates and their seed voxels
        bool cand = false;
        if (tid < nb) {
            const int64_t i = b0 + tid;
            const int len = lengths[i];
            const int sv = vox_index(g, seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
            const float base = __ldcg(flag + sv);                  // L2: earlier batches updated it with atomics
            slen[tid] = len;
            sbase[tid] = base;
            cand = len > 0 && !(base >= 3.0f);
            if (cand) {
                unsigned h = accept_hash(sv);
                for (;;) {
                    const int old = atomicCAS(&hkey[h], -1, sv);
                    if (old == -1 || old == sv) { snext[tid] = atomicExch(&hhead[h], tid); break; }
                    h = (h + 1) & (AHT - 1);
                }
            }
        }
        __syncthreads();
        // 2. dependencies: candidate j runs through the seed voxel of a later candidate i
        for (int j = warp; j < nb; j += A_THREADS / 32) {
            const bool jc = slen[j] > 0 && !(sbase[j] >= 3.0f);
            if (!jc) continue;                                     // warp-uniform
            const float* p = pts + 3 * offsets[b0 + j];
            const int len = slen[j];
            for (int k = lane; k < len; k += 32) {
                const int v = vox_index(g, p[3 * k], p[3 * k + 1], p[3 * k + 2]);
                unsigned h = accept_hash(v);
                for (;;) {
                    const int key = hkey[h];
                    if (key == -1) break;
                    if (key == v) {
                        for (int i = hhead[h]; i != -1; i = snext[i])
                            if (i > j) atomicOr(&dep[i * AB_WORDS + (j >> 5)], 1u << (j & 31));
                        break;
                    }
                    h = (h + 1) & (AHT - 1);
                }
            }
        }
        __syncthreads();
        // 3. uncontested candidates are kept outright; contested ones are resolved in index order by warp 0
        if (tid < nb && cand) {
            unsigned any = 0u;
#pragma unroll
            for (int w = 0; w < AB_WORDS; ++w) any |= dep[tid * AB_WORDS + w];
            atomicOr(any ? &contm[tid >> 5] : &keptm[tid >> 5], 1u << (tid & 31));
        }
        __syncthreads();
        if (warp == 0) {
            for (int w = 0; w < AB_WORDS; ++w) {
                unsigned word = contm[w];
                while (word) {
                    const int b = __ffs(word) - 1;
                    const int i = w * 32 + b;
                    int c = (lane < AB_WORDS) ? __popc(dep[i * AB_WORDS + lane] & keptm[lane]) : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    if (lane == 0 && sbase[i] + (float)c < 3.0f) keptm[w] |= 1u << b;
                    __syncwarp();
                    word &= word - 1;
                }
            }
        }
        __syncthreads();
        // 4. results; kept strands bump the flag once per unique voxel
        if (tid < nb) accepted[b0 + tid] = (keptm[tid >> 5] >> (tid & 31)) & 1u;
        int* vb = vbuf + warp * A_MAXLEN;
        for (int j = warp; j < nb; j += A_THREADS / 32) {
            if (!((keptm[j >> 5] >> (j & 31)) & 1u)) continue;    // warp-uniform
            const float* p = pts + 3 * offsets[b0 + j];
            const int len = min(slen[j], A_MAXLEN);
            for (int k = lane; k < len; k += 32) vb[k] = vox_index(g, p[3 * k], p[3 * k + 1], p[3 * k + 2]);
            __syncwarp();
            for (int k = lane; k < len; k += 32) {
                const int v = vb[k];
                bool first = true;
                for (int k2 = k - 1; k2 >= 0; --k2) if (vb[k2] == v) { first = false; break; }
                if (first) atomicAdd(flag + v, 1.0f);
            }
            __syncwarp();
        }
        __threadfence();
        __syncthreads();
    }
}

// ---- the ordered acceptance over all SMs ------------------------------------------------------------------------------
// Only two things in a batch depend on the batches before it: the flag at each seed voxel when the batch starts, and which
// strands end up kept.  Everything else -- the seed-voxel hash, the dependency rows (strand j runs through the seed voxel
// of a later strand i) and which points of a strand are the first visit of their voxel -- depends on geometry alone, so
// it is computed for ALL batches at once, one CTA per batch (accept_prepare_kernel), for the superset "every traced
// strand" (a strand that turns out not to be a candidate is never kept, so its edges count for nothing).  What stays
// sequential (accept_resolve_kernel, one CTA) is per batch: 512 flag reads, the resolution of the contested strands in
// index order, and the flag bumps of the kept ones.
__global__ void __launch_bounds__(A_THREADS)
accept_prepare_kernel(const float* __restrict__ pts, const int64_t* __restrict__ offsets, const int* __restrict__ lengths,
                      const float* __restrict__ seeds, int64_t n, int gx, int gy, int gz, unsigned* __restrict__ depg,
                      uint8_t* __restrict__ firstocc) {
    extern __shared__ __align__(16) unsigned char a_smem[];
    int* hkey = reinterpret_cast<int*>(a_smem);                   // [AHT] seed voxel or -1
    int* hhead = hkey + AHT;                                      // [AHT] first strand with that seed voxel
    int* snext = hhead + AHT;                                     // [AB]
    int* slen = snext + AB;                                       // [AB]
    unsigned* dep = reinterpret_cast<unsigned*>(slen + AB);       // [AB][AB_WORDS]
    int* vbuf = reinterpret_cast<int*>(dep + AB * AB_WORDS);      // [32][A_MAXLEN]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Vol g; g.v = nullptr; g.gx = gx; g.gy = gy; g.gz = gz;
    const int64_t b0 = (int64_t)blockIdx.x * AB;
    const int nb = (int)min((int64_t)AB, n - b0);
    for (int k = tid; k < AHT; k += A_THREADS) { hkey[k] = -1; hhead[k] = -1; }
    for (int k = tid; k < AB * AB_WORDS; k += A_THREADS) dep[k] = 0u;
    __syncthreads();
    if (tid < nb) {
        const int64_t i = b0 + tid;
        const int len = lengths[i];
        slen[tid] = len;
        if (len > 0) {
            const int sv = vox_index(g, seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
            unsigned h = accept_hash(sv);
            for (;;) {
                const int old = atomicCAS(&hkey[h], -1, sv);
                if (old == -1 || old == sv) { snext[tid] = atomicExch(&hhead[h], tid); break; }
                h = (h + 1) & (AHT - 1);
            }
        }
    }
    __syncthreads();
    int* vb = vbuf + warp * A_MAXLEN;
    for (int j = warp; j < nb; j += A_THREADS / 32) {
        const int len = slen[j];
        if (len <= 0) continue;                                    // warp-uniform
        const float* p = pts + 3 * offsets[b0 + j];
        for (int k = lane; k < len; k += 32) {
            const int v = vox_index(g, p[3 * k], p[3 * k + 1], p[3 * k + 2]);
            if (k < A_MAXLEN) vb[k] = v;
            unsigned h = accept_hash(v);
            for (;;) {
                const int key = hkey[h];
                if (key == -1) break;
                if (key == v) {
                    for (int i = hhead[h]; i != -1; i = snext[i])
                        if (i > j) atomicOr(&dep[i * AB_WORDS + (j >> 5)], 1u << (j & 31));
                    break;
                }
                h = (h + 1) & (AHT - 1);
            }
        }
        __syncwarp();
        uint8_t* fo = firstocc + offsets[b0 + j];
        const int l2 = min(len, A_MAXLEN);
        for (int k = lane; k < l2; k += 32) {
            const int v = vb[k];
            bool first = true;
            for (int k2 = k - 1; k2 >= 0; --k2) if (vb[k2] == v) { first = false; break; }
            fo[k] = first ? 1 : 0;
        }
        __syncwarp();
    }
    __syncthreads();
    unsigned* dg = depg + (size_t)blockIdx.x * AB * AB_WORDS;
    for (int k = tid; k < AB * AB_WORDS; k += A_THREADS) dg[k] = dep[k];
}

__global__ void __launch_bounds__(A_THREADS)
accept_resolve_kernel(const float* __restrict__ pts, const int64_t* __restrict__ offsets, const int* __restrict__ lengths,
                      const float* __restrict__ seeds, int64_t n, int gx, int gy, int gz, const unsigned* __restrict__ depg,
                      const uint8_t* __restrict__ firstocc, float* flag, uint8_t* __restrict__ accepted) {
    __shared__ int slen[AB];
    __shared__ float sbase[AB];
    __shared__ unsigned candm[AB_WORDS], keptm[AB_WORDS], contm[AB_WORDS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Vol g; g.v = nullptr; g.gx = gx; g.gy = gy; g.gz = gz;
    int64_t bi = 0;
    for (int64_t b0 = 0; b0 < n; b0 += AB, ++bi) {
        const int nb = (int)min((int64_t)AB, n - b0);
        const unsigned* dg = depg + (size_t)bi * AB * AB_WORDS;
        if (tid < AB_WORDS) { candm[tid] = 0u; keptm[tid] = 0u; contm[tid] = 0u; }
        __syncthreads();
        bool cand = false;
        if (tid < nb) {
            const int64_t i = b0 + tid;
            const int len = lengths[i];
            const int sv = vox_index(g, seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
            const float base = __ldcg(flag + sv);                  // L2: earlier batches updated it with atomics
            slen[tid] = len;
            sbase[tid] = base;
            cand = len > 0 && !(base >= 3.0f);
            if (cand) atomicOr(&candm[tid >> 5], 1u << (tid & 31));
        }
        __syncthreads();
        if (tid < nb && cand) {
            unsigned any = 0u;
#pragma unroll
            for (int w = 0; w < AB_WORDS; ++w) any |= dg[tid * AB_WORDS + w] & candm[w];
            atomicOr(any ? &contm[tid >> 5] : &keptm[tid >> 5], 1u << (tid & 31));
        }
        __syncthreads();
        if (warp == 0) {                                            // contested candidates, in index order
            for (int w = 0; w < AB_WORDS; ++w) {
                unsigned word = contm[w];
                while (word) {
                    const int b = __ffs(word) - 1;
                    const int i = w * 32 + b;
                    int c = (lane < AB_WORDS) ? __popc(dg[i * AB_WORDS + lane] & keptm[lane]) : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    if (lane == 0 && sbase[i] + (float)c < 3.0f) keptm[w] |= 1u << b;
                    __syncwarp();
                    word &= word - 1;
                }
            }
        }
        __syncthreads();
        if (tid < nb) accepted[b0 + tid] = (keptm[tid >> 5] >> (tid & 31)) & 1u;
        for (int j = warp; j < nb; j += A_THREADS / 32) {           // kept strands bump the flag once per unique voxel
            if (!((keptm[j >> 5] >> (j & 31)) & 1u)) continue;    // warp-uniform
            const float* p = pts + 3 * offsets[b0 + j];
            const uint8_t* fo = firstocc + offsets[b0 + j];
            const int len = min(slen[j], A_MAXLEN);
            for (int k = lane; k < len; k += 32)
                if (fo[k]) atomicAdd(flag + vox_index(g, p[3 * k], p[3 * k + 1], p[3 * k + 2]), 1.0f);
        }
        __threadfence();
        __syncthreads();
    }
}

}  // namespace

static int check_vol(const void* volume, int gx, int gy, int gz) {
    MH_CHECK_ARG(volume && gx > 0 && gy > 0 && gz > 0, "bad volume");
    return 0;
}

extern "C" int mh_trace_count(void* stream, const void* volume, int32_t gx, int32_t gy, int32_t gz, const float* seeds,
                              int64_t n, float thr_dot, int32_t max_steps, int32_t* n_fwd, int32_t* n_bwd) {
    if (n == 0) return 0;
    if (check_vol(volume, gx, gy, gz)) return 1;
    MH_CHECK_ARG(seeds && n_fwd && n_bwd && max_steps > 0, "bad arguments");
    Vol g{reinterpret_cast<const float4*>(volume), gx, gy, gz};
    trace_count_kernel<<<(unsigned)((2 * n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g, seeds, n, thr_dot, max_steps, n_fwd, n_bwd);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_trace_write(void* stream, const void* volume, int32_t gx, int32_t gy, int32_t gz, const float* seeds,
                              int64_t n, float thr_dot, int32_t max_steps, const int32_t* n_fwd, const int32_t* n_bwd,
                              const int64_t* offsets, int32_t min_len, float* points_out) {
    if (n == 0) return 0;
    if (check_vol(volume, gx, gy, gz)) return 1;
    MH_CHECK_ARG(seeds && n_fwd && n_bwd && offsets && points_out && max_steps > 0, "bad arguments");
    Vol g{reinterpret_cast<const float4*>(volume), gx, gy, gz};
    trace_write_kernel<<<(unsigned)((2 * n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g, seeds, n, thr_dot, max_steps, n_fwd, n_bwd, offsets, min_len, points_out);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_trace_from_scalp(void* stream, const void* volume, int32_t gx, int32_t gy, int32_t gz, const float* roots,
                                   const float* normals, int64_t n, float thr_dot, int32_t max_steps, int32_t max_inner,
                                   float* points_out, int32_t* length) {
    if (n == 0) return 0;
    if (check_vol(volume, gx, gy, gz)) return 1;
    MH_CHECK_ARG(roots && normals && points_out && length && max_steps > 0, "bad arguments");
    Vol g{reinterpret_cast<const float4*>(volume), gx, gy, gz};
    trace_scalp_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g, roots, normals, n, thr_dot, max_steps, max_inner, points_out, length);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int64_t mh_accept_strands_workspace_bytes(int64_t n, int64_t total_points) {
    const int64_t nbatch = (n + AB - 1) / AB;
    return nbatch * AB * AB_WORDS * 4 + ((total_points + 15) / 16) * 16 + 256;
}

/* mode 0 with a workspace: the geometry-only part of every batch on all SMs, then the short sequential resolution */
extern "C" int mh_accept_strands_ws(void* stream, const float* points, const int64_t* offsets, const int32_t* lengths,
                                    const float* seeds, int64_t n, int64_t total_points, int32_t gx, int32_t gy, int32_t gz,
                                    float* flag, uint8_t* accepted, void* workspace, int64_t workspace_bytes) {
    if (n == 0) return 0;
    MH_CHECK_ARG(points && offsets && lengths && flag && accepted && seeds && workspace, "null pointer");
    MH_CHECK_ARG(gx > 0 && gy > 0 && gz > 0 && total_points >= 0, "bad arguments");
    MH_CHECK_ARG(workspace_bytes >= mh_accept_strands_workspace_bytes(n, total_points), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nbatch = (n + AB - 1) / AB;
    unsigned* depg = reinterpret_cast<unsigned*>(workspace);
    uint8_t* firstocc = reinterpret_cast<uint8_t*>(depg + nbatch * AB * AB_WORDS);
    const size_t smem = sizeof(int) * (2 * AHT + 2 * AB) + sizeof(unsigned) * (AB * AB_WORDS) + sizeof(int) * 32 * A_MAXLEN;
    cudaError_t e = cudaFuncSetAttribute(accept_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mh_set_error("mh_accept_strands_ws: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 2; }
    accept_prepare_kernel<<<(unsigned)nbatch, A_THREADS, smem, st>>>(points, offsets, lengths, seeds, n, gx, gy, gz, depg, firstocc);
    MH_COUNT_LAUNCH();
    accept_resolve_kernel<<<1, A_THREADS, 0, st>>>(points, offsets, lengths, seeds, n, gx, gy, gz, depg, firstocc, flag, accepted);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_accept_strands(void* stream, const float* points, const int64_t* offsets, const int32_t* lengths,
                                 const float* seeds, int64_t n, int32_t gx, int32_t gy, int32_t gz, int32_t mode,
                                 float* flag, uint8_t* accepted) {
    if (n == 0) return 0;
    MH_CHECK_ARG(points && offsets && lengths && flag && accepted && (mode == 1 || seeds), "null pointer");
    MH_CHECK_ARG(gx > 0 && gy > 0 && gz > 0 && (mode == 0 || mode == 1), "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1) {
        accept_all_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(points, offsets, lengths, n, gx, gy, gz, flag, accepted);
    } else {
        const size_t smem = sizeof(int) * (2 * AHT + 2 * AB) + sizeof(float) * AB + sizeof(unsigned) * (AB * AB_WORDS + 2 * AB_WORDS)
                            + sizeof(int) * 32 * A_MAXLEN;
        cudaError_t e = cudaFuncSetAttribute(accept_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { mh_set_error("mh_accept_strands: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 2; }
        accept_batch_kernel<<<1, A_THREADS, smem, st>>>(points, offsets, lengths, seeds, n, gx, gy, gz, flag, accepted);
    }
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
