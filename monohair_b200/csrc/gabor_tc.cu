// Gabor orientation bank on the 5th-generation tensor cores (tcgen05 + TMEM), fused with the per-pixel epilogue.
//
// calOrientationGabor.filter (GaborFilter.py:29-94) is the contraction [HW x 289] x [289 x 180] followed by a per-pixel
// reduction over the 180 responses (max / arg-max, then a variance weighted by the circular distance to the arg-max).
// Here one CTA owns two output rows x 128 pixels:
//   * the 180 |responses| of its 256 pixels never leave the chip: two 128-lane x 192-column fp32 accumulators in TMEM;
//   * K is walked one kernel row at a time (17 taps padded to 24 = 3 MMA K-steps of 8): the im2col operand of input row j
//     ("A_j": pixel m, tap c -> in[j][m + c]) is written once into shared memory and feeds BOTH output rows (kernel row
//     r = j for the upper one, r = j - 1 for the lower one);
//   * the bank ("B_r": 192 x 24, pre-split and pre-laid-out on the host) streams from L2 with TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx) through a 4-slot ring;
//   * fp32 accuracy from tf32 tensor cores by the 3-term split x = hi + lo: hi*hi + hi*lo + lo*hi (dropped lo*lo term
//     ~2^-22 relative), accumulated in fp32 -- the arg-max over 180 near-equal responses needs it (test_gpu_gabor.py
//     gates the comparison by the top-2 margin);
//   * warp 0 / one elected thread issues the MMAs and the bulk copies, warps 1-4 build the A operands, all 8 warps run
//     the epilogue: tcgen05.ld of the pixel's 192 columns twice (arg-max, then the ordered variance sum, torch's
//     cascade order as in gabor.cu), orientation / variance out, block maximum -> global maximum.
// The normalisation by the global maximum stays in gabor_finish_kernel (gabor.cu).
// SASS: UTCHMMA-family (tcgen05.mma kind::tf32), UBLKCP (TMA bulk), LDTM (tcgen05.ld).
#include <cstdio>
#include <cstdlib>
#include "mh_common.cuh"

namespace {

constexpr int TC_GK = 17, TC_PAD = 8;
constexpr int TC_M = 128;                         // pixels per MMA (TMEM lanes)
constexpr int TC_N = 192;                         // filters padded (TMEM columns per accumulator)
constexpr int TC_KROW = 24;                       // taps per kernel row, padded: 3 K-steps of 8
constexpr int TC_ROWS = 2;                        // output rows per CTA
constexpr int TC_INROWS = TC_GK + TC_ROWS - 1;    // 18 input rows
constexpr int TC_INW = TC_M + 2 * TC_PAD;         // 144 input columns
constexpr int TC_A_BYTES = TC_M * TC_KROW * 4;    // 12288: one split part of A_j
constexpr int TC_B_BYTES = TC_N * TC_KROW * 4;    // 18432: one split part of B_r
constexpr int TC_A_SLOTS = 3, TC_B_SLOTS = 4;
constexpr int TC_THREADS = 256;
constexpr uint32_t TC_TMEM_COLS = 512;

struct __align__(16) TcSmem {
    unsigned char a[TC_A_SLOTS][2][TC_A_BYTES];   // [slot][hi, lo]
    unsigned char b[TC_B_SLOTS][2][TC_B_BYTES];
    float in[TC_INROWS][TC_INW];
    unsigned long long a_full[TC_A_SLOTS], b_full[TC_B_SLOTS], step_done[4], all_done;
    uint32_t tmem_base;
    float red[8];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    printf("gabor_tc_kernel: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
    __trap();
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // K-major, no swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
    // version 1 [46,48), layout type 0 [61,64)
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(void* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// one lane of the converged warp (elect.sync): what the single-thread tcgen05 / TMA issue hangs on
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float theta_tc(int i, int n) { return (3.14159265358979323846f * (float)i) / (float)n; }

__global__ void __launch_bounds__(TC_THREADS, 1)
gabor_tc_kernel(const float* __restrict__ img, int H, int W, const unsigned char* __restrict__ bank_tc, int nf,
                float* __restrict__ orient, float* __restrict__ var, unsigned int* __restrict__ gmax, int npass) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * TC_M, y0 = blockIdx.y * TC_ROWS;

    // ---- setup: barriers, TMEM, input tile
    if (tid == 0) {
        for (int i = 0; i < TC_A_SLOTS; ++i) mbar_init(&s.a_full[i], 128);
        for (int i = 0; i < TC_B_SLOTS; ++i) mbar_init(&s.b_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&s.step_done[i], 1);
        mbar_init(&s.all_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem_base)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        // input tile: all loads in flight before the first store (18 x 144 values, 11 per thread)
        constexpr int PER = (TC_INROWS * TC_INW + TC_THREADS - 1) / TC_THREADS;
        float vin[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = tid + k * TC_THREADS;
            const int ty = i / TC_INW, tx = i - ty * TC_INW;
            const int gy = y0 + ty - TC_PAD, gx = x0 + tx - TC_PAD;
            vin[k] = (i < TC_INROWS * TC_INW && gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(img + (size_t)gy * W + gx) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = tid + k * TC_THREADS;
            if (i < TC_INROWS * TC_INW) (&s.in[0][0])[i] = vin[k];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;

    if (warp == 0) {
        // ================= MMA issuer + bank producer: the warp stays converged, one elected lane issues =================
        // idesc: D f32 [4,6)=1, A tf32 [7,10)=2, B tf32 [10,13)=2, K-major both, N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
        // descriptors ([k group of 4][row group of 8][8 rows][4 elements]: SBO = 128 B, LBO = 16 (A) / 24 (B) row groups):
        // only the start-address field changes between operands, so each is the slot-0 descriptor plus a 16 B-unit offset
        const uint64_t da0 = umma_desc(smem_u32(&s.a[0][0][0]), (TC_M / 8) * 128, 128);
        const uint64_t db0 = umma_desc(smem_u32(&s.b[0][0][0]), (TC_N / 8) * 128, 128);
        auto load_b = [&](int r) {
            const int slot = r % TC_B_SLOTS;
            mbar_expect_tx(&s.b_full[slot], 2 * TC_B_BYTES);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(&s.b[slot][0][0])), "l"(bank_tc + (size_t)r * 2 * TC_B_BYTES), "r"(2u * TC_B_BYTES),
                            "r"(smem_u32(&s.b_full[slot])) : "memory");
        };
        if (elect_one()) { load_b(0); load_b(1); }
        __syncwarp();
        uint32_t started0 = 0, started1 = 0;
        for (int j = 0; j < TC_INROWS; ++j) {
            mbar_wait(&s.a_full[j % TC_A_SLOTS], (j / TC_A_SLOTS) & 1);
            if (j < TC_GK) mbar_wait(&s.b_full[j % TC_B_SLOTS], (j / TC_B_SLOTS) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t da = da0 + (uint64_t)((j % TC_A_SLOTS) * (2 * TC_A_BYTES / 16));
#pragma unroll
                for (int t = 0; t < TC_ROWS; ++t) {
                    const int r = j - t;
                    if (r < 0 || r >= TC_GK) continue;
                    const uint64_t db = db0 + (uint64_t)((r % TC_B_SLOTS) * (2 * TC_B_BYTES / 16));
                    const uint32_t d = tmem + (uint32_t)(t * TC_N);
                    uint32_t st = t == 0 ? started0 : started1;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        if (pass >= npass) break;
                        const uint64_t pa = da + (pass == 2 ? TC_A_BYTES / 16 : 0), pb = db + (pass == 1 ? TC_B_BYTES / 16 : 0);
#pragma unroll
                        for (int ks = 0; ks < TC_KROW / 8; ++ks) {
                            umma_tf32(d, pa + ks * (2 * (TC_M / 8) * 128 / 16), pb + ks * (2 * (TC_N / 8) * 128 / 16), idesc, st);
                            st = 1;
                        }
                    }
                }
                umma_commit(&s.step_done[j % 4]);
            }
            __syncwarp();
            if (j - 0 >= 0 && j < TC_GK) started0 = 1;
            if (j - 1 >= 0) started1 = 1;
            if (j >= 1) mbar_wait(&s.step_done[(j - 1) % 4], ((j - 1) / 4) & 1);     // B_{j-2}'s slot is free now
            if (j + 2 < TC_GK && elect_one()) load_b(j + 2);
            __syncwarp();
        }
        if (elect_one()) umma_commit(&s.all_done);    // its own single-phase barrier: a parity wait cannot alias an earlier step
        __syncwarp();
    } else if (warp <= 4) {
        // ================= A builders: thread m owns pixel column m of the tile =================
        const int m = tid - 32;
        const int mg = m >> 3, mr = m & 7;
        for (int j = 0; j < TC_INROWS; ++j) {
            if (j >= TC_A_SLOTS) mbar_wait(&s.step_done[(j - TC_A_SLOTS) % 4], ((j - TC_A_SLOTS) / 4) & 1);
            float v[TC_KROW];
#pragma unroll
            for (int c = 0; c < TC_GK; ++c) v[c] = s.in[j][m + c];
#pragma unroll
            for (int c = TC_GK; c < TC_KROW; ++c) v[c] = 0.0f;
            unsigned char* ahi = &s.a[j % TC_A_SLOTS][0][0];
            unsigned char* alo = &s.a[j % TC_A_SLOTS][1][0];
#pragma unroll
            for (int kg = 0; kg < TC_KROW / 4; ++kg) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    h[e] = tf32_rna(v[4 * kg + e]);
                    l[e] = tf32_rna(v[4 * kg + e] - __uint_as_float(h[e]));
                }
                const int off = ((kg * (TC_M / 8) + mg) * 8 + mr) * 16;
                *reinterpret_cast<uint4*>(ahi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(alo + off) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
            mbar_arrive(&s.a_full[j % TC_A_SLOTS]);
        }
    }

    // ================= epilogue: all 8 warps; warp w reads TMEM lanes 32*(w%4).., accumulator w/4 =================
    mbar_wait(&s.all_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // theta_i table (one IEEE division each, GaborFilter.py:44) in the input tile's storage, which is free now
    float* theta = &s.in[0][0];
    __syncthreads();
    if (tid < TC_N) theta[tid] = theta_tc(tid, nf);
    __syncthreads();
    const int t = warp >> 2, q = warp & 3;
    const int mpx = q * 32 + lane;
    const int x = x0 + mpx, y = y0 + t;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * TC_N);
    float mval = 0.0f;
    int am = 0;
    for (int c0 = 0; c0 < TC_N; c0 += 32) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float rv = fabsf(__uint_as_float(r[k]));
            if (c0 + k == 0) { mval = rv; am = 0; }
            else if (c0 + k < nf && rv > mval) { mval = rv; am = c0 + k; }          // first maximum, like torch.max
        }
    }
    const float best = theta_tc(am, nf);
    const float PI = 3.14159265358979323846f;
    MhCascade<1> acc;
    acc.init(nf);
    for (int c0 = 0; c0 < TC_N; c0 += 32) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const int i = c0 + k;
            if (i < nf) {
                const float th = theta[i];
                const float d = fminf(fabsf(best - th), fminf(fabsf(best - th - PI), fabsf(best - th + PI)));
                const float rd = fabsf(__uint_as_float(r[k])) - mval;
                acc.begin_row(i);
                acc.add(0, d * rd * rd);
            }
        }
    }
    float ssum;
    acc.finish(nf, &ssum);
    float v = sqrtf(ssum);
    const bool pos = v > 0.0f;
    v = pos ? v : 0.0f;
    if (x < W && y < H) {
        orient[(size_t)y * W + x] = pos ? best : 0.0f;
        var[(size_t)y * W + x] = v;
    } else {
        v = 0.0f;
    }
    float bm = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    if (lane == 0) s.red[warp] = bm;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        bm = (lane < 8) ? s.red[lane] : 0.0f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        if (lane == 0) atomicMax(gmax, __float_as_uint(bm));
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TC_TMEM_COLS) : "memory");
    }
}

}  // namespace

// bank_tc: the bank split into tf32 hi / lo parts in the shared-memory operand layout, built by the host
// (monohair_b200/gabor.py: bank_for_tensor_cores): [17 kernel rows][hi, lo][6 k-groups][24 filter groups][8][4] float32.
extern "C" int64_t mh_gabor_tc_bank_bytes(void) { return (int64_t)TC_GK * 2 * TC_B_BYTES; }

void gabor_finish_launch(cudaStream_t st, int64_t HW, const float* orient, const float* var, const unsigned int* gmax, float lo,
                         float hi, float* conf, float* two);

extern "C" int mh_gabor_orientation_tc(void* stream, const float* image, int32_t H, int32_t W, const void* bank_tc, int32_t nf,
                                       float clamp_low, float clamp_high, float* orient, float* conf, float* two_channel,
                                       void* workspace, int64_t workspace_bytes) {
    MH_CHECK_ARG(image && bank_tc && orient && conf && workspace, "null pointer");
    MH_CHECK_ARG(H > 0 && W > 0 && nf > 0 && nf <= TC_N, "bad sizes (at most 192 filters)");
    MH_CHECK_ARG(workspace_bytes >= (int64_t)4 * H * W + 256, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)H * W;
    float* var = reinterpret_cast<float*>(workspace);
    unsigned int* gmax = reinterpret_cast<unsigned int*>(var + HW);
    cudaMemsetAsync(gmax, 0, 4, st);
    static thread_local bool attr[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 15]) {
        cudaError_t e = cudaFuncSetAttribute(gabor_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem));
        if (e != cudaSuccess) { mh_set_error("mh_gabor_orientation_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 2; }
        attr[dev & 15] = true;
    }
    dim3 grid((W + TC_M - 1) / TC_M, (H + TC_ROWS - 1) / TC_ROWS);
    static const int npass = getenv("MH_GABOR_TC_PASSES") ? atoi(getenv("MH_GABOR_TC_PASSES")) : 3;   // < 3: tuning experiments only
    gabor_tc_kernel<<<grid, TC_THREADS, sizeof(TcSmem), st>>>(image, H, W, reinterpret_cast<const unsigned char*>(bank_tc), nf,
                                                              orient, var, gmax, npass);
    MH_COUNT_LAUNCH();
    gabor_finish_launch(st, HW, orient, var, gmax, clamp_low, clamp_high, conf, two_channel);
    MH_CHECK_LAUNCH();
    return 0;
}
