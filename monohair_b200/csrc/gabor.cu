// Gabor orientation filter bank.
//   mh_gabor_orientation    calOrientationGabor.filter + forward, iter=1 (GaborFilter.py:29-113), float32,
//                           zero-padded cross-correlation with the 180 x 17x17 bank (F.conv2d semantics).
//   mh_filterbank_wrap_f64  calc_orients' 180 periodic true convolutions (calc_orientation_maps.py:27-32), float64.
//   mh_dog_f64              difference_of_gaussians = two separable scipy.ndimage.gaussian_filter (mode 'nearest').
//
// The bank contraction is [HW x 289] x [289 x 180]: 104 kFLOP per pixel against 16 B of I/O, so it is FP32-FMA
// bound; it runs on the CUDA cores with register tiling (4 pixels x 6 filters per thread, operands through shared
// memory with 128-bit loads).  Tensor cores are deliberately not used in this round: the argmax over 180
// near-equal responses needs ~fp32 accuracy (3xTF32 / bf16x3 splits), see DESIGN.md.
// |responses| are staged once in a [n][H][W] workspace; the per-pixel epilogue (argmax, circular-distance
// weighted variance, global max) follows torch's reduction order over the 180 channels.
#include "mh_common.cuh"

namespace {

constexpr int GK = 17, GPAD = 8, GKP = 20;               // kernel size, padding, padded row length (5 x float4)
constexpr int TX = 32, TY = 8, PXT = 4;                  // threads, pixels per thread along x
constexpr int TILE_W = TX * PXT, TILE_H = TY;            // 128 x 8 output tile
constexpr int IN_W = TILE_W + 2 * GPAD + 4, IN_H = TILE_H + 2 * GPAD;   // 148 (144 used + 4 over-read pad) x 24
constexpr int FPT = 6;                                   // filters per pass

__global__ void __launch_bounds__(TX * TY)
gabor_resp_kernel(const float* __restrict__ img, int H, int W, const float* __restrict__ bank, int nf,
                  float* __restrict__ resp) {
    __shared__ __align__(16) float tin[IN_H][IN_W];
    __shared__ __align__(16) float wk[FPT][GK][GKP];
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * TX + threadIdx.x;
    for (int i = tid; i < IN_H * IN_W; i += TX * TY) {
        const int ty = i / IN_W, tx = i - ty * IN_W;
        const int gy = y0 + ty - GPAD, gx = x0 + tx - GPAD;
        tin[ty][tx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(img + (size_t)gy * W + gx) : 0.0f;
    }
    const int px = threadIdx.x * PXT, py = threadIdx.y;
    for (int f0 = 0; f0 < nf; f0 += FPT) {
        __syncthreads();
        for (int i = tid; i < FPT * GK * GKP; i += TX * TY) {
            const int f = i / (GK * GKP), r = (i / GKP) % GK, c = i % GKP;
            wk[f][r][c] = (c < GK && f0 + f < nf) ? __ldg(bank + ((size_t)(f0 + f) * GK + r) * GK + c) : 0.0f;
        }
        __syncthreads();
        float acc[FPT][PXT];
#pragma unroll
        for (int f = 0; f < FPT; ++f)
#pragma unroll
            for (int p = 0; p < PXT; ++p) acc[f][p] = 0.0f;
        for (int r = 0; r < GK; ++r) {
            float in[GKP + PXT];                                  // 24 inputs cover 4 pixels x 20 (padded) taps
#pragma unroll
            for (int q = 0; q < (GKP + PXT) / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(&tin[py + r][px + 4 * q]);
                in[4 * q] = t.x; in[4 * q + 1] = t.y; in[4 * q + 2] = t.z; in[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int f = 0; f < FPT; ++f) {
#pragma unroll
                for (int q = 0; q < GKP / 4; ++q) {
                    const float4 w4 = *reinterpret_cast<const float4*>(&wk[f][r][4 * q]);
                    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (4 * q + k < GK) {
#pragma unroll
                            for (int p = 0; p < PXT; ++p) acc[f][p] = fmaf(w[k], in[4 * q + k + p], acc[f][p]);
                        }
                    }
                }
            }
        }
        const int y = y0 + py;
        if (y < H) {
#pragma unroll
            for (int f = 0; f < FPT; ++f) {
                if (f0 + f >= nf) continue;
                float* o = resp + ((size_t)(f0 + f) * H + y) * W + x0 + px;
                if (x0 + px + 3 < W && ((W & 3) == 0)) {
                    *reinterpret_cast<float4*>(o) = make_float4(fabsf(acc[f][0]), fabsf(acc[f][1]), fabsf(acc[f][2]), fabsf(acc[f][3]));
                } else {
#pragma unroll
                    for (int p = 0; p < PXT; ++p) if (x0 + px + p < W) o[p] = fabsf(acc[f][p]);
                }
            }
        }
    }
}

// theta_i = ((pi_f32 * i) / n) as the reference builds it in float32 (GaborFilter.py:44, :51)
__device__ __forceinline__ float theta_of(int i, int n) { return (3.14159265358979323846f * (float)i) / (float)n; }

__global__ void gabor_epilogue_kernel(const float* __restrict__ resp, int64_t HW, int nf, float* __restrict__ orient,
                                      float* __restrict__ var, unsigned int* __restrict__ gmax) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.0f;
    if (i < HW) {
        float m = resp[i];
        int am = 0;
        for (int k = 1; k < nf; ++k) { const float r = resp[(size_t)k * HW + i]; if (r > m) { m = r; am = k; } }
        if (m != m) { /* NaN response: torch.max propagates; leave as is */ }
        const float best = theta_of(am, nf);
        const float PI = 3.14159265358979323846f;
        MhCascade<1> acc;
        acc.init(nf);
        for (int k = 0; k < nf; ++k) {
            const float th = theta_of(k, nf);
            const float d = fminf(fabsf(best - th), fminf(fabsf(best - th - PI), fabsf(best - th + PI)));
            const float rd = resp[(size_t)k * HW + i] - m;
            acc.begin_row(k);
            acc.add(0, d * rd * rd);
        }
        float s;
        acc.finish(nf, &s);
        v = sqrtf(s);
        const bool pos = v > 0.0f;
        orient[i] = pos ? best : 0.0f;
        v = pos ? v : 0.0f;
        var[i] = v;
    }
    // block max -> global max (values are >= 0 so the uint ordering equals the float ordering)
    __shared__ float red[32];
    float bm = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = bm;
    __syncthreads();
    if (threadIdx.x < 32) {
        bm = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        if (threadIdx.x == 0) atomicMax(gmax, __float_as_uint(bm));
    }
}

__global__ void gabor_finish_kernel(int64_t HW, const float* __restrict__ orient, const float* __restrict__ var,
                                    const unsigned int* __restrict__ gmax, float lo, float hi,
                                    float* __restrict__ conf, float* __restrict__ two) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const float mx = __uint_as_float(*gmax);
    const float v = var[i] / mx;
    float c = (v - lo) / (hi - lo);
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    if (c != c) c = __int_as_float(0x7fc00000);
    conf[i] = c;
    if (two) { two[i] = sinf(orient[i]); two[HW + i] = cosf(orient[i]); }
}

// ---------------------------------------------------------------------------------------------- float64 paths
__global__ void __launch_bounds__(256)
wrap_bank_kernel(const double* __restrict__ img, int H, int W, const double* __restrict__ bank, int nf, int ks,
                 double* __restrict__ out) {
    extern __shared__ double tile[];                         // [(16+ks-1)][(16+ks-1)]
    const int half = ks / 2, tw = 16 + ks - 1;
    const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
    const int tid = threadIdx.y * 16 + threadIdx.x;
    for (int i = tid; i < tw * tw; i += 256) {
        const int ty = i / tw, tx = i - ty * tw;
        int gy = (y0 + ty - half) % H, gx = (x0 + tx - half) % W;
        if (gy < 0) gy += H;
        if (gx < 0) gx += W;
        tile[i] = img[(size_t)gy * W + gx];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    for (int f = 0; f < nf; ++f) {
        const double* __restrict__ k = bank + (size_t)f * ks * ks;
        // ndi.convolve = correlate with the flipped kernel: accumulate in raster order of the flipped kernel
        double acc = 0.0;
        for (int qi = 0; qi < ks; ++qi)
            for (int qj = 0; qj < ks; ++qj) {
                const double w = __ldg(k + (ks - 1 - qi) * ks + (ks - 1 - qj));
                if (w != 0.0) acc += w * tile[(threadIdx.y + qi) * tw + threadIdx.x + qj];
            }
        out[((size_t)f * H + y) * W + x] = fabs(acc);
    }
}

// scipy correlate1d with symmetric weights, mode 'nearest':  tmp = in[c]*w[r]; for l=-r..-1: tmp += (in[c+l]+in[c-l])*w[l+r]
__global__ void gauss1d_kernel(const double* __restrict__ in, int H, int W, const double* __restrict__ w, int r, int axis,
                               double* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const int n = axis == 0 ? H : W, c = axis == 0 ? y : x;
    auto at = [&](int k) { k = min(max(k, 0), n - 1); return axis == 0 ? in[(size_t)k * W + x] : in[(size_t)y * W + k]; };
    double tmp = at(c) * w[r];
    for (int l = -r; l < 0; ++l) tmp += (at(c + l) + at(c - l)) * w[l + r];
    out[(size_t)y * W + x] = tmp;
}
__global__ void sub_kernel(const double* a, const double* b, int64_t n, double* o) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] - b[i];
}

}  // namespace

// shared with gabor_tc.cu
void gabor_finish_launch(cudaStream_t st, int64_t HW, const float* orient, const float* var, const unsigned int* gmax, float lo,
                         float hi, float* conf, float* two) {
    gabor_finish_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, st>>>(HW, orient, var, gmax, lo, hi, conf, two);
    MH_COUNT_LAUNCH();
}

// workspace: [resp nf*H*W floats][var H*W floats][gmax 64 B]
extern "C" int64_t mh_gabor_workspace_bytes(int32_t H, int32_t W, int32_t nf) {
    return (int64_t)4 * ((int64_t)nf * H * W + (int64_t)H * W) + 256;
}

extern "C" int mh_gabor_orientation(void* stream, const float* image, int32_t H, int32_t W, const float* bank,
                                    int32_t nf, int32_t ksize, float clamp_low, float clamp_high, float* orient,
                                    float* conf, float* two_channel, void* workspace, int64_t workspace_bytes) {
    MH_CHECK_ARG(image && bank && orient && conf && workspace, "null pointer");
    MH_CHECK_ARG(H > 0 && W > 0 && nf > 0, "bad sizes");
    MH_CHECK_ARG(ksize == GK, "kernel size must be 17 (GaborFilter.py:106)");
    MH_CHECK_ARG(workspace_bytes >= mh_gabor_workspace_bytes(H, W, nf), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)H * W;
    float* resp = reinterpret_cast<float*>(workspace);
    float* var = resp + (size_t)nf * HW;
    unsigned int* gmax = reinterpret_cast<unsigned int*>(var + HW);
    cudaMemsetAsync(gmax, 0, 4, st);
    dim3 grid((W + TILE_W - 1) / TILE_W, (H + TILE_H - 1) / TILE_H), block(TX, TY);
    gabor_resp_kernel<<<grid, block, 0, st>>>(image, H, W, bank, nf, resp);
    MH_COUNT_LAUNCH();
    gabor_epilogue_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, st>>>(resp, HW, nf, orient, var, gmax);
    MH_COUNT_LAUNCH();
    gabor_finish_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, st>>>(HW, orient, var, gmax, clamp_low, clamp_high, conf, two_channel);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_filterbank_wrap_f64(void* stream, const double* image, int32_t H, int32_t W, const double* bank,
                                      int32_t nf, int32_t ksize, double* out_abs) {
    MH_CHECK_ARG(image && bank && out_abs && H > 0 && W > 0 && nf > 0, "bad arguments");
    MH_CHECK_ARG((ksize & 1) && ksize >= 1 && ksize <= 33, "kernel size must be odd and <= 33");
    const int tw = 16 + ksize - 1;
    dim3 grid((W + 15) / 16, (H + 15) / 16), block(16, 16);
    wrap_bank_kernel<<<grid, block, sizeof(double) * tw * tw, (cudaStream_t)stream>>>(image, H, W, bank, nf, ksize, out_abs);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_dog_f64(void* stream, const double* image, int32_t H, int32_t W, const double* k_lo, int32_t r_lo,
                          const double* k_hi, int32_t r_hi, double* out, double* scratch) {
    MH_CHECK_ARG(image && k_lo && k_hi && out && scratch && H > 0 && W > 0 && r_lo >= 0 && r_hi >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)H * W;
    double* t0 = scratch;
    double* t1 = scratch + HW;
    dim3 grid((W + 127) / 128, H), block(128);
    // gaussian_filter filters axis 0 first, then axis 1
    gauss1d_kernel<<<grid, block, 0, st>>>(image, H, W, k_lo, r_lo, 0, t0);
    MH_COUNT_LAUNCH();
    gauss1d_kernel<<<grid, block, 0, st>>>(t0, H, W, k_lo, r_lo, 1, out);          // out = low
    gauss1d_kernel<<<grid, block, 0, st>>>(image, H, W, k_hi, r_hi, 0, t0);
    MH_COUNT_LAUNCH();
    gauss1d_kernel<<<grid, block, 0, st>>>(t0, H, W, k_hi, r_hi, 1, t1);           // t1 = high
    sub_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, st>>>(out, t1, HW, out);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
