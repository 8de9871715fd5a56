// Packing of the per-view maps into the two resident planes the PMVO kernels gather from.
//   mapC float4 {depth, mask', max_PxP conf, ori_row}   everything filter_points needs of a (point, view) pair in ONE
//                16 B texel = one 32 B sector (the PxP maximum used to sit in mapP: two sectors per pair)
//   mapP float4 {unit_row, unit_col, conf, ori_col}      one 16 B texel per patch entry; the direction is stored
//                already normalised the way torch.cosine_similarity normalises it (x / max(||x||, 1e-8)), so the
//                patch scans do not repeat a sqrt and two divisions per entry.  The RAW orientation (ori_row in mapC.w,
//                ori_col in mapP.w; the 2 px step of sample_next_3d_pos uses it un-normalised, PMVO.py:300) is only read at
//                the centre pixel, by kernels that fetch both texels there anyway
// The PxP maximum with edge clamping (get_c_patch + torch.max, PMVO.py:415-418 / :162) equals a max filter
// over the window intersected with the image, so it is computed once per view here (separable: rows then cols
// through a shared-memory tile) instead of P*P gathers per (point, view).
#include "mh_common.cuh"

namespace {

constexpr int TILE_X = 32, TILE_Y = 16, MAX_HALF = 8;

template <typename ConfLoad, typename Emit>
__global__ void __launch_bounds__(TILE_X * TILE_Y)
pack_kernel(int H, int W, int half, ConfLoad conf_at, Emit emit) {
    __shared__ float tile[TILE_Y + 2 * MAX_HALF][TILE_X + 2 * MAX_HALF + 1];
    __shared__ float rowmax[TILE_Y + 2 * MAX_HALF][TILE_X + 1];
    const int x0 = blockIdx.x * TILE_X, y0 = blockIdx.y * TILE_Y;
    const int tw = TILE_X + 2 * half, th = TILE_Y + 2 * half;
    for (int i = threadIdx.y * TILE_X + threadIdx.x; i < tw * th; i += TILE_X * TILE_Y) {
        int ty = i / tw, tx = i - ty * tw;
        int gy = min(max(y0 + ty - half, 0), H - 1), gx = min(max(x0 + tx - half, 0), W - 1);
        tile[ty][tx] = conf_at(gy, gx);
    }
    __syncthreads();
    for (int i = threadIdx.y * TILE_X + threadIdx.x; i < TILE_X * th; i += TILE_X * TILE_Y) {
        int ty = i / TILE_X, tx = i - ty * TILE_X;
        float m = tile[ty][tx];
        for (int d = 1; d <= 2 * half; ++d) m = fmaxf(m, tile[ty][tx + d]);
        rowmax[ty][tx] = m;
    }
    __syncthreads();
    int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < W && y < H) {
        float m = rowmax[threadIdx.y][threadIdx.x];
        for (int d = 1; d <= 2 * half; ++d) m = fmaxf(m, rowmax[threadIdx.y + d][threadIdx.x]);
        emit(y, x, m);
    }
}

}  // namespace

extern "C" int mh_views_pack(void* stream, int32_t v, int32_t H, int32_t W, int32_t P,
                             const float* depth, int32_t depth_stride, const float* ori, const float* conf,
                             const float* mask, int32_t mask_stride, void* mapC_, void* mapP_) {
    MH_CHECK_ARG(depth && ori && conf && mask && mapC_ && mapP_, "null pointer");
    MH_CHECK_ARG(H > 0 && W > 0 && v >= 0, "bad size");
    MH_CHECK_ARG(P >= 1 && (P & 1) && P / 2 <= MAX_HALF, "patch size must be odd and <= 17");
    float4* mapC = reinterpret_cast<float4*>(mapC_) + (size_t)v * H * W;
    float4* mapP = reinterpret_cast<float4*>(mapP_) + (size_t)v * H * W;
    dim3 grid((W + TILE_X - 1) / TILE_X, (H + TILE_Y - 1) / TILE_Y), block(TILE_X, TILE_Y);
    auto conf_at = [=] __device__(int y, int x) { return __ldg(conf + (size_t)y * W + x); };
    auto emit = [=] __device__(int y, int x, float cmax) {
        size_t i = (size_t)y * W + x;
        float m = __ldg(mask + i * mask_stride);
        m = (m > 0.2f) ? 1.0f : m;                                   // PMVO.py:427 / :124
        const float2 o = __ldg(reinterpret_cast<const float2*>(ori) + i);
        mapC[i] = make_float4(__ldg(depth + i * depth_stride), m, cmax, o.x);
        float n0, n1;
        mh_normalize2(o.x, o.y, n0, n1);
        mapP[i] = make_float4(n0, n1, __ldg(conf + i), o.y);
    };
    pack_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(H, W, P / 2, conf_at, emit);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

// Same as mh_views_pack for the dtypes the reference's loaders actually hand to PMVO.__init__ (PMVO_utils.py:255-313):
// depth float32, Ori / Conf / mask float64.  The float64 -> float32 rounding of PMVO.py:23-26 (.type(torch.float))
// happens here, so the host side only issues the raw H2D copies.
extern "C" int mh_views_pack_f64(void* stream, int32_t v, int32_t H, int32_t W, int32_t P,
                                 const float* depth, int32_t depth_stride, const double* ori, const double* conf,
                                 const double* mask, int32_t mask_stride, void* mapC_, void* mapP_) {
    MH_CHECK_ARG(depth && ori && conf && mask && mapC_ && mapP_, "null pointer");
    MH_CHECK_ARG(H > 0 && W > 0 && v >= 0, "bad size");
    MH_CHECK_ARG(P >= 1 && (P & 1) && P / 2 <= MAX_HALF, "patch size must be odd and <= 17");
    float4* mapC = reinterpret_cast<float4*>(mapC_) + (size_t)v * H * W;
    float4* mapP = reinterpret_cast<float4*>(mapP_) + (size_t)v * H * W;
    dim3 grid((W + TILE_X - 1) / TILE_X, (H + TILE_Y - 1) / TILE_Y), block(TILE_X, TILE_Y);
    auto conf_at = [=] __device__(int y, int x) { return (float)__ldg(conf + (size_t)y * W + x); };
    auto emit = [=] __device__(int y, int x, float cmax) {
        size_t i = (size_t)y * W + x;
        float m = (float)__ldg(mask + i * mask_stride);
        m = (m > 0.2f) ? 1.0f : m;
        const double2 od = __ldg(reinterpret_cast<const double2*>(ori) + i);
        const float ox = (float)od.x, oy = (float)od.y;
        mapC[i] = make_float4(__ldg(depth + i * depth_stride), m, cmax, ox);
        float n0, n1;
        mh_normalize2(ox, oy, n0, n1);
        mapP[i] = make_float4(n0, n1, (float)__ldg(conf + i), oy);
    };
    pack_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(H, W, P / 2, conf_at, emit);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_views_pack_u8(void* stream, int32_t v, int32_t H, int32_t W, int32_t P,
                                const float* depth, int32_t depth_stride, const uint8_t* ori_gray,
                                const uint8_t* conf_u8, const uint8_t* mask_u8, const float* ori_lut,
                                const float* conf_lut, const float* mask_lut, void* mapC_, void* mapP_) {
    MH_CHECK_ARG(depth && ori_gray && conf_u8 && mask_u8 && ori_lut && conf_lut && mask_lut && mapC_ && mapP_, "null pointer");
    MH_CHECK_ARG(H > 0 && W > 0 && v >= 0, "bad size");
    MH_CHECK_ARG(P >= 1 && (P & 1) && P / 2 <= MAX_HALF, "patch size must be odd and <= 17");
    float4* mapC = reinterpret_cast<float4*>(mapC_) + (size_t)v * H * W;
    float4* mapP = reinterpret_cast<float4*>(mapP_) + (size_t)v * H * W;
    dim3 grid((W + TILE_X - 1) / TILE_X, (H + TILE_Y - 1) / TILE_Y), block(TILE_X, TILE_Y);
    // conf_lut is monotone (k/255) so the max of decoded values is the decode of the max.
    auto conf_at = [=] __device__(int y, int x) { return __ldg(conf_lut + __ldg(conf_u8 + (size_t)y * W + x)); };
    auto emit = [=] __device__(int y, int x, float cmax) {
        size_t i = (size_t)y * W + x;
        float m = __ldg(mask_lut + __ldg(mask_u8 + i));               // lut already applies <50 -> 0 and /255
        m = (m > 0.2f) ? 1.0f : m;
        const int g = __ldg(ori_gray + i);
        const float ox = __ldg(ori_lut + 2 * g), oy = __ldg(ori_lut + 2 * g + 1);
        mapC[i] = make_float4(__ldg(depth + i * depth_stride), m, cmax, ox);
        float n0, n1;
        mh_normalize2(ox, oy, n0, n1);
        mapP[i] = make_float4(n0, n1, __ldg(conf_lut + __ldg(conf_u8 + i)), oy);
    };
    pack_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(H, W, P / 2, conf_at, emit);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
