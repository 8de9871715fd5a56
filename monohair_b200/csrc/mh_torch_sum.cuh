// Order-exact emulation of at::sum / at::mean over a CONTIGUOUS inner dimension of K floats on CPU
// (ATen SumKernel.cpp: cascade_sum -> vectorized_inner_sum / scalar_inner_sum -> row_sum -> multi_row_sum; the
// kernel that runs in the torch 2.11 wheel uses 8 float lanes).
//
// compute_points_similarity takes torch.mean(similar, dim=-1) and then an argmax (PMVO_utils.py:379-380); two
// candidate orientations whose mean similarities differ by an ulp are decided by the summation order, so the
// medoid kernels follow it.  row_sum over n items (scalars when K < 8, else 8-lane vectors v_i = x[8i..8i+8)):
//     p[k] = sum_{r < n/4} item[4r+k] (k = 0..3);  p[0] += item[i] for the left-over i >= 4*(n/4);
//     p[0] += p[1]; p[0] += p[2]; p[0] += p[3]
// K < 8 : total = row_sum over the K scalars.
// K >= 8: total = (scalar tail x[8*nv..K) summed from 0), then += lane l of row_sum over the nv vectors, l = 0..7.
// multi_row_sum's 16-row cascade engages from n/4 >= 16, i.e. K >= 512; this emulation is exact below that and
// keeps the same (non-cascaded) order above it.  Probe that established the order: DESIGN.md §4, checked for
// K = 1..257 against torch.sum bit-for-bit.
#pragma once
#include "mh_common.cuh"

template <typename Get>
__device__ __forceinline__ float mh_torch_inner_sum(int K, Get get) {
    constexpr int VL = 8;
    if (K < VL) {
        const int rows = K / 4;
        float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
        for (int r = 0; r < rows; ++r) { p0 += get(4 * r); p1 += get(4 * r + 1); p2 += get(4 * r + 2); p3 += get(4 * r + 3); }
        for (int i = 4 * rows; i < K; ++i) p0 += get(i);
        p0 += p1; p0 += p2; p0 += p3;
        return p0;
    }
    const int nv = K / VL;
    float total = 0.0f;
    for (int k = nv * VL; k < K; ++k) total += get(k);
    const int rows = nv / 4;
    float p0[VL];
#pragma unroll
    for (int l = 0; l < VL; ++l) p0[l] = 0.0f;
    for (int r = 0; r < rows; ++r)
#pragma unroll
        for (int l = 0; l < VL; ++l) p0[l] += get((4 * r) * VL + l);
    for (int i = 4 * rows; i < nv; ++i)
#pragma unroll
        for (int l = 0; l < VL; ++l) p0[l] += get(i * VL + l);
    for (int k = 1; k < 4; ++k) {
        float pk[VL];
#pragma unroll
        for (int l = 0; l < VL; ++l) pk[l] = 0.0f;
        for (int r = 0; r < rows; ++r)
#pragma unroll
            for (int l = 0; l < VL; ++l) pk[l] += get((4 * r + k) * VL + l);
#pragma unroll
        for (int l = 0; l < VL; ++l) p0[l] += pk[l];
    }
#pragma unroll
    for (int l = 0; l < VL; ++l) total += p0[l];
    return total;
}

// The same sum spread over 8 lanes of a warp (lane l = vl of its group of 8 takes vector lane l of every 8-float vector;
// the group's lanes then add their partial sums in lane order onto the scalar tail, exactly as above).  Every lane of the
// group returns the total.  For the last few rows of a K x K medoid matrix: with K = 100, rows 96..99 would otherwise
// occupy 4 lanes of a warp for a full 100-term row each; 4 groups of 8 lanes finish them in an eighth of the time.
// Requires K >= 8 and K < 512 (as above); all 32 lanes of the warp must call it (shuffles).
template <typename Get>
__device__ __forceinline__ float mh_torch_inner_sum_split8(int K, int lane, Get get) {
    constexpr int VL = 8;
    const int vl = lane & 7, gbase = lane & ~7;
    const int nv = K / VL, rows = nv / 4;
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
    for (int r = 0; r < rows; ++r) p0 += get((4 * r) * VL + vl);
    for (int i = 4 * rows; i < nv; ++i) p0 += get(i * VL + vl);
    for (int r = 0; r < rows; ++r) p1 += get((4 * r + 1) * VL + vl);
    for (int r = 0; r < rows; ++r) p2 += get((4 * r + 2) * VL + vl);
    for (int r = 0; r < rows; ++r) p3 += get((4 * r + 3) * VL + vl);
    p0 += p1; p0 += p2; p0 += p3;
    float total = 0.0f;
    for (int k = nv * VL; k < K; ++k) total += get(k);
#pragma unroll
    for (int l = 0; l < VL; ++l) total += __shfl_sync(0xffffffffu, p0, gbase + l);
    return total;
}
