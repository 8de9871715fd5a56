// Candidate-point sampling around the coarse mesh (SamplePointsAroundmesh, Utils/PMVO_utils.py:316-339) and the
// depth-map rasteriser (render_bust_hair_depth, Utils/Render_utils.py:310-347), device side.
#include <cstring>
#include "mh_common.cuh"

namespace {

// occ[x][y][z] = 1 for the cell of every surface sample: colmap_points[:,1:] *= -1; round((p - bbox_min) / vsize) in
// float64 (np.round: half to even), clipped to the grid (PMVO_utils.py:318-324)
__global__ void mark_cells_kernel(const double* __restrict__ pts, int64_t n, double mx, double my, double mz, double vs,
                                  int gx, int gy, int gz, uint8_t* __restrict__ occ) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = rint((pts[3 * i] - mx) / vs), y = rint((-pts[3 * i + 1] - my) / vs), z = rint((-pts[3 * i + 2] - mz) / vs);
    const int ix = (int)fmin(fmax(x, 0.0), (double)(gx - 1)), iy = (int)fmin(fmax(y, 0.0), (double)(gy - 1)),
              iz = (int)fmin(fmax(z, 0.0), (double)(gz - 1));
    occ[((size_t)ix * gy + iy) * gz + iz] = 1;
}

// sample[r * m + c] = (cell[c] + rand[r * m + c]) * vsize + bbox_min, then y, z negated (PMVO_utils.py:326-337): the
// occupied cells in np.nonzero order, tiled num_per_grid times; rand = the np.random.random draws of the reference,
// injected by the host so that its seeded stream is consumed unchanged.
__global__ void sample_cells_kernel(const int64_t* __restrict__ cells /*[m][3]*/, int64_t m, int64_t total, const double* __restrict__ rnd,
                                    double mx, double my, double mz, double vs, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t c = i % m;
    const double x = ((double)cells[3 * c] + rnd[3 * i] * 1) * vs + mx;
    const double y = ((double)cells[3 * c + 1] + rnd[3 * i + 1] * 1) * vs + my;
    const double z = ((double)cells[3 * c + 2] + rnd[3 * i + 2] * 1) * vs + mz;
    out[3 * i] = x;
    out[3 * i + 1] = y * -1;
    out[3 * i + 2] = z * -1;
}

}  // namespace

extern "C" int mh_sample_mark_cells(void* stream, const double* points, int64_t n, const double* bbox_min_host, double vsize,
                                    int32_t gx, int32_t gy, int32_t gz, uint8_t* occ) {
    MH_CHECK_ARG(occ && bbox_min_host && (n == 0 || points) && n >= 0 && gx > 0 && gy > 0 && gz > 0 && vsize > 0, "bad arguments");
    cudaMemsetAsync(occ, 0, (size_t)gx * gy * gz, (cudaStream_t)stream);
    if (n > 0) {
        mark_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, n, bbox_min_host[0], bbox_min_host[1],
                                                                                         bbox_min_host[2], vsize, gx, gy, gz, occ);
        MH_COUNT_LAUNCH();
    }
    MH_CHECK_LAUNCH();
    return 0;
}

extern "C" int mh_sample_cells(void* stream, const int64_t* cells, int64_t m, int32_t num_per_grid, const double* rnd,
                               const double* bbox_min_host, double vsize, double* out) {
    MH_CHECK_ARG(out && bbox_min_host && (m == 0 || (cells && rnd)) && m >= 0 && num_per_grid >= 0, "bad arguments");
    const int64_t total = m * num_per_grid;
    if (total == 0) return 0;
    sample_cells_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cells, m, total, rnd, bbox_min_host[0],
                                                                                           bbox_min_host[1], bbox_min_host[2], vsize, out);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}

// ---- depth-map rasteriser (render_bust_hair_depth, Utils/Render_utils.py:310-347; shaders :150-178) -------------------
// The reference draws the coarse mesh with OpenGL (moderngl / EGL): gl_Position = proj * pose * v, colour = -z_cam / 2
// (depth_range 2.0) with the varying interpolated perspective-correctly, depth test on, clear colour 1, frame flipped on
// read-back.  In pixel terms: a vertex lands at x = ((-u) + 1) / 2 * W, y = (v + 1) / 2 * H (u, v as in Camera.projection;
// row 0 on top -- the same mapping PMVO.project_points uses), pixel (row, col) is covered when its centre
// (col + .5, row + .5) is inside the triangle, and its value is the nearest covering fragment's 1 / sum(lambda_i / -z_i) / 2.
namespace {

__global__ void raster_kernel(const float* __restrict__ verts, const int* __restrict__ faces, int64_t nf, MhCam cm, int H, int W,
                              unsigned int* __restrict__ zbuf) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    float px[3], py[3], iz[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* v = verts + 3 * (int64_t)faces[3 * f + k];
        float cx, cy, cz;
        mh_world_to_cam(cm.p, v[0], v[1], v[2], cx, cy, cz);
        if (!(cz < -0.1f)) return;                                   // crosses the near plane (znear 0.1): not drawn here
        mh_cam_to_xy(cm.fx, cm.fy, cm.cx, cm.cy, (float)W, (float)H, cx, cy, cz, px[k], py[k]);
        iz[k] = 1.0f / -cz;
    }
    const float area = (px[1] - px[0]) * (py[2] - py[0]) - (px[2] - px[0]) * (py[1] - py[0]);
    if (area == 0.0f || area != area) return;
    const int x0 = max(0, (int)floorf(fminf(px[0], fminf(px[1], px[2])) - 0.5f));
    const int x1 = min(W - 1, (int)ceilf(fmaxf(px[0], fmaxf(px[1], px[2])) - 0.5f));
    const int y0 = max(0, (int)floorf(fminf(py[0], fminf(py[1], py[2])) - 0.5f));
    const int y1 = min(H - 1, (int)ceilf(fmaxf(py[0], fmaxf(py[1], py[2])) - 0.5f));
    const float inv_area = 1.0f / area;
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
            const float qx = (float)x + 0.5f, qy = (float)y + 0.5f;
            const float l0 = ((px[1] - qx) * (py[2] - qy) - (px[2] - qx) * (py[1] - qy)) * inv_area;
            const float l1 = ((px[2] - qx) * (py[0] - qy) - (px[0] - qx) * (py[2] - qy)) * inv_area;
            const float l2 = 1.0f - l0 - l1;
            if (l0 < 0.0f || l1 < 0.0f || l2 < 0.0f) continue;
            const float depth = 1.0f / (l0 * iz[0] + l1 * iz[1] + l2 * iz[2]);       // perspective-correct -z_cam
            atomicMin(zbuf + (size_t)y * W + x, __float_as_uint(depth));             // positive floats order like their bits
        }
}

__global__ void raster_resolve_kernel(const unsigned int* __restrict__ zbuf, int64_t n, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int z = zbuf[i];
    out[i] = (z == 0x7f7f7f7fu) ? 1.0f : __uint_as_float(z) / 2.0f;
}

}  // namespace

extern "C" int mh_render_depth(void* stream, const float* verts, int64_t n_verts, const int32_t* faces, int64_t n_faces,
                               const float* cam_record_host /*[MH_CAM_STRIDE]*/, int32_t H, int32_t W, float* depth /*[H][W]*/,
                               void* zbuf /*uint32 [H][W]*/, int32_t clear) {
    MH_CHECK_ARG(depth && zbuf && cam_record_host && H > 0 && W > 0 && n_faces >= 0 && (n_faces == 0 || (verts && faces)), "bad arguments");
    (void)n_verts;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)H * W;
    if (clear) cudaMemsetAsync(zbuf, 0x7f, sizeof(unsigned int) * n, st);   // 0x7f7f7f7f = 3.39e38: "nothing drawn"; clear = 0 adds another mesh
    MhCam cm;
    memcpy(&cm, cam_record_host, sizeof(cm));
    if (n_faces > 0) {
        raster_kernel<<<(unsigned)((n_faces + 127) / 128), 128, 0, st>>>(verts, faces, n_faces, cm, H, W, reinterpret_cast<unsigned int*>(zbuf));
        MH_COUNT_LAUNCH();
    }
    raster_resolve_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const unsigned int*>(zbuf), n, depth);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
