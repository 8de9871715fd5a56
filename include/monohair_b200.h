/*
 * monohair_b200 — C ABI of the B200-native MonoHair hot path (Gabor bank, PMVO, HairGrow trace).
 *
 * The reference (KeyuWu-CS/MonoHair) has no FFI: its boundary is Python.  This header is the
 * contract the thin Python host (monohair_b200/*.py, mirroring PMVO.py / HairGrow.py / GaborFilter.py) binds
 * with ctypes; each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a caller-allocated DEVICE pointer unless its name ends in _host;
 *   - the library never allocates or frees; scratch comes from a caller workspace sized by *_workspace_bytes;
 *   - sizes are int64_t, `stream` is a cudaStream_t passed as void*;
 *   - no host synchronisation inside (except functions documented as synchronous);
 *   - return 0 on success, non-zero on error; the message is mh_last_error() (thread-local).
 *   - float32 everywhere unless noted; arithmetic follows the reference's fp32 operation order.
 */
#ifndef MONOHAIR_B200_H
#define MONOHAIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MH_CAM_STRIDE 32   /* floats per view in mh_views.cam, see mh_views_pack_camera_host */
#define MH_TOPK 20         /* PMVO.py:341 */
#define MH_NUM_BASE 10     /* PMVO.py:50  range(0,20,2) */

/* Resident per-view maps (PMVO.__init__, PMVO.py:14-37).  All device pointers, owned by the caller. */
typedef struct mh_views {
    int32_t V, H, W, P;        /* views, image rows, cols, patch size (odd) */
    const void* mapC;          /* float4 [V][H][W] = {depth, mask', max_{PxP} conf, ori_row}; mask' = mask>0.2 ? 1 : mask
                                  (PMVO.py:427): all that filter_points reads of a (point, view) pair, in one 32 B sector */
    const void* mapP;          /* float4 [V][H][W] = {unit_row, unit_col, conf, ori_col}: direction pre-normalised as
                                  torch.cosine_similarity does (PMVO.py:171, 491-515); raw orientation = (mapC.w, mapP.w) */
    const float* cam;          /* [V][MH_CAM_STRIDE] */
} mh_views;

const char* mh_last_error(void);
int mh_version(void);
/* number of CUDA kernels this library has launched so far in the process (bench.py: gpu_launches) */
int64_t mh_launch_count(void);

/* ---- views ------------------------------------------------------------------------------------------- */
/* Fill cam[v] from host data: pose = world->camera 4x4 row-major (Camera.pose), ndc_prj = {fx,fy,cx,cy}
 * (Camera.get_projection_matrix, Camera_utils.py:19-36), rinv = torch.linalg.inv(pose[:3,:3]) row-major
 * (Camera_utils.py:104).  Host-side helper, writes 32 floats into cam_host. */
int mh_views_pack_camera_host(const float* pose_host, const float* ndc_prj_host, const float* rinv_host,
                              float* cam_host);

/* Pack one view's float32 maps (already on the device) into mapC/mapP planes of view v.
 * depth/mask: pixel stride in floats (3 for the reference's [H,W,3] arrays, channel 0 is read; PMVO.py:485,523).
 * Replaces the per-view tensors kept by PMVO.__init__ (PMVO.py:23-26) and precomputes the PxP confidence
 * maximum used by filter_points (PMVO.py:415-418) and compute_prj_loss (PMVO.py:162). */
int mh_views_pack(void* stream, int32_t v, int32_t H, int32_t W, int32_t P,
                  const float* depth, int32_t depth_stride, const float* ori /*[H][W][2]*/,
                  const float* conf /*[H][W]*/, const float* mask, int32_t mask_stride,
                  void* mapC /*[V][H][W] float4*/, void* mapP /*[V][H][W] float4*/);

/* Same for the dtypes the reference's loaders produce (depth float32; Ori, Conf, mask float64): the float cast of
 * PMVO.py:23-26 is fused into the pack. */
int mh_views_pack_f64(void* stream, int32_t v, int32_t H, int32_t W, int32_t P,
                      const float* depth, int32_t depth_stride, const double* ori /*[H][W][2]*/,
                      const double* conf /*[H][W]*/, const double* mask, int32_t mask_stride, void* mapC, void* mapP);

/* Same from the on-disk 8-bit formats (best_ori gray, conf, mask) with the decode of
 * Load_Ori_And_Conf / load_mask (PMVO_utils.py:255-313) fused; lut = 256x2 floats {sin o, cos o} built on host. */
int mh_views_pack_u8(void* stream, int32_t v, int32_t H, int32_t W, int32_t P,
                     const float* depth, int32_t depth_stride, const uint8_t* ori_gray, const uint8_t* conf_u8,
                     const uint8_t* mask_u8, const float* ori_lut /*[256][2]*/, const float* conf_lut /*[256]*/,
                     const float* mask_lut /*[256]*/, void* mapC, void* mapP);

/* ---- PMVO.filter_points (PMVO.py:402-459) ------------------------------------------------------------- */
/* counters [5][N] = {sum vis, sum vis*mask, sum vis*low_conf, sum vis1, sum vis1*mask} over the views of
 * `views` (a rank's shard when view-sharded; all-reduce(SUM) them across ranks). */
int mh_filter_count(void* stream, const mh_views* views, const float* points /*[N][3]*/, int64_t N,
                    float visible_threshold, float conf_threshold, float* counters /*[5][N]*/);
/* surface / filter flags from (all-reduced) counters (PMVO.py:442-454). */
int mh_filter_decide(void* stream, const float* counters, int64_t N, uint8_t* surface, uint8_t* filter);

/* PMVO.compute_unvisible_points (PMVO.py:461-480): count [N] of views with dz<=0.9 and in image;
 * unvisible = !(count > 2) is taken on the host side after the optional all-reduce. */
int mh_visible_count(void* stream, const mh_views* views, const float* points, int64_t N, float dz_threshold,
                     float* count /*[N]*/);

/* PMVO.filter_head_points per-view part (PMVO.py:110-136): counters [2][N] = {sum vis, sum vis*mask'} with
 * unvisible = dz >= visible_threshold (no out-of-image test, as in the reference). */
int mh_head_count(void* stream, const mh_views* views, const float* points, int64_t N, float visible_threshold,
                  float* counters /*[2][N]*/);

/* filter = !(sum_vis - sum_vis*mask' < sum_vis/2) && !head_top, head_top = scalp_dist < dist_threshold (0.04)
 * && point.z < z_threshold (scalp_max.z - 0.01), compared in float64 like the reference's numpy (PMVO.py:105-141). */
int mh_head_decide(void* stream, const float* counters /*[2][N]*/, const double* scalp_dist /*[N]*/,
                   const float* points, int64_t N, double dist_threshold, double z_threshold, uint8_t* filter);

/* PMVO.Compute_Visible_and_Ori centre values (PMVO.py:346-376): visible [V][N], ori [V][N][2], conf [V][N]
 * (clamped to [1e-6,1]), optional mask' [V][N] and rowcol int32 [V][N][2] (col stored as -col-1 when the
 * projection fell outside the image: the unvisible_index of project_points, PMVO.py:384-395). */
int mh_centre_gather(void* stream, const mh_views* views, const float* points, int64_t N, float* visible,
                     float* ori, float* conf, float* mask, int32_t* rowcol);

/* ---- PMVO.forward (PMVO.py:39-78) --------------------------------------------------------------------- */
int64_t mh_pmvo_optimize_workspace_bytes(const mh_views* views, int64_t N);
/* offsets: the S depth offsets of sample_next_3d_pos (PMVO.py:274-278), computed by the host with torch.arange.
 * Outputs: ori [N][3] unit direction, loss [N], high_conf [N] (0/1).
 * Optional debug outputs (may be NULL): base_idx int32 [MH_TOPK][N], base_val [MH_TOPK][N] (topk order of
 * torch.topk on CPU), best_sample [N][3], loss_b [MH_NUM_BASE][N], arg_b int32 [MH_NUM_BASE][N]. */
int mh_pmvo_optimize(void* stream, const mh_views* views, const float* points, int64_t N,
                     const float* offsets, int32_t S, float conf_threshold,
                     float* ori, float* loss, uint8_t* high_conf,
                     int32_t* dbg_base_idx, float* dbg_base_val, float* dbg_best_sample,
                     float* dbg_loss_b, int32_t* dbg_arg_b,
                     void* workspace, int64_t workspace_bytes);

/* PMVO.refine loss part (PMVO.py:82,86-90): single-sample reprojection loss for next = p + dir*0.00125.
 * The head filter (loss[filter]=-1, PMVO.py:92) is applied by the caller from mh_head_count + scalp distance. */
int mh_pmvo_refine_loss(void* stream, const mh_views* views, const float* points, const float* dir, int64_t N,
                        float conf_threshold, float* loss);

/* ---- kNN + medoid (PMVO.py:605-641, 660-686; PMVO_utils.py:366-382) ----------------------------------- */
int64_t mh_knn_workspace_bytes(int64_t n_ref, int64_t n_query, int32_t k);
/* Exact k nearest neighbours (Euclidean, float64 distances like scipy.spatial.KDTree on float64 data), sorted by
 * (distance, index).  ref/query are float32 xyz.  idx int32 [n_query][k]. */
int mh_knn(void* stream, const float* ref, int64_t n_ref, const float* query, int64_t n_query, int32_t k,
           const double* bbox_host /*{min xyz, max xyz} of ref*/, double cell_size,
           int32_t* idx, void* workspace, int64_t workspace_bytes);
/* the same with float64 queries (float32 references): PMVO.refine step (iii) queries the KDTree of the selected points with
 * the float64 candidates of filter_unvisible.npy and casts them to float32 only afterwards (PMVO.py:660-671) */
int mh_knn_q64(void* stream, const float* ref, int64_t n_ref, const double* query, int64_t n_query, int32_t k,
           const double* bbox_host /*{min xyz, max xyz} of ref*/, double cell_size,
           int32_t* idx, void* workspace, int64_t workspace_bytes);
/* Nearest-reference distance only (k=1), float64 [n_query]: scalp_tree.query(points,k=1) (PMVO.py:104). */
int mh_nn_dist(void* stream, const double* ref, int64_t n_ref, const float* query, int64_t n_query, double* dist);
/* medoid of ori[nbr[i][0..K)] under |cos| (compute_points_similarity): out [n][3], out_k int32 [n] (may be NULL). */
int mh_medoid_gather(void* stream, const float* ori /*[n_ref][3]*/, const int32_t* nbr /*[n][K]*/, int64_t n,
                     int32_t K, float* out, int32_t* out_k);

/* In-place chunk update of PMVO.refine step (i) (PMVO.py:629-641): loss = head_filter ? 0.5 : upd_loss
 * (the -1 marker of PMVO.py:92 becomes 0.5 at :639); ori = center where |cos(center, ori)| < 0.95. */
int mh_refine_update(void* stream, const float* center, const float* upd_loss, const uint8_t* head_filter,
                     int64_t n, float* ori /*in/out [n][3]*/, float* loss /*out [n]*/);

/* The whole chunk-sequential pass of PMVO.refine step (i) (PMVO.py:608-641): for each chunk of sub_num points, in
 * order: centre = medoid of the neighbours' CURRENT orientations, loss = re-score of (point, centre), ori <- centre
 * where |cos| < 0.95 -- later chunks see earlier chunks' updates (Gauss-Seidel across chunks).  Runs as one
 * persistent dependency-ordered sweep kernel + one re-score launch over all points (see pmvo_refine.cu). */
int64_t mh_refine_chunks_workspace_bytes(int64_t n, int64_t sub_num);
/* The same pass as separate steps (a multi-GPU host replicates the sweep and shards the re-score over ranks):
 * mh_refine_sweep: ori [n][3] (input, untouched) -> ori_new [n][3] (updated orientations) and center [n][3] (medoids);
 * then mh_pmvo_refine_loss(points, center) -> upd; then mh_refine_finish(upd, head_filter) -> loss. */
int64_t mh_refine_sweep_workspace_bytes(int64_t n, int64_t sub_num);
int mh_refine_sweep(void* stream, const float* ori, const int32_t* nbr, int32_t K, int64_t n, int64_t sub_num,
                    float* ori_new, float* center, void* scratch, int64_t scratch_bytes);
int mh_refine_finish(void* stream, const float* upd_loss, const uint8_t* head_filter, int64_t n, float* loss);
/* The sweep spread over the GPUs of one node (same reference lines, PMVO.py:608-641).  Rank r owns every world-th block of
 * mh_refine_sweep_dist_block() consecutive points (global index of its q-th point: ((q / B) * world + rank) * B + q % B)
 * and passes only ITS points' neighbour lists (nbr_local [local_count][K]).  peer_ori_new / peer_center: HOST arrays of
 * `world` device pointers to every rank's full [n][3] copies in symmetric (peer-mapped) memory, entry `rank` being the
 * local copy; finished points are stored into all copies over NVLink, waits spin on the local copy only, so the exchange
 * overlaps the medoid work and no collective follows the kernel.  The caller fills every word of its ori_new copy with
 * 0xffffffff and passes a cross-rank barrier on the stream BEFORE the call and another one after it.  spin_seconds bounds
 * a wait (<= 0: 2 s); *error_flag (device) is 1 afterwards if one ran out; max_blocks > 0 caps the grid (the kernels of all
 * ranks must be resident together: one per GPU in production, several on one GPU in the tests).  Bit-identical to
 * mh_refine_sweep. */
int64_t mh_refine_sweep_dist_block(void);
int64_t mh_refine_sweep_dist_local_count(int64_t n, int32_t rank, int32_t world);
int mh_refine_sweep_dist(void* stream, const float* ori, const int32_t* nbr_local, int32_t K, int64_t n, int64_t sub_num,
                         int32_t rank, int32_t world, const uint64_t* peer_ori_new, const uint64_t* peer_center,
                         double spin_seconds, int32_t max_blocks, void* scratch, int64_t scratch_bytes, int32_t* error_flag);
int mh_refine_chunks(void* stream, const mh_views* views, const float* points, const int32_t* nbr, int32_t K,
                     const uint8_t* head_filter, int64_t n, int64_t sub_num, float conf_threshold,
                     float* ori /*in/out*/, float* loss /*out*/, void* scratch, int64_t scratch_bytes);

/* ---- voxel fusion (PMVO.py:695-726, PMVO_utils.p2v :386-404) ------------------------------------------ */
/* Two buffers: a per-call workspace, and a persistent `plane` (8 B per voxel) that must be all-zero on entry and
 * is left all-zero on exit (mh_voxel_fuse_plane_init zeroes a fresh one; re-initialise it after a failed call). */
int64_t mh_voxel_fuse_workspace_bytes(int64_t n_points, int32_t gx, int32_t gy, int32_t gz);
int64_t mh_voxel_fuse_plane_bytes(int32_t gx, int32_t gy, int32_t gz);
int mh_voxel_fuse_plane_init(void* stream, void* plane, int32_t gx, int32_t gy, int32_t gz);
/* points float32 [n][3] (world), dirs float32 [n][3].  Flips dirs to dir.y<=0 (PMVO.py:702-703), voxelises with
 * float64 index math (np.round half-even), takes the per-voxel medoid in original point order and writes the
 * fused volume as float4 [gz][gy][gx] = {ori.x, -ori.y, -ori.z, occ}: the layout/sign HairGrowing.__init__
 * builds from the .mat pair (HairGrow.py:45-55).  vox_index int32 [n] (linear id x*gy*gz+y*gz+z; may be NULL).
 * valid (optional): points with valid[i] == 0 are skipped, as if they had been removed from the arrays.
 * The volume's zero fill runs on an internal auxiliary stream, concurrent with the binning kernel, and is joined back
 * into `stream` before the medoid kernel writes its winners. */
int mh_voxel_fuse(void* stream, const float* points, const float* dirs, const uint8_t* valid /*[n] or NULL*/, int64_t n,
                  const double* voxel_min_host /*[3]*/, double voxel_size, int32_t gx, int32_t gy, int32_t gz,
                  void* volume /*float4 [gz][gy][gx]*/, int32_t* vox_index, void* plane, void* workspace,
                  int64_t workspace_bytes);
/* The same fusion, stopping at the per-voxel winners: winners float4 [capacity] = {ori.x, -ori.y, -ori.z, key bits}
 * (key = (z*gy + y)*gx + x as int32 bits; entries beyond *count carry key -1), count int32 [1] (may be NULL).
 * capacity >= min(n, voxels).  This is the multi-GPU exchange format: ranks fuse disjoint voxel sets (via `valid`),
 * all-gather their winner lists (16 B per occupied voxel) and scatter the union (mh_voxel_scatter) -- exactly the
 * dense all-reduce(SUM) of disjoint volumes, without moving the zeros. */
int mh_voxel_fuse_winners(void* stream, const float* points, const float* dirs, const uint8_t* valid, int64_t n,
                          const double* voxel_min_host, double voxel_size, int32_t gx, int32_t gy, int32_t gz,
                          void* winners, int64_t capacity, int32_t* count, int32_t* vox_index, void* plane,
                          void* workspace, int64_t workspace_bytes);
/* volume[key] = {w.x, w.y, w.z, 1} for the m winners with key >= 0; zero_fill != 0 clears the volume first (TMA bulk
 * stores from a zeroed shared-memory tile, one CTA per SM). */
int mh_voxel_scatter(void* stream, const void* winners, int64_t m, int32_t gx, int32_t gy, int32_t gz, void* volume,
                     int32_t zero_fill);
/* Synchronous, informational: largest per-voxel point count seen by the last mh_voxel_fuse on this workspace if some
 * voxel held more than 32 points (the slower overflow path), else 0. */
int mh_voxel_fuse_max_points(const void* workspace, int32_t* max_k_host);
/* Overwrite voxels with given orientations, last writer wins (raw.npy merge, PMVO.py:747-749). */
int mh_voxel_overwrite(void* stream, const float* points, const float* dirs, int64_t n,
                       const double* voxel_min_host, double voxel_size, int32_t gx, int32_t gy, int32_t gz,
                       void* volume, void* winner_ws /* int32 [gz*gy*gx] scratch */);
/* float4 volume -> the float64 arrays scipy.io.savemat receives (PMVO.py:753-756):
 * Occ [gy][gx][gz], Ori [gy][gx][3*gz] (index c*gz+z), world-frame signs. */
int mh_volume_to_mat(void* stream, const void* volume, int32_t gx, int32_t gy, int32_t gz, double* occ, double* ori);
/* and back (get_ground_truth_3D_occ/ori + HairGrowing.__init__, PMVO_utils.py:86-113, HairGrow.py:45-55). */
int mh_volume_from_mat(void* stream, const double* occ, const double* ori, int32_t gx, int32_t gy, int32_t gz,
                       void* volume);

/* ---- HairGrow trace (HairGrow.py:59-299) -------------------------------------------------------------- */
/* Pass 1: lengths.  seeds float32 [n][3] are the jittered seed positions (seed + 0.5 + U[0,0.5)^3 applied by the
 * caller so RNG draws can be injected).  n_fwd/n_bwd int32 [n]: accepted steps each way. */
int mh_trace_count(void* stream, const void* volume, int32_t gx, int32_t gy, int32_t gz, const float* seeds,
                   int64_t n, float thr_dot, int32_t max_steps, int32_t* n_fwd, int32_t* n_bwd);
/* Pass 2: write strands (backward part reversed, seed, forward part) at offsets[i] (in points) for strands with
 * n_fwd+n_bwd+1 >= min_len; others are skipped.  points_out float32 [total][3] voxel coordinates. */
int mh_trace_write(void* stream, const void* volume, int32_t gx, int32_t gy, int32_t gz, const float* seeds,
                   int64_t n, float thr_dot, int32_t max_steps, const int32_t* n_fwd, const int32_t* n_bwd,
                   const int64_t* offsets, int32_t min_len, float* points_out);
/* traceFromScalp (HairGrow.py:154-223): one pass, fixed-stride output [n][max_steps+1][3]; length int32 [n]
 * (0 when the reference returns None). */
int mh_trace_from_scalp(void* stream, const void* volume, int32_t gx, int32_t gy, int32_t gz, const float* roots,
                        const float* normals, int64_t n, float thr_dot, int32_t max_steps, int32_t max_inner,
                        float* points_out, int32_t* length);
/* Ordered acceptance (GenerateGuideStrandFromScalp / randomlyGenerateSegments flag logic, HairGrow.py:235-260,
 * 280-293): strands in order; reject when flag[seed voxel] >= 3, else accept and bump flag once per unique voxel
 * (mode 0: += 1; mode 1: = 1 and no gating, the scalp pass).  flag float32 [gz][gy][gx].  accepted uint8 [n].
 * Same result as the sequential loop; mode 0 runs in batches of 512 strands whose mutual dependencies (a strand
 * running through a later strand's seed voxel) are resolved in shared memory, mode 1 is order-free. */
int mh_accept_strands(void* stream, const float* points, const int64_t* offsets, const int32_t* lengths,
                      const float* seeds, int64_t n, int32_t gx, int32_t gy, int32_t gz, int32_t mode,
                      float* flag, uint8_t* accepted);

/* ---- candidate sampling (SamplePointsAroundmesh, Utils/PMVO_utils.py:316-339) ---------------------------------------- */
/* occ uint8 [gx][gy][gz] = 1 on the cell of every surface sample (y, z negated, float64 round half to even, clipped). */
int mh_sample_mark_cells(void* stream, const double* points /*[n][3]*/, int64_t n, const double* bbox_min_host, double vsize,
                         int32_t gx, int32_t gy, int32_t gz, uint8_t* occ);
/* out float64 [m*num_per_grid][3]: (cell + rnd) * vsize + bbox_min with y, z negated; cells int64 [m][3] in np.nonzero
 * order, tiled num_per_grid times; rnd float64 [m*num_per_grid][3] = the reference's np.random.random draws. */
int mh_sample_cells(void* stream, const int64_t* cells, int64_t m, int32_t num_per_grid, const double* rnd,
                    const double* bbox_min_host, double vsize, double* out);

/* ---- depth-map producer (render_bust_hair_depth, Utils/Render_utils.py:310-347; shaders :150-178) -------------------- */
/* Rasterises a triangle mesh (verts float32 [n][3] world frame, faces int32 [m][3]) as the reference's OpenGL pass does:
 * depth [H][W] = -z_cam / 2 of the nearest fragment (perspective-correct), 1 where nothing is drawn, row 0 on top, in the
 * pixel mapping of PMVO.project_points.  cam_record_host: mh_views_pack_camera_host record.  clear = 0 draws on top of the
 * meshes already in zbuf (the reference adds the bust to the same frame).  The reference needs moderngl + EGL, absent
 * here: this stage is restated, not pinned (oracle/render_oracle.py). */
int mh_render_depth(void* stream, const float* verts, int64_t n_verts, const int32_t* faces, int64_t n_faces,
                    const float* cam_record_host, int32_t H, int32_t W, float* depth, void* zbuf, int32_t clear);

/* ---- HairGrow connect stage (HairGrowing.find_connect_info, HairGrow.py:436-505, :548-584) ------------------------- */
/* For every strand (float64 points[offsets[i] .. +lengths[i]), world frame, metres) and each of its two ends: the partner
 * strand the reference would connect it to.  End-point candidates: the 50 nearest end points strictly within
 * connect_threshold, roots first and tips only if the roots give nothing (query + find_best_connect_strands); per
 * candidate the orientation test against dot_threshold, the strand-to-strand proximity rules of :565-578 and the loss
 * distance * (1 - |cos|), first minimum.  info int32 [n][4] = {root partner, its end, tip partner, its end}, partner -1 /
 * end 0 = none, end 1 = the partner's root, 2 = its tip.  *overflow is set when more than 256 end points fall inside one
 * query radius (the result would not be exact). */
int mh_connect_find(void* stream, const double* points, const int64_t* offsets, const int32_t* lengths, int64_t n_strands,
                    double connect_threshold, double dot_threshold, int32_t* info, int32_t* overflow);
/* Share of each strand's points on occupied voxels as HairGrow.py:514-523 evaluates it (points_to_voxel with the float32
 * voxel_min, torch.round, negative-index wrap), -1 when an index leaves the grid upwards.  shift (optional float64 [n][3]) is
 * added to every point of strand i first (the random perturbation of :531); voxel_space != 0: points are voxel coordinates
 * already (random_move_strands, PMVO_utils.py:629). */
int mh_strand_occupancy(void* stream, const double* points, const int64_t* offsets, const int32_t* lengths, int64_t n_strands,
                        const double* shift, const void* volume, int32_t gx, int32_t gy, int32_t gz, int32_t voxel_space,
                        double* frac);

/* mh_accept_strands mode 0 over all SMs: what depends on geometry alone (seed-voxel hashes, dependency rows, first visits
 * of a voxel) is computed for every batch of 512 strands at once, one CTA per batch; one CTA then walks the batches in
 * order doing only the flag reads, the contested-strand resolution and the flag bumps.  Same result as mh_accept_strands. */
int64_t mh_accept_strands_workspace_bytes(int64_t n, int64_t total_points);
int mh_accept_strands_ws(void* stream, const float* points, const int64_t* offsets, const int32_t* lengths, const float* seeds,
                         int64_t n, int64_t total_points, int32_t gx, int32_t gy, int32_t gz, float* flag, uint8_t* accepted,
                         void* workspace, int64_t workspace_bytes);

/* ---- strand smoothing (Utils/Utils.py:1148-1198 smnooth_strand / smooth_strands; HairGrow.py:914, :950, :975) ---- */
/* Per strand (points[offsets[i] .. +lengths[i])) and axis: least squares of [lap*L ; pos*I] x = [0 ; pos*s] with L the
 * second-difference operator (first differences at the ends), solved in float64 through the pentadiagonal normal
 * equations, rounded to float32 like the reference's in-place store.  Strands of fewer than 2 points are copied. */
int64_t mh_smooth_strands_workspace_bytes(int64_t total_points);
int mh_smooth_strands(void* stream, const float* points /*[T][3]*/, const int64_t* offsets /*[n]*/,
                      const int32_t* lengths /*[n]*/, int64_t n_strands, double lap_constraint, double pos_constraint,
                      float* points_out /*[T][3]*/, void* workspace, int64_t workspace_bytes, int64_t total_points);

/* ---- Gabor bank (GaborFilter.py:29-145, calc_orientation_maps.py:18-49) ------------------------------- */
/* calOrientationGabor.filter+forward for iter=1: image [H][W] -> orient [H][W] (radians), conf [H][W],
int64_t mh_gabor_workspace_bytes(int32_t H, int32_t W, int32_t n_filters);
int mh_gabor_orientation(void* stream, const float* image, int32_t H, int32_t W, const float* bank, int32_t n_filters,
                         int32_t ksize, float clamp_low, float clamp_high, float* orient, float* conf,
                         float* two_channel, void* workspace, int64_t workspace_bytes);
/* The same result on the tensor cores (tcgen05.mma kind::tf32 with a 3-term hi/lo split for fp32 accuracy, fp32
 * accumulators in TMEM, bank streamed by TMA bulk copies), with the per-pixel epilogue fused: the 180 responses of a
 * pixel never leave the chip.  bank_tc: the 180 x 17 x 17 bank pre-split and pre-laid-out by the host
 * (mh_gabor_tc_bank_bytes() bytes: [17 kernel rows][hi, lo][6 k-groups][24 filter groups][8][4] float32, zero padded).
 * workspace: H*W floats + 256 B.  Orientation indices agree with mh_gabor_orientation wherever the top-2 response margin
 * exceeds ~1e-6 of the largest response. */
int64_t mh_gabor_tc_bank_bytes(void);
int mh_gabor_orientation_tc(void* stream, const float* image, int32_t H, int32_t W, const void* bank_tc, int32_t n_filters,
                            float clamp_low, float clamp_high, float* orient, float* conf, float* two_channel,
                            void* workspace, int64_t workspace_bytes);
/* Generic filter-bank responses in float64 with periodic ('wrap') true convolution: calc_orients
 * (calc_orientation_maps.py:27-32).  bank [n][k][k] zero-padded to k x k; out |response| [n][H][W]. */
int mh_filterbank_wrap_f64(void* stream, const double* image, int32_t H, int32_t W, const double* bank,
                           int32_t n_filters, int32_t ksize, double* out_abs);
/* Separable Gaussian (scipy.ndimage.gaussian_filter, mode='nearest') difference: difference_of_gaussians. */
int mh_dog_f64(void* stream, const double* image, int32_t H, int32_t W, const double* k_lo, int32_t r_lo,
               const double* k_hi, int32_t r_hi, double* out, double* scratch /*[2][H][W]*/);

/* ---- host-side test hooks (no GPU needed): the same __host__ __device__ code the kernels run ---------- */
/* order of torch.topk(k=20, dim=0, largest, sorted) on CPU for one column of V values (PMVO.py:341). */
int mh_debug_topk_host(const float* values_host, int32_t V, int32_t k, int32_t* idx_host, float* val_host);

/* GPU test hook: number of (a0, a1, b) triples for which the shared-reciprocal division of the projection code
 * (mh_common.cuh: mh_div2) differs in any bit from the IEEE quotients a0/b, a1/b; added to *mismatches (uint64). */
int mh_debug_div2_check(void* stream, const float* a0, const float* a1, const float* b, int64_t n,
                        unsigned long long* mismatches);

#ifdef __cplusplus
}
#endif
#endif
