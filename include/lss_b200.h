/*
 * lss_b200.h -- C ABI of the B200-native Lift-Splat (camera -> BEV) hot path.
 *
 * The reference (fircarpediem/LSS2_Multimodal_nu) has no FFI of its own: the
 * boundary it offers is the nn.Module method surface (get_geometry /
 * get_cam_feats / voxel_pooling / get_voxels) plus one custom-op precedent,
 * QuickCumsum, a torch.autograd.Function (reference src/tools.py:192-218).
 * This header is what a binding for that path attaches to; every entry point
 * cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C types only; every pointer named d_* is a DEVICE pointer on the
 *     current CUDA device, every other pointer is a host pointer;
 *   - `stream` is a cudaStream_t passed as void*; a call only enqueues work on
 *     that stream, never synchronises, never allocates and keeps no global
 *     mutable state, so sequences are CUDA-graph capturable and re-entrant
 *     across streams / devices (one process per GPU under DDP);
 *   - return value: LSS_OK (0) or a negative LssStatus; lss_status_string()
 *     names it, lss_last_cuda_error() holds the CUDA error text (thread local);
 *   - tensors are dense, row-major, float32 / int32 unless stated; C % 4 == 0
 *     and feature pointers 16-byte aligned (128-bit vector access over C);
 *   - a "point" p is the flat index of (b, n, d, h, w) in a (B,N,D,fH,fW)
 *     array, exactly the reference's flattening (src/model_baseline.py:89-96);
 *   - a "rank" is the reference's voxel rank
 *       x*(nx1*nx2*B) + y*(nx2*B) + z*B + b        (src/model_baseline.py:106-109)
 *     held in int32; points that fail the bounds test (:99-101) carry the
 *     sentinel rank n_cells = nx0*nx1*nx2*B, so they sort behind every kept
 *     point.  n_cells must be < 2^31 - 1 and P < 2^31;
 *   - BEV tensors (d_bev, d_dbev) hold the logical (B, C*Z, X, Y) map of
 *     src/model_baseline.py:120-124 with channels innermost, i.e. (B, X, Y, Z*C)
 *     storage = torch channels_last strides; channel index = z*C + c.  One voxel
 *     is one contiguous line of C floats (128-bit vector access over C).  The
 *     reference's contiguous NCHW buffer is one transpose away (host wrapper).
 */
#ifndef LSS_B200_H_
#define LSS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSS_ABI_VERSION 3

typedef enum LssStatus {
  LSS_OK = 0,
  LSS_ERR_NULL_POINTER = -1,
  LSS_ERR_BAD_DIMENSION = -2,
  LSS_ERR_MISALIGNED = -3,
  LSS_ERR_WORKSPACE_TOO_SMALL = -4,
  LSS_ERR_UNSUPPORTED = -5,
  LSS_ERR_CUDA = -6
} LssStatus;

/* dtype of the feature tensors at the boundary (depth / logits / feat in, their gradients out);
 * arithmetic, the BEV map and dBEV are always float32 */
typedef enum LssDtype { LSS_F32 = 0, LSS_F16 = 1, LSS_BF16 = 2 } LssDtype;

/* Grid constants, the products of gen_dx_bx (reference src/tools.py:172-178). */
typedef struct LssGrid {
  float dx[3]; /* voxel size            */
  float bx[3]; /* centre of first voxel */
  int32_t nx[3]; /* voxel counts X, Y, Z */
} LssGrid;

/* Problem shape: B samples x N cameras x D depth bins x fH x fW pixels, C channels. */
typedef struct LssShape {
  int32_t B, N, D, fH, fW, C;
} LssShape;

int lss_abi_version(void);
const char* lss_status_string(int status);
const char* lss_last_cuda_error(void);

/* ------------------------------------------------------------------------- *
 * K0  camera preparation.  Replaces torch.inverse(post_rots) and
 *     rots.matmul(torch.inverse(intrins)) (src/model_baseline.py:60,66):
 *     d_inv_post_rots[i] = inverse(d_post_rots[i]), d_combine[i] = d_rots[i] @
 *     inverse(d_intrins[i]) for i < n_cams, all (n_cams,3,3) float32.  The LU
 *     sequence is the one torch.inverse executes on the CPU (see
 *     oracle/lss_oracle.py:inverse3x3), reproduced with explicit IEEE
 *     round-to-nearest intrinsics so the bits match.
 * ------------------------------------------------------------------------- */
int lss_camera_prep(const float* d_rots, const float* d_intrins, const float* d_post_rots,
                    int32_t n_cams, float* d_inv_post_rots, float* d_combine, void* stream);

/* ------------------------------------------------------------------------- *
 * K1  quantise + bounds test + rank from a dense geometry tensor.  Replaces
 *     src/model_baseline.py:92-109 for the literal voxel_pooling(geom, x) call.
 *     d_geom   (P,3) float32 ego-frame xyz, P = B * points_per_sample
 *     d_coords (P,3) int32 truncated voxel coordinates            [optional]
 *     d_kept   (P)   uint8 bounds-test result                     [optional]
 *     d_ranks  (P)   int32 rank, or n_cells for dropped points
 *     d_cells  (P)   int32 output cell ((b*X + x)*Y + y)*Z + z, or -1 [optional]
 * ------------------------------------------------------------------------- */
int lss_quantize_rank(const float* d_geom, const LssGrid* grid, int32_t B, int64_t P,
                      int32_t* d_coords, uint8_t* d_kept, int32_t* d_ranks, int32_t* d_cells,
                      void* stream);

/* ------------------------------------------------------------------------- *
 * K1' fused frustum geometry -> rank.  Replaces get_geometry
 *     (src/model_baseline.py:50-70) followed by :92-109 without writing the
 *     (B,N,D,fH,fW,3) tensor.  The frustum is passed as its three axes
 *     (d_us[fW], d_vs[fH], d_ds[D]; src/model_baseline.py:41-44), the cameras
 *     as the K0 products plus post_trans / trans, all (B*N, ...) float32.
 *     d_geom (P,3) is written only when non-null (the get_geometry API).
 * ------------------------------------------------------------------------- */
int lss_geometry_rank(const float* d_us, const float* d_vs, const float* d_ds,
                      const float* d_inv_post_rots, const float* d_post_trans,
                      const float* d_combine, const float* d_trans, const LssGrid* grid,
                      const LssShape* shape, float* d_geom, int32_t* d_coords, uint8_t* d_kept,
                      int32_t* d_ranks, int32_t* d_cells, void* stream);

/* ------------------------------------------------------------------------- *
 * K2  stable LSD radix sort of the ranks.  Replaces ranks.argsort()
 *     (src/model_baseline.py:110) and the three gathers of :111.
 *     d_sorted_ranks[i], d_sorted_points[i] for i < P: ranks ascending, ties in
 *     ascending point index (== torch's stable order); the first K entries are
 *     the kept points, i.e. d_sorted_points[:K] == nonzero(kept)[sorts].
 *     Only ceil(log2(n_cells+1)) key bits are sorted.
 *     Workspace: lss_sort_workspace_bytes(P, n_cells) bytes, 16-byte aligned,
 *     ZERO-FILLED before first use; a successful call leaves it zero again.
 * ------------------------------------------------------------------------- */
size_t lss_sort_workspace_bytes(int64_t P, int32_t n_cells);
int lss_sort_ranks(const int32_t* d_ranks, int64_t P, int32_t n_cells, int32_t* d_sorted_ranks,
                   int32_t* d_sorted_points, void* d_workspace, size_t workspace_bytes,
                   void* stream);

/* ------------------------------------------------------------------------- *
 * K3  interval detection over the sorted ranks.  Replaces the boundary mask of
 *     QuickCumsum.forward (src/tools.py:196-197) and supplies what the cumsum /
 *     difference (:195,:200) needed it for.
 *     d_last_mask   (P) uint8, 1 at the last point of each run of equal ranks,
 *                   0 beyond the K kept points                    [optional]
 *     d_sorted_cells (P) int32 output cell ((b*X+x)*Y+y)*Z+z of each sorted point
 *                   (first K entries valid)                        [optional]
 *     d_cell_range  (n_cells,2) int32 [start, end) of every output cell's run in
 *                   the sorted order, indexed by OUTPUT cell; start >= end means
 *                   empty; must be zero-filled by the caller
 *     d_counts      (2) int32: {K kept points, V occupied cells}; zero-filled
 *                   by the caller
 * ------------------------------------------------------------------------- */
int lss_intervals(const int32_t* d_sorted_ranks, int64_t P, const LssGrid* grid, int32_t B,
                  uint8_t* d_last_mask, int32_t* d_sorted_cells, int32_t* d_cell_range,
                  int32_t* d_counts, void* stream);

/* ------------------------------------------------------------------------- *
 * The plan: everything between the calibration tensors and the per-voxel point
 * lists (K0 -> K1' -> sort -> intervals) in one call, i.e. the geometry half of
 * get_voxels (src/model_baseline.py:128-131) + the argsort of :110 + the
 * interval detection of src/tools.py:196-197.  The plan depends only on the
 * calibration, so evaluation code can build it once per rig and reuse it, and
 * training code can build it while the image backbone runs.
 *
 * The sort key is the OUTPUT CELL in tile-major order: the BEV plane of every
 * sample is cut into T x T tiles (T = lss_plan_key_tile() = 8) and
 *   key = ((((b*XT + x/T)*YT + y/T)*T + x%T)*T + y%T)*Z + z,  XT = ceil(X/T), YT = ceil(Y/T)
 * -- the digits of the reference's rank x*(Y*Z*B) + y*(Z*B) + z*B + b regrouped, a
 * bijection of the rank on the existing cells -- and the sort is stable, so each
 * voxel's run holds the reference's points in the reference's order; only the
 * order in which runs follow one another differs (neighbours in the list are
 * neighbours on the map).  lss_sort_ranks is the bit-exact argsort of the rank itself.
 * n_keys = lss_plan_key_count(grid, B) = B*XT*YT*T*T*Z >= n_cells.
 *   d_cells         (P)          output cell ((b*X + x)*Y + y)*Z + z of every point, -1 if it
 *                                fails the bounds test
 *   d_key_start     (n_keys+1)   key k owns d_sorted_rec[d_key_start[k] .. d_key_start[k+1]);
 *                                equal bounds = empty voxel; d_key_start[n_keys] = K.
 *                                This table is the interval detection (K3).
 *   d_sorted_rec    (P,2)        {output cell, point id} ordered by (key, point id): column 1 of
 *                                the first K rows is nonzero(kept)[argsort] regrouped by key,
 *                                column 0 the voxel each sorted point falls into; rows beyond the
 *                                K kept points are {-1, 0}.  16-byte aligned.
 *   d_counts        (2)          {K kept points, V occupied voxels}
 *   workspace: lss_plan_workspace_bytes(shape, grid) bytes, 16-byte aligned; its first
 *   lss_plan_workspace_control_bytes(shape, grid) bytes must be ZERO before the first use
 *   and a successful call leaves them zero again.
 * lss_build_plan_from_geom is the same from a materialised geometry tensor d_geom
 * (P,3) (the literal voxel_pooling(geom_feats, x) signature, src/model_baseline.py:84).
 * ------------------------------------------------------------------------- */
size_t lss_plan_workspace_bytes(const LssShape* shape, const LssGrid* grid);
size_t lss_plan_workspace_control_bytes(const LssShape* shape, const LssGrid* grid);
int64_t lss_plan_key_count(const LssGrid* grid, int32_t B);
int lss_plan_key_tile(void);
int lss_build_plan(const float* d_us, const float* d_vs, const float* d_ds, const float* d_rots,
                   const float* d_trans, const float* d_intrins, const float* d_post_rots,
                   const float* d_post_trans, const LssGrid* grid, const LssShape* shape,
                   int32_t* d_cells, int32_t* d_key_start, int32_t* d_sorted_rec, int32_t* d_counts,
                   void* d_workspace, size_t workspace_bytes, void* stream);
size_t lss_plan_from_geom_workspace_bytes(int64_t P, const LssGrid* grid, int32_t B);
int lss_build_plan_from_geom(const float* d_geom, const LssGrid* grid, int32_t B, int64_t P,
                             int32_t* d_cells, int32_t* d_key_start, int32_t* d_sorted_rec, int32_t* d_counts,
                             void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * K4a dense pooling.  Replaces x[kept][sorts] -> QuickCumsum -> zeros ->
 *     index_put -> cat(unbind) (src/model_baseline.py:102-124) for a
 *     materialised frustum tensor d_x (P,C), given a plan.  Every output element
 *     is written (zeros for empty voxels).  C <= 128, P * C/4 < 2^31.
 *     Backward (QuickCumsum.backward src/tools.py:211-218 + index backward):
 *     d_dx[p,:] = d_dbev[cell(p),:] for kept points, 0 otherwise.
 * ------------------------------------------------------------------------- */
int lss_pool_dense_fwd(const float* d_x, const int32_t* d_sorted_rec, const int32_t* d_key_start,
                       const LssGrid* grid, int32_t B, int32_t C, int64_t P, float* d_bev, void* stream);
int lss_pool_dense_bwd(const float* d_dbev, const int32_t* d_cells, const LssGrid* grid,
                       int32_t B, int32_t C, int64_t P, float* d_dx, void* stream);

/* ------------------------------------------------------------------------- *
 * Lift staging.  The (B*N,C,D,fH,fW) outer product of src/modules.py:84 and the
 *     permute/reshape copies of src/model_baseline.py:79-80,89 are never formed.
 *     What the pooling kernels need instead:
 *     - the context features with one contiguous row per pixel:
 *       lss_feat_stage: d_feat (BN,C,fH,fW) -> d_feat_t (BN*fH*fW, C) float32
 *       (2.2 MB at the headline config).  `feat_batch_stride` (elements) lets the
 *       input be a channel slice of ONE conv output (B*N, D+C, fH, fW) as
 *       CamEncode produces it (src/modules.py:74,82-84); `dtype` (LssDtype) is its
 *       element type (the AMP scripts hand over half tensors,
 *       train_vovnet_transformer.py:196);
 *     - the depth distribution indexed by POINT id, which is its native layout
 *       (BN, D, fH*fW): it is read IN PLACE (any LssDtype, batch stride);
 *     - only when the distribution still has to be computed from logits:
 *       lss_depth_softmax: d_logits (BN, D, fH, fW) (batch stride, dtype) ->
 *       d_depth (BN, D, fH*fW) float32 = x[:, :D].softmax(dim=1), src/modules.py:76-77.
 * ------------------------------------------------------------------------- */
int lss_feat_stage(const void* d_feat, int64_t feat_batch_stride, const LssShape* shape, int32_t dtype,
                   float* d_feat_t, void* stream);
int lss_depth_softmax(const void* d_logits, int64_t logits_batch_stride, const LssShape* shape, int32_t dtype,
                      float* d_depth, void* stream);

/* ------------------------------------------------------------------------- *
 * K4  fused lift + splat forward.  bev[cell, c] = sum over the cell's points of
 *     depth[point] * feat[pixel(point), c]; the frustum tensor never exists.
 *     Replaces src/modules.py:84 + src/model_baseline.py:79-80,89,102-124 +
 *     src/tools.py:194-208.  d_depth (BN, D*fH*fW) with batch stride (elements)
 *     and dtype, d_feat_t from lss_feat_stage, and a plan.
 * K5  fused backward.  d_ddepth (BN,D,fH,fW) = sum_c g*feat, d_dfeat
 *     (BN,C,fH,fW) = sum_d g*depth with g = dbev[cell(point), :] (an exact
 *     gather, as QuickCumsum.backward is); points that were dropped contribute 0.
 *     The outputs have batch strides (the two gradients may be channel slices of
 *     one tensor shaped like the conv output) and element type out_dtype.  With
 *     softmax != 0, d_depth holds the probabilities lss_depth_softmax wrote and
 *     the first output receives the gradient of the depth LOGITS,
 *     p * (d_depth - sum_d p * d_depth) (autograd of src/modules.py:77; D <= 128).
 *     <g, feat> is accumulated in float64.
 * ------------------------------------------------------------------------- */
int lss_liftsplat_fwd(const void* d_depth, int64_t depth_batch_stride, int32_t depth_dtype,
                      const float* d_feat_t, const int32_t* d_sorted_rec, const int32_t* d_key_start,
                      const LssGrid* grid, const LssShape* shape, float* d_bev, void* stream);
int lss_liftsplat_bwd(const float* d_dbev, const void* d_depth, int64_t depth_batch_stride, int32_t depth_dtype,
                      const float* d_feat_t, const int32_t* d_cells, const LssGrid* grid, const LssShape* shape,
                      int32_t softmax, int32_t out_dtype, void* d_ddepth_or_dlogits, int64_t ddepth_batch_stride,
                      void* d_dfeat, int64_t dfeat_batch_stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSS_B200_H_ */
