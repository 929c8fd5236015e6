// extern "C" entry points declared in include/lss_b200.h.
#include <stdlib.h>

#include "lss_common.cuh"
#include "lss_geometry.cuh"
#include "lss_pool.cuh"
#include "lss_sort.cuh"
#include "lss_plan.cuh"

namespace lss {
char* cuda_error_buffer() {
  static thread_local char buf[256] = {0};
  return buf;
}

static int check_shape(const LssShape* s) {
  if (!s) return LSS_ERR_NULL_POINTER;
  if (s->B <= 0 || s->N <= 0 || s->D <= 0 || s->fH <= 0 || s->fW <= 0 || s->C <= 0)
    return LSS_ERR_BAD_DIMENSION;
  const long long P = (long long)s->B * s->N * s->D * s->fH * s->fW;
  if (P >= (1ll << 30)) return LSS_ERR_BAD_DIMENSION;
  return LSS_OK;
}
static inline long long shape_points(const LssShape* s) {
  return (long long)s->B * s->N * s->D * s->fH * s->fW;
}
static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int launch_geometry(const GeomArgs& ga, const GridDev& g, const LssShape* sh,
                           const PointOut& out, const SortDigits* sd, cudaStream_t st) {
  const int ppc = sh->D * sh->fH * sh->fW;
  int chunks = (ppc + 2047) / 2048;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, sh->B * sh->N);
  FastDiv div_hw(sh->fH * sh->fW), div_w(sh->fW);
  if (sd && sd->hist) {
    const size_t smem = (size_t)sd->passes * sd->stride * sizeof(uint32_t);
    geometry_rank_kernel<true><<<grid, kGeomThreads, smem, st>>>(ga, g, div_hw, div_w, out, *sd);
  } else {
    SortDigits none;
    memset(&none, 0, sizeof(none));
    geometry_rank_kernel<false><<<grid, kGeomThreads, 0, st>>>(ga, g, div_hw, div_w, out, none);
  }
  LSS_LAUNCH_CHECK("geometry_rank_kernel");
  return LSS_OK;
}

static int launch_intervals(const int32_t* sorted_ranks, long long P, const GridDev& g,
                            uint8_t* last_mask, int32_t* sorted_cells, int32_t* cell_range, int32_t* counts,
                            uint32_t* wipe, long long wipe_words, cudaStream_t st) {
  IntervalArgs a;
  a.sorted_ranks = sorted_ranks; a.P = P; a.g = g;
  a.div_b = FastDiv(g.B); a.div_z = FastDiv(g.nx[2]); a.div_y = FastDiv(g.nx[1]);
  a.last_mask = last_mask; a.sorted_cells = sorted_cells;
  a.cell_range = reinterpret_cast<int2*>(cell_range); a.counts = counts;
  a.wipe = wipe; a.wipe_words = wipe_words;
  long long blocks = (P + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  intervals_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  LSS_LAUNCH_CHECK("intervals_kernel");
  return LSS_OK;
}

// ---- the plan: P1 cells -> P2 scan -> P3 scatter -> P4 order (lss_plan.cuh) ---------------
static int run_plan(const GeomArgs* ga, const float* dense_geom, const GridDev& g, long long P,
                    long long points_per_sample, int32_t* d_cells, int32_t* d_cell_start,
                    int32_t* d_sorted_points, int32_t* d_sorted_cells, int32_t* d_counts, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
  KeyMap km;
  int rck = make_keymap(g, &km);
  if (rck) return rck;
  const PlanWorkspace pw = make_plan_workspace(P, km.n_keys);
  LSS_REQUIRE(ws_bytes >= pw.total_bytes, LSS_ERR_WORKSPACE_TOO_SMALL);
  LSS_REQUIRE(aligned16(ws) && aligned16(d_cell_start), LSS_ERR_MISALIGNED);
  char* w = static_cast<char*>(ws);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(w + pw.off_cnt);
  uint32_t* state = reinterpret_cast<uint32_t*>(w + pw.off_state);
  uint32_t* ctl = reinterpret_cast<uint32_t*>(w + pw.off_ctl);
  int32_t* tmp_pt = reinterpret_cast<int32_t*>(w + pw.off_tmp_pt);
  int32_t* long_list = reinterpret_cast<int32_t*>(w + pw.off_long);

  PlanCellsArgs ca;
  memset(&ca, 0, sizeof(ca));
  ca.grid = g; ca.keys = km; ca.P = P; ca.cells = d_cells; ca.cnt = cnt; ca.counts = d_counts; ca.ctl = ctl;
  const unsigned tiles = (unsigned)((P + kPlanTile - 1) / kPlanTile);
  if (ga) {
    ca.geom = *ga;
    const int hw = ga->fH * ga->fW;
    LSS_REQUIRE((kPlanTile - 1) / (ga->D * hw) + 2 <= kPlanMaxCams, LSS_ERR_UNSUPPORTED);
    ca.div_ppc = FastDiv((uint32_t)(ga->D * hw)); ca.div_hw = FastDiv((uint32_t)hw);
    ca.div_w = FastDiv((uint32_t)ga->fW); ca.div_n = FastDiv((uint32_t)ga->N);
    plan_cells_kernel<false><<<tiles, kPlanThreads, 0, st>>>(ca);
  } else {
    ca.dense_geom = dense_geom;
    ca.div_pps = FastDiv((uint32_t)points_per_sample);
    plan_cells_kernel<true><<<tiles, kPlanThreads, 0, st>>>(ca);
  }
  LSS_LAUNCH_CHECK("plan_cells_kernel");

  PlanScanArgs sa;
  sa.cnt = cnt; sa.n = km.n_keys; sa.tiles = pw.scan_tiles; sa.cell_start = d_cell_start;
  sa.state = state; sa.ctl = ctl; sa.counts = d_counts; sa.long_list = long_list;
  plan_scan_kernel<<<(unsigned)pw.scan_tiles, kPlanThreads, 0, st>>>(sa);
  LSS_LAUNCH_CHECK("plan_scan_kernel");

  PlanScatterArgs sc;
  sc.cells = d_cells; sc.P = P; sc.keys = km; sc.cnt = cnt; sc.cell_start = d_cell_start;
  sc.tmp_pt = tmp_pt; sc.sorted_cells = d_sorted_cells; sc.state = state; sc.scan_tiles = pw.scan_tiles; sc.ctl = ctl;
  long long blocks = (P + kPlanThreads - 1) / kPlanThreads;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  plan_scatter_kernel<<<(unsigned)blocks, kPlanThreads, 0, st>>>(sc);
  LSS_LAUNCH_CHECK("plan_scatter_kernel");

  PlanOrderArgs oa;
  oa.tmp_pt = tmp_pt; oa.sorted_cells = d_sorted_cells; oa.cell_start = d_cell_start; oa.keys = km; oa.n_cells = km.n_keys;
  oa.P = P; oa.sorted_points = d_sorted_points; oa.ctl = ctl; oa.long_list = long_list;
  plan_order_kernel<<<(unsigned)blocks, kPlanThreads, 0, st>>>(oa);
  LSS_LAUNCH_CHECK("plan_order_kernel");
  return LSS_OK;
}

template <int kLanes>
static int launch_bwd(const PoolBwdArgs& a, int blocks, cudaStream_t st) {
  if (a.softmax || a.out_dtype != LSS_F32) liftsplat_bwd_nhwc_kernel<kLanes, true><<<blocks, kBwdThreads, 0, st>>>(a);
  else liftsplat_bwd_nhwc_kernel<kLanes, false><<<blocks, kBwdThreads, 0, st>>>(a);
  LSS_LAUNCH_CHECK("liftsplat_bwd_nhwc_kernel");
  return LSS_OK;
}
template <bool kFused>
static int launch_pool_fwd(const PoolFwdArgs& a, int blocks, cudaStream_t st) {
  if (a.C <= 32) pool_fwd_nhwc_kernel<kFused, 1><<<blocks, kPoolThreads, 0, st>>>(a);
  else if (a.C <= 64 && a.C % 2 == 0) pool_fwd_nhwc_kernel<kFused, 2><<<blocks, kPoolThreads, 0, st>>>(a);
  else pool_fwd_nhwc_kernel<kFused, 4><<<blocks, kPoolThreads, 0, st>>>(a);
  LSS_LAUNCH_CHECK("pool_fwd_nhwc_kernel");
  return LSS_OK;
}
}  // namespace lss

using namespace lss;

extern "C" {

int lss_abi_version(void) { return LSS_ABI_VERSION; }

const char* lss_status_string(int status) {
  switch (status) {
    case LSS_OK: return "ok";
    case LSS_ERR_NULL_POINTER: return "null pointer argument";
    case LSS_ERR_BAD_DIMENSION: return "bad dimension (non-positive, too large, or inconsistent)";
    case LSS_ERR_MISALIGNED: return "pointer not 16-byte aligned or C not a multiple of 4";
    case LSS_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case LSS_ERR_UNSUPPORTED: return "unsupported configuration";
    case LSS_ERR_CUDA: return "CUDA error (see lss_last_cuda_error)";
    default: return "unknown status";
  }
}

const char* lss_last_cuda_error(void) { return cuda_error_buffer(); }

int lss_camera_prep(const float* d_rots, const float* d_intrins, const float* d_post_rots,
                    int32_t n_cams, float* d_inv_post_rots, float* d_combine, void* stream) {
  LSS_REQUIRE(d_rots && d_intrins && d_post_rots && d_inv_post_rots && d_combine, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(n_cams > 0, LSS_ERR_BAD_DIMENSION);
  camera_prep_kernel<<<(n_cams + 63) / 64, 64, 0, as_stream(stream)>>>(
      d_rots, d_intrins, d_post_rots, n_cams, d_inv_post_rots, d_combine);
  LSS_LAUNCH_CHECK("camera_prep_kernel");
  return LSS_OK;
}

int lss_quantize_rank(const float* d_geom, const LssGrid* grid, int32_t B, int64_t P,
                      int32_t* d_coords, uint8_t* d_kept, int32_t* d_ranks, int32_t* d_cells,
                      void* stream) {
  LSS_REQUIRE(d_geom && d_ranks, LSS_ERR_NULL_POINTER);
  GridDev g;
  int rc = make_grid(grid, B, &g);
  if (rc) return rc;
  LSS_REQUIRE(P > 0 && P < (1ll << 30) && P % B == 0, LSS_ERR_BAD_DIMENSION);
  PointOut out{d_coords, d_kept, d_ranks, d_cells};
  SortDigits none;
  memset(&none, 0, sizeof(none));
  long long blocks = (P + kGeomThreads - 1) / kGeomThreads;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  quantize_rank_kernel<false><<<(unsigned)blocks, kGeomThreads, 0, as_stream(stream)>>>(
      d_geom, g, P, P / B, out, none);
  LSS_LAUNCH_CHECK("quantize_rank_kernel");
  return LSS_OK;
}

int lss_geometry_rank(const float* d_us, const float* d_vs, const float* d_ds,
                      const float* d_inv_post_rots, const float* d_post_trans,
                      const float* d_combine, const float* d_trans, const LssGrid* grid,
                      const LssShape* shape, float* d_geom, int32_t* d_coords, uint8_t* d_kept,
                      int32_t* d_ranks, int32_t* d_cells, void* stream) {
  LSS_REQUIRE(d_us && d_vs && d_ds && d_inv_post_rots && d_post_trans && d_combine && d_trans &&
                  d_ranks, LSS_ERR_NULL_POINTER);
  int rc = check_shape(shape);
  if (rc) return rc;
  GridDev g;
  rc = make_grid(grid, shape->B, &g);
  if (rc) return rc;
  GeomArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.us = d_us; ga.vs = d_vs; ga.ds = d_ds;
  ga.inv_post_rots = d_inv_post_rots; ga.post_trans = d_post_trans;
  ga.combine = d_combine; ga.trans = d_trans;
  ga.raw = 0; ga.N = shape->N; ga.D = shape->D; ga.fH = shape->fH; ga.fW = shape->fW;
  ga.geom = d_geom;
  PointOut out{d_coords, d_kept, d_ranks, d_cells};
  return launch_geometry(ga, g, shape, out, nullptr, as_stream(stream));
}

size_t lss_sort_workspace_bytes(int64_t P, int32_t n_cells) {
  if (P <= 0 || n_cells <= 0) return 0;
  return make_sort_plan(P, n_cells).total_bytes;
}

int lss_sort_ranks(const int32_t* d_ranks, int64_t P, int32_t n_cells, int32_t* d_sorted_ranks,
                   int32_t* d_sorted_points, void* d_workspace, size_t workspace_bytes,
                   void* stream) {
  LSS_REQUIRE(d_ranks && d_sorted_ranks && d_sorted_points && d_workspace, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(P > 0 && P < (1ll << 30) && n_cells > 0 && n_cells < 0x7fffffff, LSS_ERR_BAD_DIMENSION);
  LSS_REQUIRE(aligned16(d_workspace), LSS_ERR_MISALIGNED);
  const SortPlan s = make_sort_plan(P, n_cells);
  LSS_REQUIRE(workspace_bytes >= s.total_bytes, LSS_ERR_WORKSPACE_TOO_SMALL);
  cudaStream_t st = as_stream(stream);
  char* w = static_cast<char*>(d_workspace);
  LSS_CUDA_TRY(cudaMemsetAsync(w + s.off_control, 0, s.control_bytes, st), "memset sort control");
  SortDigits sd = sort_digits(s, d_workspace);
  long long blocks = (P + 256 * 8 - 1) / (256 * 8);
  const long long cap = (long long)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  histogram_kernel<<<(unsigned)blocks, 256, (size_t)sd.passes * sd.stride * 4, st>>>(d_ranks, P, sd);
  LSS_LAUNCH_CHECK("histogram_kernel");
  return run_sort_passes(s, d_ranks, d_sorted_ranks, d_sorted_points, P, d_workspace, st);
}

int lss_intervals(const int32_t* d_sorted_ranks, int64_t P, const LssGrid* grid, int32_t B,
                  uint8_t* d_last_mask, int32_t* d_sorted_cells, int32_t* d_cell_range,
                  int32_t* d_counts, void* stream) {
  LSS_REQUIRE(d_sorted_ranks && d_cell_range, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(P > 0 && P < (1ll << 30), LSS_ERR_BAD_DIMENSION);
  GridDev g;
  int rc = make_grid(grid, B, &g);
  if (rc) return rc;
  return launch_intervals(d_sorted_ranks, P, g, d_last_mask, d_sorted_cells, d_cell_range, d_counts,
                          nullptr, 0, as_stream(stream));
}

static int pool_fwd_common(bool fused, const float* d_depth_t, const float* d_feat_t,
                           const float* d_x, const int32_t* d_sorted_points, const int32_t* d_sorted_cells,
                           const int32_t* d_cell_start, const LssGrid* grid, int32_t B, int32_t C,
                           int32_t D, int32_t HW, long long dhw, long long P, int32_t layout, float* d_bev,
                           cudaStream_t st) {
  LSS_REQUIRE(d_sorted_points && d_sorted_cells && d_cell_start && d_bev, LSS_ERR_NULL_POINTER);
  GridDev g;
  int rc = make_grid(grid, B, &g);
  if (rc) return rc;
  LSS_REQUIRE(C > 0 && C % 4 == 0, LSS_ERR_MISALIGNED);
  LSS_REQUIRE(C <= 128, LSS_ERR_UNSUPPORTED);
  LSS_REQUIRE(aligned16(d_bev), LSS_ERR_MISALIGNED);
  LSS_REQUIRE(layout == LSS_BEV_NHWC, LSS_ERR_UNSUPPORTED);
  LSS_REQUIRE(P > 0 && P < (1ll << 30), LSS_ERR_BAD_DIMENSION);
  PoolFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.depth_t = d_depth_t; a.feat_t = d_feat_t; a.x = d_x;
  a.sorted_points = d_sorted_points; a.sorted_cells = d_sorted_cells; a.cell_start = d_cell_start;
  a.bev = d_bev; a.P = P;
  rc = make_keymap(g, &a.keys);
  if (rc) return rc;
  a.C = C; a.D = D; a.HW = HW;
  a.div_dhw = FastDiv((uint32_t)(dhw > 0 ? dhw : 1)); a.div_hw = FastDiv((uint32_t)(HW > 0 ? HW : 1));
  a.div_g4 = FastDiv((uint32_t)(C / 4));
  // fill CTAs: one per SM by default, so the zero stream runs in the background for the whole kernel
  long long fill = sm_count();
  const long long fill_blocks = ((long long)a.keys.n_keys + 31) / 32;
  if (fill * kPoolWarps > fill_blocks) fill = (fill_blocks + kPoolWarps - 1) / kPoolWarps;
  if (const char* e = getenv("LSS_FILL_CTAS")) fill = atoi(e);   // tuning knob
  if (fill < 1) fill = 1;
  a.fill_ctas = (int)fill;
  const long long reduce = (P + (long long)kPoolChunk * kPoolWarps - 1) / ((long long)kPoolChunk * kPoolWarps);
  const long long blocks = fill + reduce;
  LSS_REQUIRE(blocks < (1ll << 31), LSS_ERR_BAD_DIMENSION);
  return fused ? launch_pool_fwd<true>(a, (int)blocks, st) : launch_pool_fwd<false>(a, (int)blocks, st);
}

int lss_pool_dense_fwd(const float* d_x, const int32_t* d_sorted_points, const int32_t* d_sorted_cells,
                       const int32_t* d_cell_start, const LssGrid* grid, int32_t B, int32_t C, int64_t P,
                       int32_t layout, float* d_bev, void* stream) {
  LSS_REQUIRE(d_x, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(aligned16(d_x), LSS_ERR_MISALIGNED);
  LSS_REQUIRE((long long)P * (C / 4) < (1ll << 32), LSS_ERR_BAD_DIMENSION);   // 16-byte row offsets in 32 bits
  return pool_fwd_common(false, nullptr, nullptr, d_x, d_sorted_points, d_sorted_cells, d_cell_start, grid, B,
                         C, 1, 1, 1, P, layout, d_bev, as_stream(stream));
}

int lss_pool_dense_bwd(const float* d_dbev, const int32_t* d_cells, const LssGrid* grid,
                       int32_t B, int32_t C, int64_t P, int32_t layout, float* d_dx,
                       void* stream) {
  LSS_REQUIRE(d_dbev && d_cells && d_dx, LSS_ERR_NULL_POINTER);
  GridDev g;
  int rc = make_grid(grid, B, &g);
  if (rc) return rc;
  LSS_REQUIRE(C > 0 && C % 4 == 0 && aligned16(d_dbev) && aligned16(d_dx), LSS_ERR_MISALIGNED);
  LSS_REQUIRE(layout == LSS_BEV_NHWC, LSS_ERR_UNSUPPORTED);
  const int G = C / 4;
  const long long n = (long long)P * G;
  LSS_REQUIRE(P > 0 && n < (1ll << 31), LSS_ERR_BAD_DIMENSION);
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  pool_dense_bwd_nhwc_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(d_dbev), d_cells, n, G, FastDiv(G),
      reinterpret_cast<float4*>(d_dx));
  LSS_LAUNCH_CHECK("pool_dense_bwd_nhwc_kernel");
  return LSS_OK;
}

static int lift_stage_common(const void* d_depth, long long depth_bs, const void* d_feat, long long feat_bs,
                             const LssShape* shape, int softmax, int dtype, float* d_depth_t, float* d_feat_t,
                             cudaStream_t st) {
  LSS_REQUIRE(d_depth && d_feat && d_depth_t && d_feat_t, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(valid_dtype(dtype), LSS_ERR_UNSUPPORTED);
  int rc = check_shape(shape);
  if (rc) return rc;
  const int HW = shape->fH * shape->fW, BN = shape->B * shape->N;
  const int R = shape->D > shape->C ? shape->D : shape->C;
  LSS_REQUIRE(BN * 2 <= 65535, LSS_ERR_BAD_DIMENSION);
  LSS_REQUIRE(depth_bs >= (long long)shape->D * HW && feat_bs >= (long long)shape->C * HW, LSS_ERR_BAD_DIMENSION);
  dim3 grid((HW + 31) / 32, (R + 31) / 32, BN * 2);
  if (softmax) {
    LSS_REQUIRE(shape->D <= kSoftmaxMaxD, LSS_ERR_UNSUPPORTED);
    lift_stage_softmax_kernel<<<dim3((HW + 31) / 32, BN), 256, 0, st>>>(d_depth, depth_bs, dtype, shape->D, HW,
                                                                        d_depth_t);
    LSS_LAUNCH_CHECK("lift_stage_softmax_kernel");
    lift_stage_kernel<<<grid, 256, 0, st>>>(nullptr, 0, d_feat, feat_bs, dtype, shape->D, shape->C, HW, d_depth_t,
                                            d_feat_t);
  } else {
    lift_stage_kernel<<<grid, 256, 0, st>>>(d_depth, depth_bs, d_feat, feat_bs, dtype, shape->D, shape->C, HW,
                                            d_depth_t, d_feat_t);
  }
  LSS_LAUNCH_CHECK("lift_stage_kernel");
  return LSS_OK;
}

int lss_lift_stage(const float* d_depth, const float* d_feat, const LssShape* shape,
                   float* d_depth_t, float* d_feat_t, void* stream) {
  if (check_shape(shape) != LSS_OK) return check_shape(shape);
  const long long HW = (long long)shape->fH * shape->fW;
  return lift_stage_common(d_depth, shape->D * HW, d_feat, shape->C * HW, shape, 0, LSS_F32, d_depth_t, d_feat_t,
                           as_stream(stream));
}

int lss_lift_stage_ex(const void* d_depth_or_logits, int64_t depth_batch_stride, const void* d_feat,
                      int64_t feat_batch_stride, const LssShape* shape, int32_t softmax, int32_t dtype,
                      float* d_depth_t, float* d_feat_t, void* stream) {
  return lift_stage_common(d_depth_or_logits, depth_batch_stride, d_feat, feat_batch_stride, shape, softmax ? 1 : 0,
                           dtype, d_depth_t, d_feat_t, as_stream(stream));
}

int lss_liftsplat_fwd(const float* d_depth_t, const float* d_feat_t, const int32_t* d_sorted_points,
                      const int32_t* d_sorted_cells, const int32_t* d_cell_start, const LssGrid* grid,
                      const LssShape* shape, int32_t layout, float* d_bev, void* stream) {
  LSS_REQUIRE(d_depth_t && d_feat_t, LSS_ERR_NULL_POINTER);
  int rc = check_shape(shape);
  if (rc) return rc;
  LSS_REQUIRE(aligned16(d_feat_t), LSS_ERR_MISALIGNED);
  const int HW = shape->fH * shape->fW;
  return pool_fwd_common(true, d_depth_t, d_feat_t, nullptr, d_sorted_points, d_sorted_cells, d_cell_start,
                         grid, shape->B, shape->C, shape->D, HW, (long long)shape->D * HW, shape_points(shape),
                         layout, d_bev, as_stream(stream));
}

static int liftsplat_bwd_common(const float* d_dbev, const float* d_depth_t, const float* d_feat_t,
                                const int32_t* d_cells, const LssGrid* grid, const LssShape* shape,
                                int32_t layout, int softmax, int out_dtype, void* d_ddepth, long long ddepth_bs,
                                void* d_dfeat, long long dfeat_bs, cudaStream_t st) {
  LSS_REQUIRE(d_dbev && d_depth_t && d_feat_t && d_cells && d_ddepth && d_dfeat, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(valid_dtype(out_dtype), LSS_ERR_UNSUPPORTED);
  int rc = check_shape(shape);
  if (rc) return rc;
  GridDev g;
  rc = make_grid(grid, shape->B, &g);
  if (rc) return rc;
  LSS_REQUIRE(shape->C % 4 == 0 && aligned16(d_dbev) && aligned16(d_feat_t), LSS_ERR_MISALIGNED);
  LSS_REQUIRE(layout == LSS_BEV_NHWC, LSS_ERR_UNSUPPORTED);
  const long long HW = (long long)shape->fH * shape->fW;
  LSS_REQUIRE(ddepth_bs >= shape->D * HW && dfeat_bs >= shape->C * HW, LSS_ERR_BAD_DIMENSION);
  LSS_REQUIRE(!softmax || shape->D <= kBwdChunk, LSS_ERR_UNSUPPORTED);
  PoolBwdArgs a;
  a.dbev = reinterpret_cast<const float4*>(d_dbev); a.depth_t = d_depth_t;
  a.feat_t = reinterpret_cast<const float4*>(d_feat_t); a.cells = d_cells;
  a.ddepth = d_ddepth; a.dfeat = d_dfeat; a.ddepth_bs = ddepth_bs; a.dfeat_bs = dfeat_bs; a.softmax = softmax; a.out_dtype = out_dtype;
  a.D = shape->D; a.fH = shape->fH; a.fW = shape->fW; a.C = shape->C; a.G = shape->C / 4;
  a.n_pix = shape->B * shape->N * shape->fH * shape->fW;
  a.div_fh = FastDiv((uint32_t)shape->fH); a.div_fw = FastDiv((uint32_t)shape->fW);
  LSS_REQUIRE((long long)g.n_cells * a.G < (1ll << 31), LSS_ERR_BAD_DIMENSION);
  const int G = a.G;
  const int blocks = (a.n_pix + kBwdWarps - 1) / kBwdWarps;
  if (G <= 4) return launch_bwd<4>(a, blocks, st);
  if (G <= 8) return launch_bwd<8>(a, blocks, st);
  if (G <= 16) return launch_bwd<16>(a, blocks, st);
  if (G <= 32) return launch_bwd<32>(a, blocks, st);
  return LSS_ERR_UNSUPPORTED;
}

int lss_liftsplat_bwd(const float* d_dbev, const float* d_depth_t, const float* d_feat_t,
                      const int32_t* d_cells, const LssGrid* grid, const LssShape* shape,
                      int32_t layout, float* d_ddepth, float* d_dfeat, void* stream) {
  if (check_shape(shape) != LSS_OK) return check_shape(shape);
  const long long HW = (long long)shape->fH * shape->fW;
  return liftsplat_bwd_common(d_dbev, d_depth_t, d_feat_t, d_cells, grid, shape, layout, 0, LSS_F32, d_ddepth,
                              shape->D * HW, d_dfeat, shape->C * HW, as_stream(stream));
}

int lss_liftsplat_bwd_ex(const float* d_dbev, const float* d_depth_t, const float* d_feat_t,
                         const int32_t* d_cells, const LssGrid* grid, const LssShape* shape,
                         int32_t layout, int32_t softmax, int32_t out_dtype, void* d_ddepth_or_dlogits,
                         int64_t ddepth_batch_stride, void* d_dfeat, int64_t dfeat_batch_stride,
                         void* stream) {
  return liftsplat_bwd_common(d_dbev, d_depth_t, d_feat_t, d_cells, grid, shape, layout, softmax ? 1 : 0,
                              out_dtype, d_ddepth_or_dlogits, ddepth_batch_stride, d_dfeat, dfeat_batch_stride,
                              as_stream(stream));
}

size_t lss_plan_workspace_bytes(const LssShape* shape, const LssGrid* grid) {
  if (check_shape(shape) != LSS_OK) return 0;
  GridDev g;
  if (make_grid(grid, shape->B, &g) != LSS_OK) return 0;
  KeyMap km;
  if (make_keymap(g, &km) != LSS_OK) return 0;
  return make_plan_workspace(shape_points(shape), km.n_keys).total_bytes;
}

int64_t lss_plan_key_count(const LssGrid* grid, int32_t B) {
  GridDev g;
  KeyMap km;
  if (make_grid(grid, B, &g) != LSS_OK || make_keymap(g, &km) != LSS_OK) return 0;
  return km.n_keys;
}

int lss_plan_key_tile(void) { return kKeyTile; }

int lss_build_plan(const float* d_us, const float* d_vs, const float* d_ds, const float* d_rots,
                   const float* d_trans, const float* d_intrins, const float* d_post_rots,
                   const float* d_post_trans, const LssGrid* grid, const LssShape* shape,
                   int32_t* d_cells, int32_t* d_cell_start, int32_t* d_sorted_points,
                   int32_t* d_sorted_cells, int32_t* d_counts, void* d_workspace, size_t workspace_bytes,
                   void* stream) {
  LSS_REQUIRE(d_us && d_vs && d_ds && d_rots && d_trans && d_intrins && d_post_rots && d_post_trans &&
                  d_cells && d_cell_start && d_sorted_points && d_sorted_cells && d_counts && d_workspace,
              LSS_ERR_NULL_POINTER);
  int rc = check_shape(shape);
  if (rc) return rc;
  GridDev g;
  rc = make_grid(grid, shape->B, &g);
  if (rc) return rc;
  GeomArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.us = d_us; ga.vs = d_vs; ga.ds = d_ds;
  ga.post_trans = d_post_trans; ga.trans = d_trans;
  ga.rots = d_rots; ga.intrins = d_intrins; ga.post_rots = d_post_rots;
  ga.raw = 1; ga.N = shape->N; ga.D = shape->D; ga.fH = shape->fH; ga.fW = shape->fW;
  const long long P = shape_points(shape);
  return run_plan(&ga, nullptr, g, P, P / shape->B, d_cells, d_cell_start, d_sorted_points, d_sorted_cells,
                  d_counts, d_workspace, workspace_bytes, as_stream(stream));
}

int lss_build_plan_from_geom(const float* d_geom, const LssGrid* grid, int32_t B, int64_t P,
                             int32_t* d_cells, int32_t* d_cell_start, int32_t* d_sorted_points,
                             int32_t* d_sorted_cells, int32_t* d_counts, void* d_workspace,
                             size_t workspace_bytes, void* stream) {
  LSS_REQUIRE(d_geom && d_cells && d_cell_start && d_sorted_points && d_sorted_cells && d_counts && d_workspace,
              LSS_ERR_NULL_POINTER);
  GridDev g;
  int rc = make_grid(grid, B, &g);
  if (rc) return rc;
  LSS_REQUIRE(P > 0 && P < (1ll << 30) && P % B == 0, LSS_ERR_BAD_DIMENSION);
  return run_plan(nullptr, d_geom, g, P, P / B, d_cells, d_cell_start, d_sorted_points, d_sorted_cells,
                  d_counts, d_workspace, workspace_bytes, as_stream(stream));
}

size_t lss_plan_from_geom_workspace_bytes(int64_t P, const LssGrid* grid, int32_t B) {
  GridDev g;
  KeyMap km;
  if (P <= 0 || P >= (1ll << 30) || make_grid(grid, B, &g) != LSS_OK || make_keymap(g, &km) != LSS_OK) return 0;
  return make_plan_workspace(P, km.n_keys).total_bytes;
}

#ifdef LSS_PHASE_TIMING
int lss_debug_phase_ts(int kernel, unsigned long long* host_out, int n) {
  return (int)cudaMemcpyFromSymbol(host_out, g_phase_ts, (size_t)n * 8, (size_t)kernel * 4096 * 16 * 8);
}
#endif

}  // extern "C"
