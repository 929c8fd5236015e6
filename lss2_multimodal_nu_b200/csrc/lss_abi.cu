// extern "C" entry points declared in include/lss_b200.h.
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
                            cudaStream_t st) {
  IntervalArgs a;
  a.sorted_ranks = sorted_ranks; a.P = P; a.g = g;
  a.div_b = FastDiv(g.B); a.div_z = FastDiv(g.nx[2]); a.div_y = FastDiv(g.nx[1]);
  a.last_mask = last_mask; a.sorted_cells = sorted_cells;
  a.cell_range = reinterpret_cast<int2*>(cell_range); a.counts = counts;
  long long blocks = (P + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  intervals_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  LSS_LAUNCH_CHECK("intervals_kernel");
  return LSS_OK;
}

// ---- the plan: P1 cells -> P2 scan -> P3 scatter + order (lss_plan.cuh) ---------------
static int run_plan(const GeomArgs* ga, const float* dense_geom, const GridDev& g, long long P,
                    long long points_per_sample, int32_t* d_cells, int32_t* d_key_start,
                    int32_t* d_sorted_rec, int32_t* d_counts, void* ws, size_t ws_bytes, cudaStream_t st) {
  KeyMap km;
  int rck = make_keymap(g, &km);
  if (rck) return rck;
  const PlanWorkspace pw = make_plan_workspace(P, km.n_keys);
  LSS_REQUIRE(ws_bytes >= pw.total_bytes, LSS_ERR_WORKSPACE_TOO_SMALL);
  LSS_REQUIRE(aligned16(ws) && aligned16(d_key_start) && aligned16(d_sorted_rec), LSS_ERR_MISALIGNED);
  char* w = static_cast<char*>(ws);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(w + pw.off_cnt);
  uint32_t* tsum = reinterpret_cast<uint32_t*>(w + pw.off_tsum);
  int32_t* keys = reinterpret_cast<int32_t*>(w + pw.off_keys);
  int2* tmp = reinterpret_cast<int2*>(w + pw.off_tmp);

  PlanCellsArgs ca;
  memset(&ca, 0, sizeof(ca));
  ca.grid = g; ca.keys = km; ca.P = P; ca.cells = d_cells; ca.key_of_point = keys; ca.cnt = cnt; ca.tsum = tsum;
  ca.tile_shift = pw.tile_shift; ca.scan_tiles = pw.scan_tiles; ca.counts = d_counts;
  if (ga) {
    ca.geom = *ga;
    const int hw = ga->fH * ga->fW;
    ca.div_w = FastDiv((uint32_t)ga->fW); ca.div_n = FastDiv((uint32_t)ga->N);
    // CTA = as many pixels of one camera as fit 256 threads evenly, kPlanDepths depth bins each
    const int px_blocks = (hw + kPlanThreads - 1) / kPlanThreads;
    int threads = ((hw + px_blocks - 1) / px_blocks + 31) / 32 * 32;
    const long long bn = P / ((long long)ga->D * hw);
    LSS_REQUIRE(bn <= 65535 && (ga->D + kPlanDepths - 1) / kPlanDepths <= 65535, LSS_ERR_BAD_DIMENSION);
    dim3 grid(px_blocks, (ga->D + kPlanDepths - 1) / kPlanDepths, (unsigned)bn);
    LSS_CUDA_TRY(launch_chain(kPdlPlan, plan_cells_kernel, grid, dim3(threads), 0, st, ca), "plan_cells_kernel");
  } else {
    ca.dense_geom = dense_geom;
    ca.div_pps = FastDiv((uint32_t)points_per_sample);
    LSS_CUDA_TRY(launch_chain(kPdlPlan, plan_cells_dense_kernel, dim3((unsigned)((P + kPlanThreads - 1) / kPlanThreads)), dim3(kPlanThreads), 0, st, ca), "plan_cells_dense_kernel");
  }
  LSS_LAUNCH_CHECK("plan_cells_kernel");
#if defined(LSS_DBG_PLAN_UPTO) && LSS_DBG_PLAN_UPTO < 2
  return LSS_OK;
#endif

  PlanScanArgs sa;
  sa.cnt = cnt; sa.tsum = tsum; sa.n = km.n_keys; sa.tiles = pw.scan_tiles; sa.tile_shift = pw.tile_shift;
  sa.key_start = d_key_start; sa.counts = d_counts;
  LSS_CUDA_TRY(launch_chain(kPdlPlan, plan_scan_kernel, dim3((unsigned)pw.scan_tiles), dim3(kPlanThreads), 0, st, sa), "plan_scan_kernel");
  LSS_LAUNCH_CHECK("plan_scan_kernel");
#if defined(LSS_DBG_PLAN_UPTO) && LSS_DBG_PLAN_UPTO < 3
  return LSS_OK;
#endif

  PlanScatterArgs sc;
  sc.key_of_point = keys; sc.cells = d_cells; sc.P = P; sc.cnt = cnt; sc.key_start = d_key_start;
  sc.tmp = tmp; sc.tsum = tsum; sc.scan_tiles = pw.scan_tiles;
  const long long blocks = (P + kPlanThreads - 1) / kPlanThreads;
  LSS_CUDA_TRY(launch_chain(kPdlPlan, plan_scatter_kernel, dim3((unsigned)blocks), dim3(kPlanThreads), 0, st, sc), "plan_scatter_kernel");
  LSS_LAUNCH_CHECK("plan_scatter_kernel");
#if defined(LSS_DBG_PLAN_UPTO) && LSS_DBG_PLAN_UPTO < 4
  return LSS_OK;
#endif

  PlanOrderArgs oa;
  oa.tmp = tmp; oa.key_start = d_key_start; oa.n_keys = km.n_keys; oa.P = P;
  oa.rec = reinterpret_cast<int2*>(d_sorted_rec); oa.keys = km;
  LSS_CUDA_TRY(launch_chain(kPdlPlan, plan_order_kernel, dim3((unsigned)blocks), dim3(kPlanThreads), 0, st, oa), "plan_order_kernel");
  LSS_LAUNCH_CHECK("plan_order_kernel");
  return LSS_OK;
}

// ---- lane layouts of the pooling kernels -----------------------------------------------------
// a feature row of C floats = L lanes x (kNP float4 [+ one float2]); see lss_pool.cuh
struct LaneLayout { int L, np, t2, nact; };
static LaneLayout lane_layout(int C, int min_L) {
  LaneLayout l;
  if (C % 32 == 0 && C / 32 <= 4) { l.L = 8; l.np = C / 32; l.t2 = 0; l.nact = 8; return l; }
  if (C % 32 == 16 && C >= 48 && C <= 112) { l.L = 8; l.np = (C - 16) / 32; l.t2 = 1; l.nact = 8; return l; }
  int L = 1;
  while (L * 4 < C) L <<= 1;
  if (L < min_L) L = min_L;
  l.L = L; l.np = 1; l.t2 = 0; l.nact = C / 4;
  return l;
}

template <bool kFused>
static int launch_pool_fwd(const PoolFwdArgs& a0, int blocks, cudaStream_t st) {
  PoolFwdArgs a = a0;
  const LaneLayout l = lane_layout(a.C, 1);
  a.nact = l.nact;
#define LSS_FWD_CASE(LL, NP, T2)                                                                  \
  if (l.L == LL && l.np == NP && l.t2 == T2) {                                                     \
    LSS_CUDA_TRY(launch_chain(kPdlFwd, pool_fwd_kernel<kFused, LL, NP, (T2 != 0), LSS_FWD_M>, dim3(blocks), dim3(kPoolThreads), 0, st, a), "pool_fwd_kernel"); \
    LSS_LAUNCH_CHECK("pool_fwd_kernel");                                                           \
    return LSS_OK;                                                                                 \
  }
  LSS_FWD_CASE(8, 1, 0) LSS_FWD_CASE(8, 2, 0) LSS_FWD_CASE(8, 3, 0) LSS_FWD_CASE(8, 4, 0)
  LSS_FWD_CASE(8, 1, 1) LSS_FWD_CASE(8, 2, 1) LSS_FWD_CASE(8, 3, 1)
  LSS_FWD_CASE(1, 1, 0) LSS_FWD_CASE(2, 1, 0) LSS_FWD_CASE(4, 1, 0) LSS_FWD_CASE(16, 1, 0) LSS_FWD_CASE(32, 1, 0)
#undef LSS_FWD_CASE
  return LSS_ERR_UNSUPPORTED;
}

#ifndef LSS_BWD_WIDE
#define LSS_BWD_WIDE 0      // 1: C = 64 / 128 use 16-lane walkers (fewer registers, more warps in flight; measured slower)
#endif
#ifndef LSS_BWD_COLS
#define LSS_BWD_COLS 0      // 1: a CTA takes as many adjacent image columns as its warps allow (measured: no L1 gain, 5 % slower)
#endif
#ifndef LSS_BWD_BINS
#define LSS_BWD_BINS 41     // depth bins a warp should at least own (slices of D)
#endif

static int launch_bwd(const PoolBwdArgs& a0, cudaStream_t st) {
  PoolBwdArgs a = a0;
  LaneLayout l = lane_layout(a.C, 8);
  if (LSS_BWD_WIDE && l.t2 == 0 && l.L == 8 && (l.np == 2 || l.np == 4)) { l.L = 16; l.np /= 2; l.nact = 16; }
  a.nact = l.nact;
  const int G = 32 / l.L;
  int rgw = 1;
  while (rgw < kBwdMaxWarps && rgw * G < a.fH) rgw <<= 1;
  a.rg_warps = rgw;
  int slices = kBwdMaxWarps / rgw;
  while (slices > 1 && (a.D + slices - 1) / slices < LSS_BWD_BINS) slices >>= 1;
  a.slices = slices;
  a.d_per_slice = (a.D + slices - 1) / slices;
  a.row_blocks = (a.fH + rgw * G - 1) / (rgw * G);
  // the CTA's remaining warps take adjacent image columns: at one depth their rays land in the same or
  // neighbouring voxels, so the columns' gradient-line gathers meet in L1
  int cols = LSS_BWD_COLS ? kBwdMaxWarps / (rgw * slices) : 1;
  if (cols > a.fW) cols = a.fW;
  if (cols < 1) cols = 1;
  a.cols = cols;
  a.w_blocks = (a.fW + cols - 1) / cols;
  const long long blocks = (long long)a.BN * a.w_blocks * a.row_blocks;
  const int threads = 32 * rgw * slices * cols;
  LSS_REQUIRE(blocks < (1ll << 31), LSS_ERR_BAD_DIMENSION);
  const bool general = a.softmax || a.out_dtype != LSS_F32;
#define LSS_BWD_CASE(LL, NP, T2)                                                                  \
  if (l.L == LL && l.np == NP && l.t2 == T2) {                                                     \
    if (general) LSS_CUDA_TRY(launch_chain(kPdlBwd, liftsplat_bwd_kernel<LL, NP, (T2 != 0), true>, dim3((unsigned)blocks), dim3(threads), 0, st, a), "liftsplat_bwd_kernel");  \
    else LSS_CUDA_TRY(launch_chain(kPdlBwd, liftsplat_bwd_kernel<LL, NP, (T2 != 0), false>, dim3((unsigned)blocks), dim3(threads), 0, st, a), "liftsplat_bwd_kernel");         \
    LSS_LAUNCH_CHECK("liftsplat_bwd_kernel");                                                      \
    return LSS_OK;                                                                                 \
  }
  LSS_BWD_CASE(8, 1, 0) LSS_BWD_CASE(8, 2, 0) LSS_BWD_CASE(8, 3, 0) LSS_BWD_CASE(8, 4, 0)
  LSS_BWD_CASE(8, 1, 1) LSS_BWD_CASE(8, 2, 1) LSS_BWD_CASE(8, 3, 1)
  LSS_BWD_CASE(16, 1, 0) LSS_BWD_CASE(16, 2, 0) LSS_BWD_CASE(32, 1, 0)
#undef LSS_BWD_CASE
  return LSS_ERR_UNSUPPORTED;
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
                          as_stream(stream));
}

static int fill_cta_count(const KeyMap& km) {
  // fill CTAs: one per SM, so the zero stream runs in the background for the whole kernel
  long long fill = sm_count();
  const long long fill_blocks = ((long long)km.n_keys + 31) / 32;
  if (fill * kPoolWarps > fill_blocks) fill = (fill_blocks + kPoolWarps - 1) / kPoolWarps;
  if (fill < 1) fill = 1;
  return (int)fill;
}

static int pool_fwd_common(bool fused, const void* d_depth, long long depth_bs, int depth_dtype,
                           const float* d_feat_t, const float* d_x, const int32_t* d_sorted_rec,
                           const int32_t* d_key_start, const LssGrid* grid, int32_t B, int32_t C, int32_t HW,
                           long long dhw, long long P, float* d_bev, cudaStream_t st) {
  LSS_REQUIRE(d_sorted_rec && d_key_start && d_bev, LSS_ERR_NULL_POINTER);
  GridDev g;
  int rc = make_grid(grid, B, &g);
  if (rc) return rc;
  LSS_REQUIRE(C > 0 && C % 4 == 0, LSS_ERR_MISALIGNED);
  LSS_REQUIRE(C <= 128, LSS_ERR_UNSUPPORTED);
  LSS_REQUIRE(aligned16(d_bev) && aligned16(d_sorted_rec), LSS_ERR_MISALIGNED);
  LSS_REQUIRE(P > 0 && P < (1ll << 30), LSS_ERR_BAD_DIMENSION);
  PoolFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.depth = d_depth; a.depth_bs = depth_bs; a.depth_dtype = depth_dtype; a.feat_t = d_feat_t; a.x = d_x;
  a.rec = reinterpret_cast<const int2*>(d_sorted_rec); a.key_start = d_key_start;
  a.bev = d_bev; a.P = P;
  rc = make_keymap(g, &a.keys);
  if (rc) return rc;
  a.C = C; a.HW = HW;
  a.div_dhw = FastDiv((uint32_t)(dhw > 0 ? dhw : 1)); a.div_hw = FastDiv((uint32_t)(HW > 0 ? HW : 1));
  a.fill_ctas = fill_cta_count(a.keys);
  const long long per_cta = (long long)LSS_FWD_M * 32 * kPoolWarps;
  const long long blocks = a.fill_ctas + (P + per_cta - 1) / per_cta;
  LSS_REQUIRE(blocks < (1ll << 31), LSS_ERR_BAD_DIMENSION);
  return fused ? launch_pool_fwd<true>(a, (int)blocks, st) : launch_pool_fwd<false>(a, (int)blocks, st);
}

int lss_pool_dense_fwd(const float* d_x, const int32_t* d_sorted_rec, const int32_t* d_key_start,
                       const LssGrid* grid, int32_t B, int32_t C, int64_t P, float* d_bev, void* stream) {
  LSS_REQUIRE(d_x, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(aligned16(d_x), LSS_ERR_MISALIGNED);
  LSS_REQUIRE(C > 0 && (long long)P * (C / 4) < (1ll << 31), LSS_ERR_BAD_DIMENSION);   // 16-byte row offsets in 31 bits
  return pool_fwd_common(false, nullptr, 0, LSS_F32, nullptr, d_x, d_sorted_rec, d_key_start, grid, B, C, 1, 1, P,
                         d_bev, as_stream(stream));
}

int lss_pool_dense_bwd(const float* d_dbev, const int32_t* d_cells, const LssGrid* grid,
                       int32_t B, int32_t C, int64_t P, float* d_dx, void* stream) {
  LSS_REQUIRE(d_dbev && d_cells && d_dx, LSS_ERR_NULL_POINTER);
  GridDev g;
  int rc = make_grid(grid, B, &g);
  if (rc) return rc;
  LSS_REQUIRE(C > 0 && C % 4 == 0 && aligned16(d_dbev) && aligned16(d_dx), LSS_ERR_MISALIGNED);
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

int lss_feat_stage(const void* d_feat, int64_t feat_batch_stride, const LssShape* shape, int32_t dtype,
                   float* d_feat_t, void* stream) {
  LSS_REQUIRE(d_feat && d_feat_t, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(valid_dtype(dtype), LSS_ERR_UNSUPPORTED);
  int rc = check_shape(shape);
  if (rc) return rc;
  const int HW = shape->fH * shape->fW, BN = shape->B * shape->N, C = shape->C;
  LSS_REQUIRE(C % 4 == 0 && aligned16(d_feat_t), LSS_ERR_MISALIGNED);
  LSS_REQUIRE(BN <= 65535, LSS_ERR_BAD_DIMENSION);
  LSS_REQUIRE(feat_batch_stride >= (long long)C * HW, LSS_ERR_BAD_DIMENSION);
  const int vec_ok = (dtype == LSS_F32 && HW % 4 == 0 && feat_batch_stride % 4 == 0 && aligned16(d_feat)) ? 1 : 0;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, BN);
  LSS_CUDA_TRY(launch_chain(kPdlStage, feat_stage_kernel, grid, dim3(256), 0, as_stream(stream), d_feat, (long long)feat_batch_stride, (int)dtype, vec_ok, C, HW, d_feat_t), "feat_stage_kernel");
  LSS_LAUNCH_CHECK("feat_stage_kernel");
  return LSS_OK;
}

int lss_depth_softmax(const void* d_logits, int64_t logits_batch_stride, const LssShape* shape, int32_t dtype,
                      float* d_depth, void* stream) {
  LSS_REQUIRE(d_logits && d_depth, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(valid_dtype(dtype), LSS_ERR_UNSUPPORTED);
  int rc = check_shape(shape);
  if (rc) return rc;
  const int HW = shape->fH * shape->fW, BN = shape->B * shape->N;
  LSS_REQUIRE(logits_batch_stride >= (long long)shape->D * HW, LSS_ERR_BAD_DIMENSION);
  LSS_REQUIRE((long long)BN * HW < (1ll << 31), LSS_ERR_BAD_DIMENSION);
  const int n = BN * HW;
  depth_softmax_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(d_logits, logits_batch_stride, dtype,
                                                                       shape->D, HW, BN, d_depth);
  LSS_LAUNCH_CHECK("depth_softmax_kernel");
  return LSS_OK;
}

int lss_liftsplat_fwd(const void* d_depth, int64_t depth_batch_stride, int32_t depth_dtype,
                      const float* d_feat_t, const int32_t* d_sorted_rec, const int32_t* d_key_start,
                      const LssGrid* grid, const LssShape* shape, float* d_bev, void* stream) {
  LSS_REQUIRE(d_depth && d_feat_t, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(valid_dtype(depth_dtype), LSS_ERR_UNSUPPORTED);
  int rc = check_shape(shape);
  if (rc) return rc;
  LSS_REQUIRE(aligned16(d_feat_t), LSS_ERR_MISALIGNED);
  const int HW = shape->fH * shape->fW;
  LSS_REQUIRE(depth_batch_stride >= (long long)shape->D * HW, LSS_ERR_BAD_DIMENSION);
  LSS_REQUIRE((long long)shape->B * shape->N * HW * (shape->C / 4) < (1ll << 31), LSS_ERR_BAD_DIMENSION);
  return pool_fwd_common(true, d_depth, depth_batch_stride, depth_dtype, d_feat_t, nullptr, d_sorted_rec,
                         d_key_start, grid, shape->B, shape->C, HW, (long long)shape->D * HW, shape_points(shape),
                         d_bev, as_stream(stream));
}

int lss_liftsplat_bwd(const float* d_dbev, const void* d_depth, int64_t depth_batch_stride, int32_t depth_dtype,
                      const float* d_feat_t, const int32_t* d_cells, const LssGrid* grid, const LssShape* shape,
                      int32_t softmax, int32_t out_dtype, void* d_ddepth_or_dlogits, int64_t ddepth_batch_stride,
                      void* d_dfeat, int64_t dfeat_batch_stride, void* stream) {
  LSS_REQUIRE(d_dbev && d_depth && d_feat_t && d_cells && d_ddepth_or_dlogits && d_dfeat, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(valid_dtype(out_dtype) && valid_dtype(depth_dtype), LSS_ERR_UNSUPPORTED);
  int rc = check_shape(shape);
  if (rc) return rc;
  GridDev g;
  rc = make_grid(grid, shape->B, &g);
  if (rc) return rc;
  LSS_REQUIRE(shape->C % 4 == 0 && aligned16(d_dbev) && aligned16(d_feat_t), LSS_ERR_MISALIGNED);
  LSS_REQUIRE(shape->C <= 128, LSS_ERR_UNSUPPORTED);
  const long long HW = (long long)shape->fH * shape->fW;
  LSS_REQUIRE(ddepth_batch_stride >= shape->D * HW && dfeat_batch_stride >= shape->C * HW &&
                  depth_batch_stride >= shape->D * HW, LSS_ERR_BAD_DIMENSION);
  LSS_REQUIRE(!softmax || (shape->D <= kBwdMaxD && depth_dtype == LSS_F32), LSS_ERR_UNSUPPORTED);
  PoolBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.dbev = d_dbev; a.depth = d_depth; a.depth_bs = depth_batch_stride; a.depth_dtype = depth_dtype;
  a.feat_t = d_feat_t; a.cells = d_cells;
  a.ddepth = d_ddepth_or_dlogits; a.dfeat = d_dfeat; a.ddepth_bs = ddepth_batch_stride; a.dfeat_bs = dfeat_batch_stride;
  a.softmax = softmax ? 1 : 0; a.out_dtype = out_dtype;
  a.D = shape->D; a.fH = shape->fH; a.fW = shape->fW; a.C = shape->C; a.BN = shape->B * shape->N;
  return launch_bwd(a, as_stream(stream));
}

size_t lss_plan_workspace_bytes(const LssShape* shape, const LssGrid* grid) {
  if (check_shape(shape) != LSS_OK) return 0;
  GridDev g;
  if (make_grid(grid, shape->B, &g) != LSS_OK) return 0;
  KeyMap km;
  if (make_keymap(g, &km) != LSS_OK) return 0;
  return make_plan_workspace(shape_points(shape), km.n_keys).total_bytes;
}

size_t lss_plan_workspace_control_bytes(const LssShape* shape, const LssGrid* grid) {
  if (check_shape(shape) != LSS_OK) return 0;
  GridDev g;
  if (make_grid(grid, shape->B, &g) != LSS_OK) return 0;
  KeyMap km;
  if (make_keymap(g, &km) != LSS_OK) return 0;
  return make_plan_workspace(shape_points(shape), km.n_keys).control_bytes;
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
                   int32_t* d_cells, int32_t* d_key_start, int32_t* d_sorted_rec, int32_t* d_counts,
                   void* d_workspace, size_t workspace_bytes, void* stream) {
  LSS_REQUIRE(d_us && d_vs && d_ds && d_rots && d_trans && d_intrins && d_post_rots && d_post_trans &&
                  d_cells && d_key_start && d_sorted_rec && d_counts && d_workspace,
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
  return run_plan(&ga, nullptr, g, P, P / shape->B, d_cells, d_key_start, d_sorted_rec, d_counts, d_workspace,
                  workspace_bytes, as_stream(stream));
}

int lss_build_plan_from_geom(const float* d_geom, const LssGrid* grid, int32_t B, int64_t P,
                             int32_t* d_cells, int32_t* d_key_start, int32_t* d_sorted_rec, int32_t* d_counts,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
  LSS_REQUIRE(d_geom && d_cells && d_key_start && d_sorted_rec && d_counts && d_workspace, LSS_ERR_NULL_POINTER);
  GridDev g;
  int rc = make_grid(grid, B, &g);
  if (rc) return rc;
  LSS_REQUIRE(P > 0 && P < (1ll << 30) && P % B == 0, LSS_ERR_BAD_DIMENSION);
  return run_plan(nullptr, d_geom, g, P, P / B, d_cells, d_key_start, d_sorted_rec, d_counts, d_workspace,
                  workspace_bytes, as_stream(stream));
}

size_t lss_plan_from_geom_workspace_bytes(int64_t P, const LssGrid* grid, int32_t B) {
  GridDev g;
  KeyMap km;
  if (P <= 0 || P >= (1ll << 30) || make_grid(grid, B, &g) != LSS_OK || make_keymap(g, &km) != LSS_OK) return 0;
  return make_plan_workspace(P, km.n_keys).total_bytes;
}

}  // extern "C"
