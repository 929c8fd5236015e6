// extern "C" entry points declared in include/lss_b200.h.
#include <stdlib.h>

#include "lss_common.cuh"
#include "lss_geometry.cuh"
#include "lss_pool.cuh"
#include "lss_sort.cuh"
#include "lss_sort_small.cuh"
#include "lss_partition.cuh"

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

// ---- co-resident partition path (lss_partition.cuh) -------------------------------------
struct CoopPlan {
  bool ok;
  int tiles, hi_bits, lo_bits, n_buckets;
  size_t off_keys, off_vals, off_rows, off_excl, off_totals, off_bucket_start, off_barrier, total_bytes;
};

static int coop_capacity() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, partition_coop_kernel, kPartThreads, 0) != cudaSuccess)
      per_sm = 0;
    cached[dev] = per_sm * sm_count() + 1;  // +1: 0 means "not queried"
  }
  return cached[dev] - 1;
}

static CoopPlan make_coop_plan(long long P, int32_t n_cells, int ppc, bool query_device) {
  CoopPlan c;
  memset(&c, 0, sizeof(c));
  int kb = 0;
  while ((1ll << kb) <= (long long)n_cells) ++kb;
  if (kb < 2) kb = 2;
  c.hi_bits = kb - 1 < kPartMaxBits ? kb - 1 : kPartMaxBits;
  c.lo_bits = kb - c.hi_bits;
  c.n_buckets = 1 << c.hi_bits;
  const long long tiles = (P + kPartTile - 1) / kPartTile;
  c.tiles = (int)tiles;
  c.ok = c.lo_bits <= kLocalMaxBits && tiles <= kPartMaxTiles && ppc > 0 &&
         (kPartTile - 1) / ppc + 2 <= kPartMaxCams;
  if (c.ok && query_device) c.ok = tiles <= coop_capacity();
  size_t off = 0;
  c.off_keys = off; off += align_up((size_t)P * 4, 256);
  c.off_vals = off; off += align_up((size_t)P * 4, 256);
  c.off_rows = off; off += align_up((size_t)tiles * kPartMaxBins * 2, 256);
  c.off_excl = off; off += align_up((size_t)tiles * kPartMaxBins * 4, 256);
  c.off_totals = off; off += align_up((size_t)kPartMaxBins * 4, 256);
  c.off_bucket_start = off; off += align_up((size_t)(kPartMaxBins + 1) * 4, 256);
  c.off_barrier = off; off += 256;
  c.total_bytes = off;
  return c;
}

static int run_coop_plan(const CoopPlan& c, const GeomArgs& ga, const GridDev& g, long long P,
                         int32_t* d_cells, int32_t* sorted_points, int32_t* sorted_cells,
                         int32_t* cell_range, int32_t* counts, void* ws, cudaStream_t st) {
  char* w = static_cast<char*>(ws);
  PartitionArgs a;
  memset(&a, 0, sizeof(a));
  a.geom = ga; a.grid = g;
  const int hw = ga.fH * ga.fW;
  a.div_ppc = FastDiv((uint32_t)(ga.D * hw)); a.div_hw = FastDiv((uint32_t)hw);
  a.div_w = FastDiv((uint32_t)ga.fW); a.div_n = FastDiv((uint32_t)ga.N);
  a.P = P; a.tiles = c.tiles; a.shift = c.lo_bits; a.bits = c.hi_bits;
  a.cells = d_cells;
  a.part_keys = reinterpret_cast<int32_t*>(w + c.off_keys);
  a.part_vals = reinterpret_cast<int32_t*>(w + c.off_vals);
  a.rows = reinterpret_cast<uint16_t*>(w + c.off_rows);
  a.excl = reinterpret_cast<uint32_t*>(w + c.off_excl);
  a.totals = reinterpret_cast<uint32_t*>(w + c.off_totals);
  a.bucket_start = reinterpret_cast<uint32_t*>(w + c.off_bucket_start);
  a.counts = counts;
  a.barrier = reinterpret_cast<uint32_t*>(w + c.off_barrier);
  partition_coop_kernel<<<c.tiles, kPartThreads, 0, st>>>(a);
  LSS_LAUNCH_CHECK("partition_coop_kernel");
  LocalArgs l;
  memset(&l, 0, sizeof(l));
  l.keys = a.part_keys; l.vals = a.part_vals; l.bucket_start = a.bucket_start;
  l.sorted_points = sorted_points; l.sorted_ranks = nullptr; l.sorted_cells = sorted_cells;
  l.cell_range = reinterpret_cast<int2*>(cell_range); l.counts = counts;
  l.g = g;
  l.div_b = FastDiv(g.B); l.div_z = FastDiv(g.nx[2]); l.div_y = FastDiv(g.nx[1]);
  l.lo_bits = c.lo_bits;
  return launch_local_sort(l, c.n_buckets, st);
}

template <int kLanes>
static int launch_bwd(const PoolBwdArgs& a, int blocks, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    LSS_CUDA_TRY(cudaFuncSetAttribute(liftsplat_bwd_nhwc_kernel<kLanes>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(bwd smem)");
  }
  liftsplat_bwd_nhwc_kernel<kLanes><<<blocks, 256, smem, st>>>(a);
  LSS_LAUNCH_CHECK("liftsplat_bwd_nhwc_kernel");
  return LSS_OK;
}
template <bool kFused>
static int launch_pool_fwd(const PoolFwdArgs& a, int blocks, cudaStream_t st) {
  const int G = a.G;
  if (G <= 4) pool_fwd_nhwc_kernel<kFused, 4><<<blocks, kPoolThreads, 0, st>>>(a);
  else if (G <= 8) pool_fwd_nhwc_kernel<kFused, 8><<<blocks, kPoolThreads, 0, st>>>(a);
  else if (G <= 16) pool_fwd_nhwc_kernel<kFused, 16><<<blocks, kPoolThreads, 0, st>>>(a);
  else pool_fwd_nhwc_kernel<kFused, 32><<<blocks, kPoolThreads, 0, st>>>(a);
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
  if (s.small) {
    LSS_CUDA_TRY(cudaMemsetAsync(w + s.off_flags, 0, s.total_bytes - s.off_flags, st), "memset sort flags");
    return run_sort_passes_small(s, d_ranks, d_sorted_ranks, d_sorted_points, P, d_workspace, nullptr, st);
  }
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
                           const float* d_x, const int32_t* d_sorted_points,
                           const int32_t* d_sorted_cells, const int32_t* d_cell_range,
                           const int32_t* d_counts, const LssGrid* grid, int32_t B, int32_t C,
                           int32_t D, int32_t HW, long long dhw, int32_t layout, float* d_bev,
                           cudaStream_t st) {
  LSS_REQUIRE(d_sorted_points && d_sorted_cells && d_cell_range && d_counts && d_bev, LSS_ERR_NULL_POINTER);
  GridDev g;
  int rc = make_grid(grid, B, &g);
  if (rc) return rc;
  LSS_REQUIRE(C > 0 && C % 4 == 0, LSS_ERR_MISALIGNED);
  LSS_REQUIRE(C <= 128, LSS_ERR_UNSUPPORTED);
  LSS_REQUIRE(aligned16(d_bev), LSS_ERR_MISALIGNED);
  LSS_REQUIRE(layout == LSS_BEV_NHWC, LSS_ERR_UNSUPPORTED);
  PoolFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.depth_t = d_depth_t; a.feat_t = reinterpret_cast<const float4*>(d_feat_t);
  a.x = reinterpret_cast<const float4*>(d_x);
  a.sorted_points = d_sorted_points; a.sorted_cells = d_sorted_cells; a.counts = d_counts;
  a.cell_range = reinterpret_cast<const int2*>(d_cell_range);
  a.bev = reinterpret_cast<float4*>(d_bev);
  a.n_cells = (uint32_t)g.n_cells;
  a.G = C / 4; a.D = D; a.HW = HW;
  LSS_REQUIRE((long long)g.n_cells * a.G < (1ll << 31), LSS_ERR_BAD_DIMENSION);
  a.fill_warps = 1;
  int blocks_per_sm = 12;
  if (const char* e = getenv("LSS_FILL_WARPS")) a.fill_warps = atoi(e);          // tuning knobs
  if (const char* e = getenv("LSS_POOL_BLOCKS_PER_SM")) blocks_per_sm = atoi(e);
  if (a.fill_warps < 1) a.fill_warps = 1;
  if (a.fill_warps > kPoolWarps - 1) a.fill_warps = kPoolWarps - 1;
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  a.div_g = FastDiv(a.G);
  a.div_dhw = FastDiv((uint32_t)(dhw > 0 ? dhw : 1)); a.div_hw = FastDiv((uint32_t)(HW > 0 ? HW : 1));
  const int blocks = sm_count() * blocks_per_sm;
  return fused ? launch_pool_fwd<true>(a, blocks, st) : launch_pool_fwd<false>(a, blocks, st);
}

int lss_pool_dense_fwd(const float* d_x, const int32_t* d_sorted_points,
                       const int32_t* d_sorted_cells, const int32_t* d_cell_range,
                       const int32_t* d_counts, const LssGrid* grid, int32_t B, int32_t C,
                       int32_t layout, float* d_bev, void* stream) {
  LSS_REQUIRE(d_x, LSS_ERR_NULL_POINTER);
  LSS_REQUIRE(aligned16(d_x), LSS_ERR_MISALIGNED);
  return pool_fwd_common(false, nullptr, nullptr, d_x, d_sorted_points, d_sorted_cells, d_cell_range,
                         d_counts, grid, B, C, 1, 1, 1, layout, d_bev, as_stream(stream));
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

int lss_lift_stage(const float* d_depth, const float* d_feat, const LssShape* shape,
                   float* d_depth_t, float* d_feat_t, void* stream) {
  LSS_REQUIRE(d_depth && d_feat && d_depth_t && d_feat_t, LSS_ERR_NULL_POINTER);
  int rc = check_shape(shape);
  if (rc) return rc;
  const int HW = shape->fH * shape->fW, BN = shape->B * shape->N;
  const int R = shape->D > shape->C ? shape->D : shape->C;
  LSS_REQUIRE(BN * 2 <= 65535, LSS_ERR_BAD_DIMENSION);
  dim3 grid((HW + 31) / 32, (R + 31) / 32, BN * 2);
  lift_stage_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_depth, d_feat, shape->D, shape->C, HW,
                                                        d_depth_t, d_feat_t);
  LSS_LAUNCH_CHECK("lift_stage_kernel");
  return LSS_OK;
}

int lss_liftsplat_fwd(const float* d_depth_t, const float* d_feat_t,
                      const int32_t* d_sorted_points, const int32_t* d_sorted_cells,
                      const int32_t* d_cell_range, const int32_t* d_counts, const LssGrid* grid,
                      const LssShape* shape, int32_t layout, float* d_bev, void* stream) {
  LSS_REQUIRE(d_depth_t && d_feat_t, LSS_ERR_NULL_POINTER);
  int rc = check_shape(shape);
  if (rc) return rc;
  LSS_REQUIRE(aligned16(d_feat_t), LSS_ERR_MISALIGNED);
  const int HW = shape->fH * shape->fW;
  return pool_fwd_common(true, d_depth_t, d_feat_t, nullptr, d_sorted_points, d_sorted_cells,
                         d_cell_range, d_counts, grid, shape->B, shape->C, shape->D, HW,
                         (long long)shape->D * HW, layout, d_bev, as_stream(stream));
}

int lss_liftsplat_bwd(const float* d_dbev, const float* d_depth_t, const float* d_feat_t,
                      const int32_t* d_cells, const LssGrid* grid, const LssShape* shape,
                      int32_t layout, float* d_ddepth, float* d_dfeat, void* stream) {
  LSS_REQUIRE(d_dbev && d_depth_t && d_feat_t && d_cells && d_ddepth && d_dfeat, LSS_ERR_NULL_POINTER);
  int rc = check_shape(shape);
  if (rc) return rc;
  GridDev g;
  rc = make_grid(grid, shape->B, &g);
  if (rc) return rc;
  LSS_REQUIRE(shape->C % 4 == 0 && aligned16(d_dbev) && aligned16(d_feat_t), LSS_ERR_MISALIGNED);
  LSS_REQUIRE(layout == LSS_BEV_NHWC, LSS_ERR_UNSUPPORTED);
  PoolBwdArgs a;
  a.dbev = reinterpret_cast<const float4*>(d_dbev); a.depth_t = d_depth_t;
  a.feat_t = reinterpret_cast<const float4*>(d_feat_t); a.cells = d_cells;
  a.ddepth = d_ddepth; a.dfeat = d_dfeat;
  a.D = shape->D; a.fH = shape->fH; a.fW = shape->fW; a.C = shape->C; a.G = shape->C / 4;
  const int G = a.G;
  const int lanes = G <= 4 ? 4 : G <= 8 ? 8 : G <= 16 ? 16 : 32;
  const int round = (32 / lanes) * (lanes >= 8 ? 8 : lanes);   // depth bins per kernel round
  const int Dpad = (a.D + round - 1) / round * round;
  const size_t smem = ((size_t)3 * Dpad * a.fW + (size_t)a.C * (a.fW + 1)) * 4;
  LSS_REQUIRE(smem <= 200 * 1024, LSS_ERR_UNSUPPORTED);
  const int blocks = shape->B * shape->N * shape->fH;
  cudaStream_t st = as_stream(stream);
  if (G <= 4) return launch_bwd<4>(a, blocks, smem, st);
  if (G <= 8) return launch_bwd<8>(a, blocks, smem, st);
  if (G <= 16) return launch_bwd<16>(a, blocks, smem, st);
  if (G <= 32) return launch_bwd<32>(a, blocks, smem, st);
  return LSS_ERR_UNSUPPORTED;
}

size_t lss_plan_workspace_bytes(const LssShape* shape, const LssGrid* grid) {
  if (check_shape(shape) != LSS_OK) return 0;
  GridDev g;
  if (make_grid(grid, shape->B, &g) != LSS_OK) return 0;
  const long long P = shape_points(shape);
  const SortPlan s = make_sort_plan(P, g.n_cells);
  const size_t general = make_msd_plan(s).total_bytes + 2 * align_up((size_t)P * 4, 256);
  const size_t coop = make_coop_plan(P, g.n_cells, shape->D * shape->fH * shape->fW, false).total_bytes;
  return general > coop ? general : coop;
}

int lss_build_plan(const float* d_us, const float* d_vs, const float* d_ds, const float* d_rots,
                   const float* d_trans, const float* d_intrins, const float* d_post_rots,
                   const float* d_post_trans, const LssGrid* grid, const LssShape* shape,
                   int32_t* d_cells, int32_t* d_sorted_points, int32_t* d_sorted_cells,
                   int32_t* d_cell_range, int32_t* d_counts, void* d_workspace,
                   size_t workspace_bytes, void* stream) {
  LSS_REQUIRE(d_us && d_vs && d_ds && d_rots && d_trans && d_intrins && d_post_rots &&
                  d_post_trans && d_cells && d_sorted_points && d_sorted_cells && d_cell_range && d_counts &&
                  d_workspace, LSS_ERR_NULL_POINTER);
  int rc = check_shape(shape);
  if (rc) return rc;
  GridDev g;
  rc = make_grid(grid, shape->B, &g);
  if (rc) return rc;
  LSS_REQUIRE(aligned16(d_workspace), LSS_ERR_MISALIGNED);
  const long long P = shape_points(shape);
  const SortPlan s = make_sort_plan(P, g.n_cells);
  const size_t ws_sort = make_msd_plan(s).total_bytes;
  const CoopPlan coop = make_coop_plan(P, g.n_cells, shape->D * shape->fH * shape->fW, true);
  size_t need = ws_sort + 2 * align_up((size_t)P * 4, 256);
  if (coop.total_bytes > need) need = coop.total_bytes;
  LSS_REQUIRE(workspace_bytes >= need, LSS_ERR_WORKSPACE_TOO_SMALL);
  cudaStream_t st = as_stream(stream);
  char* w = static_cast<char*>(d_workspace);
  int32_t* ranks = reinterpret_cast<int32_t*>(w + ws_sort);
  int32_t* sorted_ranks = reinterpret_cast<int32_t*>(w + ws_sort + align_up((size_t)P * 4, 256));

  GeomArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.us = d_us; ga.vs = d_vs; ga.ds = d_ds;
  ga.post_trans = d_post_trans; ga.trans = d_trans;
  ga.rots = d_rots; ga.intrins = d_intrins; ga.post_rots = d_post_rots;
  ga.raw = 1; ga.N = shape->N; ga.D = shape->D; ga.fH = shape->fH; ga.fW = shape->fW;
  const int ppc = shape->D * shape->fH * shape->fW;
  if (coop.ok)  // all tiles co-resident: K0 + K1' + partition in one kernel, then local sort + intervals
    return run_coop_plan(coop, ga, g, P, d_cells, d_sorted_points, d_sorted_cells, d_cell_range, d_counts,
                         d_workspace, st);
  const MsdPlan msd = make_msd_plan(s);
  if (msd.ok && (kSortTile - 1) / ppc + 2 <= kSmallMaxCams) {
    // single wave: P1 = K0 + K1' + MSD partition, P2 = local sort + intervals (K2 + K3)
    SmallGeom sg;
    sg.enabled = true; sg.geom = ga; sg.grid = g; sg.cells = d_cells;
    return run_msd_plan(s, msd, sg, d_sorted_points, d_sorted_cells, d_cell_range, d_counts, P, d_workspace, st);
  }
  LSS_CUDA_TRY(cudaMemsetAsync(d_cell_range, 0, (size_t)g.n_cells * 8, st), "memset cell_range");
  LSS_CUDA_TRY(cudaMemsetAsync(d_counts, 0, 8, st), "memset counts");
  PointOut out{nullptr, nullptr, ranks, d_cells};
  if (s.small) {
    rc = launch_geometry(ga, g, shape, out, nullptr, st);
    if (rc) return rc;
    rc = run_sort_passes_small(s, ranks, sorted_ranks, d_sorted_points, P, d_workspace, nullptr, st);
    if (rc) return rc;
    return launch_intervals(sorted_ranks, P, g, nullptr, d_sorted_cells, d_cell_range, d_counts, nullptr, 0, st);
  }
  SortDigits sd = sort_digits(s, d_workspace);
  rc = launch_geometry(ga, g, shape, out, &sd, st);
  if (rc) return rc;
  rc = run_sort_passes(s, ranks, sorted_ranks, d_sorted_points, P, d_workspace, st);
  if (rc) return rc;
  // K3 also wipes the sort's control words so the workspace is ready for the next call
  return launch_intervals(sorted_ranks, P, g, nullptr, d_sorted_cells, d_cell_range, d_counts,
                          reinterpret_cast<uint32_t*>(w + s.off_control),
                          (long long)(s.control_bytes / 4), st);
}

#ifdef LSS_PHASE_TIMING
int lss_debug_phase_ts(int kernel, unsigned long long* host_out, int n) {
  return (int)cudaMemcpyFromSymbol(host_out, g_phase_ts, (size_t)n * 8, (size_t)kernel * 4096 * 8 * 8);
}
#endif

}  // extern "C"
