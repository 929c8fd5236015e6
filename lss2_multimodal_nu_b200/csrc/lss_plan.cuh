// The plan: everything between the calibration tensors and the per-cell point
// lists the pooling kernels consume, i.e. the geometry half of get_voxels
// (reference src/model_baseline.py:128-131) plus the argsort of :110 and the
// interval detection of src/tools.py:196-197.
//
// The voxel rank has few distinct values (n_cells = B*X*Y*Z, 320 k at the
// headline config, about one point per cell), so the sort is a ONE-DIGIT radix
// sort whose digit is the whole key -- a counting sort -- in four small kernels:
//   P1 cells    camera preparation (K0) + frustum geometry (K1') per point ->
//               output cell and key; one RED.ADD per kept point builds the key
//               histogram, one warp-aggregated RED.ADD per warp and scan tile the
//               tile totals
//   P2 scan     exclusive prefix over the keys.  A tile's base is the sum of the
//               tile totals before it (<= 1024 values, one coalesced read), so no
//               tile waits for another: no look-back chain, no tickets.
//               key_start[k] .. key_start[k+1] is key k's run in the sorted
//               order: the interval table (K3) IS the scanned histogram
//   P3 scatter  every kept point takes a slot of its key's run (atomic cursor =
//               the histogram counted back to zero) and parks {cell, id} there
//   P4 order    the slots of a run are put in ascending point order, which is the
//               order a STABLE sort gives (argsort on torch's radix path): a
//               slot finds its run by looking at its neighbours (equal cell) and
//               counts the smaller ids -- no key arithmetic, no table lookups.
//               Runs that do not fit the neighbour window are ordered by (part of)
//               the warp that holds their first slot.
// The key is the OUTPUT CELL in tile-major order (KeyMap, lss_common.cuh): the
// digits (b, x, y, z) of the reference's rank x*(Y*Z*B) + y*(Z*B) + z*B + b
// regrouped as (b, x/8, y/8, x%8, y%8, z) -- a bijection of the rank, so runs,
// per-run order and therefore every per-voxel sum are the reference's; only the
// order in which the runs follow one another differs, which is what makes
// neighbours in the sorted list neighbours on the map (and in L1).
// lss_sort_ranks (K2) remains the bit-exact argsort of the reference's rank.
#pragma once

#include "lss_common.cuh"
#include "lss_geometry.cuh"
#include "lss_sort.cuh"   // align_up

namespace lss {

constexpr int kPlanThreads = 256;
constexpr int kScanBlock = kPlanThreads * 4;           // keys per scan step (one uint4 per thread)
constexpr int kScanMaxTiles = 1024;                    // tile totals a scan CTA sums for its base
constexpr int kTsumCopies = 8;                         // replicas of the tile totals (spreads the atomics of P1)
constexpr int kOrderWin = 4;                           // neighbours a slot loads up front on either side to find its run
constexpr int kRunSerial = 128;                        // runs up to this: every slot counts the smaller ids of its run
constexpr int kRunWarp = 1024;                         // longer runs up to this: sorted by a warp in shared memory

struct PlanWorkspace {
  size_t off_cnt, off_tsum, off_ctl, off_keys, off_tmp, control_bytes, total_bytes;
  int tile_shift;   // a scan tile holds 1 << tile_shift keys
  int scan_tiles;
};

inline PlanWorkspace make_plan_workspace(long long P, int32_t n_keys) {
  PlanWorkspace w;
  int s = 10;
  while ((((long long)n_keys + (1ll << s) - 1) >> s) > kScanMaxTiles) ++s;
  w.tile_shift = s;
  w.scan_tiles = (int)(((long long)n_keys + (1ll << s) - 1) >> s);
  size_t off = 0;
  // control part: zero on entry, left zero by a successful call
  w.off_cnt = off; off += align_up((size_t)n_keys * 4, 256);
  w.off_tsum = off; off += align_up((size_t)w.scan_tiles * kTsumCopies * 4, 256);
  w.off_ctl = off; off += 256;                        // spare control words
  w.control_bytes = off;
  w.off_keys = off; off += align_up((size_t)P * 4, 256);
  w.off_tmp = off; off += align_up((size_t)P * 8, 256);   // {cell, id} in arrival order
  w.total_bytes = off;
  return w;
}

// ---------------------------------------------------------------------------
// P1: per point output cell (-1: dropped) and sort key, histogram of the keys,
// totals per scan tile.  Two variants: from the calibration (plan_cells_kernel)
// and from a materialised geometry tensor, the literal voxel_pooling(geom_feats, x)
// call (plan_cells_dense_kernel).
// ---------------------------------------------------------------------------
struct PlanCellsArgs {
  GeomArgs geom;            // raw calibration + frustum axes (fused variant)
  const float* dense_geom;  // (P,3) (dense variant)
  GridDev grid;
  KeyMap keys;
  FastDiv div_w, div_n, div_pps;
  long long P;
  int32_t* cells;           // (P)
  int32_t* key_of_point;    // (P) workspace
  uint32_t* cnt;            // (n_keys) zero on entry
  uint32_t* tsum;           // (kTsumCopies, scan_tiles) zero on entry
  int tile_shift;
  int scan_tiles;
  int32_t* counts;          // {K, V}: cleared here
};

// quantise (reference src/model_baseline.py:92, 99-101: same float32 operations as
// quantize_point_core; a division by a power of two is done as the bit-identical multiplication),
// cell, key, histogram, tile totals.  Every lane of the warp must call it (`in` = has a point).
static_assert(kKeyTile == 8, "the key arithmetic below shifts by 3");
__device__ __forceinline__ void plan_emit_point(float gx, float gy, float gz, int b, long long p, bool in,
                                                const PlanCellsArgs& a, int lane, int copy) {
  uint32_t tile = 0xffffffffu;                       // dropped / no point: no tile
  if (in) {
    const GridDev& g = a.grid;
    const float sx = __fsub_rn(gx, g.off[0]), sy = __fsub_rn(gy, g.off[1]), sz = __fsub_rn(gz, g.off[2]);
    const float qx = g.rdx[0] != 0.f ? __fmul_rn(sx, g.rdx[0]) : __fdiv_rn(sx, g.dx[0]);
    const float qy = g.rdx[1] != 0.f ? __fmul_rn(sy, g.rdx[1]) : __fdiv_rn(sy, g.dx[1]);
    const float qz = g.rdx[2] != 0.f ? __fmul_rn(sz, g.rdx[2]) : __fdiv_rn(sz, g.dx[2]);
    // .long() truncates toward zero, so (-1, 0) lands in voxel 0 and is KEPT; NaN / inf fail
    const bool keep = (qx > -1.0f) && (qx < g.nxf[0]) && (qy > -1.0f) && (qy < g.nxf[1]) &&
                      (qz > -1.0f) && (qz < g.nxf[2]);
    int32_t cell = -1, key = -1;
    if (keep) {
      const int ix = __float2int_rz(qx), iy = __float2int_rz(qy), iz = __float2int_rz(qz);
      cell = ((b * g.nx[0] + ix) * g.nx[1] + iy) * g.nx[2] + iz;
      key = ((((b * a.keys.XT + (ix >> 3)) * a.keys.YT + (iy >> 3)) * 8 + (ix & 7)) * 8 + (iy & 7)) * g.nx[2] + iz;
#ifndef LSS_DBG_P1_NOATOM
      atomicAdd(a.cnt + key, 1u);                    // result unused: RED.ADD
#endif
      tile = static_cast<uint32_t>(key) >> a.tile_shift;
    }
    a.cells[p] = cell;
    a.key_of_point[p] = key;
  }
  // tile totals: the lanes of a warp are neighbours in the image, i.e. on the map, so they fall
  // into one or two scan tiles -- one RED.ADD per distinct tile, spread over kTsumCopies replicas of
  // the totals (a few hundred addresses take every warp's adds: one copy serialises in L2)
#ifndef LSS_DBG_P1_NOTSUM
  const uint32_t peers = __match_any_sync(0xffffffffu, tile);
  if (tile != 0xffffffffu && (__ffs(peers) - 1) == lane)
    atomicAdd(a.tsum + copy * a.scan_tiles + tile, static_cast<uint32_t>(__popc(peers)));
#endif
}

// dense variant: the ego-frame points come from a materialised geometry tensor
__global__ void __launch_bounds__(kPlanThreads)
plan_cells_dense_kernel(PlanCellsArgs a) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.counts[0] = 0; a.counts[1] = 0; }
  const long long p = (long long)blockIdx.x * kPlanThreads + threadIdx.x;
  const bool in = p < a.P;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  int b = 0;
  if (in) {
    gx = __ldg(a.dense_geom + p * 3 + 0);
    gy = __ldg(a.dense_geom + p * 3 + 1);
    gz = __ldg(a.dense_geom + p * 3 + 2);
    b = static_cast<int>(a.div_pps.div(static_cast<uint32_t>(p)));
  }
  plan_emit_point(gx, gy, gz, b, p, in, a, lane, static_cast<int>((blockIdx.x * 8 + (threadIdx.x >> 5)) % kTsumCopies));
}

// fused variant: grid = (pixel blocks, depth chunks, cameras).  A thread owns one pixel (h, w) of
// one camera and walks kPlanDepths depth bins: everything that does not depend on the depth --
// the camera's matrices, the pixel's (u, v) and the first two products of every row of
// inverse(post_rots) @ p -- is computed once.  Operation order and rounding are those of
// geometry_rank_kernel (reference src/model_baseline.py:59-68): (m0*p0 + m1*p1) + m2*p2, no FMA.
#ifndef LSS_PLAN_DEPTHS
#define LSS_PLAN_DEPTHS 8
#endif
constexpr int kPlanDepths = LSS_PLAN_DEPTHS;

__global__ void __launch_bounds__(kPlanThreads)
plan_cells_kernel(PlanCellsArgs a) {
  __shared__ float s_cam[24];
  const int tid = threadIdx.x, lane = tid & 31;
  const int bn = blockIdx.z;
  pdl_wait();
  if (blockIdx.x == 0 && blockIdx.y == 0 && bn == 0 && tid == 0) { a.counts[0] = 0; a.counts[1] = 0; }
  const int second = blockDim.x >= 64 ? 32 : 1;          // the two inversions run in different warps when there are two
#ifdef LSS_DBG_P1_NOPREP
  if (tid < 24) s_cam[tid] = a.geom.rots[bn * 9 + (tid % 9)];
  if (false) {
#else
  if (tid == 0) {
#endif
    float pr[9], ipr[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) pr[j] = a.geom.post_rots[bn * 9 + j];
    inverse3x3(pr, ipr);
#pragma unroll
    for (int j = 0; j < 9; ++j) s_cam[j] = ipr[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) s_cam[18 + j] = a.geom.post_trans[bn * 3 + j];
  } else if (tid == second) {
    float r[9], k[9], ii[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) { r[j] = a.geom.rots[bn * 9 + j]; k[j] = a.geom.intrins[bn * 9 + j]; }
    inverse3x3(k, ii);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        s_cam[9 + i * 3 + j] = dot3_nofma(r[i * 3 + 0], r[i * 3 + 1], r[i * 3 + 2], ii[0 + j], ii[3 + j], ii[6 + j]);
#pragma unroll
    for (int j = 0; j < 3; ++j) s_cam[21 + j] = a.geom.trans[bn * 3 + j];
  }
  __syncthreads();
  const int HW = a.geom.fH * a.geom.fW;
  const int hw = blockIdx.x * blockDim.x + tid;
  const bool in = hw < HW;
  uint32_t h, w;
  a.div_w.divmod(static_cast<uint32_t>(in ? hw : 0), h, w);
  float c[24];
#pragma unroll
  for (int j = 0; j < 24; ++j) c[j] = s_cam[j];
  const float p0 = __fsub_rn(__ldg(a.geom.us + w), c[18]);
  const float p1 = __fsub_rn(__ldg(a.geom.vs + h), c[19]);
  const float a0 = __fadd_rn(__fmul_rn(c[0], p0), __fmul_rn(c[1], p1));
  const float a1 = __fadd_rn(__fmul_rn(c[3], p0), __fmul_rn(c[4], p1));
  const float a2 = __fadd_rn(__fmul_rn(c[6], p0), __fmul_rn(c[7], p1));
  const int b = static_cast<int>(a.div_n.div(static_cast<uint32_t>(bn)));
  const int copy = static_cast<int>((blockIdx.x + blockIdx.y + blockIdx.z + (tid >> 5)) % kTsumCopies);
  const int d0 = blockIdx.y * kPlanDepths;
  const int d1 = min(a.geom.D, d0 + kPlanDepths);
  long long p = ((long long)bn * a.geom.D + d0) * HW + hw;
#pragma unroll 4
  for (int d = d0; d < d1; ++d, p += HW) {
    const float p2 = __fsub_rn(__ldg(a.geom.ds + d), c[20]);
    const float q0 = __fadd_rn(a0, __fmul_rn(c[2], p2));
    const float q1 = __fadd_rn(a1, __fmul_rn(c[5], p2));
    const float q2 = __fadd_rn(a2, __fmul_rn(c[8], p2));
    const float r0 = __fmul_rn(q0, q2), r1 = __fmul_rn(q1, q2), r2 = q2;
    const float gx = __fadd_rn(dot3_nofma(c[9], c[10], c[11], r0, r1, r2), c[21]);
    const float gy = __fadd_rn(dot3_nofma(c[12], c[13], c[14], r0, r1, r2), c[22]);
    const float gz = __fadd_rn(dot3_nofma(c[15], c[16], c[17], r0, r1, r2), c[23]);
    plan_emit_point(gx, gy, gz, b, p, in, a, lane, copy);
  }
}

// ---------------------------------------------------------------------------
// P2: exclusive prefix over the per-key counts.  CTA = one scan tile of
// 1 << tile_shift keys; its base is the sum of the tile totals before it.
// By-products: K (kept points), V (occupied voxels).
// ---------------------------------------------------------------------------
struct PlanScanArgs {
  const uint32_t* cnt;   // (n) per-key counts
  const uint32_t* tsum;  // (kTsumCopies, tiles) totals per scan tile
  int n;                 // n_keys
  int tiles;
  int tile_shift;
  int32_t* key_start;    // (n + 1)
  int32_t* counts;       // {K, V}, zero on entry
};

__global__ void __launch_bounds__(kPlanThreads)
plan_scan_kernel(PlanScanArgs a) {
  __shared__ uint32_t s_warp[kPlanThreads / 32];
  __shared__ uint32_t s_base;
  __shared__ int s_occ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  pdl_wait();
  // base of this tile: sum of the totals of the tiles before it
  uint32_t part = 0;
  for (int j = tid; j < tile; j += kPlanThreads) {
    uint32_t v[kTsumCopies];
#pragma unroll
    for (int c = 0; c < kTsumCopies; ++c) v[c] = __ldg(a.tsum + c * a.tiles + j);     // independent loads
#pragma unroll
    for (int c = 0; c < kTsumCopies; ++c) part += v[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_warp[warp] = part;
  if (tid == 0) s_occ = 0;
  __syncthreads();
  if (tid == 0) {
    uint32_t b = 0;
#pragma unroll
    for (int w = 0; w < kPlanThreads / 32; ++w) b += s_warp[w];
    s_base = b;
  }
  __syncthreads();
  uint32_t run = s_base;
  const long long tile_lo = (long long)tile << a.tile_shift;
  long long tile_hi = tile_lo + (1ll << a.tile_shift);
  if (tile_hi > a.n) tile_hi = a.n;
  int occ = 0;
  for (long long base = tile_lo; base < tile_hi; base += kScanBlock) {
    const long long first = base + (long long)tid * 4;
    uint32_t v[4];
    const bool full = first + 4 <= a.n;
    if (full) {
      const uint4 t = *reinterpret_cast<const uint4*>(a.cnt + first);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (first + k < a.n) ? a.cnt[first + k] : 0u;
    }
    const uint32_t sum = v[0] + v[1] + v[2] + v[3];
    occ += (v[0] ? 1 : 0) + (v[1] ? 1 : 0) + (v[2] ? 1 : 0) + (v[3] ? 1 : 0);
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncthreads();                                  // s_warp is free again
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, agg = 0;
#pragma unroll
    for (int w = 0; w < kPlanThreads / 32; ++w) {
      const uint32_t x = s_warp[w];
      woff += (w < warp) ? x : 0u;
      agg += x;
    }
    uint32_t r = run + woff + incl - sum;
    if (full) {
      int4 t;
      t.x = static_cast<int>(r); r += v[0];
      t.y = static_cast<int>(r); r += v[1];
      t.z = static_cast<int>(r); r += v[2];
      t.w = static_cast<int>(r);
      *reinterpret_cast<int4*>(a.key_start + first) = t;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (first + k < a.n) a.key_start[first + k] = static_cast<int>(r);
        r += v[k];
      }
    }
    run += agg;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) occ += __shfl_xor_sync(0xffffffffu, occ, o);
  if (lane == 0 && occ) atomicAdd(&s_occ, occ);
  __syncthreads();
  if (tid == 0) {
    if (s_occ) atomicAdd(a.counts + 1, s_occ);
    if (tile == a.tiles - 1) {
      a.key_start[a.n] = static_cast<int>(run);
      a.counts[0] = static_cast<int>(run);
    }
  }
}

// ---------------------------------------------------------------------------
// P3: every kept point takes one slot of its key's run and parks {cell, id}
// there.  The cursor is the histogram itself, counted back down to zero -- the
// state the next call expects, so the workspace cleans itself.  Slot order
// inside a run is arbitrary here; P4 fixes it.
// ---------------------------------------------------------------------------
struct PlanScatterArgs {
  const int32_t* key_of_point;
  const int32_t* cells;
  long long P;
  uint32_t* cnt;
  const int32_t* key_start;
  int2* tmp;               // (P) {output cell, point id} in arrival order
  uint32_t* tsum;          // wiped for the next call
  int scan_tiles;
};

__global__ void __launch_bounds__(kPlanThreads)
plan_scatter_kernel(PlanScatterArgs a) {
  pdl_wait();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < a.scan_tiles * kTsumCopies; i += kPlanThreads) a.tsum[i] = 0u;
  const long long p = (long long)blockIdx.x * kPlanThreads + threadIdx.x;
  if (p >= a.P) return;
  const int32_t key = __ldg(a.key_of_point + p);
  if (key < 0) return;
  const int32_t cell = __ldg(a.cells + p);
  const int base = __ldg(a.key_start + key);
  const int n = __ldg(a.key_start + key + 1) - base;
  int slot = 0;
  if (n == 1) a.cnt[key] = 0u;                         // alone in its voxel: no cursor needed
  else slot = static_cast<int>(atomicSub(a.cnt + key, 1u)) - 1;
  a.tmp[base + slot] = make_int2(cell, static_cast<int32_t>(p));
}

// ---------------------------------------------------------------------------
// P4: ascending point order inside every run (== the stable sort order).
// ---------------------------------------------------------------------------
struct PlanOrderArgs {
  const int2* tmp;            // (P) {cell, id} in arrival order; first K valid
  const int32_t* key_start;   // only key_start[n_keys] = K is read here (and bounds of long runs)
  int n_keys;
  long long P;
  int2* rec;                  // (P) {cell, id} in (key, id) order; {-1, 0} beyond K
  KeyMap keys;
};

__device__ __forceinline__ void cmpxchg_asc_y(volatile int2* v, int i, int l) {
  const int32_t x = v[i].y, y = v[l].y;
  if (y < x) { v[i].y = y; v[l].y = x; }
}

__global__ void __launch_bounds__(kPlanThreads)
plan_order_kernel(PlanOrderArgs a) {
  __shared__ int32_t s_run[kPlanThreads / 32][kRunWarp];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  const int K = __ldg(a.key_start + a.n_keys);
  const long long i = (long long)blockIdx.x * kPlanThreads + tid;
  int pend_start = 0, pend_n = 0;
  if (i < a.P) {
    if (i >= K) {
      a.rec[i] = make_int2(-1, 0);
    } else {
      const int ii = static_cast<int>(i);
      const int2 me = __ldg(a.tmp + ii);
      // the run of this slot = the neighbours with the same cell; count the smaller ids on the way.  The
      // kOrderWin neighbours on either side are loaded up front (independent loads, one round trip); a
      // slot that sees both ends of its run among them has seen the whole run and is done.
      int2 lw[kOrderWin], rw[kOrderWin];
#pragma unroll
      for (int k = 0; k < kOrderWin; ++k) {
        lw[k] = make_int2(-2, 0); rw[k] = make_int2(-2, 0);          // -2: no such slot (cells are >= 0)
        if (ii - 1 - k >= 0) lw[k] = __ldg(a.tmp + ii - 1 - k);
        if (ii + 1 + k < K) rw[k] = __ldg(a.tmp + ii + 1 + k);
      }
      int rank = 0, left = 0;
      bool lgo = true, rgo = true;
#pragma unroll
      for (int k = 0; k < kOrderWin; ++k) {
        lgo = lgo && lw[k].x == me.x;
        rgo = rgo && rw[k].x == me.x;
        rank += (lgo && lw[k].y < me.y) ? 1 : 0;
        rank += (rgo && rw[k].y < me.y) ? 1 : 0;
        left += lgo ? 1 : 0;
      }
      if (!lgo && !rgo) {
        a.rec[ii - left + rank] = me;                                  // the whole run was inside the window
      } else {
        // longer run: its bounds come from the interval table; up to kRunSerial points every slot counts
        // the smaller ids itself (neighbouring lanes sit in the same run: same trip count, the loads are
        // broadcasts out of L1), beyond that the warp of the run's first slot sorts it
        const uint32_t key = a.keys.key_of_cell(static_cast<uint32_t>(me.x));
        const int s0 = __ldg(a.key_start + key), e0 = __ldg(a.key_start + key + 1);
        if (e0 - s0 <= kRunSerial) {
          int r = 0;
#pragma unroll 4
          for (int j = s0; j < e0; ++j) r += (__ldg(&a.tmp[j].y) < me.y) ? 1 : 0;
          a.rec[s0 + r] = me;
        } else if (ii == s0) {
          pend_start = s0;
          pend_n = e0 - s0;
        }
      }
    }
  }
  // ---- runs of more than kRunSerial points, by the warp that holds their first slot, one at a time ----
  uint32_t pend = __ballot_sync(0xffffffffu, pend_n > 0);
  while (pend) {
    const int src = __ffs(pend) - 1;
    pend &= pend - 1;
    const int start = __shfl_sync(0xffffffffu, pend_start, src);
    const int n = __shfl_sync(0xffffffffu, pend_n, src);
    const int32_t cell = __ldg(&a.tmp[start].x);
    int n2 = 32;
    while (n2 < n) n2 <<= 1;
    __syncwarp();
    // bitonic network whose compare-exchanges all point the same way (the +inf padding never has to
    // move): in shared memory up to kRunWarp ids, in place in the output beyond (adversarial inputs:
    // everything in a few voxels)
    if (n <= kRunWarp) {
      int32_t* v = s_run[warp];
      for (int k = lane; k < n2; k += 32) v[k] = (k < n) ? __ldg(&a.tmp[start + k].y) : 0x7fffffff;
      __syncwarp();
      for (int lk = 1; (1 << lk) <= n2; ++lk) {                 // k = 1 << lk
        const int k = 1 << lk, half = k >> 1;
        for (int t = lane; t < (n2 >> 1); t += 32) {
          const int blk = t >> (lk - 1), r0 = t & (half - 1);
          const int x = blk * k + r0, l = blk * k + (k - 1 - r0);
          const int32_t p = v[x], q = v[l];
          if (q < p) { v[x] = q; v[l] = p; }
        }
        __syncwarp();
        for (int lj = lk - 2; lj >= 0; --lj) {                    // j = 1 << lj
          const int j = 1 << lj;
          for (int t = lane; t < (n2 >> 1); t += 32) {
            const int blk = t >> lj, r0 = t & (j - 1);
            const int x = blk * 2 * j + r0, l = x + j;
            const int32_t p = v[x], q = v[l];
            if (q < p) { v[x] = q; v[l] = p; }
          }
          __syncwarp();
        }
      }
      for (int k = lane; k < n; k += 32) a.rec[start + k] = make_int2(cell, v[k]);
    } else {
      volatile int2* v = a.rec + start;
      for (int k = lane; k < n; k += 32) { v[k].x = cell; v[k].y = a.tmp[start + k].y; }
      __threadfence_block();
      __syncwarp();
      for (int k = 2; k <= n2; k <<= 1) {
        const int half = k >> 1;
        for (int t = lane; t < (n2 >> 1); t += 32) {
          const int blk = t / half, r0 = t - blk * half;
          const int x = blk * k + r0, l = blk * k + (k - 1 - r0);
          if (l < n) cmpxchg_asc_y(v, x, l);
        }
        __threadfence_block();
        __syncwarp();
        for (int j = k >> 2; j > 0; j >>= 1) {
          for (int t = lane; t < (n2 >> 1); t += 32) {
            const int blk = t / j, r0 = t - blk * j;
            const int x = blk * 2 * j + r0, l = x + j;
            if (l < n) cmpxchg_asc_y(v, x, l);
          }
          __threadfence_block();
          __syncwarp();
        }
      }
    }
  }
}

}  // namespace lss
