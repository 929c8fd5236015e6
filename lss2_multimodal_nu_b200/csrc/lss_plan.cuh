// The plan: everything between the calibration tensors and the per-cell point
// lists the pooling kernels consume, i.e. the geometry half of get_voxels
// (reference src/model_baseline.py:128-131) plus the argsort of :110 and the
// interval detection of src/tools.py:196-197.
//
// The voxel rank has few distinct values (n_cells = B*X*Y*Z, 320 k at the
// headline config, about one point per cell), so the sort is a ONE-DIGIT radix
// sort whose digit is the whole key -- a counting sort -- in four small kernels:
//   P1 cells    camera preparation (K0) + frustum geometry (K1') per point ->
//               output cell; one RED.ADD per kept point builds the histogram
//   P2 scan     exclusive prefix over the cells (single pass, decoupled look-back)
//               -> cell_start[c] .. cell_start[c+1] is cell c's run in the sorted
//               order: the interval table (K3) IS the scanned histogram
//   P3 scatter  every kept point takes a slot of its cell's run (atomic cursor)
//   P4 order    the slots of a run are put in ascending point order, which is the
//               order a STABLE sort gives (argsort on torch's radix path): each
//               point counts the smaller ids in its run.  Runs longer than
//               kPlanLongRun (adversarial inputs) are bitonic-sorted by one CTA.
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
#include "lss_sort.cuh"   // align_up, volatile helpers, look-back flags

namespace lss {

constexpr int kPlanThreads = 256;
constexpr int kPlanItems = 4;
constexpr int kPlanTile = kPlanThreads * kPlanItems;   // points per CTA in P1
constexpr int kPlanMaxCams = 24;                       // cameras one P1 tile may span
constexpr int kScanItems = 16;
constexpr int kScanTile = kPlanThreads * kScanItems;   // cells per CTA in P2
constexpr int kPlanLongRun = 1024;                     // runs above this go to the bitonic path

struct PlanWorkspace {
  size_t off_cnt, off_tmp_pt, off_state, off_ctl, off_long, total_bytes;
  int scan_tiles;
};

inline PlanWorkspace make_plan_workspace(long long P, int32_t n_keys) {
  PlanWorkspace w;
  size_t off = 0;
  w.scan_tiles = (int)(((long long)n_keys + kScanTile - 1) / kScanTile);
  w.off_cnt = off; off += align_up((size_t)n_keys * 4, 256);        // zero between calls
  w.off_state = off; off += align_up((size_t)w.scan_tiles * 4, 256); // zero between calls
  w.off_ctl = off; off += 256;                                       // [0] ticket (zero between calls), [1] n_long
  w.off_tmp_pt = off; off += align_up((size_t)P * 4, 256);
  w.off_long = off; off += align_up((size_t)(P / kPlanLongRun + 2) * 4, 256);
  w.total_bytes = off;
  return w;
}

// ---------------------------------------------------------------------------
// P1: per point output cell (-1: dropped) + histogram of the kept points.
// kDense: the ego-frame points come from a materialised geometry tensor (the
// literal voxel_pooling(geom_feats, x) call) instead of the calibration.
// ---------------------------------------------------------------------------
struct PlanCellsArgs {
  GeomArgs geom;            // raw calibration + frustum axes (fused variant)
  const float* dense_geom;  // (P,3) (dense variant)
  GridDev grid;
  KeyMap keys;
  FastDiv div_ppc, div_hw, div_w, div_n, div_pps;
  long long P;
  int32_t* cells;           // (P)
  uint32_t* cnt;            // (n_keys) zero on entry
  int32_t* counts;          // {K, V}: cleared here
  uint32_t* ctl;            // ctl[1] = n_long: cleared here
};

template <bool kDense>
__global__ void __launch_bounds__(kPlanThreads)
plan_cells_kernel(PlanCellsArgs a) {
  __shared__ float s_cam[kDense ? 1 : kPlanMaxCams * 24];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tile_base = (long long)blockIdx.x * kPlanTile;
  const long long warp_base = tile_base + (long long)warp * (32 * kPlanItems);
  if (blockIdx.x == 0 && tid == 0) {
    a.counts[0] = 0; a.counts[1] = 0;
    a.ctl[1] = 0u;
  }
  int bn0 = 0;
  if (!kDense) {
    long long last = tile_base + kPlanTile - 1;
    if (last >= a.P) last = a.P - 1;
    bn0 = static_cast<int>(a.div_ppc.div(static_cast<uint32_t>(tile_base)));
    const int bn1 = static_cast<int>(a.div_ppc.div(static_cast<uint32_t>(last)));
    // two threads per camera touched by this tile (host: <= kPlanMaxCams): warp 0 inverts
    // post_rots, warp 1 inverts intrins and forms rots @ inverse(intrins)
    const int ncam = bn1 - bn0 + 1;
    if (warp < 2 && lane < ncam) {
      const int bn = bn0 + lane;
      float* c = s_cam + lane * 24;
      if (warp == 0) {
        float pr[9], ipr[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) pr[j] = a.geom.post_rots[bn * 9 + j];
        inverse3x3(pr, ipr);
#pragma unroll
        for (int j = 0; j < 9; ++j) c[j] = ipr[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) c[18 + j] = a.geom.post_trans[bn * 3 + j];
      } else {
        float r[9], k[9], ii[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) { r[j] = a.geom.rots[bn * 9 + j]; k[j] = a.geom.intrins[bn * 9 + j]; }
        inverse3x3(k, ii);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            c[9 + i * 3 + j] = dot3_nofma(r[i * 3 + 0], r[i * 3 + 1], r[i * 3 + 2], ii[0 + j], ii[3 + j], ii[6 + j]);
#pragma unroll
        for (int j = 0; j < 3; ++j) c[21 + j] = a.geom.trans[bn * 3 + j];
      }
    }
    __syncthreads();
  }
  PointOut out{nullptr, nullptr, nullptr, a.cells};
#pragma unroll
  for (int j = 0; j < kPlanItems; ++j) {
    const long long p = warp_base + j * 32 + lane;
    if (p >= a.P) continue;
    float gx, gy, gz;
    int b;
    if (kDense) {
      gx = __ldg(a.dense_geom + p * 3 + 0);
      gy = __ldg(a.dense_geom + p * 3 + 1);
      gz = __ldg(a.dense_geom + p * 3 + 2);
      b = static_cast<int>(a.div_pps.div(static_cast<uint32_t>(p)));
    } else {
      uint32_t bn, i, d, rem, h, w;
      a.div_ppc.divmod(static_cast<uint32_t>(p), bn, i);
      a.div_hw.divmod(i, d, rem);
      a.div_w.divmod(rem, h, w);
      const float* c = s_cam + (static_cast<int>(bn) - bn0) * 24;
      // identical operation order to geometry_rank_kernel (reference model_baseline.py:59-68)
      const float p0 = __fsub_rn(__ldg(a.geom.us + w), c[18]);
      const float p1 = __fsub_rn(__ldg(a.geom.vs + h), c[19]);
      const float p2 = __fsub_rn(__ldg(a.geom.ds + d), c[20]);
      const float q0 = dot3_nofma(c[0], c[1], c[2], p0, p1, p2);
      const float q1 = dot3_nofma(c[3], c[4], c[5], p0, p1, p2);
      const float q2 = dot3_nofma(c[6], c[7], c[8], p0, p1, p2);
      const float r0 = __fmul_rn(q0, q2), r1 = __fmul_rn(q1, q2), r2 = q2;
      gx = __fadd_rn(dot3_nofma(c[9], c[10], c[11], r0, r1, r2), c[21]);
      gy = __fadd_rn(dot3_nofma(c[12], c[13], c[14], r0, r1, r2), c[22]);
      gz = __fadd_rn(dot3_nofma(c[15], c[16], c[17], r0, r1, r2), c[23]);
      b = static_cast<int>(a.div_n.div(bn));
    }
    int32_t cell;
    quantize_point_core(gx, gy, gz, b, a.grid, p, out, &cell);
    if (cell >= 0) atomicAdd(a.cnt + a.keys.key_of_cell(static_cast<uint32_t>(cell)), 1u);   // result unused: RED.ADD
  }
}

// ---------------------------------------------------------------------------
// P2: exclusive prefix over the per-cell counts, one pass.  A tile of 4096 cells
// per CTA; tiles take a ticket so that every tile a CTA waits for has started,
// publish {flag, value} in ONE 32-bit status word (K < 2^30) and look back over
// their predecessors 32 at a time.  By-products: K, V and the list of long runs.
// ---------------------------------------------------------------------------
struct PlanScanArgs {
  const uint32_t* cnt;   // (n) per-key counts
  int n;                 // n_keys
  int tiles;
  int32_t* cell_start;   // (n + 1) key_start
  uint32_t* state;       // [tiles] zero on entry
  uint32_t* ctl;         // [0] ticket (zero on entry), [1] n_long
  int32_t* counts;       // {K, V}, zero on entry
  int32_t* long_list;    // cells whose run exceeds kPlanLongRun
};

__global__ void __launch_bounds__(kPlanThreads)
plan_scan_kernel(PlanScanArgs a) {
  __shared__ uint32_t s_warp[kPlanThreads / 32];
  __shared__ uint32_t s_tile, s_excl;
  __shared__ int s_occ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_tile = atomicAdd(a.ctl, 1u); s_occ = 0; }
  __syncthreads();
  const int tile = static_cast<int>(s_tile);
  const long long base = (long long)tile * kScanTile;
  const long long first = base + (long long)tid * kScanItems;
  const bool full = base + kScanTile <= a.n;

  uint32_t v[kScanItems];
  if (full) {
    const uint4* src = reinterpret_cast<const uint4*>(a.cnt + first);
#pragma unroll
    for (int q = 0; q < kScanItems / 4; ++q) {
      const uint4 t = src[q];
      v[q * 4 + 0] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = (first + k < a.n) ? a.cnt[first + k] : 0u;
  }
  uint32_t sum = 0;
  int occ = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    sum += v[k];
    occ += v[k] ? 1 : 0;
    if (v[k] > static_cast<uint32_t>(kPlanLongRun))
      a.long_list[atomicAdd(a.ctl + 1, 1u)] = static_cast<int32_t>(first + k);
  }
  // block-wide exclusive scan of the thread sums
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) occ += __shfl_xor_sync(0xffffffffu, occ, o);
  if (lane == 31) s_warp[warp] = incl;
  if (lane == 0 && occ) atomicAdd(&s_occ, occ);
  __syncthreads();
  uint32_t woff = 0, agg = 0;
#pragma unroll
  for (int w = 0; w < kPlanThreads / 32; ++w) {
    const uint32_t x = s_warp[w];
    woff += (w < warp) ? x : 0u;
    agg += x;
  }
  if (warp == 0) {
    uint32_t excl = 0;
    if (lane == 0) st_volatile_u32(a.state + tile, (tile == 0 ? kFlagInclusive : kFlagAggregate) | agg);
    if (tile > 0) {
      int look = tile - 1;
      while (true) {
        const int idx = look - lane;
        uint32_t s = kFlagInclusive;   // before the first tile: inclusive prefix 0
        if (idx >= 0) {
          do { s = ld_volatile_u32(a.state + idx); } while ((s & kFlagMask) == 0u);
        }
        const uint32_t incl_mask = __ballot_sync(0xffffffffu, (s & kFlagMask) == kFlagInclusive);
        uint32_t val = s & kValueMask;
        if (incl_mask) val = (lane <= __ffs(incl_mask) - 1) ? val : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (incl_mask) break;
        look -= 32;
      }
      if (lane == 0) st_volatile_u32(a.state + tile, kFlagInclusive | (excl + agg));
    }
    if (lane == 0) s_excl = excl;
  }
  __syncthreads();
  uint32_t run = s_excl + woff + incl - sum;
  if (full) {
    int4* dst = reinterpret_cast<int4*>(a.cell_start + first);
#pragma unroll
    for (int q = 0; q < kScanItems / 4; ++q) {
      int4 t;
      t.x = static_cast<int>(run); run += v[q * 4 + 0];
      t.y = static_cast<int>(run); run += v[q * 4 + 1];
      t.z = static_cast<int>(run); run += v[q * 4 + 2];
      t.w = static_cast<int>(run); run += v[q * 4 + 3];
      dst[q] = t;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (first + k < a.n) a.cell_start[first + k] = static_cast<int>(run);
      run += v[k];
    }
  }
  if (tid == 0) {
    if (s_occ) atomicAdd(a.counts + 1, s_occ);
    if (tile == a.tiles - 1) {
      const int K = static_cast<int>(s_excl + agg);
      a.cell_start[a.n] = K;
      a.counts[0] = K;
    }
  }
}

// ---------------------------------------------------------------------------
// P3: every kept point takes one slot of its cell's run.  The cursor is the
// histogram itself, counted back down to zero -- which is the state the next
// call expects, so the workspace cleans itself.  Slot order is arbitrary here;
// P4 fixes it.
// ---------------------------------------------------------------------------
struct PlanScatterArgs {
  const int32_t* cells;
  long long P;
  KeyMap keys;
  uint32_t* cnt;
  const int32_t* cell_start;   // key_start
  int32_t* tmp_pt;
  int32_t* sorted_cells;   // output cell of each slot (final: the order inside a run does not change it)
  uint32_t* state;   // scan status words: wiped for the next call
  int scan_tiles;
  uint32_t* ctl;     // ticket: wiped for the next call
};

__global__ void __launch_bounds__(kPlanThreads)
plan_scatter_kernel(PlanScatterArgs a) {
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < a.scan_tiles; i += kPlanThreads) a.state[i] = 0u;
    if (threadIdx.x == 0) a.ctl[0] = 0u;
  }
  const long long stride = (long long)gridDim.x * kPlanThreads;
  for (long long p = (long long)blockIdx.x * kPlanThreads + threadIdx.x; p < a.P; p += stride) {
    const int32_t c = __ldg(a.cells + p);
    if (c >= 0) {
      const uint32_t key = a.keys.key_of_cell(static_cast<uint32_t>(c));
      const uint32_t slot = atomicSub(a.cnt + key, 1u) - 1u;
      const int32_t pos = __ldg(a.cell_start + key) + static_cast<int32_t>(slot);
      a.tmp_pt[pos] = static_cast<int32_t>(p);
      a.sorted_cells[pos] = c;
    }
  }
}

// ---------------------------------------------------------------------------
// P4: ascending point order inside every run (== the stable sort order).
// ---------------------------------------------------------------------------
struct PlanOrderArgs {
  const int32_t* tmp_pt;
  int32_t* sorted_cells;      // [K, P) is set to -1 here
  const int32_t* cell_start;  // key_start
  KeyMap keys;
  int n_cells;                // n_keys
  long long P;
  int32_t* sorted_points;
  const uint32_t* ctl;        // [1] n_long
  const int32_t* long_list;
};

__device__ __forceinline__ void cmpxchg_asc(int32_t* v, int i, int l) {
  const int32_t x = v[i], y = v[l];
  if (y < x) { v[i] = y; v[l] = x; }
}

__global__ void __launch_bounds__(kPlanThreads)
plan_order_kernel(PlanOrderArgs a) {
  const int K = __ldg(a.cell_start + a.n_cells);
  const long long stride = (long long)gridDim.x * kPlanThreads;
  for (long long i = (long long)blockIdx.x * kPlanThreads + threadIdx.x; i < a.P; i += stride) {
    if (i >= K) { a.sorted_cells[i] = -1; continue; }
    const int32_t p = __ldg(a.tmp_pt + i);
    const uint32_t c = a.keys.key_of_cell(static_cast<uint32_t>(a.sorted_cells[i]));
    const int s = __ldg(a.cell_start + c), e = __ldg(a.cell_start + c + 1);
    if (e - s <= kPlanLongRun) {
      int rank = 0;
      for (int j = s; j < e; ++j) rank += (__ldg(a.tmp_pt + j) < p) ? 1 : 0;
      a.sorted_points[s + rank] = p;
    }
  }
  // long runs (more than kPlanLongRun points in one voxel): one CTA each, bitonic network whose
  // compare-exchanges all point the same way, so the virtual +inf padding never has to move
  const int n_long = static_cast<int>(a.ctl[1]);
  for (int r = blockIdx.x; r < n_long; r += gridDim.x) {
    const int32_t c = a.long_list[r];
    const int s = a.cell_start[c], n = a.cell_start[c + 1] - s;
    int32_t* v = a.sorted_points + s;
    for (int i = threadIdx.x; i < n; i += kPlanThreads) v[i] = a.tmp_pt[s + i];
    __syncthreads();
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int k = 2; k <= n2; k <<= 1) {
      const int half = k >> 1;
      for (int t = threadIdx.x; t < (n2 >> 1); t += kPlanThreads) {
        const int blk = t / half, r0 = t - blk * half;
        const int i = blk * k + r0, l = blk * k + (k - 1 - r0);
        if (l < n) cmpxchg_asc(v, i, l);
      }
      __syncthreads();
      for (int j = k >> 2; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < (n2 >> 1); t += kPlanThreads) {
          const int blk = t / j, r0 = t - blk * j;
          const int i = blk * 2 * j + r0, l = i + j;
          if (l < n) cmpxchg_asc(v, i, l);
        }
        __syncthreads();
      }
    }
  }
}

}  // namespace lss
