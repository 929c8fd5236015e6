// The plan: everything between the calibration tensors and the per-cell point
// lists the pooling kernels consume, i.e. the geometry half of get_voxels
// (reference src/model_baseline.py:128-131) plus the argsort of :110 and the
// interval detection of src/tools.py:196-197.
//
// The voxel rank has few distinct values (n_cells = B*X*Y*Z, 320 k at the
// headline config, about one point per cell), so the sort is a ONE-DIGIT radix
// sort whose digit is the whole key -- a counting sort -- in three small kernels:
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
//               the histogram counted back to zero) and the LAST point to arrive
//               in a run puts the run in ascending point order, which is the
//               order a STABLE sort gives (argsort on torch's radix path): runs
//               of <= 8 points in registers, <= 256 with its warp, longer ones
//               (adversarial inputs) by the last CTA of the grid (bitonic).
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
constexpr int kPlanItems = 4;
constexpr int kPlanTile = kPlanThreads * kPlanItems;   // points per CTA in P1
constexpr int kPlanMaxCams = 24;                       // cameras one P1 tile may span
constexpr int kScanBlock = kPlanThreads * 4;           // keys per scan step (one uint4 per thread)
constexpr int kScanMaxTiles = 1024;                    // tile totals a scan CTA sums for its base
constexpr int kRunSerial = 8;                          // runs up to this: ordered by one thread, in registers
constexpr int kRunWarp = 256;                          // ... up to this: by the finisher's warp; longer: last CTA

struct PlanWorkspace {
  size_t off_cnt, off_done, off_tsum, off_ctl, off_keys, off_tmp, off_long, control_bytes, total_bytes;
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
  w.off_done = off; off += align_up((size_t)n_keys * 4, 256);
  w.off_tsum = off; off += align_up((size_t)w.scan_tiles * 4, 256);
  w.off_ctl = off; off += 256;                        // [0] CTA ticket of P3, [1] number of long runs
  w.control_bytes = off;
  w.off_keys = off; off += align_up((size_t)P * 4, 256);
  w.off_tmp = off; off += align_up((size_t)P * 4, 256);
  w.off_long = off; off += align_up((size_t)(P / kRunWarp + 2) * 4, 256);
  w.total_bytes = off;
  return w;
}

// ---------------------------------------------------------------------------
// P1: per point output cell (-1: dropped) and sort key, histogram of the keys,
// totals per scan tile.
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
  int32_t* key_of_point;    // (P) workspace
  uint32_t* cnt;            // (n_keys) zero on entry
  uint32_t* tsum;           // (scan_tiles) zero on entry
  int tile_shift;
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
    const bool in = p < a.P;
    uint32_t tile = 0xffffffffu;                     // dropped / out of range: no tile
    if (in) {
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
      int32_t key = -1;
      if (cell >= 0) {
        key = static_cast<int32_t>(a.keys.key_of_cell(static_cast<uint32_t>(cell)));
        atomicAdd(a.cnt + key, 1u);                  // result unused: RED.ADD
        tile = static_cast<uint32_t>(key) >> a.tile_shift;
      }
      a.key_of_point[p] = key;
    }
    // tile totals: the lanes of a warp are neighbours in the image, i.e. on the map, so they fall
    // into one or two scan tiles -- one RED.ADD per distinct tile
    const uint32_t peers = __match_any_sync(0xffffffffu, tile);
    if (tile != 0xffffffffu && (__ffs(peers) - 1) == lane) atomicAdd(a.tsum + tile, static_cast<uint32_t>(__popc(peers)));
  }
}

// ---------------------------------------------------------------------------
// P2: exclusive prefix over the per-key counts.  CTA = one scan tile of
// 1 << tile_shift keys; its base is the sum of the tile totals before it.
// By-products: K (kept points), V (occupied voxels).
// ---------------------------------------------------------------------------
struct PlanScanArgs {
  const uint32_t* cnt;   // (n) per-key counts
  const uint32_t* tsum;  // (tiles) totals per scan tile
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
  // base of this tile: sum of the totals of the tiles before it
  uint32_t part = 0;
  for (int j = tid; j < tile; j += kPlanThreads) part += __ldg(a.tsum + j);
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
// P3: scatter + order.  Every kept point takes one slot of its key's run (the
// cursor is the histogram itself, counted back down to zero -- the state the
// next call expects, so the workspace cleans itself) and parks its id there;
// the last point to arrive in a run (a second per-key counter tells) rewrites
// the run in ascending point order as {cell, point} records.
// ---------------------------------------------------------------------------
struct PlanScatterArgs {
  const int32_t* key_of_point;
  const int32_t* cells;
  long long P;
  uint32_t* cnt;
  uint32_t* done;
  const int32_t* key_start;
  int n_keys;
  int32_t* tmp;            // (P) ids in arrival order
  int2* rec;               // (P) {output cell, point id} in (key, point id) order; {-1, 0} beyond K
  uint32_t* tsum;          // wiped for the next call
  int scan_tiles;
  uint32_t* ctl;           // [0] CTA ticket, [1] n_long
  int32_t* long_list;
  KeyMap keys;
};

__device__ __forceinline__ int32_t ld_cg_i32(const int32_t* p) { return __ldcg(p); }

// ascending order of a run of <= kRunSerial ids, in registers
__device__ __forceinline__ void order_run_serial(const int32_t* tmp, int base, int n, int32_t cell, int2* rec) {
  int32_t v[kRunSerial];
#pragma unroll
  for (int i = 0; i < kRunSerial; ++i) v[i] = (i < n) ? ld_cg_i32(tmp + base + i) : 0x7fffffff;
#pragma unroll
  for (int i = 0; i < kRunSerial; ++i) {
    int r = 0;
#pragma unroll
    for (int j = 0; j < kRunSerial; ++j) r += (v[j] < v[i]) ? 1 : 0;
    if (i < n) rec[base + r] = make_int2(cell, v[i]);
  }
}

__device__ __forceinline__ void cmpxchg_asc(volatile int32_t* v, int i, int l) {
  const int32_t x = v[i], y = v[l];
  if (y < x) { v[i] = y; v[l] = x; }
}

__global__ void __launch_bounds__(kPlanThreads)
plan_scatter_kernel(PlanScatterArgs a) {
  __shared__ int32_t s_run[kPlanThreads / 32][kRunWarp];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (blockIdx.x == 0)
    for (int i = tid; i < a.scan_tiles; i += kPlanThreads) a.tsum[i] = 0u;
  const long long p = (long long)blockIdx.x * kPlanThreads + tid;
  int pend_base = 0, pend_n = 0;
  int32_t pend_cell = 0;
  if (p < a.P) {
    const int K = __ldg(a.key_start + a.n_keys);
    if (p >= K) a.rec[p] = make_int2(-1, 0);
    const int32_t key = __ldg(a.key_of_point + p);
    if (key >= 0) {
      const int base = __ldg(a.key_start + key);
      const int n = __ldg(a.key_start + key + 1) - base;
      const int32_t cell = __ldg(a.cells + p);
      if (n == 1) {                                    // alone in its voxel: nothing to order
        a.rec[base] = make_int2(cell, static_cast<int32_t>(p));
        a.cnt[key] = 0u;
      } else {
        const uint32_t slot = atomicSub(a.cnt + key, 1u) - 1u;
        __stcg(a.tmp + base + static_cast<int>(slot), static_cast<int32_t>(p));
        __threadfence();
        const uint32_t fin = atomicAdd(a.done + key, 1u) + 1u;
        if (fin == static_cast<uint32_t>(n)) {         // every id of the run is parked and visible
          a.done[key] = 0u;
          __threadfence();
          if (n <= kRunSerial) order_run_serial(a.tmp, base, n, cell, a.rec);
          else if (n <= kRunWarp) { pend_base = base; pend_n = n; pend_cell = cell; }
          else a.long_list[atomicAdd(a.ctl + 1, 1u)] = key;
        }
      }
    }
  }
  // runs of 9 .. 256 points: the finisher's warp orders them, one run at a time
  uint32_t pend = __ballot_sync(0xffffffffu, pend_n > 0);
  while (pend) {
    const int src = __ffs(pend) - 1;
    pend &= pend - 1;
    const int base = __shfl_sync(0xffffffffu, pend_base, src);
    const int n = __shfl_sync(0xffffffffu, pend_n, src);
    const int32_t cell = __shfl_sync(0xffffffffu, pend_cell, src);
    __syncwarp();
    for (int i = lane; i < n; i += 32) s_run[warp][i] = ld_cg_i32(a.tmp + base + i);
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const int32_t v = s_run[warp][i];
      int r = 0;
      for (int j = 0; j < n; ++j) r += (s_run[warp][j] < v) ? 1 : 0;
      a.rec[base + r] = make_int2(cell, v);
    }
  }
  // runs of more than 256 points (adversarial inputs: everything in a few voxels): the last CTA
  // of the grid sorts them one after the other with a bitonic network whose compare-exchanges
  // all point the same way, so the virtual +inf padding never has to move
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(a.ctl, 1u);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int n_long = static_cast<int>(*reinterpret_cast<volatile uint32_t*>(a.ctl + 1));
  for (int r = 0; r < n_long; ++r) {
    const int32_t key = ld_cg_i32(a.long_list + r);
    const int s = a.key_start[key], n = a.key_start[key + 1] - s;
    const int32_t cell = a.keys.cell_of_key(static_cast<uint32_t>(key));
    volatile int32_t* v = a.tmp + s;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int k = 2; k <= n2; k <<= 1) {
      const int half = k >> 1;
      for (int t = tid; t < (n2 >> 1); t += kPlanThreads) {
        const int blk = t / half, r0 = t - blk * half;
        const int i = blk * k + r0, l = blk * k + (k - 1 - r0);
        if (l < n) cmpxchg_asc(v, i, l);
      }
      __syncthreads();
      for (int j = k >> 2; j > 0; j >>= 1) {
        for (int t = tid; t < (n2 >> 1); t += kPlanThreads) {
          const int blk = t / j, r0 = t - blk * j;
          const int i = blk * 2 * j + r0, l = i + j;
          if (l < n) cmpxchg_asc(v, i, l);
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < n; i += kPlanThreads) a.rec[s + i] = make_int2(cell, v[i]);
    __syncthreads();
  }
  if (tid == 0) { a.ctl[0] = 0u; a.ctl[1] = 0u; }
}

}  // namespace lss
