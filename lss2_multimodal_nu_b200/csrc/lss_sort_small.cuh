// K2, single-wave variant (tiles <= kSmallMaxTiles, i.e. P <= 512 Ki points; the
// headline config has 85 tiles of 4096 keys).
//
// A chained look-back is latency-bound when every tile is resident at once (it
// measured 25 us per pass on B200 for 346k keys), so here each tile publishes
// its whole digit-count row (u16 per bin) plus ONE ready flag, waits for the
// flags of all tiles (one thread per flag, in parallel) and then reads all rows
// with independent loads: column sums give the global bin starts, the partial
// sums over lower tiles give the tile's own offsets.  No global histogram, no
// atomics on the data path, nothing to zero beforehand except the flags (wiped
// by the last CTA out).
//
// With kGeom the keys are not loaded but computed: the first pass evaluates the
// camera preparation (K0) and the frustum geometry (K1') for its 4096 points
// itself, so `ranks` never round-trips through memory.
#pragma once

#include "lss_common.cuh"
#include "lss_geometry.cuh"
#include "lss_sort.cuh"

namespace lss {

constexpr int kSmallThreads = 512;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr int kSmallItems = kSortTile / kSmallThreads;  // 8
constexpr int kSmallMaxTiles = kSmallMaxTilesPlan;
constexpr int kSmallMaxBins = 1 << kSmallMaxBitsPlan;  // 1024
constexpr int kSmallBinsPerThread = kSmallMaxBins / kSmallThreads;  // 2
constexpr int kSmallMaxCams = 48;  // cameras one tile may span in the fused first pass

struct SmallPassArgs {
  const int32_t* keys_in;  // unused when kGeom
  const int32_t* vals_in;  // null: payload = point index
  int32_t* keys_out;
  int32_t* vals_out;
  long long P;
  int shift, bits;
  int tiles;
  uint16_t* rows;   // [tiles][nbins] digit counts per tile (fully rewritten each pass)
  uint32_t* flags;  // [tiles] zero on entry, zero again on exit
  uint32_t* ctl;    // [0] ticket, [1] done counter; zero on entry and on exit
  // MSD mode (plan path): keys >= drop_from are dropped (not sorted), tile 0 publishes the
  // bucket starts (nbins + 1 words) and clears the {K, V} counters for the local pass
  int32_t drop_from;        // INT32_MAX: keep everything
  uint32_t* bucket_start;   // or null
  int32_t* counts;          // or null
  // fused geometry (kGeom)
  GeomArgs geom;
  GridDev grid;
  FastDiv div_ppc, div_hw, div_w;
  int32_t* cells;  // (P) output cell per point, written by the fused pass
};

template <bool kGeom>
__global__ void __launch_bounds__(kSmallThreads)
radix_pass_small_kernel(SmallPassArgs a) {
  __shared__ uint16_t s_wh[kSmallWarps][kSmallMaxBins + 2];
  __shared__ uint32_t s_base[kSmallMaxBins];
  __shared__ uint32_t s_warp_tot[kSmallWarps];
  __shared__ float s_cam[kGeom ? kSmallMaxCams * 24 : 1];
  __shared__ uint32_t s_tile, s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbins = 1 << a.bits;
  const uint32_t mask = static_cast<uint32_t>(nbins - 1);

  if (tid == 0) s_tile = atomicAdd(&a.ctl[0], 1u);  // tiles are claimed in start order
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(&s_wh[0][0]);
    constexpr int kWords = kSmallWarps * (kSmallMaxBins + 2) / 2;
    for (int i = tid; i < kWords; i += kSmallThreads) z[i] = 0;
  }
  __syncthreads();
  const int tile = static_cast<int>(s_tile);
  const long long tile_base = (long long)tile * kSortTile;
  const long long warp_base = tile_base + (long long)warp * (32 * kSmallItems);

  // ---- keys: load, or compute from the calibration (fused K0 + K1') ----------
  int32_t key[kSmallItems];
  if (kGeom) {
    long long last = tile_base + kSortTile - 1;
    if (last >= a.P) last = a.P - 1;
    const int bn0 = static_cast<int>(a.div_ppc.div(static_cast<uint32_t>(tile_base)));
    const int bn1 = static_cast<int>(a.div_ppc.div(static_cast<uint32_t>(last)));
    const int ncam = bn1 - bn0 + 1;  // host guarantees <= kSmallMaxCams
    if (tid < ncam) {
      const int bn = bn0 + tid;
      float r[9], k[9], pr[9], ipr[9], cmb[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        r[j] = a.geom.rots[bn * 9 + j];
        k[j] = a.geom.intrins[bn * 9 + j];
        pr[j] = a.geom.post_rots[bn * 9 + j];
      }
      camera_prep_one(r, k, pr, ipr, cmb);
      float* c = s_cam + tid * 24;
#pragma unroll
      for (int j = 0; j < 9; ++j) { c[j] = ipr[j]; c[9 + j] = cmb[j]; }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        c[18 + j] = a.geom.post_trans[bn * 3 + j];
        c[21 + j] = a.geom.trans[bn * 3 + j];
      }
    }
    __syncthreads();
    PointOut out{nullptr, nullptr, nullptr, a.cells};
#pragma unroll
    for (int j = 0; j < kSmallItems; ++j) {
      const long long p = warp_base + j * 32 + lane;
      key[j] = 0;
      if (p < a.P) {
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
        const float gx = __fadd_rn(dot3_nofma(c[9], c[10], c[11], r0, r1, r2), c[21]);
        const float gy = __fadd_rn(dot3_nofma(c[12], c[13], c[14], r0, r1, r2), c[22]);
        const float gz = __fadd_rn(dot3_nofma(c[15], c[16], c[17], r0, r1, r2), c[23]);
        key[j] = quantize_point_core(gx, gy, gz, static_cast<int>(bn) / a.geom.N, a.grid, p, out);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < kSmallItems; ++j) {
      const long long p = warp_base + j * 32 + lane;
      key[j] = (p < a.P) ? a.keys_in[p] : 0;
    }
  }

  // ---- rank inside the warp (stable multi-split) ------------------------------
  uint16_t offs[kSmallItems];
#pragma unroll
  for (int j = 0; j < kSmallItems; ++j) {
    const long long p = warp_base + j * 32 + lane;
    const uint32_t digit = (p < a.P && key[j] < a.drop_from)
                               ? ((static_cast<uint32_t>(key[j]) >> a.shift) & mask)
                               : static_cast<uint32_t>(nbins);
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const int leader = __ffs(peers) - 1;
    const uint32_t below = __popc(peers & ((1u << lane) - 1u));
    uint32_t old = 0;
    if (lane == leader) {
      old = s_wh[warp][digit];
      s_wh[warp][digit] = static_cast<uint16_t>(old + __popc(peers));
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    offs[j] = static_cast<uint16_t>(old + below);
    __syncwarp();
  }
  __syncthreads();

  // ---- per digit: scan over warps, publish the tile's row --------------------
  uint16_t* my_row = a.rows + (size_t)tile * nbins;
#pragma unroll
  for (int k = 0; k < kSmallBinsPerThread; ++k) {
    const int bin = tid + k * kSmallThreads;
    if (bin < nbins) {
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < kSmallWarps; ++w) {
        const uint32_t c = s_wh[w][bin];
        s_wh[w][bin] = static_cast<uint16_t>(run);
        run += c;
      }
      __stcg(my_row + bin, static_cast<unsigned short>(run));  // <= 4096 keys per tile
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) st_volatile_u32(a.flags + tile, 1u);
  // ---- wait for every tile's row.  Tiles are claimed in ticket order and the
  //      grid is a single wave, so every tile waited on is running or done. -----
  if (tid < a.tiles) {
    while (ld_volatile_u32(a.flags + tid) == 0u) {}
  }
  __threadfence();
  __syncthreads();

  // ---- column sums: total per bin (-> bin starts) and sum over lower tiles ----
  uint32_t tot[kSmallBinsPerThread], low[kSmallBinsPerThread];
#pragma unroll
  for (int k = 0; k < kSmallBinsPerThread; ++k) { tot[k] = 0; low[k] = 0; }
#pragma unroll 4
  for (int t = 0; t < a.tiles; ++t) {
    const unsigned short* row = reinterpret_cast<const unsigned short*>(a.rows) + (size_t)t * nbins;
#pragma unroll
    for (int k = 0; k < kSmallBinsPerThread; ++k) {
      const int bin = tid + k * kSmallThreads;
      if (bin < nbins) {
        const uint32_t c = __ldcg(row + bin);
        tot[k] += c;
        low[k] += (t < tile) ? c : 0u;
      }
    }
  }
  // exclusive scan of the totals over bins (bin = tid + k*512, so scan k-major)
  {
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < kSmallBinsPerThread; ++k) {
      const int bin = tid + k * kSmallThreads;
      const uint32_t v = (bin < nbins) ? tot[k] : 0u;
      uint32_t incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      if (lane == 31) s_warp_tot[warp] = incl;
      __syncthreads();
      uint32_t woff = 0, all = 0;
#pragma unroll
      for (int w = 0; w < kSmallWarps; ++w) {
        const uint32_t x = s_warp_tot[w];
        woff += (w < warp) ? x : 0u;
        all += x;
      }
      if (bin < nbins) {
        const uint32_t start = carry + woff + incl - v;
        s_base[bin] = start + low[k];
        if (a.bucket_start && tile == 0) a.bucket_start[bin] = start;
      }
      carry += all;
      __syncthreads();
    }
    if (tile == 0 && tid == 0) {
      if (a.bucket_start) a.bucket_start[nbins] = carry;
      if (a.counts) { a.counts[0] = 0; a.counts[1] = 0; }
    }
  }

  // ---- scatter ----------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < kSmallItems; ++j) {
    const long long p = warp_base + j * 32 + lane;
    if (p < a.P && key[j] < a.drop_from) {
      const uint32_t digit = (static_cast<uint32_t>(key[j]) >> a.shift) & mask;
      const uint32_t dst = s_base[digit] + s_wh[warp][digit] + offs[j];
      a.keys_out[dst] = key[j];
      a.vals_out[dst] = a.vals_in ? a.vals_in[p] : static_cast<int32_t>(p);
    }
  }

  // ---- last CTA out resets the control words for the next pass / call ---------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(&a.ctl[1], 1u) == static_cast<uint32_t>(a.tiles - 1)) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    for (int i = tid; i < a.tiles; i += kSmallThreads) a.flags[i] = 0u;
    if (tid < 2) a.ctl[tid] = 0u;
  }
}

struct SmallGeom {
  bool enabled;
  GeomArgs geom;
  GridDev grid;
  int32_t* cells;
};

// Enqueue all passes of the single-wave sort.  `ws` is laid out by make_sort_plan
// (s.small); its flags / ctl words must be zero (rows need no initialisation).
inline int run_sort_passes_small(const SortPlan& s, const int32_t* keys, int32_t* out_keys,
                                 int32_t* out_vals, long long P, void* ws, const SmallGeom* sg,
                                 cudaStream_t st) {
  char* c = static_cast<char*>(ws);
  int32_t* tmp_keys = reinterpret_cast<int32_t*>(c + s.off_tmp_keys);
  int32_t* tmp_vals = reinterpret_cast<int32_t*>(c + s.off_tmp_vals);
  uint16_t* rows = reinterpret_cast<uint16_t*>(c + s.off_control);
  uint32_t* flags = reinterpret_cast<uint32_t*>(c + s.off_flags);
  uint32_t* ctl = reinterpret_cast<uint32_t*>(c + s.off_ctl);
  const int32_t* src_k = keys;
  const int32_t* src_v = nullptr;
  for (int i = 0; i < s.passes; ++i) {
    const bool to_out = ((s.passes - 1 - i) % 2) == 0;
    SmallPassArgs a;
    memset(&a, 0, sizeof(a));
    a.keys_in = src_k; a.vals_in = src_v;
    a.keys_out = to_out ? out_keys : tmp_keys;
    a.vals_out = to_out ? out_vals : tmp_vals;
    a.P = P; a.shift = s.shift[i]; a.bits = s.bits[i]; a.tiles = (int)s.tiles;
    a.rows = rows; a.flags = flags; a.ctl = ctl;
    a.drop_from = 0x7fffffff;
    if (i == 0 && sg && sg->enabled) {
      a.geom = sg->geom; a.grid = sg->grid; a.cells = sg->cells;
      const int hw = sg->geom.fH * sg->geom.fW;
      a.div_ppc = FastDiv((uint32_t)(sg->geom.D * hw));
      a.div_hw = FastDiv((uint32_t)hw);
      a.div_w = FastDiv((uint32_t)sg->geom.fW);
      radix_pass_small_kernel<true><<<(unsigned)s.tiles, kSmallThreads, 0, st>>>(a);
    } else {
      radix_pass_small_kernel<false><<<(unsigned)s.tiles, kSmallThreads, 0, st>>>(a);
    }
    LSS_LAUNCH_CHECK("radix_pass_small_kernel");
    src_k = a.keys_out; src_v = a.vals_out;
  }
  return LSS_OK;
}

}  // namespace lss
