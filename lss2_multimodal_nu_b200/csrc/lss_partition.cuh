// P1 of the plan: camera preparation (K0) + frustum geometry (K1') + a stable
// partition of the ranks on their high `hi_bits` bits, in ONE co-resident kernel.
//
// Why this shape: the whole job is ~350 k points at the headline config, i.e.
// a few microseconds of work if every SM participates.  Fat sort tiles (4096
// keys / CTA -> 85 CTAs) left the GPU 85 % idle and ran 4 k instructions per
// thread; here a tile is 1024 points (339 CTAs, ~3 resident per SM) and the
// cross-tile prefix is computed cooperatively between two grid-wide barriers:
//   A  geometry -> rank per point (kept in registers), per-tile digit counts in
//      shared memory -> one u16 row per tile
//   -- barrier --
//   B  column scans: the CTAs split the digit columns, each scans its columns
//      over all tiles (exclusive prefix per (tile, digit), total per digit)
//   -- barrier --
//   C  bucket starts (scan of the totals), stable in-tile ranking (warp
//      match_any multi-split), scatter of (rank, point) to the partition buffers
// The barriers are counters in global memory; the host only takes this path when
// the occupancy calculator guarantees that all CTAs are resident at once.
#pragma once

#include "lss_common.cuh"
#include "lss_geometry.cuh"
#include "lss_sort.cuh"

namespace lss {

constexpr int kPartThreads = 256;
constexpr int kPartWarps = kPartThreads / 32;
constexpr int kPartItems = 4;
constexpr int kPartTile = kPartThreads * kPartItems;  // 1024 points per CTA
constexpr int kPartMaxBits = 10;
constexpr int kPartMaxBins = 1 << kPartMaxBits;
constexpr int kPartBinsPerThread = kPartMaxBins / kPartThreads;  // 4
constexpr int kPartMaxCams = 24;    // cameras one tile may span
constexpr int kPartMaxTiles = 2048; // column scan handles up to 8 tiles per thread

struct PartitionArgs {
  GeomArgs geom;       // raw calibration (rots, trans, intrins, post_rots, post_trans) + axes
  GridDev grid;
  FastDiv div_ppc, div_hw, div_w, div_n;
  long long P;
  int tiles;
  int shift, bits;     // partition digit = (rank >> shift) & ((1 << bits) - 1)
  int32_t* cells;      // (P) output cell per point, -1 if dropped
  int32_t* part_keys;  // (P) ranks, partitioned (first K entries)
  int32_t* part_vals;  // (P) point ids, partitioned
  uint16_t* rows;      // [tiles][nbins] per-tile digit counts
  uint32_t* excl;      // [tiles][nbins] exclusive prefix over tiles
  uint32_t* totals;    // [nbins]
  uint32_t* bucket_start;  // [nbins + 1]
  int32_t* counts;     // {K, V}: cleared here for the local pass
  uint32_t* barrier;   // [4] zero on entry, zero on exit
};

__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    // all CTAs are resident (checked by the host with the occupancy calculator); the spin is
    // bounded anyway so that a violated assumption shows up as a failed parity check, not a hang
    // (polls back off: hundreds of CTAs hammering one L2 line slow every access to that slice)
    for (uint32_t it = 0; ld_volatile_u32(counter) < expected && it < (1u << 22); ++it) __nanosleep(100);
    __threadfence();
  }
  __syncthreads();
}

// block-wide exclusive scan of one value per thread (kPartThreads threads)
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  __syncthreads();  // s_warp may still be read from a previous call
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t woff = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kPartWarps; ++w) {
    const uint32_t x = s_warp[w];
    woff += (w < warp) ? x : 0u;
    all += x;
  }
  if (total) *total = all;
  return woff + incl - v;
}

__global__ void __launch_bounds__(kPartThreads, 3)
partition_coop_kernel(PartitionArgs a) {
  __shared__ uint32_t s_cnt[kPartMaxBins];                   // A: tile digit counts; C: bucket bases
  __shared__ uint16_t s_wh[kPartWarps][kPartMaxBins + 2];    // C: per-warp digit counters
  __shared__ float s_cam[kPartMaxCams * 24];
  __shared__ uint32_t s_warp[kPartWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbins = 1 << a.bits;
  const uint32_t mask = static_cast<uint32_t>(nbins - 1);
  const int tile = blockIdx.x;
  const long long tile_base = (long long)tile * kPartTile;
  const long long warp_base = tile_base + (long long)warp * (32 * kPartItems);

  // ------------------------------ phase A -------------------------------------
  phase_stamp(0, 0);
  for (int i = tid; i < nbins; i += kPartThreads) s_cnt[i] = 0;
  long long last = tile_base + kPartTile - 1;
  if (last >= a.P) last = a.P - 1;
  const int bn0 = static_cast<int>(a.div_ppc.div(static_cast<uint32_t>(tile_base)));
  const int bn1 = static_cast<int>(a.div_ppc.div(static_cast<uint32_t>(last)));
  {
    // two threads per camera touched by this tile (host: <= kPartMaxCams cameras): warp 0 inverts
    // post_rots, warp 1 inverts intrins and forms rots @ inverse(intrins); the two are independent
    const int ncam = bn1 - bn0 + 1;
    const int cam = lane, role = warp;
    if (role < 2 && cam < ncam) {
      const int bn = bn0 + cam;
      float* c = s_cam + cam * 24;
      if (role == 0) {
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
  }
  // the per-warp digit counters of phase C are idle until then: clear them now
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(&s_wh[0][0]);
    constexpr int kWords = kPartWarps * (kPartMaxBins + 2) / 2;
    for (int i = tid; i < kWords; i += kPartThreads) z[i] = 0;
  }
  __syncthreads();
  phase_stamp(0, 1);

  int32_t key[kPartItems];
  PointOut out{nullptr, nullptr, nullptr, a.cells};
#pragma unroll
  for (int j = 0; j < kPartItems; ++j) {
    const long long p = warp_base + j * 32 + lane;
    key[j] = a.grid.n_cells;  // dropped / out of range
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
      key[j] = quantize_point_core(gx, gy, gz, static_cast<int>(a.div_n.div(bn)), a.grid, p, out);
      if (key[j] < a.grid.n_cells) atomicAdd(&s_cnt[(static_cast<uint32_t>(key[j]) >> a.shift) & mask], 1u);
    }
  }
  __syncthreads();
  {
    uint16_t* row = a.rows + (size_t)tile * nbins;
    for (int i = tid; i < nbins; i += kPartThreads) __stcg(row + i, static_cast<unsigned short>(s_cnt[i]));
  }
  phase_stamp(0, 2);
  grid_barrier(a.barrier + 0, static_cast<uint32_t>(a.tiles));
  phase_stamp(0, 3);

  // ------------------------------ phase B -------------------------------------
  // column scan, one warp per digit column, columns dealt so that every CTA gets its share.
  // Lane l owns up to kQ consecutive tiles: kQ independent loads (one L2 round trip), a serial
  // sum in registers, ONE warp scan of the lane sums, then kQ stores of the running prefix.
  {
    constexpr int kQ = 12;  // 32 * 12 = 384 tiles per sweep
    const unsigned short* rows = reinterpret_cast<const unsigned short*>(a.rows);
    for (int bin = warp * a.tiles + tile; bin < nbins; bin += kPartWarps * a.tiles) {
      uint32_t carry = 0;
      for (int t0 = 0; t0 < a.tiles; t0 += 32 * kQ) {
        uint32_t v[kQ];
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
          const int t = t0 + lane * kQ + q;
          v[q] = (t < a.tiles) ? static_cast<uint32_t>(__ldcg(rows + (size_t)t * nbins + bin)) : 0u;
        }
#pragma unroll
        for (int q = 0; q < kQ; ++q) sum += v[q];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += n;
        }
        uint32_t run = carry + incl - sum;
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
          const int t = t0 + lane * kQ + q;
          if (t < a.tiles) __stcg(a.excl + (size_t)t * nbins + bin, run);
          run += v[q];
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) __stcg(a.totals + bin, carry);
    }
  }
  phase_stamp(0, 4);
  grid_barrier(a.barrier + 1, static_cast<uint32_t>(a.tiles));
  phase_stamp(0, 5);

  // ------------------------------ phase C -------------------------------------
  // bucket starts = exclusive scan of the totals (every CTA, 4 consecutive bins per thread)
  {
    uint32_t t4[kPartBinsPerThread];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kPartBinsPerThread; ++k) {
      const int bin = tid * kPartBinsPerThread + k;
      t4[k] = (bin < nbins) ? __ldcg(a.totals + bin) : 0u;
      sum += t4[k];
    }
    uint32_t all;
    uint32_t run = block_exclusive_scan(sum, s_warp, &all);
#pragma unroll
    for (int k = 0; k < kPartBinsPerThread; ++k) {
      const int bin = tid * kPartBinsPerThread + k;
      if (bin < nbins) {
        s_cnt[bin] = run + __ldcg(a.excl + (size_t)tile * nbins + bin);
        if (tile == 0) a.bucket_start[bin] = run;
      }
      run += t4[k];
    }
    if (tile == 0 && tid == 0) {
      a.bucket_start[nbins] = all;
      a.counts[0] = 0;
      a.counts[1] = 0;
    }
  }
  __syncthreads();
  phase_stamp(0, 6);
  uint16_t offs[kPartItems];
#pragma unroll
  for (int j = 0; j < kPartItems; ++j) {
    const uint32_t digit = (key[j] < a.grid.n_cells) ? ((static_cast<uint32_t>(key[j]) >> a.shift) & mask)
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
  for (int bin = tid; bin < nbins; bin += kPartThreads) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kPartWarps; ++w) {
      const uint32_t c = s_wh[w][bin];
      s_wh[w][bin] = static_cast<uint16_t>(run);
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPartItems; ++j) {
    if (key[j] < a.grid.n_cells) {
      const uint32_t digit = (static_cast<uint32_t>(key[j]) >> a.shift) & mask;
      const uint32_t dst = s_cnt[digit] + s_wh[warp][digit] + offs[j];
      a.part_keys[dst] = key[j];
      a.part_vals[dst] = static_cast<int32_t>(warp_base + j * 32 + lane);
    }
  }

  // last CTA out clears the barrier words for the next call
  __syncthreads();
  phase_stamp(0, 7);
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(a.barrier + 2, 1u) == static_cast<uint32_t>(a.tiles - 1)) {
      a.barrier[0] = 0; a.barrier[1] = 0; a.barrier[2] = 0;
    }
  }
}

}  // namespace lss
