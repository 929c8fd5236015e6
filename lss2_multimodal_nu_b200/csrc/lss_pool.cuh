// K3 interval detection, lift staging, K4/K4a pooling forward, K5/K5a backward.
#pragma once

#include "lss_common.cuh"

namespace lss {

// --------------------------------------------------------------------------
// K3: runs of equal ranks in the sorted order.  Replaces the boundary mask of
// QuickCumsum.forward (reference src/tools.py:196-197): `last` marks the last
// point of every run.  Heads/tails also record the run's [start, end) in a
// dense table indexed by OUTPUT cell.
// --------------------------------------------------------------------------
struct IntervalArgs {
  const int32_t* sorted_ranks;
  long long P;
  GridDev g;
  FastDiv div_b, div_z, div_y;
  uint8_t* last_mask;   // or null
  int32_t* sorted_cells;  // or null: output cell of every kept sorted point
  int2* cell_range;     // (n_cells) zero on entry
  int32_t* counts;      // {K, V} zero on entry, or null
};

__global__ void __launch_bounds__(256)
intervals_kernel(IntervalArgs a) {
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int kept = 0, tails = 0;
  for (long long i = first; i < a.P; i += stride) {
    const int32_t r = a.sorted_ranks[i];
    const bool valid = r < a.g.n_cells;
    bool tail = false;
    if (valid) {
      const int32_t prev = (i > 0) ? a.sorted_ranks[i - 1] : -1;
      const int32_t next = (i + 1 < a.P) ? a.sorted_ranks[i + 1] : a.g.n_cells;
      const bool head = r != prev;
      tail = r != next;
      if (head || tail || a.sorted_cells) {
        // rank = ((x*Y + y)*Z + z)*B + b  ->  output cell ((b*X + x)*Y + y)*Z + z
        uint32_t t0, b, t1, z, x, y;
        a.div_b.divmod(static_cast<uint32_t>(r), t0, b);
        a.div_z.divmod(t0, t1, z);
        a.div_y.divmod(t1, x, y);
        const int32_t cell = ((static_cast<int32_t>(b) * a.g.nx[0] + static_cast<int32_t>(x)) * a.g.nx[1] +
                              static_cast<int32_t>(y)) * a.g.nx[2] + static_cast<int32_t>(z);
        if (head) a.cell_range[cell].x = static_cast<int>(i);
        if (tail) a.cell_range[cell].y = static_cast<int>(i + 1);
        if (a.sorted_cells) a.sorted_cells[i] = cell;
      }
      ++kept;
      tails += tail ? 1 : 0;
    }
    if (a.last_mask) a.last_mask[i] = tail ? 1 : 0;
  }
  if (a.counts) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kept += __shfl_xor_sync(0xffffffffu, kept, o);
      tails += __shfl_xor_sync(0xffffffffu, tails, o);
    }
    if (lane_id() == 0) {
      if (kept) atomicAdd(&s_cnt[0], kept);
      if (tails) atomicAdd(&s_cnt[1], tails);
    }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(&a.counts[threadIdx.x], s_cnt[threadIdx.x]);
  }
}

// --------------------------------------------------------------------------
// Lift staging.  The pooling kernels want one contiguous row of C context
// values per pixel: feat (BN, C, HW) -> feat_t (BN*HW, C) through a padded
// 32x32 shared-memory tile; 128-bit loads along HW where the input allows it,
// 128-bit stores along C always.  The depth distribution is read in place
// (it is indexed by POINT id: (BN, D, HW) is the reference's own flattening),
// unless it has to be computed first (softmax below).
// grid = (ceil(HW/32), ceil(C/32), BN)
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
feat_stage_kernel(const void* __restrict__ feat, long long feat_bs, int dtype, int vec_ok, int C, int HW,
                  float* __restrict__ feat_t) {
  __shared__ float tile[32][33];
  const int bn = blockIdx.z;
  pdl_wait();
  const size_t sbase = (size_t)bn * feat_bs;
  float* dst = feat_t + (size_t)bn * HW * C;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int t = threadIdx.x;
  if (vec_ok && p0 + 32 <= HW) {                       // float32, 16-byte aligned rows: one float4 per thread
    const int c = c0 + (t >> 3), q = t & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(feat) + sbase + (size_t)c * HW + p0) + q);
    float* row = tile[t >> 3];
    row[4 * q + 0] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
  } else {
    const int tx = t & 31, ty = t >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + k * 8, p = p0 + tx;
      tile[ty + k * 8][tx] = (c < C && p < HW) ? load_as_float(feat, sbase + (size_t)c * HW + p, dtype) : 0.0f;
    }
  }
  __syncthreads();
  // thread -> pixel t>>3, channels 4*(t&7) .. +3 of the tile (C % 4 == 0: a quad is inside or outside)
  const int p = p0 + (t >> 3), cq = 4 * (t & 7);
  if (p < HW && c0 + cq < C) {
    const float4 v = make_float4(tile[cq + 0][t >> 3], tile[cq + 1][t >> 3], tile[cq + 2][t >> 3], tile[cq + 3][t >> 3]);
    *reinterpret_cast<float4*>(dst + (size_t)p * C + c0 + cq) = v;
  }
}

// Producer fusion: the depth distribution is softmax over the D logit channels
// (reference src/modules.py:76-77, `x.softmax(dim=1)`).  One thread per pixel, lanes along HW
// (coalesced): p = exp(x - max) / sum, the expression torch evaluates, written float32 in the
// (BN, D, HW) point order the pooling kernels index by point id.
__global__ void __launch_bounds__(256)
depth_softmax_kernel(const void* __restrict__ logits, long long logits_bs, int dtype, int D, int HW, int BN,
                     float* __restrict__ depth_p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BN * HW) return;
  const int bn = i / HW, hw = i - bn * HW;
  const size_t src = (size_t)bn * logits_bs + hw;
  float* dst = depth_p + (size_t)bn * D * HW + hw;
  float m = load_as_float(logits, src, dtype);
  for (int d = 1; d < D; ++d) m = fmaxf(m, load_as_float(logits, src + (size_t)d * HW, dtype));
  float sum = 0.f;
  for (int d = 0; d < D; ++d) sum += expf(load_as_float(logits, src + (size_t)d * HW, dtype) - m);
  for (int d = 0; d < D; ++d) dst[(size_t)d * HW] = expf(load_as_float(logits, src + (size_t)d * HW, dtype) - m) / sum;
}

// --------------------------------------------------------------------------
// K4 / K4a forward, channels-innermost BEV.
//
// The BEV map is (cell, C) with cell = ((b*X + x)*Y + y)*Z + z, one voxel = one
// contiguous line, and the plan's record list {cell, point} is sorted by the
// cell's tile-major key (KeyMap), so key k owns rec[key_start[k] .. key_start[k+1])
// and neighbours in the list are neighbours on the map.  Every output element is
// written exactly once (this is the torch.zeros + index_put + cat of reference
// src/model_baseline.py:120-124) by two kinds of CTAs that run side by side:
//   * FILL CTAs (one per SM) stream zeros into the empty voxels (75 % of the map
//     at the headline config): a warp reads the 33 interval bounds of 32
//     consecutive keys with one coalesced load (the next block's are prefetched)
//     and every run of empty voxels inside a tile row -- a contiguous piece of
//     the map -- leaves as ONE bulk async store (TMA engine, cp.async.bulk
//     shared -> global from a 2 KB zero block), so the zero stream costs the SM
//     neither issue slots nor load/store queue entries;
//   * REDUCE CTAs do the warp-level segmented reduction over the sorted records.
//     The unit of work is the POINT (dense and sparse regions of the map cost the
//     same): a warp takes kM*32 consecutive records.  Its lanes first decode one
//     record each -- feature-row offset, depth value (read in place, indexed by
//     point id), interval-head flag -- into shared memory.  Then the warp splits
//     into G = 32/L WALKERS of L lanes; a walker owns the intervals that START in
//     its share of the records (the last one is followed into the records after
//     it) and walks them in order, one 128-bit gather per lane and part of the
//     feature row (L lanes x kNP x 16 bytes [+ L x 8] = one row), so one warp
//     instruction gathers / accumulates G points.  A running sum is closed
//     (stored, zeroed) at every interval head.  No cumsum, no atomics, summation
//     in sort order: bit-reproducible.
//   kFused:  acc += depth[point] * feat_t[pixel(point), :]   (K4: the frustum tensor
//            of src/modules.py:84 is never formed)
//   !kFused: acc += x[point, :]                               (K4a)
// --------------------------------------------------------------------------
struct PoolFwdArgs {
  const void* depth;              // (BN, D*HW) depth distribution, batch stride depth_bs, dtype depth_dtype  [fused]
  long long depth_bs;
  int depth_dtype;
  const float* feat_t;            // (BN*HW, C)   fused
  const float* x;                 // (P, C)       dense
  const int2* rec;                // (P) {output cell, point id} by (key, point id); cell -1 beyond the K kept points
  const int32_t* key_start;       // (n_keys + 1) interval bounds, indexed by key
  float* bev;                     // (n_cells, C)
  long long P;
  KeyMap keys;
  int fill_ctas;                  // CTAs [0, fill_ctas) zero-fill, the rest reduce
  int C, HW;
  int nact;                       // lanes of a walker that own channels (generic layouts: C/4 <= L)
  FastDiv div_dhw, div_hw;
};

constexpr int kPoolThreads = 256;
constexpr int kPoolWarps = kPoolThreads / 32;
constexpr int kZeroBytes = 2048;  // zero block in shared memory: source of the bulk zero stores

#ifndef LSS_FWD_MINB
#define LSS_FWD_MINB 3
#endif
#ifndef LSS_FWD_U
#define LSS_FWD_U 4
#endif
#ifndef LSS_FWD_M
#define LSS_FWD_M 2
#endif
#ifndef LSS_BWD_U2_MINV
#define LSS_BWD_U2_MINV 16     // 16 channels per lane (C = 128): 2 bins per round fit the registers (config 5: 1143 -> 1081 us); 10 per lane: 4 is better (181 vs 238 us)
#endif
#ifndef LSS_BWD_MINB
#define LSS_BWD_MINB 2
#endif
#ifndef LSS_BWD_F2F_MIX
#define LSS_BWD_F2F_MIX 1
#endif

// L2 residency (build with -DLSS_L2_HINTS=1): the forward streams the whole BEV map through L2 while it keeps
// re-reading the staged feature rows (22 MB at config 4 against a 205 MB map): the rows are loaded evict-last
// and the map is stored evict-first, so the rows are not pushed out to HBM by the stream.
#ifndef LSS_L2_HINTS
#define LSS_L2_HINTS 1
#endif
__device__ __forceinline__ uint64_t l2_policy(bool keep) {
  uint64_t p;
  if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_f4_keep(const float4* p, uint64_t pol) {
#if LSS_L2_HINTS
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ void st_f4_stream(float4* p, const float4& v, uint64_t pol) {
#if LSS_L2_HINTS
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
#else
  *p = v;
#endif
}

__device__ __forceinline__ void f4_fma(float d, const float4& f, float4& a) {
  a.x = fmaf(d, f.x, a.x); a.y = fmaf(d, f.y, a.y); a.z = fmaf(d, f.z, a.z); a.w = fmaf(d, f.w, a.w);
}

// zero stream of the forward (see above); returns when the CTA's share of the empty voxels is written
__device__ __forceinline__ void pool_fill_zeros(const PoolFwdArgs& a, float4* s_zero) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < kZeroBytes / 16; e += kPoolThreads) s_zero[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t zsrc = static_cast<uint32_t>(__cvta_generic_to_shared(s_zero));
  const uint64_t pol_fill = LSS_L2_HINTS ? l2_policy(false) : 0;
  (void)pol_fill;
  const int n_blocks = (a.keys.n_keys + 31) >> 5;
  const int stride = a.fill_ctas * kPoolWarps;
  const uint32_t line_bytes = static_cast<uint32_t>(a.C) * 4u;
  int blk = blockIdx.x * kPoolWarps + warp;
  if (blk >= n_blocks) return;
  auto bounds = [&](int bk, int& lo, int& hi) {
    const int k = (bk << 5) + lane;
    lo = 0; hi = 1;                                        // beyond the key space: not ours to write
    if (k < a.keys.n_keys) { lo = __ldg(a.key_start + k); hi = __ldg(a.key_start + k + 1); }
  };
  int lo, hi;
  bounds(blk, lo, hi);
  while (true) {
    const int nxt = blk + stride;
    int nlo = 0, nhi = 1;
    if (nxt < n_blocks) bounds(nxt, nlo, nhi);             // prefetch before the stores go out
    const int k = (blk << 5) + lane;
    const int mycell = (k < a.keys.n_keys) ? a.keys.cell_of_key(static_cast<uint32_t>(k)) : -1;
    const bool empty = mycell >= 0 && hi == lo;
    const uint32_t em = __ballot_sync(0xffffffffu, empty);
    // maximal runs of empty keys inside a group of 8 (= contiguous voxels): the first lane of a
    // run stores the whole run, at most kZeroBytes at a time
    if (empty) {
      const uint32_t g8 = (em >> (lane & 24)) & 0xffu;      // this group's 8 bits
      const int j = lane & 7;
      if (j == 0 || !((g8 >> (j - 1)) & 1u)) {              // run head
        const uint32_t rest = (~(g8 >> j)) & 0xffu;         // first non-empty key after the head
        const int len = rest ? __ffs(rest) - 1 : 8 - j;
        char* dst = reinterpret_cast<char*>(a.bev) + (size_t)mycell * line_bytes;
        uint32_t left = static_cast<uint32_t>(len) * line_bytes;
        while (left) {
          const uint32_t n = left < static_cast<uint32_t>(kZeroBytes) ? left : static_cast<uint32_t>(kZeroBytes);
#if LSS_L2_HINTS
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                       :: "l"(dst), "r"(zsrc), "r"(n), "l"(pol_fill) : "memory");
#else
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       :: "l"(dst), "r"(zsrc), "r"(n) : "memory");
#endif
          dst += n; left -= n;
        }
      }
    }
    if (nxt >= n_blocks) break;
    blk = nxt; lo = nlo; hi = nhi;
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the zero block must outlive the reads
}

// first set bit at a position >= pos in the kW-word mask h (kW * 32 if none)
template <int kW>
__device__ __forceinline__ int first_bit_from(const uint32_t (&h)[kW], int pos) {
  int res = kW * 32;
#pragma unroll
  for (int w = kW - 1; w >= 0; --w) {
    uint32_t bits = h[w];
    const int lo = pos - w * 32;
    if (lo >= 32) bits = 0u;
    else if (lo > 0) bits &= 0xffffffffu << lo;
    if (bits) res = w * 32 + __ffs(bits) - 1;
  }
  return res;
}

template <bool kFused, int L, int kNP, bool kT2, int kM>
__global__ void __launch_bounds__(kPoolThreads, (kNP >= 3 ? 2 : kT2 ? 4 : LSS_FWD_MINB))
pool_fwd_kernel(PoolFwdArgs a) {
  constexpr int G = 32 / L;             // walkers per warp
  constexpr int kW = kM + 1;            // staged 32-record words: the chunk + one word of look-ahead
  constexpr int NR = kW * 32;
  constexpr int S = kM * 32 / G;        // records per walker (a divisor of 32 or 64)
  // walk steps whose gathers are issued together: 4 for rows of up to 64 floats; wider rows and the
  // float4 + float2 layouts do better with 2 and the registers that frees (config 4: 260 -> 238 us,
  // config 5: 1125 -> 1021 us; at C = 64 two steps cost the kernel alone 19.6 -> 21.2 us)
  constexpr int U = (kT2 || kNP >= 3) ? 2 : LSS_FWD_U;
  static_assert(S >= 1 && (S <= 32 ? 32 % S == 0 : S == 64), "a walker's share is a whole number of mask words or a divisor of one");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  __shared__ __align__(128) float4 s_zero[kZeroBytes / 16];
  pdl_wait();
  if (static_cast<int>(blockIdx.x) < a.fill_ctas) {
#ifndef LSS_DBG_FWD_NOFILL
    pool_fill_zeros(a, s_zero);
#endif
    return;
  }
#ifdef LSS_DBG_FWD_NOREDUCE
  return;
#endif

  // ======================= REDUCE: segmented sums over the sorted records ==============
  // staged record: {feature-row offset in 16-byte units, depth bits, cell if the record opens an
  // interval that is not its walker's first else -1, unused}
  __shared__ uint4 s_rec[kPoolWarps][NR + G];
  const long long i0 = ((long long)(blockIdx.x - a.fill_ctas) * kPoolWarps + warp) * (kM * 32);
  if (i0 >= a.P) return;
  // ---- one round of loads: the chunk's and the look-ahead word's records ----
  int cell[kW], pt[kW];
#pragma unroll
  for (int w = 0; w < kW; ++w) {
    const long long i = i0 + w * 32 + lane;
    cell[w] = -1; pt[w] = 0;
    if (i < a.P) { const int2 r = __ldg(a.rec + i); cell[w] = r.x; pt[w] = r.y; }
  }
  int before = -1;
  if (lane == 0 && i0 > 0) before = __ldg(&a.rec[i0 - 1].x);
  uint32_t H[kW], HV[kW], T[kW];                            // heads (cell differs from its predecessor), valid heads, terminator
#pragma unroll
  for (int w = 0; w < kW; ++w) {
    int prev = __shfl_up_sync(0xffffffffu, cell[w], 1);
    if (lane == 0) prev = before;
    before = __shfl_sync(0xffffffffu, cell[w], 31);         // predecessor of the next word's lane 0
    H[w] = __ballot_sync(0xffffffffu, cell[w] != prev);
    HV[w] = H[w] & __ballot_sync(0xffffffffu, cell[w] >= 0);
    T[w] = H[w] & ~HV[w];
  }
  // ---- walker of this lane: records [r0, end) ----
  const int g = lane / L, sub = lane % L;
  const int lo_bit = g * S;                                  // first record of the walker's share
  int r0 = first_bit_from<kW>(HV, lo_bit);                   // first interval that starts in the share
  int end = 0;
  if (r0 < lo_bit + S) {
    // the walker runs to the first head at or after the end of its share; the kept records are a
    // prefix of the list, so the first dropped one (a head that is not valid) ends everything
    end = min(first_bit_from<kW>(H, lo_bit + S), first_bit_from<kW>(T, r0 + 1));
  } else {
    r0 = 0;
  }
  const int n_g = end - r0;
  const bool overflow = n_g > 0 && end == NR;                // the last interval runs past the look-ahead
  const int n_max = __reduce_max_sync(0xffffffffu, n_g);
  if (n_max == 0) return;                                    // earlier warps own all of it
  const int hi = __reduce_max_sync(0xffffffffu, end);
  const int lo = __reduce_min_sync(0xffffffffu, n_g > 0 ? r0 : NR);

  // ---- stage the records the walkers will touch ----
  const uint32_t row16 = static_cast<uint32_t>(a.C) >> 2;    // 16-byte units per feature row
#pragma unroll
  for (int w = 0; w < kW; ++w) {
    const int r = w * 32 + lane;
    if (r >= lo && r < hi && cell[w] >= 0) {
      uint32_t off16;
      float dv;
      if (kFused) {
        uint32_t bn, rem, d, hw;
        a.div_dhw.divmod(static_cast<uint32_t>(pt[w]), bn, rem);
        a.div_hw.divmod(rem, d, hw);
        off16 = (bn * static_cast<uint32_t>(a.HW) + hw) * row16;
        dv = load_as_float(a.depth, (size_t)bn * a.depth_bs + rem, a.depth_dtype);
      } else {
        off16 = static_cast<uint32_t>(pt[w]) * row16;
        dv = 1.0f;
      }
      // a walker's first record opens its first interval: nothing to close there
      bool head = (H[w] >> lane) & 1u;
      if (w < kM && head) {
        const int share_lo = (r / S) * S;                     // share that holds r
        if (first_bit_from<kW>(HV, share_lo) == r) head = false;
      }
      s_rec[warp][r] = make_uint4(off16, __float_as_uint(dv), head ? static_cast<uint32_t>(cell[w]) : 0xffffffffu, 0u);
    }
  }
  __syncwarp();
  // padding record of the walker: its own last record with weight zero (re-gathers a row it has
  // already summed, adds 0 * row), for the steps after its end while other walkers still run
  if (sub == 0) {
    uint4 pad = make_uint4(0u, 0u, 0xffffffffu, 0u);
    if (n_g > 0) pad.x = s_rec[warp][end - 1].x;
    s_rec[warp][NR + g] = pad;
  }
  __syncwarp();

  const uint32_t vsub = (static_cast<int>(sub) < a.nact) ? sub : 0u;   // idle lanes shadow lane 0, never store
  const bool storer = static_cast<int>(sub) < a.nact;
  const char* src = reinterpret_cast<const char*>((kFused ? a.feat_t : a.x) + vsub * 4);
  const char* src2 = reinterpret_cast<const char*>((kFused ? a.feat_t : a.x) + 4 * L * kNP + vsub * 2);
  float* out = a.bev + vsub * 4;
  float* out2 = a.bev + 4 * L * kNP + vsub * 2;
  const uint4* recs = s_rec[warp];
  const uint32_t Cw = static_cast<uint32_t>(a.C);
  const uint64_t pol_keep = LSS_L2_HINTS ? l2_policy(true) : 0, pol_stream = LSS_L2_HINTS ? l2_policy(false) : 0;
  (void)pol_keep; (void)pol_stream;

  float4 acc[kNP > 0 ? kNP : 1];
  float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int p = 0; p < kNP; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
  // output cell of the walker's first interval: its first record carries -1 (no close), so take it
  // from the cell registers of the lane that loaded it
  int cur;
  {
    int c = 0;
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      const int t = __shfl_sync(0xffffffffu, cell[w], r0 & 31);
      if ((r0 >> 5) == w) c = t;
    }
    cur = c;
  }
  auto store_acc = [&]() {
    if (storer) {
      float* o = out + (size_t)static_cast<uint32_t>(cur) * Cw;
#pragma unroll
      for (int p = 0; p < kNP; ++p) st_f4_stream(reinterpret_cast<float4*>(o + p * 4 * L), acc[p], pol_stream);
      if (kT2) *reinterpret_cast<float2*>(out2 + (size_t)static_cast<uint32_t>(cur) * Cw) = acc2;
    }
  };

  for (int t0 = 0; t0 < n_max; t0 += U) {
    uint4 rc[U];
    float4 f[U][kNP > 0 ? kNP : 1];
    float2 f2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      rc[u] = recs[(t0 + u < n_g) ? r0 + t0 + u : NR + g];
#ifdef LSS_DBG_FWD_ROW0
      rc[u].x = 0u;
#endif
      const char* row = src + ((size_t)rc[u].x << 4);
#pragma unroll
      for (int p = 0; p < kNP; ++p) f[u][p] = ldg_f4_keep(reinterpret_cast<const float4*>(row) + p * L, pol_keep);
      if (kT2) f2[u] = __ldg(reinterpret_cast<const float2*>(src2 + ((size_t)rc[u].x << 4)));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float dv = __uint_as_float(rc[u].y);
      if (static_cast<int>(rc[u].z) >= 0) {                   // the record opens a new interval: close the running one
        store_acc();
        cur = static_cast<int>(rc[u].z);
#pragma unroll
        for (int p = 0; p < kNP; ++p)
          acc[p] = make_float4(dv * f[u][p].x, dv * f[u][p].y, dv * f[u][p].z, dv * f[u][p].w);
        if (kT2) acc2 = make_float2(dv * f2[u].x, dv * f2[u].y);
      } else {
#pragma unroll
        for (int p = 0; p < kNP; ++p) f4_fma(dv, f[u][p], acc[p]);
        if (kT2) { acc2.x = fmaf(dv, f2[u].x, acc2.x); acc2.y = fmaf(dv, f2[u].y, acc2.y); }
      }
    }
  }
  // ---- an interval longer than the look-ahead (adversarial inputs): follow it to its end ----
  if (overflow) {
    for (long long j = i0 + NR; j < a.P; ++j) {
      const int2 r = __ldg(a.rec + j);
      if (r.x != cur) break;
      uint32_t off16;
      float dv;
      if (kFused) {
        uint32_t bn, rem, d, hw;
        a.div_dhw.divmod(static_cast<uint32_t>(r.y), bn, rem);
        a.div_hw.divmod(rem, d, hw);
        off16 = (bn * static_cast<uint32_t>(a.HW) + hw) * row16;
        dv = load_as_float(a.depth, (size_t)bn * a.depth_bs + rem, a.depth_dtype);
      } else {
        off16 = static_cast<uint32_t>(r.y) * row16;
        dv = 1.0f;
      }
      const size_t byte = (size_t)off16 << 4;
#pragma unroll
      for (int p = 0; p < kNP; ++p) f4_fma(dv, __ldg(reinterpret_cast<const float4*>(src + byte) + p * L), acc[p]);
      if (kT2) {
        const float2 t2 = __ldg(reinterpret_cast<const float2*>(src2 + byte));
        acc2.x = fmaf(dv, t2.x, acc2.x); acc2.y = fmaf(dv, t2.y, acc2.y);
      }
    }
  }
  if (n_g > 0) store_acc();
}

// --------------------------------------------------------------------------
// K5a dense backward: dx[p, :] = dbev[cell(p), :] (exact gather) or 0.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pool_dense_bwd_nhwc_kernel(const float4* __restrict__ dbev, const int32_t* __restrict__ cells,
                           long long n_elems, int G, FastDiv div_g, float4* __restrict__ dx) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_elems; e += stride) {
    uint32_t p, chunk;
    div_g.divmod(static_cast<uint32_t>(e), p, chunk);
    const int32_t cell = __ldg(cells + p);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cell >= 0) v = __ldg(dbev + (size_t)cell * G + chunk);
    dx[e] = v;
  }
}

// --------------------------------------------------------------------------
// K5 fused backward, channels-innermost dBEV.
//
// A walker of L lanes owns one PIXEL (bn, h, w): its context vector stays in
// registers (as float64) while it visits the pixel's depth bins; the G = 32/L
// walkers of a warp are G consecutive image rows h of one image column (bn, w),
// and the warps of a CTA are the column's row groups x slices of the depth axis.
// Along a camera ray neighbouring rows fall into the same BEV cell (Z is
// collapsed), so the G gathers of one warp instruction mostly hit the same
// line.  For every kept point the voxel gradient g (one contiguous line of the
// channels-innermost dBEV) is gathered once and used twice:
//     d_depth[d] = <g, feat>      d_feat += depth[d] * g
// Only OCCUPIED voxels of dBEV are ever read.  A round handles kU depth bins: the
// first kU lanes of a walker hold {cell, depth} of the round's bins (loaded one
// round ahead) and hand them out by shuffle, the kU gathers are issued together.
// <g, feat> has C terms of order one that cancel, so it is accumulated in
// float64 -- but without the float32 -> float64 conversion instruction, which
// runs at a sixteenth of the FMA rate on this part and was the limiter of the
// first version of this kernel: the 32 bits of g are re-packed with three integer
// operations into the float64 whose value is g * 2^-896 exactly (sign |
// exponent | mantissa shifted by three bits; zero stays zero, denormals stay
// exact), and the finished dot is scaled back by 2^896.  The kU partial dots of a
// lane are reduced together with a transposed butterfly.  There are no atomics:
// the depth slices' partial d_feat meet in shared memory in a fixed order,
// results are bit-reproducible.
// A non-finite upstream gradient still reaches the outputs: d_feat sees it
// through the float32 FMA, and a walker whose d_feat partial is not finite
// writes NaN to the d_depth bins of its slice.
// --------------------------------------------------------------------------
struct PoolBwdArgs {
  const float* dbev;        // (n_cells, C)
  const void* depth;        // (BN, D*HW) depth distribution (softmax: probabilities), batch stride depth_bs
  long long depth_bs;
  int depth_dtype;
  const float* feat_t;      // (BN*HW, C)
  const int32_t* cells;     // (BN, D, fH, fW)
  void* ddepth;             // (BN, D, fH, fW) with batch stride ddepth_bs: d_depth, or d_logits when softmax
  void* dfeat;              // (BN, C, fH, fW) with batch stride dfeat_bs
  long long ddepth_bs, dfeat_bs;
  int out_dtype;            // LssDtype of both outputs
  int softmax;              // 1: depth = softmax(logits); emit d_logits = p * (d_depth - sum_d p * d_depth)
  int D, fH, fW, C, BN;
  int nact;                 // lanes of a walker that own channels
  int cols;                 // adjacent image columns per CTA (their rays cross the same voxels: L1 reuse)
  int w_blocks;             // ceil(fW / cols)
  int rg_warps;             // warps of a column along the rows; the column's other warps slice D
  int slices;               // depth slices (warps per row group)
  int d_per_slice;
  int row_blocks;           // CTAs per image column: ceil(fH / (G * rg_warps))
};

constexpr int kBwdMaxWarps = 8;
constexpr int kBwdMaxD = 128;    // softmax backward keeps a pixel's d_depth in shared memory

// the float64 whose value is x * 2^-896 (exact for every finite float32, zero -> zero)
__device__ __forceinline__ double f32_as_scaled_f64(float x) {
  const int b = __float_as_int(x);
  return __hiloint2double((b >> 3) & static_cast<int>(0x8fffffffu), b << 29);
}

template <int L, int kNP, bool kT2, bool kGeneral>
__global__ void __launch_bounds__(32 * kBwdMaxWarps, (4 * kNP + (kT2 ? 2 : 0)) <= 4 ? 4 : LSS_BWD_MINB)
liftsplat_bwd_kernel(PoolBwdArgs a) {
  constexpr int G = 32 / L;                     // pixels (walkers) per warp
  constexpr int kV = 4 * kNP + (kT2 ? 2 : 0);   // channels per lane
  constexpr int U = (kV <= 4 && L >= 8) ? 8 : kV >= LSS_BWD_U2_MINV ? 2 : 4;   // depth bins per round (U <= L)
  constexpr int kLog = L == 32 ? 5 : L == 16 ? 4 : 3;
  constexpr int kLogU = U == 8 ? 3 : U == 4 ? 2 : 1;   // exchange levels that halve the live dots (kLogU <= kLog)
  static_assert(L >= 8 && U <= L, "a walker's first U lanes hold the round's bins");
  __shared__ float s_df[kBwdMaxWarps][G * (L * kV)];         // the warp's partial d_feat
  __shared__ float s_dd[kGeneral ? G * kBwdMaxWarps : 1][kGeneral ? kBwdMaxD : 1];   // d_depth of the CTA's pixels (softmax)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / L, sub = lane % L;
  pdl_wait();
  // CTA -> (image column, block of rows); warp -> (row group, depth slice)
  const int colblk = blockIdx.x / a.row_blocks, rb = blockIdx.x - colblk * a.row_blocks;
  const int bn = colblk / a.w_blocks;
  const int wpc = a.rg_warps * a.slices;                     // warps per column
  const int cw = warp / wpc, wl = warp - cw * wpc;
  const int w = (colblk - bn * a.w_blocks) * a.cols + cw;
  const int rg = wl % a.rg_warps, slice = wl / a.rg_warps;
  const int h = (rb * a.rg_warps + rg) * G + g;
  const bool pix_ok = h < a.fH && w < a.fW;
  const int HW = a.fH * a.fW;
  const int hw = pix_ok ? h * a.fW + w : 0;
  const int crow = (cw * a.rg_warps + rg) * G + g;           // pixel of this walker among the CTA's pixels
  const int d_lo = slice * a.d_per_slice;
  const int d_hi = min(a.D, d_lo + a.d_per_slice);
  const bool lane_act = sub < a.nact;
  const int vsub = lane_act ? sub : 0;

  // {cell, depth} of the bin this lane holds for the coming round (lanes sub < U of every walker)
  const size_t pbase = (size_t)bn * a.D * HW + hw;           // cells index of (bn, d = 0, h, w)
  const size_t dbase = (size_t)bn * a.depth_bs + hw;
  auto load_bin = [&](int d, int& c, float& dv) {
    c = -1; dv = 0.f;
    if (sub < U && d < d_hi && pix_ok) {
      c = __ldg(a.cells + pbase + (size_t)d * HW);
      dv = load_as_float(a.depth, dbase + (size_t)d * HW, a.depth_dtype);
    }
  };
  // bins of the current round (A) and the next (B); the round after is loaded inside the loop
  int cellA, cellB;
  float dvA, dvB;
  load_bin(d_lo + sub, cellA, dvA);
  load_bin(d_lo + U + sub, cellB, dvB);
#ifdef LSS_BWD_PREFETCH
  // the voxel lines of the whole slice are requested from DRAM now, all at once (L2 prefetch): the
  // rounds below then find them in L2 instead of paying one DRAM round trip per round
  for (int d = d_lo + sub; d < d_hi; d += L) {
    if (pix_ok) {
      const int c = __ldg(a.cells + pbase + (size_t)d * HW);
      if (c >= 0) {
        const char* line = reinterpret_cast<const char*>(a.dbev + (size_t)c * a.C);
        for (int b = 0; b < a.C * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(line + b));
      }
    }
  }
#endif

  // the pixel's context vector, float64
  double fd[kV];
  {
    const float* frow = a.feat_t + ((size_t)bn * HW + hw) * a.C;
#pragma unroll
    for (int p = 0; p < kNP; ++p) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(frow + vsub * 4) + p * L);
      fd[4 * p + 0] = f.x; fd[4 * p + 1] = f.y; fd[4 * p + 2] = f.z; fd[4 * p + 3] = f.w;
    }
    if (kT2) {
      const float2 f = __ldg(reinterpret_cast<const float2*>(frow + 4 * L * kNP + vsub * 2));
      fd[4 * kNP] = f.x; fd[4 * kNP + 1] = f.y;
    }
    if (!lane_act || !pix_ok) {
#pragma unroll
      for (int v = 0; v < kV; ++v) fd[v] = 0.0;
    }
    if (LSS_BWD_F2F_MIX) {
#pragma unroll
      for (int p = 0; p < kNP; ++p) { fd[4 * p + 1] *= 1.8928834978668395e-270; fd[4 * p + 3] *= 1.8928834978668395e-270; }   // 2^-896
    }
  }
  float acc[kV];
#pragma unroll
  for (int v = 0; v < kV; ++v) acc[v] = 0.f;

  // which of the U dots this lane ends up holding after the transposed butterfly
  int my_u = 0;
#pragma unroll
  for (int k = 0; k < kLogU; ++k)
    if (sub & (L >> (k + 1))) my_u += U >> (k + 1);
  const bool writer = (sub & ((L >> kLogU) - 1)) == 0;

  const float* gsrc = a.dbev + vsub * 4;
  const float* gsrc2 = a.dbev + 4 * L * kNP + vsub * 2;
  const uint32_t Cw = static_cast<uint32_t>(a.C);

  for (int d0 = d_lo; d0 < d_hi; d0 += U) {
    const int ccell = cellA;
    const float cdv = (ccell >= 0) ? dvA : 0.f;             // dropped point: weight zero
    int cellN;
    float dvN;
    load_bin(d0 + 2 * U + sub, cellN, dvN);                  // two rounds ahead: the gathers below never wait for a cell
    float4 gq[U][kNP > 0 ? kNP : 1];
    float2 g2[U];
    float dv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#ifdef LSS_DBG_BWD_CELL0
      const int c = 0 * __shfl_sync(0xffffffffu, ccell, g * L + u);
#else
      const int c = __shfl_sync(0xffffffffu, ccell, g * L + u);
#endif
      dv[u] = __shfl_sync(0xffffffffu, cdv, g * L + u);
      const size_t off = (size_t)static_cast<uint32_t>(max(c, 0)) * Cw;            // dropped: any valid line, weight zero
#pragma unroll
      for (int p = 0; p < kNP; ++p) gq[u][p] = __ldg(reinterpret_cast<const float4*>(gsrc + off) + p * L);
      if (kT2) g2[u] = __ldg(reinterpret_cast<const float2*>(gsrc2 + off));
    }
    double dot[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double s = 0.0;
#pragma unroll
      for (int p = 0; p < kNP; ++p) {
        const float gv[4] = {gq[u][p].x, gq[u][p].y, gq[u][p].z, gq[u][p].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // LSS_BWD_F2F_MIX: every other channel goes through the conversion instruction instead (its
          // multiplicand carries the 2^-896), which takes work off the integer pipe
          const double gd = (LSS_BWD_F2F_MIX && (k & 1)) ? static_cast<double>(gv[k]) : f32_as_scaled_f64(gv[k]);
          s = fma(gd, fd[4 * p + k], s);
          acc[4 * p + k] = fmaf(dv[u], gv[k], acc[4 * p + k]);
        }
      }
      if (kT2) {
        s = fma(f32_as_scaled_f64(g2[u].x), fd[4 * kNP], s);
        s = fma(f32_as_scaled_f64(g2[u].y), fd[4 * kNP + 1], s);
        acc[4 * kNP] = fmaf(dv[u], g2[u].x, acc[4 * kNP]);
        acc[4 * kNP + 1] = fmaf(dv[u], g2[u].y, acc[4 * kNP + 1]);
      }
      dot[u] = s;
    }
    // transposed butterfly over the walker's L lanes: halve the number of live dots per exchange
#pragma unroll
    for (int k = 0; k < kLog; ++k) {
      const int o = L >> (k + 1);
      if (k < kLogU) {
        const int half = U >> (k + 1);
        const bool upper = (sub & o) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const double send = upper ? dot[i] : dot[i + half];
          const double keep = upper ? dot[i + half] : dot[i];
          dot[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      } else {
        dot[0] += __shfl_xor_sync(0xffffffffu, dot[0], o);
      }
    }
    const int wcell = __shfl_sync(0xffffffffu, ccell, g * L + my_u);
    const int d = d0 + my_u;
    if (writer && pix_ok && d < d_hi) {
      // 2^896: back from the scaled domain; dropped points (weight zero, foreign line) give 0
      float r = static_cast<float>(dot[0] * 5.282945311356653e269);
      if (wcell < 0) r = 0.f;
      const size_t o = (size_t)bn * a.ddepth_bs + (size_t)d * HW + hw;
      if (!kGeneral) reinterpret_cast<float*>(a.ddepth)[o] = r;
      else if (a.softmax) s_dd[crow][d] = r;                  // D <= kBwdMaxD (host)
      else store_from_float(a.ddepth, o, r, a.out_dtype);
    }
    cellA = cellB; dvA = dvB; cellB = cellN; dvB = dvN;
  }
  // a non-finite gradient line shows in the float32 partial d_feat: poison this slice's d_depth
  {
    float chk = 0.f;
#pragma unroll
    for (int v = 0; v < kV; ++v) chk += fabsf(acc[v]);
    const bool bad = !(chk <= 3.4028235e38f);
    uint32_t bm = __ballot_sync(0xffffffffu, bad);
    if (L < 32) bm = (bm >> (g * L)) & ((1u << (L & 31)) - 1u);
    if (bm && pix_ok) {
      const float qnan = __int_as_float(0x7fc00000);
      for (int d = d_lo + sub; d < d_hi; d += L) {
        const size_t o = (size_t)bn * a.ddepth_bs + (size_t)d * HW + hw;
        if (!kGeneral) reinterpret_cast<float*>(a.ddepth)[o] = qnan;
        else if (a.softmax) s_dd[crow][d] = qnan;
        else store_from_float(a.ddepth, o, qnan, a.out_dtype);
      }
    }
  }
  // ---- d_feat: the depth slices' partials meet in shared memory, summed in slice order ----
  {
    float* mine = s_df[warp] + g * (L * kV);
#pragma unroll
    for (int p = 0; p < kNP; ++p)
      *reinterpret_cast<float4*>(mine + p * 4 * L + sub * 4) = make_float4(acc[4 * p], acc[4 * p + 1], acc[4 * p + 2], acc[4 * p + 3]);
    if (kT2) *reinterpret_cast<float2*>(mine + 4 * L * kNP + sub * 2) = make_float2(acc[4 * kNP], acc[4 * kNP + 1]);
  }
  __syncthreads();
  {
    const int rows = a.rg_warps * G;                         // pixel rows of this CTA
    // consecutive threads -> consecutive columns w (the contiguous direction of the output), then rows,
    // channels outer
    for (int i = threadIdx.x; i < a.cols * rows * a.C; i += blockDim.x) {
      const int cc = i % a.cols, t = i / a.cols;
      const int r = t % rows, c = t / rows;
      const int hh = rb * rows + r, ww = (colblk - bn * a.w_blocks) * a.cols + cc;
      if (hh >= a.fH || ww >= a.fW) continue;
      const int wg = r / G, gg = r - wg * G;                 // row group (warp along rows), walker
      float s = 0.f;
      for (int sl = 0; sl < a.slices; ++sl) s += s_df[(cc * a.slices + sl) * a.rg_warps + wg][gg * (L * kV) + c];
      const size_t o = (size_t)bn * a.dfeat_bs + (size_t)c * HW + (size_t)hh * a.fW + ww;
      if (!kGeneral) reinterpret_cast<float*>(a.dfeat)[o] = s;
      else store_from_float(a.dfeat, o, s, a.out_dtype);
    }
  }
  if (kGeneral && a.softmax) {
    // softmax backward (reference: autograd of x.softmax(dim=1), src/modules.py:77), fused:
    // d_logit[d] = p[d] * (d_depth[d] - sum_d' p[d'] * d_depth[d']); one warp per pixel
    const int rows = a.rg_warps * G;
    const int n_warps = blockDim.x >> 5;
    for (int q = warp; q < a.cols * rows; q += n_warps) {
      const int cc = q / rows, r = q - cc * rows;
      const int hh = rb * rows + r, ww = (colblk - bn * a.w_blocks) * a.cols + cc;
      if (hh >= a.fH || ww >= a.fW) continue;
      const size_t pix = (size_t)hh * a.fW + ww;
      double sp = 0.0;
      for (int d = lane; d < a.D; d += 32)
        sp += static_cast<double>(load_as_float(a.depth, (size_t)bn * a.depth_bs + (size_t)d * HW + pix, a.depth_dtype)) *
              static_cast<double>(s_dd[q][d]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sp += __shfl_xor_sync(0xffffffffu, sp, o);
      const float spf = static_cast<float>(sp);
      for (int d = lane; d < a.D; d += 32) {
        const float pd = load_as_float(a.depth, (size_t)bn * a.depth_bs + (size_t)d * HW + pix, a.depth_dtype);
        store_from_float(a.ddepth, (size_t)bn * a.ddepth_bs + (size_t)d * HW + pix, pd * (s_dd[q][d] - spf), a.out_dtype);
      }
    }
  }
}

}  // namespace lss
