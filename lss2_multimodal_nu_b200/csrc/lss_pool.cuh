// K3 interval detection, lift staging, K4/K4a pooling forward, K5/K5a backward.
#pragma once

#include "lss_common.cuh"

namespace lss {

// --------------------------------------------------------------------------
// K3: runs of equal ranks in the sorted order.  Replaces the boundary mask of
// QuickCumsum.forward (reference src/tools.py:196-197): `last` marks the last
// point of every run.  Heads/tails also record the run's [start, end) in a
// dense table indexed by OUTPUT cell, which is what lets K4 own every output
// voxel (zeros included) without a separate fill + scatter.
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
  // control words of the sort to wipe for the next call (fused plan path)
  uint32_t* wipe;
  long long wipe_words;
};

__global__ void __launch_bounds__(256)
intervals_kernel(IntervalArgs a) {
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long w = first; w < a.wipe_words; w += stride) a.wipe[w] = 0u;
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
// Lift staging: (BN, R, HW) -> (BN*HW, R) for R = D (depth) and R = C (context)
// through a padded 32x32 shared-memory tile; reads and writes are coalesced.
// grid = (ceil(HW/32), ceil(max(D,C)/32), BN * 2)  [z even: depth, odd: feat]
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lift_stage_kernel(const float* __restrict__ depth, const float* __restrict__ feat, int D, int C,
                  int HW, float* __restrict__ depth_t, float* __restrict__ feat_t) {
  __shared__ float tile[32][33];
  const int which = blockIdx.z & 1;
  const int bn = blockIdx.z >> 1;
  const int R = which ? C : D;
  const float* src = (which ? feat : depth) + (size_t)bn * R * HW;
  float* dst = (which ? feat_t : depth_t) + (size_t)bn * HW * R;
  const int r0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  if (r0 >= R) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + k * 8, p = p0 + tx;
    tile[ty + k * 8][tx] = (r < R && p < HW) ? src[(size_t)r * HW + p] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = p0 + ty + k * 8, r = r0 + tx;
    if (p < HW && r < R) dst[(size_t)p * R + r] = tile[tx][ty + k * 8];
  }
}

// --------------------------------------------------------------------------
// K4 / K4a forward, channels-innermost BEV.
//
// The BEV map is addressed as (cell, G) float4 vectors, cell = ((b*X + x)*Y + y)*Z + z
// and G = C/4, so one voxel is one contiguous line.  Every output element is
// written exactly once (this replaces torch.zeros + index_put + cat of
// reference src/model_baseline.py:120-124) by two kinds of warps that run side
// by side in every CTA:
//   * FILL warps stream zeros into the empty voxels (about 75 % of the map at
//     the headline config), one contiguous 512-byte line per warp store;
//   * REDUCE warps do the warp-level segmented reduction over the sorted point
//     list: kLanes lanes (a power of two >= G) own one voxel interval at a time,
//     each lane one float4 of channels.  A group loads kLanes consecutive sorted
//     points cooperatively (point id, output cell, depth), broadcasts them with
//     shuffles and walks them in order: a change of cell closes the running sum
//     (one 16*G-byte store) and opens the next.  A group owns the intervals that
//     START inside its chunk and follows the last one past the chunk end.
// The walk order is the sort order, so per-voxel sums are bit-reproducible.
//   kFused:  acc += depth_t[pixel*D + d] * feat_t[pixel, :]   (K4: the frustum
//            tensor of src/modules.py:84 is never formed)
//   !kFused: acc += x[point, :]                                (K4a)
// --------------------------------------------------------------------------
struct PoolFwdArgs {
  const float* depth_t;           // (BN*HW, D)        fused
  const float4* feat_t;           // (BN*HW, G)        fused
  const float4* x;                // (P, G)            dense
  const int32_t* sorted_points;   // (P) first K valid
  const int32_t* sorted_cells;    // (P) output cell of each sorted point, first K valid
  const int32_t* counts;          // {K, V}
  const int2* cell_range;         // (n_cells) start >= end: empty
  float4* bev;
  uint32_t n_cells;
  int G, D, HW;
  int fill_warps;                 // warps per CTA that zero-fill (the rest reduce)
  FastDiv div_g, div_dhw, div_hw;
};

constexpr int kPoolThreads = 256;
constexpr int kPoolWarps = kPoolThreads / 32;

template <bool kFused, int kLanes>
__global__ void __launch_bounds__(kPoolThreads)
pool_fwd_nhwc_kernel(PoolFwdArgs a) {
  constexpr int kGroups = 32 / kLanes;
  constexpr uint32_t kLaneBits = (kLanes == 32) ? 0xffffffffu : ((1u << kLanes) - 1u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // ---- FILL: zeros into empty voxels (dedicated warps; pure stores that overlap the gathers) ----
  // One coalesced load fetches the intervals of 32 consecutive cells (the next block of 32 is
  // prefetched before the stores go out); the emptiness bits are shared with a ballot and the
  // cells' 32*G float4 are covered by G full-warp stores.
  if (warp < a.fill_warps) {
    if (warp == 0 && lane == 0) phase_stamp_any(2, 0);
    const uint32_t n_fill = gridDim.x * static_cast<uint32_t>(a.fill_warps);
    uint32_t c0 = (blockIdx.x * a.fill_warps + warp) * 32u;
    int2 r = (c0 + lane < a.n_cells) ? __ldg(a.cell_range + c0 + lane) : make_int2(0, 1);
    while (c0 < a.n_cells) {
      const uint32_t c1 = c0 + n_fill * 32u;
      const int2 rn = (c1 + lane < a.n_cells) ? __ldg(a.cell_range + c1 + lane) : make_int2(0, 1);
      const uint32_t empty = __ballot_sync(0xffffffffu, r.x >= r.y);
      if (empty) {
        float4* dst = a.bev + (size_t)c0 * a.G;
        for (int k = 0; k < a.G; ++k) {
          const uint32_t e = k * 32u + lane;
          const uint32_t cl = a.div_g.div(e);
          if ((empty >> cl) & 1u) st_stream_f4(dst + e, zero4);
        }
      }
      r = rn;
      c0 = c1;
    }
    if (warp == 0 && lane == 0) phase_stamp_any(2, 1);
    return;
  }

  // ---- REDUCE: segmented sums over the sorted points --------------------------
  // A group of kLanes lanes owns the intervals that START inside its chunk of kLanes sorted
  // points.  Per pass: one cooperative load (point id, output cell, depth); the decoded records
  // {cell, feature-row offset, depth} go to shared memory and two group ballots (valid points,
  // interval heads) drive the walk, which reads one broadcast LDS.128 per point and has no
  // per-point bounds logic.  The tail of the last interval is followed into the next chunk in
  // steps of 4 points, where it only accumulates up to the first foreign head.  The kernel is
  // latency-bound (dependent gathers), so it is kept lean in registers: 8 CTAs stay resident
  // per SM and supply the parallelism.
  __shared__ int4 s_rec[kPoolWarps][32];
  const int rw = warp - a.fill_warps, n_rw = kPoolWarps - a.fill_warps;
  const int grp = lane / kLanes, sub = lane % kLanes;
  const uint32_t gmask = kLaneBits << (grp * kLanes);
  const int K = __ldg(a.counts);
  const int n_chunks = (K + kLanes - 1) / kLanes;
  const int slot = (blockIdx.x * n_rw + rw) * kGroups + grp;
  const int n_slots = gridDim.x * n_rw * kGroups;
  const bool lane_active = sub < a.G;  // kLanes may exceed G (e.g. C = 80: 20 of 32 lanes)
  int4* rec = &s_rec[warp][grp * kLanes];
  const float4* src_lane = (kFused ? a.feat_t : a.x) + sub;   // this lane's float4 column of a row
  float4* bev_lane = a.bev + sub;
  const uint32_t G = static_cast<uint32_t>(a.G);

  for (int chunk = slot; chunk < n_chunks; chunk += n_slots) {
    int base = chunk * kLanes;
    int width = kLanes;      // points loaded per pass: a full chunk first, then 4 at a time
    int cur_cell = -1;
    float4 acc = zero4;
    bool first = true;
    while (true) {
      // cooperative load of `width` consecutive sorted points
      const int i = base + sub;
      const bool valid = (sub < width) && (i < K);
      const int32_t pt = valid ? __ldg(a.sorted_points + i) : 0;
      const int32_t cell = valid ? __ldg(a.sorted_cells + i) : -1;
      int32_t prev = __shfl_up_sync(gmask, cell, 1, kLanes);
      if (sub == 0) prev = (valid && i > 0) ? __ldg(a.sorted_cells + i - 1) : -1;
      const bool head = valid && (cell != prev);
      uint32_t row;   // feature row (pixel, or point for the dense variant)
      float dv = 0.f;
      if (kFused) {
        uint32_t bn, rem, d, hw;
        a.div_dhw.divmod(static_cast<uint32_t>(pt), bn, rem);
        a.div_hw.divmod(rem, d, hw);
        row = bn * a.HW + hw;
        if (valid) dv = __ldg(a.depth_t + (size_t)row * a.D + d);
      } else {
        row = static_cast<uint32_t>(pt);
        dv = 1.f;
      }
      __syncwarp(gmask);   // the previous pass is done reading the records
      rec[sub] = make_int4(cell, static_cast<int>(row * G), __float_as_int(dv), 0);
      const uint32_t vbits = (__ballot_sync(gmask, valid) >> (grp * kLanes)) & kLaneBits;
      const uint32_t hbits = (__ballot_sync(gmask, head) >> (grp * kLanes)) & kLaneBits;
      __syncwarp(gmask);
      const int n_valid = __popc(vbits);   // valid points are a prefix of the pass
      int j_begin, j_end;
      if (first) {        // own chunk: start at the first head, run to the end of the data in it
        j_begin = hbits ? (__ffs(hbits) - 1) : n_valid;
        j_end = n_valid;
      } else {            // continuation: only the points before the first foreign head
        j_begin = 0;
        j_end = hbits ? (__ffs(hbits) - 1) : n_valid;
      }
#pragma unroll 4
      for (int j = j_begin; j < j_end; ++j) {
        const int4 r = rec[j];
        const float4 f = lane_active ? ldg_f4(src_lane + static_cast<uint32_t>(r.y)) : zero4;
        if (first && ((hbits >> j) & 1u)) {      // a new interval starts: close the running one
          if (cur_cell >= 0 && lane_active) st_stream_f4(bev_lane + static_cast<uint32_t>(cur_cell) * G, acc);
          acc = zero4;
          cur_cell = r.x;
        }
        const float d = __int_as_float(r.z);
        acc.x = fmaf(d, f.x, acc.x);
        acc.y = fmaf(d, f.y, acc.y);
        acc.z = fmaf(d, f.z, acc.z);
        acc.w = fmaf(d, f.w, acc.w);
      }
      // the open interval continues iff this pass was full and ended without meeting its end
      if (cur_cell < 0 || j_end < width) break;
      first = false;
      base += width;
      width = 4;
    }
    if (cur_cell >= 0 && lane_active) st_stream_f4(bev_lane + static_cast<uint32_t>(cur_cell) * G, acc);
  }
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
    if (cell >= 0) v = ldg_f4(dbev + (size_t)cell * G + chunk);
    dx[e] = v;
  }
}

// --------------------------------------------------------------------------
// K5 fused backward, channels-innermost dBEV.
//
// One CTA per feature-map row (bn, h): its fW pixels x D depth bins.  A warp
// owns one pixel at a time and keeps that pixel's context vector in registers;
// kLanes lanes (a power of two >= C/4) cooperate on one point, 32/kLanes points
// sit side by side in the warp and kUnroll such steps are issued back to back,
// so up to kUnroll*32/kLanes voxel-gradient lines are in flight per warp.
// For every kept point the voxel gradient g (C floats, one contiguous line of
// the channels-innermost dBEV) is gathered once and used twice:
//     d_depth[d] = <g, feat>      d_feat += depth[d] * g
// Only OCCUPIED voxels of dBEV are ever read.  <g, feat> has C terms of order
// one that cancel, so it is accumulated in float64 (B200 has the FP64 pipe) and
// the kUnroll partial dots of a lane are reduced together with a transposed
// butterfly (2*kUnroll shuffles instead of kUnroll*log2(kLanes)).  Results are
// staged in shared memory and written as whole (d, :) / (c, :) rows.  No
// atomics anywhere: bit-reproducible.
// --------------------------------------------------------------------------
struct PoolBwdArgs {
  const float4* dbev;       // (n_cells, G)
  const float* depth_t;     // (BN*HW, D)
  const float4* feat_t;     // (BN*HW, G)
  const int32_t* cells;     // (BN, D, fH, fW)
  float* ddepth;            // (BN, D, fH, fW)
  float* dfeat;             // (BN, C, fH, fW)
  int D, fH, fW, C, G;
};

template <int kLanes>
__global__ void __launch_bounds__(256, 3)
liftsplat_bwd_nhwc_kernel(PoolBwdArgs a) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr int kPts = 32 / kLanes;                 // points per warp step
  constexpr int kUnroll = kLanes >= 8 ? 8 : kLanes; // steps in flight
  constexpr int kRound = kPts * kUnroll;            // depth bins per round
  constexpr int kLog = kLanes == 32 ? 5 : kLanes == 16 ? 4 : kLanes == 8 ? 3 : 2;
  constexpr int kLogU = kUnroll == 8 ? 3 : 2;
  const int bn = blockIdx.x / a.fH, h = blockIdx.x % a.fH;
  const int HW = a.fH * a.fW;
  // depth bins are padded to a whole number of rounds ({-1, 0} = dropped point), so the
  // main loop needs no bounds checks at all
  const int Dpad = (a.D + kRound - 1) / kRound * kRound;
  int2* s_cd = reinterpret_cast<int2*>(s_raw);            // [Dpad][fW] {output cell, depth bits}
  float* s_dd = reinterpret_cast<float*>(s_cd + Dpad * a.fW);  // [Dpad][fW]
  float* s_df = s_dd + Dpad * a.fW;                       // [C][fW+1]
  const int dfs = a.fW + 1;
  for (int i = threadIdx.x; i < Dpad * a.fW; i += blockDim.x) {
    const int d = i / a.fW, w = i - d * a.fW;
    int2 v = make_int2(-1, 0);
    if (d < a.D) {
      v.x = __ldg(a.cells + ((size_t)(bn * a.D + d) * a.fH + h) * a.fW + w);
      v.y = __float_as_int(__ldg(a.depth_t + ((size_t)bn * HW + h * a.fW + w) * a.D + d));
    }
    s_cd[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int grp = lane / kLanes, sub = lane % kLanes;
  const bool lane_active = sub < a.G;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // which of the kUnroll dots this lane ends up holding after the transposed butterfly
  int my_u = 0;
#pragma unroll
  for (int k = 0; k < kLogU; ++k)
    if (sub & (kLanes >> (k + 1))) my_u += kUnroll >> (k + 1);
  const bool writer = (sub & ((kLanes >> kLogU) - 1)) == 0;
  const float4* gbase = a.dbev + sub;                 // this lane's float4 column of every voxel line
  const uint32_t G = static_cast<uint32_t>(a.G);
  const int step = kPts * a.fW;                       // shared-memory stride between unrolled steps

  for (int w = warp; w < a.fW; w += nwarps) {
    const uint32_t pix = static_cast<uint32_t>(bn * HW + h * a.fW + w);
    const float4 f = lane_active ? ldg_f4(a.feat_t + pix * G + sub) : zero4;
    const double fx = f.x, fy = f.y, fz = f.z, fw = f.w;
    float4 acc = zero4;
    const int2* cd = s_cd + grp * a.fW + w;
    float* dd = s_dd + (my_u * kPts + grp) * a.fW + w;
    for (int d0 = 0; d0 < Dpad; d0 += kRound, cd += kRound * a.fW, dd += kRound * a.fW) {
      float4 g[kUnroll];
      float dv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int2 v = cd[u * step];
        dv[u] = __int_as_float(v.y);
        const uint32_t off = static_cast<uint32_t>(v.x) * G;   // n_cells * G < 2^31 (checked by the host)
        g[u] = (v.x >= 0 && lane_active) ? ldg_f4(gbase + off) : zero4;
      }
      double dot[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        double t = static_cast<double>(g[u].x) * fx;
        t = fma(static_cast<double>(g[u].y), fy, t);
        t = fma(static_cast<double>(g[u].z), fz, t);
        t = fma(static_cast<double>(g[u].w), fw, t);
        dot[u] = t;
        acc.x = fmaf(dv[u], g[u].x, acc.x);
        acc.y = fmaf(dv[u], g[u].y, acc.y);
        acc.z = fmaf(dv[u], g[u].z, acc.z);
        acc.w = fmaf(dv[u], g[u].w, acc.w);
      }
      // transposed butterfly: halve the number of live values at every exchange
#pragma unroll
      for (int k = 0; k < kLog; ++k) {
        const int o = kLanes >> (k + 1);
        if (k < kLogU) {
          const int half = kUnroll >> (k + 1);
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
      if (writer) *dd = static_cast<float>(dot[0]);   // rows beyond D are padding
    }
    // fold the kPts point-groups of the warp together
#pragma unroll
    for (int o = kLanes; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (grp == 0 && lane_active) {
      float* df = s_df + (sub * 4) * dfs + w;
      df[0] = acc.x; df[dfs] = acc.y; df[2 * dfs] = acc.z; df[3 * dfs] = acc.w;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.D * a.fW; i += blockDim.x) {
    const int d = i / a.fW, w = i - d * a.fW;
    a.ddepth[((size_t)(bn * a.D + d) * a.fH + h) * a.fW + w] = s_dd[i];
  }
  for (int i = threadIdx.x; i < a.C * a.fW; i += blockDim.x) {
    const int c = i / a.fW, w = i - c * a.fW;
    a.dfeat[((size_t)(bn * a.C + c) * a.fH + h) * a.fW + w] = s_df[c * dfs + w];
  }
}

}  // namespace lss
