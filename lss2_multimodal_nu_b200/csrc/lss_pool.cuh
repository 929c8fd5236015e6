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
      if (head || tail) {
        // rank = ((x*Y + y)*Z + z)*B + b  ->  output cell ((b*X + x)*Y + y)*Z + z
        uint32_t t0, b, t1, z, x, y;
        a.div_b.divmod(static_cast<uint32_t>(r), t0, b);
        a.div_z.divmod(t0, t1, z);
        a.div_y.divmod(t1, x, y);
        const int32_t cell = ((static_cast<int32_t>(b) * a.g.nx[0] + static_cast<int32_t>(x)) * a.g.nx[1] +
                              static_cast<int32_t>(y)) * a.g.nx[2] + static_cast<int32_t>(z);
        if (head) a.cell_range[cell].x = static_cast<int>(i);
        if (tail) a.cell_range[cell].y = static_cast<int>(i + 1);
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
// The output is addressed as a flat array of float4: element e belongs to
// output cell e / G (G = C/4 vectors per cell) and holds channels 4*(e % G)..+3.
// Cells are numbered ((b*X + x)*Y + y)*Z + z, so consecutive elements are
// consecutive in memory: every warp store is one contiguous 512-byte line, and
// every output element -- empty voxels included -- is written exactly once
// (this replaces torch.zeros + index_put + cat of src/model_baseline.py:120-124).
// Each lane walks its cell's run of sorted points in ascending order, so the
// per-voxel sum order is fixed: results are bit-reproducible run to run.
//   kFused:  acc += depth_t[pixel*D + d] * feat_t[pixel, 4*chunk..]   (K4)
//   !kFused: acc += x[point, 4*chunk..]                                (K4a)
// --------------------------------------------------------------------------
struct PoolFwdArgs {
  const float* depth_t;        // (BN*HW, D)        fused
  const float4* feat_t;        // (BN*HW, G)        fused
  const float4* x;             // (P, G)            dense
  const int32_t* sorted_points;
  const int2* cell_range;
  float4* bev;
  long long n_elems;           // n_cells * G
  int G, D, HW;
  FastDiv div_g, div_dhw, div_hw;
};

template <bool kFused>
__global__ void __launch_bounds__(256)
pool_fwd_nhwc_kernel(PoolFwdArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < a.n_elems; e += stride) {
    uint32_t cell, chunk;
    a.div_g.divmod(static_cast<uint32_t>(e), cell, chunk);
    const int2 range = __ldg(a.cell_range + cell);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = range.x; i < range.y; i += 4) {
      const int n = min(4, range.y - i);
      int32_t pt[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) pt[k] = (k < n) ? __ldg(a.sorted_points + i + k) : 0;
      float dv[4];
      float4 f[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kFused) {
          uint32_t bn, rem, d, hw;
          a.div_dhw.divmod(static_cast<uint32_t>(pt[k]), bn, rem);
          a.div_hw.divmod(rem, d, hw);
          const uint32_t pix = bn * a.HW + hw;
          dv[k] = (k < n) ? __ldg(a.depth_t + (size_t)pix * a.D + d) : 0.f;
          f[k] = (k < n) ? ldg_f4(a.feat_t + (size_t)pix * a.G + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          dv[k] = (k < n) ? 1.f : 0.f;
          f[k] = (k < n) ? ldg_f4(a.x + (size_t)pt[k] * a.G + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc.x = fmaf(dv[k], f[k].x, acc.x);
        acc.y = fmaf(dv[k], f[k].y, acc.y);
        acc.z = fmaf(dv[k], f[k].z, acc.z);
        acc.w = fmaf(dv[k], f[k].w, acc.w);
      }
    }
    st_stream_f4(a.bev + e, acc);
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
// are in flight per step.  For every kept point the voxel gradient g (C floats,
// one contiguous line of the channels-innermost dBEV) is gathered once and used
// twice:   d_depth[d] = <g, feat>   (shuffle reduction over the point's lanes)
//          d_feat    += depth[d] * g (register accumulation over d)
// Only OCCUPIED voxels of dBEV are ever read.  Results are staged in shared
// memory and written as whole (d, :) / (c, :) rows.  No atomics: deterministic.
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

template <int kLanes, int kChunks>
__global__ void __launch_bounds__(256)
liftsplat_bwd_nhwc_kernel(PoolBwdArgs a) {
  extern __shared__ float s_mem[];
  constexpr int kPts = 32 / kLanes;  // points per warp step
  constexpr int kUnroll = 4;
  const int bn = blockIdx.x / a.fH, h = blockIdx.x % a.fH;
  const int HW = a.fH * a.fW;
  int32_t* s_cells = reinterpret_cast<int32_t*>(s_mem);  // [D][fW]
  float* s_dd = s_mem + a.D * a.fW;                       // [D][fW]
  float* s_df = s_dd + a.D * a.fW;                        // [C][fW+1]
  const int dfs = a.fW + 1;
  for (int i = threadIdx.x; i < a.D * a.fW; i += blockDim.x) {
    const int d = i / a.fW, w = i - d * a.fW;
    s_cells[i] = a.cells[((size_t)(bn * a.D + d) * a.fH + h) * a.fW + w];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int grp = lane / kLanes, sub = lane % kLanes;
  for (int w = warp; w < a.fW; w += nwarps) {
    const size_t pix = (size_t)bn * HW + h * a.fW + w;
    float4 f[kChunks], acc[kChunks];
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const int chunk = sub + k * kLanes;
      f[k] = (chunk < a.G) ? ldg_f4(a.feat_t + pix * a.G + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int d0 = 0; d0 < a.D; d0 += kPts * kUnroll) {
      float4 g[kUnroll][kChunks];
      float dv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int d = d0 + u * kPts + grp;
        const int32_t cell = (d < a.D) ? s_cells[d * a.fW + w] : -1;
        dv[u] = (d < a.D) ? __ldg(a.depth_t + pix * a.D + d) : 0.f;
#pragma unroll
        for (int k = 0; k < kChunks; ++k) {
          const int chunk = sub + k * kLanes;
          g[u][k] = (cell >= 0 && chunk < a.G) ? ldg_f4(a.dbev + (size_t)cell * a.G + chunk)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int d = d0 + u * kPts + grp;
        // <g, feat> has C terms of order one that cancel: accumulate it in float64 so the
        // result is the correctly rounded sum (abs 1e-6 parity bar); B200 has the FP64 pipe.
        double dot = 0.0;
#pragma unroll
        for (int k = 0; k < kChunks; ++k) {
          dot = fma((double)g[u][k].x, (double)f[k].x, dot);
          dot = fma((double)g[u][k].y, (double)f[k].y, dot);
          dot = fma((double)g[u][k].z, (double)f[k].z, dot);
          dot = fma((double)g[u][k].w, (double)f[k].w, dot);
          acc[k].x = fmaf(dv[u], g[u][k].x, acc[k].x);
          acc[k].y = fmaf(dv[u], g[u][k].y, acc[k].y);
          acc[k].z = fmaf(dv[u], g[u][k].z, acc[k].z);
          acc[k].w = fmaf(dv[u], g[u][k].w, acc[k].w);
        }
#pragma unroll
        for (int o = kLanes / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (sub == 0 && d < a.D) s_dd[d * a.fW + w] = static_cast<float>(dot);
      }
    }
    // fold the kPts point-groups of the warp together
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
#pragma unroll
      for (int o = kLanes; o < 32; o <<= 1) {
        acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o);
        acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
        acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, o);
        acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, o);
      }
      const int chunk = sub + k * kLanes;
      if (grp == 0 && chunk < a.G) {
        s_df[(chunk * 4 + 0) * dfs + w] = acc[k].x;
        s_df[(chunk * 4 + 1) * dfs + w] = acc[k].y;
        s_df[(chunk * 4 + 2) * dfs + w] = acc[k].z;
        s_df[(chunk * 4 + 3) * dfs + w] = acc[k].w;
      }
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
