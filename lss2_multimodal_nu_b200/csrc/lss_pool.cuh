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
lift_stage_kernel(const void* __restrict__ depth, long long depth_bs, const void* __restrict__ feat,
                  long long feat_bs, int dtype, int D, int C, int HW, float* __restrict__ depth_t,
                  float* __restrict__ feat_t) {
  __shared__ float tile[32][33];
  const int which = blockIdx.z & 1;
  const int bn = blockIdx.z >> 1;
  const int R = which ? C : D;
  if (!which && depth == nullptr) return;              // depth staged elsewhere (softmax variant)
  const void* src = which ? feat : depth;
  const size_t sbase = (size_t)bn * (which ? feat_bs : depth_bs);
  float* dst = (which ? feat_t : depth_t) + (size_t)bn * HW * R;
  const int r0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  if (r0 >= R) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + k * 8, p = p0 + tx;
    tile[ty + k * 8][tx] = (r < R && p < HW) ? load_as_float(src, sbase + (size_t)r * HW + p, dtype) : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = p0 + ty + k * 8, r = r0 + tx;
    if (p < HW && r < R) dst[(size_t)p * R + r] = tile[tx][ty + k * 8];
  }
}

// Producer fusion: the depth distribution is softmax over the D logit channels
// (reference src/modules.py:76-77, `x.softmax(dim=1)`), computed here while staging, so the
// probabilities are written once, pixel-major, and the (B*N, D, fH, fW) probability tensor of the
// reference never exists.  One CTA = 32 pixels x all D logits (tile in shared memory):
// p = exp(x - max) / sum, the expression torch evaluates.
constexpr int kSoftmaxMaxD = 128;
__global__ void __launch_bounds__(256)
lift_stage_softmax_kernel(const void* __restrict__ logits, long long logits_bs, int dtype, int D, int HW,
                          float* __restrict__ depth_t) {
  __shared__ float tile[kSoftmaxMaxD][33];
  __shared__ float s_inv[32];
  const int bn = blockIdx.y, p0 = blockIdx.x * 32;
  const size_t sbase = (size_t)bn * logits_bs;
  float* dst = depth_t + (size_t)bn * HW * D;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < D; r += 8) {
    const int p = p0 + tx;
    tile[r][tx] = (p < HW) ? load_as_float(logits, sbase + (size_t)r * HW + p, dtype) : 0.0f;
  }
  __syncthreads();
  if (ty == 0) {                                      // one lane per pixel: max, then the sum of exp
    float m = tile[0][tx];
    for (int r = 1; r < D; ++r) m = fmaxf(m, tile[r][tx]);
    float sum = 0.f;
    for (int r = 0; r < D; ++r) {
      const float e = expf(tile[r][tx] - m);
      tile[r][tx] = e;
      sum += e;
    }
    s_inv[tx] = sum;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {                  // pixel k of the tile, lanes along d: coalesced rows
    const int p = p0 + k;
    if (p >= HW) continue;
    const float sum = s_inv[k];
    for (int r = tx; r < D; r += 32) dst[(size_t)p * D + r] = tile[r][k] / sum;
  }
}

// --------------------------------------------------------------------------
// K4 / K4a forward, channels-innermost BEV.
//
// The BEV map is (cell, C) with cell = ((b*X + x)*Y + y)*Z + z, one voxel = one
// contiguous line, and the plan's point list is sorted by the cell's tile-major
// key (KeyMap), so key k owns sorted_points[key_start[k] .. key_start[k+1]) and
// neighbours in the list are neighbours on the map.  Every output element is
// written exactly once (this is the torch.zeros + index_put + cat of reference
// src/model_baseline.py:120-124) by two kinds of CTAs that run side by side:
//   * FILL CTAs (one per SM) stream zeros into the empty voxels (75 % of the map
//     at the headline config): a warp reads the 33 interval bounds of 32
//     consecutive keys with one coalesced load (the next block's are prefetched)
//     and every run of empty voxels inside a tile row -- a contiguous piece of
//     the map -- leaves as ONE bulk async store (TMA engine, cp.async.bulk
//     shared -> global from a 2 KB zero block), so the zero stream costs the SM
//     neither issue slots nor load/store queue entries;
//   * REDUCE CTAs do the warp-level segmented reduction over the sorted point
//     list.  A warp takes 32 consecutive sorted points -- the unit of work is the
//     POINT, so dense and sparse regions of the map cost the same -- and owns
//     the intervals that START among them; the last one is followed into the
//     next chunk.  One coalesced load brings point ids and cells, the lanes
//     decode them, gather the depths and stage {feature-row offset, depth, cell}
//     in shared memory; then the warp walks the records in order with
//     (kS - 1) * kU feature-row gathers in flight (software pipeline), every lane
//     owning kVec channels (C = 64: 32 lanes x 8 bytes = one 256-byte row per
//     load), closing the running sum at every interval head.  No cumsum, no
//     atomics, summation in sort order: bit-reproducible.
//   kFused:  acc += depth_t[pixel, d] * feat_t[pixel, :]   (K4: the frustum tensor
//            of src/modules.py:84 is never formed)
//   !kFused: acc += x[point, :]                             (K4a)
// --------------------------------------------------------------------------
struct PoolFwdArgs {
  const float* depth_t;           // (BN*HW, D)   fused
  const float* feat_t;            // (BN*HW, C)   fused
  const float* x;                 // (P, C)       dense
  const int32_t* sorted_points;   // (P) sorted by output cell, ascending point id inside a cell
  const int32_t* sorted_cells;    // (P) output cell of each sorted point, -1 beyond the K kept points
  const int32_t* cell_start;      // (n_keys + 1) interval bounds, indexed by key
  float* bev;                     // (n_cells, C)
  long long P;
  KeyMap keys;
  int fill_ctas;                  // CTAs [0, fill_ctas) zero-fill, the rest reduce
  int C, D, HW;
  FastDiv div_dhw, div_hw, div_g4;
};

constexpr int kPoolThreads = 256;
constexpr int kPoolWarps = kPoolThreads / 32;
constexpr int kPoolChunk = 32;    // sorted points per reduce warp
constexpr int kZeroBytes = 2048;  // zero block in shared memory: source of the bulk zero stores
constexpr int kPoolRec = 64;      // staged records per warp: the chunk + the tail of its last interval

template <int kVec> struct VecOf;
template <> struct VecOf<1> { using type = float; };
template <> struct VecOf<2> { using type = float2; };
template <> struct VecOf<4> { using type = float4; };

template <int kVec> __device__ __forceinline__ void vec_zero(typename VecOf<kVec>::type& v);
template <> __device__ __forceinline__ void vec_zero<1>(float& v) { v = 0.f; }
template <> __device__ __forceinline__ void vec_zero<2>(float2& v) { v = make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ void vec_zero<4>(float4& v) { v = make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ void vec_fma(float d, const float& f, float& a) { a = fmaf(d, f, a); }
__device__ __forceinline__ void vec_fma(float d, const float2& f, float2& a) {
  a.x = fmaf(d, f.x, a.x); a.y = fmaf(d, f.y, a.y);
}
__device__ __forceinline__ void vec_fma(float d, const float4& f, float4& a) {
  a.x = fmaf(d, f.x, a.x); a.y = fmaf(d, f.y, a.y); a.z = fmaf(d, f.z, a.z); a.w = fmaf(d, f.w, a.w);
}

#ifndef LSS_FWD_MINB
#define LSS_FWD_MINB 4
#endif
#ifndef LSS_FWD_STAGES
#define LSS_FWD_STAGES 4
#endif
#ifndef LSS_FWD_STAGES4
#define LSS_FWD_STAGES4 2
#endif
#ifndef LSS_BWD_MINB
#define LSS_BWD_MINB 4
#endif
template <bool kFused, int kVec>
__global__ void __launch_bounds__(kPoolThreads, LSS_FWD_MINB)
pool_fwd_nhwc_kernel(PoolFwdArgs a) {
  using V = typename VecOf<kVec>::type;
  constexpr int kU = 4;                                      // gathers per pipeline stage
  constexpr int kS = kVec == 4 ? LSS_FWD_STAGES4 : LSS_FWD_STAGES;   // stages: (kS - 1) * kU gathers in flight
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) phase_stamp_any(2, warp * 2);

  // ======================= FILL: zeros into the empty voxels ==========================
  // A warp takes 32 consecutive KEYS (tile-major order, KeyMap): their interval bounds are one
  // coalesced load; every group of 8 keys is 8 consecutive voxels of one tile row, i.e. one
  // contiguous piece of the map, covered with 128-bit stores wherever the voxel is empty.
  if (static_cast<int>(blockIdx.x) < a.fill_ctas) {
    // the zeros come from a 2 KB block of shared memory and leave through the TMA engine (bulk
    // async stores, one instruction per contiguous piece), so they occupy neither the warps' issue
    // slots nor the load/store queues the REDUCE warps' gathers go through
    __shared__ __align__(128) float4 s_zero[kZeroBytes / 16];
    for (int e = threadIdx.x; e < kZeroBytes / 16; e += kPoolThreads) s_zero[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t zsrc = static_cast<uint32_t>(__cvta_generic_to_shared(s_zero));
    const int n_blocks = (a.keys.n_keys + 31) >> 5;
    const int stride = a.fill_ctas * kPoolWarps;
    const uint32_t line_bytes = static_cast<uint32_t>(a.C) * 4u;
    int blk = blockIdx.x * kPoolWarps + warp;
    if (blk >= n_blocks) return;
    auto bounds = [&](int bk, int& lo, int& hi) {
      const int k = (bk << 5) + lane;
      lo = 0; hi = 1;                                        // beyond the key space: not ours to write
      if (k < a.keys.n_keys) { lo = __ldg(a.cell_start + k); hi = __ldg(a.cell_start + k + 1); }
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
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst), "r"(zsrc), "r"(n) : "memory");
            dst += n; left -= n;
          }
        }
      }
      if (nxt >= n_blocks) break;
      blk = nxt; lo = nlo; hi = nhi;
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the zero block must outlive the reads
    if (lane == 0) phase_stamp_any(2, warp * 2 + 1);
    return;
  }

  // ======================= REDUCE: segmented sums over the sorted points ==============
  __shared__ uint4 s_off[kPoolWarps][kPoolRec / 4];          // feature-row offset (16-byte units)
  __shared__ float4 s_dep[kPoolWarps][kPoolRec / 4];         // depth probability
  __shared__ int s_cell[kPoolWarps][kPoolChunk];             // output cell (heads only lie in the chunk)
  const long long i0 = ((long long)(blockIdx.x - a.fill_ctas) * kPoolWarps + warp) * kPoolChunk;
  if (i0 >= a.P) return;
  const long long i = i0 + lane;
  // one round of loads: this chunk's and the next chunk's ids and cells, and the cell before the chunk
  int cell = -1, ncell = -1, pt = 0, npt = 0;
  if (i < a.P) cell = __ldg(a.sorted_cells + i);
  if (i + kPoolChunk < a.P) ncell = __ldg(a.sorted_cells + i + kPoolChunk);
  int prev = (lane == 0 && i0 > 0) ? __ldg(a.sorted_cells + i0 - 1) : -1;
  if (i < a.P) pt = __ldg(a.sorted_points + i);
  if (i + kPoolChunk < a.P) npt = __ldg(a.sorted_points + i + kPoolChunk);
  {
    const int up = __shfl_up_sync(0xffffffffu, cell, 1);
    if (lane > 0) prev = up;
  }
  const uint32_t hb = __ballot_sync(0xffffffffu, cell >= 0 && cell != prev);   // interval heads
  if (hb == 0u) { if (lane == 0) phase_stamp_any(2, warp * 2 + 1); return; }   // an earlier warp owns all of it
  const uint32_t vb = __ballot_sync(0xffffffffu, cell >= 0);  // kept points are a prefix of the chunk
  const int nv = __popc(vb);
  const int h0 = __ffs(hb) - 1;                              // first owned point
  const int last_cell = __shfl_sync(0xffffffffu, cell, nv - 1);
  // tail of the last interval inside the next chunk (a prefix of it)
  const uint32_t cont = __ballot_sync(0xffffffffu, nv == kPoolChunk && ncell == last_cell);
  const int tail = (cont == 0xffffffffu) ? 32 : __ffs(~cont) - 1;
  const int n_rec = nv - h0 + tail;                          // records to walk: [h0, nv) + tail

  const uint32_t nact = static_cast<uint32_t>(a.C / kVec);   // lanes that own channels
  const bool active = lane < nact;                           // idle lanes (C < 32 * kVec) shadow lane 0:
  const uint32_t vlane = active ? lane : 0u;                 // they gather valid data and never store
  const char* src = reinterpret_cast<const char*>(reinterpret_cast<const V*>(kFused ? a.feat_t : a.x) + vlane);
  const uint32_t row16 = static_cast<uint32_t>(a.C) >> 2;    // 16-byte units per feature row
  auto gather = [&](uint32_t off16) { return __ldg(reinterpret_cast<const V*>(src + ((size_t)off16 << 4))); };
  V* out = reinterpret_cast<V*>(a.bev) + vlane;
  uint32_t* offs = reinterpret_cast<uint32_t*>(s_off[warp]);
  float* deps = reinterpret_cast<float*>(s_dep[warp]);

  // ---- stage the records, shifted so that the first owned point is record 0 ----
  auto record = [&](int p, uint32_t& off16, float& dv) {
    if (kFused) {
      uint32_t bn, rem, d, hw;
      a.div_dhw.divmod(static_cast<uint32_t>(p), bn, rem);
      a.div_hw.divmod(rem, d, hw);
      const uint32_t row = bn * a.HW + hw;
      off16 = row * row16;
      dv = __ldg(a.depth_t + (size_t)row * a.D + d);
    } else {
      off16 = static_cast<uint32_t>(p) * row16;
      dv = 1.0f;
    }
  };
  {
    uint32_t o0 = 0, o1 = 0;
    float d0 = 0.f, d1 = 0.f;
    const bool own0 = lane >= h0 && lane < nv, own1 = lane < tail;
    if (own0) record(pt, o0, d0);
    if (own1) record(npt, o1, d1);
    if (own0) { offs[lane - h0] = o0; deps[lane - h0] = d0; }
    if (own1) { offs[nv - h0 + lane] = o1; deps[nv - h0 + lane] = d1; }
    s_cell[warp][lane] = cell;
  }
  __syncwarp();

  // ---- walk: software pipeline over groups of kU records ----
  const uint32_t heads = hb >> h0;                           // bit r: record r starts an interval (bit 0 set)
  int cur = s_cell[warp][h0];
  V acc;
  vec_zero<kVec>(acc);
  auto issue = [&](int r, V (&f)[kU]) {
    const uint4 o4 = s_off[warp][r >> 2];
    f[0] = gather(o4.x);
    f[1] = gather(o4.y);
    f[2] = gather(o4.z);
    f[3] = gather(o4.w);
  };
  auto consume = [&](int r, const V (&f)[kU]) {
    const float4 d4 = s_dep[warp][r >> 2];
    const float d[kU] = {d4.x, d4.y, d4.z, d4.w};
    const uint32_t hm = (r < 32) ? ((heads >> r) & 0xfu) : 0u;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (((hm >> u) & 1u) && (r + u) > 0) {                 // warp-uniform: a new interval starts
        if (active) out[(size_t)cur * nact] = acc;
        vec_zero<kVec>(acc);
        cur = s_cell[warp][h0 + r + u];
      }
      vec_fma(d[u], f[u], acc);
    }
  };
  const int n_full = n_rec & ~(kU - 1);
  if (n_full > 0) {
    V f[kS][kU];
#pragma unroll
    for (int st = 0; st < kS - 1; ++st)
      if (st * kU < n_full) issue(st * kU, f[st]);
    for (int r0 = 0; r0 < n_full; r0 += kS * kU) {
#pragma unroll
      for (int st = 0; st < kS; ++st) {
        const int r = r0 + st * kU;
        if (r < n_full) {
          if (r + (kS - 1) * kU < n_full) issue(r + (kS - 1) * kU, f[(st + kS - 1) % kS]);
          consume(r, f[st]);
        }
      }
    }
  }
  for (int r = n_full; r < n_rec; ++r) {                     // at most kU - 1 records
    const V f = gather(offs[r]);
    if (r < 32 && ((heads >> r) & 1u) && r > 0) {
      if (active) out[(size_t)cur * nact] = acc;
      vec_zero<kVec>(acc);
      cur = s_cell[warp][h0 + r];
    }
    vec_fma(deps[r], f, acc);
  }
  // ---- an interval longer than the look-ahead (adversarial inputs): follow it to its end ----
  if (tail == 32) {
    for (long long j = i0 + 2 * kPoolChunk; j < a.P; ++j) {
      if (__ldg(a.sorted_cells + j) != last_cell) break;
      uint32_t o;
      float dv;
      record(__ldg(a.sorted_points + j), o, dv);
      vec_fma(dv, gather(o), acc);
    }
  }
  if (active) out[(size_t)cur * nact] = acc;
  if (lane == 0) phase_stamp_any(2, warp * 2 + 1);
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
// A warp owns one pixel (bn, h, w): its context vector stays in registers while
// the warp walks the pixel's D depth bins.  kLanes lanes (a power of two >= C/4)
// cooperate on one point, 32/kLanes points sit side by side in the warp and
// kUnroll such steps are issued back to back, so kUnroll*32/kLanes voxel-gradient
// lines are in flight per warp.  For every kept point the voxel gradient g (one
// contiguous line of the channels-innermost dBEV) is gathered once and used twice:
//     d_depth[d] = <g, feat>      d_feat += depth[d] * g
// Only OCCUPIED voxels of dBEV are ever read.  <g, feat> has C terms of order one
// that cancel, so it is accumulated in float64 and the kUnroll partial dots of a
// lane are reduced together with a transposed butterfly (2*kUnroll shuffles
// instead of kUnroll*log2(kLanes)).  The warps of a CTA are the fH pixels of one
// image column (bn, w): along a camera ray they fall into the same BEV cells
// (Z is collapsed), so the column's voxel lines are fetched from L2 once and
// re-used out of L1.  There is no CTA-wide phase and no atomics: warps run
// independently and results are bit-reproducible.
// --------------------------------------------------------------------------
struct PoolBwdArgs {
  const float4* dbev;       // (n_cells, G)
  const float* depth_t;     // (BN*HW, D)
  const float4* feat_t;     // (BN*HW, G)
  const int32_t* cells;     // (BN, D, fH, fW)
  void* ddepth;             // (BN, D, fH, fW) with batch stride ddepth_bs: d_depth, or d_logits when softmax
  void* dfeat;              // (BN, C, fH, fW) with batch stride dfeat_bs
  long long ddepth_bs, dfeat_bs;
  int out_dtype;            // LssDtype of both outputs
  int softmax;              // 1: depth_t = softmax(logits); emit d_logits = p * (d_depth - sum_d p * d_depth)
  int D, fH, fW, C, G;
  int n_pix;                // BN * fH * fW
  FastDiv div_fh, div_fw;
};

#ifndef LSS_BWD_THREADS
#define LSS_BWD_THREADS 256
#endif
constexpr int kBwdThreads = LSS_BWD_THREADS;
constexpr int kBwdWarps = kBwdThreads / 32;
constexpr int kBwdChunk = 128;   // depth bins staged per warp at a time

// kGeneral = false: float32 gradients, no fused softmax (the plain K5); true: output dtype and the
// fused softmax backward are run-time options (kept out of the plain kernel's inner loop)
template <int kLanes, bool kGeneral>
__global__ void __launch_bounds__(kBwdThreads, LSS_BWD_MINB)
liftsplat_bwd_nhwc_kernel(PoolBwdArgs a) {
  constexpr int kPts = 32 / kLanes;                 // points per warp step
  constexpr int kUnroll = kLanes >= 8 ? 8 : kLanes; // steps in flight
  constexpr int kRound = kPts * kUnroll;            // depth bins per round
  constexpr int kLog = kLanes == 32 ? 5 : kLanes == 16 ? 4 : kLanes == 8 ? 3 : 2;
  constexpr int kLogU = kUnroll == 8 ? 3 : 2;
  static_assert(kBwdChunk % kRound == 0, "chunk must hold whole rounds");
  __shared__ int2 s_cd[kBwdWarps][kBwdChunk];       // {output cell, depth bits} of the staged bins
  __shared__ float s_dd[kGeneral ? kBwdWarps : 1][kGeneral ? kBwdChunk : 1];   // d_depth of the pixel (softmax backward)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pixel of this warp: h fastest, so a CTA is one image column (bn, w) when fH == 8
  const uint32_t q = blockIdx.x * kBwdWarps + warp;
  if (q >= static_cast<uint32_t>(a.n_pix)) return;
  uint32_t t, h, bn, w;
  a.div_fh.divmod(q, t, h);
  a.div_fw.divmod(t, bn, w);
  const int HW = a.fH * a.fW;
  const uint32_t pix = bn * HW + h * a.fW + w;
  const int grp = lane / kLanes, sub = lane % kLanes;
  const bool lane_active = sub < a.G;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // which of the kUnroll dots this lane ends up holding after the transposed butterfly
  int my_u = 0;
#pragma unroll
  for (int k = 0; k < kLogU; ++k)
    if (sub & (kLanes >> (k + 1))) my_u += kUnroll >> (k + 1);
  const bool writer = (sub & ((kLanes >> kLogU) - 1)) == 0;
  const uint32_t G = static_cast<uint32_t>(a.G);
  const float4* gbase = a.dbev + sub;                 // this lane's float4 column of every voxel line
  const float4 f = lane_active ? ldg_f4(a.feat_t + pix * G + sub) : zero4;
  const double fx = f.x, fy = f.y, fz = f.z, fw = f.w;
  float4 acc = zero4;
  int2* cd = s_cd[warp];
  const size_t col = (size_t)h * a.fW + w;            // offset of the pixel inside a (fH, fW) slice

  for (int dc = 0; dc < a.D; dc += kBwdChunk) {
    const int nd = min(kBwdChunk, a.D - dc);
    const int nd_pad = (nd + kRound - 1) / kRound * kRound;
    __syncwarp();
    for (int l = lane; l < nd_pad; l += 32) {
      int2 v = make_int2(-1, 0);                      // padding = dropped point
      if (l < nd) {
        const int d = dc + l;
        v.x = __ldg(a.cells + ((size_t)bn * a.D + d) * HW + col);
        v.y = __float_as_int(__ldg(a.depth_t + (size_t)pix * a.D + d));
      }
      cd[l] = v;
    }
    __syncwarp();
    for (int d0 = 0; d0 < nd_pad; d0 += kRound) {
      float4 g[kUnroll];
      float dv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int2 v = cd[d0 + u * kPts + grp];
        dv[u] = __int_as_float(v.y);
        const uint32_t off = static_cast<uint32_t>(v.x) * G;   // n_cells * G < 2^31 (checked by the host)
        g[u] = (v.x >= 0 && lane_active) ? ldg_f4(gbase + off) : zero4;
      }
      double dot[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        double s = static_cast<double>(g[u].x) * fx;
        s = fma(static_cast<double>(g[u].y), fy, s);
        s = fma(static_cast<double>(g[u].z), fz, s);
        s = fma(static_cast<double>(g[u].w), fw, s);
        dot[u] = s;
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
      const int d = dc + d0 + my_u * kPts + grp;
      if (writer && d < a.D) {
        const size_t o = (size_t)bn * a.ddepth_bs + (size_t)d * HW + col;
        if (!kGeneral) reinterpret_cast<float*>(a.ddepth)[o] = static_cast<float>(dot[0]);
        else if (a.softmax) s_dd[warp][d] = static_cast<float>(dot[0]);   // D <= kBwdChunk (host)
        else store_from_float(a.ddepth, o, static_cast<float>(dot[0]), a.out_dtype);
      }
    }
  }
  if (kGeneral && a.softmax) {
    // softmax backward (reference: autograd of x.softmax(dim=1), src/modules.py:77), fused:
    // d_logit[d] = p[d] * (d_depth[d] - sum_d' p[d'] * d_depth[d'])
    __syncwarp();
    double sp = 0.0;
    for (int l = lane; l < a.D; l += 32) sp += static_cast<double>(__int_as_float(cd[l].y)) * static_cast<double>(s_dd[warp][l]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sp += __shfl_xor_sync(0xffffffffu, sp, o);
    const float spf = static_cast<float>(sp);
    for (int l = lane; l < a.D; l += 32)
      store_from_float(a.ddepth, (size_t)bn * a.ddepth_bs + (size_t)l * HW + col,
                       __int_as_float(cd[l].y) * (s_dd[warp][l] - spf), a.out_dtype);
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
    const size_t df = (size_t)bn * a.dfeat_bs + (size_t)(sub * 4) * HW + col;
    if (!kGeneral) {
      float* o = reinterpret_cast<float*>(a.dfeat) + df;
      o[0] = acc.x; o[HW] = acc.y; o[2 * (size_t)HW] = acc.z; o[3 * (size_t)HW] = acc.w;
    } else {
      store_from_float(a.dfeat, df, acc.x, a.out_dtype);
      store_from_float(a.dfeat, df + HW, acc.y, a.out_dtype);
      store_from_float(a.dfeat, df + 2 * (size_t)HW, acc.z, a.out_dtype);
      store_from_float(a.dfeat, df + 3 * (size_t)HW, acc.w, a.out_dtype);
    }
  }
}

}  // namespace lss
