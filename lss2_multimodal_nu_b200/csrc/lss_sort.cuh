// K2: stable least-significant-digit radix sort of voxel ranks (int32 keys,
// int32 point-index payload).  Replaces ranks.argsort() + the gathers of
// reference src/model_baseline.py:110-111.
//
// Only the ceil(log2(n_cells+1)) significant key bits are sorted, in passes of
// at most 11 bits (19 bits -> 10+9 at the headline config instead of the 64-bit
// keys torch sorts).  Each pass is ONE kernel ("onesweep"): a tile of 4096 keys
// is ranked inside the CTA with warp match_any multi-split (stable by
// construction: warp-striped loads keep ascending point order), tiles chain
// their per-digit counts through a decoupled look-back on a status array, and
// keys + payloads are scattered straight to their final slot of the pass (tiles
// take tickets, so every tile a CTA waits for has started: no co-residency
// assumption).  The
// digit histograms of all passes come from whoever produced the keys (K1/K1'
// accumulate them on the fly) or from histogram_kernel below.
#pragma once

#include "lss_common.cuh"
#include "lss_geometry.cuh"  // SortDigits, kMaxHistBins

namespace lss {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per CTA
constexpr int kSortMaxBits = 11;
constexpr int kSortMaxBins = 1 << kSortMaxBits;  // == kMaxHistBins
constexpr int kSortBinsPerThread = kSortMaxBins / kSortThreads;  // 8
constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValueMask = ~kFlagMask;

static_assert(kSortMaxBins == kMaxHistBins, "histogram stride mismatch");

struct SortPlan {
  int key_bits;
  int passes;
  int bits[4];
  int shift[4];
  long long tiles;
  // workspace layout (byte offsets)
  size_t off_tmp_keys, off_tmp_vals, off_control, off_hist, off_ticket, off_status[4];
  size_t control_bytes, total_bytes;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline SortPlan make_sort_plan(long long P, int32_t n_cells) {
  SortPlan s;
  memset(&s, 0, sizeof(s));
  int kb = 0;
  while ((1ll << kb) <= (long long)n_cells) ++kb;  // bit length of the sentinel rank
  if (kb < 1) kb = 1;
  s.key_bits = kb;
  s.tiles = (P + kSortTile - 1) / kSortTile;
  if (s.tiles < 1) s.tiles = 1;
  s.passes = (kb + kSortMaxBits - 1) / kSortMaxBits;
  const int base = kb / s.passes, rem = kb % s.passes;
  int sh = 0;
  for (int i = 0; i < s.passes; ++i) {
    s.bits[i] = base + (i < rem ? 1 : 0);
    s.shift[i] = sh;
    sh += s.bits[i];
  }
  size_t off = 0;
  s.off_tmp_keys = off; off += align_up((size_t)P * 4, 256);
  s.off_tmp_vals = off; off += align_up((size_t)P * 4, 256);
  s.off_control = off;
  s.off_hist = off; off += (size_t)4 * kSortMaxBins * 4;
  s.off_ticket = off; off += 256;
  for (int i = 0; i < s.passes; ++i) {
    s.off_status[i] = off;
    off += align_up((size_t)s.tiles * (size_t)(1 << s.bits[i]) * 4, 256);
  }
  s.control_bytes = off - s.off_control;
  s.total_bytes = off;
  return s;
}

inline SortDigits sort_digits(const SortPlan& s, void* ws) {
  SortDigits d;
  d.hist = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + s.off_hist);
  d.passes = s.passes;
  for (int i = 0; i < 4; ++i) { d.bits[i] = s.bits[i]; d.shift[i] = s.shift[i]; }
  d.stride = kSortMaxBins;
  return d;
}

// digit histograms of every pass (used when the keys were produced elsewhere)
__global__ void __launch_bounds__(256)
histogram_kernel(const int32_t* __restrict__ keys, long long P, SortDigits sd) {
  extern __shared__ uint32_t s_hist[];
  for (int i = threadIdx.x; i < sd.passes * sd.stride; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P;
       p += (long long)gridDim.x * blockDim.x) {
    const uint32_t k = static_cast<uint32_t>(keys[p]);
#pragma unroll 1
    for (int ps = 0; ps < sd.passes; ++ps)
      atomicAdd(&s_hist[ps * sd.stride + ((k >> sd.shift[ps]) & ((1u << sd.bits[ps]) - 1u))], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < sd.passes * sd.stride; i += blockDim.x) {
    const uint32_t v = s_hist[i];
    if (v) atomicAdd(&sd.hist[i], v);
  }
}

struct SortPassArgs {
  const int32_t* keys_in;
  const int32_t* vals_in;  // null on the first pass: payload = point index
  int32_t* keys_out;
  int32_t* vals_out;
  long long P;
  int shift, bits;
  const uint32_t* hist;  // [nbins] global digit counts of this pass
  uint32_t* status;      // [tiles][nbins], zero on entry
  uint32_t* ticket;      // zero on entry
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  return *reinterpret_cast<const volatile uint32_t*>(p);
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  *reinterpret_cast<volatile uint32_t*>(p) = v;
}

__global__ void __launch_bounds__(kSortThreads)
radix_pass_kernel(SortPassArgs a) {
  // per-warp digit counters; column nbins collects out-of-range lanes
  __shared__ uint16_t s_wh[kSortWarps][kSortMaxBins + 2];
  __shared__ uint32_t s_base[kSortMaxBins];
  __shared__ uint32_t s_warp_tot[kSortWarps];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbins = 1 << a.bits;
  const uint32_t mask = static_cast<uint32_t>(nbins - 1);

  if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);  // tiles are claimed in launch order
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(&s_wh[0][0]);
    constexpr int kWords = kSortWarps * (kSortMaxBins + 2) / 2;
    for (int i = tid; i < kWords; i += kSortThreads) z[i] = 0;
  }
  // ---- exclusive scan of the global digit counts -> s_base ---------------
  {
    const int per = nbins >= kSortThreads ? nbins / kSortThreads : 1;
    uint32_t local[kSortBinsPerThread];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kSortBinsPerThread; ++k) {
      const int bin = tid * per + k;
      local[k] = (k < per && bin < nbins) ? a.hist[bin] : 0u;
      sum += local[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_warp_tot[w];
    uint32_t run = woff + incl - sum;
#pragma unroll
    for (int k = 0; k < kSortBinsPerThread; ++k) {
      const int bin = tid * per + k;
      if (k < per && bin < nbins) s_base[bin] = run;
      run += local[k];
    }
  }
  __syncthreads();
  const long long tile = s_tile;

  // ---- load keys (warp-striped) and rank them inside the warp ------------
  int32_t key[kSortItems];
  uint16_t offs[kSortItems];
  const long long warp_base = tile * kSortTile + (long long)warp * (32 * kSortItems);
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const long long idx = warp_base + j * 32 + lane;
    key[j] = (idx < a.P) ? a.keys_in[idx] : 0;
  }
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const long long idx = warp_base + j * 32 + lane;
    const uint32_t digit = (idx < a.P) ? ((static_cast<uint32_t>(key[j]) >> a.shift) & mask)
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

  // ---- per digit: scan over warps, publish, look back --------------------
  uint32_t count[kSortBinsPerThread], excl[kSortBinsPerThread];
  long long look[kSortBinsPerThread];
  bool done[kSortBinsPerThread];
  uint32_t* my_status = a.status + tile * nbins;
#pragma unroll
  for (int k = 0; k < kSortBinsPerThread; ++k) {
    const int bin = tid + k * kSortThreads;
    count[k] = 0; excl[k] = 0; look[k] = tile - 1;
    done[k] = (bin >= nbins) || (tile == 0);
    if (bin < nbins) {
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t c = s_wh[w][bin];
        s_wh[w][bin] = static_cast<uint16_t>(run);
        run += c;
      }
      count[k] = run;
      st_volatile_u32(my_status + bin, (tile == 0 ? kFlagInclusive : kFlagAggregate) | run);
    }
  }
  bool pending = true;
  while (pending) {
    pending = false;
#pragma unroll
    for (int k = 0; k < kSortBinsPerThread; ++k) {
      if (!done[k]) {
        const int bin = tid + k * kSortThreads;
        const uint32_t v = ld_volatile_u32(a.status + look[k] * nbins + bin);
        if (v & kFlagMask) {
          excl[k] += v & kValueMask;
          if (v & kFlagInclusive) done[k] = true; else --look[k];
        }
        pending |= !done[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kSortBinsPerThread; ++k) {
    const int bin = tid + k * kSortThreads;
    if (bin < nbins) {
      if (tile != 0) st_volatile_u32(my_status + bin, kFlagInclusive | (excl[k] + count[k]));
      s_base[bin] += excl[k];
    }
  }
  __syncthreads();

  // ---- scatter ------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const long long idx = warp_base + j * 32 + lane;
    if (idx < a.P) {
      const uint32_t digit = (static_cast<uint32_t>(key[j]) >> a.shift) & mask;
      const uint32_t dst = s_base[digit] + s_wh[warp][digit] + offs[j];
      a.keys_out[dst] = key[j];
      a.vals_out[dst] = a.vals_in ? a.vals_in[idx] : static_cast<int32_t>(idx);
    }
  }
}

// Enqueue the passes.  The control region (hist, tickets, status) must hold the
// digit histograms and zeros elsewhere.
inline int run_sort_passes(const SortPlan& s, const int32_t* keys, int32_t* out_keys,
                           int32_t* out_vals, long long P, void* ws, cudaStream_t st) {
  char* w = static_cast<char*>(ws);
  int32_t* tmp_k = reinterpret_cast<int32_t*>(w + s.off_tmp_keys);
  int32_t* tmp_v = reinterpret_cast<int32_t*>(w + s.off_tmp_vals);
  uint32_t* hist = reinterpret_cast<uint32_t*>(w + s.off_hist);
  uint32_t* ticket = reinterpret_cast<uint32_t*>(w + s.off_ticket);
  // ping-pong so that the last pass lands in (out_keys, out_vals)
  const int32_t* src_k = keys;
  const int32_t* src_v = nullptr;
  for (int i = 0; i < s.passes; ++i) {
    const bool to_out = ((s.passes - 1 - i) % 2) == 0;
    SortPassArgs a;
    a.keys_in = src_k; a.vals_in = src_v;
    a.keys_out = to_out ? out_keys : tmp_k;
    a.vals_out = to_out ? out_vals : tmp_v;
    a.P = P; a.shift = s.shift[i]; a.bits = s.bits[i];
    a.hist = hist + (size_t)i * kSortMaxBins;
    a.status = reinterpret_cast<uint32_t*>(w + s.off_status[i]);
    a.ticket = ticket + i * 4;
    radix_pass_kernel<<<(unsigned)s.tiles, kSortThreads, 0, st>>>(a);
    LSS_LAUNCH_CHECK("radix_pass_kernel");
    src_k = a.keys_out; src_v = a.vals_out;
  }
  return LSS_OK;
}

}  // namespace lss
