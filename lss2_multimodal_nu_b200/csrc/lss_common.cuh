// Shared helpers for the Lift-Splat sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "lss_b200.h"

namespace lss {

// ---- error plumbing -------------------------------------------------------
char* cuda_error_buffer();  // thread-local, defined in lss_abi.cu

inline int record_cuda_error(cudaError_t e, const char* where) {
  snprintf(cuda_error_buffer(), 256, "%s: %s", where, cudaGetErrorString(e));
  return LSS_ERR_CUDA;
}

#define LSS_LAUNCH_CHECK(where)                                     \
  do {                                                              \
    cudaError_t e__ = cudaGetLastError();                           \
    if (e__ != cudaSuccess) return lss::record_cuda_error(e__, where); \
  } while (0)

#define LSS_CUDA_TRY(expr, where)                                   \
  do {                                                              \
    cudaError_t e__ = (expr);                                       \
    if (e__ != cudaSuccess) return lss::record_cuda_error(e__, where); \
  } while (0)

#define LSS_REQUIRE(cond, status) \
  do {                            \
    if (!(cond)) return (status); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- division by a runtime constant ---------------------------------------
// q = n / d for 0 <= n < 2^31, 1 <= d < 2^31:  q = (n * m) >> s with
// s = 31 + ceil(log2 d), m = floor(2^s / d) + 1 (fits 32 bits).
struct FastDiv {
  uint32_t m, s, d;
  FastDiv() : m(0), s(0), d(1) {}
  explicit FastDiv(uint32_t div) : d(div) {
    uint32_t L = 0;
    while ((1ull << L) < div) ++L;
    s = 31 + L;
    m = static_cast<uint32_t>(((1ull << s) / div) + 1ull);
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    return static_cast<uint32_t>((static_cast<unsigned long long>(n) * m) >> s);
#else
    return static_cast<uint32_t>((static_cast<uint64_t>(n) * m) >> s);
#endif
  }
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = div(n);
    r = n - q * d;
  }
};

// ---- grid constants as kernels see them -----------------------------------
struct GridDev {
  float off[3];  // bx - dx/2   (reference src/model_baseline.py:92)
  float dx[3];
  float nxf[3];
  int32_t nx[3];
  int32_t B;
  int32_t n_cells;  // nx0*nx1*nx2*B, also the sentinel rank
};

inline int make_grid(const LssGrid* g, int32_t B, GridDev* out) {
  if (!g) return LSS_ERR_NULL_POINTER;
  if (B <= 0) return LSS_ERR_BAD_DIMENSION;
  int64_t cells = B;
  for (int i = 0; i < 3; ++i) {
    if (g->nx[i] <= 0 || g->nx[i] >= (1 << 24)) return LSS_ERR_BAD_DIMENSION;
    cells *= g->nx[i];
    // the same two float32 operations torch performs: (dx / 2.) then (bx - .)
    volatile float half = g->dx[i] / 2.0f;
    volatile float off = g->bx[i] - half;
    out->off[i] = off;
    out->dx[i] = g->dx[i];
    out->nx[i] = g->nx[i];
    out->nxf[i] = static_cast<float>(g->nx[i]);
  }
  if (cells >= 0x7fffffffLL) return LSS_ERR_BAD_DIMENSION;
  out->B = B;
  out->n_cells = static_cast<int32_t>(cells);
  return LSS_OK;
}

// ---- the plan's sort key ----------------------------------------------------
// Output cells are sorted in TILE-major order: the BEV plane of every sample is cut into T x T
// tiles (T = kKeyTile) and key = ((((b*XT + xt)*YT + yt)*T + xi)*T + yi)*Z + z.  Points that are
// neighbours in the sorted list then fall into the same few square metres of the map, i.e. they
// come from the same few camera rays, and the feature rows they gather are re-used out of L1.
// The key is a bijection of the reference's rank on the cells that exist (n_keys >= n_cells:
// when X or Y is no multiple of T the last tiles hold keys without a cell, which never occur).
constexpr int kKeyTile = 8;

struct KeyMap {
  int32_t X, Y, Z, XT, YT, n_keys;
  FastDiv div_z, div_y, div_x;        // cell -> (b, x, y, z)
  FastDiv div_tz, div_yt, div_xt;     // key  -> (b, xt, yt, xi, yi, z)
  __host__ __device__ __forceinline__ uint32_t key_of_cell(uint32_t cell) const {
    uint32_t t, z, t2, y, b, x;
    div_z.divmod(cell, t, z);
    div_y.divmod(t, t2, y);
    div_x.divmod(t2, b, x);
    const uint32_t xt = x / kKeyTile, xi = x % kKeyTile, yt = y / kKeyTile, yi = y % kKeyTile;
    return ((((b * XT + xt) * YT + yt) * kKeyTile + xi) * kKeyTile + yi) * Z + z;
  }
  // output cell of a key, or -1 when the key has no cell (x >= X or y >= Y in the last tiles)
  __host__ __device__ __forceinline__ int32_t cell_of_key(uint32_t key) const {
    uint32_t t, z, tile, in, t2, yt, b, xt;
    div_z.divmod(key, t, z);
    div_tz.divmod(t, tile, in);                        // in = xi*T + yi
    div_yt.divmod(tile, t2, yt);
    div_xt.divmod(t2, b, xt);
    const uint32_t x = xt * kKeyTile + in / kKeyTile, y = yt * kKeyTile + in % kKeyTile;
    if (x >= static_cast<uint32_t>(X) || y >= static_cast<uint32_t>(Y)) return -1;
    return static_cast<int32_t>(((b * X + x) * Y + y) * Z + z);
  }
};

inline int make_keymap(const GridDev& g, KeyMap* k) {
  k->X = g.nx[0]; k->Y = g.nx[1]; k->Z = g.nx[2];
  k->XT = (g.nx[0] + kKeyTile - 1) / kKeyTile;
  k->YT = (g.nx[1] + kKeyTile - 1) / kKeyTile;
  const long long n = (long long)g.B * k->XT * k->YT * kKeyTile * kKeyTile * k->Z;
  if (n >= 0x7fffffffLL) return LSS_ERR_BAD_DIMENSION;
  k->n_keys = static_cast<int32_t>(n);
  k->div_z = FastDiv((uint32_t)k->Z); k->div_y = FastDiv((uint32_t)k->Y); k->div_x = FastDiv((uint32_t)k->X);
  k->div_tz = FastDiv((uint32_t)(kKeyTile * kKeyTile)); k->div_yt = FastDiv((uint32_t)k->YT);
  k->div_xt = FastDiv((uint32_t)k->XT);
  return LSS_OK;
}

// ---- feature dtypes at the boundary (LssDtype): arithmetic is always float32 ----
__device__ __forceinline__ float load_as_float(const void* base, size_t i, int dtype) {
  if (dtype == LSS_F16) return __half2float(reinterpret_cast<const __half*>(base)[i]);
  if (dtype == LSS_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
  return reinterpret_cast<const float*>(base)[i];
}
__device__ __forceinline__ void store_from_float(void* base, size_t i, float v, int dtype) {
  if (dtype == LSS_F16) reinterpret_cast<__half*>(base)[i] = __float2half_rn(v);
  else if (dtype == LSS_BF16) reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(base)[i] = v;
}
inline bool valid_dtype(int d) { return d == LSS_F32 || d == LSS_F16 || d == LSS_BF16; }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Optional phase timestamps (build with -DLSS_PHASE_TIMING; tools/phase_timing.py reads them).
#ifdef LSS_PHASE_TIMING
__device__ unsigned long long g_phase_ts[3][4096 * 16];
__device__ __forceinline__ void phase_stamp_any(int kernel, int slot) {   // caller picks the thread
  if (blockIdx.x < 4096) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_phase_ts[kernel][blockIdx.x * 16 + slot] = t;
  }
}
__device__ __forceinline__ void phase_stamp(int kernel, int slot) {
  if (threadIdx.x == 0 && blockIdx.x < 4096) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_phase_ts[kernel][blockIdx.x * 16 + slot] = t;
  }
}
#else
__device__ __forceinline__ void phase_stamp_any(int, int) {}
__device__ __forceinline__ void phase_stamp(int, int) {}
#endif

__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

// 128-bit read-only load that the compiler may not sink towards its first use: a batch of these
// stays a batch, i.e. all of them are in flight before the first result is consumed.
__device__ __forceinline__ float4 ldg_f4_issue(const float4* p, bool pred) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "l"(p), "r"(static_cast<int>(pred)));
  return v;
}

// L2 residency control.  The forward streams 82 MB of BEV through a 126 MB L2 while it keeps
// re-reading ~9 MB of small tables (staged features, sorted points, intervals): the stream is
// marked evict-first and the tables evict-last, so the tables stay resident instead of being
// pushed out to HBM (where their reloads would queue behind the write-back traffic).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_f4_hint(float4* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
// streaming 128-bit store (no hint)
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) { __stcs(p, v); }

// predicated 128-bit read-only load with an L2 policy; like ldg_f4_issue it is not sunk
__device__ __forceinline__ float4 ldg_f4_issue_hint(const float4* p, bool pred, uint64_t pol) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %6;\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "l"(p), "r"(static_cast<int>(pred)), "l"(pol));
  return v;
}
__device__ __forceinline__ int32_t ldg_i32_hint(const int32_t* p, uint64_t pol) {
  int32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ldg_f32_hint(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}

}  // namespace lss
