// Shared helpers for the Lift-Splat sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <math.h>
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
  float rdx[3];  // 1/dx where dx is a power of two (then x / dx == x * rdx bit for bit), else 0
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
    {
      int e = 0;
      const float m = frexpf(g->dx[i], &e);
      out->rdx[i] = (m == 0.5f && e > -120 && e < 120) ? ldexpf(1.0f, 1 - e) : 0.0f;
    }
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

// ---- programmatic dependent launch (build with -DLSS_PDL=1) -------------------------------------
// The kernels of one step form a chain on one stream.  Launched with the programmatic-serialization
// attribute, a kernel may be SCHEDULED while its predecessor still runs; pdl_wait() then holds it until
// the predecessor has completed and its writes are visible, so what overlaps is launch latency, block
// scheduling and the prologue before the wait.  Every kernel of the chain calls pdl_wait() before its
// first global-memory access (reads and writes: the predecessor may still be reading what this kernel
// overwrites) and pdl_trigger() right after, which lets ITS successor be scheduled once all of its own
// blocks have started.  Without the attribute both are no-ops.
#ifndef LSS_PDL
#define LSS_PDL 5
#endif
__device__ __forceinline__ void pdl_wait() {
#if LSS_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// LSS_PDL is a bit mask of the kernels launched as programmatic dependents
constexpr int kPdlPlan = 1, kPdlFwd = 2, kPdlBwd = 4, kPdlStage = 8;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(int which, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (LSS_PDL & which) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace lss
