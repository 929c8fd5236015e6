// K0 camera preparation, K1 quantise+rank, K1' fused frustum geometry -> rank.
//
// Index parity with the reference needs bit-exact float32 arithmetic, so every
// operation that feeds the voxel index is an explicit IEEE round-to-nearest
// intrinsic (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn / __fmaf_rn /
// __frcp_rn): nvcc can neither contract them into FMAs nor replace the divide
// by a reciprocal multiply.  The operation order is the one torch executes on
// the CPU for the reference's statements (oracle/lss_oracle.py has the probe
// results):
//   * 3x3 @ 3x1 and 3x3 @ 3x3 products: (m0*p0 + m1*p1) + m2*p2, no FMA
//   * torch.inverse: LAPACK LU with partial pivoting, see inverse3x3 below
//   * quantisation: (g - (bx - dx/2)) / dx, then truncation toward zero
#pragma once

#include "lss_common.cuh"

namespace lss {

// --------------------------------------------------------------------------
// 3x3 inverse, op for op the getrf + getrs sequence behind torch.inverse
// (reference src/model_baseline.py:60,66).  a and x are row-major.
// --------------------------------------------------------------------------
__device__ __forceinline__ void swap_rows3(float (*m)[3], int r, int p) {
  if (r == p) return;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t = m[r][j];
    m[r][j] = m[p][j];
    m[p][j] = t;
  }
}

__device__ inline void inverse3x3(const float* a, float* x) {
  float m[3][3], b[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      m[i][j] = a[i * 3 + j];
      b[i][j] = (i == j) ? 1.0f : 0.0f;
    }
  // column 0: first maximum of |.|
  int p0 = 0;
  float best = fabsf(m[0][0]);
  if (fabsf(m[1][0]) > best) { best = fabsf(m[1][0]); p0 = 1; }
  if (fabsf(m[2][0]) > best) { p0 = 2; }
  swap_rows3(m, 0, p0);
  swap_rows3(b, 0, p0);
  const float r0 = __frcp_rn(m[0][0]);
  m[1][0] = __fmul_rn(m[1][0], r0);
  m[2][0] = __fmul_rn(m[2][0], r0);
  m[1][1] = __fmaf_rn(-m[1][0], m[0][1], m[1][1]);
  m[2][1] = __fmaf_rn(-m[2][0], m[0][1], m[2][1]);
  // column 1
  const int p1 = (fabsf(m[2][1]) > fabsf(m[1][1])) ? 2 : 1;
  swap_rows3(m, 1, p1);
  swap_rows3(b, 1, p1);
  m[2][1] = __fdiv_rn(m[2][1], m[1][1]);
  m[1][2] = __fmaf_rn(-m[1][0], m[0][2], m[1][2]);
  m[2][2] = __fmaf_rn(-m[2][1], m[1][2], __fmaf_rn(-m[2][0], m[0][2], m[2][2]));

  const float l10 = m[1][0], l20 = m[2][0], l21 = m[2][1];
  const float u00 = m[0][0], u01 = m[0][1], u02 = m[0][2];
  const float u11 = m[1][1], u12 = m[1][2], u22 = m[2][2];
  const float rd0 = __frcp_rn(u00), rd1 = __frcp_rn(u11), rd2 = __frcp_rn(u22);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float y0 = b[0][c];
    const float y1 = __fsub_rn(b[1][c], __fmul_rn(l10, y0));
    const float y2 = __fsub_rn(__fsub_rn(b[2][c], __fmul_rn(l20, y0)), __fmul_rn(l21, y1));
    float x0, x1, x2;
    if (c < 2) {  // first two right-hand sides: scaled by reciprocals
      x2 = __fmul_rn(y2, rd2);
      x1 = __fmul_rn(__fmaf_rn(-u12, x2, y1), rd1);
      const float s = __fmaf_rn(u02, x2, __fmul_rn(u01, x1));
      x0 = __fmul_rn(__fsub_rn(y0, s), rd0);
    } else {  // last right-hand side: true divisions
      x2 = __fdiv_rn(y2, u22);
      x1 = __fdiv_rn(__fmaf_rn(-u12, x2, y1), u11);
      const float s = __fmaf_rn(u02, x2, __fmul_rn(u01, x1));
      x0 = __fdiv_rn(__fsub_rn(y0, s), u00);
    }
    x[0 * 3 + c] = x0;
    x[1 * 3 + c] = x1;
    x[2 * 3 + c] = x2;
  }
}

__device__ __forceinline__ float dot3_nofma(float a0, float a1, float a2, float b0, float b1,
                                            float b2) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// one camera: inv_post_rots = inverse(post_rots), combine = rots @ inverse(intrins)
__device__ inline void camera_prep_one(const float* rots, const float* intrins,
                                       const float* post_rots, float* inv_post_rots,
                                       float* combine) {
  float ii[9];
  inverse3x3(post_rots, inv_post_rots);
  inverse3x3(intrins, ii);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      combine[i * 3 + j] = dot3_nofma(rots[i * 3 + 0], rots[i * 3 + 1], rots[i * 3 + 2], ii[0 + j],
                                      ii[3 + j], ii[6 + j]);
}

__global__ void camera_prep_kernel(const float* __restrict__ rots,
                                   const float* __restrict__ intrins,
                                   const float* __restrict__ post_rots, int n_cams,
                                   float* __restrict__ inv_post_rots,
                                   float* __restrict__ combine) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cams) return;
  float r[9], k[9], pr[9], ipr[9], cmb[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    r[j] = rots[i * 9 + j];
    k[j] = intrins[i * 9 + j];
    pr[j] = post_rots[i * 9 + j];
  }
  camera_prep_one(r, k, pr, ipr, cmb);
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    inv_post_rots[i * 9 + j] = ipr[j];
    combine[i * 9 + j] = cmb[j];
  }
}

// --------------------------------------------------------------------------
// quantise one ego-frame point and emit everything derived from it
// --------------------------------------------------------------------------
struct PointOut {
  int32_t* coords;  // (P,3) or null
  uint8_t* kept;    // (P) or null
  int32_t* ranks;   // (P)
  int32_t* cells;   // (P) or null
};

// Digit histograms of the ranks for the radix sort, accumulated on the fly so
// the sort needs no pass of its own over the keys (hist == null: disabled).
struct SortDigits {
  uint32_t* hist;  // [passes][nbins_max] global, zero on entry
  int passes;
  int bits[4];
  int shift[4];
  int stride;  // nbins_max
};

__device__ __forceinline__ int32_t quantize_point_core(float gx, float gy, float gz, int b,
                                                       const GridDev& g, long long p,
                                                       const PointOut& out, int32_t* cell_ret = nullptr) {
  // ((geom - (bx - dx/2)) / dx).long()   reference src/model_baseline.py:92
  const float qx = __fdiv_rn(__fsub_rn(gx, g.off[0]), g.dx[0]);
  const float qy = __fdiv_rn(__fsub_rn(gy, g.off[1]), g.dx[1]);
  const float qz = __fdiv_rn(__fsub_rn(gz, g.off[2]), g.dx[2]);
  // .long() truncates toward zero, so (-1, 0) lands in voxel 0 and is KEPT
  // (SURVEY.md 7.3-2); NaN / inf fail every comparison and are dropped.
  const bool keep = (qx > -1.0f) && (qx < g.nxf[0]) && (qy > -1.0f) && (qy < g.nxf[1]) &&
                    (qz > -1.0f) && (qz < g.nxf[2]);
  const int ix = __float2int_rz(qx), iy = __float2int_rz(qy), iz = __float2int_rz(qz);
  if (out.coords) {
    out.coords[p * 3 + 0] = ix;
    out.coords[p * 3 + 1] = iy;
    out.coords[p * 3 + 2] = iz;
  }
  if (out.kept) out.kept[p] = keep ? 1 : 0;
  int32_t rank = g.n_cells, cell = -1;
  if (keep) {
    // x*(ny*nz*B) + y*(nz*B) + z*B + b     reference src/model_baseline.py:106-109
    rank = ((ix * g.nx[1] + iy) * g.nx[2] + iz) * g.B + b;
    cell = ((b * g.nx[0] + ix) * g.nx[1] + iy) * g.nx[2] + iz;
  }
  if (out.ranks) out.ranks[p] = rank;
  if (out.cells) out.cells[p] = cell;
  if (cell_ret) *cell_ret = cell;
  return rank;
}

__device__ __forceinline__ int32_t quantize_point(float gx, float gy, float gz, int b,
                                                  const GridDev& g, long long p,
                                                  const PointOut& out) {
  return quantize_point_core(gx, gy, gz, b, g, p, out);
}

constexpr int kGeomThreads = 256;
constexpr int kMaxHistBins = 2048;

template <bool kHist>
__device__ __forceinline__ void hist_add(uint32_t* s_hist, const SortDigits& sd, int32_t rank) {
  if (kHist) {
#pragma unroll 1
    for (int ps = 0; ps < sd.passes; ++ps) {
      const uint32_t digit = (static_cast<uint32_t>(rank) >> sd.shift[ps]) & ((1u << sd.bits[ps]) - 1u);
      atomicAdd(&s_hist[ps * sd.stride + digit], 1u);
    }
  }
}

template <bool kHist>
__device__ __forceinline__ void hist_flush(uint32_t* s_hist, const SortDigits& sd) {
  if (kHist) {
    __syncthreads();
    for (int i = threadIdx.x; i < sd.passes * sd.stride; i += blockDim.x) {
      const uint32_t v = s_hist[i];
      if (v) atomicAdd(&sd.hist[i], v);
    }
  }
}

// K1: dense geom -> rank
template <bool kHist>
__global__ void __launch_bounds__(kGeomThreads)
quantize_rank_kernel(const float* __restrict__ geom, GridDev g, long long P,
                     long long points_per_sample, PointOut out, SortDigits sd) {
  extern __shared__ uint32_t s_hist[];
  if (kHist) {
    for (int i = threadIdx.x; i < sd.passes * sd.stride; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
  }
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P;
       p += (long long)gridDim.x * blockDim.x) {
    const int b = static_cast<int>(p / points_per_sample);
    const int32_t r = quantize_point(geom[p * 3 + 0], geom[p * 3 + 1], geom[p * 3 + 2], b, g, p, out);
    hist_add<kHist>(s_hist, sd, r);
  }
  hist_flush<kHist>(s_hist, sd);
}

// K1': frustum axes + camera matrices -> rank.  grid = (chunks per camera, B*N);
// a block never straddles two cameras, so its 24 camera constants sit in shared
// memory and each thread walks points of one camera in flat (d, h, w) order.
struct GeomArgs {
  const float* us;  // fW
  const float* vs;  // fH
  const float* ds;  // D
  const float* inv_post_rots;  // (BN,9)
  const float* post_trans;     // (BN,3)
  const float* combine;        // (BN,9)
  const float* trans;          // (BN,3)
  // when raw != 0 the block computes K0 itself from these (fused plan path)
  const float* rots;
  const float* intrins;
  const float* post_rots;
  int raw;
  int N, D, fH, fW;
  float* geom;  // (P,3) or null
};

template <bool kHist>
__global__ void __launch_bounds__(kGeomThreads)
geometry_rank_kernel(GeomArgs a, GridDev g, FastDiv div_hw, FastDiv div_w, PointOut out,
                     SortDigits sd) {
  extern __shared__ uint32_t s_hist[];
  __shared__ float s_cam[24];  // ipr[9] combine[9] post_trans[3] trans[3]
  const int bn = blockIdx.y;
  if (kHist) {
    for (int i = threadIdx.x; i < sd.passes * sd.stride; i += blockDim.x) s_hist[i] = 0;
  }
  if (a.raw) {
    if (threadIdx.x == 0) {
      float r[9], k[9], pr[9], ipr[9], cmb[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        r[j] = a.rots[bn * 9 + j];
        k[j] = a.intrins[bn * 9 + j];
        pr[j] = a.post_rots[bn * 9 + j];
      }
      camera_prep_one(r, k, pr, ipr, cmb);
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        s_cam[j] = ipr[j];
        s_cam[9 + j] = cmb[j];
      }
    }
    if (threadIdx.x >= 32 && threadIdx.x < 35) s_cam[18 + threadIdx.x - 32] = a.post_trans[bn * 3 + threadIdx.x - 32];
    if (threadIdx.x >= 64 && threadIdx.x < 67) s_cam[21 + threadIdx.x - 64] = a.trans[bn * 3 + threadIdx.x - 64];
  } else {
    if (threadIdx.x < 9) s_cam[threadIdx.x] = a.inv_post_rots[bn * 9 + threadIdx.x];
    else if (threadIdx.x < 18) s_cam[threadIdx.x] = a.combine[bn * 9 + threadIdx.x - 9];
    else if (threadIdx.x < 21) s_cam[threadIdx.x] = a.post_trans[bn * 3 + threadIdx.x - 18];
    else if (threadIdx.x < 24) s_cam[threadIdx.x] = a.trans[bn * 3 + threadIdx.x - 21];
  }
  __syncthreads();
  float m[9], c[9], pt[3], tr[3];
#pragma unroll
  for (int j = 0; j < 9; ++j) { m[j] = s_cam[j]; c[j] = s_cam[9 + j]; }
#pragma unroll
  for (int j = 0; j < 3; ++j) { pt[j] = s_cam[18 + j]; tr[j] = s_cam[21 + j]; }

  const int hw = a.fH * a.fW;
  const int ppc = a.D * hw;  // points per camera
  const int b = bn / a.N;
  const long long base = (long long)bn * ppc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ppc; i += gridDim.x * blockDim.x) {
    uint32_t d, rem, h, w;
    div_hw.divmod(static_cast<uint32_t>(i), d, rem);
    div_w.divmod(rem, h, w);
    // points = frustum - post_trans                         model_baseline.py:59
    const float p0 = __fsub_rn(__ldg(a.us + w), pt[0]);
    const float p1 = __fsub_rn(__ldg(a.vs + h), pt[1]);
    const float p2 = __fsub_rn(__ldg(a.ds + d), pt[2]);
    // points = inverse(post_rots) @ points                  :60
    const float q0 = dot3_nofma(m[0], m[1], m[2], p0, p1, p2);
    const float q1 = dot3_nofma(m[3], m[4], m[5], p0, p1, p2);
    const float q2 = dot3_nofma(m[6], m[7], m[8], p0, p1, p2);
    // (x*z, y*z, z)                                         :63-65
    const float r0 = __fmul_rn(q0, q2), r1 = __fmul_rn(q1, q2), r2 = q2;
    // combine @ points + trans                              :66-68
    const float gx = __fadd_rn(dot3_nofma(c[0], c[1], c[2], r0, r1, r2), tr[0]);
    const float gy = __fadd_rn(dot3_nofma(c[3], c[4], c[5], r0, r1, r2), tr[1]);
    const float gz = __fadd_rn(dot3_nofma(c[6], c[7], c[8], r0, r1, r2), tr[2]);
    const long long p = base + i;
    if (a.geom) {
      a.geom[p * 3 + 0] = gx;
      a.geom[p * 3 + 1] = gy;
      a.geom[p * 3 + 2] = gz;
    }
    const int32_t r = quantize_point(gx, gy, gz, b, g, p, out);
    hist_add<kHist>(s_hist, sd, r);
  }
  hist_flush<kHist>(s_hist, sd);
}

}  // namespace lss
