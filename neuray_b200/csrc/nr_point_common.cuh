// Device helpers shared by the SIMT point kernel (nr_point_kernel.cu) and the tensor-core point kernel
// (nr_point_kernel_tc.cu): swizzled column-major activation tiles, weight staging, register-tiled fp32 GEMM fragment.
#pragma once
#include "nr_common.cuh"

namespace nr {
namespace pk {

constexpr int NT = 256;        // compute threads per CTA (one per (point,view) row)

__device__ __forceinline__ int swz(int col) { return ((col >> 2) & 7) << 2; }
// element (col,row) of a column-major tile with row stride ld
template <int ld>
__device__ __forceinline__ float& at(float* tile, int col, int row) { return tile[col * ld + (row ^ swz(col))]; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct Ctx {
  float* sm;     // arena
  float* wbuf;   // staged weights
  int tid, lane, warp;
};

__device__ __forceinline__ void stage(const Ctx& c, const float* __restrict__ g, int n) {
  for (int i = c.tid * 4; i < n; i += NT * 4) *reinterpret_cast<float4*>(c.wbuf + i) = ldg4(g + i);
}

// ------------------------------------------------------------------------------------------------------------
// Register-tiled GEMM fragment: TR rows x (4*TCH) output columns per thread.
//   out[r][j] = sum_k A[k][r] * WT[k][j]
// Thread mapping inside a warp: NCG = OUT/(4*TCH) column groups x 32/NCG row groups, so that one warp-wide
// LDS.128 of the activations touches 32/NCG distinct 16-byte chunks (<= 128 B, one wavefront) and the weight
// loads touch NCG consecutive chunks (broadcast to the other lanes).
template <int OUT, int TR, int TCH>
struct Frag {
  static constexpr int NC = 4 * TCH;
  static constexpr int NCG = OUT / NC;
  static constexpr int RGW = 32 / NCG;
  static_assert(OUT % NC == 0 && 32 % NCG == 0 && (TR == 4 || TR == 8), "bad fragment shape");
  float acc[TR][NC];
  int r0, ja, jb;

  __device__ __forceinline__ void setup(const Ctx& c) {
    const int cg = c.lane % NCG;
    r0 = (c.warp * RGW + c.lane / NCG) * TR;
    ja = 4 * cg;
    jb = OUT / 2 + 4 * cg;   // only used when TCH == 2
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[i][j] = 0.f;
  }
  // bias[j] into every row
  __device__ __forceinline__ void init_bias(const float* __restrict__ b) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float v = b[(j < 4 ? ja : jb - 4) + j];
#pragma unroll
      for (int i = 0; i < TR; ++i) acc[i][j] = v;
    }
  }
  // one k step: activations at ap (rows r0..) / ap2 (rows r0+4..), weights at wp (row k of WT)
  __device__ __forceinline__ void step(const float* __restrict__ ap, const float* __restrict__ ap2, const float* __restrict__ wp) {
    float a[TR];
    {
      const float4 t = *reinterpret_cast<const float4*>(ap);
      a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
    }
    if constexpr (TR == 8) {
      const float4 t = *reinterpret_cast<const float4*>(ap2);
      a[4] = t.x; a[5] = t.y; a[6] = t.z; a[7] = t.w;
    }
    float w[NC];
    {
      const float4 t = *reinterpret_cast<const float4*>(wp + ja);
      w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
    }
    if constexpr (TCH == 2) {
      const float4 t = *reinterpret_cast<const float4*>(wp + jb);
      w[4] = t.x; w[5] = t.y; w[6] = t.z; w[7] = t.w;
    }
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
  // accumulate k = 0..KN-1; activations from tile `tA` (row stride ld), columns colA0+k (colA0 % 4 == 0, so the
  // row swizzle is constant inside each group of 4 columns); weights sW[k*OUT + j].  The k loop is a real loop
  // (groups of 4, unrolled by UNR groups): fully unrolling every layer made the kernel ~700 KB of SASS and the
  // warps starved on instruction fetch (ncu: stall_no_instruction 3.1 per issue, profiles/r1_point_kernel_v0).
  template <int KN, int ld, int UNR = 2>
  __device__ __forceinline__ void mac(const float* __restrict__ tA, int colA0, const float* __restrict__ sW) {
    constexpr int KQ = KN / 4, KT = KN % 4;
#pragma unroll UNR
    for (int kq = 0; kq < KQ; ++kq) {
      const int col0 = colA0 + 4 * kq;
      const int sw = swz(col0);
      const float* ap = tA + col0 * ld + (r0 ^ sw);
      const float* ap2 = tA + col0 * ld + ((r0 + 4) ^ sw);
      const float* wp = sW + 4 * kq * OUT;
#pragma unroll
      for (int i = 0; i < 4; ++i) step(ap + i * ld, ap2 + i * ld, wp + i * OUT);
    }
    if constexpr (KT > 0) {
      const int col0 = colA0 + 4 * KQ;
      const int sw = swz(col0);
      const float* ap = tA + col0 * ld + (r0 ^ sw);
      const float* ap2 = tA + col0 * ld + ((r0 + 4) ^ sw);
      const float* wp = sW + 4 * KQ * OUT;
#pragma unroll
      for (int i = 0; i < KT; ++i) step(ap + i * ld, ap2 + i * ld, wp + i * OUT);
    }
  }
  // epi(col j, first row r (multiple of 4), float4 of the 4 consecutive rows)
  template <class Epi>
  __device__ __forceinline__ void store(Epi epi) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int col = (j < 4 ? ja : jb - 4) + j;
      epi(col, r0, make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]));
      if constexpr (TR == 8) epi(col, r0 + 4, make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]));
    }
  }
};

template <int ld>
__device__ __forceinline__ float4& at4(float* tile, int col, int row) {
  return *reinterpret_cast<float4*>(tile + col * ld + (row ^ swz(col)));
}
__device__ __forceinline__ float4 elu4(float4 v) { return make_float4(elu(v.x), elu(v.y), elu(v.z), elu(v.w)); }


}  // namespace pk
}  // namespace nr
