// Point kernel (namespace nr::pkt): everything of a render pass that is per (point, view) row, the cross-view pooling and
// the per-point layers -- reference renderer.py:171-178 up to geometry_fc, ibrnet.py:315-354.  One kernel
// (nr_point_kernel_pm3.cuh, three 128-row blocks per SM); this file holds what its pieces share and the launcher:
//
//   * thread r owns (point,view) row r of a 128-row block = TMEM lane r (a warp may only touch its own 32-lane TMEM
//     quadrant, which is exactly "its" rows); row = point * G + view, G = lanes per point = power of two >= rfn
//   * a layer  D[128 x N] += A[128 x K] * W^T :  A (hi and lo parts of the 3xTF32 split) is written to TMEM by the row
//     owners with tcgen05.st, W^T (hi / lo, pre-split and pre-swizzled by nr_pack_weights) sits in shared memory, D comes
//     back with tcgen05.ld for the bias / activation epilogue, whose result goes straight back to TMEM as the next A
//   * weights stream through a 3 x 16 KB shared-memory ring filled by TMA bulk copies (cp.async.bulk); full/empty
//     mbarriers, the "empty" arrivals are tcgen05.commit of the consuming MMAs
#include <type_traits>

#include "nr_common.cuh"
#include "nr_tc.cuh"

namespace nr {
namespace pkt {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

constexpr int REC = NR_POINT_REC;
constexpr int RING_STAGE = 4096;    // 16 KB
constexpr int NBUF = 4;             // depth of the shared weight ring.  The three blocks of a CTA consume every stage, so the depth bounds how far
                                    // they can drift apart: 3 -> 4 slots = +1.9 % (166.6 -> 169.8 M); more does not fit next to the resident
                                    // base_fc.0 tile, and streaming that tile too (19 stages per tile, 6..9 slots) costs 13..18 %.
                                    // Also measured and rejected (profiles/README.md): feeding from all three blocks through a shared
                                    // claim counter (-9 %), feeding from a non-issuing warp of block 0 (-2.5 %), a non-inlined feed (-12 %)
constexpr int SW = 2048;

// ---- small resident weights (floats inside `sw`) ----
constexpr int SW_HEAD = 0, SW_HEAD_STRIDE = 136;   // per head: L0 bias 32 | L1 bias 32 | L2 W[2][32] | L2 bias 4
constexpr int SW_PE0B = 544, SW_PE1B = 576, SW_RD0W = 608, SW_RD0B = 672, SW_RD1W = 688, SW_RD1B = 1264, SW_NF0W = 1300,
              SW_NF0B = 1556, SW_NF1W = 1564, SW_NF1B = 1572, SW_B1B = 1576, SW_V0B = 1608, SW_V1B = 1640, SW_V1LW = 1672,
              SW_V1LB = 1704, SW_V20B = 1708, SW_V21W = 1740, SW_V21B = 1772, SW_RGB0B = 1776, SW_RGB1W = 1792, SW_RGB1B = 1920,
              SW_RGB2W = 1928, SW_RGB2B = 1936,
              SW_NF0C = 1944;   // neuray_fc.0 bias seen through prob_embed.2: W_nf0 @ b_pe2 + b_nf0 (pm3 kernel)
static_assert(SW_NF0C + 8 <= SW, "small weights overflow");

struct KParams {
  NrPassParams p;
  float* dbg;
  long long* timing;   // optional: clock64() at phase boundaries, CTA 0, threads 0 and 128: [tile][2][32]
  int P, n_tiles, n_heads;
};

// The weight stream, driven by ONE thread (thread 0) from inside its own MMA-issue path: before it waits for a stage it
// issues every stage up to that one (blocking on the ring slot if it has to) and opportunistically up to NBUF-1 further.
struct Producer {
  const float* w_tc;
  float* ring;
  uint64_t *wfull, *wempty;
  const int2* table;      // shared: per stage of a tile (source offset in floats, bytes), built once at kernel start
  uint32_t next;          // next stage index to issue
  uint32_t total;         // iters * stages_per_tile
  int stages_per_tile;
  int s;                  // next % stages_per_tile, kept as a running counter (this code runs on one lane, inside block 0's MMA-issue
                          // path: every instruction counts)

  // (source offset, bytes) of stage s of a tile: heads, prob_embed x2, base_fc.0 x3, base_fc.2, vis, vis2+rgb, geometry_fc.0 x3
  __device__ static int2 stage_source(int s, int n_heads) {
    if (s < n_heads) return make_int2(tcl::HEAD0 + s * RING_STAGE, RING_STAGE * 4);
    const int t = s - n_heads;
    const int src = t == 0 ? tcl::PE0 : t == 1 ? tcl::PE1 : t <= 4 ? tcl::B0 + (t - 2) * RING_STAGE : t == 5 ? tcl::B1 : t == 6 ? tcl::V01
                  : t == 7 ? tcl::V2R : tcl::G0 + (t - 8) * RING_STAGE;
    return make_int2(src, t == 1 ? 3072 * 4 : RING_STAGE * 4);
  }
  __device__ __forceinline__ void issue() {
    const int2 e = table[s];
    const uint32_t buf = next % NBUF;
    tc::mbar_arrive_expect_tx(wfull + buf, e.y);
    tc::bulk_g2s(ring + buf * RING_STAGE, w_tc + e.x, e.y, wfull + buf);
    ++next;
    if (++s == stages_per_tile) s = 0;
  }
  // make sure stages [.., last] are in flight; then try to run ahead without blocking
  __device__ __forceinline__ void feed(uint32_t last) {
    while (next <= last && next < total) {
      if (next >= NBUF) tc::mbar_wait(wempty + (next % NBUF), ((next / NBUF) - 1) & 1);
      issue();
    }
    while (next < total && next <= last + (NBUF - 1)) {
      if (next >= NBUF && !tc::mbar_try_wait(wempty + (next % NBUF), ((next / NBUF) - 1) & 1)) break;
      issue();
    }
  }
};

// Everything a compute thread needs to drive its block's tensor-core layers.
struct Blk {
  uint32_t tAhi, tAlo, tD;      // TMEM addresses of this thread's lane quadrant (lane field included)
  uint32_t mAhi, mAlo, mD;      // same columns, lane 0: what the MMA instruction takes
  uint64_t* mma_bar;
  uint64_t *wfull, *wempty;     // arrays [NBUF]
  uint32_t ring_addr;           // shared-space byte address of the ring
  uint32_t phase;               // parity of the next mma_bar completion
  uint32_t wi;                  // weight stages consumed so far (same sequence in every thread)
  int blk;
  bool leader;
  bool issuer_warp;             // warp 0 of the block (warp-uniform): one elected lane of it issues the MMAs
  Producer* prod;               // non-null in the one thread that also feeds the weight ring (point-major kernel)
  long long* tk;                // diagnostics: where run_layer drops clock64() stamps (nullptr: off)
};

// bias vector (shared memory, 16-byte aligned) added to 32 values with 128-bit loads
__device__ __forceinline__ void add_bias32(float* x, const float* __restrict__ bias) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(bias + 4 * q);
    x[4 * q] += t.x; x[4 * q + 1] += t.y; x[4 * q + 2] += t.z; x[4 * q + 3] += t.w;
  }
}

namespace pm {
template <int G>
__device__ __forceinline__ float bsum(float x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
template <int G>
__device__ __forceinline__ float bmax(float x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// N independent butterflies advanced together: the N shuffles of a step are independent, so they pipeline instead of
// forming one latency chain per value (the scalar form ran the 140 pooled statistics at ~80 cycles each).
template <int G, int N>
__device__ __forceinline__ void bsum_vec(float (&x)[N]) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    float t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = __shfl_xor_sync(0xffffffffu, x[i], o);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] += t[i];
  }
}

// Sum of N (16 or 32) per-row values over the G rows (views) of each point, result in every lane of the point -- through
// the warp-private transposition buffer instead of N butterflies (3 shuffles + 3 adds per value at G = 8):
//   every lane stores its row [N] (128-bit stores, row stride 36 floats: conflict-free), lane v of a point then adds up its
//   N/G columns over the point's G rows (the G lanes of a point read one contiguous N*4-byte stretch per row), writes the
//   partial row into the point's first row, and every lane reads the N sums back (broadcast reads).  ~57 instructions for
//   32 values at G = 8 against 192.  `stg`: this warp's [32][36] buffer; lane0: first lane of the point.
template <int G, int N>
__device__ __forceinline__ void pool_rows(float* stg, int lane, int lane0, int v, const float (&in)[N], float (&out)[N]) {
  static_assert(N == 16 || N == 32, "row length");
  constexpr int C = N / G > 0 ? N / G : 1;          // columns per lane (G = 32, N = 16: lanes >= 16 idle)
  constexpr int ROW = 36;
#pragma unroll
  for (int q = 0; q < N / 4; ++q) *reinterpret_cast<float4*>(stg + lane * ROW + 4 * q) = make_float4(in[4 * q], in[4 * q + 1], in[4 * q + 2], in[4 * q + 3]);
  __syncwarp();
  float acc[C];
#pragma unroll
  for (int j = 0; j < C; ++j) acc[j] = 0.f;
  const bool active = v * C < N;
  if (active) {
#pragma unroll
    for (int t = 0; t < G; ++t) {
      const float* __restrict__ src = stg + (lane0 + t) * ROW + v * C;
      if constexpr (C >= 4) {
#pragma unroll
        for (int j = 0; j < C; j += 4) {
          const float4 x = *reinterpret_cast<const float4*>(src + j);
          acc[j] += x.x; acc[j + 1] += x.y; acc[j + 2] += x.z; acc[j + 3] += x.w;
        }
      } else if constexpr (C == 2) {
        const float2 x = *reinterpret_cast<const float2*>(src);
        acc[0] += x.x; acc[1] += x.y;
      } else {
        acc[0] += src[0];
      }
    }
  }
  __syncwarp();
  if (active) {
#pragma unroll
    for (int j = 0; j < C; ++j) stg[lane0 * ROW + v * C + j] = acc[j];
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    const float4 x = *reinterpret_cast<const float4*>(stg + lane0 * ROW + 4 * q);
    out[4 * q] = x.x; out[4 * q + 1] = x.y; out[4 * q + 2] = x.z; out[4 * q + 3] = x.w;
  }
  __syncwarp();
}

// ---- lane groups that are not a power of two (G = 5, 6, 10: PPW = 32 / G points per warp, the last 32 - PPW * G lanes idle) ----
// Butterflies do not exist for them; a group's scalars are summed through columns 32..35 of the group's rows in the same
// warp-private buffer (pool_rows only touches columns 0..31).  `idle` lanes take part in the warp syncs only.
template <int G>
__device__ __forceinline__ float gsum_np2(float* stg, int lane, int lane0, bool idle, float x) {
  stg[lane * 36 + 32] = x;
  __syncwarp();
  float s = 0.f;
  if (!idle) {
#pragma unroll
    for (int t = 0; t < G; ++t) s += stg[(lane0 + t) * 36 + 32];
  }
  __syncwarp();
  return s;
}
template <int G>
__device__ __forceinline__ float gmax_np2(float* stg, int lane, int lane0, bool idle, float x) {
  stg[lane * 36 + 32] = x;
  __syncwarp();
  float s = x;
  if (!idle) {
#pragma unroll
    for (int t = 0; t < G; ++t) s = fmaxf(s, stg[(lane0 + t) * 36 + 32]);
  }
  __syncwarp();
  return s;
}
template <int G>
__device__ __forceinline__ void gsum4_np2(float* stg, int lane, int lane0, bool idle, float (&x)[4]) {
  *reinterpret_cast<float4*>(stg + lane * 36 + 32) = make_float4(x[0], x[1], x[2], x[3]);
  __syncwarp();
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!idle) {
#pragma unroll
    for (int t = 0; t < G; ++t) {
      const float4 y = *reinterpret_cast<const float4*>(stg + (lane0 + t) * 36 + 32);
      s.x += y.x; s.y += y.y; s.z += y.z; s.w += y.w;
    }
  }
  __syncwarp();
  x[0] = s.x; x[1] = s.y; x[2] = s.z; x[3] = s.w;
}
// pool_rows for such a group: lane v adds up columns v, v + G, v + 2G, ... of the group's G rows
template <int G, int N>
__device__ __forceinline__ void pool_rows_np2(float* stg, int lane, int lane0, int v, bool idle, const float (&in)[N], float (&out)[N]) {
  constexpr int ROW = 36;
  constexpr int C = (N + G - 1) / G;
#pragma unroll
  for (int q = 0; q < N / 4; ++q) *reinterpret_cast<float4*>(stg + lane * ROW + 4 * q) = make_float4(in[4 * q], in[4 * q + 1], in[4 * q + 2], in[4 * q + 3]);
  __syncwarp();
  float acc[C];
#pragma unroll
  for (int j = 0; j < C; ++j) acc[j] = 0.f;
  if (!idle) {
#pragma unroll
    for (int t = 0; t < G; ++t) {
#pragma unroll
      for (int j = 0; j < C; ++j)
        if (v + G * j < N) acc[j] += stg[(lane0 + t) * ROW + v + G * j];
    }
  }
  __syncwarp();
  if (!idle) {
#pragma unroll
    for (int j = 0; j < C; ++j)
      if (v + G * j < N) stg[lane0 * ROW + v + G * j] = acc[j];
  }
  __syncwarp();
  if (!idle) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 x = *reinterpret_cast<const float4*>(stg + lane0 * ROW + 4 * q);
      out[4 * q] = x.x; out[4 * q + 1] = x.y; out[4 * q + 2] = x.z; out[4 * q + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) out[q] = 0.f;
  }
  __syncwarp();
}

// The 64/G output columns a lane owns in the streamed per-point layers.  Lane v takes, for q = 0..CPL/4-1, the float4
// at columns q*4G + 4v: the G lanes of a point then read G consecutive 16-byte chunks (one conflict-free wavefront) for
// every q, instead of 32-byte chunks whose second halves collide in the banks.  (G = 32: two columns 2v, 2v+1.)
template <int G>
__device__ __forceinline__ int own_col(int v, int j) {
  constexpr int CPL = 64 / G;
  if constexpr (CPL == 2) return 2 * v + j;
  else return (j >> 2) * 4 * G + 4 * v + (j & 3);
}
// this lane's CPL values of a 64-wide row at `row`
template <int G>
__device__ __forceinline__ void ld_cols(const float* __restrict__ row, int v, float* w) {
  constexpr int CPL = 64 / G;
  if constexpr (CPL == 2) {
    const float2 t = *reinterpret_cast<const float2*>(row + 2 * v);
    w[0] = t.x; w[1] = t.y;
  } else {
#pragma unroll
    for (int q = 0; q < CPL / 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(row + q * 4 * G + 4 * v);
      w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
    }
  }
}

}  // namespace pm

#include "nr_point_kernel_pm3.cuh"

}  // namespace pkt

static int device_sms() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// The dynamic shared-memory opt-in is a per-device function attribute; it is set on every launch (a few hundred
// nanoseconds against a kernel of milliseconds) so that the library keeps no per-process state: a process may render on
// any number of devices from any number of threads.
template <int G, bool DEBUG>
static int launch_pm3_inst(const pkt::KParams& kp, int grid, cudaStream_t stream) {
  auto* fn = pkt::pm3::point_kernel_pm3<G, DEBUG>;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pkt::pm3::SMEM_BYTES)) != cudaSuccess) {
    nr::set_error("point kernel: cannot opt in to %zu bytes of shared memory: %s", pkt::pm3::SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
    return NR_E_CUDA;
  }
  fn<<<grid, pkt::pm3::NTHR, pkt::pm3::SMEM_BYTES, stream>>>(kp);
  NR_CHECK_LAUNCH("point_kernel_pm3");
  return NR_OK;
}

template <int G>
static int launch_pm3(pkt::KParams kp, cudaStream_t stream) {
  constexpr int PB = 4 * (32 / G);       // points per 128-row block: 32 / G per warp (G not a power of two leaves 32 % G lanes idle)
  const long long N = (long long)kp.p.rn * kp.p.dn;
  kp.P = PB;
  kp.n_tiles = int((N + PB - 1) / PB);
  const int groups = (kp.n_tiles + pkt::pm3::NBLK - 1) / pkt::pm3::NBLK;
  const int sms = device_sms();
  const int grid = groups < sms ? groups : sms;
  return (kp.dbg || kp.timing) ? launch_pm3_inst<G, true>(kp, grid, stream) : launch_pm3_inst<G, false>(kp, grid, stream);
}

// `timing`: optional clock64() stamps at the phase boundaries (nr_point_kernel_timing), passed per call
int launch_point_kernel(const NrPassParams* p, float* dbg, long long* timing, cudaStream_t stream) {
  NR_CHECK_ARG(p != nullptr, "params");
  NR_CHECK_ARG(p->coords && p->que_depth && p->que_cam && p->feat && p->rgb && p->view_params && p->w_point && p->point_rec,
               "null device pointer");
  NR_CHECK_ARG(p->w_tc != nullptr, "w_tc (tensor-core weight buffer, nr_pack_weights) is required");
  NR_CHECK_ARG(p->rfn >= 1 && p->rfn <= NR_MAX_VIEWS, "rfn out of range");
  NR_CHECK_ARG(p->dn >= 3 && p->dn <= NR_MAX_SAMPLES, "dn out of range");
  NR_CHECK_ARG(p->rn >= 0, "rn");
  NR_CHECK_ARG(p->h > 1 && p->w > 1 && p->fh > 0 && p->fw > 0, "map shape");
  NR_CHECK_ARG((long long)p->rn * p->dn < (1ll << 31) / NR_POINT_REC, "too many points for one call; chunk the rays");
  NR_CHECK_ARG((long long)p->rfn * p->fh * p->fw * 64 < (1ll << 31), "feature maps too large for 32-bit texel offsets");
  if (p->rn == 0) return NR_OK;
  pkt::KParams kp;
  kp.p = *p;
  kp.dbg = dbg;
  kp.timing = timing;
  kp.n_heads = p->use_vis ? 4 : 3;
  // lanes per point: the smallest group size that holds the views and wastes the fewest of a warp's 32 lanes
  if (p->rfn <= 4) return launch_pm3<4>(kp, stream);
  if (p->rfn == 5) return launch_pm3<5>(kp, stream);      // 6 points per warp (30 lanes) instead of 4 groups of 8
  if (p->rfn == 6) return launch_pm3<6>(kp, stream);      // 5 points per warp
  if (p->rfn <= 8) return launch_pm3<8>(kp, stream);
  if (p->rfn <= 10) return launch_pm3<10>(kp, stream);    // 3 points per warp instead of 2 groups of 16 (cfg4: 10 views)
  if (p->rfn <= 16) return launch_pm3<16>(kp, stream);
  return launch_pm3<32>(kp, stream);
}

}  // namespace nr
