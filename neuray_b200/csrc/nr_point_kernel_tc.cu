// Point kernel, tensor-core version: same work, tile shape and outputs as nr_point_kernel.cu (see its header), but
// the 32/64-wide dense layers run on tcgen05 with fp32-accurate 3xTF32:
//
//   * thread r owns (point,view) row r of the 256-row tile = TMEM lane r % 128 of block r / 128 (warps 0-3 -> block 0,
//     warps 4-7 -> block 1: a warp may only touch its own 32-lane TMEM quadrant, which is exactly "its" rows)
//   * a layer  D[128 x N] += A[128 x K] * W^T :  A (hi and lo parts) is written to TMEM by the row owners with
//     tcgen05.st, W^T (hi / lo, pre-split and pre-swizzled on the host) sits in shared memory, D comes back with
//     tcgen05.ld for the bias / activation epilogue, whose result goes straight back to TMEM as the next layer's A.
//     Hidden activations never touch shared memory; TMEM holds what shared memory could not (hi+lo of everything).
//   * weights stream through a 3 x 16 KB shared-memory ring filled by TMA bulk copies (cp.async.bulk) issued by a
//     dedicated producer warp; full/empty mbarriers, the "empty" arrivals are tcgen05.commit of the consuming MMAs
//   * the two 128-row blocks run the layer chain independently (own named barrier + MMA mbarrier), so one block's
//     epilogue overlaps the other block's MMAs; CTA-wide syncs only where views are reduced
//   * gather, projections, per-view scalar heads, the cross-view reductions and the per-point layers (hoisted
//     base_fc.0, geometry_fc) stay SIMT exactly as in the SIMT kernel
#include <stdlib.h>
#include <type_traits>

#include "nr_common.cuh"
#include "nr_point_common.cuh"
#include "nr_tc.cuh"

namespace nr {
namespace pkt {

using pk::at;
using pk::at4;
using pk::elu4;
using pk::Frag;
using pk::ldg4;
using pk::NT;

constexpr int NTHREADS = NT + 32;   // 8 compute warps + 1 producer warp
constexpr int LD = 256;
constexpr int LDP = 32;             // points per tile <= 32
constexpr int REC = NR_POINT_REC;

// ---- shared memory map (floats from the 1024-byte aligned base) ----
constexpr int RING_STAGE = 4096;    // 16 KB
constexpr int NBUF = 3;
constexpr int OFF_RING = 0;
constexpr int C_SCAL = 0, C_A = 24, C_RF = 60, C_PP = 100, N_COLS = 126;
constexpr int OFF_ARENA = OFF_RING + NBUF * RING_STAGE;
constexpr int OFF_WBUF = OFF_ARENA + N_COLS * LD;
constexpr int WBUF = 4608;
constexpr int OFF_WBUF2 = OFF_WBUF + WBUF;
constexpr int OFF_SW = OFF_WBUF2 + WBUF;
constexpr int SW = 2048;
constexpr int OFF_GEO = OFF_SW + SW;              // per-point ray geometry, double buffered: [2][8][LDP]
constexpr int GEO = 2 * 8 * LDP;
constexpr int OFF_BAR = OFF_GEO + GEO;            // 12 mbarriers (96 B) + tmem base
constexpr int SMEM_FLOATS = OFF_BAR + 32;
constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * 4;

enum { S_MASK = 0, S_Z, S_HIT, S_VIS, S_W1, S_W0, S_VISA, S_VIS2, S_W2, S_DD0, S_DD1, S_DD2, S_DD3, S_R, S_G, S_B,
       S_LOGIT, S_IX, S_IY, S_PT0 };
enum { P_X = 0, P_Y, P_Z, P_QX, P_QY, P_QZ, P_IHP, P_IHC, P_NV, P_NARR };
static_assert(P_NARR * LDP <= 4 * LD, "per-point arrays overflow");
// per-point tiles inside the PP region (floats from the region start, column stride LDP)
constexpr int PP_GLOB = 0, PP_G = 140 * LDP, PP_GVEC = 0, PP_GHID = 68 * LDP, PP_GOUT = PP_GHID + 64 * LDP;
static_assert(PP_G + 64 * LDP <= 26 * LD && PP_GOUT + 20 * LDP <= 26 * LD, "PP region overflow");

// ---- small resident weights (floats inside `sw`) ----
constexpr int SW_HEAD = 0, SW_HEAD_STRIDE = 136;   // per head: L0 bias 32 | L1 bias 32 | L2 W[2][32] | L2 bias 4
constexpr int SW_PE0B = 544, SW_PE1B = 576, SW_RD0W = 608, SW_RD0B = 672, SW_RD1W = 688, SW_RD1B = 1264, SW_NF0W = 1300,
              SW_NF0B = 1556, SW_NF1W = 1564, SW_NF1B = 1572, SW_B1B = 1576, SW_V0B = 1608, SW_V1B = 1640, SW_V1LW = 1672,
              SW_V1LB = 1704, SW_V20B = 1708, SW_V21W = 1740, SW_V21B = 1772, SW_RGB0B = 1776, SW_RGB1W = 1792, SW_RGB1B = 1920,
              SW_RGB2W = 1928, SW_RGB2B = 1936,
              SW_NF0C = 1944;   // neuray_fc.0 bias seen through prob_embed.2: W_nf0 @ b_pe2 + b_nf0 (pm3 kernel)
static_assert(SW_NF0C + 8 <= SW, "small weights overflow");

// ---- TMEM columns per 128-row block (block b starts at column 256*b) ----
constexpr int T_AHI = 0, T_ALO = 80, T_D = 160;

struct KParams {
  NrPassParams p;
  float* dbg;
  long long* timing;   // optional: clock64() at phase boundaries, CTA 0, threads 0 and 128: [tile][2][32]
  int P, n_tiles, n_heads;
};

__device__ __forceinline__ void sync_compute() { tc::named_sync(1, NT); }

// asynchronous staging of SIMT-layer weights (cp.async, 16 B per request): issued a phase ahead, waited for at use
__device__ __forceinline__ void stage_async(float* dst, const float* __restrict__ src, int n, int tid) {
  for (int i = tid * 4; i < n; i += NT * 4)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(tc::smem_u32(dst + i)), "l"(src + i) : "memory");
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// The weight stream, driven by ONE thread (thread 0) from inside its own MMA-issue path: before it waits for a stage it
// issues every stage up to that one (blocking on the ring slot if it has to) and opportunistically up to NBUF-1 further.
struct Producer {
  const float* w_tc;
  float* ring;
  uint64_t *wfull, *wempty;
  uint32_t next;          // next stage index to issue
  uint32_t total;         // iters * stages_per_tile
  int stages_per_tile, n_heads;

  __device__ __forceinline__ void issue() {
    const int s = int(next % uint32_t(stages_per_tile));
    int src, bytes = RING_STAGE * 4;
    if (s < n_heads) src = tcl::HEAD0 + s * RING_STAGE;
    else {
      const int t = s - n_heads;
      src = t == 0 ? tcl::PE0 : t == 1 ? tcl::PE1 : t <= 4 ? tcl::B0 + (t - 2) * RING_STAGE : t == 5 ? tcl::B1 : t == 6 ? tcl::V01 : t == 7 ? tcl::V2R
                                                                                                                       : tcl::G0 + (t - 8) * RING_STAGE;   // pm3 only
      if (t == 1) bytes = 3072 * 4;
    }
    const uint32_t buf = next % NBUF;
    tc::mbar_arrive_expect_tx(wfull + buf, bytes);
    tc::bulk_g2s(ring + buf * RING_STAGE, w_tc + src, bytes, wfull + buf);
    ++next;
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

// Layer issue + completion wait, executed by all 128 threads of the block after they wrote their A columns.
//   NCH K-chunks of 8; chunks c < NLIN read A columns A0 + 8c, the rest ATAIL + 8(c - NLIN); B chunk c lives in slab c/4
//   at byte offset (c%4)*32 inside the slab; a layer reads NSTG ring stages: one (all slabs in it, SLAB bytes apart) or
//   one slab per stage (base_fc.0).  Everything is a template constant so that the single issuing lane executes ~4
//   instructions per MMA (a runtime-parameterised loop cost ~140 cycles per MMA: dependent scalar code on one lane).
template <int N, int NCH, int A0, int NLIN, int ATAIL, int NSTG, uint32_t OFF_HI, uint32_t OFF_LO, uint32_t SLAB, bool WAIT_FULL, bool RELEASE>
__device__ __forceinline__ void run_layer(Blk& b) {
  if (b.tk && b.leader) b.tk[0] = clock64();
  tc::tmem_st_wait();
  tc::fence_before_thread_sync();
  tc::named_sync(2 + b.blk, 128);
  if (b.tk && b.leader) b.tk[2] = clock64();
  if (b.issuer_warp) {                       // warp-uniform branch; one elected lane issues
    tc::fence_after_thread_sync();
    const uint32_t st0 = b.wi % NBUF;
    if (WAIT_FULL && b.prod != nullptr) b.prod->feed(b.wi + NSTG - 1);   // only thread 0 of the point-major kernel
    __syncwarp();
    if (WAIT_FULL) {
#pragma unroll
      for (int s = 0; s < NSTG; ++s) tc::mbar_wait(b.wfull + ((st0 + s) % NBUF), ((b.wi + s) / NBUF) & 1);
    }
    if (b.tk && b.leader) b.tk[3] = clock64();
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_tf32(N);
      uint64_t dhi[NSTG], dlo[NSTG];
#pragma unroll
      for (int s = 0; s < NSTG; ++s) {
        const uint32_t base = b.ring_addr + ((st0 + s) % NBUF) * (RING_STAGE * 4);
        dhi[s] = tc::smem_desc_sw128(base + OFF_HI);
        dlo[s] = tc::smem_desc_sw128(base + OFF_LO);
      }
#pragma unroll
      for (int ps = 0; ps < 3; ++ps) {
        const uint32_t abase = (ps == 1) ? b.mAlo : b.mAhi;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          constexpr int dummy = 0;
          const int acol = c < NLIN ? A0 + 8 * c : ATAIL + 8 * (c - NLIN);
          const int slab = c >> 2;
          const int stg = NSTG == 1 ? 0 : slab;
          const uint32_t inc = ((NSTG == 1 ? slab * SLAB : 0u) + (c & 3) * 32) >> 4;
          tc::mma_tf32_ts(b.mD, abase + acol, (ps == 2 ? dlo[stg] : dhi[stg]) + inc, idesc, (ps | c) != 0);
          (void)dummy;
        }
      }
      if (RELEASE) {
#pragma unroll
        for (int s = 0; s < NSTG; ++s) tc::mma_commit(b.wempty + ((st0 + s) % NBUF));
      }
      tc::mma_commit(b.mma_bar);
    }
    __syncwarp();
    if (b.tk && b.leader) b.tk[4] = clock64();
  }
  if (RELEASE) b.wi += NSTG;
  tc::mbar_wait(b.mma_bar, b.phase);
  b.phase ^= 1;
  tc::fence_after_thread_sync();
  if (b.tk && b.leader) b.tk[5] = clock64();
}

// D columns [c0, c0+16) of this thread's row
__device__ __forceinline__ void load_d16(const Blk& b, int c0, float* v) {
  tc::tmem_ld16(b.tD + c0, v);
  tc::tmem_ld_wait();
}
// D columns [c0, c0+32): both loads in flight before the single wait
__device__ __forceinline__ void load_d32(const Blk& b, int c0, float* v) {
  tc::tmem_ld16(b.tD + c0, v);
  tc::tmem_ld16(b.tD + c0 + 16, v + 16);
  tc::tmem_ld_wait();
}
__device__ __forceinline__ void store_a32(const Blk& b, int col, const float* v);
// 16 activations -> A columns [col, col+16) (hi and lo parts)
__device__ __forceinline__ void store_a16(const Blk& b, int col, const float* v) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) tc::split_tf32(v[j], hi[j], lo[j]);
  tc::tmem_st16(b.tAhi + col, hi);
  tc::tmem_st16(b.tAlo + col, lo);
}

__device__ __forceinline__ void store_a32(const Blk& b, int col, const float* v) {
  store_a16(b, col, v);
  store_a16(b, col + 16, v + 16);
}
// bias vector (shared memory, 16-byte aligned) added to 32 values with 128-bit loads
__device__ __forceinline__ void add_bias32(float* x, const float* __restrict__ bias) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(bias + 4 * q);
    x[4 * q] += t.x; x[4 * q + 1] += t.y; x[4 * q + 2] += t.z; x[4 * q + 3] += t.w;
  }
}

template <bool DEBUG>
__global__ void __launch_bounds__(NTHREADS, 1) point_kernel_tc(const KParams kp) {
  // 1024-byte alignment (SWIZZLE_128B tiles) comes from the declaration: the kernel has no static shared memory, so
  // the dynamic window starts at the CTA's shared base.  (Rounding the pointer up by hand made every access a generic
  // LD/ST instead of LDS/STS.)
  extern __shared__ __align__(1024) float smem[];
  const NrPassParams& pp = kp.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  float* const ring = smem + OFF_RING;
  float* const arena = smem + OFF_ARENA;
  float* const sw = smem + OFF_SW;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* const wfull = bars;             // [NBUF]
  uint64_t* const wempty = bars + NBUF;     // [NBUF]
  uint64_t* const mma_bars = bars + 2 * NBUF;   // [2]
  uint64_t* const pfull = bars + 2 * NBUF + 2;  // [2] per-point geometry of the next tile (producer warp -> compute warps)
  uint64_t* const pempty = bars + 2 * NBUF + 4; // [2]
  uint32_t* const tmem_base_s = reinterpret_cast<uint32_t*>(bars + 2 * NBUF + 6);
  float* const geo = smem + OFF_GEO;

  const int P = kp.P, rfn = pp.rfn, ROWS = P * rfn, dn = pp.dn;
  const int N = pp.rn * dn;
  const int n_heads = kp.n_heads;
  const int stages_per_tile = n_heads + 8;
  const float* __restrict__ W = pp.w_point;

  // ---------------- one-time setup ----------------
  if (tid == 0) {
    for (int i = 0; i < NBUF; ++i) { tc::mbar_init(wfull + i, 1); tc::mbar_init(wempty + i, 2); }
    tc::mbar_init(mma_bars + 0, 1);
    tc::mbar_init(mma_bars + 1, 1);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(pfull + i, 1); tc::mbar_init(pempty + i, 1); }
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(tmem_base_s);
  if (tid < NT) {   // resident small weights
    auto cp = [&](int dst, int src, int n) { for (int i = tid; i < n; i += NT) sw[dst + i] = __ldg(W + src + i); };
    for (int h = 0; h < 4; ++h) {
      const int g = lay::DD_HEAD + h * lay::DD_HEAD_STRIDE, d = SW_HEAD + h * SW_HEAD_STRIDE;
      cp(d, g + lay::DD_L0_B, 32); cp(d + 32, g + lay::DD_L1_B, 32); cp(d + 64, g + lay::DD_L2_W, 64); cp(d + 128, g + lay::DD_L2_B, 4);
    }
    const int gb = lay::GRP_B, gd = lay::GRP_D1;
    cp(SW_PE0B, gb + lay::PE0_B, 32); cp(SW_PE1B, gb + lay::PE1_B, 32); cp(SW_RD0W, gb + lay::RD0_W, 64); cp(SW_RD0B, gb + lay::RD0_B, 16);
    cp(SW_RD1W, gb + lay::RD1_W, 576); cp(SW_RD1B, gb + lay::RD1_B, 36); cp(SW_NF0W, gb + lay::NF0_W, 256); cp(SW_NF0B, gb + lay::NF0_B, 8);
    cp(SW_NF1W, gb + lay::NF1_W, 8); cp(SW_NF1B, gb + lay::NF1_B, 4); cp(SW_B1B, lay::BASE1_B, 32);
    cp(SW_V0B, gd + lay::VIS0_B, 32); cp(SW_V1B, gd + lay::VIS1_B, 32); cp(SW_V1LW, gd + lay::VIS1L_W, 32); cp(SW_V1LB, gd + lay::VIS1L_B, 4);
    cp(SW_V20B, gd + lay::V20_B, 32); cp(SW_V21W, gd + lay::V21_W, 32); cp(SW_V21B, gd + lay::V21_B, 4); cp(SW_RGB0B, gd + lay::RGB0_B, 16);
    cp(SW_RGB1W, gd + lay::RGB1_W, 128); cp(SW_RGB1B, gd + lay::RGB1_B, 8); cp(SW_RGB2W, gd + lay::RGB2_W, 8); cp(SW_RGB2B, gd + lay::RGB2_B, 4);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_base_s;

  // ---------------- producer warp: per-point ray geometry one tile ahead + the tensor-core weight stream ----------------
  if (warp == NT / 32) {
    // depth2points / depth2inv_dists for the P points of a tile (reference render_ops.py:27-52), one lane per point
    auto geometry = [&](int tile, int j) {
      const int buf = j & 1;
      if (j >= 2) tc::mbar_wait(pempty + buf, ((j >> 1) - 1) & 1);
      float* __restrict__ g = geo + buf * 8 * LDP;
      const int n = tile * P + lane;
      float px = 0.f, py = 0.f, pz = 0.f, qx = 0.f, qy = 0.f, qz = 0.f, ihp = 0.f, ihc = 0.f;
      if (lane < P && n < N) {
        const float* __restrict__ cam = pp.que_cam;
        const int ray = n / dn, s = n - ray * dn;
        const float cx = __ldg(pp.coords + 2 * ray), cy = __ldg(pp.coords + 2 * ray + 1);
        float cm[3], d[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) cm[i] = fmaf(cam[12 + 3 * i + 1], cy, cam[12 + 3 * i] * cx) + cam[12 + 3 * i + 2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float wld = fmaf(cam[3 * i + 2], cm[2], fmaf(cam[3 * i + 1], cm[1], cam[3 * i] * cm[0])) + cam[9 + i];
          d[i] = wld - cam[9 + i];   // the reference adds the centre and subtracts it again (render_ops.py:22-23)
        }
        const float z = __ldg(pp.que_depth + n);
        px = fmaf(d[0], z, cam[9]); py = fmaf(d[1], z, cam[10]); pz = fmaf(d[2], z, cam[11]);
        const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        qx = -d[0] / nrm; qy = -d[1] / nrm; qz = -d[2] / nrm;
        const float a = -1.f / cam[21], bb = -1.f / cam[22];
        const float tcur = (-1.f / z - a) / (bb - a);
        float dc = 1e6f;
        if (s + 1 < dn) dc = (-1.f / __ldg(pp.que_depth + n + 1) - a) / (bb - a) - tcur;
        float dp = dc;
        if (s > 0) dp = tcur - (-1.f / __ldg(pp.que_depth + n - 1) - a) / (bb - a);
        ihc = dc * 0.5f; ihp = dp * 0.5f;
      }
      g[P_X * LDP + lane] = px; g[P_Y * LDP + lane] = py; g[P_Z * LDP + lane] = pz;
      g[P_QX * LDP + lane] = qx; g[P_QY * LDP + lane] = qy; g[P_QZ * LDP + lane] = qz;
      g[P_IHP * LDP + lane] = ihp; g[P_IHC * LDP + lane] = ihc;
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(pfull + buf);
    };
    uint32_t i = 0;
    int j = 0;
    if (blockIdx.x < kp.n_tiles) geometry(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < kp.n_tiles; tile += gridDim.x, ++j) {
      if (tile + int(gridDim.x) < kp.n_tiles) geometry(tile + gridDim.x, j + 1);
      if (lane == 0) {
        for (int s = 0; s < stages_per_tile; ++s, ++i) {
          // stage sequence: heads [0,n_heads) | pe0 | pe1 | base0 slab 0,1,2 | base1 | vis0+vis1 | vis_fc2.0+rgb_fc.0
          int src, bytes = RING_STAGE * 4;
          if (s < n_heads) src = tcl::HEAD0 + s * RING_STAGE;
          else {
            const int t = s - n_heads;
            src = t == 0 ? tcl::PE0 : t == 1 ? tcl::PE1 : t <= 4 ? tcl::B0 + (t - 2) * RING_STAGE : t == 5 ? tcl::B1 : t == 6 ? tcl::V01 : tcl::V2R;
            if (t == 1) bytes = 3072 * 4;
          }
          const uint32_t buf = i % NBUF;
          if (i >= NBUF) tc::mbar_wait(wempty + buf, ((i / NBUF) - 1) & 1);
          tc::mbar_arrive_expect_tx(wfull + buf, bytes);
          tc::bulk_g2s(ring + buf * RING_STAGE, pp.w_tc + src, bytes, wfull + buf);
        }
      }
      __syncwarp();
    }
  } else {
    // ---------------- compute warps ----------------
    pk::Ctx c;
    c.sm = arena;
    c.wbuf = smem + OFF_WBUF;
    float* const wbufA = smem + OFF_WBUF;
    float* const wbufB = smem + OFF_WBUF2;
    c.tid = tid; c.lane = lane; c.warp = warp;

    float* const tS = arena + C_SCAL * LD;
    float* const tA = arena + C_A * LD;
    float* const tRF = arena + C_RF * LD;
    float* const pnv = tS + S_PT0 * LD;          // per-point #valid views
    float* const tPP = arena + C_PP * LD;
    float* const tGLOB = tPP + PP_GLOB;
    float* const tG = tPP + PP_G;
    float* const tGVEC = tPP + PP_GVEC;
    float* const tGHID = tPP + PP_GHID;
    float* const tGOUT = tPP + PP_GOUT;

    const int fh = pp.fh, fw = pp.fw, h = pp.h, w = pp.w;
    const bool feat_align = (fh == h && fw == w);
    const int r = tid;
    const bool row_ok = r < ROWS;
    const int v = row_ok ? r / P : 0;
    const int p = row_ok ? r - v * P : 0;   // padding rows (r >= ROWS) alias point 0: they only ever produce unused values

    Blk b;
    b.blk = warp >> 2;
    b.leader = (tid & 127) == 0;
    b.issuer_warp = (__shfl_sync(0xffffffffu, warp, 0) & 3) == 0;
    {
      const uint32_t col0 = tmem_base + 256 * b.blk;
      const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
      b.mAhi = col0 + T_AHI; b.mAlo = col0 + T_ALO; b.mD = col0 + T_D;
      b.tAhi = b.mAhi + lane_off; b.tAlo = b.mAlo + lane_off; b.tD = b.mD + lane_off;
    }
    b.mma_bar = mma_bars + b.blk;
    b.wfull = wfull; b.wempty = wempty;
    b.ring_addr = tc::smem_u32(ring);
    b.phase = 0;
    b.wi = 0;
    b.tk = nullptr;
    b.prod = nullptr;

    int tile_it = 0;
#define NR_TICK(id)                                                                                   \
  if (kp.timing != nullptr && blockIdx.x == 0 && (tid & 127) == 0 && tile_it < 64)                    \
    kp.timing[(tile_it * 2 + (tid >> 7)) * 32 + (id)] = clock64();
    for (int tile = blockIdx.x; tile < kp.n_tiles; tile += gridDim.x, ++tile_it) {
      const int n0 = tile * P;
      sync_compute();
      NR_TICK(0)
      // SIMT-layer weights of this tile: hoisted base_fc.0 halves now (group 1 -> buffer A, group 2 -> buffer B)
      stage_async(wbufA, W + lay::HOIST_W, 72 * 64, tid);
      stage_async(wbufB, W + lay::HOIST_W + 72 * 64, 68 * 64 + 64, tid);

      // ---------------- phase 0: per-point ray geometry comes from the producer warp, one tile ahead ----------------
      if (tile_it >= 1 && tid == 0) tc::mbar_arrive(pempty + ((tile_it - 1) & 1));   // previous tile's buffer is free
      tc::mbar_wait(pfull + (tile_it & 1), (tile_it >> 1) & 1);
      const float* __restrict__ parr = geo + (tile_it & 1) * 8 * LDP;
      NR_TICK(1)

      // ---------------- phase 1: projection into the row's view + rgb taps ----------------
      float dbg_px = 0.f, dbg_py = 0.f, dbg_dir[3] = {0.f, 0.f, 0.f};
      float mrow = 0.f, zrow = 1.f, dd[4] = {0.f, 0.f, 0.f, 0.f};
      if (row_ok) {
        const float* __restrict__ vp = pp.view_params + v * 20;
        const float X = parr[P_X * LDP + p], Y = parr[P_Y * LDP + p], Z = parr[P_Z * LDP + p];
        const float xh = fmaf(__ldg(vp + 2), Z, fmaf(__ldg(vp + 1), Y, __ldg(vp + 0) * X)) + __ldg(vp + 3);
        const float yh = fmaf(__ldg(vp + 6), Z, fmaf(__ldg(vp + 5), Y, __ldg(vp + 4) * X)) + __ldg(vp + 7);
        float zh = fmaf(__ldg(vp + 10), Z, fmaf(__ldg(vp + 9), Y, __ldg(vp + 8) * X)) + __ldg(vp + 11);
        const bool degenerate = fabsf(zh) < 1e-4f;
        if (degenerate) zh = 1e-3f;
        const float ux = xh / zh, uy = yh / zh;
        const bool outside = (ux < -0.5f) || (ux >= float(w) - 0.5f) || (uy < -0.5f) || (uy >= float(h) - 0.5f);
        const bool valid = (n0 + p < N) && !degenerate && !outside;
        mrow = valid ? 1.f : 0.f;
        zrow = zh;
        const float dx = X - __ldg(vp + 12), dy = Y - __ldg(vp + 13), dz = Z - __ldg(vp + 14);
        const float inv = -1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-5f);
        const float ex = dx * inv, ey = dy * inv, ez = dz * inv;
        const float qx = parr[P_QX * LDP + p], qy = parr[P_QY * LDP + p], qz = parr[P_QZ * LDP + p];
        dd[0] = ex - qx; dd[1] = ey - qy; dd[2] = ez - qz; dd[3] = ex * qx + ey * qy + ez * qz;
        at<LD>(tS, S_MASK, r) = mrow;
        at<LD>(tS, S_Z, r) = zh;
        if (DEBUG) { dbg_px = ux; dbg_py = uy; dbg_dir[0] = ex; dbg_dir[1] = ey; dbg_dir[2] = ez; }
        const float gx = ux / float(w - 1) * 2.f - 1.f, gy = uy / float(h - 1) * 2.f - 1.f;
        float fx = feat_align ? (gx + 1.f) / 2.f * float(fw - 1) : ((gx + 1.f) * float(fw) - 1.f) / 2.f;
        float fy = feat_align ? (gy + 1.f) / 2.f * float(fh - 1) : ((gy + 1.f) * float(fh) - 1.f) / 2.f;
        fx = fminf(fmaxf(fx, 0.f), float(fw - 1)); fy = fminf(fmaxf(fy, 0.f), float(fh - 1));
        at<LD>(tS, S_IX, r) = fx; at<LD>(tS, S_IY, r) = fy;
        float cr = 0.f, cg = 0.f, cb = 0.f;
        if (valid) {
          float ix = (gx + 1.f) / 2.f * float(w - 1), iy = (gy + 1.f) / 2.f * float(h - 1);
          ix = fminf(fmaxf(ix, 0.f), float(w - 1)); iy = fminf(fmaxf(iy, 0.f), float(h - 1));
          const float x0f = floorf(ix), y0f = floorf(iy);
          const int x0 = int(x0f), y0 = int(y0f);
          const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
          const float we = ix - x0f, ww = (x0f + 1.f) - ix, ws = iy - y0f, wn = (y0f + 1.f) - iy;
          const float* __restrict__ base = pp.rgb + size_t(v) * h * w * 4;
          const float4 t00 = ldg4(base + (size_t(y0) * w + x0) * 4), t01 = ldg4(base + (size_t(y0) * w + x1) * 4);
          const float4 t10 = ldg4(base + (size_t(y1) * w + x0) * 4), t11 = ldg4(base + (size_t(y1) * w + x1) * 4);
          const float w00 = ww * wn, w01 = we * wn, w10 = ww * ws, w11 = we * ws;
          cr = t00.x * w00 + t01.x * w01 + t10.x * w10 + t11.x * w11;
          cg = t00.y * w00 + t01.y * w01 + t10.y * w10 + t11.y * w11;
          cb = t00.z * w00 + t01.z * w01 + t10.z * w10 + t11.z * w11;
        }
        at<LD>(tS, S_R, r) = cr; at<LD>(tS, S_G, r) = cg; at<LD>(tS, S_B, r) = cb;
        at<LD>(tA, 0, r) = cr; at<LD>(tA, 1, r) = cg; at<LD>(tA, 2, r) = cb;
      }
      sync_compute();
      NR_TICK(2)

      // ---------------- phase 2: 64-channel bilinear gather, 16 lanes x float4 per texel ----------------
      // Each half-warp fetches one row's texels (4 taps x 256 contiguous bytes); four rows are batched so that 16
      // independent 128-bit loads are in flight per lane before the first blend (the un-batched loop exposed one L2
      // round trip per row: 12.9 k cycles per tile, profiles/r1_phase_timing.md).
      {
        const int hw = lane >> 4, l = lane & 15;
        for (int rb = warp * 2 + hw; rb < ROWS; rb += 64) {
          float4 t[4][4];
          float wq[4][4];
          bool on[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int rr = rb + 16 * u;
            on[u] = rr < ROWS && at<LD>(tS, S_MASK, rr < ROWS ? rr : 0) != 0.f;
            if (on[u]) {
              const int vv = rr / P;
              const float ix = at<LD>(tS, S_IX, rr), iy = at<LD>(tS, S_IY, rr);
              const float x0f = floorf(ix), y0f = floorf(iy);
              const int x0 = int(x0f), y0 = int(y0f);
              const int x1 = min(x0 + 1, fw - 1), y1 = min(y0 + 1, fh - 1);
              const float we = ix - x0f, ww = (x0f + 1.f) - ix, ws = iy - y0f, wn = (y0f + 1.f) - iy;
              const float* __restrict__ base = pp.feat + size_t(vv) * fh * fw * 64 + 4 * l;
              t[u][0] = ldg4(base + (size_t(y0) * fw + x0) * 64); t[u][1] = ldg4(base + (size_t(y0) * fw + x1) * 64);
              t[u][2] = ldg4(base + (size_t(y1) * fw + x0) * 64); t[u][3] = ldg4(base + (size_t(y1) * fw + x1) * 64);
              wq[u][0] = ww * wn; wq[u][1] = we * wn; wq[u][2] = ww * ws; wq[u][3] = we * ws;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int rr = rb + 16 * u;
            if (rr < ROWS) {
              float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
              if (on[u]) {
                o.x = t[u][0].x * wq[u][0] + t[u][1].x * wq[u][1] + t[u][2].x * wq[u][2] + t[u][3].x * wq[u][3];
                o.y = t[u][0].y * wq[u][0] + t[u][1].y * wq[u][1] + t[u][2].y * wq[u][2] + t[u][3].y * wq[u][3];
                o.z = t[u][0].z * wq[u][0] + t[u][1].z * wq[u][1] + t[u][2].z * wq[u][2] + t[u][3].z * wq[u][3];
                o.w = t[u][0].w * wq[u][0] + t[u][1].w * wq[u][1] + t[u][2].w * wq[u][2] + t[u][3].w * wq[u][3];
              }
              if (l < 8) {
                at<LD>(tRF, 4 * l + 0, rr) = o.x; at<LD>(tRF, 4 * l + 1, rr) = o.y;
                at<LD>(tRF, 4 * l + 2, rr) = o.z; at<LD>(tRF, 4 * l + 3, rr) = o.w;
              } else {
                const int cc = 3 + 4 * (l - 8);
                at<LD>(tA, cc + 0, rr) = o.x; at<LD>(tA, cc + 1, rr) = o.y;
                at<LD>(tA, cc + 2, rr) = o.z; at<LD>(tA, cc + 3, rr) = o.w;
              }
            }
          }
        }
      }
      sync_compute();
      NR_TICK(3)

      // ---------------- phase 3: dist decoder on the tensor cores ----------------
      // A[0:32] <- ray_feats of this row
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = at<LD>(tRF, c0 + j, r);
        store_a16(b, c0, x);
      }
      NR_TICK(4)
      float hv[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 1
      for (int hd = 0; hd < n_heads; ++hd) {
        const float* __restrict__ hw_ = sw + SW_HEAD + hd * SW_HEAD_STRIDE;
        run_layer<32, 4, 0, 4, 0, 1, 0, 1024 * 4, 0, true, false>(b);           // L0: A[0:32]
        {
          float x[32];
          load_d32(b, 0, x);
          add_bias32(x, hw_);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = elu(x[j]);
          store_a32(b, 32, x);
        }
        run_layer<32, 4, 32, 4, 0, 1, 2048 * 4, 3072 * 4, 0, false, true>(b);   // L1: A[32:64]
        float o0 = hw_[128], o1 = hw_[129];
        {
          float x[32];
          load_d32(b, 0, x);
          add_bias32(x, hw_ + 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 wa = *reinterpret_cast<const float4*>(hw_ + 64 + 4 * q);
            const float4 wb = *reinterpret_cast<const float4*>(hw_ + 96 + 4 * q);
            const float a0 = elu(x[4 * q]), a1 = elu(x[4 * q + 1]), a2 = elu(x[4 * q + 2]), a3 = elu(x[4 * q + 3]);
            o0 = fmaf(wa.w, a3, fmaf(wa.z, a2, fmaf(wa.y, a1, fmaf(wa.x, a0, o0))));
            o1 = fmaf(wb.w, a3, fmaf(wb.z, a2, fmaf(wb.y, a1, fmaf(wb.x, a0, o1))));
          }
        }
        // (explicit selects keep hv in registers: a dynamically indexed array would live in local memory)
        if (hd == 0) { hv[0][0] = o0; hv[0][1] = o1; }
        else if (hd == 1) { hv[1][0] = o0; hv[1][1] = o1; }
        else if (hd == 2) { hv[2][0] = o0; hv[2][1] = o1; }
        else { hv[3][0] = o0; hv[3][1] = o1; }
      }
      NR_TICK(5)
      float hit = 0.f, visib = 0.f;
      {
        const float* __restrict__ vp = pp.view_params + v * 20;
        const float zc = fmaxf(zrow, 1e-5f);
        const float a = __ldg(vp + 15), bb = __ldg(vp + 16);
        const float tz = (-1.f / zc - a) / (bb - a);
        const float lo = tz - parr[P_IHP * LDP + p], hi = tz + parr[P_IHC * LDP + p];
        const float aw = sigmoidf_(hv[2][0]);
        const float vd = pp.use_vis ? sigmoidf_(hv[3][0]) : 1.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float mean = softplusf_(hv[0][i]);
          const float var = softplusf_(hv[1][i]) + pp.var_bias;
          const float c0 = logistic_cdf((lo - mean) * var) * vd, c1 = logistic_cdf((hi - mean) * var) * vd;
          const float mix = i == 0 ? aw : 1.f - aw;
          visib = fmaf(1.f - c0, mix, visib);
          hit = fmaf(c1 - c0, mix, hit);
        }
        visib *= mrow; hit *= mrow;
        if (DEBUG && kp.dbg != nullptr && row_ok && n0 + p < N) {
          float* __restrict__ o = kp.dbg + (size_t(v) * N + n0 + p) * 76;
          o[0] = mrow; o[1] = zrow; o[2] = hit; o[3] = visib; o[4] = dbg_px; o[5] = dbg_py;
          o[6] = dbg_dir[0]; o[7] = dbg_dir[1]; o[8] = dbg_dir[2];
          o[9] = at<LD>(tS, S_R, r); o[10] = at<LD>(tS, S_G, r); o[11] = at<LD>(tS, S_B, r);
          for (int k = 0; k < 32; ++k) o[12 + k] = at<LD>(tRF, k, r);
          for (int k = 0; k < 32; ++k) o[44 + k] = at<LD>(tA, 3 + k, r);
        }
      }

      // ---------------- phase 4: prob_embed (tensor cores), ray_dir_fc, neuray_fc ----------------
      {
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = 0.f;
        x[0] = (hit - 0.5f) * 2.f; x[1] = (visib - 0.5f) * 2.f;
        store_a16(b, 64, x);                                                         // A[64:72] = hit', vis', 0...
      }
      run_layer<32, 5, 0, 4, 64, 1, 0, 2048 * 4, 4096, true, true>(b);            // prob_embed.0: K = 32 + 8
      {
        float x[32];
        load_d32(b, 0, x);
        add_bias32(x, sw + SW_PE0B);
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        store_a32(b, 32, x);
      }
      NR_TICK(6)
      run_layer<32, 4, 32, 4, 0, 1, 0, 1536 * 4, 0, true, true>(b);               // prob_embed.2
      float gate;
      {
        float h8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h8[j] = sw[SW_NF0B + j];
        {
          float x[32];
          load_d32(b, 0, x);
          add_bias32(x, sw + SW_PE1B);                                               // neuray_feat
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float4 wa = *reinterpret_cast<const float4*>(sw + SW_NF0W + j * 8);
            const float4 wb = *reinterpret_cast<const float4*>(sw + SW_NF0W + j * 8 + 4);
            h8[0] = fmaf(wa.x, x[j], h8[0]); h8[1] = fmaf(wa.y, x[j], h8[1]); h8[2] = fmaf(wa.z, x[j], h8[2]); h8[3] = fmaf(wa.w, x[j], h8[3]);
            h8[4] = fmaf(wb.x, x[j], h8[4]); h8[5] = fmaf(wb.y, x[j], h8[5]); h8[6] = fmaf(wb.z, x[j], h8[6]); h8[7] = fmaf(wb.w, x[j], h8[7]);
          }
          store_a32(b, 40, x);                                                       // base_fc.0 input columns 40..71
        }
        gate = sw[SW_NF1B];
#pragma unroll
        for (int j = 0; j < 8; ++j) gate = fmaf(sw[SW_NF1W + j], elu(h8[j]), gate);
      }
      {   // ray_dir_fc 4 -> 16 -> 35, rgb_feat = [rgb | img_feats] + it ; -> arena (for the view reductions) and A[0:40]
        float h16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float* __restrict__ w0 = sw + SW_RD0W;
          h16[j] = elu(fmaf(w0[48 + j], dd[3], fmaf(w0[32 + j], dd[2], fmaf(w0[16 + j], dd[1], fmaf(w0[j], dd[0], sw[SW_RD0B + j])))));
        }
        float o[40];
#pragma unroll
        for (int j4 = 0; j4 < 36; j4 += 4) {
          float4 acc = *reinterpret_cast<const float4*>(sw + SW_RD1B + j4);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float4 wv = *reinterpret_cast<const float4*>(sw + SW_RD1W + k * 36 + j4);
            acc.x = fmaf(wv.x, h16[k], acc.x); acc.y = fmaf(wv.y, h16[k], acc.y); acc.z = fmaf(wv.z, h16[k], acc.z); acc.w = fmaf(wv.w, h16[k], acc.w);
          }
          o[j4] = elu(acc.x); o[j4 + 1] = elu(acc.y); o[j4 + 2] = elu(acc.z); o[j4 + 3] = elu(acc.w);
        }
#pragma unroll
        for (int j = 0; j < 35; ++j) {
          o[j] += at<LD>(tA, j, r);
          at<LD>(tA, j, r) = o[j];
        }
#pragma unroll
        for (int j = 35; j < 40; ++j) o[j] = 0.f;
        store_a16(b, 0, o);
        store_a16(b, 16, o + 16);
        {
          uint32_t hi8[8], lo8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) tc::split_tf32(o[32 + j], hi8[j], lo8[j]);
          tc::tmem_st8(b.tAhi + 32, hi8);
          tc::tmem_st8(b.tAlo + 32, lo8);
        }
      }
      NR_TICK(7)
      sync_compute();
      {
        float msum = 0.f;
        for (int vv = 0; vv < rfn; ++vv) msum += at<LD>(tS, S_MASK, vv * P + p);
        const float w1 = mrow / (msum + 1e-8f);
        at<LD>(tS, S_W1, r) = w1;
        at<LD>(tS, S_W0, r) = sigmoidf_(gate) * w1;
        if (v == 0 && row_ok) pnv[p] = msum;
      }
      sync_compute();

      NR_TICK(8)
      // ---------------- phase 5: weighted mean/var over views of rgb_feat, twice ----------------
      for (int it = tid; it < P * 35; it += NT) {
        const int f = it / P, q = it - f * P;
        float m0 = 0.f, m1 = 0.f;
#pragma unroll 8
        for (int vv = 0; vv < rfn; ++vv) {
          const int rr = vv * P + q;
          const float x = at<LD>(tA, f, rr);
          m0 = fmaf(x, at<LD>(tS, S_W0, rr), m0);
          m1 = fmaf(x, at<LD>(tS, S_W1, rr), m1);
        }
        float v0 = 0.f, v1 = 0.f;
#pragma unroll 8
        for (int vv = 0; vv < rfn; ++vv) {
          const int rr = vv * P + q;
          const float x = at<LD>(tA, f, rr);
          v0 = fmaf(at<LD>(tS, S_W0, rr), (x - m0) * (x - m0), v0);
          v1 = fmaf(at<LD>(tS, S_W1, rr), (x - m1) * (x - m1), v1);
        }
        at<LDP>(tGLOB, f, q) = m0; at<LDP>(tGLOB, 35 + f, q) = v0;
        at<LDP>(tGLOB, 70 + f, q) = m1; at<LDP>(tGLOB, 105 + f, q) = v1;
      }

      NR_TICK(9)
      // ---------------- phase 6: hoisted base_fc.0 on the 140 view-invariant inputs (SIMT, per point) ----------------
      // 32 points x 64 outputs x K 140: every thread takes a 4-point x 8-output register tile over one quarter of K
      // (K split 36|36|36|32 across the four warp pairs), then the three upper partial sums are added in a fixed order
      // (deterministic).  The previous 64-thread version was 11.7 k cycles of the 96 k-cycle tile.
      {
        Frag<64, 4, 2> f;
        const int ks = tid >> 6;                      // K quarter (warp-uniform)
        f.r0 = ((tid & 63) >> 3) * 4;                 // points r0..r0+3
        f.ja = (tid & 7) * 4;
        f.jb = 32 + f.ja;
        f.zero();
        stage_wait<0>();                              // groups 1,2: both halves of W_hoist^T have landed
        sync_compute();                               // (also: the view reductions above are complete)
        if (f.r0 < P) {
          if (ks == 0) f.mac<36, LDP>(tGLOB, 0, wbufA);
          else if (ks == 1) f.mac<36, LDP>(tGLOB, 36, wbufA + 36 * 64);
          else if (ks == 2) f.mac<36, LDP>(tGLOB, 72, wbufB);
          else f.mac<32, LDP>(tGLOB, 108, wbufB + 36 * 64);
        }
        float hb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) hb[j] = wbufB[68 * 64 + (j < 4 ? f.ja : f.jb - 4) + j];
        sync_compute();                               // all reads of the weight buffers and of GLOB are done
        float* const part = wbufA;                    // partials of K quarters 1,2 -> wbufA, quarter 3 -> the dead GLOB tile
        if (f.r0 < P && ks > 0) {
          float* dst = ks == 3 ? tGLOB : part + (ks - 1) * 64 * LDP;
          f.store([&](int col, int r4, float4 v4) { at4<LDP>(dst, col, r4) = v4; });
        }
        sync_compute();
        if (f.r0 < P && ks == 0) {
          int jj = 0;
          f.store([&](int col, int r4, float4 v4) {
            const float4 p1 = at4<LDP>(part, col, r4), p2 = at4<LDP>(part + 64 * LDP, col, r4), p3 = at4<LDP>(tGLOB, col, r4);
            const float bv = hb[jj++];
            at4<LDP>(tG, col, r4) = make_float4(((v4.x + p1.x) + p2.x) + p3.x + bv, ((v4.y + p1.y) + p2.y) + p3.y + bv,
                                                ((v4.z + p1.z) + p2.z) + p3.z + bv, ((v4.w + p1.w) + p2.w) + p3.w + bv);
          });
        }
      }
      sync_compute();
      stage_async(wbufA, W + lay::GRP_D2 + lay::GEO0_W, 65 * 64 + 64, tid);      // group 3: geometry_fc.0 -> buffer A
      stage_async(wbufB, W + lay::GRP_D2 + lay::GEO1_W, 64 * 16 + 16, tid);      // group 4: geometry_fc.2 -> buffer B

      NR_TICK(10)
      // ---------------- phase 7: base_fc on the tensor cores ----------------
      run_layer<64, 9, 0, 9, 0, 3, 0, 2048 * 4, 0, true, true>(b);               // base_fc.0: K = 72, three ring stages
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float x[32];
        load_d32(b, c0, x);
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = elu(x[j] + at<LDP>(tG, c0 + j, p));
        store_a32(b, c0, x);
      }
      NR_TICK(11)
      run_layer<32, 8, 0, 8, 0, 1, 0, 2048 * 4, 4096, true, true>(b);            // base_fc.2: K = 64
      float xr[32];
      load_d32(b, 0, xr);
      add_bias32(xr, sw + SW_B1B);
#pragma unroll
      for (int j = 0; j < 32; ++j) xr[j] = elu(xr[j]);
      store_a32(b, 0, xr);

      NR_TICK(12)
      // ---------------- phase 8: vis_fc, vis_fc2, rgb_fc ----------------
      const float w1row = at<LD>(tS, S_W1, r);
      if (kp.timing != nullptr && blockIdx.x == 0 && tile_it < 64) b.tk = kp.timing + (tile_it * 2 + (tid >> 7)) * 32 + 22;
      run_layer<32, 4, 0, 4, 0, 1, 0, 1024 * 4, 0, true, false>(b);               // vis_fc.0 (row scale folded into the epilogue)
      b.tk = nullptr;
      float lg = sw[SW_V1LB];
      {
        float x[32];
        load_d32(b, 0, x);
        NR_TICK(29)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = *reinterpret_cast<const float4*>(sw + SW_V0B + 4 * q);
          const float4 wl = *reinterpret_cast<const float4*>(sw + SW_V1LW + 4 * q);
          x[4 * q] = elu(fmaf(w1row, x[4 * q], bb.x)); x[4 * q + 1] = elu(fmaf(w1row, x[4 * q + 1], bb.y));
          x[4 * q + 2] = elu(fmaf(w1row, x[4 * q + 2], bb.z)); x[4 * q + 3] = elu(fmaf(w1row, x[4 * q + 3], bb.w));
          lg = fmaf(wl.w, x[4 * q + 3], fmaf(wl.z, x[4 * q + 2], fmaf(wl.y, x[4 * q + 1], fmaf(wl.x, x[4 * q], lg))));
        }
        NR_TICK(28)
        store_a32(b, 32, x);
      }
      NR_TICK(13)
      const float visa = sigmoidf_(elu(lg)) * mrow;
      run_layer<32, 4, 32, 4, 0, 1, 2048 * 4, 3072 * 4, 0, false, true>(b);       // vis_fc.2 outputs 0..31 (residual)
      {
        float x[32];
        load_d32(b, 0, x);
        add_bias32(x, sw + SW_V1B);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          xr[j] += elu(x[j]);
          at<LD>(tRF, j, r) = xr[j];                                                 // x for the second view reduction
        }
        store_a32(b, 0, xr);
      }
      NR_TICK(14)
      run_layer<32, 4, 0, 4, 0, 1, 0, 1024 * 4, 0, true, false>(b);               // vis_fc2.0
      float vis2;
      {
        float l2 = sw[SW_V21B];
        {
          float x[32];
          load_d32(b, 0, x);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bb = *reinterpret_cast<const float4*>(sw + SW_V20B + 4 * q);
            const float4 wl = *reinterpret_cast<const float4*>(sw + SW_V21W + 4 * q);
            l2 = fmaf(wl.x, elu(fmaf(visa, x[4 * q], bb.x)), l2); l2 = fmaf(wl.y, elu(fmaf(visa, x[4 * q + 1], bb.y)), l2);
            l2 = fmaf(wl.z, elu(fmaf(visa, x[4 * q + 2], bb.z)), l2); l2 = fmaf(wl.w, elu(fmaf(visa, x[4 * q + 3], bb.w)), l2);
          }
        }
        vis2 = sigmoidf_(l2) * mrow;
        at<LD>(tS, S_VIS2, r) = vis2;
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = 0.f;
        x[0] = vis2; x[1] = dd[0]; x[2] = dd[1]; x[3] = dd[2]; x[4] = dd[3];
        store_a16(b, 32, x);                                                         // A[32:40] = vis, ray_diff, 0
      }
      NR_TICK(15)
      run_layer<16, 5, 0, 4, 32, 1, 2048 * 4, 3072 * 4, 2048, false, true>(b);    // rgb_fc.0: K = 32 + 8, N = 16
      {
        float x[16];
        load_d16(b, 0, x);
        float h8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h8[j] = sw[SW_RGB1B + j];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float a = elu(x[k] + sw[SW_RGB0B + k]);
          const float4 wa = *reinterpret_cast<const float4*>(sw + SW_RGB1W + k * 8);
          const float4 wb = *reinterpret_cast<const float4*>(sw + SW_RGB1W + k * 8 + 4);
          h8[0] = fmaf(wa.x, a, h8[0]); h8[1] = fmaf(wa.y, a, h8[1]); h8[2] = fmaf(wa.z, a, h8[2]); h8[3] = fmaf(wa.w, a, h8[3]);
          h8[4] = fmaf(wb.x, a, h8[4]); h8[5] = fmaf(wb.y, a, h8[5]); h8[6] = fmaf(wb.z, a, h8[6]); h8[7] = fmaf(wb.w, a, h8[7]);
        }
        float l3 = sw[SW_RGB2B];
#pragma unroll
        for (int j = 0; j < 8; ++j) l3 = fmaf(sw[SW_RGB2W + j], elu(h8[j]), l3);
        at<LD>(tS, S_LOGIT, r) = mrow == 0.f ? -1e9f : l3;
      }
      NR_TICK(16)
      sync_compute();
      {
        float s = 0.f;
        for (int vv = 0; vv < rfn; ++vv) s += at<LD>(tS, S_VIS2, vv * P + p);
        at<LD>(tS, S_W2, r) = vis2 / (s + 1e-8f);
      }
      sync_compute();

      NR_TICK(17)
      // ---------------- phase 9: per-point softmax blend + second weighted mean/var ----------------
      if (tid < P) {
        const int q = tid;
        float mx = -3.4e38f;
        for (int vv = 0; vv < rfn; ++vv) mx = fmaxf(mx, at<LD>(tS, S_LOGIT, vv * P + q));
        float den = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
        for (int vv = 0; vv < rfn; ++vv) {
          const int rr = vv * P + q;
          const float e = expf(at<LD>(tS, S_LOGIT, rr) - mx);
          den += e;
          cr = fmaf(e, at<LD>(tS, S_R, rr), cr); cg = fmaf(e, at<LD>(tS, S_G, rr), cg); cb = fmaf(e, at<LD>(tS, S_B, rr), cb);
        }
        at<LDP>(tGOUT, 16, q) = cr / den; at<LDP>(tGOUT, 17, q) = cg / den; at<LDP>(tGOUT, 18, q) = cb / den;
        at<LDP>(tGOUT, 19, q) = pnv[q];
      }
      for (int it = tid; it < P * 33; it += NT) {
        const int f = it / P, q = it - f * P;
        if (f < 32) {
          float m = 0.f;
#pragma unroll 8
          for (int vv = 0; vv < rfn; ++vv) m = fmaf(at<LD>(tRF, f, vv * P + q), at<LD>(tS, S_W2, vv * P + q), m);
          float vr = 0.f;
#pragma unroll 8
          for (int vv = 0; vv < rfn; ++vv) {
            const float x = at<LD>(tRF, f, vv * P + q);
            vr = fmaf(at<LD>(tS, S_W2, vv * P + q), (x - m) * (x - m), vr);
          }
          at<LDP>(tGVEC, f, q) = m; at<LDP>(tGVEC, 32 + f, q) = vr;
        } else {
          float s = 0.f;
          for (int vv = 0; vv < rfn; ++vv) s += at<LD>(tS, S_W2, vv * P + q);
          at<LDP>(tGVEC, 64, q) = s / float(rfn);
        }
      }

      NR_TICK(18)
      // ---------------- phase 10: geometry_fc per point (SIMT, K split across warp pairs / warps) ----------------
      stage_wait<0>();                         // groups 3,4
      sync_compute();
      {
        Frag<64, 4, 2> f;                      // geometry_fc.0: 65 -> 64, K split 20|16|16|13
        const int ks = tid >> 6;
        f.r0 = ((tid & 63) >> 3) * 4;
        f.ja = (tid & 7) * 4;
        f.jb = 32 + f.ja;
        f.zero();
        if (f.r0 < P) {
          if (ks == 0) f.mac<20, LDP>(tGVEC, 0, wbufA);
          else if (ks == 1) f.mac<16, LDP>(tGVEC, 20, wbufA + 20 * 64);
          else if (ks == 2) f.mac<16, LDP>(tGVEC, 36, wbufA + 36 * 64);
          else f.mac<13, LDP>(tGVEC, 52, wbufA + 52 * 64);
        }
        float gb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) gb[j] = wbufA[65 * 64 + (j < 4 ? f.ja : f.jb - 4) + j];
        sync_compute();                        // weights (wbufA) and GVEC fully consumed
        float* const part = wbufA;             // quarters 1,2 -> wbufA, quarter 3 -> the dead GVEC tile
        if (f.r0 < P && ks > 0) {
          float* dst = ks == 3 ? tGVEC : part + (ks - 1) * 64 * LDP;
          f.store([&](int col, int r4, float4 v4) { at4<LDP>(dst, col, r4) = v4; });
        }
        sync_compute();
        if (f.r0 < P && ks == 0) {
          int jj = 0;
          f.store([&](int col, int r4, float4 v4) {
            const float4 p1 = at4<LDP>(part, col, r4), p2 = at4<LDP>(part + 64 * LDP, col, r4), p3 = at4<LDP>(tGVEC, col, r4);
            const float bv = gb[jj++];
            at4<LDP>(tGHID, col, r4) = elu4(make_float4(((v4.x + p1.x) + p2.x) + p3.x + bv, ((v4.y + p1.y) + p2.y) + p3.y + bv,
                                                         ((v4.z + p1.z) + p2.z) + p3.z + bv, ((v4.w + p1.w) + p2.w) + p3.w + bv));
          });
        }
      }
      NR_TICK(19)
      sync_compute();
      {
        Frag<16, 4, 1> f;                      // geometry_fc.2: 64 -> 16, one K slice of 8 per warp
        const int ks = warp;
        f.r0 = (lane >> 2) * 4;
        f.ja = (lane & 3) * 4;
        f.jb = 0;
        f.zero();
        if (f.r0 < P) f.mac<8, LDP>(tGHID, 8 * ks, wbufB + 8 * ks * 16);
        const float4 gb = *reinterpret_cast<const float4*>(wbufB + 64 * 16 + f.ja);
        float* const part = wbufA;             // 7 partial tiles [16][LDP]
        if (f.r0 < P && ks > 0) f.store([&](int col, int r4, float4 v4) { at4<LDP>(part + (ks - 1) * 16 * LDP, col, r4) = v4; });
        sync_compute();
        if (f.r0 < P && ks == 0) {
          int jj = 0;
          f.store([&](int col, int r4, float4 v4) {
            float4 acc = v4;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
              const float4 pk_ = at4<LDP>(part + k * 16 * LDP, col, r4);
              acc.x += pk_.x; acc.y += pk_.y; acc.z += pk_.z; acc.w += pk_.w;
            }
            const float bv = jj == 0 ? gb.x : jj == 1 ? gb.y : jj == 2 ? gb.z : gb.w;
            ++jj;
            at4<LDP>(tGOUT, col, r4) = elu4(make_float4(acc.x + bv, acc.y + bv, acc.z + bv, acc.w + bv));
          });
        }
      }
      sync_compute();
      NR_TICK(20)
      {
        const int cnt = min(P, N - n0) * REC;
        float* __restrict__ dst = pp.point_rec + size_t(n0) * REC;
        for (int i = tid; i < cnt; i += NT) {
          const int q = i / REC, cc = i - q * REC;
          dst[i] = at<LDP>(tGOUT, cc, q);
        }
      }
      NR_TICK(21)
    }
  }

  // ---------------- teardown ----------------
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

#include "nr_point_kernel_pm.cuh"
#include "nr_point_kernel_pm3.cuh"

}  // namespace pkt

static long long* g_timing = nullptr;   // diagnostics only (nr_point_kernel_timing)
void set_point_kernel_timing(long long* buf) { g_timing = buf; }


template <int G>
static int launch_pm(const pkt::KParams& kp0, int sms, cudaStream_t stream) {
  pkt::KParams kp = kp0;
  constexpr int PB = 128 / G;
  const long long N = (long long)kp.p.rn * kp.p.dn;
  kp.P = PB;
  kp.n_tiles = int((N + PB - 1) / PB);
  const int pairs = (kp.n_tiles + pkt::pm::NBLK - 1) / pkt::pm::NBLK;
  const int grid = pairs < sms ? pairs : sms;
  static bool attr_done[2] = {false, false};
  if (kp.dbg) {
    if (!attr_done[1]) {
      cudaFuncSetAttribute(pkt::pm::point_kernel_pm<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pkt::pm::SMEM_BYTES));
      attr_done[1] = true;
    }
    pkt::pm::point_kernel_pm<G, true><<<grid, pkt::pm::NTHR, pkt::pm::SMEM_BYTES, stream>>>(kp);
  } else {
    if (!attr_done[0]) {
      cudaFuncSetAttribute(pkt::pm::point_kernel_pm<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pkt::pm::SMEM_BYTES));
      attr_done[0] = true;
    }
    pkt::pm::point_kernel_pm<G, false><<<grid, pkt::pm::NTHR, pkt::pm::SMEM_BYTES, stream>>>(kp);
  }
  NR_CHECK_LAUNCH("point_kernel_pm");
  return NR_OK;
}

template <int G>
static int launch_pm3(const pkt::KParams& kp0, int sms, cudaStream_t stream) {
  pkt::KParams kp = kp0;
  constexpr int PB = 128 / G;
  const long long N = (long long)kp.p.rn * kp.p.dn;
  kp.P = PB;
  kp.n_tiles = int((N + PB - 1) / PB);
  const int groups = (kp.n_tiles + pkt::pm3::NBLK - 1) / pkt::pm3::NBLK;
  const int grid = groups < sms ? groups : sms;
  static bool attr_done[2] = {false, false};
  if (kp.dbg) {
    if (!attr_done[1]) {
      cudaFuncSetAttribute(pkt::pm3::point_kernel_pm3<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pkt::pm3::SMEM_BYTES));
      attr_done[1] = true;
    }
    pkt::pm3::point_kernel_pm3<G, true><<<grid, pkt::pm3::NTHR, pkt::pm3::SMEM_BYTES, stream>>>(kp);
  } else {
    if (!attr_done[0]) {
      cudaFuncSetAttribute(pkt::pm3::point_kernel_pm3<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pkt::pm3::SMEM_BYTES));
      attr_done[0] = true;
    }
    pkt::pm3::point_kernel_pm3<G, false><<<grid, pkt::pm3::NTHR, pkt::pm3::SMEM_BYTES, stream>>>(kp);
  }
  NR_CHECK_LAUNCH("point_kernel_pm3");
  return NR_OK;
}

// point-major kernel (nr_point_kernel_pm.cuh); the row-per-(view,point) kernel above stays selectable for A/B runs
int launch_point_kernel_pm(const NrPassParams* p, float* dbg, cudaStream_t stream) {
  pkt::KParams kp;
  kp.p = *p;
  kp.dbg = dbg;
  kp.timing = g_timing;
  kp.n_heads = p->use_vis ? 4 : 3;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // default: three blocks per SM (nr_point_kernel_pm3.cuh); NR_POINT_KERNEL=pm2 keeps the two-block kernel for A/B runs
  const char* sel = getenv("NR_POINT_KERNEL");
  if (!(sel != nullptr && sel[0] == 'p' && sel[1] == 'm' && sel[2] == '2')) {
    if (p->rfn <= 4) return launch_pm3<4>(kp, sms, stream);
    if (p->rfn <= 8) return launch_pm3<8>(kp, sms, stream);
    if (p->rfn <= 16) return launch_pm3<16>(kp, sms, stream);
    return launch_pm3<32>(kp, sms, stream);
  }
  if (p->rfn <= 4) return launch_pm<4>(kp, sms, stream);
  if (p->rfn <= 8) return launch_pm<8>(kp, sms, stream);
  if (p->rfn <= 16) return launch_pm<16>(kp, sms, stream);
  return launch_pm<32>(kp, sms, stream);
}

int launch_point_kernel_tc(const NrPassParams* p, float* dbg, cudaStream_t stream) {
  {
    // default: the point-major kernel; NR_POINT_KERNEL=tc selects the row-per-(view,point) tensor-core kernel and
    // NR_POINT_KERNEL=simt (host side: w_tc = NULL) the fp32 SIMT kernel -- development A/B switches only
    const char* sel = getenv("NR_POINT_KERNEL");
    if (sel == nullptr || !(sel[0] == 't' && sel[1] == 'c')) return launch_point_kernel_pm(p, dbg, stream);
  }
  pkt::KParams kp;
  kp.p = *p;
  kp.dbg = dbg;
  kp.timing = g_timing;
  kp.P = (pkt::LD / p->rfn) & ~3;
  if (kp.P > pkt::LDP) kp.P = pkt::LDP;
  const long long N = (long long)p->rn * p->dn;
  kp.n_tiles = int((N + kp.P - 1) / kp.P);
  kp.n_heads = p->use_vis ? 4 : 3;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = kp.n_tiles < sms ? kp.n_tiles : sms;
  static bool attr_done[2] = {false, false};
  if (dbg) {
    if (!attr_done[1]) {
      cudaFuncSetAttribute(pkt::point_kernel_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pkt::SMEM_BYTES));
      attr_done[1] = true;
    }
    pkt::point_kernel_tc<true><<<grid, pkt::NTHREADS, pkt::SMEM_BYTES, stream>>>(kp);
  } else {
    if (!attr_done[0]) {
      cudaFuncSetAttribute(pkt::point_kernel_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pkt::SMEM_BYTES));
      attr_done[0] = true;
    }
    pkt::point_kernel_tc<false><<<grid, pkt::NTHREADS, pkt::SMEM_BYTES, stream>>>(kp);
  }
  NR_CHECK_LAUNCH("point_kernel_tc");
  return NR_OK;
}

}  // namespace nr
