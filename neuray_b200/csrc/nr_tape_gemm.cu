// Weight gradients from the backward tapes: for every Linear layer, dW[out][in] = sum over rows of dz[out] x[in] and
// db[out] = sum of dz[out] (one extra "input" that is constantly 1).  All layers of a pass in ONE launch.
//
// The tapes are slot-major (a slot's values for consecutive rows are contiguous), so a K-chunk of 64 rows of a layer's
// operands is n_in + n_out coalesced 256-byte reads.  A CTA takes (layer, every gridDim.y-th chunk): the chunk goes to
// shared memory (row stride 65 floats: no bank conflicts between the operand rows a warp touches), each thread
// accumulates up to three 4x4 output tiles in registers over all its chunks and adds them to the result once.
// cuBLAS runs these shapes (M, N <= 64, K = 262 144) as sgemm_largek at ~320 us each, 26 per pass.
#include "nr_train_math.cuh"

namespace nr {
namespace tg {

constexpr int MAXD = 48;
constexpr int KC = 64, LDK = KC + 1;   // half a tape tile per K chunk; odd row stride: the operand rows a warp touches fall into distinct banks
constexpr int MAXT = 3;      // 4x4 tiles per thread: 16 x 36 tiles at most (64 outputs, 140 + 1 inputs)

struct Args {
  NrGemmDesc d[MAXD];
  const float* tape[4];      // 0 row tape, 1 row gradients, 2 point tape, 3 point gradients
  long long rows[4];
  int slots[4];
  float* out;
};

__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = valid ? 4 : 0;                                   // src-size 0: the destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256) tape_gemm_kernel(const __grid_constant__ Args a) {
  extern __shared__ float sm[];
  const NrGemmDesc& d = a.d[blockIdx.x];
  const int n_in = d.n_in + 1, n_out = d.n_out;                  // + the constant-1 input (bias)
  const float* __restrict__ X = a.tape[d.x_tape] + (long long)d.x_slot * tr::TILE;      // + tile * slots * 128
  const float* __restrict__ Z = a.tape[d.g_tape] + (long long)d.g_slot * tr::TILE;
  const long long xt = (long long)a.slots[d.x_tape] * tr::TILE, zt = (long long)a.slots[d.g_tape] * tr::TILE;   // floats per tape tile
  const long long M = a.rows[d.x_tape];
  const int buf_floats = (n_in + n_out) * LDK;
  const int ti_n = (n_in + 3) >> 2, to_n = (n_out + 3) >> 2, tiles = ti_n * to_n;
  // thread -> (tile, k group): a layer with few output tiles splits the chunk's rows over the spare threads instead of
  // leaving them idle (a 32x33 layer has 72 tiles: three k groups); the partial sums meet in the final atomic adds
  const int kgroups = tiles >= 256 ? 1 : min(256 / tiles, KC);
  const int tile0 = tiles >= 256 ? int(threadIdx.x) : int(threadIdx.x) % tiles;
  const int kg = tiles >= 256 ? 0 : int(threadIdx.x) / tiles;
  const bool active = kg < kgroups;
  float acc[MAXT][16];
#pragma unroll
  for (int t = 0; t < MAXT; ++t)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[t][e] = 0.f;
  // a CTA walks a contiguous range of 64-row chunks (sequential streams per operand slot), double buffered: the
  // cp.async copies of chunk c+1 are in flight while chunk c is multiplied (a plain load-then-compute loop spent most of
  // its time waiting for the loads: one chunk is only ~17 KB)
  const long long chunks = (M + KC - 1) / KC;
  const long long per = (chunks + gridDim.y - 1) / gridDim.y;
  const long long ch0 = (long long)blockIdx.y * per, ch_end = min(chunks, ch0 + per);
  auto stage = [&](long long ch, float* buf) {
    const long long tile = ch >> 1;
    const int half = int(ch & 1) * KC;
    const long long k0 = ch * KC;
    const int kn = int(M - k0 < KC ? M - k0 : KC);
    const float* __restrict__ xsrc = X + tile * xt + half;
    const float* __restrict__ zsrc = Z + tile * zt + half;
    const int n = (n_in - 1 + n_out) * KC;
    for (int idx = threadIdx.x; idx < n; idx += 256) {
      const int row = idx / KC, k = idx - row * KC;
      const float* src = row < n_in - 1 ? xsrc + row * tr::TILE + k : zsrc + (row - (n_in - 1)) * tr::TILE + k;
      cp_async4(buf + (row < n_in - 1 ? row : row + 1) * LDK + k, src, k < kn);       // row n_in - 1 is the constant input
    }
  };
  for (int i = threadIdx.x; i < 2 * KC; i += 256) sm[(i / KC) * buf_floats + (n_in - 1) * LDK + (i % KC)] = 1.f;
  if (ch0 < ch_end) stage(ch0, sm);
  cp_async_commit();
  for (long long ch = ch0; ch < ch_end; ++ch) {
    float* buf = sm + int((ch - ch0) & 1) * buf_floats;
    if (ch + 1 < ch_end) stage(ch + 1, sm + int((ch - ch0 + 1) & 1) * buf_floats);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* xs = buf;                       // [n_in][LDK]
    const float* zs = buf + n_in * LDK;          // [n_out][LDK]
    if (active) {
#pragma unroll
      for (int t = 0; t < MAXT; ++t) {
        const int tile = tile0 + t * 256;
        if (tile < tiles && (t == 0 || tiles > 256)) {
          const int to = tile / ti_n, ti = tile - to * ti_n;
          const float* zr[4];
          const float* xr[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            zr[e] = zs + min(4 * to + e, n_out - 1) * LDK;
            xr[e] = xs + min(4 * ti + e, n_in - 1) * LDK;
          }
          for (int k = kg; k < KC; k += kgroups) {
            float zv[4], xv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { zv[e] = zr[e][k]; xv[e] = xr[e][k]; }
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[t][4 * o + i] = fmaf(zv[o], xv[i], acc[t][4 * o + i]);
          }
        }
      }
    }
    __syncthreads();                             // the buffer is refilled by the prefetch of the iteration after next
  }
  float* out = a.out + d.out_off;
  if (active) {
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int tile = tile0 + t * 256;
      if (tile < tiles && (t == 0 || tiles > 256)) {
        const int to = tile / ti_n, ti = tile - to * ti_n;
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (4 * to + o < n_out && 4 * ti + i < n_in) atomicAdd(out + (4 * to + o) * n_in + 4 * ti + i, acc[t][4 * o + i]);
      }
    }
  }
}

}  // namespace tg
}  // namespace nr

extern "C" int nr_tape_gemms(const NrGemmDesc* descs, int n_desc, const float* tape_row, const float* grad_row, long long rows,
                             const float* tape_point, const float* grad_point, long long points, float* out, void* stream) {
  using namespace nr;
  NR_CHECK_ARG(descs != nullptr && n_desc >= 1 && n_desc <= tg::MAXD, "descriptor count");
  NR_CHECK_ARG(tape_row && grad_row && tape_point && grad_point && out, "null device pointer");
  tg::Args a;
  int max_rows = 0;
  for (int i = 0; i < n_desc; ++i) {
    const NrGemmDesc& d = descs[i];
    NR_CHECK_ARG(d.n_out >= 1 && d.n_out <= 64 && d.n_in >= 1 && d.n_in <= 140 && d.x_tape >= 0 && d.x_tape < 4 && d.g_tape >= 0 && d.g_tape < 4,
                 "descriptor");
    NR_CHECK_ARG(((d.n_out + 3) / 4) * ((d.n_in + 4) / 4) <= tg::MAXT * 256, "layer too large");
    a.d[i] = d;
    if (d.n_in + 1 + d.n_out > max_rows) max_rows = d.n_in + 1 + d.n_out;
  }
  a.tape[0] = tape_row; a.tape[1] = grad_row; a.tape[2] = tape_point; a.tape[3] = grad_point;
  a.rows[0] = rows; a.rows[1] = rows; a.rows[2] = points; a.rows[3] = points;
  a.slots[0] = tr::R_SLOTS; a.slots[1] = tr::G_SLOTS; a.slots[2] = tr::P_SLOTS; a.slots[3] = tr::GP_SLOTS;
  a.out = out;
  const size_t smem = 2 * size_t(max_rows) * tg::LDK * sizeof(float);     // two chunk buffers
  if (smem > 48 * 1024)   // per-device attribute, set per launch: no per-process state
    cudaFuncSetAttribute(tg::tape_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  const long long chunks = (rows + tg::KC - 1) / tg::KC;
  int gy = int(chunks < 48 ? chunks : 48);
  if (gy < 1) gy = 1;
  tg::tape_gemm_kernel<<<dim3(n_desc, gy), 256, smem, (cudaStream_t)stream>>>(a);
  NR_CHECK_LAUNCH("tape_gemms");
  return NR_OK;
}
