// Point kernel: everything of one render_by_depth pass that is computed per (point, reference view) pair and
// the cross-view reductions that follow, fused in one persistent kernel (reference renderer.py:171-178 =
// depth2inv_dists + depth2points + project_points_dict + predict_proj_ray_prob + get_img_feats + the per-view
// part of IBRNetWithNeuRay.forward, ibrnet.py:315-353, + geometry_fc :354).
//
// Work decomposition
//   * a tile = P consecutive points (ray-major, sample-minor) x all rfn views = ROWS = P*rfn <= 256 rows;
//     row r = v*P + p (view-major), one CTA of 256 threads per SM walks tiles persistently
//   * all per-row activations live in shared memory, column-major [feature][256 rows] with an XOR swizzle on
//     the row index (conflict-free for the transposing gather stores, for per-row accesses and for 4-row float4
//     accesses), ~196 KB; weights are staged from L2 per layer group into a 21 KB buffer
//   * dense layers are register-tiled fp32 SIMT GEMMs (4x8 / 8x8 outputs per thread, 128-bit LDS for both
//     operands); tiny heads and odd-sized layers run thread-per-row
//   * the feature gather reads the NHWC-packed maps with one 128-bit load per lane, 16 lanes per texel
//     (256 contiguous bytes = ray_feats|img_feats), 4 taps
//   * output: 20 floats per point (geometry feature 16, blended rgb 3, #valid views 1) for the ray kernel
#include "nr_common.cuh"
#include "nr_point_common.cuh"

namespace nr {
namespace pk {

constexpr int LD = 256;        // rows per activation column
constexpr int LDP = 96;        // rows per per-point column (P <= 84)
constexpr int REC = NR_POINT_REC;

// arena columns (each column = LD floats = 1 KB)
constexpr int C_SCAL = 0;      // 24 columns of per-row scalars + per-point arrays
constexpr int C_A = 24;        // 68: rgb_feat 0..34 | neuray_feat 35..66
constexpr int C_RF = 92;       // 40: ray_feats 0..31 | hit' 32 | vis' 33   -> later x 0..31 | vis2 32 | ray_diff 33..36
constexpr int C_H = 132;       // 64: hidden activations
constexpr int N_COLS = 196;
constexpr int WBUF = 5280;     // floats of staged weights
constexpr int SMEM_FLOATS = N_COLS * LD + WBUF;
constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * 4;

// per-row scalar columns (relative to C_SCAL)
enum { S_MASK = 0, S_Z, S_HIT, S_VIS, S_W1, S_W0, S_VISA, S_VIS2, S_W2, S_DD0, S_DD1, S_DD2, S_DD3, S_R, S_G, S_B,
       S_LOGIT, S_IX, S_IY, S_PT0 /* 19..22: raw per-point arrays */ };
// per-point arrays (raw, stride LDP) inside columns S_PT0..S_PT0+3
enum { P_X = 0, P_Y, P_Z, P_QX, P_QY, P_QZ, P_IHP, P_IHC, P_NV, P_NARR };
static_assert(P_NARR * LDP <= 4 * LD, "per-point arrays overflow their columns");

// per-point tiles (raw float offsets inside the arena, column stride LDP)
constexpr int OFF_G = C_RF * LD;                    // [64][LDP]  hoisted base_fc.0 partial sums
constexpr int OFF_GLOB = C_H * LD;                  // [140][LDP] mean0|var0|mean1|var1
constexpr int OFF_GVEC = C_A * LD;                  // [68][LDP]  mean|var|weight-mean (65 used)
constexpr int OFF_GHID = OFF_GVEC + 68 * LDP;       // [64][LDP]
constexpr int OFF_GOUT = OFF_GHID + 64 * LDP;       // [20][LDP]
static_assert(64 * LDP <= 40 * LD && 140 * LDP <= 64 * LD && OFF_GOUT + 20 * LDP <= (C_A + 68) * LD, "per-point tiles overflow");

struct KParams {
  NrPassParams p;
  float* dbg;       // optional stage tap [rfn][N][76]
  int P;            // points per tile
  int n_tiles;
};

template <bool DEBUG>
__global__ void __launch_bounds__(NT, 1) point_kernel(const KParams kp) {
  extern __shared__ __align__(16) float smem[];
  const NrPassParams& pp = kp.p;
  Ctx c;
  c.sm = smem;
  c.wbuf = smem + N_COLS * LD;
  c.tid = threadIdx.x;
  c.lane = c.tid & 31;
  c.warp = c.tid >> 5;

  float* const tS = smem + C_SCAL * LD;
  float* const tA = smem + C_A * LD;
  float* const tRF = smem + C_RF * LD;
  float* const tH = smem + C_H * LD;
  float* const parr = tS + S_PT0 * LD;   // raw per-point arrays, stride LDP
  float* const tG = smem + OFF_G;
  float* const tGLOB = smem + OFF_GLOB;
  float* const tGVEC = smem + OFF_GVEC;
  float* const tGHID = smem + OFF_GHID;
  float* const tGOUT = smem + OFF_GOUT;

  const int P = kp.P, rfn = pp.rfn, ROWS = P * rfn, dn = pp.dn;
  const int N = pp.rn * dn;
  const int fh = pp.fh, fw = pp.fw, h = pp.h, w = pp.w;
  const float* __restrict__ W = pp.w_point;
  const bool feat_align = (fh == h && fw == w);   // interpolate_feature_map's align_corners rule (render_ops.py:64-68)

  // this thread's row (thread-per-row phases)
  const int r = c.tid;
  const bool row_ok = r < ROWS;
  const int v = row_ok ? r / P : 0;
  const int p = r - v * P;

  for (int tile = blockIdx.x; tile < kp.n_tiles; tile += gridDim.x) {
    const int n0 = tile * P;
    __syncthreads();   // previous tile's output copy is done with the arena

    // ---------------- phase 0: per-point ray geometry (depth2points, depth2inv_dists) ----------------
    if (c.tid < P) {
      const int n = n0 + c.tid;
      float px = 0.f, py = 0.f, pz = 0.f, qx = 0.f, qy = 0.f, qz = 0.f, ihp = 0.f, ihc = 0.f;
      if (n < N) {
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
        const float a = -1.f / cam[21], b = -1.f / cam[22];
        const float tc = (-1.f / z - a) / (b - a);
        float dc = 1e6f;
        if (s + 1 < dn) dc = (-1.f / __ldg(pp.que_depth + n + 1) - a) / (b - a) - tc;
        float dp = dc;
        if (s > 0) dp = tc - (-1.f / __ldg(pp.que_depth + n - 1) - a) / (b - a);
        ihc = dc * 0.5f; ihp = dp * 0.5f;
      }
      parr[P_X * LDP + c.tid] = px; parr[P_Y * LDP + c.tid] = py; parr[P_Z * LDP + c.tid] = pz;
      parr[P_QX * LDP + c.tid] = qx; parr[P_QY * LDP + c.tid] = qy; parr[P_QZ * LDP + c.tid] = qz;
      parr[P_IHP * LDP + c.tid] = ihp; parr[P_IHC * LDP + c.tid] = ihc;
    }
    __syncthreads();

    // ---------------- phase 1: projection into the row's view + rgb taps (project_points_ref_views) ----------------
    float dbg_px = 0.f, dbg_py = 0.f, dbg_dir[3] = {0.f, 0.f, 0.f};
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
      const float m = valid ? 1.f : 0.f;
      // project_points_directions
      const float dx = X - __ldg(vp + 12), dy = Y - __ldg(vp + 13), dz = Z - __ldg(vp + 14);
      const float inv = -1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-5f);
      const float ex = dx * inv, ey = dy * inv, ez = dz * inv;
      const float qx = parr[P_QX * LDP + p], qy = parr[P_QY * LDP + p], qz = parr[P_QZ * LDP + p];
      at<LD>(tS, S_MASK, r) = m;
      at<LD>(tS, S_Z, r) = zh;
      at<LD>(tS, S_DD0, r) = ex - qx; at<LD>(tS, S_DD1, r) = ey - qy; at<LD>(tS, S_DD2, r) = ez - qz;
      at<LD>(tS, S_DD3, r) = ex * qx + ey * qy + ez * qz;
      if (DEBUG) { dbg_px = ux; dbg_py = uy; dbg_dir[0] = ex; dbg_dir[1] = ey; dbg_dir[2] = ez; }
      // grid_sample coordinates (ops.py:29-31 + F.grid_sample unnormalise + border clip)
      const float gx = ux / float(w - 1) * 2.f - 1.f, gy = uy / float(h - 1) * 2.f - 1.f;
      float fx = feat_align ? (gx + 1.f) / 2.f * float(fw - 1) : ((gx + 1.f) * float(fw) - 1.f) / 2.f;
      float fy = feat_align ? (gy + 1.f) / 2.f * float(fh - 1) : ((gy + 1.f) * float(fh) - 1.f) / 2.f;
      fx = fminf(fmaxf(fx, 0.f), float(fw - 1)); fy = fminf(fmaxf(fy, 0.f), float(fh - 1));
      at<LD>(tS, S_IX, r) = fx; at<LD>(tS, S_IY, r) = fy;
      // rgb: full-resolution map, align_corners=True
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
    __syncthreads();

    // ---------------- phase 2: 64-channel bilinear gather, 16 lanes x float4 per texel ----------------
    {
      const int hw = c.lane >> 4, l = c.lane & 15;
      for (int rr = c.warp * 2 + hw; rr < ROWS; rr += 16) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (at<LD>(tS, S_MASK, rr) != 0.f) {
          const int vv = rr / P;
          const float ix = at<LD>(tS, S_IX, rr), iy = at<LD>(tS, S_IY, rr);
          const float x0f = floorf(ix), y0f = floorf(iy);
          const int x0 = int(x0f), y0 = int(y0f);
          const int x1 = min(x0 + 1, fw - 1), y1 = min(y0 + 1, fh - 1);
          const float we = ix - x0f, ww = (x0f + 1.f) - ix, ws = iy - y0f, wn = (y0f + 1.f) - iy;
          const float* __restrict__ base = pp.feat + size_t(vv) * fh * fw * 64 + 4 * l;
          const float4 t00 = ldg4(base + (size_t(y0) * fw + x0) * 64), t01 = ldg4(base + (size_t(y0) * fw + x1) * 64);
          const float4 t10 = ldg4(base + (size_t(y1) * fw + x0) * 64), t11 = ldg4(base + (size_t(y1) * fw + x1) * 64);
          const float w00 = ww * wn, w01 = we * wn, w10 = ww * ws, w11 = we * ws;
          o.x = t00.x * w00 + t01.x * w01 + t10.x * w10 + t11.x * w11;
          o.y = t00.y * w00 + t01.y * w01 + t10.y * w10 + t11.y * w11;
          o.z = t00.z * w00 + t01.z * w01 + t10.z * w10 + t11.z * w11;
          o.w = t00.w * w00 + t01.w * w01 + t10.w * w10 + t11.w * w11;
        }
        if (l < 8) {   // ray_feats channels 4l..4l+3
          at<LD>(tRF, 4 * l + 0, rr) = o.x; at<LD>(tRF, 4 * l + 1, rr) = o.y;
          at<LD>(tRF, 4 * l + 2, rr) = o.z; at<LD>(tRF, 4 * l + 3, rr) = o.w;
        } else {       // img_feats channels -> rgb_feat columns 3..34
          const int cc = 3 + 4 * (l - 8);
          at<LD>(tA, cc + 0, rr) = o.x; at<LD>(tA, cc + 1, rr) = o.y;
          at<LD>(tA, cc + 2, rr) = o.z; at<LD>(tA, cc + 3, rr) = o.w;
        }
      }
    }

    // ---------------- phase 3: dist decoder heads + compute_prob (dist_decoder.py:99-140) ----------------
    float hv[4][2];   // head outputs of this thread's row
    const int n_heads = pp.use_vis ? 4 : 3;
#pragma unroll 1
    for (int hd = 0; hd < n_heads; ++hd) {
      __syncthreads();   // gather stores visible / previous head done with wbuf + H
      stage(c, W + lay::DD_HEAD + hd * lay::DD_HEAD_STRIDE, lay::DD_HEAD_STRIDE);
      __syncthreads();
      {
        Frag<32, 4, 2> f;
        f.setup(c);
        if (f.r0 < ROWS) {
          f.init_bias(c.wbuf + lay::DD_L0_B);
          f.mac<32, LD>(tRF, 0, c.wbuf + lay::DD_L0_W);
          f.store([&](int col, int r4, float4 v4) { at4<LD>(tH, col, r4) = elu4(v4); });
        }
      }
      __syncthreads();
      {
        Frag<32, 4, 2> f;
        f.setup(c);
        if (f.r0 < ROWS) {
          f.init_bias(c.wbuf + lay::DD_L1_B);
          f.mac<32, LD>(tH, 0, c.wbuf + lay::DD_L1_W);
          f.store([&](int col, int r4, float4 v4) { at4<LD>(tH, 32 + col, r4) = elu4(v4); });
        }
      }
      __syncthreads();
      if (row_ok) {
        float o0 = c.wbuf[lay::DD_L2_B], o1 = c.wbuf[lay::DD_L2_B + 1];
        const float* __restrict__ w2 = c.wbuf + lay::DD_L2_W;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float a = at<LD>(tH, 32 + k, r);
          o0 = fmaf(w2[k], a, o0);
          o1 = fmaf(w2[32 + k], a, o1);
        }
        hv[hd][0] = o0; hv[hd][1] = o1;
      }
    }
    if (row_ok) {
      const float* __restrict__ vp = pp.view_params + v * 20;
      const float m = at<LD>(tS, S_MASK, r);
      const float zc = fmaxf(at<LD>(tS, S_Z, r), 1e-5f);
      const float a = __ldg(vp + 15), b = __ldg(vp + 16);
      const float tz = (-1.f / zc - a) / (b - a);
      const float lo = tz - parr[P_IHP * LDP + p], hi = tz + parr[P_IHC * LDP + p];
      const float aw = sigmoidf_(hv[2][0]);
      const float vd = pp.use_vis ? sigmoidf_(hv[3][0]) : 1.f;
      float visib = 0.f, hit = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float mean = softplusf_(hv[0][i]);
        const float var = softplusf_(hv[1][i]) + pp.var_bias;
        const float c0 = logistic_cdf((lo - mean) * var) * vd, c1 = logistic_cdf((hi - mean) * var) * vd;
        const float mix = i == 0 ? aw : 1.f - aw;
        visib = fmaf(1.f - c0, mix, visib);
        hit = fmaf(c1 - c0, mix, hit);
      }
      visib *= m; hit *= m;
      at<LD>(tS, S_HIT, r) = hit; at<LD>(tS, S_VIS, r) = visib;
      at<LD>(tRF, 32, r) = (hit - 0.5f) * 2.f;      // aggregate_net.py:47-48
      at<LD>(tRF, 33, r) = (visib - 0.5f) * 2.f;
      if (DEBUG && kp.dbg != nullptr && n0 + p < N) {
        float* __restrict__ o = kp.dbg + (size_t(v) * N + n0 + p) * 76;
        o[0] = m; o[1] = at<LD>(tS, S_Z, r); o[2] = hit; o[3] = visib; o[4] = dbg_px; o[5] = dbg_py;
        o[6] = dbg_dir[0]; o[7] = dbg_dir[1]; o[8] = dbg_dir[2];
        o[9] = at<LD>(tS, S_R, r); o[10] = at<LD>(tS, S_G, r); o[11] = at<LD>(tS, S_B, r);
        for (int k = 0; k < 32; ++k) o[12 + k] = at<LD>(tRF, k, r);
        for (int k = 0; k < 32; ++k) o[44 + k] = at<LD>(tA, 3 + k, r);
      }
    }

    // ---------------- phase 4: prob_embed, ray_dir_fc, neuray_fc (aggregate_net.py:53, ibrnet.py:325-336) ----------------
    __syncthreads();
    stage(c, W + lay::GRP_B, lay::GRP_B_SIZE);
    __syncthreads();
    {
      Frag<32, 4, 2> f;
      f.setup(c);
      if (f.r0 < ROWS) {
        f.init_bias(c.wbuf + lay::PE0_B);
        f.mac<34, LD>(tRF, 0, c.wbuf + lay::PE0_W);
        f.store([&](int col, int r4, float4 v4) {
          at4<LD>(tH, col, r4) = make_float4(fmaxf(v4.x, 0.f), fmaxf(v4.y, 0.f), fmaxf(v4.z, 0.f), fmaxf(v4.w, 0.f));
        });
      }
    }
    if (row_ok) {   // ray_dir_fc: 4 -> 16 -> 35, added onto [rgb | img_feats]
      const float d0 = at<LD>(tS, S_DD0, r), d1 = at<LD>(tS, S_DD1, r), d2 = at<LD>(tS, S_DD2, r), d3 = at<LD>(tS, S_DD3, r);
      float h16[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float* __restrict__ w0 = c.wbuf + lay::RD0_W;
        h16[j] = elu(fmaf(w0[48 + j], d3, fmaf(w0[32 + j], d2, fmaf(w0[16 + j], d1, fmaf(w0[j], d0, c.wbuf[lay::RD0_B + j])))));
      }
#pragma unroll
      for (int j4 = 0; j4 < 36; j4 += 4) {
        float4 o = *reinterpret_cast<const float4*>(c.wbuf + lay::RD1_B + j4);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float4 wv = *reinterpret_cast<const float4*>(c.wbuf + lay::RD1_W + k * 36 + j4);
          o.x = fmaf(wv.x, h16[k], o.x); o.y = fmaf(wv.y, h16[k], o.y); o.z = fmaf(wv.z, h16[k], o.z); o.w = fmaf(wv.w, h16[k], o.w);
        }
        at<LD>(tA, j4 + 0, r) += elu(o.x); at<LD>(tA, j4 + 1, r) += elu(o.y); at<LD>(tA, j4 + 2, r) += elu(o.z);
        if (j4 + 3 < 35) at<LD>(tA, j4 + 3, r) += elu(o.w);
      }
    }
    __syncthreads();
    {
      Frag<32, 4, 2> f;
      f.setup(c);
      if (f.r0 < ROWS) {
        f.init_bias(c.wbuf + lay::PE1_B);
        f.mac<32, LD>(tH, 0, c.wbuf + lay::PE1_W);
        f.store([&](int col, int r4, float4 v4) { at4<LD>(tA, 35 + col, r4) = v4; });
      }
    }
    __syncthreads();
    if (row_ok) {   // neuray_fc 32 -> 8 -> 1, weight = mask / (sum mask + 1e-8)
      float h8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h8[j] = c.wbuf[lay::NF0_B + j];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float a = at<LD>(tA, 35 + k, r);
        const float4 wa = *reinterpret_cast<const float4*>(c.wbuf + lay::NF0_W + k * 8);
        const float4 wb = *reinterpret_cast<const float4*>(c.wbuf + lay::NF0_W + k * 8 + 4);
        h8[0] = fmaf(wa.x, a, h8[0]); h8[1] = fmaf(wa.y, a, h8[1]); h8[2] = fmaf(wa.z, a, h8[2]); h8[3] = fmaf(wa.w, a, h8[3]);
        h8[4] = fmaf(wb.x, a, h8[4]); h8[5] = fmaf(wb.y, a, h8[5]); h8[6] = fmaf(wb.z, a, h8[6]); h8[7] = fmaf(wb.w, a, h8[7]);
      }
      float gate = c.wbuf[lay::NF1_B];
#pragma unroll
      for (int j = 0; j < 8; ++j) gate = fmaf(c.wbuf[lay::NF1_W + j], elu(h8[j]), gate);
      float msum = 0.f;
      for (int vv = 0; vv < rfn; ++vv) msum += at<LD>(tS, S_MASK, vv * P + p);
      const float w1 = at<LD>(tS, S_MASK, r) / (msum + 1e-8f);
      at<LD>(tS, S_W1, r) = w1;
      at<LD>(tS, S_W0, r) = sigmoidf_(gate) * w1;
      if (v == 0) parr[P_NV * LDP + p] = msum;
    }
    __syncthreads();

    // ---------------- phase 5: weighted mean/var over views of rgb_feat, twice (ibrnet.py:337-339) ----------------
    for (int it = c.tid; it < P * 35; it += NT) {
      const int f = it / P, q = it - f * P;
      float m0 = 0.f, m1 = 0.f;
      for (int vv = 0; vv < rfn; ++vv) {
        const int rr = vv * P + q;
        const float x = at<LD>(tA, f, rr);
        m0 = fmaf(x, at<LD>(tS, S_W0, rr), m0);
        m1 = fmaf(x, at<LD>(tS, S_W1, rr), m1);
      }
      float v0 = 0.f, v1 = 0.f;
      for (int vv = 0; vv < rfn; ++vv) {
        const int rr = vv * P + q;
        const float x = at<LD>(tA, f, rr);
        v0 = fmaf(at<LD>(tS, S_W0, rr), (x - m0) * (x - m0), v0);
        v1 = fmaf(at<LD>(tS, S_W1, rr), (x - m1) * (x - m1), v1);
      }
      at<LDP>(tGLOB, f, q) = m0; at<LDP>(tGLOB, 35 + f, q) = v0;
      at<LDP>(tGLOB, 70 + f, q) = m1; at<LDP>(tGLOB, 105 + f, q) = v1;
    }

    // ---------------- phase 6: base_fc.0 on the 140 view-invariant inputs, once per point ----------------
    {
      Frag<64, 4, 2> f;
      f.setup(c);
      f.zero();
      __syncthreads();
      stage(c, W + lay::HOIST_W, 72 * 64);
      __syncthreads();
      if (f.r0 < P) f.mac<72, LDP>(tGLOB, 0, c.wbuf);
      __syncthreads();
      stage(c, W + lay::HOIST_W + 72 * 64, 68 * 64 + 64);
      __syncthreads();
      if (f.r0 < P) {
        f.mac<68, LDP>(tGLOB, 72, c.wbuf);
        const float* __restrict__ hb = c.wbuf + 68 * 64;
        f.store([&](int col, int r4, float4 v4) {
          const float b = hb[col];
          at4<LDP>(tG, col, r4) = make_float4(v4.x + b, v4.y + b, v4.z + b, v4.w + b);
        });
      }
    }

    // ---------------- phase 7: base_fc on the per-view inputs (ibrnet.py:342-343) ----------------
    __syncthreads();
    stage(c, W + lay::BASE0_W, 67 * 64);
    __syncthreads();
    {
      Frag<64, 8, 2> f;
      f.setup(c);
      if (f.r0 < ROWS) {
        // start from the hoisted partial sums of the rows' points
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int rr = f.r0 + 4 * half;
          const int q = rr % P;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = (j < 4 ? f.ja : f.jb - 4) + j;
            const float4 g = at4<LDP>(tG, col, q);
            f.acc[4 * half + 0][j] = g.x; f.acc[4 * half + 1][j] = g.y; f.acc[4 * half + 2][j] = g.z; f.acc[4 * half + 3][j] = g.w;
          }
        }
        f.mac<67, LD, 1>(tA, 0, c.wbuf);
        f.store([&](int col, int r4, float4 v4) { at4<LD>(tH, col, r4) = elu4(v4); });
      }
    }
    __syncthreads();
    stage(c, W + lay::BASE1_W, 64 * 32 + 32);
    __syncthreads();
    {
      Frag<32, 4, 2> f;
      f.setup(c);
      if (f.r0 < ROWS) {
        f.init_bias(c.wbuf + 64 * 32);
        f.mac<64, LD>(tH, 0, c.wbuf);
        f.store([&](int col, int r4, float4 v4) { at4<LD>(tRF, col, r4) = elu4(v4); });   // x
      }
    }

    // ---------------- phase 8: vis_fc, vis_fc2, rgb_fc (ibrnet.py:345-350, 363-364) ----------------
    __syncthreads();
    stage(c, W + lay::GRP_D1, lay::GRP_D1_SIZE);
    __syncthreads();
    {
      Frag<32, 4, 2> f;   // vis_fc.0 on x*weight: the per-row scale commutes with the matmul
      f.setup(c);
      if (f.r0 < ROWS) {
        f.zero();
        f.mac<32, LD>(tRF, 0, c.wbuf + lay::VIS0_W);
        const float4 s = at4<LD>(tS, S_W1, f.r0);
        f.store([&](int col, int r4, float4 v4) {
          const float b = c.wbuf[lay::VIS0_B + col];
          at4<LD>(tH, col, r4) = elu4(make_float4(fmaf(s.x, v4.x, b), fmaf(s.y, v4.y, b), fmaf(s.z, v4.z, b), fmaf(s.w, v4.w, b)));
        });
      }
    }
    __syncthreads();
    {
      Frag<32, 4, 2> f;   // vis_fc.2 outputs 0..31: residual onto x
      f.setup(c);
      if (f.r0 < ROWS) {
        f.init_bias(c.wbuf + lay::VIS1_B);
        f.mac<32, LD>(tH, 0, c.wbuf + lay::VIS1_W);
        f.store([&](int col, int r4, float4 v4) {
          float4& x = at4<LD>(tRF, col, r4);
          const float4 e = elu4(v4);
          x = make_float4(x.x + e.x, x.y + e.y, x.z + e.z, x.w + e.w);
        });
      }
    }
    if (row_ok) {   // vis_fc.2 output 32 -> vis = sigmoid(.) * mask ; stage ray_diff next to x for rgb_fc
      float lg = c.wbuf[lay::VIS1L_B];
#pragma unroll
      for (int k = 0; k < 32; ++k) lg = fmaf(c.wbuf[lay::VIS1L_W + k], at<LD>(tH, k, r), lg);
      at<LD>(tS, S_VISA, r) = sigmoidf_(elu(lg)) * at<LD>(tS, S_MASK, r);
      at<LD>(tRF, 33, r) = at<LD>(tS, S_DD0, r); at<LD>(tRF, 34, r) = at<LD>(tS, S_DD1, r);
      at<LD>(tRF, 35, r) = at<LD>(tS, S_DD2, r); at<LD>(tRF, 36, r) = at<LD>(tS, S_DD3, r);
    }
    __syncthreads();
    {
      Frag<32, 4, 2> f;   // vis_fc2.0 on x*vis
      f.setup(c);
      if (f.r0 < ROWS) {
        f.zero();
        f.mac<32, LD>(tRF, 0, c.wbuf + lay::V20_W);
        const float4 s = at4<LD>(tS, S_VISA, f.r0);
        f.store([&](int col, int r4, float4 v4) {
          const float b = c.wbuf[lay::V20_B + col];
          at4<LD>(tH, 32 + col, r4) = elu4(make_float4(fmaf(s.x, v4.x, b), fmaf(s.y, v4.y, b), fmaf(s.z, v4.z, b), fmaf(s.w, v4.w, b)));
        });
      }
    }
    __syncthreads();
    if (row_ok) {
      float lg = c.wbuf[lay::V21_B];
#pragma unroll
      for (int k = 0; k < 32; ++k) lg = fmaf(c.wbuf[lay::V21_W + k], at<LD>(tH, 32 + k, r), lg);
      const float v2 = sigmoidf_(lg) * at<LD>(tS, S_MASK, r);
      at<LD>(tS, S_VIS2, r) = v2;
      at<LD>(tRF, 32, r) = v2;
    }
    __syncthreads();
    if (row_ok) {
      float s = 0.f;
      for (int vv = 0; vv < rfn; ++vv) s += at<LD>(tS, S_VIS2, vv * P + p);
      at<LD>(tS, S_W2, r) = at<LD>(tS, S_VIS2, r) / (s + 1e-8f);
    }
    {
      Frag<16, 4, 1> f;   // rgb_fc.0: [x | vis | ray_diff] 37 -> 16
      f.setup(c);
      if (f.r0 < ROWS) {
        f.init_bias(c.wbuf + lay::RGB0_B);
        f.mac<37, LD>(tRF, 0, c.wbuf + lay::RGB0_W);
        f.store([&](int col, int r4, float4 v4) { at4<LD>(tH, col, r4) = elu4(v4); });
      }
    }
    __syncthreads();
    if (row_ok) {   // rgb_fc.2, rgb_fc.4 and the mask fill (ibrnet.py:364-365)
      float h8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h8[j] = c.wbuf[lay::RGB1_B + j];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float a = at<LD>(tH, k, r);
        const float4 wa = *reinterpret_cast<const float4*>(c.wbuf + lay::RGB1_W + k * 8);
        const float4 wb = *reinterpret_cast<const float4*>(c.wbuf + lay::RGB1_W + k * 8 + 4);
        h8[0] = fmaf(wa.x, a, h8[0]); h8[1] = fmaf(wa.y, a, h8[1]); h8[2] = fmaf(wa.z, a, h8[2]); h8[3] = fmaf(wa.w, a, h8[3]);
        h8[4] = fmaf(wb.x, a, h8[4]); h8[5] = fmaf(wb.y, a, h8[5]); h8[6] = fmaf(wb.z, a, h8[6]); h8[7] = fmaf(wb.w, a, h8[7]);
      }
      float lg = c.wbuf[lay::RGB2_B];
#pragma unroll
      for (int j = 0; j < 8; ++j) lg = fmaf(c.wbuf[lay::RGB2_W + j], elu(h8[j]), lg);
      at<LD>(tS, S_LOGIT, r) = at<LD>(tS, S_MASK, r) == 0.f ? -1e9f : lg;
    }
    __syncthreads();

    // ---------------- phase 9: per-point softmax blend + second weighted mean/var (ibrnet.py:350-353, 366-367) ----------------
    if (c.tid < P) {
      const int q = c.tid;
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
      at<LDP>(tGOUT, 19, q) = parr[P_NV * LDP + q];
    }
    for (int it = c.tid; it < P * 33; it += NT) {
      const int f = it / P, q = it - f * P;
      if (f < 32) {
        float m = 0.f;
        for (int vv = 0; vv < rfn; ++vv) m = fmaf(at<LD>(tRF, f, vv * P + q), at<LD>(tS, S_W2, vv * P + q), m);
        float vr = 0.f;
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

    // ---------------- phase 10: geometry_fc per point (ibrnet.py:354) ----------------
    __syncthreads();
    stage(c, W + lay::GRP_D2, lay::GRP_D2_SIZE);
    __syncthreads();
    {
      Frag<64, 4, 2> f;
      f.setup(c);
      if (f.r0 < P) {
        f.init_bias(c.wbuf + lay::GEO0_B);
        f.mac<65, LDP>(tGVEC, 0, c.wbuf + lay::GEO0_W);
        f.store([&](int col, int r4, float4 v4) { at4<LDP>(tGHID, col, r4) = elu4(v4); });
      }
    }
    __syncthreads();
    {
      Frag<16, 4, 1> f;
      f.setup(c);
      if (f.r0 < P) {
        f.init_bias(c.wbuf + lay::GEO1_B);
        f.mac<64, LDP>(tGHID, 0, c.wbuf + lay::GEO1_W);
        f.store([&](int col, int r4, float4 v4) { at4<LDP>(tGOUT, col, r4) = elu4(v4); });
      }
    }
    __syncthreads();
    {
      const int cnt = min(P, N - n0) * REC;
      float* __restrict__ dst = pp.point_rec + size_t(n0) * REC;
      for (int i = c.tid; i < cnt; i += NT) {
        const int q = i / REC, cc = i - q * REC;
        dst[i] = at<LDP>(tGOUT, cc, q);
      }
    }
  }
}

}  // namespace pk

int launch_point_kernel_tc(const NrPassParams* p, float* dbg, cudaStream_t stream);   // nr_point_kernel_tc.cu

int launch_point_kernel(const NrPassParams* p, float* dbg, cudaStream_t stream) {
  NR_CHECK_ARG(p != nullptr, "params");
  NR_CHECK_ARG(p->coords && p->que_depth && p->que_cam && p->feat && p->rgb && p->view_params && p->w_point && p->point_rec,
               "null device pointer");
  NR_CHECK_ARG(p->rfn >= 1 && p->rfn <= NR_MAX_VIEWS, "rfn out of range");
  NR_CHECK_ARG(p->dn >= 3 && p->dn <= NR_MAX_SAMPLES, "dn out of range");
  NR_CHECK_ARG(p->rn >= 0, "rn");
  if (p->rn == 0) return NR_OK;
  NR_CHECK_ARG((long long)p->rn * p->dn < (1ll << 31) / NR_POINT_REC, "too many points for one call; chunk the rays");
  if (p->w_tc != nullptr) return launch_point_kernel_tc(p, dbg, stream);
  pk::KParams kp;
  kp.p = *p;
  kp.dbg = dbg;
  kp.P = (pk::LD / p->rfn) & ~3;
  if (kp.P > 84) kp.P = 84;
  const long long N = (long long)p->rn * p->dn;
  NR_CHECK_ARG(N < (1ll << 31) / NR_POINT_REC, "too many points for one call; chunk the rays");
  kp.n_tiles = int((N + kp.P - 1) / kp.P);
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = kp.n_tiles < sms ? kp.n_tiles : sms;
  static bool attr_done[2] = {false, false};
  if (dbg) {
    if (!attr_done[1]) {
      cudaFuncSetAttribute(pk::point_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pk::SMEM_BYTES));
      attr_done[1] = true;
    }
    pk::point_kernel<true><<<grid, pk::NT, pk::SMEM_BYTES, stream>>>(kp);
  } else {
    if (!attr_done[0]) {
      cudaFuncSetAttribute(pk::point_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pk::SMEM_BYTES));
      attr_done[0] = true;
    }
    pk::point_kernel<false><<<grid, pk::NT, pk::SMEM_BYTES, stream>>>(kp);
  }
  NR_CHECK_LAUNCH("point_kernel");
  return NR_OK;
}

}  // namespace nr
