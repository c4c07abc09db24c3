// The two encoder networks as layer graphs over the building blocks of nr_conv.cuh.  The graph is a template over an
// `Ops` backend: csrc/nr_encoder.cu instantiates it with kernel launches on a stream (the product), and
// tests/cpu_harness/conv_cpu_harness.cu with host loops over the same __host__ __device__ routines (CPU parity tests of the
// wiring, the weight layout and the index math against the reference's torch modules).
//
//   image_encoder = ResUNetLight(3, [1,2,6,4], 32, inplanes=16)   reference network/ops.py:150-230 (renderer.py:59)
//   vis_encoder   = DefaultVisEncoder                              reference network/vis_encoder.py:6-21
#pragma once
#include <stddef.h>

#include "nr_conv.cuh"

namespace nr {
namespace enc {

// ---- parameter tensors in state_dict() order, and where each lands in the packed buffer ----------------------------
struct TensorSpec {
  int kind;        // 0 conv weight [cout][cin][ks][ks] -> packed [tap][cin][cout]; 1 vector [n] copied
  int cout, cin, ks;
  int rot;         // packed input channel c reads reference channel (c + rot) % cin
  int cin_major;   // packed as [cin][tap][cout] (the 7x7 first layer)
  int cin_ref;     // input channels of the reference tensor (< cin: the packed tensor has zero rows for the padding channels)
  long long off;   // floats, multiple of 4
  long long n;     // elements
};
struct NetSpec {
  TensorSpec t[160];
  int count;
  long long total;
  int conv(int cout, int cin, int ks, int rot = 0, int cin_major = 0, int cin_ref = 0) {
    TensorSpec& s = t[count];
    s.kind = 0; s.cout = cout; s.cin = cin; s.ks = ks; s.rot = rot; s.cin_major = cin_major; s.cin_ref = cin_ref > 0 ? cin_ref : cin;
    s.off = total; s.n = (long long)cout * cin * ks * ks;
    total += (s.n + 3) & ~3LL;
    return count++;
  }
  int vec(int n) {
    TensorSpec& s = t[count];
    s.kind = 1; s.cout = n; s.cin = 1; s.ks = 1; s.rot = 0; s.cin_major = 0; s.cin_ref = 1;
    s.off = total; s.n = n;
    total += (s.n + 3) & ~3LL;
    return count++;
  }
};

// packed element e of tensor s <- index into the reference tensor (-1: a padding channel, packed as zero)
NR_HD long long pack_source(const TensorSpec& s, long long e) {
  if (s.kind == 1) return e;
  const int taps = s.ks * s.ks;
  int tap, c, o;
  o = int(e % s.cout);
  const long long r = e / s.cout;
  if (s.cin_major) { tap = int(r % taps); c = int(r / taps); }
  else { c = int(r % s.cin); tap = int(r / s.cin); }
  const int c_ref = (c + s.rot) % s.cin;
  if (c_ref >= s.cin_ref) return -1;
  return ((long long)o * s.cin_ref + c_ref) * taps + tap;
}

struct BasicBlockIdx { int c1, n1w, n1b, c2, n2w, n2b, ds, dsw, dsb; int cin, cout, stride; };
struct ConvBnIdx { int w, b, nw, nb; int cin, cout; };

// The U-shaped residual encoder both networks share (ops.py:150-230 ResUNetLight, :232-312 ResEncoder): first conv + IN + ReLU,
// three stages of BasicBlocks (the first block of a stage has stride 2 and a 1x1 downsample branch), two
// upsample-conv / skip / conv decoder steps, a 1x1 output conv.
struct UNet {
  NetSpec spec;
  int conv1, bn1w, bn1b, inplanes;
  BasicBlockIdx blocks[12];
  int nb[3];                   // blocks per stage
  ConvBnIdx upconv3, iconv3, upconv2, iconv2;
  int out_w, out_b, out_planes;
};
typedef UNet ImageNet;

inline BasicBlockIdx basic_block(NetSpec& s, int cin, int cout, int stride) {
  BasicBlockIdx b;
  b.cin = cin; b.cout = cout; b.stride = stride;
  b.c1 = s.conv(cout, cin, 3); b.n1w = s.vec(cout); b.n1b = s.vec(cout);
  b.c2 = s.conv(cout, cout, 3); b.n2w = s.vec(cout); b.n2b = s.vec(cout);
  if (stride != 1 || cin != cout) { b.ds = s.conv(cout, cin, 1); b.dsw = s.vec(cout); b.dsb = s.vec(cout); }
  else b.ds = b.dsw = b.dsb = -1;
  return b;
}
inline ConvBnIdx conv_bn(NetSpec& s, int cin, int cout) {
  ConvBnIdx c;
  c.cin = cin; c.cout = cout;
  c.w = s.conv(cout, cin, 3); c.b = s.vec(cout); c.nw = s.vec(cout); c.nb = s.vec(cout);
  return c;
}
// everything after the first conv + norm, in state_dict() order
inline void build_unet_tail(UNet& n, int inplanes, int nb0, int nb1, int nb2, int out_planes) {
  NetSpec& s = n.spec;
  const int planes[3] = {32, 64, 128};
  n.inplanes = inplanes; n.nb[0] = nb0; n.nb[1] = nb1; n.nb[2] = nb2; n.out_planes = out_planes;
  int k = 0, cin = inplanes;
  for (int st = 0; st < 3; ++st)
    for (int i = 0; i < n.nb[st]; ++i) {
      n.blocks[k++] = basic_block(s, cin, planes[st], i == 0 ? 2 : 1);
      cin = planes[st];
    }
  n.upconv3 = conv_bn(s, 128, 64);
  n.iconv3 = conv_bn(s, 128, 64);
  n.upconv2 = conv_bn(s, 64, 32);
  n.iconv2 = conv_bn(s, 64, out_planes);
  n.out_w = s.conv(out_planes, out_planes, 1); n.out_b = s.vec(out_planes);
}

// image_encoder = ResUNetLight(3, [1,2,6,4], 32, inplanes=16): layers [1,2,6] are used (the fourth entry never is, ops.py:166-170)
inline void build_image_net(ImageNet& n) {
  NetSpec& s = n.spec;
  s.count = 0; s.total = 0;
  n.conv1 = s.conv(16, 3, 7, 0, 1); n.bn1w = s.vec(16); n.bn1b = s.vec(16);
  build_unet_tail(n, 16, 1, 2, 6, 32);
}

// DepthInitNet (init_net.py:78-101): res_net = ResEncoder (ops.py:232-312: conv1 8x8 stride 2 pad 2 on 12 channels = imgs 3 |
// normalised depth 1 | diff_feats 8, packed with 4 zero channels; stages [2,2,2]), depth_skip, conv_out (1x1, 48 -> 32)
struct DepthInitNet {
  UNet res;                    // its spec holds ALL tensors of the module in state_dict() order
  int skip_w0, skip_b0, skip_w1, skip_b1, out_w, out_b;
};
inline void build_depth_init_net(DepthInitNet& n) {
  NetSpec& s = n.res.spec;
  s.count = 0; s.total = 0;
  n.res.conv1 = s.conv(32, 16, 8, 0, 0, 12); n.res.bn1w = s.vec(32); n.res.bn1b = s.vec(32);
  build_unet_tail(n.res, 32, 2, 2, 2, 32);
  n.skip_w0 = s.vec(8 * 1 * 2 * 2); n.skip_b0 = s.vec(8);          // depth_skip.0 (copied as is: the SIMT routine reads PyTorch layout)
  n.skip_w1 = s.vec(16 * 8 * 2 * 2); n.skip_b1 = s.vec(16);        // depth_skip.2
  n.out_w = s.conv(32, 48, 1); n.out_b = s.vec(32);                // conv_out on cat([depth_feats 16, feats 32])
}

struct ResidualIdx { int n0w, n0b, c0, n1w, n1b, c1; };
struct VisNet {
  NetSpec spec;
  int conv0;                  // conv3x3(64, 32): the packed input order is [ray_feats | img_feats] (the frame pack's), the reference
                              // concatenates [img_feats, ray_feats] (vis_encoder.py:20) -> rot 32
  ResidualIdx rb[2];
  int conv_out;
};
inline void build_vis_net(VisNet& n) {
  NetSpec& s = n.spec;
  s.count = 0; s.total = 0;
  n.conv0 = s.conv(32, 64, 3, 32);
  for (int i = 0; i < 2; ++i) {
    ResidualIdx& r = n.rb[i];
    r.n0w = s.vec(32); r.n0b = s.vec(32); r.c0 = s.conv(32, 32, 3);
    r.n1w = s.vec(32); r.n1b = s.vec(32); r.c1 = s.conv(32, 32, 3);
  }
  n.conv_out = s.conv(32, 32, 1);
}

// conv3x3(cin, 32) -> ResidualBlock(32, 32) x nrb -> conv1x1(32, 32), all reflect / bias-free (the three heads of
// CostVolumeInitNet, init_net.py:230-249; the vis encoder is the same stack with two blocks)
struct StackNet { int conv0; ResidualIdx rb[2]; int nrb; int conv_out; int cin; };
inline StackNet build_stack(NetSpec& s, int cin, int cin_ref, int nrb) {
  StackNet n;
  n.cin = cin; n.nrb = nrb;
  n.conv0 = s.conv(32, cin, 3, 0, 0, cin_ref);
  for (int i = 0; i < nrb; ++i) {
    ResidualIdx& r = n.rb[i];
    r.n0w = s.vec(32); r.n0b = s.vec(32); r.c0 = s.conv(32, 32, 3);
    r.n1w = s.vec(32); r.n1b = s.vec(32); r.c1 = s.conv(32, 32, 3);
  }
  n.conv_out = s.conv(32, 32, 1);
  return n;
}
// CostVolumeInitNet without its frozen MVSNet (init_net.py:227-254): res_net = ResUNetLight(out_dim=32) (3 -> 32, stages
// [2,3,6]), volume_conv2d on the softmaxed cost volume (64 planes), depth_conv on the normalised regressed depth (1 channel,
// packed into 16), out_conv on cat([ref_feats, volume_feats, depth_feats]); one spec, state_dict() order
struct CostVolumeHead {
  UNet res;
  StackNet volume, depthc, outc;
};
inline void build_cost_volume_head(CostVolumeHead& n, int cost_volume_sn) {
  NetSpec& s = n.res.spec;
  s.count = 0; s.total = 0;
  n.res.conv1 = s.conv(32, 3, 7, 0, 1); n.res.bn1w = s.vec(32); n.res.bn1b = s.vec(32);
  build_unet_tail(n.res, 32, 2, 3, 6, 32);
  n.volume = build_stack(s, cost_volume_sn, cost_volume_sn, 1);
  n.depthc = build_stack(s, 16, 1, 1);
  n.outc = build_stack(s, 96, 96, 1);
}

// ---- geometry ---------------------------------------------------------------------------------------------------------
inline int conv_out(int n, int ks, int stride, int pad = -1) { return (n + 2 * (pad < 0 ? (ks - 1) / 2 : pad) - ks) / stride + 1; }
struct ImageDims {
  int h0, w0, h1, w1, h2, w2, h3, w3;   // after conv1 (/2), layer1 (/4), layer2 (/8), layer3 (/16)
  int u3h, u3w, u2h, u2w;               // after the two x2 upsamplings; (u2h, u2w) is the output size
};
inline ImageDims unet_dims(int H, int W, int ks, int pad) {        // first conv: ks x ks, stride 2, padding pad
  ImageDims d;
  d.h0 = conv_out(H, ks, 2, pad); d.w0 = conv_out(W, ks, 2, pad);
  d.h1 = conv_out(d.h0, 3, 2); d.w1 = conv_out(d.w0, 3, 2);
  d.h2 = conv_out(d.h1, 3, 2); d.w2 = conv_out(d.w1, 3, 2);
  d.h3 = conv_out(d.h2, 3, 2); d.w3 = conv_out(d.w2, 3, 2);
  d.u3h = 2 * d.h3; d.u3w = 2 * d.w3;
  d.u2h = 2 * d.u3h; d.u2w = 2 * d.u3w;
  return d;
}
inline bool dims_ok(const ImageDims& d) {      // the skip connections pad the encoder feature up to the decoder's size, never crop (ops.py:199-208)
  return d.h3 >= 2 && d.w3 >= 2 && d.u3h >= d.h2 && d.u3w >= d.w2 && d.u2h >= d.h1 && d.u2w >= d.w1;
}
inline ImageDims image_dims(int H, int W) { return unet_dims(H, W, 7, 3); }

struct CopyP {     // y[n, yo, xo, y_off + c] = (yo - py, xo - px) inside the source ? x[n, yo - py, xo - px, x_off + c] : 0
  const float* x; float* y;
  int N, H, W, Ho, Wo, C, py, px, x_stride, x_off, y_stride, y_off;
};

// bump allocator over the caller's workspace (256-byte granules)
struct Arena {
  char* base; size_t size, used;
  bool ok;
  size_t peak = 0;     // high-water mark (graphs that release intermediates reset `used`)
  void* take(size_t bytes) {
    const size_t a = (used + 255) & ~size_t(255);
    if (base != nullptr && a + bytes > size) { ok = false; return base; }
    used = a + bytes;
    if (used > peak) peak = used;
    return base != nullptr ? base + a : nullptr;
  }
  float* floats(long long n) { return (float*)take(size_t(n) * sizeof(float)); }
  double* doubles(long long n) { return (double*)take(size_t(n) * sizeof(double)); }
};

template <class Ops>
struct Builder {
  Ops& ops;
  Arena& ar;
  const float* W;      // packed parameters
  int N;
  double* stats_base; long long stats_used, stats_cap;

  const float* w(const NetSpec& s, int i) const { return i < 0 ? nullptr : W + s.t[i].off; }
  double* stats(int C) {
    double* p = stats_base != nullptr ? stats_base + stats_used : nullptr;
    stats_used += (long long)N * C * 2;
    if (stats_base != nullptr && stats_used > stats_cap) ar.ok = false;
    return p;
  }
  // y (dense [N,Ho,Wo,cout] unless y/y_stride/y_off say otherwise) = conv(x)
  float* conv(const float* x, int x_stride, int x_off, int H, int Wd, int cin, int cout, int ks, int stride, const float* wt, const float* bias,
              const float* res, int res_stride, int res_off, double* st, float* y, int y_stride, int y_off, int& Ho, int& Wo, int pad = -1) {
    if (pad < 0) pad = (ks - 1) / 2;
    Ho = conv_out(H, ks, stride, pad); Wo = conv_out(Wd, ks, stride, pad);
    if (y == nullptr) { y = ar.floats((long long)N * Ho * Wo * cout); y_stride = cout; y_off = 0; }
    cv::ConvP p;
    p.x = x; p.w = wt; p.bias = bias; p.res = res; p.y = y; p.stats = st;
    p.N = N; p.H = H; p.W = Wd; p.Ho = Ho; p.Wo = Wo; p.Cin = cin; p.Cout = cout; p.ks = ks; p.stride = stride; p.pad = pad; p.reflect = 1;
    p.x_stride = x_stride; p.x_off = x_off; p.y_stride = y_stride; p.y_off = y_off; p.res_stride = res_stride; p.res_off = res_off;
    p.tf32x1 = 0; p.bm = 0;      // the backend decides (StreamOps)
    ops.conv(p);
    return y;
  }
  // y = act(IN(x) gamma + beta [+ res | + IN(res) rg + rb])
  float* norm(const float* x, int C, int HW, const double* st, const float* g, const float* b, const float* res, int res_stride, int res_off,
              const double* rst, const float* rg, const float* rb, int act, float* y, int y_stride, int y_off) {
    if (y == nullptr) { y = ar.floats((long long)N * HW * C); y_stride = C; y_off = 0; }
    cv::NormP p;
    p.x = x; p.stats = st; p.gamma = g; p.beta = b; p.res = res; p.res_stats = rst; p.res_gamma = rg; p.res_beta = rb; p.y = y;
    p.N = N; p.HW = HW; p.C = C; p.act = act; p.x_stride = C; p.x_off = 0; p.res_stride = res_stride; p.res_off = res_off;
    p.y_stride = y_stride; p.y_off = y_off; p.eps = 1e-5f;
    ops.norm(p);
    return y;
  }
};

// BasicBlock (ops.py:86-124): relu(IN(conv2(relu(IN(conv1(x))))) + identity), identity = x or IN(conv1x1(x))
template <class Ops>
float* run_basic_block(Builder<Ops>& b, const NetSpec& s, const BasicBlockIdx& k, const float* x, int& H, int& Wd) {
  int Ho, Wo, h2, w2;
  double* st1 = b.stats(k.cout);
  float* t1 = b.conv(x, k.cin, 0, H, Wd, k.cin, k.cout, 3, k.stride, b.w(s, k.c1), nullptr, nullptr, 0, 0, st1, nullptr, 0, 0, Ho, Wo);
  float* a1 = b.norm(t1, k.cout, Ho * Wo, st1, b.w(s, k.n1w), b.w(s, k.n1b), nullptr, 0, 0, nullptr, nullptr, nullptr, 1, t1, k.cout, 0);
  double* st2 = b.stats(k.cout);
  float* t2 = b.conv(a1, k.cout, 0, Ho, Wo, k.cout, k.cout, 3, 1, b.w(s, k.c2), nullptr, nullptr, 0, 0, st2, nullptr, 0, 0, h2, w2);
  float* out;
  if (k.ds >= 0) {
    double* st3 = b.stats(k.cout);
    int h3, w3;
    float* d = b.conv(x, k.cin, 0, H, Wd, k.cin, k.cout, 1, k.stride, b.w(s, k.ds), nullptr, nullptr, 0, 0, st3, nullptr, 0, 0, h3, w3);
    out = b.norm(t2, k.cout, Ho * Wo, st2, b.w(s, k.n2w), b.w(s, k.n2b), d, k.cout, 0, st3, b.w(s, k.dsw), b.w(s, k.dsb), 1, t2, k.cout, 0);
  } else {
    out = b.norm(t2, k.cout, Ho * Wo, st2, b.w(s, k.n2w), b.w(s, k.n2b), x, k.cin, 0, nullptr, nullptr, nullptr, 1, t2, k.cout, 0);
  }
  H = Ho; Wd = Wo;
  return out;
}

// `conv` of ops.py:126-138: ELU(IN(conv3x3 reflect + bias)); the result goes to channels [y_off, y_off + cout) of y
template <class Ops>
void run_conv_bn_elu(Builder<Ops>& b, const NetSpec& s, const ConvBnIdx& k, const float* x, int H, int Wd, float* y, int y_stride, int y_off) {
  int Ho, Wo;
  double* st = b.stats(k.cout);
  float* t = b.conv(x, k.cin, 0, H, Wd, k.cin, k.cout, 3, 1, b.w(s, k.w), b.w(s, k.b), nullptr, 0, 0, st, nullptr, 0, 0, Ho, Wo);
  b.norm(t, k.cout, Ho * Wo, st, b.w(s, k.nw), b.w(s, k.nb), nullptr, 0, 0, nullptr, nullptr, nullptr, 2, y, y_stride, y_off);
}

// Everything of the U-shaped encoder after conv1 + bn1 + relu (ops.py:213-228): x [N,h0,w0,inplanes] -> out slot
template <class Ops>
void unet_body(Builder<Ops>& b, Arena& ar, const UNet& net, const ImageDims& d, float* x, float* out, int out_stride, int out_off) {
  Ops& ops = b.ops;
  const NetSpec& s = net.spec;
  const int N = b.N;
  int h = d.h0, w = d.w0, k = 0;
  float* x1 = x;
  for (int i = 0; i < net.nb[0]; ++i) x1 = run_basic_block(b, s, net.blocks[k++], x1, h, w);     // layer1: [N,h1,w1,32]
  const int h1 = h, w1 = w;
  float* x2 = x1;
  for (int i = 0; i < net.nb[1]; ++i) x2 = run_basic_block(b, s, net.blocks[k++], x2, h, w);     // layer2: [N,h2,w2,64]
  const int h2 = h, w2 = w;
  float* x3 = x2;
  for (int i = 0; i < net.nb[2]; ++i) x3 = run_basic_block(b, s, net.blocks[k++], x3, h, w);     // layer3: [N,h3,w3,128]
  // upconv3 -> skipconnect(x2, .) -> iconv3
  float* up3 = ar.floats((long long)N * d.u3h * d.u3w * 128);
  cv::UpP u;
  u.x = x3; u.y = up3; u.N = N; u.H = h; u.W = w; u.Ho = d.u3h; u.Wo = d.u3w; u.C = 128; u.x_stride = 128; u.x_off = 0; u.y_stride = 128; u.y_off = 0;
  ops.upsample(u);
  float* cat3 = ar.floats((long long)N * d.u3h * d.u3w * 128);               // [upconv3 out (64) | x2 (64)]  (torch.cat([x2_up, x1_skip]), ops.py:207)
  run_conv_bn_elu(b, s, net.upconv3, up3, d.u3h, d.u3w, cat3, 128, 0);
  CopyP cp;
  cp.x = x2; cp.y = cat3; cp.N = N; cp.H = h2; cp.W = w2; cp.Ho = d.u3h; cp.Wo = d.u3w; cp.C = 64;
  cp.py = (d.u3h - h2) / 2; cp.px = (d.u3w - w2) / 2; cp.x_stride = 64; cp.x_off = 0; cp.y_stride = 128; cp.y_off = 64;
  ops.copy_pad(cp);
  float* i3 = ar.floats((long long)N * d.u3h * d.u3w * 64);
  run_conv_bn_elu(b, s, net.iconv3, cat3, d.u3h, d.u3w, i3, 64, 0);
  // upconv2 -> skipconnect(x1, .) -> iconv2
  float* up2 = ar.floats((long long)N * d.u2h * d.u2w * 64);
  u.x = i3; u.y = up2; u.H = d.u3h; u.W = d.u3w; u.Ho = d.u2h; u.Wo = d.u2w; u.C = 64; u.x_stride = 64; u.y_stride = 64;
  ops.upsample(u);
  float* cat2 = ar.floats((long long)N * d.u2h * d.u2w * 64);
  run_conv_bn_elu(b, s, net.upconv2, up2, d.u2h, d.u2w, cat2, 64, 0);
  cp.x = x1; cp.y = cat2; cp.H = h1; cp.W = w1; cp.Ho = d.u2h; cp.Wo = d.u2w; cp.C = 32;
  cp.py = (d.u2h - h1) / 2; cp.px = (d.u2w - w1) / 2; cp.x_stride = 32; cp.y_stride = 64; cp.y_off = 32;
  ops.copy_pad(cp);
  float* i2 = ar.floats((long long)N * d.u2h * d.u2w * net.out_planes);
  run_conv_bn_elu(b, s, net.iconv2, cat2, d.u2h, d.u2w, i2, net.out_planes, 0);
  // out_conv (1x1 + bias) straight into the caller's channel-last destination
  int Ho, Wo;
  b.conv(i2, net.out_planes, 0, d.u2h, d.u2w, net.out_planes, net.out_planes, 1, 1, b.w(s, net.out_w), b.w(s, net.out_b), nullptr, 0, 0, nullptr, out,
         out_stride, out_off, Ho, Wo);
}

// ResUNetLight.forward (ops.py:210-228).  imgs [N,3,H,W] (NCHW, as the reference holds them) -> out[n, y, x, out_off + c],
// c < 32, at image_dims(H, W).u2h x u2w (= H/4 x W/4 for sizes that are multiples of 16).
template <class Ops>
bool image_encoder_graph(Ops& ops, Arena& ar, const ImageNet& net, const float* packed, const float* imgs, int N, int H, int Wd, float* out,
                         int out_stride, int out_off, long long stats_cap, long long* stats_used) {
  const NetSpec& s = net.spec;
  const ImageDims d = image_dims(H, Wd);
  if (!dims_ok(d)) return false;
  Builder<Ops> b{ops, ar, packed, N, nullptr, 0, stats_cap};
  // InstanceNorm sums of every normalised conv output in one block, zeroed once (stats_cap == 0: a dry run that counts)
  if (stats_cap > 0) {
    b.stats_base = ar.doubles(stats_cap);
    ops.zero(b.stats_base, size_t(stats_cap) * sizeof(double));
  }
  // conv1 + bn1 + relu
  const int c0 = net.inplanes;      // 16 (image_encoder) or 32 (CostVolumeInitNet.res_net)
  double* st0 = b.stats(c0);
  float* c1 = ar.floats((long long)N * d.h0 * d.w0 * c0);
  cv::Conv7P p7;
  p7.img = imgs; p7.w = b.w(s, net.conv1); p7.y = c1; p7.stats = st0; p7.N = N; p7.H = H; p7.W = Wd; p7.Ho = d.h0; p7.Wo = d.w0; p7.Cout = net.inplanes;
  ops.conv7(p7);
  float* x = b.norm(c1, c0, d.h0 * d.w0, st0, b.w(s, net.bn1w), b.w(s, net.bn1b), nullptr, 0, 0, nullptr, nullptr, nullptr, 1, c1, c0, 0);
  unet_body(b, ar, net, d, x, out, out_stride, out_off);
  if (stats_used != nullptr) *stats_used = b.stats_used;
  return ar.ok;
}

// DepthInitNet.forward after extract_depth_for_init and get_diff_feats (init_net.py:93-101).  x16 [N,H,W,16] = imgs (3) |
// normalised depth (1) | diff_feats (8) | zeros (4), channel-last (the caller assembles it); depth [N,H,W] = the
// normalised depth again for depth_skip.  out[n, y, x, out_off + c], c < 32, at unet_dims(H, W, 8, 2).u2h x u2w.
inline ImageDims depth_init_dims(int H, int W) { return unet_dims(H, W, 8, 2); }
template <class Ops>
bool depth_init_graph(Ops& ops, Arena& ar, const DepthInitNet& net, const float* packed, const float* x16, const float* depth, int N, int H, int Wd,
                      float* out, int out_stride, int out_off, long long stats_cap, long long* stats_used) {
  const NetSpec& s = net.res.spec;
  const ImageDims d = depth_init_dims(H, Wd);
  // depth_skip halves twice without padding; torch.cat (init_net.py:101) needs the two maps to agree
  if (!dims_ok(d) || (H / 2) / 2 != d.u2h || (Wd / 2) / 2 != d.u2w) return false;
  Builder<Ops> b{ops, ar, packed, N, nullptr, 0, stats_cap};
  if (stats_cap > 0) {
    b.stats_base = ar.doubles(stats_cap);
    ops.zero(b.stats_base, size_t(stats_cap) * sizeof(double));
  }
  int Ho, Wo;
  double* st0 = b.stats(32);
  float* c1 = b.conv(x16, 16, 0, H, Wd, 16, 32, 8, 2, b.w(s, net.res.conv1), nullptr, nullptr, 0, 0, st0, nullptr, 0, 0, Ho, Wo, 2);
  float* x = b.norm(c1, 32, Ho * Wo, st0, b.w(s, net.res.bn1w), b.w(s, net.res.bn1b), nullptr, 0, 0, nullptr, nullptr, nullptr, 1, c1, 32, 0);
  float* cat = ar.floats((long long)N * d.u2h * d.u2w * 48);                 // [depth_feats 16 | feats 32]
  unet_body(b, ar, net.res, d, x, cat, 48, 16);
  cv::DepthSkipP ds;
  ds.depth = depth; ds.w0 = b.w(s, net.skip_w0); ds.b0 = b.w(s, net.skip_b0); ds.w1 = b.w(s, net.skip_w1); ds.b1 = b.w(s, net.skip_b1);
  ds.y = cat; ds.N = N; ds.H = H; ds.W = Wd; ds.Ho = d.u2h; ds.Wo = d.u2w; ds.y_stride = 48; ds.y_off = 0;
  ops.depth_skip(ds);
  b.conv(cat, 48, 0, d.u2h, d.u2w, 48, 32, 1, 1, b.w(s, net.out_w), b.w(s, net.out_b), nullptr, 0, 0, nullptr, out, out_stride, out_off, Ho, Wo);
  if (stats_used != nullptr) *stats_used = b.stats_used;
  return ar.ok;
}

// DefaultVisEncoder.forward (vis_encoder.py:19-21) on the channel-last frame pack: feat [N,fh,fw,64] holds the init-net
// ray_feats in channels 0..31 and the image encoder's img_feats in 32..63; the result overwrites channels 0..31.
template <class Ops>
bool vis_encoder_graph(Ops& ops, Arena& ar, const VisNet& net, const float* packed, float* feat, int N, int fh, int fw, long long stats_cap,
                       long long* stats_used) {
  const NetSpec& s = net.spec;
  Builder<Ops> b{ops, ar, packed, N, nullptr, 0, stats_cap};
  if (stats_cap > 0) {
    b.stats_base = ar.doubles(stats_cap);
    ops.zero(b.stats_base, size_t(stats_cap) * sizeof(double));
  }
  int Ho, Wo;
  const int HW = fh * fw;
  double* st = b.stats(32);
  float* x = b.conv(feat, 64, 0, fh, fw, 64, 32, 3, 1, b.w(s, net.conv0), nullptr, nullptr, 0, 0, st, nullptr, 0, 0, Ho, Wo);
  for (int i = 0; i < 2; ++i) {       // ResidualBlock (ops.py:43-76): x + conv(relu(IN(conv(relu(IN(x))))))
    const ResidualIdx& r = net.rb[i];
    float* a0 = b.norm(x, 32, HW, st, b.w(s, r.n0w), b.w(s, r.n0b), nullptr, 0, 0, nullptr, nullptr, nullptr, 1, nullptr, 0, 0);
    double* st1 = b.stats(32);
    float* t = b.conv(a0, 32, 0, fh, fw, 32, 32, 3, 1, b.w(s, r.c0), nullptr, nullptr, 0, 0, st1, nullptr, 0, 0, Ho, Wo);
    b.norm(t, 32, HW, st1, b.w(s, r.n1w), b.w(s, r.n1b), nullptr, 0, 0, nullptr, nullptr, nullptr, 1, t, 32, 0);
    st = i == 0 ? b.stats(32) : nullptr;   // the second block's output is not normalised again
    x = b.conv(t, 32, 0, fh, fw, 32, 32, 3, 1, b.w(s, r.c1), nullptr, x, 32, 0, st, a0, 32, 0, Ho, Wo);   // a0 is free by now
  }
  b.conv(x, 32, 0, fh, fw, 32, 32, 1, 1, b.w(s, net.conv_out), nullptr, nullptr, 0, 0, nullptr, feat, 64, 0, Ho, Wo);
  if (stats_used != nullptr) *stats_used = b.stats_used;
  return ar.ok;
}

// x [N,H,W,cin] (stride / offset) -> conv0 -> residual blocks -> conv_out into the output slot
template <class Ops>
void run_stack(Builder<Ops>& b, const NetSpec& s, const StackNet& net, const float* xin, int x_stride, int x_off, int H, int Wd, float* y, int y_stride,
               int y_off) {
  int Ho, Wo;
  const int HW = H * Wd;
  double* st = b.stats(32);
  float* x = b.conv(xin, x_stride, x_off, H, Wd, net.cin, 32, 3, 1, b.w(s, net.conv0), nullptr, nullptr, 0, 0, st, nullptr, 0, 0, Ho, Wo);
  for (int i = 0; i < net.nrb; ++i) {       // ResidualBlock (ops.py:43-76)
    const ResidualIdx& r = net.rb[i];
    float* a0 = b.norm(x, 32, HW, st, b.w(s, r.n0w), b.w(s, r.n0b), nullptr, 0, 0, nullptr, nullptr, nullptr, 1, nullptr, 0, 0);
    double* st1 = b.stats(32);
    float* t = b.conv(a0, 32, 0, H, Wd, 32, 32, 3, 1, b.w(s, r.c0), nullptr, nullptr, 0, 0, st1, nullptr, 0, 0, Ho, Wo);
    b.norm(t, 32, HW, st1, b.w(s, r.n1w), b.w(s, r.n1b), nullptr, 0, 0, nullptr, nullptr, nullptr, 1, t, 32, 0);
    st = i + 1 < net.nrb ? b.stats(32) : nullptr;
    x = b.conv(t, 32, 0, H, Wd, 32, 32, 3, 1, b.w(s, r.c1), nullptr, x, 32, 0, st, a0, 32, 0, Ho, Wo);
  }
  b.conv(x, 32, 0, H, Wd, 32, 32, 1, 1, b.w(s, net.conv_out), nullptr, nullptr, 0, 0, nullptr, y, y_stride, y_off, Ho, Wo);
}

// CostVolumeInitNet.forward after construct_cost_volume_with_src (init_net.py:247-254): imgs [N,3,H,W] (NCHW), prob
// [N,fh,fw,sn] (softmaxed cost volume, channel-last), depth16 [N,fh,fw,16] (normalised regressed depth in channel 0, zeros
// elsewhere) -> out[n, y, x, out_off + c], c < 32
template <class Ops>
bool cost_volume_head_graph(Ops& ops, Arena& ar, const CostVolumeHead& net, const float* packed, const float* imgs, const float* prob, int sn,
                            const float* depth16, int N, int H, int Wd, int fh, int fw, float* out, int out_stride, int out_off, long long stats_cap,
                            long long* stats_used) {
  const NetSpec& s = net.res.spec;
  const ImageDims d = image_dims(H, Wd);
  if (!dims_ok(d) || d.u2h != fh || d.u2w != fw || sn != net.volume.cin) return false;
  Builder<Ops> b{ops, ar, packed, N, nullptr, 0, stats_cap};
  if (stats_cap > 0) {
    b.stats_base = ar.doubles(stats_cap);
    ops.zero(b.stats_base, size_t(stats_cap) * sizeof(double));
  }
  float* cat = ar.floats((long long)N * fh * fw * 96);           // [ref_feats 32 | volume_feats 32 | depth_feats 32]
  double* st0 = b.stats(32);
  float* c1 = ar.floats((long long)N * d.h0 * d.w0 * 32);
  cv::Conv7P p7;
  p7.img = imgs; p7.w = b.w(s, net.res.conv1); p7.y = c1; p7.stats = st0; p7.N = N; p7.H = H; p7.W = Wd; p7.Ho = d.h0; p7.Wo = d.w0; p7.Cout = 32;
  ops.conv7(p7);
  float* x = b.norm(c1, 32, d.h0 * d.w0, st0, b.w(s, net.res.bn1w), b.w(s, net.res.bn1b), nullptr, 0, 0, nullptr, nullptr, nullptr, 1, c1, 32, 0);
  unet_body(b, ar, net.res, d, x, cat, 96, 0);
  run_stack(b, s, net.volume, prob, sn, 0, fh, fw, cat, 96, 32);
  run_stack(b, s, net.depthc, depth16, 16, 0, fh, fw, cat, 96, 64);
  run_stack(b, s, net.outc, cat, 96, 0, fh, fw, out, out_stride, out_off);
  if (stats_used != nullptr) *stats_used = b.stats_used;
  return ar.ok;
}

// dry runs: the arena hands out fake (never dereferenced) addresses so that in-place reuse is counted once
static char* const DRY_BASE = (char*)0x100000;
struct NullOps {
  void conv(const cv::ConvP&) {}
  void conv7(const cv::Conv7P&) {}
  void norm(const cv::NormP&) {}
  void upsample(const cv::UpP&) {}
  void copy_pad(const CopyP&) {}
  void depth_skip(const cv::DepthSkipP&) {}
  void zero(void*, size_t) {}
};
// doubles of InstanceNorm sums a forward needs (dry run of the graph)
inline long long image_stats_doubles(const ImageNet& net, int N, int H, int W) {
  NullOps ops;
  Arena ar{DRY_BASE, ~size_t(0) / 2, 0, true};
  long long used = 0;
  image_encoder_graph(ops, ar, net, nullptr, nullptr, N, H, W, (float*)DRY_BASE, 32, 0, 0, &used);
  return used;
}
inline long long vis_stats_doubles(const VisNet& net, int N, int fh, int fw) {
  NullOps ops;
  Arena ar{DRY_BASE, ~size_t(0) / 2, 0, true};
  long long used = 0;
  vis_encoder_graph(ops, ar, net, nullptr, (float*)DRY_BASE, N, fh, fw, 0, &used);
  return used;
}
inline long long depth_init_stats_doubles(const DepthInitNet& net, int N, int H, int W) {
  NullOps ops;
  Arena ar{DRY_BASE, ~size_t(0) / 2, 0, true};
  long long used = 0;
  depth_init_graph(ops, ar, net, nullptr, nullptr, nullptr, N, H, W, (float*)DRY_BASE, 32, 0, 0, &used);
  return used;
}
inline size_t depth_init_workspace_bytes(const DepthInitNet& net, int N, int H, int W) {
  NullOps ops;
  Arena ar{DRY_BASE, ~size_t(0) / 2, 0, true};
  depth_init_graph(ops, ar, net, nullptr, nullptr, nullptr, N, H, W, (float*)DRY_BASE, 32, 0, depth_init_stats_doubles(net, N, H, W), nullptr);
  return ar.used + 256;
}
inline long long cv_head_stats_doubles(const CostVolumeHead& net, int N, int H, int W, int fh, int fw) {
  NullOps ops;
  Arena ar{DRY_BASE, ~size_t(0) / 2, 0, true};
  long long used = 0;
  cost_volume_head_graph(ops, ar, net, nullptr, nullptr, nullptr, net.volume.cin, nullptr, N, H, W, fh, fw, (float*)DRY_BASE, 32, 0, 0, &used);
  return used;
}
inline size_t cv_head_workspace_bytes(const CostVolumeHead& net, int N, int H, int W, int fh, int fw) {
  NullOps ops;
  Arena ar{DRY_BASE, ~size_t(0) / 2, 0, true};
  cost_volume_head_graph(ops, ar, net, nullptr, nullptr, nullptr, net.volume.cin, nullptr, N, H, W, fh, fw, (float*)DRY_BASE, 32, 0,
                         cv_head_stats_doubles(net, N, H, W, fh, fw), nullptr);
  return ar.used + 256;
}
inline size_t image_workspace_bytes(const ImageNet& net, int N, int H, int W) {
  NullOps ops;
  Arena ar{DRY_BASE, ~size_t(0) / 2, 0, true};
  image_encoder_graph(ops, ar, net, nullptr, nullptr, N, H, W, (float*)DRY_BASE, 32, 0, image_stats_doubles(net, N, H, W), nullptr);
  return ar.used + 256;
}
inline size_t vis_workspace_bytes(const VisNet& net, int N, int fh, int fw) {
  NullOps ops;
  Arena ar{DRY_BASE, ~size_t(0) / 2, 0, true};
  vis_encoder_graph(ops, ar, net, nullptr, (float*)DRY_BASE, N, fh, fw, vis_stats_doubles(net, N, fh, fw), nullptr);
  return ar.used + 256;
}

}  // namespace enc
}  // namespace nr
