// MVSNet (frozen) as a layer graph over the routines of nr_mvs.cuh, templated over an `Ops` backend like the encoder graphs:
// csrc/nr_mvs.cu launches kernels, tests/cpu_harness/mvs_cpu_harness.cu runs host loops.
//   reference network/mvsnet/mvsnet.py:7-66 (FeatureNet, CostRegNet), :124-141 (construct_cost_volume_with_src),
//   network/init_net.py:113-160 (resize rule, projection matrices, depth planes, nan -> 0, softmax, depth regression)
#pragma once
#include "nr_encoder_graph.cuh"
#include "nr_mvs.cuh"

namespace nr {
namespace mvs {

// ---- parameters: the module's state_dict() order (conv weight, then the ABN's weight, bias, running_mean, running_var) ------
struct LayerSpec {
  int cin, cout, k, stride, dims;   // dims 2 or 3
  int transposed, has_bn, has_bias;
  long long w_off, scale_off, shift_off;   // packed: w [tap][cin][cout], scale [cout], shift [cout]
};
struct MvsNet {
  LayerSpec L[20];
  int n_layers, n_tensors;
  long long total;
};
inline void add_layer(MvsNet& n, int cin, int cout, int k, int stride, int dims, int transposed, int has_bn, int has_bias) {
  LayerSpec& l = n.L[n.n_layers++];
  l.cin = cin; l.cout = cout; l.k = k; l.stride = stride; l.dims = dims; l.transposed = transposed; l.has_bn = has_bn; l.has_bias = has_bias;
  const long long taps = dims == 3 ? (long long)k * k * k : (long long)k * k;
  l.w_off = n.total; n.total += (taps * cin * cout + 3) & ~3LL;
  l.scale_off = n.total; n.total += (cout + 3) & ~3LL;
  l.shift_off = n.total; n.total += (cout + 3) & ~3LL;
  n.n_tensors += 1 + (has_bn ? 4 : 0) + (has_bias ? 1 : 0);
}
// layer indices
enum { F0 = 0, F1, F2, F3, F4, F5, F6, FEAT, C0, C1, C2, C3, C4, C5, C6, C7, C9, C11, PROB, N_LAYERS };
inline void build_mvsnet(MvsNet& n) {
  n.n_layers = 0; n.n_tensors = 0; n.total = 0;
  add_layer(n, 3, 8, 3, 1, 2, 0, 1, 0); add_layer(n, 8, 8, 3, 1, 2, 0, 1, 0);                                      // feature.conv0, conv1
  add_layer(n, 8, 16, 5, 2, 2, 0, 1, 0); add_layer(n, 16, 16, 3, 1, 2, 0, 1, 0); add_layer(n, 16, 16, 3, 1, 2, 0, 1, 0);   // conv2..4
  add_layer(n, 16, 32, 5, 2, 2, 0, 1, 0); add_layer(n, 32, 32, 3, 1, 2, 0, 1, 0);                                  // conv5, conv6
  add_layer(n, 32, 32, 3, 1, 2, 0, 0, 1);                                                                          // feature.feature (bias)
  add_layer(n, 32, 8, 3, 1, 3, 0, 1, 0);                                                                           // cost_regularization.conv0
  add_layer(n, 8, 16, 3, 2, 3, 0, 1, 0); add_layer(n, 16, 16, 3, 1, 3, 0, 1, 0);                                   // conv1, conv2
  add_layer(n, 16, 32, 3, 2, 3, 0, 1, 0); add_layer(n, 32, 32, 3, 1, 3, 0, 1, 0);                                  // conv3, conv4
  add_layer(n, 32, 64, 3, 2, 3, 0, 1, 0); add_layer(n, 64, 64, 3, 1, 3, 0, 1, 0);                                  // conv5, conv6
  add_layer(n, 64, 32, 3, 2, 3, 1, 1, 0); add_layer(n, 32, 16, 3, 2, 3, 1, 1, 0); add_layer(n, 16, 8, 3, 2, 3, 1, 1, 0);   // conv7, conv9, conv11
  add_layer(n, 8, 1, 3, 1, 3, 0, 0, 1);                                                                            // prob (bias)
}

// Element e of a layer's packed weight <- index into the reference tensor: Conv [cout][cin][taps], ConvTranspose3d [cin][cout][taps]
NR_HD long long weight_source(const LayerSpec& l, long long e) {
  const int taps = l.dims == 3 ? l.k * l.k * l.k : l.k * l.k;
  const int o = int(e % l.cout);
  const long long r = e / l.cout;
  const int c = int(r % l.cin), tap = int(r / l.cin);
  (void)taps;
  return l.transposed ? ((long long)c * l.cout + o) * taps + tap : ((long long)o * l.cin + c) * taps + tap;
}
// folded BatchNorm (ABN in eval mode: F.batch_norm with the running statistics, eps 1e-5)
NR_HD void fold_bn(float gamma, float beta, float mean, float var, float& scale, float& shift) {
  scale = gamma / sqrtf(var + 1e-5f);
  shift = beta - mean * scale;
}

struct MvsDims {
  int hr, wr;          // images as MVSNet sees them (after the optional resize)
  int h2, w2, h4, w4;  // feature resolutions
  float ratio;         // hr / h
};
// init_net.py:120-139: evaluation at >= 800 pixels resizes 768x1024 -> 576x768 and 800x800 -> 640x640, nothing else
inline MvsDims mvs_dims(int h, int w, int is_train) {
  MvsDims d;
  d.hr = h; d.wr = w; d.ratio = 1.f;
  if (!is_train && (h > w ? h : w) >= 800) {
    if (h == 768 && w == 1024) { d.hr = 576; d.wr = 768; d.ratio = 576.f / 768.f; }
    else if (h == 800 && w == 800) { d.hr = 640; d.wr = 640; d.ratio = 640.f / 800.f; }
  }
  d.h2 = (d.hr + 4 - 5) / 2 + 1; d.w2 = (d.wr + 4 - 5) / 2 + 1;
  d.h4 = (d.h2 + 4 - 5) / 2 + 1; d.w4 = (d.w2 + 4 - 5) / 2 + 1;
  return d;
}

template <class Ops>
struct MvsBuilder {
  Ops& ops;
  enc::Arena& ar;
  const MvsNet& net;
  const float* W;
  // y = layer(x) on a [D,H,W,cin] volume (D = 1 for the 2-D layers); returns the output geometry
  float* layer(int li, const float* x, int D, int H, int Wd, const float* skip, float* y, long long y_sd, long long y_sh, long long y_sw, int& Do, int& Ho,
               int& Wo, const float* in_mean = nullptr, const float* in_istd = nullptr) {
    const LayerSpec& l = net.L[li];
    ConvP p;
    p.x = x; p.w = W + l.w_off; p.scale = W + l.scale_off; p.shift = W + l.shift_off; p.skip = skip;
    p.D = D; p.H = H; p.W = Wd; p.Cin = l.cin; p.Cout = l.cout;
    const int pad = l.k / 2;
    p.kd = l.dims == 3 ? l.k : 1; p.kh = l.k; p.kw = l.k;
    p.sd = l.dims == 3 ? l.stride : 1; p.sh = l.stride; p.sw = l.stride;
    p.pd = l.dims == 3 ? pad : 0; p.ph = pad; p.pw = pad;
    p.transposed = l.transposed;
    if (l.transposed) { Do = 2 * D; Ho = 2 * H; Wo = 2 * Wd; }
    else { Do = (D + 2 * p.pd - p.kd) / p.sd + 1; Ho = (H + 2 * pad - l.k) / l.stride + 1; Wo = (Wd + 2 * pad - l.k) / l.stride + 1; }
    p.Do = Do; p.Ho = Ho; p.Wo = Wo;
    p.slope = l.has_bn ? 0.01f : 1.f;
    if (y == nullptr) { y = ar.floats((long long)Do * Ho * Wo * l.cout); y_sw = l.cout; y_sh = (long long)Wo * l.cout; y_sd = (long long)Ho * Wo * l.cout; }
    p.y = y; p.y_sd = y_sd; p.y_sh = y_sh; p.y_sw = y_sw;
    p.in_mean = in_mean; p.in_istd = in_istd;
    ops.mvs_conv(p);
    return y;
  }
};

// FeatureNet.forward (mvsnet.py:25-29) on n images [n,hr,wr,3] (raw 0..1 values; ImageNet normalisation folded into conv0) ->
// [n,h4,w4,32].  The batch rides on the depth axis of the direct convolution (kernel depth 1): eight launches for all images.
template <class Ops>
void feature_net(MvsBuilder<Ops>& b, const float* imgs, int n, int hr, int wr, const float* mean, const float* istd, float* out) {
  int D, H, W;
  const size_t mark = b.ar.used;
  float* x = b.layer(F0, imgs, n, hr, wr, nullptr, nullptr, 0, 0, 0, D, H, W, mean, istd);
  x = b.layer(F1, x, n, H, W, nullptr, nullptr, 0, 0, 0, D, H, W);
  x = b.layer(F2, x, n, H, W, nullptr, nullptr, 0, 0, 0, D, H, W);
  x = b.layer(F3, x, n, H, W, nullptr, nullptr, 0, 0, 0, D, H, W);
  x = b.layer(F4, x, n, H, W, nullptr, nullptr, 0, 0, 0, D, H, W);
  x = b.layer(F5, x, n, H, W, nullptr, nullptr, 0, 0, 0, D, H, W);
  x = b.layer(F6, x, n, H, W, nullptr, nullptr, 0, 0, 0, D, H, W);
  int Ho, Wo;
  Ho = H; Wo = W;
  b.layer(FEAT, x, n, H, W, nullptr, out, (long long)Ho * Wo * 32, (long long)Wo * 32, 32, D, H, W);
  b.ar.used = mark;        // the intermediate maps are free again (stream order keeps later layers from overwriting them early)
}

// CostRegNet.forward (mvsnet.py:52-66) on the variance volume [dn,h,w,32]; the logits go to out[(y*w + x)*dn + d]
template <class Ops>
void cost_reg_net(MvsBuilder<Ops>& b, const float* vol, int dn, int h, int w, float* out) {
  int D0, H0, W0, D1, H1, W1, D2, H2, W2, D3, H3, W3, D, H, W;
  const size_t mark = b.ar.used;
  float* c0 = b.layer(C0, vol, dn, h, w, nullptr, nullptr, 0, 0, 0, D0, H0, W0);
  float* t = b.layer(C1, c0, D0, H0, W0, nullptr, nullptr, 0, 0, 0, D1, H1, W1);
  float* c2 = b.layer(C2, t, D1, H1, W1, nullptr, nullptr, 0, 0, 0, D1, H1, W1);
  t = b.layer(C3, c2, D1, H1, W1, nullptr, nullptr, 0, 0, 0, D2, H2, W2);
  float* c4 = b.layer(C4, t, D2, H2, W2, nullptr, nullptr, 0, 0, 0, D2, H2, W2);
  t = b.layer(C5, c4, D2, H2, W2, nullptr, nullptr, 0, 0, 0, D3, H3, W3);
  t = b.layer(C6, t, D3, H3, W3, nullptr, nullptr, 0, 0, 0, D3, H3, W3);
  t = b.layer(C7, t, D3, H3, W3, c4, nullptr, 0, 0, 0, D, H, W);          // conv4 + conv7(x)
  t = b.layer(C9, t, D, H, W, c2, nullptr, 0, 0, 0, D, H, W);             // conv2 + conv9(x)
  t = b.layer(C11, t, D, H, W, c0, nullptr, 0, 0, 0, D, H, W);            // conv0 + conv11(x)
  b.layer(PROB, t, D, H, W, nullptr, out, 1, (long long)W * dn, dn, D, H, W);      // [d,y,x] -> [(y,x),d]
  b.ar.used = mark;
}

struct MvsIn {
  const float* ref_imgs; const float* src_imgs;        // [rfn,3,h,w], [sn,3,h,w] (NCHW, values 0..1)
  const float* ref_Ks; const float* ref_poses;         // [rfn,3,3], [rfn,3,4]
  const float* src_Ks; const float* src_poses;         // [sn,3,3], [sn,3,4]
  const float* depth_range;                            // [rfn,2]
  const int* nn_ids;                                   // [rfn,nn] indices into the source views
  int rfn, sn, nn, h, w, dn, is_train;
};
struct TransformsP { MvsIn in; float ratio; float* transforms; float* depth_vals; };   // [rfn,nn,12], [rfn,dn]

// construct_cost_volume_with_src (init_net.py:113-160): prob [rfn,ho,wo,dn] (the softmaxed cost volume, channel-last) and
// depth [rfn,ho,wo] (metric), ho x wo = h/4 x w/4 (of the UNRESIZED images when the evaluation resize applies)
template <class Ops>
bool mvsnet_graph(Ops& ops, enc::Arena& ar, const MvsNet& net, const float* packed, const float* consts /*mean[3] | istd[3]*/, const MvsIn& in,
                  float* prob, float* depth) {
  const MvsDims d = mvs_dims(in.h, in.w, in.is_train);
  const int resized = d.hr != in.h || d.wr != in.w;
  const int ho = resized ? in.h / 4 : d.h4, wo = resized ? in.w / 4 : d.w4;
  if (in.dn % 8 != 0 || d.h4 % 8 != 0 || d.w4 % 8 != 0) return false;      // three stride-2 / x2 stages must restore the volume's size
  MvsBuilder<Ops> b{ops, ar, net, packed};
  const long long n4 = (long long)d.h4 * d.w4;
  float* ref_feats = ar.floats((long long)in.rfn * n4 * 32);
  float* src_feats = ar.floats((long long)in.sn * n4 * 32);
  float* imgs_r = ar.floats((long long)(in.rfn + in.sn) * d.hr * d.wr * 3);
  ResizeP rp;
  rp.img = in.ref_imgs; rp.out = imgs_r; rp.N = in.rfn; rp.H = in.h; rp.W = in.w; rp.Ho = d.hr; rp.Wo = d.wr;
  ops.mvs_resize(rp);
  rp.img = in.src_imgs; rp.out = imgs_r + (long long)in.rfn * d.hr * d.wr * 3; rp.N = in.sn;
  ops.mvs_resize(rp);
  feature_net(b, imgs_r, in.rfn, d.hr, d.wr, consts, consts + 3, ref_feats);
  feature_net(b, imgs_r + (long long)in.rfn * d.hr * d.wr * 3, in.sn, d.hr, d.wr, consts, consts + 3, src_feats);
  TransformsP tp;
  tp.in = in; tp.ratio = 0.25f * d.ratio;
  tp.transforms = ar.floats((long long)in.rfn * in.nn * 12);
  tp.depth_vals = ar.floats((long long)in.rfn * in.dn);
  ops.mvs_transforms(tp);
  float* vol = ar.floats((long long)in.dn * n4 * 32);
  float* logits = ar.floats((long long)in.dn * n4);
  for (int i = 0; i < in.rfn; ++i) {
    VolumeP vp;
    vp.ref_feat = ref_feats + (long long)i * n4 * 32; vp.src_feats = src_feats; vp.nn_ids = in.nn_ids + (long long)i * in.nn;
    vp.transforms = tp.transforms + (long long)i * in.nn * 12; vp.depth_vals = tp.depth_vals + (long long)i * in.dn; vp.vol = vol;
    vp.nn = in.nn; vp.dn = in.dn; vp.h = d.h4; vp.w = d.w4;
    ops.mvs_volume(vp);
    cost_reg_net(b, vol, in.dn, d.h4, d.w4, logits);
    SoftmaxP sp;
    sp.logits = logits; sp.depth_vals = vp.depth_vals; sp.prob = prob + (long long)i * ho * wo * in.dn; sp.depth = depth + (long long)i * ho * wo;
    sp.hr = d.h4; sp.wr = d.w4; sp.ho = ho; sp.wo = wo; sp.dn = in.dn;
    ops.mvs_softmax(sp);
  }
  return ar.ok;
}

}  // namespace mvs
}  // namespace nr
