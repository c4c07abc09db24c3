// Test infrastructure: csrc/nr_mvs_graph.cuh (MVSNet behind construct_cost_volume_with_src) executed on the HOST through the
// same __host__ __device__ per-voxel routines the CUDA kernels call (csrc/nr_mvs.cuh), including the weight re-layout and
// BatchNorm folding of the pack step.  tests/test_mvsnet.py compares it with the unmodified reference's golden.
#include <string.h>

#include <vector>

#include "../../neuray_b200/csrc/nr_mvs_graph.cuh"

namespace nr {
void set_error(const char*, ...) {}
}  // namespace nr

using namespace nr;
using namespace nr::mvs;

struct HostOps {
  template <int CG>
  static void conv_t(const ConvP& p) {
    const int groups = p.Cout / CG;
    const long long total = (long long)p.Do * p.Ho * p.Wo * groups;
    for (long long i = 0; i < total; ++i) {
      float acc[CG];
      conv_voxel<CG>(p, i / groups, int(i % groups), acc);
    }
  }
  void mvs_conv(const ConvP& p) { group_size(p.Cout) == 1 ? conv_t<1>(p) : conv_t<8>(p); }
  void mvs_volume(const VolumeP& p) { for (long long v = 0; v < (long long)p.dn * p.h * p.w; ++v) volume_voxel(p, v); }
  void mvs_softmax(const SoftmaxP& p) {
    std::vector<float> tmp(p.dn);
    for (int pix = 0; pix < p.ho * p.wo; ++pix) softmax_pixel(p, pix, tmp.data());
  }
  void mvs_resize(const ResizeP& p) { for (long long i = 0; i < (long long)p.N * p.Ho * p.Wo; ++i) resize_pixel(p, i); }
  void mvs_transforms(const TransformsP& p) {
    for (int i = 0; i < p.in.rfn; ++i) {
      for (int k = 0; k < p.in.nn; ++k) {
        const int s = p.in.nn_ids[i * p.in.nn + k];
        pair_transform(p.in.ref_Ks + 9 * i, p.in.ref_poses + 12 * i, p.in.src_Ks + 9 * s, p.in.src_poses + 12 * s, p.ratio, p.transforms + (i * p.in.nn + k) * 12);
      }
      for (int j = 0; j < p.in.dn; ++j) p.depth_vals[i * p.in.dn + j] = depth_val(p.in.depth_range[2 * i], p.in.depth_range[2 * i + 1], j, p.in.dn);
    }
  }
};

extern "C" int nr_cpu_mvsnet_tensors() {
  MvsNet net;
  build_mvsnet(net);
  return net.n_tensors;
}

// params: MVSNet.state_dict() order (89 tensors); outputs prob [rfn,ho,wo,dn], depth [rfn,ho,wo]
extern "C" int nr_cpu_mvsnet(const float* const* params, int n_params, const float* ref_imgs, const float* src_imgs, const float* ref_Ks,
                             const float* ref_poses, const float* src_Ks, const float* src_poses, const float* depth_range, const int* nn_ids, int rfn,
                             int sn, int nn, int h, int w, int dn, int is_train, float* prob, float* depth) {
  MvsNet net;
  build_mvsnet(net);
  if (n_params != net.n_tensors) return -1;
  std::vector<float> packed(net.total + 8, 0.f);
  int t = 0;
  for (int i = 0; i < N_LAYERS; ++i) {
    const LayerSpec& l = net.L[i];
    const long long taps = l.dims == 3 ? (long long)l.k * l.k * l.k : (long long)l.k * l.k;
    const float* w = params[t++];
    for (long long e = 0; e < taps * l.cin * l.cout; ++e) packed[l.w_off + e] = w[weight_source(l, e)];
    if (l.has_bn) {
      const float *g = params[t], *b = params[t + 1], *m = params[t + 2], *v = params[t + 3];
      t += 4;
      for (int o = 0; o < l.cout; ++o) fold_bn(g[o], b[o], m[o], v[o], packed[l.scale_off + o], packed[l.shift_off + o]);
    } else {
      const float* b = l.has_bias ? params[t++] : nullptr;
      for (int o = 0; o < l.cout; ++o) { packed[l.scale_off + o] = 1.f; packed[l.shift_off + o] = b ? b[o] : 0.f; }
    }
  }
  const float consts[6] = {0.485f, 0.456f, 0.406f, 1.f / 0.229f, 1.f / 0.224f, 1.f / 0.225f};
  memcpy(&packed[net.total], consts, sizeof(consts));
  MvsIn in;
  in.ref_imgs = ref_imgs; in.src_imgs = src_imgs; in.ref_Ks = ref_Ks; in.ref_poses = ref_poses; in.src_Ks = src_Ks; in.src_poses = src_poses;
  in.depth_range = depth_range; in.nn_ids = nn_ids; in.rfn = rfn; in.sn = sn; in.nn = nn; in.h = h; in.w = w; in.dn = dn; in.is_train = is_train;
  // size the workspace with a dry run
  struct Null {
    void mvs_conv(const ConvP&) {} void mvs_volume(const VolumeP&) {} void mvs_softmax(const SoftmaxP&) {} void mvs_resize(const ResizeP&) {}
    void mvs_transforms(const TransformsP&) {}
  } nops;
  enc::Arena dry{enc::DRY_BASE, ~size_t(0) / 2, 0, true};
  if (!mvsnet_graph(nops, dry, net, nullptr, nullptr, in, (float*)enc::DRY_BASE, (float*)enc::DRY_BASE)) return -2;
  std::vector<char> ws(dry.peak + 512);
  char* base = (char*)((uintptr_t(ws.data()) + 255) & ~uintptr_t(255));
  enc::Arena ar{base, dry.peak + 256, 0, true};
  HostOps ops;
  return mvsnet_graph(ops, ar, net, packed.data(), packed.data() + net.total, in, prob, depth) ? 0 : -3;
}
