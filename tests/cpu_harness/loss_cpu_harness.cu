// Test infrastructure: the per-element routines of csrc/nr_loss_math.cuh on the HOST (predict_mean_for_depth_loss and the
// three losses, forward + backward), with the reductions of csrc/nr_losses.cu written as plain loops, so that
// tests/test_losses.py can check them against the reference's golden values without a GPU.  Not part of the product library.
#include "../../neuray_b200/csrc/nr_loss_math.cuh"

using namespace nr;
using namespace nr::ls;

extern "C" int nr_cpu_depth_mean(const NrDepthMeanParams* p) {
  for (long long i = 0; i < (long long)p->rfn * p->pn; ++i) depth_mean_point(*p, i);
  return 0;
}

extern "C" int nr_cpu_render_loss(const float* pr, const float* gt, const uint8_t* mask, int qn, int rn, float* loss, const float* g, float* d_pr) {
  for (int q = 0; q < qn; ++q) {
    double num = 0.0, den = 0.0;
    for (int r = 0; r < rn; ++r) {
      const long long i = (long long)q * rn + r;
      const float m = mask != nullptr ? float(mask[i] != 0) : 1.f;
      num += double(render_loss_term(pr, gt, i) * m);
      den += double(m);
    }
    const float d = mask != nullptr ? float(den) + 1e-3f : float(rn);
    if (g == nullptr) { loss[q] = mask != nullptr ? float(num) / d : float(num / rn); continue; }
    const float scale = g[q] * 2.f / d;
    for (int r = 0; r < rn; ++r) {
      const long long i = (long long)q * rn + r;
      const float m = mask != nullptr ? float(mask[i] != 0) : 1.f;
      for (int c = 0; c < 3; ++c) d_pr[3 * i + c] = scale * m * (pr[3 * i + c] - gt[3 * i + c]);
    }
  }
  return 0;
}

extern "C" int nr_cpu_depth_loss(const NrDepthLossParams* p) {
  for (int view = 0; view < p->rfn; ++view) {
    double num = 0.0, den = 0.0;
    for (int j = 0; j < p->pn; ++j) {
      float term, dterm, m;
      depth_loss_term(*p, (long long)view * p->pn + j, p->depth_pr, term, dterm, m);
      num += double(term * m);
      den += double(m);
    }
    const float d = p->aug_depth != nullptr ? float(den) + 1e-4f : float(p->pn);
    if (p->g == nullptr) { p->loss[view] = float(num) / d; continue; }
    const float scale = p->g[view] / d;
    for (int j = 0; j < p->pn; ++j) {
      const long long i = (long long)view * p->pn + j;
      float term, dterm, m;
      depth_loss_term(*p, i, p->depth_pr, term, dterm, m);
      p->d_depth_pr[i] = scale * m * dterm;
    }
  }
  return 0;
}

extern "C" int nr_cpu_consistency_loss(const float* p0, const float* p1, int qn, int rn, int dn, float* loss, const float* g, float* d_p1) {
  const long long n = (long long)rn * dn;
  for (int q = 0; q < qn; ++q) {
    const long long base = (long long)q * n;
    if (g == nullptr) {
      double num = 0.0;
      for (long long i = 0; i < n; ++i) num += double(consist_term(p0[base + i], p1[base + i]));
      loss[q] = float(num / double(n));
    } else {
      for (long long i = 0; i < n; ++i) d_p1[base + i] = g[q] / float(n) * consist_dterm(p0[base + i], p1[base + i]);
    }
  }
  return 0;
}
