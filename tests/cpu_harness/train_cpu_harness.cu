// Test infrastructure: runs the per-point / per-sample / per-ray routines of csrc/nr_train_math.cuh on the HOST (nvcc
// compiles the same __host__ __device__ code for the CPU) so that tests/test_backward_cpu.py can check the hand-written
// backward against PyTorch autograd without a GPU.  Not part of the product library.
#include "../../neuray_b200/csrc/nr_train_math.cuh"

extern "C" int nr_self_cpu(const NrSelfParams* p) {
  for (long long r = 0; r < p->rn; ++r) nr::tr::self_hit_prob_ray(*p, r);
  return 0;
}

extern "C" int nr_train_cpu(const NrPassParams* p, const NrBwdParams* b) {
  using namespace nr::tr;
  const long long N = (long long)p->rn * p->dn, R = N * p->rfn;
  Ctx c;
  c.p = *p;
  c.n_heads = p->use_vis ? 4 : 3;
  c.W = p->w_point;
  c.Wr = p->w_ray;
  c.tr = {b->tape_row, R, R_SLOTS};
  c.gr = {b->grad_row, R, G_SLOTS};
  c.tp = {b->tape_point, N, P_SLOTS};
  c.gp = {b->grad_point, N, GP_SLOTS};
  c.d_feat = b->d_feat;
  c.d_pix = b->d_pixel_colors;
  c.d_hit = b->d_hit_prob;
  c.d_depth = b->d_render_depth;
  // the same phase order as nr_render_pass_bwd's kernel launches (csrc/nr_train.cu)
  for (long long r = 0; r < R; ++r) row_forward_a(c, r);
  for (long long n = 0; n < N; ++n) point_forward_b(c, n);
  for (long long r = 0; r < R; ++r) row_forward_c(c, r);
  for (long long n = 0; n < N; ++n) point_forward_d(c, n);
  for (long long n = 0; n < N; ++n) sample_forward(c, n);
  for (long long r = 0; r < p->rn; ++r) ray_backward(c, r);
  for (long long n = 0; n < N; ++n) sample_backward_q(c, n);
  for (long long n = 0; n < N; ++n) { sample_backward_kv(c, n); point_backward_a(c, n); }
  for (long long r = 0; r < R; ++r) row_backward_b(c, r);
  for (long long n = 0; n < N; ++n) point_backward_c(c, n);
  for (long long r = 0; r < R; ++r) row_backward_d(c, r);
  return 0;
}
