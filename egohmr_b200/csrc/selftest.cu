// On-box bring-up harness (not part of the product path): drives the C ABI with seeded random weights and compares
// the tcgen05 fp16x3 GEMM path against the fp32 FFMA check path on the same GPU, then times the hot kernels.
//   ehb_selftest [n_img] [samples_per_img] [hid] [n_blocks]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/egohmr_b200.h"

#define CK(x)                                                            \
  do {                                                                   \
    if ((x) != 0) {                                                      \
      std::fprintf(stderr, "FAIL %s: %s\n", #x, ehb_last_error());       \
      return 2;                                                          \
    }                                                                    \
  } while (0)
#define CU(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      std::fprintf(stderr, "CUDA FAIL %s: %s\n", #x, cudaGetErrorString(e_));    \
      return 3;                                                                  \
    }                                                                            \
  } while (0)

static std::mt19937 rng(1234);
static std::vector<float> randn(size_t n, float s) {
  std::normal_distribution<float> d(0.f, s);
  std::vector<float> v(n);
  for (auto& x : v) x = d(rng);
  return v;
}
static std::vector<float> randu(size_t n, float lo, float hi) {
  std::uniform_real_distribution<float> d(lo, hi);
  std::vector<float> v(n);
  for (auto& x : v) x = d(rng);
  return v;
}
template <typename T>
static T* to_dev(const std::vector<T>& h) {
  T* p = nullptr;
  cudaMalloc(&p, h.size() * sizeof(T));
  cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  return p;
}
static double maxabs_diff(const std::vector<float>& a, const std::vector<float>& b, double* maxref) {
  double m = 0, r = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    const double d = std::fabs(double(a[i]) - double(b[i]));
    if (!(d <= m)) m = d;  // propagates NaN
    r = std::fmax(r, std::fabs(double(b[i])));
  }
  *maxref = r;
  return m;
}

int main(int argc, char** argv) {
  const int n_img = argc > 1 ? std::atoi(argv[1]) : 3;
  const int S = argc > 2 ? std::atoi(argv[2]) : 3;
  const int C = argc > 3 ? std::atoi(argv[3]) : 1024;
  const int nblk = argc > 4 ? std::atoi(argv[4]) : 4;
  const int B = n_img * S;
  const int img_dim = 2048, cond_dim = 2694, xf = 512, te = 512, in_dim = cond_dim + xf + te;
  std::printf("selftest: n_img=%d S=%d bodies=%d hid=%d blocks=%d\n", n_img, S, B, C, nblk);

  ehb_ctx* ctx = nullptr;
  CK(ehb_ctx_create(0, &ctx));

  // ---- random denoiser
  std::vector<float> adj(24 * 24, 0.f);
  for (int j = 0; j < 24; ++j) {
    adj[j * 24 + j] = 1.f;
    adj[j * 24 + (j + 1) % 24] = 0.5f;
    adj[j * 24 + (j + 5) % 24] = 0.5f;
  }
  const int n_layers = 2 * nblk + 2;
  std::vector<ehb_gconv> layers(n_layers);
  std::vector<std::vector<float>> keep;
  auto hold = [&](std::vector<float> v) {
    keep.push_back(std::move(v));
    return keep.back().data();
  };
  for (int l = 0; l < n_layers; ++l) {
    ehb_gconv& g = layers[l];
    g.in_dim = l == 0 ? in_dim : C;
    g.out_dim = l == n_layers - 1 ? 6 : C;
    const float bound = 1.414f * std::sqrt(6.f / (g.in_dim + g.out_dim));
    g.W = hold(randu(size_t(2) * g.in_dim * g.out_dim, -bound, bound));
    g.M = hold(randu(size_t(24) * g.out_dim, -0.5f, 0.5f));
    g.adj2 = hold(randn(24 * 24, 0.05f));
    g.bias = hold(randu(g.out_dim, -0.03f, 0.03f));
    if (l != n_layers - 1) {
      g.bn_weight = hold(randu(g.out_dim, 0.5f, 1.5f));
      g.bn_bias = hold(randn(g.out_dim, 0.1f));
      g.bn_mean = hold(randn(g.out_dim, 0.1f));
      g.bn_var = hold(randu(g.out_dim, 0.5f, 1.5f));
    } else {
      g.bn_weight = g.bn_bias = g.bn_mean = g.bn_var = nullptr;
    }
    g.bn_eps = 1e-5f;
  }
  ehb_gcn_weights w{};
  w.hid = C;
  w.n_blocks = nblk;
  w.img_dim = img_dim;
  w.cond_dim = cond_dim;
  w.xfeat_dim = xf;
  w.temb_dim = te;
  w.diffuse_fuse = 1;
  w.adj = adj.data();
  w.inproc_w = hold(randu(size_t(xf) * 6, -0.4f, 0.4f));
  w.inproc_b = hold(randu(xf, -0.4f, 0.4f));
  w.layers = layers.data();
  w.n_layers = n_layers;
  CK(ehb_gcn_load(ctx, &w));

  const int n_steps = 5;
  std::vector<float> coef(size_t(n_steps) * 8, 0.f);
  for (int i = 0; i < n_steps; ++i) {
    coef[i * 8 + 0] = 1.1f + i;
    coef[i * 8 + 1] = 0.5f + i;
    coef[i * 8 + 2] = 0.9f;
    coef[i * 8 + 3] = 0.3f;
  }
  CK(ehb_set_schedule(ctx, 0, n_steps, coef.data()));

  float* d_img = to_dev(randu(size_t(n_img) * img_dim, 0.f, 1.f));
  float* d_rest = to_dev(randn(size_t(n_img) * (cond_dim - img_dim), 0.5f));
  float* d_temb = to_dev(randn(size_t(n_steps) * te, 0.5f));
  std::vector<uint8_t> vis(size_t(n_img) * 24);
  for (auto& v : vis) v = (rng() % 10) < 6;
  uint8_t* d_vis = to_dev(vis);
  CK(ehb_set_cond(ctx, n_img, d_img, d_rest, d_vis, n_steps, d_temb, nullptr));
  std::vector<int32_t> iob(B);
  for (int b = 0; b < B; ++b) iob[b] = b / S;
  CK(ehb_set_bodies(ctx, B, iob.data()));

  const size_t nx = size_t(B) * 144;
  float* d_x = to_dev(randn(nx, 1.f));
  float *d_xp, *d_x0, *d_oc, *d_ou;
  CU(cudaMalloc(&d_xp, nx * 4));
  CU(cudaMalloc(&d_x0, nx * 4));
  CU(cudaMalloc(&d_oc, nx * 4));
  CU(cudaMalloc(&d_ou, nx * 4));

  std::vector<float> ref_c(nx), ref_u(nx), ref_x0(nx), ref_xp(nx), got_c(nx), got_u(nx), got_x0(nx), got_xp(nx);
  // fp32 check path
  CK(ehb_debug_set_gemm_mode(ctx, 1));
  CK(ehb_denoise_step_debug(ctx, 2, d_x, nullptr, nullptr, d_xp, d_x0, d_oc, d_ou, nullptr));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(ref_c.data(), d_oc, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ref_u.data(), d_ou, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ref_x0.data(), d_x0, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ref_xp.data(), d_xp, nx * 4, cudaMemcpyDeviceToHost));
  std::printf("check path: overflow=%d  x0[0..3] = %g %g %g %g\n", ehb_check_overflow(ctx, nullptr), ref_x0[0],
              ref_x0[1], ref_x0[2], ref_x0[3]);
  // tcgen05 path
  CK(ehb_debug_set_gemm_mode(ctx, 0));
  CU(cudaMemset(d_oc, 0, nx * 4));
  CU(cudaMemset(d_ou, 0, nx * 4));
  CK(ehb_denoise_step_debug(ctx, 2, d_x, nullptr, nullptr, d_xp, d_x0, d_oc, d_ou, nullptr));
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) {
    std::printf("tcgen05 path: CUDA error after step: %s\n", cudaGetErrorString(se));
    return 4;
  }
  CU(cudaMemcpy(got_c.data(), d_oc, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(got_u.data(), d_ou, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(got_x0.data(), d_x0, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(got_xp.data(), d_xp, nx * 4, cudaMemcpyDeviceToHost));
  std::printf("tcgen05 path: overflow=%d  x0[0..3] = %g %g %g %g\n", ehb_check_overflow(ctx, nullptr), got_x0[0],
              got_x0[1], got_x0[2], got_x0[3]);
  double r;
  const double dc = maxabs_diff(got_c, ref_c, &r);
  std::printf("max|umma - fp32| out_cond   = %.3e (max|ref| %.3e)\n", dc, r);
  const double du = maxabs_diff(got_u, ref_u, &r);
  std::printf("max|umma - fp32| out_uncond = %.3e (max|ref| %.3e)\n", du, r);
  const double d0 = maxabs_diff(got_x0, ref_x0, &r);
  std::printf("max|umma - fp32| x0         = %.3e (max|ref| %.3e)\n", d0, r);
  const double dp = maxabs_diff(got_xp, ref_xp, &r);
  std::printf("max|umma - fp32| x_prev     = %.3e (max|ref| %.3e)\n", dp, r);
  const bool ok = dc < 2e-4 * (1 + r) && du < 2e-4 * (1 + r) && std::isfinite(dc) && std::isfinite(du);
  std::printf("SELFTEST %s\n", ok ? "PASS" : "FAIL");

  // ---- timing
  for (int layer : {1, 2}) {
    float ms = 0;
    CK(ehb_time_hidden_layer(ctx, layer, 3, &ms, nullptr));
    CK(ehb_time_hidden_layer(ctx, layer, 10, &ms, nullptr));
    const double rows = double(B) * 2 * 24;
    const double flop = 2.0 * rows * C * (2.0 * C);
    std::printf("hidden layer %d: %.3f ms  -> %.1f TFLOP/s fp32-equivalent (%.1f issued fp16)\n", layer, ms,
                flop / ms * 1e-9, 3 * flop / ms * 1e-9);
  }
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) CK(ehb_denoise_step(ctx, 1, d_x, nullptr, nullptr, d_xp, d_x0, nullptr));
    cudaEventRecord(e0);
    const int it = 5;
    for (int i = 0; i < it; ++i) CK(ehb_denoise_step(ctx, 1, d_x, nullptr, nullptr, d_xp, d_x0, nullptr));
    cudaEventRecord(e1);
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::printf("full denoise step: %.3f ms  (%.0f bodies/s at 5 steps)\n", ms / it, B / (5.0 * ms / it * 1e-3));
  }
  ehb_ctx_destroy(ctx);
  return ok ? 0 : 1;
}
