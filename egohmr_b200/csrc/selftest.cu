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


#include <cuda_fp16.h>
// CPU double-precision evaluation of hidden layer `which_layer` (1 = gconv1 of block 0, 2 = gconv2 of block 0 with
// the residual) at sampled (slot, channel) pairs, from the layer's actual fp16 hi/lo input operand.
static int g_only_slot = -1, g_only_c = -1;
static void spot_check_layer(ehb_ctx* ctx, int which_layer, const ehb_gconv& g, const float* adj, int C, int n_slots,
                             float act_scale, const char* tag) {
  void *p0 = nullptr, *p1 = nullptr, *pr = nullptr;
  uint64_t b0 = 0, b1 = 0, br = 0;
  ehb_debug_get_buffer(ctx, 1, &p0, &b0);
  ehb_debug_get_buffer(ctx, 2, &p1, &b1);
  ehb_debug_get_buffer(ctx, 0, &pr, &br);
  std::vector<__half> h_a(b0 / 2), h_b(b1 / 2);
  std::vector<float> res(br / 4);
  cudaMemcpy(h_a.data(), p0, b0, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_b.data(), p1, b1, cudaMemcpyDeviceToHost);
  cudaMemcpy(res.data(), pr, br, cudaMemcpyDeviceToHost);
  const std::vector<__half>& in = which_layer == 1 ? h_a : h_b;
  auto hl = [&](const std::vector<__half>& v, size_t row, int c) {
    return (double(__half2float(v[row * 2 * C + c])) + double(__half2float(v[row * 2 * C + C + c]))) / act_scale;
  };
  float a[24][24];
  for (int i = 0; i < 24; ++i)
    for (int j = 0; j < 24; ++j) a[i][j] = adj[i * 24 + j] + g.adj2[i * 24 + j];
  double worst = 0, worst_ref = 0;
  int wslot = -1, wc = -1, wj = -1;
  double by_j[24] = {0};
  for (int slot = 0; slot < n_slots; ++slot)
    for (int c = (g_only_c >= 0 ? g_only_c : (C <= 256 ? 0 : (slot * 37) % 64)); c < C; c += (g_only_c >= 0 ? C : (C <= 256 ? 1 : 61))) {
      if (g_only_slot >= 0 && slot != g_only_slot) continue;
      const size_t row0 = size_t(slot / 5) * 128 + size_t(slot % 5) * 24;
      double h0[24], h1[24];
      for (int j = 0; j < 24; ++j) {
        double s0 = 0, s1 = 0;
        for (int kk = 0; kk < C; ++kk) {
          const double av = hl(in, row0 + j, kk);
          s0 += av * g.W[(size_t(0) * C + kk) * C + c];
          s1 += av * g.W[(size_t(1) * C + kk) * C + c];
        }
        h0[j] = s0;
        h1[j] = s1;
      }
      const double sc = double(g.bn_weight[c]) / std::sqrt(double(g.bn_var[c]) + double(g.bn_eps));
      const double sh = double(g.bn_bias[c]) + (double(g.bias[c]) - double(g.bn_mean[c])) * sc;
      for (int j = 0; j < 24; ++j) {
        double y = 0;
        for (int i = 0; i < 24; ++i) {
          const double s = (double(a[i][j]) + double(a[j][i])) / 2.0;
          y += s * g.M[i * C + c] * (i == j ? h0[i] : h1[i]);
        }
        double ref = std::fmax(y * sc + sh, 0.0);
        double got;
        if (which_layer == 1) {
          got = hl(h_b, row0 + j, c);
        } else {
          ref += hl(h_a, row0 + j, c);  // residual = block input = input-layer output
          got = res[(row0 + j) * C + c];
        }
        const double d = std::fabs(got - ref);
        if (g_only_c >= 0) std::printf("      j=%d ref=%.6f got=%.6f\n", j, ref, got);
        by_j[j] = std::fmax(by_j[j], d);
        if (d > worst) {
          worst = d; worst_ref = ref; wslot = slot; wc = c; wj = j;
        }
      }
    }
  std::printf("[%s] layer-%d vs CPU fp64: max err %.3e (ref %.4f at slot %d ch %d joint %d)\n   by joint:", tag,
              which_layer, worst, worst_ref, wslot, wc, wj);
  for (int j = 0; j < 24; ++j) std::printf(" %.0e", by_j[j]);
  std::printf("\n");
}

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
  CK(ehb_set_cond(ctx, n_img, d_img, d_rest, d_vis, nullptr));
  CK(ehb_set_temb(ctx, n_steps, d_temb, nullptr));
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
  void* res_ptr = nullptr;
  uint64_t res_bytes = 0;
  CK(ehb_debug_get_buffer(ctx, 0, &res_ptr, &res_bytes));
  std::vector<float> res_ref(res_bytes / 4), res_got(res_bytes / 4);
  // fp32 check path
  CK(ehb_debug_set_gemm_mode(ctx, 1));
  CK(ehb_denoise_step_debug(ctx, 2, d_x, nullptr, nullptr, d_xp, d_x0, d_oc, d_ou, nullptr));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(ref_c.data(), d_oc, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ref_u.data(), d_ou, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ref_x0.data(), d_x0, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ref_xp.data(), d_xp, nx * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(res_ref.data(), res_ptr, res_bytes, cudaMemcpyDeviceToHost));
  spot_check_layer(ctx, 1, layers[1], adj.data(), C, 2 * B, 8.f, "fp32 check path");
  if (nblk == 1) spot_check_layer(ctx, 2, layers[2], adj.data(), C, 2 * B, 8.f, "fp32 check path");
  std::printf("check path: overflow=%d  x0[0..3] = %g %g %g %g\n", ehb_check_overflow(ctx, nullptr), ref_x0[0],
              ref_x0[1], ref_x0[2], ref_x0[3]);
  bool all_ok = true;
  for (int gm : {2, 0}) {
  std::printf("---- tcgen05 gemm_mode=%d (%s)\n", gm, gm == 0 ? "CTA pairs, cta_group::2" : "single CTA");
  // tcgen05 path
  CK(ehb_debug_set_gemm_mode(ctx, gm));
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
  CU(cudaMemcpy(res_got.data(), res_ptr, res_bytes, cudaMemcpyDeviceToHost));
  spot_check_layer(ctx, 1, layers[1], adj.data(), C, 2 * B, 8.f, "tcgen05 path");
  if (nblk == 1) spot_check_layer(ctx, 2, layers[2], adj.data(), C, 2 * B, 8.f, "tcgen05 path");
  {
    // error map of the last block's fp32 activations: by row-in-tile (groups of 8) and by channel (groups of 32)
    const size_t rows = res_ref.size() / C;
    std::vector<double> by_row(16, 0.0), by_col(C / 32, 0.0), by_tile(rows / 128, 0.0);
    double mx = 0, mr = 0;
    size_t wr = 0;
    int wcc = 0;
    for (size_t r = 0; r < rows; ++r)
      for (int c = 0; c < C; ++c) {
        const double d = std::fabs(double(res_got[r * C + c]) - double(res_ref[r * C + c]));
        if (d > mx) { wr = r; wcc = c; }
        mr = std::fmax(mr, std::fabs(double(res_ref[r * C + c])));
        if (!(d <= mx)) mx = d;
        by_row[(r % 128) / 8] = std::fmax(by_row[(r % 128) / 8], d);
        by_col[c / 32] = std::fmax(by_col[c / 32], d);
        by_tile[r / 128] = std::fmax(by_tile[r / 128], d);
      }
    std::printf("worst res location: row %zu (tile %zu, slot-in-tile %zu, joint %zu) channel %d: check=%.6f umma=%.6f\n", wr,
                wr / 128, (wr % 128) / 24, (wr % 128) % 24, wcc, res_ref[wr * C + wcc], res_got[wr * C + wcc]);
    if (nblk == 1) {
      g_only_slot = int((wr / 128) * 5 + (wr % 128) / 24);
      g_only_c = wcc;
      spot_check_layer(ctx, 1, layers[1], adj.data(), C, 2 * B, 8.f, "tcgen05 path @worst");
      spot_check_layer(ctx, 2, layers[2], adj.data(), C, 2 * B, 8.f, "tcgen05 path @worst");
      g_only_slot = g_only_c = -1;
    }
    std::printf("res (last block fp32 activations): max err %.3e (max|ref| %.3e)\n  by row-in-tile/8:", mx, mr);
    for (double v : by_row) std::printf(" %.1e", v);
    std::printf("\n  by channel/32:");
    for (double v : by_col) std::printf(" %.1e", v);
    std::printf("\n  by m-tile:");
    for (size_t i = 0; i < by_tile.size() && i < 12; ++i) std::printf(" %.1e", by_tile[i]);
    std::printf("\n");
  }
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
  std::printf("SELFTEST gemm_mode=%d %s\n", gm, ok ? "PASS" : "FAIL");
  all_ok = all_ok && ok;

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
  }
  ehb_ctx_destroy(ctx);
  return all_ok ? 0 : 1;
}
