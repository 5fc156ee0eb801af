// Data movers around the ResNet-50 convolution GEMMs (conv_umma.cu): im2col of hi/lo activations, the stem's
// space-to-depth transform of the fp32 image, max-pool and global average pool on hi/lo activations.  All HBM-bound, 128-bit accesses.
#include "kernels.cuh"
#include "ptx.cuh"

namespace ehb {
namespace {

// thread = 8 channels (16 B of hi + 16 B of lo) of one (output pixel, tap)
__global__ void __launch_bounds__(256) im2col_hl_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int H,
                                                        int W, int C8, int KH, int KW, int stride, int pad, int Ho,
                                                        int Wo) {
  ptx::pdl_launch_dependents();   // the convolution GEMM that follows may run its set-up while this kernel drains
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int taps = KH * KW;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * taps * C8;
  if (i >= total) return;
  const int c = static_cast<int>(i % C8);
  size_t r = i / C8;
  const int tap = static_cast<int>(r % taps);
  r /= taps;                                   // output pixel index (n, ho, wo)
  const int wo = static_cast<int>(r % Wo);
  const int ho = static_cast<int>((r / Wo) % Ho);
  const int n = static_cast<int>(r / (static_cast<size_t>(Wo) * Ho));
  const int y = ho * stride - pad + tap / KW, x = wo * stride - pad + tap % KW;
  uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
  if (y >= 0 && y < H && x >= 0 && x < W) {
    const size_t sp = ((static_cast<size_t>(n) * H + y) * W + x) * (2 * C8);   // row of [hi(C) | lo(C)] in uint4 units
    hi = __ldg(src + sp + c);
    lo = __ldg(src + sp + C8 + c);
  }
  const size_t K8 = static_cast<size_t>(taps) * C8;                            // K / 8
  const size_t dp = r * (2 * K8) + static_cast<size_t>(tap) * C8 + c;
  dst[dp] = hi;
  dst[dp + K8] = lo;
}

// Space-to-depth of the image (thread = one 2 x 2 input block of one image): the 7x7 / stride-2 stem becomes a 4x4 /
// stride-1 convolution on 16 channels (12 used).  The four horizontal taps of an output pixel are 4 consecutive pixels =
// 64 contiguous halves of a plane, so the implicit-GEMM kernel reads one ordinary [rows][64] TMA box per vertical tap
// through an overlapping-window tensor map (pixel stride 32 B, box row 128 B); the 2 zero pad columns on either side
// replace the horizontal out-of-bounds fill, which a windowed map cannot express.
__global__ void __launch_bounds__(256) stem_s2d_kernel(const float* __restrict__ img, uint4* __restrict__ dst, int N, int H,
                                                       int W, int H2, int W2, float act_scale) {
  ptx::pdl_launch_dependents();   // the convolution GEMM that follows may run its set-up while this kernel drains
  const int W2p = W2 + 4;     // two zero pad columns on either side, written here too (no dependence on a memset)
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(N) * H2 * W2p) return;
  const int xp = static_cast<int>(i % W2p), y2 = static_cast<int>((i / W2p) % H2), n = static_cast<int>(i / (static_cast<size_t>(W2p) * H2));
  const int x2 = xp - 2;
  __align__(16) __half hi[16], lo[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    float v = 0.f;
    if (q < 12 && x2 >= 0 && x2 < W2) {
      const int c = q % 3, dx = (q / 3) & 1, dy = q / 6;
      const int y = 2 * y2 + dy, x = 2 * x2 + dx;
      if (y < H && x < W) v = __ldg(img + ((static_cast<size_t>(n) * 3 + c) * H + y) * W + x);
    }
    const float sv = v * act_scale;
    hi[q] = __float2half_rn(sv);
    lo[q] = __float2half_rn(sv - __half2float(hi[q]));
  }
  // planes [hi | lo][N][H2][W2 + 4][16 halves]: 32 bytes per pixel and plane
  const size_t plane = static_cast<size_t>(N) * H2 * W2p;
  uint4* dh = dst + i * 2;
  uint4* dl = dst + (plane + i) * 2;
  dh[0] = reinterpret_cast<const uint4*>(hi)[0];
  dh[1] = reinterpret_cast<const uint4*>(hi)[1];
  dl[0] = reinterpret_cast<const uint4*>(lo)[0];
  dl[1] = reinterpret_cast<const uint4*>(lo)[1];
}

// thread = 8 channels of one output pixel; the (hi, lo) pair with the largest value hi + lo wins (the sum of an fp16
// hi/lo split is exact in fp32, so this is exactly max over the represented values)
__global__ void __launch_bounds__(256) maxpool_hl_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int H,
                                                         int W, int C8, int Ho, int Wo) {
  ptx::pdl_launch_dependents();   // the convolution GEMM that follows may run its set-up while this kernel drains
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C8;
  if (i >= total) return;
  const int c = static_cast<int>(i % C8);
  const size_t r = i / C8;
  const int wo = static_cast<int>(r % Wo);
  const int ho = static_cast<int>((r / Wo) % Ho);
  const int n = static_cast<int>(r / (static_cast<size_t>(Wo) * Ho));
  float best[8];
  __align__(16) __half bh[8], bl[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    best[e] = -INFINITY;
    bh[e] = __float2half_rn(0.f);
    bl[e] = __float2half_rn(0.f);
  }
  for (int dy = 0; dy < 3; ++dy) {
    const int y = ho * 2 - 1 + dy;
    if (y < 0 || y >= H) continue;
    for (int dx = 0; dx < 3; ++dx) {
      const int x = wo * 2 - 1 + dx;
      if (x < 0 || x >= W) continue;
      const size_t sp = ((static_cast<size_t>(n) * H + y) * W + x) * (2 * C8);
      const uint4 h4 = __ldg(src + sp + c), l4 = __ldg(src + sp + C8 + c);
      const __half* hh = reinterpret_cast<const __half*>(&h4);
      const __half* ll = reinterpret_cast<const __half*>(&l4);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float v = __half2float(hh[e]) + __half2float(ll[e]);
        if (v > best[e]) {
          best[e] = v;
          bh[e] = hh[e];
          bl[e] = ll[e];
        }
      }
    }
  }
  const size_t dp = r * (2 * C8) + c;
  dst[dp] = *reinterpret_cast<const uint4*>(bh);
  dst[dp + C8] = *reinterpret_cast<const uint4*>(bl);
}

// x.mean(dim=(2, 3)) (models/resnet.py:148-149 avgpool + flatten): thread = one (image, channel), coalesced over channels
__global__ void __launch_bounds__(256) avgpool_hl_kernel(const __half* __restrict__ src, float* __restrict__ dst, int N, int HW,
                                                         int C, float inv_scale_hw) {
  ptx::pdl_launch_dependents();   // the convolution GEMM that follows may run its set-up while this kernel drains
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(N) * C) return;
  const int c = static_cast<int>(i % C);
  const size_t n = i / C;
  float acc = 0.f;
  for (int p = 0; p < HW; ++p) {
    const __half* row = src + (n * HW + p) * (2 * static_cast<size_t>(C));
    acc += __half2float(row[c]) + __half2float(row[C + c]);
  }
  dst[i] = acc * inv_scale_hw;
}

inline unsigned blocks_for(size_t total) { return static_cast<unsigned>((total + 255) / 256); }

}  // namespace

cudaError_t launch_im2col_hl(const __half* src, __half* dst, int N, int H, int W, int C, int KH, int KW, int stride,
                             int pad, int Ho, int Wo, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(N) * Ho * Wo * KH * KW * (C / 8);
  if (total == 0) return cudaSuccess;
  im2col_hl_kernel<<<blocks_for(total), 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), N,
                                                          H, W, C / 8, KH, KW, stride, pad, Ho, Wo);
  return cudaGetLastError();
}

cudaError_t launch_stem_s2d(const float* img, __half* dst, int N, int H, int W, float act_scale, cudaStream_t stream) {
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const size_t total = static_cast<size_t>(N) * H2 * (W2 + 4);
  if (total == 0) return cudaSuccess;
  stem_s2d_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(img, reinterpret_cast<uint4*>(dst), N, H, W,
                                                                                  H2, W2, act_scale);
  return cudaGetLastError();
}

cudaError_t launch_maxpool_hl(const __half* src, __half* dst, int N, int H, int W, int C, cudaStream_t stream) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * (C / 8);
  if (total == 0) return cudaSuccess;
  maxpool_hl_kernel<<<blocks_for(total), 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), N,
                                                           H, W, C / 8, Ho, Wo);
  return cudaGetLastError();
}

cudaError_t launch_avgpool_hl(const __half* src, float* dst, int N, int HW, int C, float act_scale, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(N) * C;
  if (total == 0) return cudaSuccess;
  avgpool_hl_kernel<<<blocks_for(total), 256, 0, stream>>>(src, dst, N, HW, C, 1.f / (act_scale * HW));
  return cudaGetLastError();
}

}  // namespace ehb
