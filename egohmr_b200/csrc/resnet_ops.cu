// Data movers around the ResNet-50 convolution GEMMs (conv_umma.cu): im2col of hi/lo activations, the stem's im2col
// from the fp32 image, max-pool and global average pool on hi/lo activations.  All HBM-bound, 128-bit accesses.
#include "kernels.cuh"

namespace ehb {
namespace {

// thread = 8 channels (16 B of hi + 16 B of lo) of one (output pixel, tap)
__global__ void __launch_bounds__(256) im2col_hl_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int H,
                                                        int W, int C8, int KH, int KW, int stride, int pad, int Ho,
                                                        int Wo) {
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

// The 7x7 / stride 2 / pad 3 stem (models/resnet.py:109-110).  Block = 32 consecutive output pixels of one output row:
// the 7 x 69 x 3 input patch they share is staged in shared memory with coalesced reads (each input value is used by
// up to 4 x 7 outputs), then every thread emits 8 consecutive k of one pixel, so the 384-byte hi / lo row segments of a
// pixel are written by 24 consecutive threads.
constexpr int STEM_TW = 32;
constexpr int STEM_PW = 2 * STEM_TW + 5;   // 69 input columns
__global__ void __launch_bounds__(256) im2col_stem_kernel(const float* __restrict__ img, uint4* __restrict__ dst, int N, int H,
                                                          int W, int Ho, int Wo, int Kp8, float act_scale) {
  __shared__ float patch[3][7][STEM_PW + 1];
  __shared__ int16_t koff[192];   // k -> offset of (c, ky, kx) inside `patch` (-1: K padding), instead of div/mod per element
  for (int k = threadIdx.x; k < 192; k += blockDim.x) {
    const int tap = k / 3, c = k % 3;
    koff[k] = k < 147 ? static_cast<int16_t>((c * 7 + tap / 7) * (STEM_PW + 1) + tap % 7) : static_cast<int16_t>(-1);
  }
  const int wo0 = blockIdx.x * STEM_TW, ho = blockIdx.y, n = blockIdx.z;
  const int y0 = ho * 2 - 3, x0 = wo0 * 2 - 3;
  for (int e = threadIdx.x; e < 3 * 7 * STEM_PW; e += blockDim.x) {
    const int px = e % STEM_PW, py = (e / STEM_PW) % 7, c = e / (STEM_PW * 7);
    const int y = y0 + py, x = x0 + px;
    patch[c][py][px] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + ((static_cast<size_t>(n) * 3 + c) * H + y) * W + x) : 0.f;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < STEM_TW * Kp8; item += blockDim.x) {
    const int kg = item % Kp8, pw = item / Kp8;
    const int wo = wo0 + pw;
    if (wo >= Wo) continue;
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int off = koff[kg * 8 + e];
      const float v = off >= 0 ? (&patch[0][0][0])[off + pw * 2] : 0.f;
      const float sv = v * act_scale;
      hi[e] = __float2half_rn(sv);
      lo[e] = __float2half_rn(sv - __half2float(hi[e]));
    }
    const size_t r = (static_cast<size_t>(n) * Ho + ho) * Wo + wo;
    const size_t dp = r * (2 * static_cast<size_t>(Kp8)) + kg;
    dst[dp] = *reinterpret_cast<const uint4*>(hi);
    dst[dp + Kp8] = *reinterpret_cast<const uint4*>(lo);
  }
}

// thread = 8 channels of one output pixel; the (hi, lo) pair with the largest value hi + lo wins (the sum of an fp16
// hi/lo split is exact in fp32, so this is exactly max over the represented values)
__global__ void __launch_bounds__(256) maxpool_hl_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int H,
                                                         int W, int C8, int Ho, int Wo) {
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

cudaError_t launch_im2col_stem(const float* img, __half* dst, int N, int H, int W, int Ho, int Wo, int Kp, float act_scale,
                               cudaStream_t stream) {
  if (N <= 0 || Ho <= 0 || Wo <= 0) return cudaSuccess;
  const dim3 grid((Wo + STEM_TW - 1) / STEM_TW, Ho, N);
  im2col_stem_kernel<<<grid, 256, 0, stream>>>(img, reinterpret_cast<uint4*>(dst), N, H, W, Ho, Wo, Kp / 8, act_scale);
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
