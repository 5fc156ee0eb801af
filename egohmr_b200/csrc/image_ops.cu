// Small NHWC image ops around the cuDNN convolutions of the step-invariant ResNet-50 feature provider.
#include "kernels.cuh"

namespace ehb {
namespace {

// nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (models/resnet.py:113,142) on an NHWC fp32 tensor.  HBM-bound:
// one thread = four channels of one output pixel, 128-bit loads/stores, consecutive threads walk the channel axis.
__global__ void __launch_bounds__(256) maxpool3x3s2_nhwc_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                                                int N, int H, int W, int C4, int HO, int WO) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(N) * HO * WO * C4;
  if (i >= total) return;
  const int c = static_cast<int>(i % C4);
  size_t r = i / C4;
  const int wo = static_cast<int>(r % WO);
  r /= WO;
  const int ho = static_cast<int>(r % HO);
  const int n = static_cast<int>(r / HO);
  const float ninf = -INFINITY;
  float4 m = make_float4(ninf, ninf, ninf, ninf);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int y = ho * 2 - 1 + dy;
    if (y < 0 || y >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int x = wo * 2 - 1 + dx;
      if (x < 0 || x >= W) continue;
      const float4 v = __ldg(in + ((static_cast<size_t>(n) * H + y) * W + x) * C4 + c);
      m.x = fmaxf(m.x, v.x);
      m.y = fmaxf(m.y, v.y);
      m.z = fmaxf(m.z, v.z);
      m.w = fmaxf(m.w, v.w);
    }
  }
  out[i] = m;
}

}  // namespace

cudaError_t launch_maxpool3x3s2_nhwc(const float* in, float* out, int N, int H, int W, int C, cudaStream_t stream) {
  const int HO = (H + 2 - 3) / 2 + 1, WO = (W + 2 - 3) / 2 + 1, C4 = C / 4;
  const size_t total = static_cast<size_t>(N) * HO * WO * C4;
  if (total == 0) return cudaSuccess;
  maxpool3x3s2_nhwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), N, H, W, C4, HO, WO);
  return cudaGetLastError();
}

}  // namespace ehb
