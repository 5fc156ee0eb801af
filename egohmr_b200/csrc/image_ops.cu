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

// ------------------------------------------------------------------------------------------------ scene crop
// guide_coll / eval_coll hand the collision model only the scene points inside each body's axis-aligned bounding box
// (egohmr.py:550-554, 504-508): bb = [min_v verts, max_v verts]; inds = all(p >= bb_min) & all(p <= bb_max).  The
// reference does this body by body from Python (two reductions, a compare and an `.any()` host sync per body); here one
// launch handles every body: block = body, phase 1 reduces the bounding box, phase 2 classifies the image's points.
namespace ehb {
namespace {

__global__ void __launch_bounds__(256) scene_crop_kernel(const float* __restrict__ verts, int V,
                                                         const float* __restrict__ scene, int n_pts,
                                                         const int32_t* __restrict__ img_of_body,
                                                         uint8_t* __restrict__ mask, int32_t* __restrict__ count,
                                                         float* __restrict__ bbox) {
  __shared__ float red[8][6];
  __shared__ float bb[6];
  __shared__ int cnt;
  const int b = blockIdx.x;
  const float* v = verts + static_cast<size_t>(b) * V * 3;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float x = v[i * 3 + k];
      mn[k] = fminf(mn[k], x);
      mx[k] = fmaxf(mx[k], x);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], s));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], s));
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      red[warp][k] = mn[k];
      red[warp][3 + k] = mx[k];
    }
  }
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  if (threadIdx.x < 6) {
    float r = red[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) r = threadIdx.x < 3 ? fminf(r, red[w][threadIdx.x]) : fmaxf(r, red[w][threadIdx.x]);
    bb[threadIdx.x] = r;
    if (bbox) bbox[b * 6 + threadIdx.x] = r;
  }
  __syncthreads();
  const float* p = scene + static_cast<size_t>(img_of_body ? img_of_body[b] : b) * n_pts * 3;
  int local = 0;
  for (int i = threadIdx.x; i < n_pts; i += blockDim.x) {
    const float x = p[i * 3], y = p[i * 3 + 1], z = p[i * 3 + 2];
    const bool in = x >= bb[0] && y >= bb[1] && z >= bb[2] && x <= bb[3] && y <= bb[4] && z <= bb[5];
    mask[static_cast<size_t>(b) * n_pts + i] = in ? 1 : 0;
    local += in ? 1 : 0;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if (lane == 0 && local) atomicAdd(&cnt, local);
  __syncthreads();
  if (threadIdx.x == 0) count[b] = cnt;
}

}  // namespace

cudaError_t launch_scene_crop(const float* verts, int n_bodies, int V, const float* scene, int n_pts,
                              const int32_t* img_of_body, uint8_t* mask, int32_t* count, float* bbox,
                              cudaStream_t stream) {
  if (n_bodies <= 0) return cudaSuccess;
  scene_crop_kernel<<<n_bodies, 256, 0, stream>>>(verts, V, scene, n_pts, img_of_body, mask, count, bbox);
  return cudaGetLastError();
}

}  // namespace ehb
