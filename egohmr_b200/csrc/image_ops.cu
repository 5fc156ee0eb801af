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

// EgoHMR.forward's step-invariant glue (egohmr.py:186-205, 220-223, 262-263) for every image in one launch:
//   vis[j]  = openpose confidence of the joint mapped to SMPL joint j > 0, OpenPose joint 8 forced visible (:186-189)
//   rest    = [scene_feats | transl_feat | cam_cx/f, cam_cy/f | box_cx/f, box_cy/f, box_size/f | fx]   (f = fx * coeff)
//   ctxfull = [img_feats | rest]   (the beta head's input, :263)
// The reference builds these with ~18 small torch kernels per call; the arithmetic is the same IEEE fp32 ops.
struct CondInputsParams {
  const float* kp2d;        // [n][25][3]
  const float* scene_feat;  // [n][sf]
  const float* transl_feat; // [n][tf]
  const float* img_feat;    // [n][img_dim]
  const float* fx;          // [n] (normalised), or null when no camera flag is set
  const float* box_center;  // [n][2] or null
  const float* box_size;    // [n] or null
  const float* cam_cx;      // [n] or null
  const float* cam_cy;      // [n] or null
  uint8_t* vis;             // [n][24]
  float* rest;              // [n][rest_dim]
  float* ctxfull;           // [n][img_dim + rest_dim]
  int sf, tf, img_dim, with_focal, with_bbox, with_center;
  float coeff;
  int o2s[NJ];
};

__global__ void __launch_bounds__(256) cond_inputs_kernel(const __grid_constant__ CondInputsParams p) {
  const int n = blockIdx.x, t = threadIdx.x;
  const int ncam = (p.with_center ? 2 : 0) + (p.with_bbox ? 3 : 0) + (p.with_focal ? 1 : 0);
  const int rest_dim = p.sf + p.tf + ncam;
  float* rest = p.rest + static_cast<size_t>(n) * rest_dim;
  float* full = p.ctxfull + static_cast<size_t>(n) * (p.img_dim + rest_dim);
  if (t < NJ) {
    const int jo = p.o2s[t];
    p.vis[n * NJ + t] = (jo == 8 || p.kp2d[(static_cast<size_t>(n) * 25 + jo) * 3 + 2] > 0.f) ? 1 : 0;
  }
  for (int i = t; i < p.img_dim; i += blockDim.x) full[i] = p.img_feat[static_cast<size_t>(n) * p.img_dim + i];
  for (int i = t; i < rest_dim; i += blockDim.x) {
    float v;
    if (i < p.sf) {
      v = p.scene_feat[static_cast<size_t>(n) * p.sf + i];
    } else if (i < p.sf + p.tf) {
      v = p.transl_feat[static_cast<size_t>(n) * p.tf + i - p.sf];
    } else {
      int c = i - p.sf - p.tf;
      const float f = p.fx ? __fmul_rn(p.fx[n], p.coeff) : 1.f;
      if (p.with_center) {
        if (c < 2) { v = __fdiv_rn(c == 0 ? p.cam_cx[n] : p.cam_cy[n], f); c = -1; }
        else c -= 2;
      }
      if (c >= 0 && p.with_bbox) {
        if (c < 3) { v = __fdiv_rn(c < 2 ? p.box_center[n * 2 + c] : p.box_size[n], f); c = -1; }
        else c -= 3;
      }
      if (c >= 0) v = p.fx[n];
    }
    rest[i] = v;
    full[p.img_dim + i] = v;
  }
}

// egohmr.py:277-301 for the final joints: full-frame 3-D joints, their perspective projection
// (utils/geometry.py:78-116 with identity rotation) normalised to the 1920 x 1080 frame, plus the per-body focal length /
// camera centre rows compute_loss reads.  Thread = (body, joint).
__global__ void project_joints_kernel(const float* __restrict__ joints, const float* __restrict__ transl,
                                      const float* __restrict__ fx, const float* __restrict__ cam_cx,
                                      const float* __restrict__ cam_cy, const int32_t* __restrict__ img_of_body, int n_bodies,
                                      int J, float coeff, float default_focal, float* __restrict__ kp3d_full,
                                      float* __restrict__ kp2d, float* __restrict__ focal_out, float* __restrict__ center_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bodies * J) return;
  const int b = i / J, j = i % J, img = img_of_body ? img_of_body[b] : b;
  const float f = fx ? __fmul_rn(fx[img], coeff) : default_focal;
  const float cx = fx ? cam_cx[img] : 960.f, cy = fx ? cam_cy[img] : 540.f;
  const float x = __fadd_rn(joints[i * 3 + 0], transl[img * 3 + 0]);
  const float y = __fadd_rn(joints[i * 3 + 1], transl[img * 3 + 1]);
  const float z = __fadd_rn(joints[i * 3 + 2], transl[img * 3 + 2]);
  kp3d_full[i * 3 + 0] = x;
  kp3d_full[i * 3 + 1] = y;
  kp3d_full[i * 3 + 2] = z;
  const float px = __fdiv_rn(x, z), py = __fdiv_rn(y, z), pz = __fdiv_rn(z, z);
  const float u = __fadd_rn(__fmul_rn(px, f), __fmul_rn(cx, pz));
  const float v = __fadd_rn(__fmul_rn(py, f), __fmul_rn(cy, pz));
  kp2d[i * 2 + 0] = __fsub_rn(__fdiv_rn(u, 1920.f), 0.5f);
  kp2d[i * 2 + 1] = __fsub_rn(__fdiv_rn(v, 1080.f), 0.5f);
  if (j == 0) {
    focal_out[b * 2 + 0] = f;
    focal_out[b * 2 + 1] = f;
    center_out[b * 2 + 0] = cx;
    center_out[b * 2 + 1] = cy;
  }
}

}  // namespace

cudaError_t launch_cond_inputs(const float* kp2d, const float* scene_feat, const float* transl_feat, const float* img_feat,
                               const float* fx, const float* box_center, const float* box_size, const float* cam_cx,
                               const float* cam_cy, int n, int sf, int tf, int img_dim, int with_focal, int with_bbox,
                               int with_center, const int32_t* o2s, float coeff, uint8_t* vis, float* rest, float* ctxfull,
                               cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  CondInputsParams p{kp2d, scene_feat, transl_feat, img_feat, fx, box_center, box_size, cam_cx, cam_cy, vis, rest, ctxfull,
                     sf, tf, img_dim, with_focal, with_bbox, with_center, coeff, {}};
  for (int j = 0; j < NJ; ++j) p.o2s[j] = o2s[j];
  cond_inputs_kernel<<<n, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_project_joints(const float* joints, const float* transl, const float* fx, const float* cam_cx,
                                  const float* cam_cy, const int32_t* img_of_body, int n_bodies, int J, float coeff,
                                  float default_focal, float* kp3d_full, float* kp2d, float* focal_out, float* center_out,
                                  cudaStream_t stream) {
  const int total = n_bodies * J;
  if (total <= 0) return cudaSuccess;
  project_joints_kernel<<<(total + 255) / 256, 256, 0, stream>>>(joints, transl, fx, cam_cx, cam_cy, img_of_body, n_bodies, J,
                                                                 coeff, default_focal, kp3d_full, kp2d, focal_out, center_out);
  return cudaGetLastError();
}

cudaError_t launch_scene_crop(const float* verts, int n_bodies, int V, const float* scene, int n_pts,
                              const int32_t* img_of_body, uint8_t* mask, int32_t* count, float* bbox,
                              cudaStream_t stream) {
  if (n_bodies <= 0) return cudaSuccess;
  scene_crop_kernel<<<n_bodies, 256, 0, stream>>>(verts, V, scene, n_pts, img_of_body, mask, count, bbox);
  return cudaGetLastError();
}

}  // namespace ehb
