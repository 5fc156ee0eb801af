// K6 — backward of (de-normalise -> rot6d -> SMPL forward -> rotation_matrix_to_angle_axis) for the collision-guided
// sampler: given dL/dvertices, dL/djoints, dL/dfull_pose(axis-angle) from the pluggable collision term, produce
// the [B][144] gradient guide_coll returns (w.r.t. the de-normalised 6-D pose, see the note in smpl_chain_bwd_kernel).
//
// Reference: EgoHMR.guide_coll (models/egohmr/egohmr.py:517-570) obtains this gradient with torch.autograd through
// x_t*std+mean (:528), rot6d_to_rotmat (utils/geometry.py:47-66), self.smpl(...) (:537, smplx lbs),
// rotation_matrix_to_angle_axis (utils/konia_transform.py:316-339,349-443,560-630).  Here it is written out:
//   smpl_skin_bwd_kernel      dL/dv -> dL/dv_posed [B][V][3]  and  dL/dA_j [B][24][3x4] (skinning transforms)
//   smpl_posedirs_bwd_kernel  dL/dpose_feature [B][207] = posedirs . dL/dv_posed
//   smpl_chain_bwd_kernel     kinematic chain reverse pass + pose-feature + angle-axis + rot6d Jacobians -> dL/dx_t
// The small per-joint Jacobians (9->3 and 6->9) are evaluated with forward-mode dual numbers over the very same
// templated functions the forward kernels use, so branch selection (torch.where) and clamps (clamp_min) differentiate
// exactly like autograd does.
#include "kernels.cuh"

namespace ehb {
namespace {

// ------------------------------------------------------------------ forward-mode scalar
struct Dual {
  float v, d;
};
__device__ __forceinline__ Dual mk(float v, float d = 0.f) { return Dual{v, d}; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const float q = a.v / b.v;
  return {q, (a.d - q * b.d) / b.v};
}
__device__ __forceinline__ Dual operator+(Dual a, float b) { return {a.v + b, a.d}; }
__device__ __forceinline__ Dual operator*(Dual a, float b) { return {a.v * b, a.d * b}; }
__device__ __forceinline__ Dual dsqrt(Dual a) {
  const float s = sqrtf(a.v);
  return {s, a.d / (2.f * s)};
}
__device__ __forceinline__ float dsqrt(float a) { return sqrtf(a); }
// x.clamp_min(m): value max(x, m); gradient passes only where x >= m (torch semantics)
__device__ __forceinline__ Dual clamp_min(Dual a, float m) { return a.v >= m ? a : Dual{m, 0.f}; }
__device__ __forceinline__ float clamp_min(float a, float m) { return fmaxf(a, m); }
__device__ __forceinline__ Dual datan2(Dual y, Dual x) {
  const float r2 = x.v * x.v + y.v * y.v;
  return {atan2f(y.v, x.v), (x.v * y.d - y.v * x.d) / r2};
}
__device__ __forceinline__ float datan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ float val(Dual a) { return a.v; }
__device__ __forceinline__ float val(float a) { return a; }
template <typename T> __device__ __forceinline__ T cst(float c);
template <> __device__ __forceinline__ float cst<float>(float c) { return c; }
template <> __device__ __forceinline__ Dual cst<Dual>(float c) { return Dual{c, 0.f}; }

// safe_zero_division (konia_transform.py:343-346): denominators with |d| < eps get eps ADDED (gradient still flows)
template <typename T>
__device__ __forceinline__ T safe_div(T num, T den, float eps = 1e-6f) {
  if (fabsf(val(den)) < eps) den = den + cst<T>(eps);
  return num / den;
}

// rotation_matrix_to_quaternion (WXYZ, eps 1e-6) followed by quaternion_to_angle_axis (konia_transform.py:349-443,560-630)
template <typename T>
__device__ __forceinline__ void rotmat_to_aa(const T m[9], T aa[3]) {
  const float eps = 1e-6f;
  const T m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
  const T trace = m00 + m11 + m22;
  T qw, qx, qy, qz;
  if (val(trace) > 0.f) {
    const T sq = dsqrt(clamp_min(trace + cst<T>(1.f), eps)) * cst<T>(2.f);
    qw = sq * cst<T>(0.25f);
    qx = safe_div(m21 - m12, sq);
    qy = safe_div(m02 - m20, sq);
    qz = safe_div(m10 - m01, sq);
  } else if (val(m00) > val(m11) && val(m00) > val(m22)) {
    const T sq = dsqrt(clamp_min(cst<T>(1.f) + m00 - m11 - m22, eps)) * cst<T>(2.f);
    qw = safe_div(m21 - m12, sq);
    qx = sq * cst<T>(0.25f);
    qy = safe_div(m01 + m10, sq);
    qz = safe_div(m02 + m20, sq);
  } else if (val(m11) > val(m22)) {
    const T sq = dsqrt(clamp_min(cst<T>(1.f) + m11 - m00 - m22, eps)) * cst<T>(2.f);
    qw = safe_div(m02 - m20, sq);
    qx = safe_div(m01 + m10, sq);
    qy = sq * cst<T>(0.25f);
    qz = safe_div(m12 + m21, sq);
  } else {
    const T sq = dsqrt(clamp_min(cst<T>(1.f) + m22 - m00 - m11, eps)) * cst<T>(2.f);
    qw = safe_div(m10 - m01, sq);
    qx = safe_div(m02 + m20, sq);
    qy = safe_div(m12 + m21, sq);
    qz = sq * cst<T>(0.25f);
  }
  const T s2 = qx * qx + qy * qy + qz * qz;
  const T sin_t = dsqrt(clamp_min(s2, eps));
  // torch_safe_atan2 (konia_transform.py:44-47): y += eps when both |y| and |x| are below eps
  T y = val(qw) < 0.f ? -sin_t : sin_t;
  T x = val(qw) < 0.f ? -qw : qw;
  if (fabsf(val(y)) < eps && fabsf(val(x)) < eps) y = y + cst<T>(eps);
  const T two_theta = datan2(y, x) * cst<T>(2.f);
  T k;
  if (val(s2) > 0.f) k = safe_div(two_theta, sin_t);
  else k = cst<T>(2.f);
  aa[0] = qx * k;
  aa[1] = qy * k;
  aa[2] = qz * k;
}

// F.normalize + Gram-Schmidt, mode 'diffusion' (utils/geometry.py:47-66); same op order as rot6d_one in smpl_lbs.cu
template <typename T>
__device__ __forceinline__ void normalize3_t(const T a[3], T b[3]) {
  const T n = dsqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const T d = clamp_min(n, 1e-12f);   // v / max(||v||, eps): below eps the denominator is the constant eps
  b[0] = a[0] / d;
  b[1] = a[1] / d;
  b[2] = a[2] / d;
}
template <typename T>
__device__ __forceinline__ void rot6d_t(const T v[6], T R[9]) {
  const T a1[3] = {v[0], v[2], v[4]};
  const T a2[3] = {v[1], v[3], v[5]};
  T b1[3], b2[3], u[3];
  normalize3_t(a1, b1);
  const T dp = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  u[0] = a2[0] - dp * b1[0];
  u[1] = a2[1] - dp * b1[1];
  u[2] = a2[2] - dp * b1[2];
  normalize3_t(u, b2);
  const T b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    R[i * 3 + 0] = b1[i];
    R[i * 3 + 1] = b2[i];
    R[i * 3 + 2] = b3[i];
  }
}

__global__ void rotmat_to_aa_kernel(const float* __restrict__ R, float* __restrict__ aa, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float m[9], o[3];
#pragma unroll
  for (int e = 0; e < 9; ++e) m[e] = R[static_cast<size_t>(i) * 9 + e];
  rotmat_to_aa<float>(m, o);
  aa[static_cast<size_t>(i) * 3 + 0] = o[0];
  aa[static_cast<size_t>(i) * 3 + 1] = o[1];
  aa[static_cast<size_t>(i) * 3 + 2] = o[2];
}

// ------------------------------------------------------------------ skinning backward
constexpr int SB_THREADS = 128;

// One block = 128 vertices of ONE body.  Recomputes v_posed and the blended transform T_v, then
//   dL/dv_posed = T_v[:, :3]^T g      dL/dA_j += w_vj * (g (x) [v_posed; 1])
// dA is reduced in shared memory and flushed with one atomicAdd per entry per block.
__global__ void __launch_bounds__(SB_THREADS) smpl_skin_bwd_kernel(const __grid_constant__ SmplDevice m,
                                                                   const float* __restrict__ betas,
                                                                   const int32_t* __restrict__ beta_index,
                                                                   const float* __restrict__ A,
                                                                   const float* __restrict__ posefeat,
                                                                   const float* __restrict__ g_verts,
                                                                   const float* __restrict__ g_joints,
                                                                   float* __restrict__ d_vposed,
                                                                   float* __restrict__ dA, int n_bodies) {
  __shared__ float pf_s[207];
  __shared__ __align__(16) float A_s[NJ][12];
  __shared__ float dA_s[NJ][12];
  __shared__ float beta_s[16];
  const int b = blockIdx.y;
  const int nj = NJ + m.n_extra;
  for (int e = threadIdx.x; e < 207; e += SB_THREADS) pf_s[e] = posefeat[static_cast<size_t>(b) * 207 + e];
  for (int e = threadIdx.x; e < NJ * 12; e += SB_THREADS) {
    (&A_s[0][0])[e] = A[static_cast<size_t>(b) * NJ * 12 + e];
    (&dA_s[0][0])[e] = 0.f;
  }
  if (threadIdx.x < 16)
    beta_s[threadIdx.x] = threadIdx.x < m.NB ? betas[static_cast<size_t>(beta_index ? beta_index[b] : b) * m.NB + threadIdx.x] : 0.f;
  __syncthreads();
  const int v = blockIdx.x * SB_THREADS + threadIdx.x;
  if (v < m.V) {
    float g[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = g_verts ? g_verts[(static_cast<size_t>(b) * m.V + v) * 3 + k] : 0.f;
    if (g_joints) {  // the vertex-picked extra joints are copies of vertices: their gradient lands on those vertices
      for (int e = 0; e < m.n_extra; ++e)
        if (m.extra_vids[e] == v) {
#pragma unroll
          for (int k = 0; k < 3; ++k) g[k] += g_joints[(static_cast<size_t>(b) * nj + NJ + e) * 3 + k];
        }
    }
    // forward recompute of v_posed
    float p[3] = {m.v_template[v * 3 + 0], m.v_template[v * 3 + 1], m.v_template[v * 3 + 2]};
    for (int l = 0; l < m.NB; ++l) {
#pragma unroll
      for (int k = 0; k < 3; ++k) p[k] = fmaf(m.shapedirs[(static_cast<size_t>(v) * 3 + k) * m.NB + l], beta_s[l], p[k]);
    }
    float po[3] = {0.f, 0.f, 0.f};
    const float* pd = m.posedirs + static_cast<size_t>(v) * 3;
    const size_t pd_ld = static_cast<size_t>(m.V) * 3;
    for (int k = 0; k < 207; ++k) {
      const float f = pf_s[k];
      po[0] = fmaf(f, __ldg(pd + k * pd_ld), po[0]);
      po[1] = fmaf(f, __ldg(pd + k * pd_ld + 1), po[1]);
      po[2] = fmaf(f, __ldg(pd + k * pd_ld + 2), po[2]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] += po[k];
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    float w[NJ];
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      w[jj] = m.lbs_weights[static_cast<size_t>(v) * NJ + jj];
      if (w[jj] != 0.f) {
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = fmaf(w[jj], A_s[jj][e], T[e]);
      }
    }
    // dL/dv_posed
#pragma unroll
    for (int c = 0; c < 3; ++c)
      d_vposed[(static_cast<size_t>(b) * m.V + v) * 3 + c] = T[0 * 4 + c] * g[0] + T[1 * 4 + c] * g[1] + T[2 * 4 + c] * g[2];
    // dL/dA_j
    if (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f) {
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        if (w[jj] != 0.f) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float wg = w[jj] * g[r];
            atomicAdd(&dA_s[jj][r * 4 + 0], wg * p[0]);
            atomicAdd(&dA_s[jj][r * 4 + 1], wg * p[1]);
            atomicAdd(&dA_s[jj][r * 4 + 2], wg * p[2]);
            atomicAdd(&dA_s[jj][r * 4 + 3], wg);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NJ * 12; e += SB_THREADS) {
    const float s = (&dA_s[0][0])[e];
    if (s != 0.f) atomicAdd(dA + static_cast<size_t>(b) * NJ * 12 + e, s);
  }
}

// dL/dpose_feature[b][k] = sum_n posedirs[k][n] * dL/dv_posed[b][n]; one warp per (k, 8-body chunk)
constexpr int PB_BODIES = 8;
__global__ void __launch_bounds__(128) smpl_posedirs_bwd_kernel(const __grid_constant__ SmplDevice m,
                                                                const float* __restrict__ d_vposed,
                                                                float* __restrict__ d_pf, int n_bodies) {
  const int k = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int b0 = blockIdx.y * PB_BODIES;
  if (k >= 207) return;
  const int N = m.V * 3;
  const float* prow = m.posedirs + static_cast<size_t>(k) * N;
  float acc[PB_BODIES];
#pragma unroll
  for (int bb = 0; bb < PB_BODIES; ++bb) acc[bb] = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float pv = __ldg(prow + n);
#pragma unroll
    for (int bb = 0; bb < PB_BODIES; ++bb)
      if (b0 + bb < n_bodies) acc[bb] = fmaf(pv, d_vposed[static_cast<size_t>(b0 + bb) * N + n], acc[bb]);
  }
#pragma unroll
  for (int bb = 0; bb < PB_BODIES; ++bb) {
    float s = acc[bb];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && b0 + bb < n_bodies) d_pf[static_cast<size_t>(b0 + bb) * 207 + k] = s;
  }
}

// One warp per body, lane j < 24 owns joint j: forward recompute (rot6d, joints, chain), then the reverse pass.
__global__ void __launch_bounds__(128) smpl_chain_bwd_kernel(const __grid_constant__ SmplDevice m,
                                                             const float* __restrict__ x, const float* __restrict__ mean,
                                                             const float* __restrict__ std_,
                                                             const float* __restrict__ betas,
                                                             const int32_t* __restrict__ beta_index,
                                                             const float* __restrict__ dA, const float* __restrict__ d_pf,
                                                             const float* __restrict__ g_joints,
                                                             const float* __restrict__ g_aa, float* __restrict__ grad_x,
                                                             int n_bodies) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= n_bodies) return;
  const int j = lane < NJ ? lane : NJ - 1;
  const int nj = NJ + m.n_extra;
  const float* beta = betas + static_cast<size_t>(beta_index ? beta_index[b] : b) * m.NB;
  // ---- forward recompute
  float v6[6], sd6[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    sd6[d] = std_[j * 6 + d];
    v6[d] = __fadd_rn(__fmul_rn(x[static_cast<size_t>(b) * XDIM + j * 6 + d], sd6[d]), mean[j * 6 + d]);
  }
  float Rj[9];
  rot6d_t<float>(v6, Rj);
  float J[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float s = 0.f;
    for (int l = 0; l < m.NB; ++l) s = fmaf(m.j_shapedirs[(j * 3 + k) * m.NB + l], beta[l], s);
    J[k] = m.j_template[j * 3 + k] + s;
  }
  const int par = m.parents[j];
  float rel[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float jp = __shfl_sync(0xffffffffu, J[k], par < 0 ? 0 : par);
    rel[k] = par < 0 ? J[k] : J[k] - jp;
  }
  float Gr[9], Gt[3];
#pragma unroll
  for (int e = 0; e < 9; ++e) Gr[e] = Rj[e];
#pragma unroll
  for (int k = 0; k < 3; ++k) Gt[k] = rel[k];
  for (int i = 1; i < NJ; ++i) {
    const int pi = m.parents[i];
    float Pr[9], Pt[3];
#pragma unroll
    for (int e = 0; e < 9; ++e) Pr[e] = __shfl_sync(0xffffffffu, Gr[e], pi);
#pragma unroll
    for (int k = 0; k < 3; ++k) Pt[k] = __shfl_sync(0xffffffffu, Gt[k], pi);
    if (lane == i) {
      float Nr[9], Nt[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          Nr[r * 3 + c] = Pr[r * 3 + 0] * Gr[0 * 3 + c] + Pr[r * 3 + 1] * Gr[1 * 3 + c] + Pr[r * 3 + 2] * Gr[2 * 3 + c];
        Nt[r] = Pr[r * 3 + 0] * Gt[0] + Pr[r * 3 + 1] * Gt[1] + Pr[r * 3 + 2] * Gt[2] + Pt[r];
      }
#pragma unroll
      for (int e = 0; e < 9; ++e) Gr[e] = Nr[e];
#pragma unroll
      for (int k = 0; k < 3; ++k) Gt[k] = Nt[k];
    }
  }
  // ---- seeds: A_j = [G_j.R | G_j.t - G_j.R J_j], joints24_j = G_j.t
  float dGr[9], dGt[3];
  {
    const float* dAj = dA + (static_cast<size_t>(b) * NJ + j) * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float dt = lane < NJ ? dAj[r * 4 + 3] : 0.f;
      dGt[r] = dt + ((g_joints && lane < NJ) ? g_joints[(static_cast<size_t>(b) * nj + j) * 3 + r] : 0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) dGr[r * 3 + c] = (lane < NJ ? dAj[r * 4 + c] : 0.f) - dt * J[c];
    }
  }
  // ---- reverse pass over the chain: G_i = G_p * [R_i | rel_i]
  float dR[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) dR[e] = 0.f;
  for (int i = NJ - 1; i >= 1; --i) {
    const int pi = m.parents[i];
    float cGr[9], cGt[3], cR[9], crel[3], pGr[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      cGr[e] = __shfl_sync(0xffffffffu, dGr[e], i);
      cR[e] = __shfl_sync(0xffffffffu, Rj[e], i);
      pGr[e] = __shfl_sync(0xffffffffu, Gr[e], pi);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      cGt[k] = __shfl_sync(0xffffffffu, dGt[k], i);
      crel[k] = __shfl_sync(0xffffffffu, rel[k], i);
    }
    if (lane == i) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          dR[a * 3 + c] = pGr[0 * 3 + a] * cGr[0 * 3 + c] + pGr[1 * 3 + a] * cGr[1 * 3 + c] + pGr[2 * 3 + a] * cGr[2 * 3 + c];
    }
    if (lane == pi) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
          dGr[r * 3 + a] += cGr[r * 3 + 0] * cR[a * 3 + 0] + cGr[r * 3 + 1] * cR[a * 3 + 1] + cGr[r * 3 + 2] * cR[a * 3 + 2] +
                            cGt[r] * crel[a];
        dGt[r] += cGt[r];
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int e = 0; e < 9; ++e) dR[e] = dGr[e];
  }
  if (lane >= NJ) return;
  // ---- pose feature (R_j - I for j >= 1)
  if (lane >= 1 && d_pf) {
#pragma unroll
    for (int e = 0; e < 9; ++e) dR[e] += d_pf[static_cast<size_t>(b) * 207 + (j - 1) * 9 + e];
  }
  // ---- axis-angle of every joint rotation (full_pose): J^T g via 9 forward-mode passes
  if (g_aa) {
    const float ga[3] = {g_aa[(static_cast<size_t>(b) * NJ + j) * 3 + 0], g_aa[(static_cast<size_t>(b) * NJ + j) * 3 + 1],
                         g_aa[(static_cast<size_t>(b) * NJ + j) * 3 + 2]};
    if (ga[0] != 0.f || ga[1] != 0.f || ga[2] != 0.f) {
#pragma unroll 1
      for (int e = 0; e < 9; ++e) {
        Dual md[9], o[3];
#pragma unroll
        for (int q = 0; q < 9; ++q) md[q] = mk(Rj[q], q == e ? 1.f : 0.f);
        rotmat_to_aa<Dual>(md, o);
        dR[e] += ga[0] * o[0].d + ga[1] * o[1].d + ga[2] * o[2].d;
      }
    }
  }
  // ---- rot6d Jacobian (6 forward-mode passes) and the de-normalisation x*std + mean
#pragma unroll 1
  for (int d = 0; d < 6; ++d) {
    Dual vd[6], Rd[9];
#pragma unroll
    for (int q = 0; q < 6; ++q) vd[q] = mk(v6[q], q == d ? 1.f : 0.f);
    rot6d_t<Dual>(vd, Rd);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 9; ++e) s = fmaf(dR[e], Rd[e].d, s);
    // NOTE (reference quirk, egohmr.py:523-528,562): guide_coll rebinds the name x_t to x_t*std+mean BEFORE calling
    // autograd.grad(..., [x_t]), so the gradient it returns is w.r.t. the de-normalised pose: no std factor here.
    grad_x[static_cast<size_t>(b) * XDIM + j * 6 + d] = s;
  }
}

}  // namespace

cudaError_t launch_rotmat_to_aa(const float* R, float* aa, int n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  rotmat_to_aa_kernel<<<(n + 127) / 128, 128, 0, stream>>>(R, aa, n);
  return cudaGetLastError();
}

cudaError_t launch_smpl_backward(const SmplDevice& m, const float* x, const float* mean, const float* std_,
                                 const float* betas, const int32_t* beta_index, const float* A, const float* posefeat,
                                 const float* g_verts, const float* g_joints, const float* g_aa, float* d_vposed,
                                 float* dA, float* d_pf, float* grad_x, int n_bodies, cudaStream_t stream) {
  if (n_bodies <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(dA, 0, static_cast<size_t>(n_bodies) * NJ * 12 * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  const bool need_skin = g_verts != nullptr || (g_joints != nullptr && m.n_extra > 0);
  if (need_skin) {
    dim3 grid((m.V + SB_THREADS - 1) / SB_THREADS, n_bodies);
    smpl_skin_bwd_kernel<<<grid, SB_THREADS, 0, stream>>>(m, betas, beta_index, A, posefeat, g_verts, g_joints, d_vposed,
                                                          dA, n_bodies);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    dim3 g2((207 + 3) / 4, (n_bodies + PB_BODIES - 1) / PB_BODIES);
    smpl_posedirs_bwd_kernel<<<g2, 128, 0, stream>>>(m, d_vposed, d_pf, n_bodies);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  smpl_chain_bwd_kernel<<<(n_bodies + 3) / 4, 128, 0, stream>>>(m, x, mean, std_, betas, beta_index, dA,
                                                                need_skin ? d_pf : nullptr, g_joints, g_aa, grad_x,
                                                                n_bodies);
  return cudaGetLastError();
}

}  // namespace ehb
