// K4/K5 — rot6d -> rotation matrices and the SMPL forward pass (shape blend, joint regression, pose blend,
// kinematic chain, linear-blend skinning, 24 + n_extra joints).
//
// Reference call sites: utils/geometry.py:47-66 (rot6d_to_rotmat, mode 'diffusion'), models/egohmr/egohmr.py:258-260
// (de-normalise + convert), egohmr.py:276 (self.smpl(global_orient, body_pose, betas, pose2rot=False)).
// The SMPL arithmetic itself lives in the un-vendored dependency smplx==0.1.28 (environment.yml:197); this file
// restates the published algorithm of smplx/lbs.py::lbs and smplx/body_models.py::SMPL.forward (see oracle/smpl.py).
#include "kernels.cuh"

namespace ehb {
namespace {

// F.normalize(v, dim=1): v / max(||v||_2, 1e-12)
__device__ __forceinline__ void normalize3(const float a[3], float b[3]) {
  const float n = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const float d = fmaxf(n, 1e-12f);
  b[0] = a[0] / d;
  b[1] = a[1] / d;
  b[2] = a[2] / d;
}

// x.reshape(-1,3,2): a1 = x[:, :, 0] = (v0,v2,v4), a2 = x[:, :, 1] = (v1,v3,v5); R = stack((b1,b2,b3), dim=-1)
__device__ __forceinline__ void rot6d_one(const float v[6], float R[9]) {
  const float a1[3] = {v[0], v[2], v[4]};
  const float a2[3] = {v[1], v[3], v[5]};
  float b1[3], b2[3], u[3];
  normalize3(a1, b1);
  const float dp = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  u[0] = a2[0] - dp * b1[0];
  u[1] = a2[1] - dp * b1[1];
  u[2] = a2[2] - dp * b1[2];
  normalize3(u, b2);
  const float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    R[i * 3 + 0] = b1[i];
    R[i * 3 + 1] = b2[i];
    R[i * 3 + 2] = b3[i];
  }
}

__global__ void rot6d_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                             const float* __restrict__ std_, float* __restrict__ R, int n_rot, int period) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rot) return;
  float v[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    float t = x[static_cast<size_t>(i) * 6 + d];
    if (mean) {
      const int e = (i % period) * 6 + d;
      t = __fadd_rn(__fmul_rn(t, std_[e]), mean[e]);  // egohmr.py:258, two roundings like torch
    }
    v[d] = t;
  }
  float r[9];
  rot6d_one(v, r);
#pragma unroll
  for (int e = 0; e < 9; ++e) R[static_cast<size_t>(i) * 9 + e] = r[e];
}

// One warp per body, lane j < 24 owns joint j.
__global__ void __launch_bounds__(128) smpl_pose_kernel(const __grid_constant__ SmplDevice m,
                                                        const float* __restrict__ R, const float* __restrict__ betas,
                                                        const int32_t* __restrict__ beta_index, float* __restrict__ A,
                                                        float* __restrict__ joints24, float* __restrict__ posefeat,
                                                        int n_bodies) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= n_bodies) return;
  const int j = lane < NJ ? lane : NJ - 1;
  const float* beta = betas + static_cast<size_t>(beta_index ? beta_index[b] : b) * m.NB;
  float Rj[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) Rj[e] = R[(static_cast<size_t>(b) * NJ + j) * 9 + e];
  float J[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float s = 0.f;
    for (int l = 0; l < m.NB; ++l) s = fmaf(m.j_shapedirs[(j * 3 + k) * m.NB + l], beta[l], s);
    J[k] = m.j_template[j * 3 + k] + s;
  }
  if (lane >= 1 && lane < NJ) {
#pragma unroll
    for (int e = 0; e < 9; ++e)
      posefeat[static_cast<size_t>(b) * 207 + (j - 1) * 9 + e] = Rj[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
  }
  const int par = m.parents[j];
  float rel[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float jp = __shfl_sync(0xffffffffu, J[k], par < 0 ? 0 : par);
    rel[k] = par < 0 ? J[k] : J[k] - jp;
  }
  // G = [Gr | Gt], starts as the local transform; joints are topologically ordered (parents[i] < i)
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
  if (lane < NJ) {
    float* Ab = A + (static_cast<size_t>(b) * NJ + j) * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Ab[r * 4 + 0] = Gr[r * 3 + 0];
      Ab[r * 4 + 1] = Gr[r * 3 + 1];
      Ab[r * 4 + 2] = Gr[r * 3 + 2];
      Ab[r * 4 + 3] = Gt[r] - (Gr[r * 3 + 0] * J[0] + Gr[r * 3 + 1] * J[1] + Gr[r * 3 + 2] * J[2]);
      joints24[(static_cast<size_t>(b) * NJ + j) * 3 + r] = Gt[r];
    }
  }
}

constexpr int SKIN_BODIES = 8;
constexpr int SKIN_THREADS = 128;

// Thread = vertex, SKIN_BODIES bodies per block share the posedirs stream (pose-blend is the only real contraction of
// the SMPL pass: [B,207] x [207, 3V]).
__global__ void __launch_bounds__(SKIN_THREADS) smpl_skin_kernel(const __grid_constant__ SmplDevice m,
                                                                 const float* __restrict__ betas,
                                                                 const int32_t* __restrict__ beta_index,
                                                                 const float* __restrict__ A,
                                                                 const float* __restrict__ posefeat,
                                                                 const float* __restrict__ transl,
                                                                 float* __restrict__ verts, int n_bodies) {
  __shared__ __align__(16) float pf_s[207][SKIN_BODIES];
  __shared__ __align__(16) float A_s[SKIN_BODIES][NJ][12];
  __shared__ float beta_s[SKIN_BODIES][16];
  const int b0 = blockIdx.y * SKIN_BODIES;
  const int nb = min(SKIN_BODIES, n_bodies - b0);
  for (int e = threadIdx.x; e < 207 * SKIN_BODIES; e += SKIN_THREADS) {
    const int k = e / SKIN_BODIES, bb = e % SKIN_BODIES;
    pf_s[k][bb] = bb < nb ? posefeat[static_cast<size_t>(b0 + bb) * 207 + k] : 0.f;
  }
  for (int e = threadIdx.x; e < SKIN_BODIES * NJ * 12; e += SKIN_THREADS) {
    const int bb = e / (NJ * 12);
    (&A_s[0][0][0])[e] = bb < nb ? A[static_cast<size_t>(b0) * NJ * 12 + e] : 0.f;
  }
  for (int e = threadIdx.x; e < SKIN_BODIES * 16; e += SKIN_THREADS) {
    const int bb = e / 16, l = e % 16;
    float v = 0.f;
    if (bb < nb && l < m.NB) v = betas[static_cast<size_t>(beta_index ? beta_index[b0 + bb] : (b0 + bb)) * m.NB + l];
    beta_s[bb][l] = v;
  }
  __syncthreads();
  const int v = blockIdx.x * SKIN_THREADS + threadIdx.x;
  if (v >= m.V) return;

  // v_shaped = v_template + shapedirs . beta
  float vs[SKIN_BODIES][3];
#pragma unroll
  for (int bb = 0; bb < SKIN_BODIES; ++bb)
#pragma unroll
    for (int k = 0; k < 3; ++k) vs[bb][k] = 0.f;
  for (int l = 0; l < m.NB; ++l) {
    const float s0 = m.shapedirs[(static_cast<size_t>(v) * 3 + 0) * m.NB + l];
    const float s1 = m.shapedirs[(static_cast<size_t>(v) * 3 + 1) * m.NB + l];
    const float s2 = m.shapedirs[(static_cast<size_t>(v) * 3 + 2) * m.NB + l];
#pragma unroll
    for (int bb = 0; bb < SKIN_BODIES; ++bb) {
      const float be = beta_s[bb][l];
      vs[bb][0] = fmaf(s0, be, vs[bb][0]);
      vs[bb][1] = fmaf(s1, be, vs[bb][1]);
      vs[bb][2] = fmaf(s2, be, vs[bb][2]);
    }
  }
  const float t0 = m.v_template[v * 3 + 0], t1 = m.v_template[v * 3 + 1], t2 = m.v_template[v * 3 + 2];

  // pose_offsets = pose_feature . posedirs
  float po[SKIN_BODIES][3];
#pragma unroll
  for (int bb = 0; bb < SKIN_BODIES; ++bb)
#pragma unroll
    for (int k = 0; k < 3; ++k) po[bb][k] = 0.f;
  const float* pd = m.posedirs + static_cast<size_t>(v) * 3;
  const size_t pd_ld = static_cast<size_t>(m.V) * 3;
#pragma unroll 3
  for (int k = 0; k < 207; ++k) {
    const float p0 = __ldg(pd + k * pd_ld), p1 = __ldg(pd + k * pd_ld + 1), p2 = __ldg(pd + k * pd_ld + 2);
    const float4 fa = *reinterpret_cast<const float4*>(&pf_s[k][0]);
    const float4 fb = *reinterpret_cast<const float4*>(&pf_s[k][4]);
    const float f[SKIN_BODIES] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll
    for (int bb = 0; bb < SKIN_BODIES; ++bb) {
      po[bb][0] = fmaf(f[bb], p0, po[bb][0]);
      po[bb][1] = fmaf(f[bb], p1, po[bb][1]);
      po[bb][2] = fmaf(f[bb], p2, po[bb][2]);
    }
  }

  float w[NJ];
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) w[jj] = m.lbs_weights[static_cast<size_t>(v) * NJ + jj];

#pragma unroll
  for (int bb = 0; bb < SKIN_BODIES; ++bb) {
    if (bb >= nb) break;
    const float px = po[bb][0] + (t0 + vs[bb][0]);
    const float py = po[bb][1] + (t1 + vs[bb][1]);
    const float pz = po[bb][2] + (t2 + vs[bb][2]);
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      if (w[jj] != 0.f) {
        const float4* ap = reinterpret_cast<const float4*>(&A_s[bb][jj][0]);
        const float4 r0 = ap[0], r1 = ap[1], r2 = ap[2];
        T[0] = fmaf(w[jj], r0.x, T[0]); T[1] = fmaf(w[jj], r0.y, T[1]); T[2] = fmaf(w[jj], r0.z, T[2]);
        T[3] = fmaf(w[jj], r0.w, T[3]); T[4] = fmaf(w[jj], r1.x, T[4]); T[5] = fmaf(w[jj], r1.y, T[5]);
        T[6] = fmaf(w[jj], r1.z, T[6]); T[7] = fmaf(w[jj], r1.w, T[7]); T[8] = fmaf(w[jj], r2.x, T[8]);
        T[9] = fmaf(w[jj], r2.y, T[9]); T[10] = fmaf(w[jj], r2.z, T[10]); T[11] = fmaf(w[jj], r2.w, T[11]);
      }
    }
    float ox = T[0] * px + T[1] * py + T[2] * pz + T[3];
    float oy = T[4] * px + T[5] * py + T[6] * pz + T[7];
    float oz = T[8] * px + T[9] * py + T[10] * pz + T[11];
    if (transl) {
      ox += transl[(b0 + bb) * 3 + 0];
      oy += transl[(b0 + bb) * 3 + 1];
      oz += transl[(b0 + bb) * 3 + 2];
    }
    float* o = verts + (static_cast<size_t>(b0 + bb) * m.V + v) * 3;
    o[0] = ox;
    o[1] = oy;
    o[2] = oz;
  }
}

// ---- tensor-core route (n_bodies >= 64): the pose blend [B,207] x [207,3V] — the one dense contraction of the SMPL
// pass — runs as a conv_umma.cu GEMM (A = posedirs^T hi/lo, B = pose features hi/lo) into Y[3V][Bpad] fp32; this
// kernel turns the pose features into that B operand ...
__global__ void __launch_bounds__(256) smpl_pf_operand_kernel(const float* __restrict__ posefeat, __half* __restrict__ pf_hl,
                                                              int n_bodies, int n_pad, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (body row, k) over [n_pad][256]
  if (i >= n_pad * 256) return;
  const int b = i >> 8, k = i & 255;
  const float v = (b < n_bodies && k < 207) ? posefeat[static_cast<size_t>(b) * 207 + k] * scale : 0.f;
  const __half hi = __float2half_rn(v);
  pf_hl[static_cast<size_t>(b) * 512 + k] = hi;
  pf_hl[static_cast<size_t>(b) * 512 + 256 + k] = __float2half_rn(v - __half2float(hi));
}

// ... and this one skins: thread = vertex, block = 128 vertices x 32 bodies.  SMPL's skinning weights are sparse (at most
// four joints per vertex), so the model carries a compact (joint, weight) list per vertex, in ascending joint order: the
// blend sums exactly the terms the dense loop of smpl_skin_kernel keeps, in the same order (identical bits), with 4
// instead of 24 warp-divergent iterations.  Every global access is coalesced along the vertex axis: the GEMM writes the
// pose offsets body-major (Y[b][3v + k]), the blend-shape basis is read from a transposed copy ([3 * NB][V]), a body's
// vertices leave as 384 contiguous bytes per warp; the vertex's constants sit in registers for all 32 bodies and the
// bodies' skinning transforms are staged in shared memory with one linear copy.
constexpr int SKT_V = 128, SKT_B = 32;
__global__ void __launch_bounds__(SKT_V) smpl_skin_tiled_kernel(const __grid_constant__ SmplDevice m,
                                                                const float* __restrict__ betas,
                                                                const int32_t* __restrict__ beta_index,
                                                                const float* __restrict__ A, const float* __restrict__ Y,
                                                                size_t ldy, const float* __restrict__ transl,
                                                                float* __restrict__ verts, int n_bodies) {
  __shared__ __align__(16) float A_s[SKT_B][NJ][12];
  __shared__ float beta_s[SKT_B][10];
  const int b_begin = blockIdx.y * SKT_B;
  const int nb = min(SKT_B, n_bodies - b_begin);
  {
    const float4* src = reinterpret_cast<const float4*>(A + static_cast<size_t>(b_begin) * NJ * 12);
    float4* dst = reinterpret_cast<float4*>(&A_s[0][0][0]);
    for (int e = threadIdx.x; e < nb * NJ * 3; e += blockDim.x) dst[e] = __ldg(src + e);
    for (int e = threadIdx.x; e < SKT_B * 10; e += blockDim.x) {
      const int bb = e / 10, l = e % 10;
      float v = 0.f;
      if (bb < nb && l < m.NB) v = betas[static_cast<size_t>(beta_index ? beta_index[b_begin + bb] : (b_begin + bb)) * m.NB + l];
      beta_s[bb][l] = v;
    }
  }
  __syncthreads();
  const int v = blockIdx.x * SKT_V + threadIdx.x;
  if (v >= m.V) return;
  float sd[3][10];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int l = 0; l < 10; ++l) sd[k][l] = l < m.NB ? __ldg(m.shapedirs_t + static_cast<size_t>(k * m.NB + l) * m.V + v) : 0.f;
  const float4 wq = __ldg(reinterpret_cast<const float4*>(m.skin_w4) + v);
  const uchar4 jq = __ldg(reinterpret_cast<const uchar4*>(m.skin_j4) + v);
  const float w4[4] = {wq.x, wq.y, wq.z, wq.w};
  const int j4[4] = {jq.x, jq.y, jq.z, jq.w};
  const float t0 = m.v_template[v * 3 + 0], t1 = m.v_template[v * 3 + 1], t2 = m.v_template[v * 3 + 2];
#pragma unroll 4
  for (int bb = 0; bb < nb; ++bb) {
    const int b = b_begin + bb;
    const float* yp = Y + static_cast<size_t>(b) * ldy + static_cast<size_t>(v) * 3;   // pose offsets of (body, vertex)
    const float y0 = __ldg(yp), y1 = __ldg(yp + 1), y2 = __ldg(yp + 2);
    float vs0 = 0.f, vs1 = 0.f, vs2 = 0.f;   // same summation order as smpl_skin_kernel
#pragma unroll
    for (int l = 0; l < 10; ++l) {
      const float be = beta_s[bb][l];
      vs0 = fmaf(sd[0][l], be, vs0);
      vs1 = fmaf(sd[1][l], be, vs1);
      vs2 = fmaf(sd[2][l], be, vs2);
    }
    const float px = y0 + (t0 + vs0);
    const float py = y1 + (t1 + vs1);
    const float pz = y2 + (t2 + vs2);
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (w4[q] != 0.f) {
        const float4* ap = reinterpret_cast<const float4*>(&A_s[bb][j4[q]][0]);
        const float4 r0 = ap[0], r1 = ap[1], r2 = ap[2];
        T[0] = fmaf(w4[q], r0.x, T[0]); T[1] = fmaf(w4[q], r0.y, T[1]); T[2] = fmaf(w4[q], r0.z, T[2]);
        T[3] = fmaf(w4[q], r0.w, T[3]); T[4] = fmaf(w4[q], r1.x, T[4]); T[5] = fmaf(w4[q], r1.y, T[5]);
        T[6] = fmaf(w4[q], r1.z, T[6]); T[7] = fmaf(w4[q], r1.w, T[7]); T[8] = fmaf(w4[q], r2.x, T[8]);
        T[9] = fmaf(w4[q], r2.y, T[9]); T[10] = fmaf(w4[q], r2.z, T[10]); T[11] = fmaf(w4[q], r2.w, T[11]);
      }
    }
    float ox = T[0] * px + T[1] * py + T[2] * pz + T[3];
    float oy = T[4] * px + T[5] * py + T[6] * pz + T[7];
    float oz = T[8] * px + T[9] * py + T[10] * pz + T[11];
    if (transl) {
      ox += transl[b * 3 + 0];
      oy += transl[b * 3 + 1];
      oz += transl[b * 3 + 2];
    }
    float* o = verts + (static_cast<size_t>(b) * m.V + v) * 3;
    o[0] = ox;
    o[1] = oy;
    o[2] = oz;
  }
}

__global__ void smpl_joints_kernel(const __grid_constant__ SmplDevice m, const float* __restrict__ joints24,
                                   const float* __restrict__ verts, const float* __restrict__ transl,
                                   float* __restrict__ joints, int n_bodies) {
  const int nj = NJ + m.n_extra;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bodies * nj) return;
  const int b = i / nj, j = i % nj;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v;
    if (j < NJ) {
      v = joints24[(static_cast<size_t>(b) * NJ + j) * 3 + k];
      if (transl) v += transl[b * 3 + k];
    } else {
      v = verts[(static_cast<size_t>(b) * m.V + m.extra_vids[j - NJ]) * 3 + k];  // already translated
    }
    joints[static_cast<size_t>(i) * 3 + k] = v;
  }
}

}  // namespace

cudaError_t launch_rot6d(const float* x, const float* mean, const float* std_, float* R, int n_bodies,
                         cudaStream_t stream) {
  const int n = n_bodies * NJ;
  if (n <= 0) return cudaSuccess;
  rot6d_kernel<<<(n + 127) / 128, 128, 0, stream>>>(x, mean, std_, R, n, NJ);
  return cudaGetLastError();
}

cudaError_t launch_rot6d_flat(const float* x6, float* R, int n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  rot6d_kernel<<<(n + 127) / 128, 128, 0, stream>>>(x6, nullptr, nullptr, R, n, 1);
  return cudaGetLastError();
}

cudaError_t launch_smpl_pose(const SmplDevice& m, const float* R, const float* betas, const int32_t* beta_index,
                             float* A, float* joints24, float* posefeat, int n_bodies, cudaStream_t stream) {
  if (n_bodies <= 0) return cudaSuccess;
  smpl_pose_kernel<<<(n_bodies + 3) / 4, 128, 0, stream>>>(m, R, betas, beta_index, A, joints24, posefeat, n_bodies);
  return cudaGetLastError();
}

cudaError_t launch_smpl_skin(const SmplDevice& m, const float* betas, const int32_t* beta_index, const float* A,
                             const float* posefeat, const float* transl, float* verts, int n_bodies,
                             cudaStream_t stream) {
  if (n_bodies <= 0) return cudaSuccess;
  if (m.NB > 16) return cudaErrorInvalidValue;
  dim3 grid((m.V + SKIN_THREADS - 1) / SKIN_THREADS, (n_bodies + SKIN_BODIES - 1) / SKIN_BODIES);
  smpl_skin_kernel<<<grid, SKIN_THREADS, 0, stream>>>(m, betas, beta_index, A, posefeat, transl, verts, n_bodies);
  return cudaGetLastError();
}

cudaError_t launch_smpl_pf_operand(const float* posefeat, __half* pf_hl, int n_bodies, int n_pad, float scale,
                                   cudaStream_t stream) {
  if (n_pad <= 0) return cudaSuccess;
  smpl_pf_operand_kernel<<<(n_pad * 256 + 255) / 256, 256, 0, stream>>>(posefeat, pf_hl, n_bodies, n_pad, scale);
  return cudaGetLastError();
}

cudaError_t launch_smpl_skin_tiled(const SmplDevice& m, const float* betas, const int32_t* beta_index, const float* A,
                                   const float* Y, size_t ldy, const float* transl, float* verts, int n_bodies,
                                   cudaStream_t stream) {
  if (n_bodies <= 0) return cudaSuccess;
  if (m.NB > 10 || !m.skin_w4 || !m.shapedirs_t) return cudaErrorInvalidValue;
  dim3 grid((m.V + SKT_V - 1) / SKT_V, (n_bodies + SKT_B - 1) / SKT_B);
  smpl_skin_tiled_kernel<<<grid, SKT_V, 0, stream>>>(m, betas, beta_index, A, Y, ldy, transl, verts, n_bodies);
  return cudaGetLastError();
}

cudaError_t launch_smpl_joints(const SmplDevice& m, const float* joints24, const float* verts, const float* transl,
                               float* joints, int n_bodies, cudaStream_t stream) {
  const int n = n_bodies * (NJ + m.n_extra);
  if (n <= 0) return cudaSuccess;
  smpl_joints_kernel<<<(n + 127) / 128, 128, 0, stream>>>(m, joints24, verts, transl, joints, n_bodies);
  return cudaGetLastError();
}

}  // namespace ehb
