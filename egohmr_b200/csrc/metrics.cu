// Evaluation-side kernels: batched Procrustes alignment for PA-MPJPE.
//
// Reference: utils/pose_utils.py:11-73 compute_similarity_transform(_batch) (and :75-105 with a visibility mask) — a
// Python loop over samples calling numpy's SVD, fed by `.cpu().numpy()` copies (test_egohmr.py:420-433).  After the
// sampler got fast that loop dominates an evaluation run; here block = one (S1, S2) problem, the 3 x 3 algebra in
// double precision on one thread (Jacobi eigen-decomposition of K^T K), everything else block-parallel.
#include "kernels.cuh"

namespace ehb {
namespace {

__device__ inline void jacobi_eigen_sym3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {   // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {   // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {   // V <- V J
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

__device__ inline double det3(const double M[3][3]) {
  return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// scale * R and t of the similarity transform that maps S1 onto S2 (pose_utils.py:25-52), from the centred moments:
// K = X1 X2^T (3x3), var1 = sum |X1|^2, mu1, mu2.
__device__ void solve_similarity(const double K[3][3], double var1, const double mu1[3], const double mu2[3],
                                 double sR[3][3], double t[3]) {
  // K = U S V^T  =>  K^T K = V S^2 V^T
  double A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double a = 0;
      for (int k = 0; k < 3; ++k) a += K[k][i] * K[k][j];
      A[i][j] = a;
    }
  jacobi_eigen_sym3(A, V);
  int ord[3] = {0, 1, 2};   // descending eigenvalues (numpy orders singular values that way; Z acts on the smallest)
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (A[ord[j]][ord[j]] > A[ord[i]][ord[i]]) {
        const int tmp = ord[i];
        ord[i] = ord[j];
        ord[j] = tmp;
      }
  double Vs[3][3], U[3][3], sv[3];
  for (int c = 0; c < 3; ++c) {
    sv[c] = sqrt(fmax(A[ord[c]][ord[c]], 0.0));
    for (int r = 0; r < 3; ++r) Vs[r][c] = V[r][ord[c]];
  }
  const double tol = 1e-12 * fmax(sv[0], 1e-300);
  for (int c = 0; c < 3; ++c) {
    double u[3];
    if (sv[c] > tol) {
      for (int r = 0; r < 3; ++r) u[r] = (K[r][0] * Vs[0][c] + K[r][1] * Vs[1][c] + K[r][2] * Vs[2][c]) / sv[c];
    } else if (c == 2) {   // rank-2 K: complete the basis (its sign is irrelevant, Z fixes det(R) = +1)
      u[0] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
      u[1] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
      u[2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    } else {               // rank <= 1: any unit vector orthogonal to the previous columns
      const double a0 = c == 0 ? 1.0 : -U[1][0], a1 = c == 0 ? 0.0 : U[0][0];
      const double nrm = sqrt(a0 * a0 + a1 * a1);
      u[0] = nrm > 1e-12 ? a0 / nrm : 0.0;
      u[1] = nrm > 1e-12 ? a1 / nrm : 0.0;
      u[2] = nrm > 1e-12 ? 0.0 : 1.0;
    }
    for (int r = 0; r < 3; ++r) U[r][c] = u[r];
  }
  double UVt[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) UVt[i][j] = U[i][0] * Vs[j][0] + U[i][1] * Vs[j][1] + U[i][2] * Vs[j][2];
  const double d = det3(UVt);
  const double z = d > 0 ? 1.0 : (d < 0 ? -1.0 : 0.0);   // np.sign
  double R[3][3];                                          // R = V Z U^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + z * Vs[i][2] * U[j][2];
  double tr = 0;                                           // trace(R K)
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) tr += R[i][k] * K[k][i];
  const double scale = tr / var1;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) sR[i][j] = scale * R[i][j];
    t[i] = mu2[i] - (sR[i][0] * mu1[0] + sR[i][1] * mu1[1] + sR[i][2] * mu1[2]);
  }
}

constexpr int PT = 128;

__device__ inline double block_sum(double v, double* red) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
  for (int w = 0; w < PT / 32; ++w) r += red[w];
  return r;
}

__global__ void __launch_bounds__(PT) procrustes_kernel(const float* __restrict__ S1, const float* __restrict__ S2,
                                                        const float* __restrict__ mask, int n_pts,
                                                        float* __restrict__ S1_hat, float* __restrict__ err) {
  __shared__ double red[PT / 32];
  __shared__ double sRt[12];
  const size_t base = static_cast<size_t>(blockIdx.x) * n_pts;
  const float* a = S1 + base * 3;
  const float* b = S2 + base * 3;
  const float* m = mask ? mask + base * 3 : nullptr;
  // masked copies exactly as pose_utils.py:78-79 (S*vis_mask, the mean still divides by N)
  double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < n_pts; i += PT)
    for (int k = 0; k < 3; ++k) {
      const double w = m ? m[i * 3 + k] : 1.0;
      s1[k] += w * a[i * 3 + k];
      s2[k] += w * b[i * 3 + k];
    }
  double mu1[3], mu2[3];
  for (int k = 0; k < 3; ++k) {
    mu1[k] = block_sum(s1[k], red) / n_pts;
    mu2[k] = block_sum(s2[k], red) / n_pts;
  }
  double Kp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, vp = 0;
  for (int i = threadIdx.x; i < n_pts; i += PT) {
    double x1[3], x2[3];
    for (int k = 0; k < 3; ++k) {
      const double w = m ? m[i * 3 + k] : 1.0;
      x1[k] = w * a[i * 3 + k] - mu1[k];
      x2[k] = w * b[i * 3 + k] - mu2[k];
      vp += x1[k] * x1[k];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Kp[r * 3 + c] += x1[r] * x2[c];
  }
  double K[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) K[r][c] = block_sum(Kp[r * 3 + c], red);
  const double var1 = block_sum(vp, red);
  if (threadIdx.x == 0) {
    double sR[3][3], t[3];
    solve_similarity(K, var1, mu1, mu2, sR, t);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) sRt[r * 3 + c] = sR[r][c];
      sRt[9 + r] = t[r];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_pts; i += PT) {   // S1_hat = scale R S1 + t on the UNMASKED points (:100)
    const double x = a[i * 3], y = a[i * 3 + 1], z = a[i * 3 + 2];
    float h[3];
    double e2 = 0;
    for (int r = 0; r < 3; ++r) {
      const double v = sRt[r * 3] * x + sRt[r * 3 + 1] * y + sRt[r * 3 + 2] * z + sRt[9 + r];
      h[r] = static_cast<float>(v);
      const double d = static_cast<double>(h[r]) - b[i * 3 + r];
      e2 += d * d;
    }
    if (S1_hat) {
      S1_hat[(base + i) * 3] = h[0];
      S1_hat[(base + i) * 3 + 1] = h[1];
      S1_hat[(base + i) * 3 + 2] = h[2];
    }
    if (err) err[base + i] = static_cast<float>(sqrt(e2));
  }
}

}  // namespace

cudaError_t launch_procrustes(const float* S1, const float* S2, const float* mask, int n_problems, int n_pts,
                              float* S1_hat, float* err, cudaStream_t stream) {
  if (n_problems <= 0 || n_pts <= 0) return cudaSuccess;
  procrustes_kernel<<<n_problems, PT, 0, stream>>>(S1, S2, mask, n_pts, S1_hat, err);
  return cudaGetLastError();
}

}  // namespace ehb

// ------------------------------------------------------------------------------------------------ 1-NN distances
// The contact score of the evaluation driver (test_egohmr.py:496-506) takes, for every predicted body, the squared
// distance from each vertex to its nearest scene point: `chamfer_distance(verts, scene)` of
// utils/pytorch3d_chamfer_distance.py:69-222, whose arithmetic is pytorch3d's brute-force `knn_points(K=1)` (third-party
// CUDA, not vendored in the reference).  Block = (pair, 128 query points); the reference cloud streams through shared
// memory in 1024-point tiles, every thread keeps the running minimum of its query point.  Compute-bound (8 flop per
// point pair); the clouds may be indexed (all samples of an image share one scene cloud — no `.repeat`).
namespace ehb {
namespace {

constexpr int NN_Q = 128, NN_TILE = 1024;

__global__ void __launch_bounds__(NN_Q) nn_dist_sq_kernel(const float* __restrict__ q, const int32_t* __restrict__ q_index,
                                                          int n_q, const float* __restrict__ r,
                                                          const int32_t* __restrict__ r_index, int n_r,
                                                          float* __restrict__ out) {
  __shared__ float rs[NN_TILE * 3];
  const int pair = blockIdx.y;
  const float* qc = q + static_cast<size_t>(q_index ? q_index[pair] : pair) * n_q * 3;
  const float* rc = r + static_cast<size_t>(r_index ? r_index[pair] : pair) * n_r * 3;
  const int i = blockIdx.x * NN_Q + threadIdx.x;
  const bool live = i < n_q;
  const float x = live ? qc[i * 3] : 0.f, y = live ? qc[i * 3 + 1] : 0.f, z = live ? qc[i * 3 + 2] : 0.f;
  float best = INFINITY;
  for (int t0 = 0; t0 < n_r; t0 += NN_TILE) {
    const int nt = min(NN_TILE, n_r - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < nt * 3; e += NN_Q) rs[e] = rc[static_cast<size_t>(t0) * 3 + e];
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < nt; ++j) {
      const float dx = x - rs[j * 3], dy = y - rs[j * 3 + 1], dz = z - rs[j * 3 + 2];
      best = fminf(best, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    }
  }
  if (live) out[static_cast<size_t>(pair) * n_q + i] = best;
}

}  // namespace

cudaError_t launch_nn_dist_sq(const float* q, const int32_t* q_index, int n_q, const float* r, const int32_t* r_index,
                              int n_r, int n_pairs, float* out, cudaStream_t stream) {
  if (n_pairs <= 0 || n_q <= 0) return cudaSuccess;
  dim3 grid((n_q + NN_Q - 1) / NN_Q, n_pairs);
  nn_dist_sq_kernel<<<grid, NN_Q, 0, stream>>>(q, q_index, n_q, r, r_index, n_r, out);
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------ evaluation reductions
// The metric block of the reference driver (test_egohmr.py:373-494) after the sampler: visibility masks of the ground
// truth (:375-388), G-MPJPE / MPJPE / V2V per (image, sample) with their visible / invisible splits (:398-447), per-joint
// standard deviation and average pairwise distance over the samples of an image (:449-494).  The reference evaluates them
// with ~60 small torch launches, `.cpu().numpy()` copies and Python loops over the images of a batch.
namespace {

// mask[b][i] = 1 iff gt point i of image b projects into the 1920 x 1080 image (perspective_projection with identity
// rotation and zero translation, utils/geometry.py:78-116; test_egohmr.py:375-388)
__global__ void vis_mask_kernel(const float* __restrict__ pts, const float* __restrict__ focal, const float* __restrict__ cx,
                                const float* __restrict__ cy, uint8_t* __restrict__ mask, int n_pts, float W, float H) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const float* p = pts + (static_cast<size_t>(b) * n_pts + i) * 3;
  const float z = p[2];
  const float px = __fdiv_rn(p[0], z), py = __fdiv_rn(p[1], z), pz = __fdiv_rn(z, z);
  const float f = focal[b];
  // K . [px, py, pz]: x = f*px + 0*py + cx*pz (einsum order), y = 0*px + f*py + cy*pz
  const float x = __fadd_rn(__fadd_rn(__fmul_rn(f, px), __fmul_rn(0.f, py)), __fmul_rn(cx[b], pz));
  const float y = __fadd_rn(__fadd_rn(__fmul_rn(0.f, px), __fmul_rn(f, py)), __fmul_rn(cy[b], pz));
  mask[static_cast<size_t>(b) * n_pts + i] = (x >= 0.f && x < W && y >= 0.f && y < H) ? 1 : 0;
}

constexpr int EM_T = 256;

__device__ inline float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < EM_T / 32; ++i) t += red[i];
  return t;
}

// block = (image, sample): out[b*S+s][0..8] = g_mpjpe (mean), g_vis (sum), g_invis (sum), mpjpe, mpjpe_vis, mpjpe_invis,
// v2v (mean), v2v_vis (sum), v2v_invis (sum)
__global__ void __launch_bounds__(EM_T) pose_error_kernel(const float* __restrict__ pj, const float* __restrict__ pv,
                                                          const float* __restrict__ transl, const float* __restrict__ gj,
                                                          const float* __restrict__ gv, const uint8_t* __restrict__ jmask,
                                                          const uint8_t* __restrict__ vmask, float* __restrict__ out, int S,
                                                          int J, int V) {
  __shared__ float red[EM_T / 32];
  const int bs = blockIdx.x, b = bs / S, t = threadIdx.x;
  const float* pjb = pj + static_cast<size_t>(bs) * J * 3;
  const float* gjb = gj + static_cast<size_t>(b) * J * 3;
  const float pp[3] = {pjb[0], pjb[1], pjb[2]}, gp[3] = {gjb[0], gjb[1], gjb[2]};   // pelvis = joint 0
  const float tr[3] = {transl[b * 3], transl[b * 3 + 1], transl[b * 3 + 2]};
  float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = t; j < J; j += EM_T) {
    float dg = 0.f, da = 0.f;
    for (int c = 0; c < 3; ++c) {
      const float p = pjb[j * 3 + c], g = gjb[j * 3 + c];
      const float eg = (p + tr[c]) - g;                    // pred_keypoints_3d_full - gt_keypoints_3d
      const float ea = (p - pp[c]) - (g - gp[c]);          // pelvis-aligned
      dg += eg * eg;
      da += ea * ea;
    }
    dg = sqrtf(dg);
    da = sqrtf(da);
    const bool vis = jmask[b * J + j] != 0;
    acc[0] += dg; acc[1] += vis ? dg : 0.f; acc[2] += vis ? 0.f : dg;
    acc[3] += da; acc[4] += vis ? da : 0.f; acc[5] += vis ? 0.f : da;
  }
  const float* pvb = pv + static_cast<size_t>(bs) * V * 3;
  const float* gvb = gv + static_cast<size_t>(b) * V * 3;
  for (int i = t; i < V; i += EM_T) {
    float d = 0.f;
    for (int c = 0; c < 3; ++c) {
      const float e = (pvb[i * 3 + c] - pp[c]) - (gvb[i * 3 + c] - gp[c]);
      d += e * e;
    }
    d = sqrtf(d);
    const bool vis = vmask[static_cast<size_t>(b) * V + i] != 0;
    acc[6] += d; acc[7] += vis ? d : 0.f; acc[8] += vis ? 0.f : d;
  }
  for (int k = 0; k < 9; ++k) {
    const float s = block_sum(acc[k], red);
    if (t == 0) out[static_cast<size_t>(bs) * 9 + k] = (k == 0 || k == 3) ? s / J : (k == 6 ? s / V : s);
  }
}

// block = image: out[b][0..5] = std_joints (all / visible / invisible joints), apd_joints (all / visible / invisible).
// thread j < J handles joint j: unbiased std over the S samples of each coordinate of the pelvis-aligned joint, and the
// sum over ordered sample pairs of the joint's distance.  Empty joint sets give NaN like the reference (mean of nothing).
__global__ void __launch_bounds__(EM_T) diversity_kernel(const float* __restrict__ pj, const uint8_t* __restrict__ jmask,
                                                         float* __restrict__ out, int S, int J) {
  __shared__ float red[EM_T / 32];
  const int b = blockIdx.x, t = threadIdx.x;
  float sd = 0.f, pd = 0.f;
  bool vis = false;
  if (t < J) {
    vis = jmask[b * J + t] != 0;
    const float* base = pj + static_cast<size_t>(b) * S * J * 3;
    for (int c = 0; c < 3; ++c) {
      float mean = 0.f;
      for (int s = 0; s < S; ++s) mean += base[(s * J + t) * 3 + c] - base[(s * J) * 3 + c];
      mean /= S;
      float var = 0.f;
      for (int s = 0; s < S; ++s) {
        const float d = (base[(s * J + t) * 3 + c] - base[(s * J) * 3 + c]) - mean;
        var += d * d;
      }
      sd += sqrtf(var / (S - 1));
    }
    sd /= 3.f;
    for (int s1 = 0; s1 < S; ++s1)
      for (int s2 = 0; s2 < S; ++s2) {
        float d = 0.f;
        for (int c = 0; c < 3; ++c) {
          const float e = (base[(s1 * J + t) * 3 + c] - base[(s1 * J) * 3 + c]) - (base[(s2 * J + t) * 3 + c] - base[(s2 * J) * 3 + c]);
          d += e * e;
        }
        pd += sqrtf(d);
      }
  }
  const float in = t < J ? 1.f : 0.f, v = (t < J && vis) ? 1.f : 0.f;
  const float n_vis = block_sum(v, red), n_inv = block_sum(in - v, red);
  const float s_all = block_sum(sd * in, red), s_vis = block_sum(sd * v, red), s_inv = block_sum(sd * (in - v), red);
  const float p_all = block_sum(pd * in, red), p_vis = block_sum(pd * v, red), p_inv = block_sum(pd * (in - v), red);
  if (t == 0) {
    const float pairs = static_cast<float>(S) * (S - 1) * 2.f;      // the reference divides by n (n - 1) and by 2 (:472)
    float* o = out + static_cast<size_t>(b) * 6;
    o[0] = s_all / J;
    o[1] = s_vis / n_vis;
    o[2] = s_inv / n_inv;
    o[3] = p_all / J / pairs;
    o[4] = p_vis / n_vis / pairs;
    o[5] = p_inv / n_inv / pairs;
  }
}

}  // namespace

cudaError_t launch_vis_mask(const float* pts, const float* focal, const float* cx, const float* cy, uint8_t* mask, int n_img,
                            int n_pts, float W, float H, cudaStream_t stream) {
  if (n_img <= 0 || n_pts <= 0) return cudaSuccess;
  vis_mask_kernel<<<dim3((n_pts + 255) / 256, n_img), 256, 0, stream>>>(pts, focal, cx, cy, mask, n_pts, W, H);
  return cudaGetLastError();
}

cudaError_t launch_pose_errors(const float* pj, const float* pv, const float* transl, const float* gj, const float* gv,
                               const uint8_t* jmask, const uint8_t* vmask, float* out, int n_img, int S, int J, int V,
                               cudaStream_t stream) {
  if (n_img <= 0 || S <= 0) return cudaSuccess;
  pose_error_kernel<<<n_img * S, EM_T, 0, stream>>>(pj, pv, transl, gj, gv, jmask, vmask, out, S, J, V);
  return cudaGetLastError();
}

cudaError_t launch_diversity(const float* pj, const uint8_t* jmask, float* out, int n_img, int S, int J, cudaStream_t stream) {
  if (n_img <= 0) return cudaSuccess;
  if (J > EM_T) return cudaErrorInvalidValue;
  diversity_kernel<<<n_img, EM_T, 0, stream>>>(pj, jmask, out, S, J);
  return cudaGetLastError();
}

}  // namespace ehb
