// Evaluation metrics on the GPU: the consumer of pred_vertices / pred_keypoints_3d.
//   reference: Evaluator.__call__ (lib/utils/pose_utils.py:201-275), eval_pose (:129-143),
//              reconstruction_error (:116-127), compute_similarity_transform (:61-114),
//              cam_crop_to_full (lib/utils/renderer.py:13-23).
// The reference pulls every batch to the host (torch.svd + .cpu().numpy()); here one block per sample does the pelvis
// alignment, MPJPE, the 3x3 Procrustes fit (fp64 Jacobi, one thread) and the per-vertex error without leaving the GPU.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace thmr {

// joints[b, j, :] = sum_v jreg[j, v] * verts[b, v, :]   (torch.matmul(J_regressor_24_SMPL, vertices), pose_utils.py:213,219)
// one block per (sample, joint); the regressor row and the vertices stream through L2.
__global__ void __launch_bounds__(256)
regress_joints_kernel(const float* __restrict__ jreg, const float* __restrict__ verts, float* __restrict__ out, int J,
                      int V) {
  __shared__ float red[3][8];
  const int b = blockIdx.x / J, j = blockIdx.x % J;
  const float* w = jreg + static_cast<size_t>(j) * V;
  const float* x = verts + static_cast<size_t>(b) * V * 3;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float wv = w[v];
    a0 = fmaf(wv, x[v * 3], a0);
    a1 = fmaf(wv, x[v * 3 + 1], a1);
    a2 = fmaf(wv, x[v * 3 + 2], a2);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = a0; red[1][wid] = a1; red[2][wid] = a2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    out[(static_cast<size_t>(b) * J + j) * 3 + threadIdx.x] = s;
  }
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi (fp64): A = V diag(e) V^T, eigenvalues sorted descending.
__device__ inline void jacobi_eig3(double A[3][3], double V[3][3], double e[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
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
  e[0] = A[0][0]; e[1] = A[1][1]; e[2] = A[2][2];
  for (int i = 0; i < 2; ++i)          // sort descending (columns of V follow)
    for (int j = 0; j < 2 - i; ++j)
      if (e[j] < e[j + 1]) {
        const double te = e[j]; e[j] = e[j + 1]; e[j + 1] = te;
        for (int k = 0; k < 3; ++k) { const double tv = V[k][j]; V[k][j] = V[k][j + 1]; V[k][j + 1] = tv; }
      }
}

// Rotation of the orthogonal Procrustes problem for K = X1 X2^T (3x3): with K = U S V^T,
// R = V diag(1, 1, sign det(U V^T)) U^T (pose_utils.py:94-104).  Written without the third left singular vector:
//   R = v1 u1^T + v2 u2^T + det(V) v3 (u1 x u2)^T, which equals the reference's R whatever sign its SVD picked for
// u3 and stays defined when the smallest singular value vanishes (planar point sets).
__device__ inline void procrustes_rotation(const double K[3][3], double R[3][3]) {
  double A[3][3], V[3][3], e[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = K[0][i] * K[0][j] + K[1][i] * K[1][j] + K[2][i] * K[2][j];   // K^T K
  jacobi_eig3(A, V, e);
  double u[2][3];
  for (int c = 0; c < 2; ++c) {
    for (int i = 0; i < 3; ++i) u[c][i] = K[i][0] * V[0][c] + K[i][1] * V[1][c] + K[i][2] * V[2][c];   // K v_c
    if (c == 1) {   // Gram-Schmidt against u1 (exact in exact arithmetic; guards the near-degenerate case)
      const double d = u[1][0] * u[0][0] + u[1][1] * u[0][1] + u[1][2] * u[0][2];
      for (int i = 0; i < 3; ++i) u[1][i] -= d * u[0][i];
    }
    const double n = sqrt(u[c][0] * u[c][0] + u[c][1] * u[c][1] + u[c][2] * u[c][2]);
    const double inv = n > 0.0 ? 1.0 / n : 0.0;
    for (int i = 0; i < 3; ++i) u[c][i] *= inv;
  }
  const double u3[3] = {u[0][1] * u[1][2] - u[0][2] * u[1][1], u[0][2] * u[1][0] - u[0][0] * u[1][2],
                        u[0][0] * u[1][1] - u[0][1] * u[1][0]};
  const double detV = V[0][0] * (V[1][1] * V[2][2] - V[1][2] * V[2][1]) - V[0][1] * (V[1][0] * V[2][2] - V[1][2] * V[2][0]) +
                      V[0][2] * (V[1][0] * V[2][1] - V[1][1] * V[2][0]);
  const double sg = detV >= 0.0 ? 1.0 : -1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = V[i][0] * u[0][j] + V[i][1] * u[1][j] + sg * V[i][2] * u3[j];
}

constexpr int kEvalMaxKp = 64;

// One block per sample.
//   pred_kp (B, J, 3); gt_kp (B, J, gt_stride) (gt_stride 4 = with the confidence column the reference slices off,
//   pose_utils.py:230); pelvis = (kp[pelvis_a] + kp[pelvis_b]) / 2 of each set (a == b: 3DPW branch :235-238,
//   a,b = 1,2: EMDB branch :214,220); list = keypoint_list (K <= 64 indices into J).
//   mpjpe[b] = 1000 * mean_k |p_k - g_k|,  re[b] = 1000 * mean_k |sR p_k + t - g_k|  (eval_pose :129-143)
//   pve[b]   = 1000 * mean_v |(pv - p_pelvis) - (gv - g_pelvis)|                     (:248-250), skipped if null.
__global__ void __launch_bounds__(256)
eval_pose_kernel(const float* __restrict__ pred_kp, const float* __restrict__ gt_kp, int gt_stride, int J,
                 const int* __restrict__ list, int K, int pelvis_a, int pelvis_b, const float* __restrict__ pred_v,
                 const float* __restrict__ gt_v, int V, float* __restrict__ mpjpe, float* __restrict__ re,
                 float* __restrict__ pve) {
  __shared__ float sp[kEvalMaxKp][3], sgt[kEvalMaxKp][3];
  __shared__ float pel[2][3];
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* pk = pred_kp + static_cast<size_t>(b) * J * 3;
  const float* gk = gt_kp + static_cast<size_t>(b) * J * gt_stride;
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    pel[0][c] = (pk[pelvis_a * 3 + c] + pk[pelvis_b * 3 + c]) / 2.0f;
    pel[1][c] = (gk[pelvis_a * gt_stride + c] + gk[pelvis_b * gt_stride + c]) / 2.0f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * 3; i += blockDim.x) {
    const int k = i / 3, c = i % 3, j = list[k];
    sp[k][c] = pk[j * 3 + c] - pel[0][c];
    sgt[k][c] = gk[j * gt_stride + c] - pel[1][c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // ---- MPJPE
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const float dx = sp[k][0] - sgt[k][0], dy = sp[k][1] - sgt[k][1], dz = sp[k][2] - sgt[k][2];
      acc += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    mpjpe[b] = 1000.f * (acc / K);
    // ---- similarity transform S1 -> S2 (compute_similarity_transform)
    double mu1[3] = {0, 0, 0}, mu2[3] = {0, 0, 0};
    for (int k = 0; k < K; ++k)
      for (int c = 0; c < 3; ++c) { mu1[c] += sp[k][c]; mu2[c] += sgt[k][c]; }
    for (int c = 0; c < 3; ++c) { mu1[c] /= K; mu2[c] /= K; }
    double var1 = 0.0, Km[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < K; ++k) {
      double x1[3], x2[3];
      for (int c = 0; c < 3; ++c) { x1[c] = sp[k][c] - mu1[c]; x2[c] = sgt[k][c] - mu2[c]; }
      var1 += x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Km[i][j] += x1[i] * x2[j];
    }
    double R[3][3];
    procrustes_rotation(Km, R);
    double trace = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) trace += R[i][j] * Km[j][i];
    const double scale = trace / var1;
    double t[3];
    for (int i = 0; i < 3; ++i) t[i] = mu2[i] - scale * (R[i][0] * mu1[0] + R[i][1] * mu1[1] + R[i][2] * mu1[2]);
    double err = 0.0;
    for (int k = 0; k < K; ++k) {
      double d2 = 0.0;
      for (int i = 0; i < 3; ++i) {
        const double h = scale * (R[i][0] * sp[k][0] + R[i][1] * sp[k][1] + R[i][2] * sp[k][2]) + t[i] - sgt[k][i];
        d2 += h * h;
      }
      err += sqrt(d2);
    }
    re[b] = static_cast<float>(1000.0 * err / K);
  }
  if (pve != nullptr) {
    const float* pv = pred_v + static_cast<size_t>(b) * V * 3;
    const float* gv = gt_v + static_cast<size_t>(b) * V * 3;
    float acc = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float dx = (pv[v * 3] - pel[0][0]) - (gv[v * 3] - pel[1][0]);
      const float dy = (pv[v * 3 + 1] - pel[0][1]) - (gv[v * 3 + 1] - pel[1][1]);
      const float dz = (pv[v * 3 + 2] - pel[0][2]) - (gv[v * 3 + 2] - pel[1][2]);
      acc += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i];
      pve[b] = 1000.f * (s / V);
    }
  }
}

// cam_crop_to_full (renderer.py:13-23): weak-perspective crop camera -> full-image translation.
//   cam (B,3) = [s, tx, ty]; center (B,2); size (B); img_size (B,2) = [w, h]; out (B,3) = [tx, ty, tz].
__global__ void cam_crop_to_full_kernel(const float* __restrict__ cam, const float* __restrict__ center,
                                        const float* __restrict__ size, const float* __restrict__ img_size,
                                        float focal, float* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float w_2 = img_size[b * 2] / 2.f, h_2 = img_size[b * 2 + 1] / 2.f;
  const float bs = size[b] * cam[b * 3] + 1e-9f;
  out[b * 3] = (2.f * (center[b * 2] - w_2) / bs) + cam[b * 3 + 1];
  out[b * 3 + 1] = (2.f * (center[b * 2 + 1] - h_2) / bs) + cam[b * 3 + 2];
  out[b * 3 + 2] = 2.f * focal / bs;
}

}  // namespace thmr
