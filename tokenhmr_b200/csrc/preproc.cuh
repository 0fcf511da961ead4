// Input pre-processing on the GPU: detector boxes -> normalised 256 x 256 crops, i.e.
// ViTDetDataset.__getitem__ (tokenhmr/lib/datasets/vitdet_dataset.py:44-88) with generate_image_patch_cv2
// (tokenhmr/lib/datasets/utils.py:317-361) for every person of one frame.
//
// Everything here is byte / integer work bound by HBM traffic (786 KB written per person, at most the box area x 3
// bytes read): one thread per output pixel, coalesced plane stores, no shared memory needed.
//   * 8-bit path (box <= 2.2 x 256 px): cv2.warpAffine's fixed-point bilinear remap, bit exact: inverse map in
//     double, 10-bit fixed-point source coordinates rounded to 1/32 px, integer weights 32*a*b (sum 2^15),
//     (acc + 2^14) >> 15; then BGR -> RGB and (v - mean) / std through a 3 x 256 table the host fills in double.
//   * blurred path (larger boxes): skimage.filters.gaussian == separable float64 Gaussian with replicated edges
//     (rows, then columns), evaluated only over the box's source region and stored as fp32 (accumulated in double),
//     then the same remap with cv2's float32 weight table and a double sum.
#pragma once
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace thmr {

constexpr int kPreMaxTaps = 129;     // Gaussian radius <= 64 (sigma <= 16: a 9000-pixel box)

struct PrePerson {
  double iM[6];        // inverse affine map (dst -> src), cv2 invertAffineTransform
};

__device__ __forceinline__ int pre_sat_int(double v) {
  long long r = __double2ll_rn(v);                      // cvRound: round half to even
  r = r < -2147483648LL ? -2147483648LL : (r > 2147483647LL ? 2147483647LL : r);
  return static_cast<int>(r);
}

// Source coordinate of destination pixel (x, y): integer part and 5-bit fractions (imgwarp.cpp WarpAffineInvoker).
// Explicit _rn intrinsics: no FMA contraction, so the roundings are the host library's.
__device__ __forceinline__ void pre_src_coord(const double* iM, int x, int y, int* sx, int* sy, int* fx, int* fy) {
  const int ad = pre_sat_int(__dmul_rn(__dmul_rn(iM[0], static_cast<double>(x)), 1024.0));
  const int bd = pre_sat_int(__dmul_rn(__dmul_rn(iM[3], static_cast<double>(x)), 1024.0));
  const int X0 = pre_sat_int(__dmul_rn(__dadd_rn(__dmul_rn(iM[1], static_cast<double>(y)), iM[2]), 1024.0)) + 16;
  const int Y0 = pre_sat_int(__dmul_rn(__dadd_rn(__dmul_rn(iM[4], static_cast<double>(y)), iM[5]), 1024.0)) + 16;
  const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
  *sx = X >> 5; *sy = Y >> 5; *fx = X & 31; *fy = Y & 31;
}

// img: BGR uint8 [H, pitch]; out: fp32 [N,3,S,S] (RGB planes); patch (nullable): uint8 [N,S,S,3] (BGR, cv2's result)
// which[n] = index into persons / output slot of the n-th 8-bit person; lut: [3][256] fp32, RGB order.
__global__ void preproc_warp_u8_kernel(const uint8_t* __restrict__ img, int H, int W, long long pitch,
                                       const PrePerson* __restrict__ persons, const int* __restrict__ which,
                                       const float* __restrict__ lut, int S, float* __restrict__ out,
                                       uint8_t* __restrict__ patch) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * S) return;
  const int n = which[blockIdx.y];
  const int x = idx % S, y = idx / S;
  int sx, sy, fx, fy;
  pre_src_coord(persons[n].iM, x, y, &sx, &sy, &fx, &fy);
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const bool x0 = sx >= 0 && sx < W, x1 = sx + 1 >= 0 && sx + 1 < W;
  const bool y0 = sy >= 0 && sy < H, y1 = sy + 1 >= 0 && sy + 1 < H;
  const uint8_t* r0 = img + static_cast<long long>(sy) * pitch + 3LL * sx;
  const uint8_t* r1 = r0 + pitch;
  const size_t plane = static_cast<size_t>(S) * S;
  float* o = out + static_cast<size_t>(n) * 3 * plane + idx;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int p00 = (y0 && x0) ? __ldg(r0 + c) : 0, p01 = (y0 && x1) ? __ldg(r0 + 3 + c) : 0;
    const int p10 = (y1 && x0) ? __ldg(r1 + c) : 0, p11 = (y1 && x1) ? __ldg(r1 + 3 + c) : 0;
    const int v = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15;
    if (patch) patch[(static_cast<size_t>(n) * plane + idx) * 3 + c] = static_cast<uint8_t>(v);
    o[(2 - c) * plane] = __ldg(lut + (2 - c) * 256 + v);          // BGR -> RGB (vitdet_dataset.py:75)
  }
}

// One pass of the separable Gaussian over the region rows [y0,y1) x cols [x0,x1) of a [H,W,3] image.
// AXIS 0: in = uint8 image (pitch bytes), taps along y;  AXIS 1: in = fp32 [H,W,3], taps along x.
// Tap order as scipy's correlate1d for a symmetric kernel: centre, then mirrored pairs from the outside in.
template <int AXIS>
__global__ void preproc_gauss_kernel(const void* __restrict__ in, long long in_pitch, int H, int W,
                                     const double* __restrict__ wts, int radius, int x0, int x1, int y0, int y1,
                                     float* __restrict__ out) {
  const int rw = (x1 - x0) * 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(rw) * (y1 - y0)) return;
  const int y = y0 + static_cast<int>(idx / rw);
  const int xc = static_cast<int>(idx % rw);
  const int x = x0 + xc / 3, c = xc % 3;
  auto at = [&](int yy, int xx) -> double {
    if (AXIS == 0) return static_cast<double>(__ldg(static_cast<const uint8_t*>(in) + yy * in_pitch + 3LL * xx + c));
    return static_cast<double>(__ldg(static_cast<const float*>(in) + (static_cast<long long>(yy) * W + xx) * 3 + c));
  };
  double acc = __dmul_rn(at(y, x), wts[radius]);
  for (int k = -radius; k < 0; ++k) {
    double lo, hi;
    if (AXIS == 0) {
      lo = at(max(y + k, 0), x); hi = at(min(y - k, H - 1), x);
    } else {
      lo = at(y, max(x + k, 0)); hi = at(y, min(x - k, W - 1));
    }
    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(lo, hi), wts[k + radius]));
  }
  out[(static_cast<long long>(y) * W + x) * 3 + c] = static_cast<float>(acc);
}

// Remap of the blurred fp32 image (one person): cv2's float32 weight table, double sum, normalisation in double.
__global__ void preproc_warp_f32_kernel(const float* __restrict__ img, int H, int W, PrePerson person, int n, int S,
                                        double m0, double m1, double m2, double s0, double s1, double s2,
                                        float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * S) return;
  const int x = idx % S, y = idx / S;
  int sx, sy, fx, fy;
  pre_src_coord(person.iM, x, y, &sx, &sy, &fx, &fy);
  const float tx1 = static_cast<float>(fx) * (1.0f / 32), ty1 = static_cast<float>(fy) * (1.0f / 32);
  const float tx0 = 1.0f - tx1, ty0 = 1.0f - ty1;
  const double w00 = __fmul_rn(ty0, tx0), w01 = __fmul_rn(ty0, tx1), w10 = __fmul_rn(ty1, tx0), w11 = __fmul_rn(ty1, tx1);
  const bool x0 = sx >= 0 && sx < W, x1 = sx + 1 >= 0 && sx + 1 < W;
  const bool y0 = sy >= 0 && sy < H, y1 = sy + 1 >= 0 && sy + 1 < H;
  const float* r0 = img + (static_cast<long long>(sy) * W + sx) * 3;
  const float* r1 = r0 + 3LL * W;
  const size_t plane = static_cast<size_t>(S) * S;
  float* o = out + static_cast<size_t>(n) * 3 * plane + idx;
  const double mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};   // RGB order
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double p00 = (y0 && x0) ? r0[c] : 0.0, p01 = (y0 && x1) ? r0[3 + c] : 0.0;
    const double p10 = (y1 && x0) ? r1[c] : 0.0, p11 = (y1 && x1) ? r1[3 + c] : 0.0;
    double v = __dmul_rn(p00, w00);
    v = __dadd_rn(v, __dmul_rn(p01, w01));
    v = __dadd_rn(v, __dmul_rn(p10, w10));
    v = __dadd_rn(v, __dmul_rn(p11, w11));
    const float vf = static_cast<float>(v);                      // convert_cvimg_to_tensor: astype(float32)
    o[(2 - c) * plane] = static_cast<float>(__ddiv_rn(__dsub_rn(static_cast<double>(vf), mean[2 - c]), sd[2 - c]));
  }
}

// ---------------------------------------------------------------------------------------------- host logic
// float32 arithmetic step by step, as NumPy does for float32 arrays against Python scalars (vitdet_dataset.py:35-38,
// 51-53, 61-66; utils.py:14-33).  `volatile` keeps x87-style excess precision / re-association out of the picture.
inline float pre_bbox_size(float w, float h, int bw, int bh) {
  if (bw <= 0 || bh <= 0) return w > h ? w : h;
  volatile float ratio = h / w;
  volatile float target = static_cast<float>(static_cast<double>(bh) / static_cast<double>(bw));
  float w_new = w, h_new = h;
  if (ratio < target) {
    volatile float t = w * static_cast<float>(bh);
    volatile float t2 = t / static_cast<float>(bw);
    h_new = t2 > h ? t2 : h;
  } else {
    volatile float t = h * static_cast<float>(bw);
    volatile float t2 = t / static_cast<float>(bh);
    w_new = t2 > w ? t2 : w;
  }
  return w_new > h_new ? w_new : h_new;
}

// cv2.getAffineTransform: 6x6 system, OpenCV's LU with partial pivoting (hal::LU64f), double.
inline bool pre_get_affine(const float src[3][2], const float dst[3][2], double M[6]) {
  double a[6][6] = {{0}}, b[6];
  for (int i = 0; i < 3; ++i) {
    const int r0 = 2 * i, r1 = 2 * i + 1;
    a[r0][0] = a[r1][3] = src[i][0];
    a[r0][1] = a[r1][4] = src[i][1];
    a[r0][2] = a[r1][5] = 1.0;
    b[r0] = dst[i][0]; b[r1] = dst[i][1];
  }
  const int m = 6;
  for (int i = 0; i < m; ++i) {
    int k = i;
    for (int j = i + 1; j < m; ++j)
      if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
    if (fabs(a[k][i]) < 2.220446049250313e-15) return false;   // DBL_EPSILON * 10, as OpenCV
    if (k != i) {
      for (int c = i; c < m; ++c) std::swap(a[i][c], a[k][c]);
      std::swap(b[i], b[k]);
    }
    volatile double d = -1.0 / a[i][i];
    for (int j = i + 1; j < m; ++j) {
      volatile double alpha = a[j][i] * d;
      for (int c = i + 1; c < m; ++c) { volatile double t = alpha * a[i][c]; a[j][c] += t; }
      volatile double t = alpha * b[i];
      b[j] += t;
    }
  }
  for (int i = m - 1; i >= 0; --i) {
    double s = b[i];
    for (int c = i + 1; c < m; ++c) { volatile double t = a[i][c] * b[c]; s -= t; }
    b[i] = s / a[i][i];
  }
  for (int i = 0; i < 6; ++i) M[i] = b[i];
  return true;
}

inline void pre_invert_affine(const double M[6], double iM[6]) {
  volatile double t0 = M[0] * M[4], t1 = M[1] * M[3];
  double D = t0 - t1;
  D = D != 0 ? 1.0 / D : 0.0;
  const double A11 = M[4] * D, A22 = M[0] * D;
  iM[0] = A11; iM[1] = M[1] * (-D);
  iM[3] = M[3] * (-D); iM[4] = A22;
  volatile double u0 = -iM[0] * M[2], u1 = iM[1] * M[5];
  iM[2] = u0 - u1;
  volatile double v0 = -iM[3] * M[2], v1 = iM[4] * M[5];
  iM[5] = v0 - v1;
}

struct PreLayout {
  size_t persons_off, which_off, lut_off, wts_off, tmp_off, blur_off, total;
};
inline PreLayout pre_layout(int H, int W, int n) {
  PreLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { off = (off + 255) & ~size_t(255); const size_t o = off; off += bytes; return o; };
  L.persons_off = take(sizeof(PrePerson) * static_cast<size_t>(n > 0 ? n : 1));
  L.which_off = take(sizeof(int) * static_cast<size_t>(n > 0 ? n : 1));
  L.lut_off = take(sizeof(float) * 3 * 256);
  L.wts_off = take(sizeof(double) * kPreMaxTaps * static_cast<size_t>(n > 0 ? n : 1));   // one slot per person
  L.tmp_off = take(sizeof(float) * 3 * static_cast<size_t>(H) * W);
  L.blur_off = take(sizeof(float) * 3 * static_cast<size_t>(H) * W);
  L.total = (off + 255) & ~size_t(255);
  return L;
}

}  // namespace thmr
