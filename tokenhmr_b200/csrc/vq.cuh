// VQ codebook kernels (tokenization/models/quantize_cnn.py:80-93).
//
// Hard quantisation (`quantize`, :80-86):  idx = argmin_k ( sum x^2 - 2 x.c_k + sum c_k^2 ), first minimum,
// int64.  The 2048 x 256 distance contraction runs on tcgen05 tensor cores in split precision: each fp32
// operand is pre-scaled by 2^6 and split into fp16 hi + lo, and  x.c ~= hi.hi + lo.hi + hi.lo  is one GEMM
// with K = 3*256 whose fp32-accumulated result differs from an fp32 dot product by ~2^-21 relative.  The
// distance matrix is never materialised: the GEMM epilogue keeps a running first-minimum per query row
// (GemmParams::argmin_out).
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace thmr {

constexpr float kVqScale = 64.0f;

// rows (R, D) fp32 -> split fp16 (R, 3*D) + per-row sum of squares (fp32).
// order_a: [hi | lo | hi] (query side);  !order_a: [hi | hi | lo] (codebook side)
__global__ void __launch_bounds__(256)
vq_split_rows_kernel(const float* __restrict__ x, __half* __restrict__ out, float* __restrict__ sq, long R, int D,
                     int order_a) {
  const long row = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + row * D;
  __half* o = out + row * (3 * D);
  float s = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    const float f[4] = {v.x * kVqScale, v.y * kVqScale, v.z * kVqScale, v.w * kVqScale};
    __half hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = __float2half_rn(f[e]);
      lo[e] = __float2half_rn(f[e] - __half2float(hi[e]));
    }
    uint2 phi, plo;
    phi.x = (static_cast<uint32_t>(__half_as_ushort(hi[1])) << 16) | __half_as_ushort(hi[0]);
    phi.y = (static_cast<uint32_t>(__half_as_ushort(hi[3])) << 16) | __half_as_ushort(hi[2]);
    plo.x = (static_cast<uint32_t>(__half_as_ushort(lo[1])) << 16) | __half_as_ushort(lo[0]);
    plo.y = (static_cast<uint32_t>(__half_as_ushort(lo[3])) << 16) | __half_as_ushort(lo[2]);
    *reinterpret_cast<uint2*>(o + c) = phi;
    *reinterpret_cast<uint2*>(o + D + c) = order_a ? plo : phi;
    *reinterpret_cast<uint2*>(o + 2 * D + c) = order_a ? phi : plo;
  }
  s = warp_sum(s);
  if (lane == 0) sq[row] = s;
}

// dequantize (F.embedding, quantize_cnn.py:88-90): out[r] = codebook[idx[r]]
__global__ void vq_gather_kernel(const long long* __restrict__ idx, const float* __restrict__ codebook,
                                 float* __restrict__ out, long R, int D4) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= R * D4) return;
  const long r = i / D4;
  reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(codebook)[idx[r] * D4 + i % D4];
}

}  // namespace thmr
