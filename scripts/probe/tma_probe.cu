// How long does the TMA engine need for one head's Q/K/V (3 x [192 x 80] fp16) with different box shapes?
//   mode 0: 3 loads of a 3-D box {16, 192, 5}, SWIZZLE_32B  (960 32-byte rows each)       <- attention kernels r1
//   mode 1: per matrix: 2-D box {64, 192} SWIZZLE_128B + 2-D box {16, 192} SWIZZLE_32B   (192 x 128 B + 192 x 32 B)
#include <cstdio>
#include <vector>
#include "../../tokenhmr_b200/csrc/common.cuh"
#include "../../tokenhmr_b200/csrc/ptx.cuh"
using namespace thmr;
constexpr int kIters = 7;
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tm3, const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16,
      int mode, int heads, int nprob, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int i = 0;
    for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x, ++i) {
      const int b = prob / heads, h = prob % heads;
      mbar_arrive_expect_tx(&bar, 3 * 30720);
      for (int m = 0; m < 3; ++m) {
        uint8_t* dst = smem + m * 30720;
        if (mode == 0) {
          tma_load_3d(dst, &tm3, &bar, 0, b * 192, (m * heads + h) * 5);
        } else {
          tma_load_2d(dst, &tm64, &bar, (m * heads + h) * 80, b * 192);
          tma_load_2d(dst + 24576, &tm16, &bar, (m * heads + h) * 80 + 64, b * 192);
        }
      }
      mbar_wait(&bar, i & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  const int B = 64, H = 16, M = B * 192, LD = 3 * H * 80;
  __half* qkv; cudaMalloc(&qkv, size_t(M) * LD * 2); cudaMemset(qkv, 0, size_t(M) * LD * 2);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  CUtensorMap tm3, tm64, tm16;
  EncodeTiledFn fn = encode_tiled_fn();
  { cuuint64_t gd[3] = {16, (cuuint64_t)M, (cuuint64_t)LD / 16}; cuuint64_t gs[2] = {(cuuint64_t)LD * 2, 32}; cuuint32_t bx[3] = {16, 192, 5}; cuuint32_t es[3] = {1, 1, 1};
    fn(&tm3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, qkv, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  make_tmap_2d_f16(&tm64, qkv, M, LD, LD, 192, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  make_tmap_2d_f16(&tm16, qkv, M, LD, LD, 192, 16, CU_TENSOR_MAP_SWIZZLE_32B);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      probe<<<148, 128, 100 * 1024>>>(tm3, tm64, tm16, mode, H, B * H, cyc);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> h(148); cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (auto v : h) avg += v; avg /= 148;
      printf("mode %d rep %d: %.1f us for all heads of one layer (serial per SM, no overlap), %.0f cycles per head, err=%s\n", mode, rep, ms * 1e3, avg / 6.92, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
