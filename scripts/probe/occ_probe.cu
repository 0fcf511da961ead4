// Does using tcgen05.alloc cap occupancy at one CTA per SM?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __maxnreg__(80) k_plain(float* out) {
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) slot = 1;
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = slot;
}
template <int COLS>
__global__ void __maxnreg__(80) k_tmem(float* out) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(&slot));
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = slot;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(COLS) : "memory");
}
int main() {
  int n;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_plain, 320, 0); printf("plain 320thr 0B: %d\n", n);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tmem<256>, 320, 0); printf("tmem256 320thr 0B: %d\n", n);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tmem<128>, 320, 0); printf("tmem128 320thr 0B: %d\n", n);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tmem<32>, 320, 0); printf("tmem32 320thr 0B: %d\n", n);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tmem<256>, 256, 0); printf("tmem256 256thr 0B: %d\n", n);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tmem<256>, 128, 0); printf("tmem256 128thr 0B: %d\n", n);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_tmem<256>); printf("regs %d\n", fa.numRegs);
  return 0;
}
