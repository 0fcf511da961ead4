// Single translation unit of libtokenhmr_b200.so: device kernels (*.cuh) + the extern "C" ABI
// declared in include/tokenhmr_b200.h.
#include "common.cuh"
#include "gemm_host.cuh"

using namespace thmr;

extern "C" {

int thmr_abi_version(void) { return 1; }

const char* thmr_last_error(void) { return last_error_buf(); }

int thmr_check_device_flags(void) {
  THMR_CUDA(cudaDeviceSynchronize());
  unsigned int flag = 0;
  THMR_CUDA(cudaMemcpyFromSymbol(&flag, g_pipeline_timeout, sizeof(flag)));
  if (flag) {
    unsigned int zero = 0;
    THMR_CUDA(cudaMemcpyToSymbol(g_pipeline_timeout, &zero, sizeof(zero)));
    return fail(THMR_ERR_TIMEOUT, "device pipeline wait timed out (mbarrier never completed)");
  }
  return THMR_OK;
}

int thmr_gemm_f16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                  const float* resid, int ldr, int act, float* out32, int ld32, void* out16, int ld16, int block_n,
                  void* stream) {
  THMR_CHECK(A && B, "gemm: null operand");
  GemmDesc d;
  d.A = static_cast<const __half*>(A); d.lda = lda; d.a_rows = M;
  d.B = static_cast<const __half*>(B); d.ldb = ldb;
  d.M = M; d.N = N; d.K = K;
  d.bias = bias; d.resid = resid; d.ldr = ldr; d.act = act;
  d.out32 = out32; d.ld32 = ld32; d.out16 = static_cast<__half*>(out16); d.ld16 = ld16;
  d.force_bn = block_n;
  GemmPlan plan;
  THMR_TRY(gemm_make_plan(d, &plan));
  return gemm_launch(plan, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
