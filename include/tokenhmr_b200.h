/*
 * tokenhmr_b200 — C ABI of the B200-native TokenHMR inference engine.
 *
 * The reference (saidwivedi/TokenHMR @ 198645f) has no FFI layer: its seam is the Python method
 * TokenHMR.forward(batch) (tokenhmr/lib/models/tokenhmr.py:330-338 -> forward_step :135-188) and the
 * sub-module calls inside it.  Every entry point below names the reference call it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - functions return THMR_OK (0) or a negative thmr_status; thmr_last_error() holds the message
 *     (thread-local); nothing throws across this boundary;
 *   - the caller owns every buffer it passes (inputs, outputs, workspace); the engine owns only the
 *     repacked weights it creates in thmr_engine_create;
 *   - calls on one engine must be serialised by the caller; kernels are stream-ordered and
 *     CUDA-graph capturable (no host synchronisation inside thmr_engine_forward);
 *   - "f16" buffers hold IEEE binary16 (__half).
 */
#ifndef TOKENHMR_B200_H_
#define TOKENHMR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum thmr_status {
  THMR_OK = 0,
  THMR_ERR_INVALID = -1, /* bad argument */
  THMR_ERR_CUDA = -2,    /* CUDA runtime / driver error */
  THMR_ERR_TIMEOUT = -3, /* a device-side pipeline wait expired (kernel bug, not a data error) */
  THMR_ERR_NOMEM = -4
} thmr_status;

/* ABI version (bumped on any signature change) and last error message of the calling thread. */
int thmr_abi_version(void);
const char* thmr_last_error(void);
/* Reads and clears the device-side pipeline-timeout flag (synchronises the device). */
int thmr_check_device_flags(void);

/* ------------------------------------------------------------------------------------------------
 * Standalone operators (each is also a stage of thmr_engine_forward)
 * ---------------------------------------------------------------------------------------------- */

enum { THMR_ACT_NONE = 0, THMR_ACT_GELU = 1, THMR_ACT_RELU = 2 };

/* nn.Linear as one tcgen05 GEMM:  y = x @ W^T (+ bias) (+ resid)   [vit.py:82-86,112,123; every F.linear on the path]
 *   A [M,K] f16 row-major (lda), B = weight [N,K] f16 row-major (ldb), fp32 accumulate.
 *   out32 (nullable) receives acc+bias+resid in fp32; out16 (nullable) receives act(acc+bias+resid) in f16.
 *   resid (nullable, fp32 [M,N], pitch ldr) may alias out32.  block_n: 0 = auto, else 32/64/128/256. */
int thmr_gemm_f16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                  const float* resid, int ldr, int act, float* out32, int ld32, void* out16, int ld16, int block_n,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOKENHMR_B200_H_ */
