// Tokenizer encoder + hard quantisation: EncodeTokens.forward
// (tokenization/models/vanilla_pose_vqvae.py:304-346) = PoseSPEncoderV1.encoder (:42-111) followed by
// QuantizeEMAReset.preprocess / quantize (tokenization/models/quantize_cnn.py:74-86).
//
//   pose6d (B,21,6) -> Conv1d(6,W,3)+ReLU -> Upsample(40) Conv+ReLU -> [Upsample(x2) Conv+ReLU] x (mul-1)
//   -> Conv1d(W,W,4,stride 2,pad 1) -> Resnet1D(depth, dilation rate^d, reversed) -> Conv1d(W,code_dim,3)
//   -> (B*T, code_dim) -> argmin_k |x - c_k|^2 -> code_idx (B*T,) int64        (T = 20 * 2^mul / 2 = 160)
//
// Every convolution runs on the tcgen05 GEMM of the forward path as an implicit GEMM over channels-last, zero-padded
// sequences [B, L + 2*kEncPad, C] (the tokenizer decoder's layout, engine.cuh).  Two layers need a gather first:
//   * the 6 input channels are zero-padded to 64 (one 128-byte TMA row per tap);
//   * the stride-2, 4-tap down-sampling conv reads rows 2j-1 .. 2j+2: those four rows are gathered side by side
//     into one [B*(Lout+2 pad), 4W] operand (tap-major, like the repacked weight) and the conv becomes a plain GEMM.
// The quantiser is the split-fp16 distance GEMM with the running arg-min in its epilogue (thmr_vq_argmin).
#pragma once
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm_host.cuh"

struct thmr_tok_encoder {
  thmr_tok_encoder_desc d;
};

namespace thmr {

constexpr int kEncPad = 3;      // zero rows either side of every sequence (max dilation of the release tokenizer)
constexpr int kEncCin0 = 64;    // the 6 input channels padded to one 64-wide k-block

// pose6d fp32 [B, J, in_dim] -> fp16 [B, J + 2 pad, 64], channels >= in_dim and pad rows zero.
__global__ void enc_input_kernel(const float* __restrict__ pose, __half* __restrict__ dst, int B, int J, int in_dim,
                                 int pad) {
  const long t = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long total = static_cast<long>(B) * (J + 2 * pad) * kEncCin0;
  if (t >= total) return;
  const int c = t % kEncCin0;
  const long row = t / kEncCin0;
  const int r = row % (J + 2 * pad);
  const int b = row / (J + 2 * pad);
  float v = 0.f;
  if (r >= pad && r < pad + J && c < in_dim) v = pose[(static_cast<long>(b) * J + (r - pad)) * in_dim + c];
  dst[t] = __float2half_rn(v);
}

// Operand of Conv1d(C, C, 4, stride 2, padding 1): dst[b, pad + j, tap*C + c] = src[b, pad + 2j - 1 + tap, c]
// (rows -1 and Lin are pad rows of src, i.e. zero).  Pad rows of dst are zeroed.  C8 = C / 8.
__global__ void enc_gather_s2_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int B, int Lin, int Lout,
                                     int pad, int C8) {
  const long total = static_cast<long>(B) * (Lout + 2 * pad) * 4 * C8;
  const long t = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (t >= total) return;
  const int c8 = t % C8;
  const int tap = (t / C8) % 4;
  const long row = t / (4L * C8);
  const int r = row % (Lout + 2 * pad);
  const int b = row / (Lout + 2 * pad);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (r >= pad && r < pad + Lout) {
    const int sr = pad + 2 * (r - pad) - 1 + tap;          // in [pad - 1, pad + Lin]: always inside the padded row range
    v = reinterpret_cast<const uint4*>(src)[(static_cast<long>(b) * (Lin + 2 * pad) + sr) * C8 + c8];
  }
  reinterpret_cast<uint4*>(dst)[t] = v;
}

struct EncWs {
  __half *in0, *bufA, *bufB, *gat;
  float *x32, *lat_pad, *lat;
  void* vq;
  size_t total;
};
inline int enc_seq_lens(const thmr_tok_encoder_desc& d, int* L_up_max, int* T) {
  int L = ((d.joints * 2) / 10) * 10;                       // vanilla_pose_vqvae.py:69
  for (int i = 1; i < d.size_mul; ++i) L *= 2;
  *L_up_max = L;
  *T = (L + 2 - 4) / 2 + 1;                                 // Conv1d(k=4, s=2, p=1)
  return THMR_OK;
}
inline void enc_carve(const thmr_tok_encoder_desc& d, void* base, int B, EncWs* ws) {
  Bump bp(base);
  int Lmax, T;
  enc_seq_lens(d, &Lmax, &T);
  const size_t rows_max = static_cast<size_t>(B) * (Lmax + 2 * kEncPad);
  const size_t rows_T = static_cast<size_t>(B) * (T + 2 * kEncPad);
  ws->in0 = bp.take<__half>(static_cast<size_t>(B) * (d.joints + 2 * kEncPad) * kEncCin0);
  ws->bufA = bp.take<__half>(rows_max * d.width);
  ws->bufB = bp.take<__half>(rows_max * d.width);
  ws->gat = bp.take<__half>(rows_T * 4 * d.width);
  ws->x32 = bp.take<float>(rows_T * d.width);
  ws->lat_pad = bp.take<float>(rows_T * d.code_dim);
  ws->lat = bp.take<float>(static_cast<size_t>(B) * T * d.code_dim);
  const size_t vq = thmr_vq_workspace_bytes(static_cast<int64_t>(B) * T, d.nb_code, d.code_dim);
  ws->vq = bp.take<uint8_t>(vq);
  ws->total = (bp.off + 1023) & ~size_t(1023);
}

inline int enc_conv(const __half* in, int B, int Lcur, int cin, int taps, int dil, const thmr_tok_conv& cw, int cout, int act,
                    float* o32, __half* o16, const float* resid, cudaStream_t st) {
  const int Lp = Lcur + 2 * kEncPad;
  GemmDesc d;
  d.A = in; d.lda = cin; d.a_rows = static_cast<long long>(B) * Lp;
  d.B = static_cast<const __half*>(cw.w); d.ldb = taps * cin;
  d.M = B * Lp; d.N = cout; d.K = taps * cin;
  d.bias = cw.b; d.act = act;
  d.resid = resid; d.ldr = cout;
  d.out32 = o32; d.ld32 = cout; d.out16 = o16; d.ld16 = cout;
  if (taps > 1) { d.taps = taps; d.cin = cin; d.tap_row0 = -dil; d.tap_stride = dil; }
  d.seq_pitch = Lp; d.seq_lo = kEncPad; d.seq_hi = kEncPad + Lcur;
  GemmPlan plan;
  THMR_TRY(gemm_make_plan(d, &plan));
  return gemm_launch(plan, st);
}

inline int enc_run(const thmr_tok_encoder* e, const float* pose6d, int B, int64_t* code_idx, float* latent, void* workspace,
                   cudaStream_t st) {
  const thmr_tok_encoder_desc& d = e->d;
  EncWs ws;
  enc_carve(d, workspace, B, &ws);
  const int W = d.width, PAD = kEncPad;
  auto blocks = [](long n) { return static_cast<unsigned>((n + 255) / 256); };
  int L = d.joints;
  enc_input_kernel<<<blocks(static_cast<long>(B) * (L + 2 * PAD) * kEncCin0), 256, 0, st>>>(pose6d, ws.in0, B, L, d.in_dim,
                                                                                           PAD);
  THMR_CUDA(cudaGetLastError());
  THMR_TRY(enc_conv(ws.in0, B, L, kEncCin0, 3, 1, d.conv_in, W, kActRelu, nullptr, ws.bufA, nullptr, st));
  int Lout = ((d.joints * 2) / 10) * 10;
  for (int u = 0; u < d.size_mul; ++u) {                    // Upsample -> Conv1d(W,W,3) -> ReLU
    upsample_rows_kernel<<<blocks(static_cast<long>(B) * (Lout + 2 * PAD) * (W / 8)), 256, 0, st>>>(ws.bufA, ws.bufB, B, L,
                                                                                                    Lout, PAD, W / 8);
    THMR_CUDA(cudaGetLastError());
    L = Lout;
    THMR_TRY(enc_conv(ws.bufB, B, L, W, 3, 1, d.conv_up[u], W, kActRelu, nullptr, ws.bufA, nullptr, st));
    Lout = 2 * L;
  }
  // down-sampling conv: gather the four taps, then one GEMM; fp32 output = residual stream of the Resnet1D,
  // fp16 output = ReLU of it (ResConv1DBlock applies its activation first, resnet.py:51-60)
  const int T = (L + 2 - 4) / 2 + 1;
  enc_gather_s2_kernel<<<blocks(static_cast<long>(B) * (T + 2 * PAD) * 4 * (W / 8)), 256, 0, st>>>(ws.bufA, ws.gat, B, L, T,
                                                                                                  PAD, W / 8);
  THMR_CUDA(cudaGetLastError());
  THMR_TRY(enc_conv(ws.gat, B, T, 4 * W, 1, 1, d.conv_down, W, kActRelu, ws.x32, ws.bufA, nullptr, st));
  for (int dd = 0; dd < d.depth; ++dd) {                    // stored order = dilation descending (reverse_dilation)
    int dil = 1;
    for (int k = 0; k < d.depth - 1 - dd; ++k) dil *= d.dilation_rate;
    THMR_CHECK(dil <= PAD, "tok_encoder: dilation %d exceeds the sequence padding %d", dil, PAD);
    THMR_TRY(enc_conv(ws.bufA, B, T, W, 3, dil, d.res_conv1[dd], W, kActRelu, nullptr, ws.bufB, nullptr, st));
    const bool last = dd == d.depth - 1;
    THMR_TRY(enc_conv(ws.bufB, B, T, W, 1, 1, d.res_conv2[dd], W, last ? kActNone : kActRelu, ws.x32, ws.bufA, ws.x32, st));
  }
  THMR_TRY(enc_conv(ws.bufA, B, T, W, 3, 1, d.conv_out, d.code_dim, kActNone, ws.lat_pad, nullptr, nullptr, st));
  // QuantizeEMAReset.preprocess: (B, C, T) -> (B*T, C): drop the pad rows
  float* lat = latent ? latent : ws.lat;
  const size_t row_bytes = sizeof(float) * d.code_dim;
  THMR_CUDA(cudaMemcpy2DAsync(lat, row_bytes * T, ws.lat_pad + static_cast<size_t>(PAD) * d.code_dim,
                              row_bytes * (T + 2 * PAD), row_bytes * T, B, cudaMemcpyDeviceToDevice, st));
  return thmr_vq_argmin(lat, static_cast<int64_t>(B) * T, d.codebook, d.nb_code, d.code_dim, code_idx, ws.vq, st);
}

}  // namespace thmr
