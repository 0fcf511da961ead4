// The whole one-token transformer decoder (pose_transformer.py:191-201, 349-357: 6 x [self-attention on one token,
// cross-attention over the 192 image tokens, feed-forward]) as ONE persistent kernel.
//
// Why: with a single query token per image every Linear of the decoder is a 64-row GEMM (bs = 64): 66 launches of
// 8-40 us each (~0.9 ms, 5 % of the forward) that move 48 MB of weights in total, i.e. ~8 us of HBM time.  They are
// pure launch / pipeline-fill latency.  Here 128 CTAs stay resident, split the OUTPUT FEATURES of each Linear between
// them (every weight byte is read exactly once, 8 output columns per CTA-tile) and meet at a grid barrier between the
// seven dependent steps of a layer:
//     1  y = LN0(tok)            v   = y Wv^T                       (softmax over one key = 1: attention output = v)
//     2                          tok += v Wo^T + bo
//     3  y = LN1(tok)            q   = y Wq^T
//     4  per (image, head) warp  att = softmax(q k^T / 8) v         (k, v rows of the batched to_kv GEMM)
//     5                          tok += att Wco^T + bco
//     6  y = LN2(tok)            hid = GELU(y W1^T + b1)
//     7                          tok += hid W2^T + b2
// Each GEMM step stages its 64 x K fp16 A operand in shared memory (LayerNorm recomputed per CTA from the fp32 token
// rows: 256 KB of L2 reads beat another barrier), splits K over the 8 warps, multiplies with mma.sync m16n8k16
// (fp16 operands, fp32 accumulate: the engine's numeric contract) and reduces the 8 partial tiles through shared
// memory.  These are weight-streaming products (64 rows: HBM / latency bound), not tensor-pipe work, so the legacy
// warp-level MMA is the right tool; the big GEMMs stay on tcgen05.
// Data exchanged between CTAs inside the kernel (tok, v, q, att, hid) is read with ld.global.cg (L1 is not coherent).
// The grid barrier is a monotonically increasing counter (zeroed by a memset node before the launch); every wait is
// bounded and raises the pipeline-timeout flag instead of hanging.  All CTAs are co-resident by construction
// (grid <= SM count, one CTA per SM, nothing else runs on the stream).
#pragma once
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm_tcgen05.cuh"   // gelu_erf, g_pipeline_timeout

namespace thmr {

constexpr int kDfRows = 64;          // token rows (images) per launch
constexpr int kDfThreads = 256;      // 8 warps
constexpr int kDfWarps = kDfThreads / 32;
constexpr int kDfMaxK = 1024;
constexpr int kDfPitch = kDfMaxK + 8;   // halves per A row in smem (+8: conflict-free fragment loads)
constexpr int kDfCtas = 128;
constexpr uint32_t kDfSmemA = kDfRows * kDfPitch * 2;                 // 132,096 B
constexpr uint32_t kDfSmemRed = kDfWarps * kDfRows * 8 * 4;           // 16 KB: per-warp 64 x 8 partial tiles
constexpr uint32_t kDfSmemGB = 2 * kDfMaxK * 4;                       // gamma, beta
constexpr uint32_t kDfSmemSc = kDfWarps * 192 * 4;                    // attention scores per warp
constexpr uint32_t kDfSmemTotal = kDfSmemA + kDfSmemRed + kDfSmemGB + kDfSmemSc + 128;

struct DecFusedLayer {
  const float *ln0_g, *ln0_b; const __half* sa_v_w; const __half* sa_out_w; const float* sa_out_b;
  const float *ln1_g, *ln1_b; const __half* ca_q_w; const __half* ca_out_w; const float* ca_out_b;
  const float *ln2_g, *ln2_b; const __half* ff1_w; const float* ff1_b; const __half* ff2_w; const float* ff2_b;
};
struct DecFusedParams {
  DecFusedLayer L[8];
  int depth, B, E, inner, mlp, heads, T;
  float eps, scale;
  const float* token0;    // [E] initial token (to_token_embedding.bias + pos_embedding)
  float* tok;             // [B, E] fp32 residual stream (output)
  __half* v16;            // [B, inner]
  float* q32;             // [B, inner]
  __half* att16;          // [B, inner]
  __half* hid16;          // [B, mlp]
  const __half* kv;       // [(b*T + j), kv_ld]: K of layer l at column l*2*inner, V at +inner
  int kv_ld;
  unsigned* barrier;      // zeroed before the launch
};

__device__ __forceinline__ void df_grid_sync(unsigned* bar, unsigned& target, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    uint32_t it = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++it > (1u << 22)) { atomicExch(&g_pipeline_timeout, 1u); break; }
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void df_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// A operand <- LayerNorm(tok) (fp32 [B, K] -> fp16 smem), same arithmetic as layernorm_reg_kernel (elementwise.cuh).
__device__ __forceinline__ void df_stage_ln(__half* As, float* sgb, const float* tok, const float* gamma,
                                            const float* beta, int B, int K, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* sg = reinterpret_cast<float4*>(sgb);
  float4* sb = sg + K / 4;
  for (int c = threadIdx.x; c < K / 4; c += kDfThreads) {
    sg[c] = __ldg(reinterpret_cast<const float4*>(gamma) + c);
    sb[c] = __ldg(reinterpret_cast<const float4*>(beta) + c);
  }
  __syncthreads();
  const int vec = K / 128;                    // float4 per lane (8 for K = 1024)
  auto load_row = [&](float4 (&v)[8], int r) {
    const float* xr = tok + static_cast<size_t>(r) * K;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = (i < vec && r < B) ? __ldcg(reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  // this warp's rows: warp, warp + 8, ...; the next row's loads are issued before this row's arithmetic
  float4 v[8], nx[8];
  load_row(v, warp);
#pragma unroll 1
  for (int r = warp; r < kDfRows; r += kDfWarps) {
    if (r + kDfWarps < kDfRows) load_row(nx, r + kDfWarps);
    __half* arow = As + static_cast<size_t>(r) * kDfPitch;
    if (r >= B) {
      for (int i = 0; i < vec; ++i) *reinterpret_cast<uint2*>(arow + (i * 32 + lane) * 4) = make_uint2(0u, 0u);
    } else {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
      const float mean = warp_sum(s) / K;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < vec) {
          const float a = v[i].x - mean, b = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
          q += a * a + b * b + d * d + e * e;
        }
      }
      const float rstd = rsqrtf(warp_sum(q) / K + eps);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < vec) {
          const int c = (i * 32 + lane) * 4;
          const float4 g = sg[c >> 2];
          const float4 bb = sb[c >> 2];
          __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
          __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(arow + c) = pk;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = nx[i];
  }
}

// A operand <- fp16 [B, K] produced by other CTAs in this kernel: cp.async.cg (L2 -> smem, no L1), every 16-byte
// piece in flight at once (a load-then-store loop exposes one L2 round trip per iteration).
__device__ __forceinline__ void df_stage_f16(__half* As, const __half* src, int B, int K) {
  const int per_row = K / 8;                  // 16-byte pieces per row
  for (int i = threadIdx.x; i < kDfRows * per_row; i += kDfThreads) {
    const int r = i / per_row, c = (i % per_row) * 8;
    __half* dst = As + static_cast<size_t>(r) * kDfPitch + c;
    if (r < B) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)),
                   "l"(src + static_cast<size_t>(r) * K + c)
                   : "memory");
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

enum { kDfOutF16 = 0, kDfOutF32 = 1, kDfOutAcc = 2 };

// out[B, N] (mode) = act(A[64, K] W[N, K]^T + bias); A staged by the caller's `stage` functor once the weight fragments
// of this CTA's first column tile are in flight.
template <typename Stage>
__device__ __forceinline__ void df_gemm(__half* As, float* red, Stage stage, const __half* __restrict__ W,
                                        const float* __restrict__ bias, int B, int K, int N, int act, int mode,
                                        void* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int kw = K / kDfWarps;                // k range of this warp
  const int ksteps = kw / 16;                 // 8 (K = 1024) or 4 (K = 512)
  bool staged = false;
  for (int nt = blockIdx.x; nt < N / 8; nt += gridDim.x) {
    // weight fragments of this warp's k range: b0 = W[n][k0 + 2t .. +1], b1 = W[n][k0 + 2t + 8 .. +9]
    uint32_t bf[8][2];
    const __half* wrow = W + static_cast<size_t>(nt * 8 + g) * K + warp * kw + 2 * t;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s < ksteps) {
        bf[s][0] = __ldg(reinterpret_cast<const uint32_t*>(wrow + s * 16));
        bf[s][1] = __ldg(reinterpret_cast<const uint32_t*>(wrow + s * 16 + 8));
      }
    }
    if (!staged) {
      stage();
      staged = true;
    }
    __syncthreads();        // A operand visible (and `red` of the previous tile fully consumed)
    float acc[4][4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][e] = 0.f;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s < ksteps) {
        const int k0 = warp * kw + s * 16 + 2 * t;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          const __half* a0 = As + static_cast<size_t>(mi * 16 + g) * kDfPitch + k0;
          uint32_t af[4];
          af[0] = *reinterpret_cast<const uint32_t*>(a0);
          af[1] = *reinterpret_cast<const uint32_t*>(a0 + 8 * kDfPitch);
          af[2] = *reinterpret_cast<const uint32_t*>(a0 + 8);
          af[3] = *reinterpret_cast<const uint32_t*>(a0 + 8 * kDfPitch + 8);
          df_mma(acc[mi], af, bf[s]);
        }
      }
    }
    // partial tile of this warp: rows mi*16 + g (+8), columns 2t, 2t+1
    float* rw = red + static_cast<size_t>(warp) * kDfRows * 8;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      *reinterpret_cast<float2*>(rw + (mi * 16 + g) * 8 + 2 * t) = make_float2(acc[mi][0], acc[mi][1]);
      *reinterpret_cast<float2*>(rw + (mi * 16 + g + 8) * 8 + 2 * t) = make_float2(acc[mi][2], acc[mi][3]);
    }
    __syncthreads();
    // 64 x 8 outputs, two per thread: sum the 8 partials in warp order (deterministic), bias, activation, store
    const int row = threadIdx.x >> 2, c2 = (threadIdx.x & 3) * 2;
    float2 sum = make_float2(0.f, 0.f);
#pragma unroll
    for (int w = 0; w < kDfWarps; ++w) {
      const float2 pr = *reinterpret_cast<const float2*>(red + (static_cast<size_t>(w) * kDfRows + row) * 8 + c2);
      sum.x += pr.x; sum.y += pr.y;
    }
    if (row < B) {
      const int col = nt * 8 + c2;
      if (bias) { sum.x += __ldg(bias + col); sum.y += __ldg(bias + col + 1); }
      if (mode == kDfOutAcc) {
        float2* o = reinterpret_cast<float2*>(static_cast<float*>(out) + static_cast<size_t>(row) * N + col);
        const float2 old = __ldcg(o);
        *o = make_float2(old.x + sum.x, old.y + sum.y);
      } else if (mode == kDfOutF32) {
        *reinterpret_cast<float2*>(static_cast<float*>(out) + static_cast<size_t>(row) * N + col) = sum;
      } else {
        if (act == kActGelu) { sum.x = gelu_erf(sum.x); sum.y = gelu_erf(sum.y); }
        *reinterpret_cast<__half2*>(static_cast<__half*>(out) + static_cast<size_t>(row) * N + col) =
            __floats2half2_rn(sum.x, sum.y);
      }
    }
  }
}

// One-query cross-attention (pose_transformer.py:111-124), one warp per (image, head), dim_head = 64, T keys (T % 4 == 0,
// T <= 192).  8 lanes share a key / value row (16 bytes each), 4 rows per step.
__device__ __forceinline__ void df_cross_attn(float* sc_all, const DecFusedParams& p, int layer) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane & 7, grp = lane >> 3;
  float* sc = sc_all + warp * 192;
  const int koff = layer * 2 * p.inner, voff = koff + p.inner;
  for (int pr = blockIdx.x * kDfWarps + warp; pr < p.B * p.heads; pr += gridDim.x * kDfWarps) {
    const int b = pr / p.heads, h = pr % p.heads;
    const float4 q0 = __ldcg(reinterpret_cast<const float4*>(p.q32 + static_cast<size_t>(b) * p.inner + h * 64 + sub * 8));
    const float4 q1 = __ldcg(reinterpret_cast<const float4*>(p.q32 + static_cast<size_t>(b) * p.inner + h * 64 + sub * 8 + 4));
    const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    const __half* base = p.kv + static_cast<size_t>(b) * p.T * p.kv_ld + h * 64 + sub * 8;
    // scores: 8 rows in flight per lane (one L2 / HBM round trip per batch instead of one per row)
    for (int it0 = 0; it0 < p.T / 4; it0 += 8) {
      uint4 pk[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = (it0 + u) * 4 + grp;
        pk[u] = (it0 + u < p.T / 4) ? __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(j) * p.kv_ld + koff))
                                    : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&pk[u]);
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          s += qv[2 * e] * f.x + qv[2 * e + 1] * f.y;
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const int j = (it0 + u) * 4 + grp;
        if (sub == 0 && it0 + u < p.T / 4) sc[j] = s * p.scale;
      }
    }
    __syncwarp();
    float m = -INFINITY;
    for (int j = lane; j < p.T; j += 32) m = fmaxf(m, sc[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < p.T; j += 32) {
      const float e = expf(sc[j] - m);
      sc[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int it0 = 0; it0 < p.T / 4; it0 += 8) {
      uint4 pk[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = (it0 + u) * 4 + grp;
        pk[u] = (it0 + u < p.T / 4) ? __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(j) * p.kv_ld + voff))
                                    : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (it0 + u < p.T / 4) {
          const float pj = sc[(it0 + u) * 4 + grp] / sum;
          const __half2* h2 = reinterpret_cast<const __half2*>(&pk[u]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h2[e]);
            o[2 * e] += pj * f.x;
            o[2 * e + 1] += pj * f.y;
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
      o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
    }
    if (grp == 0) {
      uint4 pk;
      __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
      __half2 h2 = __floats2half2_rn(o[4], o[5]), h3 = __floats2half2_rn(o[6], o[7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(p.att16 + static_cast<size_t>(b) * p.inner + h * 64 + sub * 8) = pk;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kDfThreads, 1) dec_stack_kernel(const DecFusedParams p) {
  extern __shared__ uint8_t df_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(df_smem_raw) + 127) & ~uintptr_t(127));
  __half* As = reinterpret_cast<__half*>(smem);
  float* red = reinterpret_cast<float*>(smem + kDfSmemA);
  float* sgb = reinterpret_cast<float*>(smem + kDfSmemA + kDfSmemRed);
  float* sc = reinterpret_cast<float*>(smem + kDfSmemA + kDfSmemRed + kDfSmemGB);
  unsigned target = 0;
  const unsigned nb = gridDim.x;
  // token init: every image starts from the same learned token
  for (int i = blockIdx.x * kDfThreads + threadIdx.x; i < p.B * p.E; i += nb * kDfThreads) p.tok[i] = __ldg(p.token0 + i % p.E);
  df_grid_sync(p.barrier, target, nb);
  for (int l = 0; l < p.depth; ++l) {
    const DecFusedLayer& w = p.L[l];
    // 1: v = LN0(tok) Wv^T
    df_gemm(As, red, [&] { df_stage_ln(As, sgb, p.tok, w.ln0_g, w.ln0_b, p.B, p.E, p.eps); }, w.sa_v_w, nullptr, p.B, p.E,
            p.inner, kActNone, kDfOutF16, p.v16);
    df_grid_sync(p.barrier, target, nb);
    // 2: tok += v Wo^T + bo
    df_gemm(As, red, [&] { df_stage_f16(As, p.v16, p.B, p.inner); }, w.sa_out_w, w.sa_out_b, p.B, p.inner, p.E, kActNone,
            kDfOutAcc, p.tok);
    df_grid_sync(p.barrier, target, nb);
    // 3: q = LN1(tok) Wq^T
    df_gemm(As, red, [&] { df_stage_ln(As, sgb, p.tok, w.ln1_g, w.ln1_b, p.B, p.E, p.eps); }, w.ca_q_w, nullptr, p.B, p.E,
            p.inner, kActNone, kDfOutF32, p.q32);
    df_grid_sync(p.barrier, target, nb);
    // 4: cross-attention
    df_cross_attn(sc, p, l);
    df_grid_sync(p.barrier, target, nb);
    // 5: tok += att Wco^T + bco
    df_gemm(As, red, [&] { df_stage_f16(As, p.att16, p.B, p.inner); }, w.ca_out_w, w.ca_out_b, p.B, p.inner, p.E, kActNone,
            kDfOutAcc, p.tok);
    df_grid_sync(p.barrier, target, nb);
    // 6: hid = GELU(LN2(tok) W1^T + b1)
    df_gemm(As, red, [&] { df_stage_ln(As, sgb, p.tok, w.ln2_g, w.ln2_b, p.B, p.E, p.eps); }, w.ff1_w, w.ff1_b, p.B, p.E,
            p.mlp, kActGelu, kDfOutF16, p.hid16);
    df_grid_sync(p.barrier, target, nb);
    // 7: tok += hid W2^T + b2
    df_gemm(As, red, [&] { df_stage_f16(As, p.hid16, p.B, p.mlp); }, w.ff2_w, w.ff2_b, p.B, p.mlp, p.E, kActNone, kDfOutAcc,
            p.tok);
    df_grid_sync(p.barrier, target, nb);
  }
}

inline bool dec_fused_supported(int B, int E, int inner, int mlp, int heads, int dim_head, int T, int depth) {
  return B >= 1 && B <= kDfRows && E % 128 == 0 && E <= kDfMaxK && inner % 128 == 0 && inner <= kDfMaxK &&
         mlp % 128 == 0 && mlp <= kDfMaxK && dim_head == 64 && heads * 64 == inner && T % 4 == 0 && T <= 192 &&
         depth <= 8;
}

inline int dec_fused_launch(const DecFusedParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    THMR_CUDA(cudaFuncSetAttribute(dec_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDfSmemTotal));
    configured = true;
  }
  THMR_CUDA(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), st));
  const int grid = num_sms() < kDfCtas ? num_sms() : kDfCtas;
  dec_stack_kernel<<<grid, kDfThreads, kDfSmemTotal, st>>>(p);
  THMR_CUDA(cudaGetLastError());
  return THMR_OK;
}

}  // namespace thmr
