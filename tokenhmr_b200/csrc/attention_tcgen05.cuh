// Fused ViT multi-head attention core on tcgen05 (reference: Attention.forward, vit.py:116-122).
//   per (image b, head h):  S = Q K^T (fp32 in TMEM) -> softmax rows (fp32, scale folded into exp2)
//                           -> P (fp16, 128B-swizzled smem) -> O = P V (fp32 in TMEM) -> O / rowsum -> fp16
// N = 192 tokens, head_dim = 80, the whole K/V of one head is resident in shared memory, so the softmax
// is a plain single-block softmax (no online rescaling).  Q/K/V are read straight out of the QKV GEMM's
// [B*192, 3*H*80] fp16 output with 3-D TMA boxes (16 dims x 192 tokens x 5 chunks, 32B swizzle).
//
// 192 query rows = two M=128 UMMA tiles: rows 0..127 and rows 64..191; the second tile's first 64 rows are
// duplicates and are skipped by the softmax/epilogue warps (only their MMA cycles are spent).
//
//   warp 0      TMA producer (Q,K double-buffered across heads; V single-buffered)
//   warp 1      MMA issuer, TMEM owner (512 columns: S tile t at columns [192 t, 192 t + 192), O_t aliases S_t)
//   warps 2..9  softmax + epilogue: thread = one query row x 96 of the 192 key columns
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace thmr {

constexpr int kAttTokens = 192;
constexpr int kAttHeadDim = 80;
constexpr int kAttChunks = kAttHeadDim / 16;                     // 16-element (32 B) K-chunks
constexpr int kAttThreads = 320;
constexpr uint32_t kAttChunkBytes = kAttTokens * 32;             // 6144
constexpr uint32_t kAttMatBytes = kAttChunks * kAttChunkBytes;   // 30720 (one of Q / K / V)
constexpr uint32_t kAttPBytes = 3 * 128 * 128;                   // 49152: P tile [128 x 192] fp16, SW128 atoms
constexpr uint32_t kAttOffQK = kAttPBytes;
constexpr uint32_t kAttOffV = kAttOffQK + 4 * kAttMatBytes;
constexpr uint32_t kAttOffStats = kAttOffV + kAttMatBytes;       // smax[2][2][128], ssum[2][2][128]
constexpr uint32_t kAttOffBars = kAttOffStats + 2 * 2 * 2 * 128 * 4;
constexpr uint32_t kAttSmemBytes = kAttOffBars + 16 * 8 + 16 + 1024;

struct AttnParams {
  int num_problems;  // B * H
  int heads;
  float scale_log2e; // head_dim^-0.5 * log2(e)
  __half* out;       // [B*192, ldo], head h at columns [80 h, 80 h + 80)
  int ldo;
  float* dbg_s;      // optional [B*H, 192, 192] raw scores (tests only)
  unsigned long long* dbg_counters;  // optional [gridDim.x][16] cycle counters (THMR_ATTN_COUNTERS)
  int p_in_tmem;     // experiment (THMR_ATTN_TS=1): P also written to TMEM cols [384,480) and PV issued in TS mode
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kAttThreads, 1)
vit_attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;
  uint8_t* sV = smem + kAttOffV;
  float* smax = reinterpret_cast<float*>(smem + kAttOffStats);          // [tile][half][128]
  float* ssum = smax + 2 * 2 * 128;                                      // [tile][half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAttOffBars);
  uint64_t* qk_full = bars;        // [2]
  uint64_t* qk_empty = bars + 2;   // [2]
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;
  uint64_t* s_full = bars + 6;     // [2]
  uint64_t* p_full = bars + 8;
  uint64_t* p_empty = bars + 9;
  uint64_t* o_full = bars + 10;    // [2]
  uint64_t* o_empty = bars + 12;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&qk_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 8);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      int i = 0;
      for (int prob = blockIdx.x; prob < p.num_problems; prob += gridDim.x, ++i) {
        const int b = prob / p.heads, h = prob % p.heads;
        const int buf = i & 1;
        uint8_t* sQ = smem + kAttOffQK + buf * 2 * kAttMatBytes;
        uint8_t* sK = sQ + kAttMatBytes;
        mbar_wait(&qk_empty[buf], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qk_full[buf], 2 * kAttMatBytes);
        tma_load_3d(sQ, &tmQKV, &qk_full[buf], 0, b * kAttTokens, h * kAttChunks);
        tma_load_3d(sK, &tmQKV, &qk_full[buf], 0, b * kAttTokens, (p.heads + h) * kAttChunks);
        mbar_wait(v_empty, (i & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, kAttMatBytes);
        tma_load_3d(sV, &tmQKV, v_full, 0, b * kAttTokens, (2 * p.heads + h) * kAttChunks);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = make_idesc_f16(128, kAttTokens);            // Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_f16(128, kAttHeadDim, 0, 1);     // P V   : V is N-major (d contiguous)
      int i = 0;
      long long w_qk = 0, w_oe = 0, w_v = 0, w_p = 0;
      const long long t_begin = clock64();
      for (int prob = blockIdx.x; prob < p.num_problems; prob += gridDim.x, ++i) {
        const int buf = i & 1;
        const uint32_t sQ = smem_u32(smem + kAttOffQK + buf * 2 * kAttMatBytes);
        const uint32_t sK = sQ + kAttMatBytes;
        long long t0 = clock64();
        mbar_wait(&qk_full[buf], (i >> 1) & 1);
        w_qk += clock64() - t0;
        for (int t = 0; t < 2; ++t) {
          t0 = clock64();
          mbar_wait(&o_empty[t], (i & 1) ^ 1);   // epilogue of the previous head has drained O_t (aliases S_t)
          w_oe += clock64() - t0;
          tc_fence_after();
#pragma unroll
          for (int kc = 0; kc < kAttChunks; ++kc) {
            const uint64_t da = make_smem_desc(sQ + kc * kAttChunkBytes + t * 64 * 32, 16, 256, kSwz32);
            const uint64_t db = make_smem_desc(sK + kc * kAttChunkBytes, 16, 256, kSwz32);
            umma_f16_ss(tmem_base + t * kAttTokens, da, db, idesc_s, kc != 0);
          }
          umma_commit(&s_full[t]);
        }
        umma_commit(&qk_empty[buf]);
        t0 = clock64();
        mbar_wait(v_full, i & 1);
        w_v += clock64() - t0;
        const uint32_t sPa = smem_u32(sP), sVa = smem_u32(sV);
        for (int t = 0; t < 2; ++t) {
          t0 = clock64();
          mbar_wait(p_full, t);                  // completion #(2i + t)
          w_p += clock64() - t0;
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < kAttTokens / 16; ++ks) {
            const uint64_t da = make_smem_desc(sPa + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, kSwz128);
            const uint64_t db = make_smem_desc(sVa + ks * 512, kAttChunkBytes, 256, kSwz32);
            if (p.p_in_tmem) umma_f16_ts(tmem_base + t * kAttTokens, tmem_base + 384 + ks * 8, db, idesc_o, ks != 0);
            else umma_f16_ss(tmem_base + t * kAttTokens, da, db, idesc_o, ks != 0);
          }
          umma_commit(&o_full[t]);
          umma_commit(p_empty);
        }
        umma_commit(v_empty);
      }
      if (p.dbg_counters) {
        unsigned long long* c = p.dbg_counters + blockIdx.x * 16;
        c[0] = w_qk; c[1] = w_oe; c[2] = w_v; c[3] = w_p; c[4] = clock64() - t_begin;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warps
    const int q = warp & 3;              // TMEM lane quarter
    const int half = (warp - 2) >> 2;    // which 96 of the 192 key columns
    const int trow = q * 32 + lane;      // row inside the 128-row tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    int i = 0;
    long long w_s = 0, w_pe = 0, w_of = 0, w_bar = 0, w_ld = 0, w_max = 0, w_exp = 0, w_sts = 0, w_epi = 0;
    const long long t_begin = clock64();
    for (int prob = blockIdx.x; prob < p.num_problems; prob += gridDim.x, ++i) {
      const int b = prob / p.heads, h = prob % p.heads;
      for (int t = 0; t < 2; ++t) {
        const bool active = (t == 0) || (q >= 2);
        long long t0 = clock64();
        mbar_wait(&s_full[t], i & 1);
        w_s += clock64() - t0;
        tc_fence_after();
        uint32_t pk[48];
        float sum = 0.f;
        if (active) {
          float s[96];
          long long tph = clock64();
          {
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              tmem_ld_x32(tmem_base + lane_addr + t * kAttTokens + half * 96 + j * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) s[j * 32 + e] = __uint_as_float(v[e]);
            }
          }
          if (p.dbg_s) {
            float* d = p.dbg_s + (static_cast<size_t>(prob) * kAttTokens + t * 64 + trow) * kAttTokens + half * 96;
#pragma unroll
            for (int e = 0; e < 96; ++e) d[e] = s[e];
          }
          { const long long tn = clock64(); w_ld += tn - tph; tph = tn; }
          float m = s[0];
#pragma unroll
          for (int e = 1; e < 96; ++e) m = fmaxf(m, s[e]);
          smax[(t * 2 + half) * 128 + trow] = m;
          t0 = clock64();
          named_bar_sync(1 + q, 64);
          w_bar += clock64() - t0;
          m = fmaxf(m, smax[(t * 2 + (half ^ 1)) * 128 + trow]);
          { const long long tn = clock64(); w_max += tn - tph; tph = tn; }
          const float mo = m * p.scale_log2e;
#pragma unroll
          for (int e = 0; e < 96; e += 2) {
            const float e0 = fast_exp2(fmaf(s[e], p.scale_log2e, -mo));
            const float e1 = fast_exp2(fmaf(s[e + 1], p.scale_log2e, -mo));
            sum += e0 + e1;
            __half2 h2 = __floats2half2_rn(e0, e1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h2);
          }
          ssum[(t * 2 + half) * 128 + trow] = sum;
          { const long long tn = clock64(); w_exp += tn - tph; tph = tn; }
        }
        t0 = clock64();
        mbar_wait(p_empty, t ^ 1);       // PV of the previous tile has finished reading P
        w_pe += clock64() - t0;
        if (active) {
          const uint32_t rbase = smem_u32(sP) + (trow >> 3) * 1024 + (trow & 7) * 128;
#pragma unroll
          for (int c = 0; c < 12; ++c) {
            const int col = half * 96 + c * 8;
            const uint32_t addr = rbase + (col >> 6) * 16384 + ((((col & 63) >> 3) ^ (trow & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                         "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                         : "memory");
          }
        }
        if (p.p_in_tmem) {
          // all 32 lanes participate (.sync.aligned); rows of inactive warps carry stale registers (results unused)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            uint32_t w16[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) w16[e] = pk[c * 16 + e];
            tmem_st_x16(tmem_base + lane_addr + 384 + half * 48 + c * 16, w16);
          }
          tmem_st_wait();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (active) w_sts += clock64() - t0 ;
      }
      for (int t = 0; t < 2; ++t) {
        const bool active = (t == 0) || (q >= 2);
        const long long t0 = clock64();
        mbar_wait(&o_full[t], i & 1);
        w_of += clock64() - t0;
        tc_fence_after();
        if (active) {
          const float inv = 1.0f / (ssum[(t * 2) * 128 + trow] + ssum[(t * 2 + 1) * 128 + trow]);
          const int row = b * kAttTokens + t * 64 + trow;
          const int c0 = half ? 48 : 0;
          const int nchunk = half ? 2 : 3;   // 16-column chunks: cols [0,48) | [48,80)
          __half* o = p.out + static_cast<size_t>(row) * p.ldo + h * kAttHeadDim + c0;
          for (int j = 0; j < nchunk; ++j) {
            uint32_t v[16];
            tmem_ld_x16(tmem_base + lane_addr + t * kAttTokens + c0 + j * 16, v);
            tmem_ld_wait();
            uint4 w0, w1;
            __half2 hh[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              hh[e] = __floats2half2_rn(__uint_as_float(v[2 * e]) * inv, __uint_as_float(v[2 * e + 1]) * inv);
            w0.x = *reinterpret_cast<uint32_t*>(&hh[0]); w0.y = *reinterpret_cast<uint32_t*>(&hh[1]);
            w0.z = *reinterpret_cast<uint32_t*>(&hh[2]); w0.w = *reinterpret_cast<uint32_t*>(&hh[3]);
            w1.x = *reinterpret_cast<uint32_t*>(&hh[4]); w1.y = *reinterpret_cast<uint32_t*>(&hh[5]);
            w1.z = *reinterpret_cast<uint32_t*>(&hh[6]); w1.w = *reinterpret_cast<uint32_t*>(&hh[7]);
            *reinterpret_cast<uint4*>(o + j * 16) = w0;
            *reinterpret_cast<uint4*>(o + j * 16 + 8) = w1;
          }
        }
        if (active) w_epi += clock64() - t0;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[t]);
      }
    }
    if (p.dbg_counters && warp == 4 && lane == 0) {
      unsigned long long* c = p.dbg_counters + blockIdx.x * 16;
      c[8] = w_s; c[9] = w_pe; c[10] = w_of; c[11] = w_bar; c[12] = clock64() - t_begin;
      c[5] = w_ld; c[6] = w_max; c[7] = w_exp; c[13] = w_sts; c[14] = w_epi;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// 3-D view of the QKV activation [rows, 3*H*80] fp16 as (16 elements, rows, chunks of 16 columns).
inline int make_tmap_qkv(CUtensorMap* out, const void* qkv, uint64_t rows, uint64_t ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(THMR_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  THMR_CHECK((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && ld % 16 == 0, "attention: qkv pointer/pitch alignment");
  cuuint64_t gdim[3] = {16, rows, ld / 16};
  cuuint64_t gstr[2] = {ld * 2, 32};
  cuuint32_t box[3] = {16, static_cast<cuuint32_t>(kAttTokens), static_cast<cuuint32_t>(kAttChunks)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(THMR_ERR_CUDA, "cuTensorMapEncodeTiled(qkv) failed (%d)", (int)r);
  return THMR_OK;
}

struct AttnPlan {
  CUtensorMap tm;
  CUtensorMap tm_out;   // [B*192, H*80] fp16 output, 32-row x 80-column boxes (third-generation kernel)
  AttnParams p;
  int grid;
};

inline int attention_make_plan(const __half* qkv, int ld_qkv, int B, int heads, __half* out, int ldo, float* dbg_s,
                               AttnPlan* plan) {
  THMR_CHECK(B > 0 && heads > 0, "attention: bad shape");
  THMR_CHECK(ld_qkv >= 3 * heads * kAttHeadDim, "attention: qkv pitch %d < %d", ld_qkv, 3 * heads * kAttHeadDim);
  THMR_CHECK(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "attention: output alignment");
  THMR_TRY(make_tmap_qkv(&plan->tm, qkv, static_cast<uint64_t>(B) * kAttTokens, ld_qkv));
  THMR_TRY(make_tmap_2d_f16(&plan->tm_out, out, static_cast<uint64_t>(B) * kAttTokens,
                            static_cast<uint64_t>(heads) * kAttHeadDim, ldo, 32, kAttHeadDim, CU_TENSOR_MAP_SWIZZLE_NONE));
  plan->p.num_problems = B * heads;
  plan->p.heads = heads;
  plan->p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(kAttHeadDim));
  plan->p.out = out;
  plan->p.ldo = ldo;
  plan->p.dbg_s = dbg_s;
  { const char* e = getenv("THMR_ATTN_TS"); plan->p.p_in_tmem = e ? atoi(e) : 0; }
  { const char* e = getenv("THMR_ATTN_COUNTERS"); plan->p.dbg_counters = e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr; }
  plan->grid = plan->p.num_problems < num_sms() ? plan->p.num_problems : num_sms();
  return THMR_OK;
}

inline int attention_launch(const AttnPlan& plan, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    THMR_CUDA(cudaFuncSetAttribute(vit_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmemBytes));
    configured = true;
  }
  vit_attention_kernel<<<plan.grid, kAttThreads, kAttSmemBytes, st>>>(plan.tm, plan.p);
  THMR_CUDA(cudaGetLastError());
  return THMR_OK;
}

}  // namespace thmr
