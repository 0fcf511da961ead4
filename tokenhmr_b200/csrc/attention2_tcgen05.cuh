// Fused ViT attention, second generation: TWO heads in flight per CTA, probabilities kept in TMEM.
//
// The first kernel (attention_tcgen05.cuh) processes one head at a time and every head is a serial chain
// S-MMA -> softmax -> P (smem) -> PV-MMA -> epilogue; its counters show each role idle ~60 % of the time.
// Running two CTAs per SM would hide that, but a kernel that executes tcgen05.alloc is limited to ONE resident
// CTA per SM on this toolchain (scripts/probe/occ_probe.cu: occupancy 1 even with 12 registers and no smem).
// So the CTA is split into two independent "slots" of 10 warps; each slot owns 256 TMEM columns, ~93 KB of
// shared memory, its own barriers and its own stream of heads, and the warp schedulers interleave one slot's
// MMAs / TMA loads with the other slot's exp-heavy softmax.
//
//   per (image, head), per 128-row tile t (rows 0..127, then rows 64..191 whose first 64 rows are duplicates):
//     S_t = Q_t K^T            tcgen05.mma SS, fp32 in TMEM columns [0,192)
//     pass 1: row max          thread = (row, half of the 192 keys); S re-read from TMEM in pass 2 (register budget)
//     pass 2: P = exp2(...)    fp16 pairs written back IN PLACE over the consumed S columns (tcgen05.st):
//                              keys 0..95 -> columns [0,48), keys 96..191 -> columns [96,144)
//     O_t = P V                tcgen05.mma TS (A operand = P from TMEM), V N-major from smem, fp32 in columns [176,256)
//     epilogue                 O / rowsum -> fp16 -> global
//
//   per slot (warp w = warp % 10):
//   w 0 : TMA producer (Q,K and V single-buffered; the next head's Q,K load is issued as soon as S_1 retires)
//   w 1 : MMA issuer (slot 0's also allocates the 512 TMEM columns)
//   w 2..9 : softmax / epilogue
#pragma once
#include <stdlib.h>

#include "attention_tcgen05.cuh"

namespace thmr {

constexpr uint32_t kAtt2OffK = kAttMatBytes;
constexpr uint32_t kAtt2OffV = 2 * kAttMatBytes;
constexpr uint32_t kAtt2OffStats = 3 * kAttMatBytes;                 // smax[2][128], ssum[2][128]
constexpr uint32_t kAtt2OffBars = kAtt2OffStats + 2 * 2 * 128 * 4;
constexpr uint32_t kAtt2SlotBytes = (kAtt2OffBars + 128 + 1023) / 1024 * 1024;
constexpr int kAtt2SlotWarps = 10;
constexpr uint32_t kAtt2TmemCols = 256;                           // per slot
constexpr uint32_t kAtt2ColO = 176;

template <int kAtt2Slots>
__global__ void __launch_bounds__(kAtt2Slots * kAtt2SlotWarps * 32, 1)
vit_attention2_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int slot = (threadIdx.x >> 5) / kAtt2SlotWarps;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023)) +
                  slot * kAtt2SlotBytes;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kAtt2OffK;
  uint8_t* sV = smem + kAtt2OffV;
  float* smax = reinterpret_cast<float*>(smem + kAtt2OffStats);   // [half][128]
  float* ssum = smax + 2 * 128;                                    // [half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtt2OffBars);
  uint64_t* qk_full = bars + 0;
  uint64_t* qk_empty = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* o_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = (threadIdx.x >> 5) % kAtt2SlotWarps;   // role index inside the slot
  const int lane = threadIdx.x & 31;
  const int first_prob = blockIdx.x * kAtt2Slots + slot;
  const int prob_stride = gridDim.x * kAtt2Slots;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    fence_mbar_init();
  }
  uint32_t* tmem_slot0 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(tmem_slot) - slot * kAtt2SlotBytes);
  if (warp == 1 && slot == 0) {
    tmem_alloc(tmem_slot0, kAtt2Slots * kAtt2TmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot0 + slot * kAtt2TmemCols;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      int i = 0;
      for (int prob = first_prob; prob < p.num_problems; prob += prob_stride, ++i) {
        const int b = prob / p.heads, h = prob % p.heads;
        mbar_wait(qk_empty, (i & 1) ^ 1);
        mbar_arrive_expect_tx(qk_full, 2 * kAttMatBytes);
        tma_load_3d(sQ, &tmQKV, qk_full, 0, b * kAttTokens, h * kAttChunks);
        tma_load_3d(sK, &tmQKV, qk_full, 0, b * kAttTokens, (p.heads + h) * kAttChunks);
        mbar_wait(v_empty, (i & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, kAttMatBytes);
        tma_load_3d(sV, &tmQKV, v_full, 0, b * kAttTokens, (2 * p.heads + h) * kAttChunks);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = make_idesc_f16(128, kAttTokens);
      constexpr uint32_t idesc_o = make_idesc_f16(128, kAttHeadDim, 0, 1);
      const uint32_t sQa = smem_u32(sQ), sKa = smem_u32(sK), sVa = smem_u32(sV);
      int i = 0;
      uint32_t n = 0;   // tile counter (2 per head): parity of the per-tile barriers
      for (int prob = first_prob; prob < p.num_problems; prob += prob_stride, ++i) {
        mbar_wait(qk_full, i & 1);
        for (int t = 0; t < 2; ++t, ++n) {
          mbar_wait(o_empty, (n & 1) ^ 1);          // previous tile's epilogue has drained O (aliases S)
          tc_fence_after();
#pragma unroll
          for (int kc = 0; kc < kAttChunks; ++kc) {
            const uint64_t da = make_smem_desc(sQa + kc * kAttChunkBytes + t * 64 * 32, 16, 256, kSwz32);
            const uint64_t db = make_smem_desc(sKa + kc * kAttChunkBytes, 16, 256, kSwz32);
            umma_f16_ss(tmem_base, da, db, idesc_s, kc != 0);
          }
          umma_commit(s_full);
          if (t == 1) umma_commit(qk_empty);        // Q,K free for the next head
          if (t == 0) mbar_wait(v_full, i & 1);
          mbar_wait(p_full, n & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < kAttTokens / 16; ++ks) {
            const uint32_t pa = tmem_base + (ks < 6 ? ks * 8 : 96 + (ks - 6) * 8);
            const uint64_t db = make_smem_desc(sVa + ks * 512, kAttChunkBytes, 256, kSwz32);
            umma_f16_ts(tmem_base + kAtt2ColO, pa, db, idesc_o, ks != 0);
          }
          umma_commit(o_full);
          if (t == 1) umma_commit(v_empty);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warps
    const int q = (threadIdx.x >> 5) & 3;   // TMEM lane quarter is fixed by the PHYSICAL warp id (slot 1 starts at warp 10)
    const int half = (warp - 2) >> 2;
    const int trow = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + half * 96;
    int i = 0;
    uint32_t n = 0;
    for (int prob = first_prob; prob < p.num_problems; prob += prob_stride, ++i) {
      const int b = prob / p.heads, h = prob % p.heads;
      for (int t = 0; t < 2; ++t, ++n) {
        const bool active = (t == 0) || (q >= 2);
        if (lane == 0) mbar_wait(s_full, n & 1);
        __syncwarp();
        tc_fence_after();
        if (active) {
          // ---- pass 1: row maximum over this thread's 96 keys
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;   // 4 chains, not one of 96
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            uint32_t v[16];
            tmem_ld_x16(s_addr + j * 16, v);
            tmem_ld_wait();
            if (p.dbg_s) {
              float* d = p.dbg_s + (static_cast<size_t>(prob) * kAttTokens + t * 64 + trow) * kAttTokens + half * 96 + j * 16;
#pragma unroll
              for (int e = 0; e < 16; ++e) d[e] = __uint_as_float(v[e]);
            }
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[e])); m1 = fmaxf(m1, __uint_as_float(v[e + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[e + 2])); m3 = fmaxf(m3, __uint_as_float(v[e + 3]));
            }
          }
          float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          smax[half * 128 + trow] = m;
          // constant barrier ids: a run-time id makes ptxas reserve all 16 hardware barriers, which caps the SM at
          // one resident CTA
          switch (slot * 4 + q) {
            case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
            case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
            case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
            case 3: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
            case 4: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
            case 5: asm volatile("bar.sync 6, 64;" ::: "memory"); break;
            case 6: asm volatile("bar.sync 7, 64;" ::: "memory"); break;
            default: asm volatile("bar.sync 8, 64;" ::: "memory"); break;
          }
          m = fmaxf(m, smax[(half ^ 1) * 128 + trow]);
          const float mo = m * p.scale_log2e;
          // ---- pass 2: exponentials, row sum, P (fp16) written in place over the consumed S columns
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            uint32_t v[16];
            tmem_ld_x16(s_addr + j * 16, v);
            tmem_ld_wait();
            uint32_t w8[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const float e0 = fast_exp2(fmaf(__uint_as_float(v[e]), p.scale_log2e, -mo));
              const float e1 = fast_exp2(fmaf(__uint_as_float(v[e + 1]), p.scale_log2e, -mo));
              sum0 += e0;
              sum1 += e1;
              __half2 h2 = __floats2half2_rn(e0, e1);
              w8[e >> 1] = *reinterpret_cast<uint32_t*>(&h2);
            }
            tmem_st_x8(s_addr + j * 8, w8);          // keys [half*96 + 16j, +16) -> columns half*96 + 8j .. +8
          }
          const float sum = sum0 + sum1;
          tmem_st_wait();
          ssum[half * 128 + trow] = sum;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // ---- epilogue of this tile
        if (lane == 0) mbar_wait(o_full, n & 1);
        __syncwarp();
        tc_fence_after();
        if (active) {
          const float inv = 1.0f / (ssum[trow] + ssum[128 + trow]);
          const int row = b * kAttTokens + t * 64 + trow;
          const int c0 = half ? 48 : 0;
          const int nchunk = half ? 2 : 3;
          __half* o = p.out + static_cast<size_t>(row) * p.ldo + h * kAttHeadDim + c0;
          for (int j = 0; j < nchunk; ++j) {
            uint32_t v[16];
            tmem_ld_x16(tmem_base + lane_addr + kAtt2ColO + c0 + j * 16, v);
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
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1 && slot == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAtt2Slots * kAtt2TmemCols);
  }
}

template <int SLOTS>
inline int attention2_launch_t(const AttnPlan& plan, cudaStream_t st) {
  constexpr int threads = SLOTS * kAtt2SlotWarps * 32;
  constexpr uint32_t smem = SLOTS * kAtt2SlotBytes + 1024;
  static bool configured = false;
  if (!configured) {
    THMR_CUDA(cudaFuncSetAttribute(vit_attention2_kernel<SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int groups = (plan.p.num_problems + SLOTS - 1) / SLOTS;
  const int grid = groups < num_sms() ? groups : num_sms();
  vit_attention2_kernel<SLOTS><<<grid, threads, smem, st>>>(plan.tm, plan.p);
  THMR_CUDA(cudaGetLastError());
  return THMR_OK;
}

// THMR_ATTN_SLOTS = heads in flight per CTA (1 or 2)
inline int attention2_launch(const AttnPlan& plan, cudaStream_t st) {
  static const int slots = [] { const char* e = getenv("THMR_ATTN_SLOTS"); return e ? atoi(e) : 1; }();
  return slots == 2 ? attention2_launch_t<2>(plan, st) : attention2_launch_t<1>(plan, st);
}

}  // namespace thmr
