// CTA-quad tcgen05 GEMM: a cluster of FOUR CTAs = two CTA pairs (cta_group::2) that compute two vertically adjacent
// 256 x 256 output tiles and share the weight tile through TMA multicast.
//
// Why: with the warp-converged MMA issuer the CTA-pair kernel moves ~10 TB/s from L2 to the SMs, i.e. it sits at the
// L2 slice throughput cap (B200_MICROARCH: ~6300 B/clk chip-wide); cutting the last round of tiles into quarters
// changed nothing, which is what a bandwidth bound predicts.  The only lever left is bytes per FLOP.  The two pairs of
// a quad work on tiles (2i, n) and (2i+1, n): same 256 weight rows.  Each of the four CTAs fetches ONE QUARTER of the
// B k-block (64 rows, 8 KB) and multicasts it to the CTA of the other pair that holds the same half of B, so a pair
// receives A 32 KB + B 32 KB per k-block as before but only A 32 KB + B 16 KB leave the L2: 25 % less L2 -> SM traffic.
//
//   rank r of the cluster: pair = r >> 1 (m-tile 2i + pair), prank = r & 1 (rows prank*128.. of the pair's A tile,
//                          rows prank*128.. of the B tile)
//   producer (every CTA) : A box 128 x 64 -> own smem;  B box 64 x 64 (rows prank*128 + pair*64..) -> smem of ranks
//                          {prank, prank + 2} at offset pair * 8 KB;  bytes are credited to each destination pair's
//                          leader barrier (64 KB per stage per pair, as in the CTA-pair kernel)
//   smem slot reuse      : a slot may be overwritten by the OTHER pair's producer, so every tcgen05.commit that frees a
//                          slot is multicast to all four CTAs and the "slot free" barriers expect two arrivals
//   everything else      : per pair exactly as gemm2_tcgen05.cuh (TMEM double buffer, epilogue warps, TMA epilogues)
//
// Clusters of four strand part of the chip (GPC sizes are not multiples of four: ~132 of 148 SMs get work), which the
// traffic saving has to pay for first; the host picks this kernel only when THMR_GEMM_QUAD allows it.
#pragma once
#include "gemm2_tcgen05.cuh"

namespace thmr {

constexpr uint32_t kG4BQuarterRows = kG2BN / 4;                       // 64 weight rows fetched per CTA per k-block
constexpr uint32_t kG4BQuarterBytes = kG4BQuarterRows * kGemmBK * 2;  // 8 KB

// 2-SM TMA load multicast to the CTAs in `mask` (same CTA-relative destination); in every destination the bytes are
// credited to the barrier of that CTA's pair leader (peer bit cleared).
__device__ __forceinline__ void tma_load_2d_2sm_mcast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                                      int32_t c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
        "r"(c1), "h"(mask)
      : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_f16_tn_4cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  static_assert(EPI == kEpiStore16 || EPI == kEpiAdd32 || EPI == kEpiStore32, "4-CTA GEMM supports the TMA epilogues only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kG2BarOffset);   // used in the pair leaders only
  uint64_t* empty_bar = full_bar + kG2Stages;                               // 2 arrivals: one commit per pair
  uint64_t* tfull_bar = empty_bar + kG2Stages;
  uint64_t* tempty_bar = tfull_bar + 2;                                     // used in the pair leaders only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = rank >> 1, prank = rank & 1;
  const int cluster_id = blockIdx.x >> 2;
  const int num_clusters = gridDim.x >> 2;

  const int tiles_m = (p.M + 2 * kGemmBM - 1) / (2 * kGemmBM);
  const int tiles_n = (p.N + kG2BN - 1) / kG2BN;
  const int num_kb = (p.K + kGemmBK - 1) / kGemmBK;
  const int num_super = ((tiles_m + 1) / 2) * tiles_n;      // 512 x 256 super tiles (an odd last m-tile is all padding)

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int s = 0; s < kG2Stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * kGemmEpiWarps);   // epilogue warps of both CTAs of the pair
    }
    fence_mbar_init();
  }
  if (warp == kWarpAlloc) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // every CTA's barriers initialised, all TMEM allocations done
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kWarpTma) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (every CTA)
      int stage = 0;
      uint32_t phase = 0;
      const uint16_t bmask = static_cast<uint16_t>((1u << prank) | (1u << (prank + 2)));
      for (int st = cluster_id; st < num_super; st += num_clusters) {
        const int m0 = ((st / tiles_n) * 2 + static_cast<int>(pair)) * 2 * kGemmBM + static_cast<int>(prank) * kGemmBM;
        const int n0 = (st % tiles_n) * kG2BN + static_cast<int>(prank) * (kG2BN / 2) +
                       static_cast<int>(pair) * static_cast<int>(kG4BQuarterRows);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);        // both pairs have retired their MMAs on this slot
          uint8_t* sa = smem + stage * kG2StageBytes;
          uint8_t* sb = sa + kG2ABytes + pair * kG4BQuarterBytes;
          if (prank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kG2StageBytes);
          tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * kGemmBK, m0);
          tma_load_2d_2sm_mcast(sb, &tmB, &full_bar[stage], kb * kGemmBK, n0, bmask);
          if (++stage == kG2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kWarpMma) {
    if (prank == 0) {
      // ---------------------------------------------------------- MMA issuer (pair leaders), warp-converged
      constexpr uint32_t idesc = make_idesc_f16(2 * kGemmBM, kG2BN);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pair));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool ready = false;
      for (int st = cluster_id; st < num_super; st += num_clusters) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + acc * kG2BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!ready) mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kG2StageBytes;
          const uint32_t sb = sa + kG2ABytes;
          const int nstage = (stage + 1 == kG2Stages) ? 0 : stage + 1;
          const uint32_t nphase = (stage + 1 == kG2Stages) ? (phase ^ 1) : phase;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kGemmBK / 16; ++k) {
              const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024, kSwz128);
              const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024, kSwz128);
              umma_f16_ss_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm_mcast(&empty_bar[stage], 0xF);          // slot free: tell the producers of all four CTAs
            if (kb == num_kb - 1) umma_commit_2sm_mcast(&tfull_bar[acc], pair_mask);
          }
          __syncwarp();
          ready = __all_sync(0xffffffffu, mbar_try_wait(&full_bar[nstage], nphase));   // peek the next stage (a hint)
          stage = nstage;
          phase = nphase;
        }
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < kGemmEpiWarps) {
    // -------------------------------------------------------------- epilogue (all CTAs, own 128 rows; as the CTA-pair kernel)
    const int q = warp & 3;
    const int half = warp >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    constexpr int kChunkCols = (EPI == kEpiStore16) ? 64 : 32;
    constexpr int kChunksPerHalf = kG2BN / kChunkCols / 2;
    uint8_t* stage_buf = smem + kG2StagingOffset + warp * 4096;
    const uint32_t srow = smem_u32(stage_buf) + lane * 128;
    const int sw = lane & 7;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int st = cluster_id; st < num_super; st += num_clusters) {
      const int m0 = ((st / tiles_n) * 2 + static_cast<int>(pair)) * 2 * kGemmBM + static_cast<int>(prank) * kGemmBM;
      const int n0 = (st % tiles_n) * kG2BN;
      const bool add_bias = p.bias != nullptr;
      // this warp's 128 bias values (lane l: columns 4l..4l+3 of its column half), fetched while the MMAs run
      float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        const int bc = n0 + half * (kG2BN / 2) + lane * 4;
        if (add_bias && lane * 4 < kG2BN / 2 && bc < p.N) bq = __ldg(reinterpret_cast<const float4*>(p.bias + bc));
      }
      if (lane == 0) mbar_wait(&tfull_bar[acc], acc_phase);   // one polling lane per warp
      __syncwarp();
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < ((p.dbg & 1) ? 0 : kChunksPerHalf); ++cc) {
        const int c = half * kChunksPerHalf + cc;
        const int col0 = n0 + c * kChunkCols;
        uint32_t pk[32];
        if constexpr (EPI == kEpiStore16) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld_x32(tmem_base + lane_addr + acc * kG2BN + c * 64 + hh * 32, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int bl = (cc * 64 + hh * 32 + j) >> 2;   // lane holding these 4 columns' bias
              float4 b4;
              b4.x = __shfl_sync(0xffffffffu, bq.x, bl); b4.y = __shfl_sync(0xffffffffu, bq.y, bl);
              b4.z = __shfl_sync(0xffffffffu, bq.z, bl); b4.w = __shfl_sync(0xffffffffu, bq.w, bl);
              f[j] = __uint_as_float(v[j]) + b4.x; f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
              f[j + 2] = __uint_as_float(v[j + 2]) + b4.z; f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
            }
            // warp-uniform branch OUTSIDE the element loop (otherwise the compiler if-converts it and every element
            // pays for GELU and ReLU even when no activation is requested)
            if (p.act == kActGelu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
            } else if (p.act == kActRelu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              __half2 h2 = __floats2half2_rn(f[j], f[j + 1]);
              pk[hh * 16 + (j >> 1)] = *reinterpret_cast<uint32_t*>(&h2);
            }
          }
        } else {
          uint32_t v[32];
          tmem_ld_x32(tmem_base + lane_addr + acc * kG2BN + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int bl = (cc * 32 + j) >> 2;
            float4 b4;
            b4.x = __shfl_sync(0xffffffffu, bq.x, bl); b4.y = __shfl_sync(0xffffffffu, bq.y, bl);
            b4.z = __shfl_sync(0xffffffffu, bq.z, bl); b4.w = __shfl_sync(0xffffffffu, bq.w, bl);
            pk[j] = __float_as_uint(fmaf(p.alpha, __uint_as_float(v[j]), b4.x));
            pk[j + 1] = __float_as_uint(fmaf(p.alpha, __uint_as_float(v[j + 1]), b4.y));
            pk[j + 2] = __float_as_uint(fmaf(p.alpha, __uint_as_float(v[j + 2]), b4.z));
            pk[j + 3] = __float_as_uint(fmaf(p.alpha, __uint_as_float(v[j + 3]), b4.w));
          }
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((j ^ sw) << 4)), "r"(pk[4 * j]),
                       "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && col0 < p.N && m0 + q * 32 < p.M && !(p.dbg & 2)) {
          if constexpr (EPI == kEpiStore16) tma_store_2d(&tmC, stage_buf, col0, m0 + q * 32);
          else if constexpr (EPI == kEpiStore32) tma_store_2d(&tmC, stage_buf, col0, m0 + q * 32);
          else tma_reduce_add_2d(&tmC, stage_buf, col0, m0 + q * 32);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], pair * 2);   // accumulator buffer free (pair leader's barrier)
      if ((acc ^= 1) == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // no CTA may exit (or free TMEM) while another can still reach it
  if (warp == kWarpAlloc) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace thmr
