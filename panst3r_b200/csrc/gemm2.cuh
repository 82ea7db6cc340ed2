// 2-CTA (cta_group::2) variant of the bf16 GEMM for large problems: a cluster of two CTAs (one SM pair) computes a
// 256 x 256 output tile.  CTA r of the pair loads its own 128 x 64 A tile and HALF of the 256 x 64 B tile per k-block;
// the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16) which reads A/B from both
// CTAs' shared memory and accumulates 128 rows into each CTA's TMEM.  Per SM this moves 32 KB per k-block instead of
// 48 KB for the same 128 x 256 x 64 of tensor work — the 1-CTA kernel is bound by exactly that L2->smem traffic
// (~65 % of the cuBLAS peak).  Same epilogues (epilogue_chunk) as the 1-CTA kernel.
//   TMA completion of both CTAs lands on the LEADER's full barrier (cta_group::2 TMA form, peer bit masked);
//   smem slots / accumulators are released to both CTAs by multicast tcgen05.commit;
//   the peer's epilogue warps release accumulators with a remote mbarrier arrive on the leader.
#pragma once

namespace pst3r {

constexpr int G2_BN = 256;
constexpr int G2_STAGES = 6;
constexpr uint32_t G2_A_BYTES = GEMM_BM * GEMM_BK * 2;          // 16 KB
constexpr uint32_t G2_B_BYTES = (G2_BN / 2) * GEMM_BK * 2;      // 16 KB (half of the B tile)
constexpr uint32_t G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr uint32_t G2_BAR_OFFSET = G2_STAGES * G2_STAGE_BYTES;
constexpr uint32_t G2_TOTAL = G2_BAR_OFFSET + (2 * G2_STAGES + 4) * 8 + 16;
constexpr uint32_t G2_DYN_BYTES = G2_TOTAL + 1024;
constexpr uint32_t G2_OUT_OFFSET = (G2_TOTAL + 127) & ~127u;  // TMA-store staging tiles: requested only by launches using them
constexpr uint32_t G2_DYN_BYTES_TMA = G2_OUT_OFFSET + GEMM_OUT_STAGE_BYTES + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
// Remote arrive on the leader's barrier.  RELAXED on purpose: the only thing this arrive has to order is the epilogue's
// tcgen05.ld of the accumulator against the leader's next tcgen05.mma, and that is done by tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync on this side and tcgen05.fence::after_thread_sync on the waiter's.  The default
// .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR, i.e. every epilogue warp drained its 16 KB of global
// stores per tile before it could signal (ncu: stall_membar 8.7 % of all samples of the fc1 GEMM).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose mbarrier lives in the leader CTA of the pair (shared::cluster address with the peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the issuing thread's prior MMAs completed) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

// PROMOTE: accumulator promotion for split-mode reductions (GemmEpi::promote): every epilogue thread then keeps 128 fp32
// partial sums in registers.  320 threads put 3 warps on one SM sub-partition (16 K registers), which caps a thread at
// 168 registers: ptxas spills 64 bytes of that variant's epilogue (a launch with more registers per thread is refused).
template <int MODE>  // as for gemm_bf16_tn_kernel: 0 all-bf16 hot path, 1 split operands / outputs, 2 split + promotion
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmOut, const GemmEpi ep, const int M, const int N, const int K) {
  constexpr int BN = G2_BN;
  constexpr int STAGES = G2_STAGES;
  constexpr bool PROMOTE = MODE == 2;
  constexpr bool SPLIT_IO = MODE != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int num_m_blocks = (M + 255) / 256;
  const int num_n_blocks = N / BN;  // host guarantees N % 256 == 0
  const int num_tiles = num_m_blocks * num_n_blocks;
  const int num_k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  const int terms = (SPLIT_IO && ep.split_terms > 1) ? ep.split_terms : 1;  // split-bf16 mode: see GemmEpi::split_terms
  const int num_k_iters = num_k_blocks * terms;
  const int chunk_iters = PROMOTE ? ep.promote : num_k_iters;
  const int num_chunks = (num_k_iters + chunk_iters - 1) / chunk_iters;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's: its own arrive.expect_tx; bytes of both CTAs' TMA loads
      mbar_init(&empty_bar[s], 1);  // per CTA: multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);    // per CTA: multicast commit
      mbar_init(&tempty_bar[s], 16);  // leader's: 8 epilogue warps of each CTA
    }
    mbar_fence_init();
  }
  cluster_sync_all();  // barriers of both CTAs initialised before any remote traffic
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile % num_m_blocks;
        const int n_blk = tile / num_m_blocks;
        for (int term = 0; term < terms; ++term) {
        const int pa = (terms == 3 && term == 1) ? 1 : 0;
        const int pb = (terms > 1 && term == terms - 1) ? 1 : 0;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * G2_STAGE_BYTES);
          uint8_t* a_dst = smem + s * G2_STAGE_BYTES;
          uint8_t* b_dst = a_dst + G2_A_BYTES;
          if (terms > 1) {
            tma_load_3d_2sm(a_dst, &tmA, &full_bar[s], kb * GEMM_BK, m_blk * 256 + (int)rank * 128, pa);
            tma_load_3d_2sm(b_dst, &tmB, &full_bar[s], kb * GEMM_BK, n_blk * BN + (int)rank * (BN / 2), pb);
          } else {
            tma_load_2d_2sm(a_dst, &tmA, &full_bar[s], kb * GEMM_BK, m_blk * 256 + (int)rank * 128);
            tma_load_2d_2sm(b_dst, &tmB, &full_bar[s], kb * GEMM_BK, n_blk * BN + (int)rank * (BN / 2));
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) ---------------
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int local = 0;  // accumulator hand-overs so far: one per (tile, chunk)
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        for (int ch = 0; ch < num_chunks; ++ch, ++local) {
          const int acc = local & 1;
          const uint32_t acc_ph = (local >> 1) & 1;
          mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          const int k0 = ch * chunk_iters;
          const int k1 = min(k0 + chunk_iters, num_k_iters);
          for (int kb = k0; kb < k1; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + s * G2_STAGE_BYTES);
            const uint32_t b_addr = a_addr + G2_A_BYTES;
            const uint64_t a_desc = make_smem_desc_sw128(a_addr, 0, 1024);
            const uint64_t b_desc = make_smem_desc_sw128(b_addr, 0, 1024);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              umma_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, ((kb - k0) | k) != 0);
            umma_commit_2sm(&empty_bar[s]);
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          umma_commit_2sm(&tfull_bar[acc]);
        }
      }
    }
  } else {
    // ------------------------------- epilogue (both CTAs) -----------------------
    const int quad = warp & 3;
    const int cgrp = (warp - 2) >> 2;
    float* out_stage = reinterpret_cast<float*>(smem + G2_OUT_OFFSET) + (warp - 2) * 1024;
    int local = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_blk = tile % num_m_blocks;
      const int n_blk = tile / num_m_blocks;
      const int row = m_blk * 256 + (int)rank * 128 + quad * 32 + lane;
      const float2 ln = ln_row_stats(ep, row, M);
      if constexpr (PROMOTE) {
        // this warp owns the 32-column chunks cgrp, cgrp + 2, ...: BN / 64 register accumulators of 32 columns each
        constexpr int NCH = BN / 64;
        float accv[NCH][32];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
          for (int i = 0; i < 32; ++i) accv[j][i] = 0.0f;
        for (int ch = 0; ch < num_chunks; ++ch, ++local) {
          const int acc = local & 1;
          mbar_wait(&tfull_bar[acc], (local >> 1) & 1);
          tc_fence_after();
          const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            uint32_t r[32];
            tmem_ld32(t_base + (cgrp + 2 * j) * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) accv[j][i] += __uint_as_float(r[i]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));  // leader's barrier
        }
#pragma unroll
        for (int j = 0; j < NCH; ++j) epilogue_chunk<true>(ep, accv[j], row, n_blk * BN + (cgrp + 2 * j) * 32, M, N, 0, ln);
      } else {
        const int acc = local & 1;
        const uint32_t acc_ph = (local >> 1) & 1;
        mbar_wait(&tfull_bar[acc], acc_ph);
        tc_fence_after();
        const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = cgrp; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld32(t_base + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (ep.tma_store)
            epilogue_chunk_tma(ep, &tmOut, out_stage, v, m_blk * 256 + (int)rank * 128 + quad * 32, n_blk * BN + c * 32, lane);
          else
            epilogue_chunk<SPLIT_IO>(ep, v, row, n_blk * BN + c * 32, M, N, 0, ln);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));  // leader's barrier
        ++local;
      }
    }
    if (ep.tma_store && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer must not exit (or free TMEM) while the leader's MMAs still read its smem / write its TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

static bool gemm2_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PST3R_GEMM2");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const GemmEpi& ep, int M,
                        int N, int K, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    PST3R_CHECK_CUDA(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_DYN_BYTES_TMA));
    PST3R_CHECK_CUDA(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_DYN_BYTES_TMA));
    PST3R_CHECK_CUDA(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_DYN_BYTES_TMA));
    configured = true;
  }
  const int tiles = ((M + 255) / 256) * (N / G2_BN);
  const int max_pairs = sm_budget() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = ep.tma_store ? G2_DYN_BYTES_TMA : G2_DYN_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];  // the cluster shape (2,1,1) is compiled into the kernel (__cluster_dims__)
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (ep.promote)
    PST3R_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm2_bf16_tn_kernel<2>, tmA, tmB, tmOut, ep, M, N, K));
  else if (ep.split_terms || ep.out_kind == KIND_SPLIT || (ep.residual && ep.res_kind != KIND_BF16))
    PST3R_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm2_bf16_tn_kernel<1>, tmA, tmB, tmOut, ep, M, N, K));
  else
    PST3R_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm2_bf16_tn_kernel<0>, tmA, tmB, tmOut, ep, M, N, K));
  return PST3R_OK;
}

}  // namespace pst3r
