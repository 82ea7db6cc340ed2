// Split-K variant of the bf16 GEMM for SMALL problems: a cluster of two CTAs computes ONE 128 x 64 output tile, each over half
// of the reduction; the peer ships its fp32 partial tile into the leader's shared memory (distributed shared memory stores),
// the leader adds it to its own accumulator and runs the usual fused epilogue.
//
// Why: the sequential memory build (reference engine/must3r.py:40-54) is a chain of M = 768 GEMMs with 72 output tiles of
// 128 x 64 — half of the machine — each of which streams its own K x 64 weight slice from L2 at a rate set by one SM's ingest
// (profiles/r01_stage_times.md: 7.0 us for K = 768, 17.1 us for the K = 3072 fc2; profiles/r02_ab_same_box.md: these GEMMs want
// MANY CTAs pulling, the 2-CTA 256 x 256 kernel is slower).  Splitting the reduction puts 144 CTAs on the same 72 tiles and halves
// the length of every CTA's k-loop; the price is one 32 KB DSMEM transfer and a cluster barrier per tile.
// Restrictions (checked by the dispatcher): all-bf16 mode (MODE 0 epilogues), no batches / convolution / TMA-store planes,
// at least 4 k-blocks, one tile per cluster in a single wave.
#pragma once

namespace pst3r {

constexpr int GSK_BN = 64;
constexpr int GSK_STAGES = 6;

__device__ __forceinline__ void st_shared_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
// release at cluster scope: the partial tile written above is visible to the leader once it has seen this arrive
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > PST3R_SPIN_LIMIT) {
      printf("pst3r: split-K reduce barrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_splitk2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmEpi ep,
                    const int M, const int N, const int K) {
  constexpr int BN = GSK_BN;
  constexpr int STAGES = GSK_STAGES;
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [0]: own accumulator complete; [1]: the peer's partial tile has landed (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 4);
  float* red = reinterpret_cast<float*>(smem + L::OUT_OFFSET);  // [64 columns][128 rows] fp32: the peer's partial tile (leader)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.x >> 1;
  const int num_m_blocks = (M + GEMM_BM - 1) / GEMM_BM;
  const int m_blk = tile % num_m_blocks;
  const int n_blk = tile / num_m_blocks;
  const int num_k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  const int kb_mid = (num_k_blocks + 1) >> 1;
  const int kb0 = rank == 0 ? 0 : kb_mid;
  const int kb1 = rank == 0 ? kb_mid : num_k_blocks;
  constexpr uint32_t TMEM_COLS = BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 8);  // one elected lane of each of the peer's 8 epilogue warps
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();  // the leader's barriers are initialised before the peer's remote arrives
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        uint8_t* a_dst = smem + s * L::STAGE_BYTES;
        tma_load_2d(a_dst, &tmA, &full_bar[s], kb * GEMM_BK, m_blk * GEMM_BM);
        tma_load_2d(a_dst + L::A_BYTES, &tmB, &full_bar[s], kb * GEMM_BK, n_blk * BN);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int n_rem = N - n_blk * BN;
      const int n_mma = n_rem >= BN ? BN : ((n_rem + 15) & ~15);
      const uint32_t idesc = make_idesc_bf16(GEMM_BM, n_mma, 0, 0);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t a_desc = make_smem_desc_sw128(a_addr, 0, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(a_addr + L::A_BYTES, 0, 1024);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) umma_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, ((kb - kb0) | k) != 0);
        umma_commit(&empty_bar[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(&tfull_bar[0]);
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int quad = warp & 3;         // TMEM lane quadrant this warp may access
    const int c = (warp - 2) >> 2;     // which 32-column chunk of the 64
    const int r_in_tile = quad * 32 + lane;
    const int row = m_blk * GEMM_BM + r_in_tile;
    const int n_rem = N - n_blk * BN;
    const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + c * 32;
    if (rank != 0) {
      mbar_wait(&tfull_bar[0], 0);
      tc_fence_after();
      if (c * 32 < n_rem) {  // warp-uniform
        uint32_t r[32];
        tmem_ld32(t_addr, r);
        tmem_ld_wait();
        // column-major in the leader's buffer: the 32 lanes of a warp (32 rows) write one 128-byte line per column
        const uint32_t dst = mapa_u32(smem_u32(red + (c * 32) * GEMM_BM + r_in_tile), 0);
#pragma unroll
        for (int i = 0; i < 32; ++i) st_shared_cluster_f32(dst + i * GEMM_BM * 4, __uint_as_float(r[i]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_release_cluster(mapa_u32(smem_u32(&tfull_bar[1]), 0));
    } else {
      const float2 ln = ln_row_stats(ep, row, M);  // before the accumulator wait: overlaps the main loop
      mbar_wait(&tfull_bar[0], 0);
      tc_fence_after();
      if (c * 32 < n_rem) {
        uint32_t r[32];
        tmem_ld32(t_addr, r);
        tmem_ld_wait();
        mbar_wait_acquire_cluster(&tfull_bar[1], 0);
        float v[32];
        const float* src = red + (c * 32) * GEMM_BM + r_in_tile;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + src[i * GEMM_BM];
        epilogue_chunk<false>(ep, v, row, n_blk * BN + c * 32, M, N, 0, ln);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  cluster_sync_all();  // the leader's shared memory stays valid until the peer is done with it (and vice versa at exit)
}

static int launch_gemm_splitk(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpi& ep, int M, int N, int K,
                              cudaStream_t stream) {
  using L = GemmSmem<GSK_BN, GSK_STAGES>;
  static bool configured = false;
  if (!configured) {
    PST3R_CHECK_CUDA(cudaFuncSetAttribute(gemm_splitk2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES_TMA));
    configured = true;
  }
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + GSK_BN - 1) / GSK_BN);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * tiles);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = L::DYN_BYTES_TMA;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];  // the cluster shape (2,1,1) is compiled into the kernel (__cluster_dims__)
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  PST3R_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_splitk2_kernel, tmA, tmB, ep, M, N, K));
  return PST3R_OK;
}

}  // namespace pst3r
