// tcgen05 flash attention for head_dim 64 without a block mask: 256 queries per CTA as two 128-row sub-tiles that
// ping-pong on the tensor core, with the probabilities kept in TENSOR MEMORY instead of shared memory.
//
//   S_t = Q_t K_j^T       SS MMA (Q, K from swizzled smem)       -> TMEM columns [t*128, t*128+128)   fp32
//   P_t = 2^((S_t - m) c) softmax warps: tcgen05.ld -> ex2 -> bf16x2 -> tcgen05.st -> TMEM [384 + t*64, +64)
//   O_t = P_t V_j         TS MMA (A = P from TMEM, B = V from smem, MN-major) -> TMEM [256 + t*64, +64)
//
// The previous generation (round 1, removed) wrote P into a swizzled smem tile and appended a ones column to V for the
// row sums: per 128x128 sub-tile step the tensor core then fetched 84 KB of operands from shared memory and the softmax
// warps stored another 32 KB into it, and its timeline (profiles/r01_attention_analysis.md) shows the MMA issue stream
// waiting on exactly that port (P V alone needs 162 B/clk against the 128 B/clk an SM can read).  With P in TMEM
// the P V product reads only V (16 KB per step), the softmax warps issue 4 tcgen05.st instead of 16 st.shared,
// and the freed 80 KB of smem deepen the K/V ring to 4 stages.  Row sums are accumulated by the softmax threads (fp32).
// CTA: 640 threads = TMA warp, MMA warp, 2 idle warps, 8 + 8 softmax warps: TWO threads per score row (one per 64-key half
// of the tile; they exchange their half-row maxima through shared memory and a 64-thread named barrier once per tile, their
// row sums once at the end) and each folds half of the output columns.  Round 1 ran one thread per row (4 + 4 warps): with a
// single softmax warp of each sub-tile per SM sub-partition nobody filled the issue slots a warp left empty while it waited
// for TMEM, a barrier or the MUFU (profiles/r02_attention_single_sweep.md: removing a quarter of the instructions changed
// nothing); four independent warps per sub-partition do.
#pragma once

namespace pst3r {

constexpr int AT3_THREADS = 640;
constexpr int AT3_STAGES = 4;
constexpr uint32_t AT3_OFF_Q = 0;                                       // 2 x 16 KB
constexpr uint32_t AT3_OFF_K = AT3_OFF_Q + 2 * ATT_ATOM_BYTES;          // 4 x 16 KB
constexpr uint32_t AT3_OFF_V = AT3_OFF_K + AT3_STAGES * ATT_ATOM_BYTES; // 4 x 16 KB
constexpr uint32_t AT3_OFF_BAR = AT3_OFF_V + AT3_STAGES * ATT_ATOM_BYTES;
// barriers: q_full, k_full[ST], v_full[ST], kv_empty[ST], s_full[2], p_full[2], o_full[2]
constexpr int AT3_NUM_BARS = 1 + 3 * AT3_STAGES + 6;
constexpr uint32_t AT3_OFF_X = AT3_OFF_BAR + 256;  // fp32 exchange area [2 buffers][2 sub-tiles][2 halves][128 rows]
constexpr uint32_t AT3_DYN_BYTES = AT3_OFF_X + 2 * 2 * 2 * 128 * 4 + 1024;
constexpr uint32_t AT3_TMEM_S = 0;     // S_A at 0, S_B at 128 (fp32)
constexpr uint32_t AT3_TMEM_O = 256;   // O_A at 256, O_B at 320 (fp32)
constexpr uint32_t AT3_TMEM_P = 384;   // P_A at 384, P_B at 448 (bf16 pairs: column c of row r = keys 2c, 2c+1)

// TRACE: development aid (PST3R_ATT_TRACE=<file>): CTA (0,0,0) records clock64() at the hand-over points of the MMA thread
// and of one softmax thread per (sub-tile, half) into p.trace[(slot * 16 + event) * 128 + tile]
#define AT3_TR(slot, ev, j) do { if (TRACE && tr_on && (j) < 128) p.trace[(((slot) * 16 + (ev)) << 7) + (j)] = clock64(); } while (0)

template <bool TRACE>
__global__ void __launch_bounds__(AT3_THREADS, 1)
attention3_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int HD = 64;
  constexpr int ST = AT3_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT3_OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + ST;
  uint64_t* kv_empty = v_full + ST;
  uint64_t* s_full = kv_empty + ST;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AT3_NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blk = blockIdx.x;  // 256 queries
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh - b * p.H;
  const int split = blockIdx.z;
  const int kvb = p.kv_shared ? 0 : b;
  const bool tr_on = TRACE && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;

  const int total_tiles = (p.Nk + ATT_BN - 1) / ATT_BN;
  const int tiles_per_split = (total_tiles + p.splits - 1) / p.splits;
  const int t0 = split * tiles_per_split;
  const int t1 = min(total_tiles, t0 + tiles_per_split);
  const int n_tiles = max(0, t1 - t0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&o_full[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0 && n_tiles > 0) {
      mbar_expect_tx(q_full, 2 * ATT_ATOM_BYTES);
      tma_load_4d(smem + AT3_OFF_Q, &tmQ, q_full, 0, q_blk * 256, h, b);
      tma_load_4d(smem + AT3_OFF_Q + ATT_ATOM_BYTES, &tmQ, q_full, 0, q_blk * 256 + 128, h, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
        const int key0 = (t0 + j) * ATT_BN;
        mbar_expect_tx(&k_full[s], ATT_ATOM_BYTES);
        tma_load_4d(smem + AT3_OFF_K + s * ATT_ATOM_BYTES, &tmK, &k_full[s], 0, key0, h, kvb);
        mbar_expect_tx(&v_full[s], ATT_ATOM_BYTES);
        tma_load_4d(smem + AT3_OFF_V + s * ATT_ATOM_BYTES, &tmV, &v_full[s], 0, key0, h, kvb);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, HD, 0, 1);  // A = P (TMEM, K-major), B = V (MN-major)
      const uint32_t q_addr = smem_u32(smem + AT3_OFF_Q);
      auto issue_s = [&](int t, int j) {  // caller has waited for k_full(j) and knows S_t is free
        const uint32_t k_addr = smem_u32(smem + AT3_OFF_K + (j % ST) * ATT_ATOM_BYTES);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_ss(tmem_base + AT3_TMEM_S + t * ATT_BN, make_smem_desc_sw128(q_addr + t * ATT_ATOM_BYTES + ks * 32, 0, 1024),
                  make_smem_desc_sw128(k_addr + ks * 32, 0, 1024), idesc_s, ks != 0);
        umma_commit(&s_full[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        if (j + 1 < n_tiles) mbar_wait(&k_full[(j + 1) % ST], ((j + 1) / ST) & 1);
        mbar_wait(&v_full[s], (j / ST) & 1);
        const uint32_t v_addr = smem_u32(smem + AT3_OFF_V + s * ATT_ATOM_BYTES);
        AT3_TR(0, 0, j);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          // p_full(t, j): P_t(j) is in TMEM, S_t(j) has been fully read, O_t(j-1) has been consumed
          mbar_wait(&p_full[t], j & 1);
          AT3_TR(0, 1 + 2 * t, j);
          tc_fence_after();
          if (j + 1 < n_tiles) issue_s(t, j + 1);  // next scores first: the softmax of t restarts soonest
#pragma unroll
          for (int ks = 0; ks < ATT_BN / 16; ++ks)  // 16 keys per step = 8 TMEM columns of P, 16 k-rows (2 KB) of V
            umma_ts(tmem_base + AT3_TMEM_O + t * HD, tmem_base + AT3_TMEM_P + t * (ATT_BN / 2) + ks * 8,
                    make_smem_desc_sw128(v_addr + ks * 16 * 128, ATT_ATOM_BYTES, 1024), idesc_pv, ks != 0);
          umma_commit(&o_full[t]);
          AT3_TR(0, 2 + 2 * t, j);
        }
        umma_commit(&kv_empty[s]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- softmax / accumulate -----------------------
    const int t = (warp - 4) >> 3;         // sub-tile
    const int quad = warp & 3;             // TMEM lane quadrant
    const int half = ((warp - 4) >> 2) & 1;  // which 64 keys of a tile / which 32 output columns
    const int r = quad * 32 + lane;
    const int q = q_blk * 256 + t * 128 + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t s_addr = lane_addr + AT3_TMEM_S + t * ATT_BN + half * 64;
    const uint32_t o_addr = lane_addr + AT3_TMEM_O + t * HD + half * 32;
    const uint32_t p_addr = lane_addr + AT3_TMEM_P + t * (ATT_BN / 2) + half * 32;
    float* xchg = reinterpret_cast<float*>(smem + AT3_OFF_X);
    const int bar_id = 1 + t * 4 + quad;   // named barrier of the two warps that share these 32 rows
    const int tr_slot = 1 + t + 2 * half;
    const bool tr_me = quad == 0 && lane == 0;
#define AT3_TRS(ev, j) do { if (tr_me) AT3_TR(tr_slot, ev, j); } while (0)
    float o_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o_acc[i] = 0.0f;
    float m_run = -CUDART_INF_F, l_run = 0.0f, a_prev = 0.0f;
    const float c = p.scale_log2;

    auto consume_o = [&](int j, float alpha) {  // O_acc = O_acc * alpha + O_t(j), this thread's 32 columns
      mbar_wait(&o_full[t], j & 1);
      AT3_TRS(3, j + 1);
      tc_fence_after();
      if (TRACE && (p.trace_mode & 2)) return;
      uint32_t rr[32];
      tmem_ld32(o_addr, rr);
      tmem_ld_wait();
      const float2 al2 = make_float2(alpha, alpha);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 r2 = __ffma2_rn(make_float2(o_acc[i], o_acc[i + 1]), al2,
                                     make_float2(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1])));
        o_acc[i] = r2.x; o_acc[i + 1] = r2.y;
      }
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int kv = p.Nk - (t0 + j) * ATT_BN - half * 64;  // valid keys among this thread's 64 (may exceed 64 / be <= 0)
      mbar_wait(&s_full[t], j & 1);
      AT3_TRS(0, j);
      tc_fence_after();
      float mx = -CUDART_INF_F;
      if (TRACE && (p.trace_mode & 1)) {
        mx = 8.0f;
      } else if (kv >= 64) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t rr[32];
          tmem_ld32(s_addr + ch * 32, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(rr[i]));
        }
      } else {
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t rr[32];
          tmem_ld32(s_addr + ch * 32, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ch * 32 + i < kv) mx = fmaxf(mx, __uint_as_float(rr[i]));
        }
      }
      // the other half of the row: exchange the half-row maxima (double buffered across tiles)
      float* xb = xchg + (((j & 1) * 2 + t) * 2) * 128;
      xb[half * 128 + r] = mx;
      AT3_TRS(1, j);
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      AT3_TRS(2, j);
      mx = fmaxf(fmaxf(mx, xb[(1 - half) * 128 + r]), m_run);
      const float m_use = (mx == -CUDART_INF_F) ? 0.0f : mx;
      const float alpha = (m_run == -CUDART_INF_F) ? 0.0f : ex2_approx((m_run - m_use) * c);
      const float neg_m = -m_use * c;
      m_run = mx;
      if (j > 0) consume_o(j - 1, a_prev);  // also guarantees P V of tile j-1 is done reading P_t from TMEM
      a_prev = alpha;
      AT3_TRS(4, j);
      float sum0 = 0.0f, sum1 = 0.0f;
      float2 sA = make_float2(0.0f, 0.0f), sB = make_float2(0.0f, 0.0f);
      const float2 c2 = make_float2(c, c), nm2 = make_float2(neg_m, neg_m);
      if (TRACE && (p.trace_mode & 4)) {
        // timing experiment: no sweep at all
      } else if (TRACE && (p.trace_mode & 8)) {
        // timing experiment: tensor-memory traffic of the sweep without its arithmetic
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t rr[32];
          tmem_ld32(s_addr + ch * 32, rr);
          tmem_ld_wait();
          tmem_st16(p_addr + ch * 16, rr);
        }
      } else if (kv >= 64) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t rr[32];
          if (TRACE && (p.trace_mode & 16)) {
            // timing experiment: the arithmetic of the sweep without reading the scores
#pragma unroll
            for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(o_acc[i] + (float)ch);
          } else {
            tmem_ld32(s_addr + ch * 32, rr);
            tmem_ld_wait();
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            // packed fp32x2 scale / shift and row-sum adds (sm_100): 2.5 instead of 3.5 issue slots per score
            const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1])), c2, nm2);
            const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3])), c2, nm2);
            const float e0 = ex2_approx(x01.x);
            const float e1 = ex2_approx(x01.y);
            const float e2 = ex2_approx(x23.x);
            const float e3 = ex2_approx(x23.y);
            sA = __fadd2_rn(sA, make_float2(e0, e1));
            sB = __fadd2_rn(sB, make_float2(e2, e3));
            pk[i >> 1] = pack_bf16x2(e0, e1);
            pk[(i >> 1) + 1] = pack_bf16x2(e2, e3);
          }
          tmem_st16(p_addr + ch * 16, pk);
        }
      } else {
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t rr[32];
          tmem_ld32(s_addr + ch * 32, rr);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float e0 = (ch * 32 + i < kv) ? ex2_approx(fmaf(__uint_as_float(rr[i]), c, neg_m)) : 0.0f;
            const float e1 = (ch * 32 + i + 1 < kv) ? ex2_approx(fmaf(__uint_as_float(rr[i + 1]), c, neg_m)) : 0.0f;
            sum0 += e0; sum1 += e1;
            pk[i >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st16(p_addr + ch * 16, pk);
        }
      }
      l_run = fmaf(l_run, alpha, (sum0 + sum1) + ((sA.x + sA.y) + (sB.x + sB.y)));
      AT3_TRS(5, j);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      AT3_TRS(6, j);
    }
    if (n_tiles > 0) consume_o(n_tiles - 1, a_prev);

    // row sum of both halves (same reference maximum on both sides)
    {
      float* xb = xchg + ((n_tiles & 1) * 2 + t) * 2 * 128;
      xb[half * 128 + r] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      l_run += xb[(1 - half) * 128 + r];
    }
    if (q < p.Nq) {
      if (p.splits == 1) {
        const float inv = l_run > 0.0f ? 1.0f / l_run : 0.0f;
        bf16* o = p.o + (long long)b * p.o_sb + (long long)q * p.o_sn + h * HD + half * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(o_acc[8 * i + 0] * inv, o_acc[8 * i + 1] * inv);
          u.y = pack_bf16x2(o_acc[8 * i + 2] * inv, o_acc[8 * i + 3] * inv);
          u.z = pack_bf16x2(o_acc[8 * i + 4] * inv, o_acc[8 * i + 5] * inv);
          u.w = pack_bf16x2(o_acc[8 * i + 6] * inv, o_acc[8 * i + 7] * inv);
          reinterpret_cast<uint4*>(o)[i] = u;
        }
      } else {
        const long long row = ((long long)split * p.B * p.H + bh) * p.Nq + q;
        float* wo = p.ws_o + row * HD + half * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(wo)[i] = make_float4(o_acc[4 * i], o_acc[4 * i + 1], o_acc[4 * i + 2], o_acc[4 * i + 3]);
        if (half == 0) {
          p.ws_ml[row * 2 + 0] = m_run;
          p.ws_ml[row * 2 + 1] = l_run;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
#undef AT3_TRS
#undef AT3_TR

}  // namespace pst3r
