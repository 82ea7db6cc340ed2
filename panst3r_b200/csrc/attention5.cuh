// tcgen05 flash attention for head_dim 64 without a block mask: 256 queries per CTA as two 128-row sub-tiles whose exp sweeps
// alternate on the MUFU, probabilities AND the output accumulator in TENSOR MEMORY.
//
//   S_t = Q_t K_j^T       SS MMA (Q, K from swizzled smem)                    -> TMEM columns [t*128, +128)   fp32
//   P_t = 2^((S_t - m) c) softmax warps: tcgen05.ld -> ex2 -> bf16x2 -> tcgen05.st -> TMEM [384 + t*64, +64)
//   O_t += P_t V_j        TS MMA (A = P from TMEM, B = V from smem, MN-major), accumulating in TMEM [256 + t*64, +64) over all
//                         key tiles; read back once at the end
//
// What the previous structure of this kernel spent its time on (in-kernel timeline + tcgen05.mma rate microbenchmark,
// profiles/r02_attention_timeline.md): one sub-tile step was a single dependent chain
//   S MMA visible (550 clk after the request) -> max pass (2 tcgen05.ld, 580) -> wait for P V of the previous tile (340) ->
//   fold O into registers (tcgen05.ld, 230) -> exp sweep with 2 more tcgen05.ld (1520) -> hand-over (230)  = 3 590 clk per key tile,
// every tcgen05.ld + wait costs ~300 clk under load, every M = 128 tcgen05.mma occupies its issuing thread for >= 80 clk whatever
// its N, and the MUFU (8 clk per warp instruction and sub-partition, 2 048 clk per key tile for both sub-tiles) idled for 40 % of
// the period.  Here:
//   * the tcgen05.ld latencies leave the chain: the maximum pass has both of its loads in flight together; the sweep's loads
//     (16 columns each) fly during the exchange / hand-over waits and behind the exponentials of earlier chunks.  (Holding all
//     64 scores of a thread across both passes does not fit the 96 registers a 640-thread CTA leaves: measured, it spills and
//     runs at 900 instead of 700 us.)
//   * the S buffer is handed back (s_free) as soon as the sweep's last load has landed (after its second chunk), so S_t(j+1) is
//     computed under the second half of sweep j and the other sub-tile's sweep, and the softmax warps hardly wait for S;
//   * the row sums add the probabilities AS ROUNDED to bf16, one mixed-precision add (FHADD.BF16) per element reading the packed
//     register that goes to tensor memory (12 instructions per four scores: 2 FFMA2, 4 MUFU, 2 F2FP, 4 FHADD);
//   * O is never read per tile.  The reference maximum m is raised lazily: only when a tile's maximum exceeds it by more than 2^8
//     are the row sums and the O rows rescaled in place (tcgen05.ld / st behind the P V of the previous tile); P <= 256 in bf16;
//   * the two sub-tiles' sweeps alternate strictly (a sub-tile starts its exponentials when the other has handed its P over):
//     left alone they fall into phase, sweep together at half the MUFU rate each and then wait together;
//   * barrier arrivals are one elected lane per warp (8 per hand-over instead of 256 serialised shared-memory atomics).
// CTA: 640 threads = TMA warp, two MMA-issuing warps (scores / P V), 1 idle warp, 8 + 8 softmax warps: two threads per score
// row (one per 64-key half of the tile and 32 output columns; half-row maxima exchanged through shared memory and a 64-thread
// named barrier).  Measured: render cross-attention (B16 H12 Nq768 Nk12288) 642 us = 722 TFLOP/s (round-1 kernel: 719 us).
#pragma once

namespace pst3r {

constexpr int AT5_THREADS = 640;
constexpr int AT5_STAGES = 4;
constexpr uint32_t AT5_OFF_Q = 0;                                       // 2 x 16 KB
constexpr uint32_t AT5_OFF_K = AT5_OFF_Q + 2 * ATT_ATOM_BYTES;          // 4 x 16 KB
constexpr uint32_t AT5_OFF_V = AT5_OFF_K + AT5_STAGES * ATT_ATOM_BYTES; // 4 x 16 KB
constexpr uint32_t AT5_OFF_BAR = AT5_OFF_V + AT5_STAGES * ATT_ATOM_BYTES;
// barriers: q_full, k_full[ST], v_full[ST], kv_empty[ST], s_full[2], s_free[2], p_full[2], o_full[2], tok[2]
constexpr int AT5_NUM_BARS = 1 + 3 * AT5_STAGES + 10;
constexpr uint32_t AT5_OFF_X = AT5_OFF_BAR + 256;  // fp32 exchange area [2 buffers][2 sub-tiles][2 halves][128 rows]
constexpr uint32_t AT5_DYN_BYTES = AT5_OFF_X + 2 * 2 * 2 * 128 * 4 + 1024;
constexpr uint32_t AT5_TMEM_S = 0;     // S_0 at 0, S_1 at 128 (fp32)
constexpr uint32_t AT5_TMEM_O = 256;   // O_0 at 256, O_1 at 320 (fp32)
constexpr uint32_t AT5_TMEM_P = 384;   // P_0 at 384, P_1 at 448 (bf16 pairs: column c of row r = keys 2c, 2c+1)
constexpr float AT5_LAZY = 8.0f;       // the reference maximum is raised when a tile exceeds it by more than 2^8

// 16 scores (keys base .. base+15 of a thread's 64) -> 8 bf16x2 registers of P in tensor memory, row-sum partials.
// FULL: all 64 keys of the thread are valid (no predication; the four chunks of a tile then form one basic block).
template <bool FULL>
__device__ __forceinline__ void at5_sweep16(const uint32_t (&sc)[16], int base, int kv, float c, float neg_m, float2& sA,
                                            float2& sB, uint32_t taddr) {
  uint32_t pk[8];
  if (FULL) {
    const float2 c2 = make_float2(c, c), nm2 = make_float2(neg_m, neg_m);
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      // packed fp32x2 scale / shift and row-sum adds (sm_100): 2.5 instead of 3.5 issue slots per score
      const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(sc[i]), __uint_as_float(sc[i + 1])), c2, nm2);
      const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(sc[i + 2]), __uint_as_float(sc[i + 3])), c2, nm2);
      const float e0 = ex2_approx(x01.x);
      const float e1 = ex2_approx(x01.y);
      const float e2 = ex2_approx(x23.x);
      const float e3 = ex2_approx(x23.y);
      const uint32_t pa = pack_bf16x2(e0, e1), pb = pack_bf16x2(e2, e3);
      // The row sum adds the probabilities AS ROUNDED to bf16 (what the P V product sees): numerator and denominator of O / l
      // carry the same rounding.  Under the lazy maximum the largest probability of a row is 2^x, x in [0, 8], not exactly 1: with
      // row sums of the unrounded values a row dominated by one key showed that key's 2^-9 rounding (errors up to 6.6e-3 of the
      // largest output instead of 3.6e-3).  The halves of the packed register go straight into the fp32 sums (FHADD.BF16).
      acc_bf16x2(pa, sA.x, sA.y);
      acc_bf16x2(pb, sB.x, sB.y);
      pk[i >> 1] = pa;
      pk[(i >> 1) + 1] = pb;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float e0 = (base + i < kv) ? ex2_approx(fmaf(__uint_as_float(sc[i]), c, neg_m)) : 0.0f;
      const float e1 = (base + i + 1 < kv) ? ex2_approx(fmaf(__uint_as_float(sc[i + 1]), c, neg_m)) : 0.0f;
      const uint32_t pa = pack_bf16x2(e0, e1);
      acc_bf16x2(pa, sA.x, sA.y);
      pk[i >> 1] = pa;
    }
  }
  tmem_st8(taddr, pk);
}

// TRACE: development aid (PST3R_ATT_TRACE=<file>, tools/attn_trace.py): CTA (0,0,0) records clock64() at the hand-over points of the
// MMA thread and of one softmax thread per (sub-tile, half) into p.trace[(slot * 16 + event) * 128 + tile]
#define AT5_TR(slot, ev, j) do { if (TRACE && tr_on && (j) < 128) p.trace[(((slot) * 16 + (ev)) << 7) + (j)] = clock64(); } while (0)

template <bool TRACE>
__global__ void __launch_bounds__(AT5_THREADS, 1)
attention5_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int HD = 64;
  constexpr int ST = AT5_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT5_OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + ST;
  uint64_t* kv_empty = v_full + ST;
  uint64_t* s_full = kv_empty + ST;
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* tok = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AT5_NUM_BARS);

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
      mbar_init(&s_free[i], 8);   // one elected lane of each of the sub-tile's 8 softmax warps
      mbar_init(&p_full[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&tok[i], 8);
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
      tma_load_4d(smem + AT5_OFF_Q, &tmQ, q_full, 0, q_blk * 256, h, b);
      tma_load_4d(smem + AT5_OFF_Q + ATT_ATOM_BYTES, &tmQ, q_full, 0, q_blk * 256 + 128, h, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
        const int key0 = (t0 + j) * ATT_BN;
        mbar_expect_tx(&k_full[s], ATT_ATOM_BYTES);
        tma_load_4d(smem + AT5_OFF_K + s * ATT_ATOM_BYTES, &tmK, &k_full[s], 0, key0, h, kvb);
        mbar_expect_tx(&v_full[s], ATT_ATOM_BYTES);
        tma_load_4d(smem + AT5_OFF_V + s * ATT_ATOM_BYTES, &tmV, &v_full[s], 0, key0, h, kvb);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer 1: scores ------------------------
    // Two issuing threads: one thread gets a tcgen05.mma out every ~80 cycles whatever its shape, two threads together twice that
    // (tools/mma_rate.cu); the 8 S + 16 P V instructions of a key tile would keep a single issuer busy for ~2 400 cycles.
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      const uint32_t q_addr = smem_u32(smem + AT5_OFF_Q);
      auto issue_s = [&](int t, int j) {  // caller has waited for k_full(j) and knows S_t is free
        const uint32_t k_addr = smem_u32(smem + AT5_OFF_K + (j % ST) * ATT_ATOM_BYTES);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_ss(tmem_base + AT5_TMEM_S + t * ATT_BN, make_smem_desc_sw128(q_addr + t * ATT_ATOM_BYTES + ks * 32, 0, 1024),
                  make_smem_desc_sw128(k_addr + ks * 32, 0, 1024), idesc_s, ks != 0);
        umma_commit(&s_full[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j + 1 < n_tiles; ++j) {
        mbar_wait(&k_full[(j + 1) % ST], ((j + 1) / ST) & 1);
        AT5_TR(0, 0, j);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          // s_free(t, j): the sweep of sub-tile t has read the last of S_t(j)
          mbar_wait(&s_free[t], j & 1);
          tc_fence_after();
          issue_s(t, j + 1);
          AT5_TR(0, 1 + 3 * t, j);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------- MMA issuer 2: P V ---------------------------
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, HD, 0, 1);  // A = P (TMEM, K-major), B = V (MN-major)
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        mbar_wait(&v_full[s], (j / ST) & 1);
        const uint32_t v_addr = smem_u32(smem + AT5_OFF_V + s * ATT_ATOM_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          // p_full(t, j): P_t(j) is in TMEM and O_t has been rescaled if it had to be
          mbar_wait(&p_full[t], j & 1);
          AT5_TR(0, 2 + 3 * t, j);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < ATT_BN / 16; ++ks)  // 16 keys per step = 8 TMEM columns of P, 16 k-rows (2 KB) of V
            umma_ts(tmem_base + AT5_TMEM_O + t * HD, tmem_base + AT5_TMEM_P + t * (ATT_BN / 2) + ks * 8,
                    make_smem_desc_sw128(v_addr + ks * 16 * 128, ATT_ATOM_BYTES, 1024), idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&o_full[t]);
          AT5_TR(0, 3 + 3 * t, j);
        }
        // both sub-tiles have handed P(j) over, so their S(j) (other issuer) had been read: K_j and V_j may be overwritten
        // once these P V have completed
        umma_commit(&kv_empty[s]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- softmax ------------------------------------
    const int t = (warp - 4) >> 3;           // sub-tile
    const int quad = warp & 3;               // TMEM lane quadrant
    const int half = ((warp - 4) >> 2) & 1;  // which 64 keys of a tile / which 32 output columns
    const int r = quad * 32 + lane;
    const int q = q_blk * 256 + t * 128 + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t s_addr = lane_addr + AT5_TMEM_S + t * ATT_BN + half * 64;
    const uint32_t o_addr = lane_addr + AT5_TMEM_O + t * HD + half * 32;
    const uint32_t p_addr = lane_addr + AT5_TMEM_P + t * (ATT_BN / 2) + half * 32;
    float* xchg = reinterpret_cast<float*>(smem + AT5_OFF_X);
    const int bar_id = 1 + t * 4 + quad;   // named barrier of the two warps that share these 32 rows
    const int tr_slot = 1 + t + 2 * half;
    const bool tr_me = quad == 0 && lane == 0;
#define AT5_TRS(ev, j) do { if (tr_me) AT5_TR(tr_slot, ev, j); } while (0)
    float m_run = -CUDART_INF_F, l_run = 0.0f;
    const float c = p.scale_log2;

    float2 sA, sB;  // row-sum partials of the current tile
    for (int j = 0; j < n_tiles; ++j) {
      const int kv = p.Nk - (t0 + j) * ATT_BN - half * 64;  // valid keys among this thread's 64 (may exceed 64 / be <= 0)
      mbar_wait_relaxed(&s_full[t], j & 1);
      AT5_TRS(0, j);
      tc_fence_after();
      // pass 1: the half-row maximum, both 32-column loads in flight together
      float mx = -CUDART_INF_F;
      {
        uint32_t ra[32], rb[32];
        tmem_ld32(s_addr, ra);
        tmem_ld32(s_addr + 32, rb);
        tmem_ld_wait();
        if (kv >= 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(ra[i]), __uint_as_float(rb[i])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < kv) mx = fmaxf(mx, __uint_as_float(ra[i]));
            if (32 + i < kv) mx = fmaxf(mx, __uint_as_float(rb[i]));
          }
        }
      }
      AT5_TRS(1, j);
      // pass 2 reads the scores again 16 columns at a time: three loads fly during the exchange and the waits below, the
      // fourth behind the exponentials of the second and third chunk, so that the sweep itself never waits for tensor memory
      uint32_t c0[16], c1[16], c2[16];
      tmem_ld16(s_addr, c0);
      tmem_ld16(s_addr + 16, c1);
      tmem_ld16(s_addr + 32, c2);
      // the other half of the row: exchange the half-row maxima (double buffered across tiles)
      float* xb = xchg + (((j & 1) * 2 + t) * 2) * 128;
      xb[half * 128 + r] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      mx = fmaxf(mx, xb[(1 - half) * 128 + r]);
      AT5_TRS(2, j);
      // lazily raised reference maximum (identical in the two threads of a row)
      bool raise = false;
      if (m_run == -CUDART_INF_F) {
        m_run = mx;  // nothing accumulated yet
      } else if ((mx - m_run) * c > AT5_LAZY) {
        raise = true;
      }
      // P_t(j) overwrites P_t(j-1), a rescale touches O_t: both behind the P V of the previous tile
      if (j > 0) mbar_wait_relaxed(&o_full[t], (j - 1) & 1);
      tc_fence_after();
      AT5_TRS(3, j);
      if (__any_sync(0xffffffffu, raise)) {
        const float alpha = raise ? ex2_approx((m_run - mx) * c) : 1.0f;
        if (raise) m_run = mx;
        l_run *= alpha;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t oo[16];
          tmem_ld16(o_addr + hh * 16, oo);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) oo[i] = __float_as_uint(__uint_as_float(oo[i]) * alpha);
          tmem_st16(o_addr + hh * 16, oo);
        }
      }
      const float m_use = (m_run == -CUDART_INF_F) ? 0.0f : m_run;
      const float neg_m = -m_use * c;
      // strict alternation of the two sub-tiles' sweeps: 0:j, 1:j, 0:j+1, ...
      if (t == 1) mbar_wait_relaxed(&tok[0], j & 1);
      else if (j > 0) mbar_wait_relaxed(&tok[1], (j - 1) & 1);
      AT5_TRS(4, j);
      sA = make_float2(0.0f, 0.0f); sB = make_float2(0.0f, 0.0f);
      tmem_ld_wait();
      if (kv >= 64) {
        at5_sweep16<true>(c0, 0, kv, c, neg_m, sA, sB, p_addr);
        tmem_ld16(s_addr + 48, c0);
        at5_sweep16<true>(c1, 16, kv, c, neg_m, sA, sB, p_addr + 8);
        tmem_ld_wait();
        // all scores of this tile have been read: S_t may take the next tile while the second half is swept
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        at5_sweep16<true>(c2, 32, kv, c, neg_m, sA, sB, p_addr + 16);
        at5_sweep16<true>(c0, 48, kv, c, neg_m, sA, sB, p_addr + 24);
      } else {
        at5_sweep16<false>(c0, 0, kv, c, neg_m, sA, sB, p_addr);
        tmem_ld16(s_addr + 48, c0);
        at5_sweep16<false>(c1, 16, kv, c, neg_m, sA, sB, p_addr + 8);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        at5_sweep16<false>(c2, 32, kv, c, neg_m, sA, sB, p_addr + 16);
        at5_sweep16<false>(c0, 48, kv, c, neg_m, sA, sB, p_addr + 24);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tok[t]);
      l_run += (sA.x + sA.y) + (sB.x + sB.y);
      AT5_TRS(5, j);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      AT5_TRS(6, j);
    }

    // row sum of both halves (same reference maximum on both sides), then this thread's 32 output columns
    {
      float* xb = xchg + ((n_tiles & 1) * 2 + t) * 2 * 128;
      xb[half * 128 + r] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      l_run += xb[(1 - half) * 128 + r];
    }
    float o_acc[32];
    if (n_tiles > 0) {
      mbar_wait(&o_full[t], (n_tiles - 1) & 1);
      tc_fence_after();
      uint32_t oo[32];
      tmem_ld32(o_addr, oo);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] = __uint_as_float(oo[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] = 0.0f;
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
#undef AT5_TRS
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
#undef AT5_TR

}  // namespace pst3r
