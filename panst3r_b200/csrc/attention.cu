// tcgen05 flash attention (forward, non-causal), head_dim 64 / 96, optional per-(query,key) block mask,
// optional split-KV.  O = softmax(scale * Q K^T + mask) V.
//
// Replaces xformers.memory_efficient_attention / croco Attention+CrossAttention / nn.MultiheadAttention on the
// PanSt3R forward path: encoder + DINOv2 self-attention (hd 64), MUSt3R decoder self- and cross-attention over
// keyframe memory tokens (hd 64, up to 49 152 keys), LoftUp cross-attention (hd 96), query-decoder masked
// cross-attention and self-attention (hd 96).
//
// One CTA = one 128-query tile of one (batch, head) [x one KV split]; 192 threads:
//   warp 0     : TMA producer (Q once; K and V tiles of 128 keys through a KV_STAGES ring, SWIZZLE_128B)
//   warp 1     : TMEM allocator + single-thread tcgen05.mma issuer
//                  S[b]  = Q K_j^T      (SS, both K-major, fp32 128x128 in TMEM, double buffered)
//                  Ot[b] = P_j V_j      (SS, P K-major from smem, V MN-major, fp32 128xHD in TMEM, double buffered)
//   warps 2..5 : softmax + accumulate; thread r owns query row r: two passes over S in TMEM
//                (row max, then exp2 / row sum / bf16 P -> swizzled smem), then O_reg = (O_reg + Ot_{j-1}) * alpha_j
//   The S MMA of tile j+1 is issued before the PV MMA of tile j so the tensor core overlaps the softmax.
#include "common.cuh"
#include "host_util.h"
#include "../../include/panst3r_b200.h"

#include <math_constants.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

namespace pst3r {

constexpr int ATT_BM = 128;
constexpr int ATT_BN = 128;
constexpr int ATT_THREADS = 192;
constexpr uint32_t ATT_ATOM_BYTES = 128 * 128;  // 128 rows x 128 B (64 bf16)

template <int HD>
struct AttnCfg {
  static constexpr int KATOMS = (HD + 63) / 64;
  static constexpr uint32_t Q_BYTES = KATOMS * ATT_ATOM_BYTES;
  static constexpr uint32_t K_BYTES = Q_BYTES;
  static constexpr uint32_t V_BYTES = Q_BYTES;
  static constexpr int KV_STAGES = (HD == 64) ? 3 : 2;
  static constexpr uint32_t P_BYTES = 2 * ATT_ATOM_BYTES;
  static constexpr uint32_t OFF_Q = 0;
  static constexpr uint32_t OFF_K = OFF_Q + Q_BYTES;
  static constexpr uint32_t OFF_V = OFF_K + KV_STAGES * K_BYTES;
  static constexpr uint32_t OFF_P = OFF_V + KV_STAGES * V_BYTES;
  static constexpr uint32_t OFF_BAR = OFF_P + 2 * P_BYTES;
  // barriers: q_full, k_full[ST], v_full[ST], kv_empty[ST], s_full[2], s_empty[2], p_full[2], o_full[2], o_empty[2]
  static constexpr int NUM_BARS = 1 + 3 * KV_STAGES + 10;
  static constexpr uint32_t TOTAL = OFF_BAR + NUM_BARS * 8 + 16;
  static constexpr uint32_t DYN_BYTES = TOTAL + 1024;
  static constexpr uint32_t TMEM_S = 0;        // 2 x 128 columns
  static constexpr uint32_t TMEM_O = 256;      // 2 x HD columns
  static constexpr uint32_t TMEM_COLS = 512;
};

struct AttnParams {
  bf16* o;
  long long o_sb, o_sn;
  int B, H, Nq, Nk;
  float scale_log2;  // scale * log2(e)
  const uint32_t* mask_bits;
  long long mask_sb, mask_sq;
  int kv_shared;  // K/V batch stride 0: always read batch 0
  int splits;
  float* ws_o;   // [splits][B*H][Nq][HD] fp32 (unnormalised)
  float* ws_ml;  // [splits][B*H][Nq][2]   (m in raw-score units, l)
  long long* trace;  // development aid (attention5.cuh, tools/attn_trace.py), normally null
};

template <int HD, bool HAS_MASK>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = AttnCfg<HD>;
  constexpr int ST = C::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + ST;
  uint64_t* kv_empty = v_full + ST;
  uint64_t* s_full = kv_empty + ST;
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh - b * p.H;
  const int split = blockIdx.z;
  const int kvb = p.kv_shared ? 0 : b;

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
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
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
      mbar_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int a = 0; a < C::KATOMS; ++a)
        tma_load_4d(smem + C::OFF_Q + a * ATT_ATOM_BYTES, &tmQ, q_full, a * 64, q_tile * ATT_BM, h, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        const uint32_t ph = (j / ST) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        const int key0 = (t0 + j) * ATT_BN;
        mbar_expect_tx(&k_full[s], C::K_BYTES);
#pragma unroll
        for (int a = 0; a < C::KATOMS; ++a)
          tma_load_4d(smem + C::OFF_K + s * C::K_BYTES + a * ATT_ATOM_BYTES, &tmK, &k_full[s], a * 64, key0, h, kvb);
        mbar_expect_tx(&v_full[s], C::V_BYTES);
#pragma unroll
        for (int a = 0; a < C::KATOMS; ++a)
          tma_load_4d(smem + C::OFF_V + s * C::V_BYTES + a * ATT_ATOM_BYTES, &tmV, &v_full[s], a * 64, key0, h, kvb);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, HD, 0, 1);  // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(smem + C::OFF_Q);
      auto issue_s = [&](int j) {
        const int s = j % ST;
        const int sb = j & 1;
        mbar_wait(&k_full[s], (j / ST) & 1);
        mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + C::OFF_K + s * C::K_BYTES);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * ATT_ATOM_BYTES + (ks & 3) * 32;
          umma_ss(tmem_base + C::TMEM_S + sb * ATT_BN, make_smem_desc_sw128(q_addr + off, 0, 1024),
                  make_smem_desc_sw128(k_addr + off, 0, 1024), idesc_s, ks != 0);
        }
        umma_commit(&s_full[sb]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);
        const int s = j % ST;
        const int ob = j & 1;
        mbar_wait(&p_full[ob], (j >> 1) & 1);
        mbar_wait(&v_full[s], (j / ST) & 1);
        mbar_wait(&o_empty[ob], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(smem + C::OFF_P + ob * C::P_BYTES);
        const uint32_t v_addr = smem_u32(smem + C::OFF_V + s * C::V_BYTES);
#pragma unroll
        for (int ks = 0; ks < ATT_BN / 16; ++ks) {
          const uint32_t p_off = (ks >> 2) * ATT_ATOM_BYTES + (ks & 3) * 32;
          const uint32_t v_off = ks * 16 * 128;  // 16 keys x 128 B rows
          umma_ss(tmem_base + C::TMEM_O + ob * HD, make_smem_desc_sw128(p_addr + p_off, 0, 1024),
                  make_smem_desc_sw128(v_addr + v_off, ATT_ATOM_BYTES, 1024), idesc_pv, ks != 0);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(&o_full[ob]);
      }
    }
  } else {
    // ------------------------------- softmax / accumulate -----------------------
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // row within the tile == TMEM lane
    const int q = q_tile * ATT_BM + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float o_acc[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o_acc[i] = 0.0f;
    float m_run = -CUDART_INF_F;
    float l_run = 0.0f;
    const uint32_t* mrow = nullptr;
    if (HAS_MASK) mrow = p.mask_bits + (long long)b * p.mask_sb + (long long)min(q, p.Nq - 1) * p.mask_sq;

    auto consume_o = [&](int j, float alpha) {
      const int ob = j & 1;
      mbar_wait(&o_full[ob], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t rr[32];
        tmem_ld32(lane_addr + C::TMEM_O + ob * HD + c * 32, rr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = (o_acc[c * 32 + i] + __uint_as_float(rr[i])) * alpha;
      }
      tc_fence_before();
      mbar_arrive(&o_empty[ob]);
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j & 1;
      const int key0 = (t0 + j) * ATT_BN;
      const int kmax = p.Nk - key0;  // number of valid keys in this tile (may exceed 128)
      uint32_t mw[4] = {0u, 0u, 0u, 0u};
      if (HAS_MASK) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(mrow + (t0 + j) * 4));
        mw[0] = w.x; mw[1] = w.y; mw[2] = w.z; mw[3] = w.w;
      }
      if (kmax < ATT_BN) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int lo = kmax - c * 32;
          mw[c] |= (lo >= 32) ? 0u : (lo <= 0 ? 0xffffffffu : (0xffffffffu << lo));
        }
      }
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = lane_addr + C::TMEM_S + sb * ATT_BN;
      // pass 1: row max
      float mx = m_run;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t rr[32];
        tmem_ld32(s_addr + c * 32, rr);
        tmem_ld_wait();
        const uint32_t w = mw[c];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = ((w >> i) & 1u) ? -CUDART_INF_F : __uint_as_float(rr[i]);
          mx = fmaxf(mx, sv);
        }
      }
      const float m_use = (mx == -CUDART_INF_F) ? 0.0f : mx;
      const float alpha = (m_run == -CUDART_INF_F) ? 0.0f : exp2f((m_run - m_use) * p.scale_log2);
      const float neg_m = -m_use * p.scale_log2;
      m_run = mx;
      // pass 2: p = exp2(s*c - m*c); row sum; bf16 P into the swizzled K-major smem tile
      float rs = 0.0f;
      uint8_t* p_tile = smem + C::OFF_P + sb * C::P_BYTES;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t rr[32];
        tmem_ld32(s_addr + c * 32, rr);
        tmem_ld_wait();
        const uint32_t w = mw[c];
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0 = exp2f(fmaf(__uint_as_float(rr[i]), p.scale_log2, neg_m));
          float e1 = exp2f(fmaf(__uint_as_float(rr[i + 1]), p.scale_log2, neg_m));
          e0 = ((w >> i) & 1u) ? 0.0f : e0;
          e1 = ((w >> (i + 1)) & 1u) ? 0.0f : e1;
          // the row sum uses the bf16-rounded probabilities the tensor core will actually consume
          const uint32_t u = pack_bf16x2(e0, e1);
          const float2 f = unpack_bf16x2(u);
          rs += f.x + f.y;
          pk[i >> 1] = u;
        }
        // keys [c*32, c*32+32) -> atom (c>>1), 16-byte chunks ((c&1)*4 .. +3) of row r, XOR-swizzled with (r & 7)
        uint8_t* row_base = p_tile + (c >> 1) * ATT_ATOM_BYTES + r * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = ((c & 1) * 4 + ch) ^ (r & 7);
          *reinterpret_cast<uint4*>(row_base + chunk * 16) =
              make_uint4(pk[ch * 4 + 0], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
      }
      l_run = l_run * alpha + rs;
      tc_fence_before();
      mbar_arrive(&s_empty[sb]);
      fence_proxy_async_smem();
      mbar_arrive(&p_full[sb]);
      if (j > 0) consume_o(j - 1, alpha);
    }
    if (n_tiles > 0) consume_o(n_tiles - 1, 1.0f);

    if (q < p.Nq) {
      if (p.splits == 1) {
        const float inv = l_run > 0.0f ? 1.0f / l_run : 0.0f;
        bf16* o = p.o + (long long)b * p.o_sb + (long long)q * p.o_sn + h * HD;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
          uint4 u;
          u.x = pack_bf16x2(o_acc[8 * i + 0] * inv, o_acc[8 * i + 1] * inv);
          u.y = pack_bf16x2(o_acc[8 * i + 2] * inv, o_acc[8 * i + 3] * inv);
          u.z = pack_bf16x2(o_acc[8 * i + 4] * inv, o_acc[8 * i + 5] * inv);
          u.w = pack_bf16x2(o_acc[8 * i + 6] * inv, o_acc[8 * i + 7] * inv);
          reinterpret_cast<uint4*>(o)[i] = u;
        }
      } else {
        const long long row = ((long long)split * p.B * p.H + bh) * p.Nq + q;
        float* wo = p.ws_o + row * HD;
#pragma unroll
        for (int i = 0; i < HD / 4; ++i)
          reinterpret_cast<float4*>(wo)[i] =
              make_float4(o_acc[4 * i], o_acc[4 * i + 1], o_acc[4 * i + 2], o_acc[4 * i + 3]);
        p.ws_ml[row * 2 + 0] = m_run;
        p.ws_ml[row * 2 + 1] = l_run;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// Merge split-KV partials: one warp per (b, h, q) row.
template <int HD>
__global__ void attention_combine_kernel(const float* __restrict__ ws_o, const float* __restrict__ ws_ml, bf16* o,
                                         long long o_sb, long long o_sn, int B, int H, int Nq, int splits,
                                         float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  const int warps_per_block = blockDim.x >> 5;
  const long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const long long rows = (long long)B * H * Nq;
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int q = row % Nq;
  const int bh = row / Nq;
  const int b = bh / H, h = bh % H;
  float mmax = -CUDART_INF_F;
  for (int s = 0; s < splits; ++s) mmax = fmaxf(mmax, ws_ml[((long long)s * rows + row) * 2]);
  float lsum = 0.0f;
  float acc[(HD + 31) / 32];
#pragma unroll
  for (int i = 0; i < (HD + 31) / 32; ++i) acc[i] = 0.0f;
  for (int s = 0; s < splits; ++s) {
    const long long rr = (long long)s * rows + row;
    const float m = ws_ml[rr * 2];
    const float l = ws_ml[rr * 2 + 1];
    const float w = (m == -CUDART_INF_F) ? 0.0f : exp2f((m - mmax) * scale_log2);
    lsum += w * l;
#pragma unroll
    for (int i = 0; i < (HD + 31) / 32; ++i) {
      const int d = i * 32 + lane;
      if (d < HD) acc[i] += w * ws_o[rr * HD + d];
    }
  }
  const float inv = lsum > 0.0f ? 1.0f / lsum : 0.0f;
  bf16* op = o + (long long)b * o_sb + (long long)q * o_sn + h * HD;
#pragma unroll
  for (int i = 0; i < (HD + 31) / 32; ++i) {
    const int d = i * 32 + lane;
    if (d < HD) op[d] = __float2bfloat16(acc[i] * inv);
  }
}

// One query per (batch, head) on the CUDA cores: DINOv2's CLS token (model/dino.py:59-71 keeps it through all 24
// layers and drops it at the end).  769 = 3 x 256 + 1 tokens would otherwise cost a fourth 256-query CTA per (batch, head)
// for a single row — a quarter of the attention work of every DINOv2 layer.  One block per (b, h); thread per key for the
// scores, then 4 key groups x 64 value columns.
__global__ void __launch_bounds__(256)
attention_q1_kernel(const bf16* __restrict__ q, long long q_sb, long long q_sh, const bf16* __restrict__ k, long long k_sb,
                    long long k_sn, long long k_sh, const bf16* __restrict__ v, long long v_sb, long long v_sn, long long v_sh,
                    bf16* __restrict__ o, long long o_sb, int H, int Nk, float scale_log2) {
  constexpr int HD = 64;
  extern __shared__ float sm[];  // [Nk] probabilities | [4][64] partial outputs | [16] reductions
  float* prob = sm;
  float* part = sm + ((Nk + 3) & ~3);
  float* red = part + 4 * HD;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float qr[HD];
  {
    const uint4* q4 = reinterpret_cast<const uint4*>(q + b * q_sb + h * q_sh);
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const uint4 u = __ldg(q4 + i);
      float2 f;
      f = unpack_bf16x2(u.x); qr[8 * i + 0] = f.x; qr[8 * i + 1] = f.y;
      f = unpack_bf16x2(u.y); qr[8 * i + 2] = f.x; qr[8 * i + 3] = f.y;
      f = unpack_bf16x2(u.z); qr[8 * i + 4] = f.x; qr[8 * i + 5] = f.y;
      f = unpack_bf16x2(u.w); qr[8 * i + 6] = f.x; qr[8 * i + 7] = f.y;
    }
  }
  float mx = -CUDART_INF_F;
  for (int key = tid; key < Nk; key += 256) {
    const uint4* k4 = reinterpret_cast<const uint4*>(k + b * k_sb + (long long)key * k_sn + h * k_sh);
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const uint4 u = __ldg(k4 + i);
      float2 f;
      f = unpack_bf16x2(u.x); s = fmaf(qr[8 * i + 0], f.x, fmaf(qr[8 * i + 1], f.y, s));
      f = unpack_bf16x2(u.y); s = fmaf(qr[8 * i + 2], f.x, fmaf(qr[8 * i + 3], f.y, s));
      f = unpack_bf16x2(u.z); s = fmaf(qr[8 * i + 4], f.x, fmaf(qr[8 * i + 5], f.y, s));
      f = unpack_bf16x2(u.w); s = fmaf(qr[8 * i + 6], f.x, fmaf(qr[8 * i + 7], f.y, s));
    }
    prob[key] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.0f;
  for (int key = tid; key < Nk; key += 256) {
    const float e = ex2_approx((prob[key] - mx) * scale_log2);
    prob[key] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[8 + warp] = sum;
  __syncthreads();
  float tot = 0.0f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[8 + w];
  // output: thread = (key group g, column d); 4 groups stride the keys
  const int d = tid & 63, g = tid >> 6;
  float acc = 0.0f;
  const bf16* vb = v + b * v_sb + h * v_sh + d;
  for (int key = g; key < Nk; key += 4) acc = fmaf(prob[key], __bfloat162float(vb[(long long)key * v_sn]), acc);
  part[g * HD + d] = acc;
  __syncthreads();
  if (tid < HD) {
    const float r = (part[tid] + part[HD + tid]) + (part[2 * HD + tid] + part[3 * HD + tid]);
    o[b * o_sb + h * HD + tid] = __float2bfloat16(tot > 0.0f ? r / tot : 0.0f);
  }
}

}  // namespace pst3r
#include "attention5.cuh"
namespace pst3r {

static int make_qkv_map(CUtensorMap* m, const void* ptr, int hd, long long n, int H, int B, long long sn,
                        long long sh, long long sb) {
  uint64_t dims[4] = {(uint64_t)hd, (uint64_t)n, (uint64_t)H, (uint64_t)B};
  uint64_t str[4] = {2, (uint64_t)sn * 2, (uint64_t)sh * 2, (uint64_t)(sb ? sb : sn * n) * 2};
  uint32_t box[4] = {64, 128, 1, 1};
  return encode_tmap(m, ptr, 2, 4, dims, str, box);
}

template <int HD>
static int launch_attention(const pst3r_attn_args* a, int splits, cudaStream_t stream) {
  using C = AttnCfg<HD>;
  CUtensorMap tmQ, tmK, tmV;
  int r;
  const int kv_shared = (a->k_sb == 0) ? 1 : 0;
  if ((r = make_qkv_map(&tmQ, a->q, HD, a->Nq, a->H, a->B, a->q_sn, a->q_sh, a->q_sb))) return r;
  if ((r = make_qkv_map(&tmK, a->k, HD, a->Nk, a->H, kv_shared ? 1 : a->B, a->k_sn, a->k_sh, a->k_sb))) return r;
  if ((r = make_qkv_map(&tmV, a->v, HD, a->Nk, a->H, kv_shared ? 1 : a->B, a->v_sn, a->v_sh, a->v_sb))) return r;

  AttnParams p;
  p.o = reinterpret_cast<bf16*>(a->o);
  p.o_sb = a->o_sb; p.o_sn = a->o_sn;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.mask_bits = a->mask_bits; p.mask_sb = a->mask_sb; p.mask_sq = a->mask_sq;
  p.kv_shared = kv_shared;
  p.splits = splits;
  p.ws_o = nullptr; p.ws_ml = nullptr;
  p.trace = nullptr;
  if (splits > 1) {
    const long long rows = (long long)a->B * a->H * a->Nq;
    p.ws_o = reinterpret_cast<float*>(a->workspace);
    p.ws_ml = p.ws_o + (long long)splits * rows * HD;
  }
  dim3 grid((a->Nq + ATT_BM - 1) / ATT_BM, a->B * a->H, splits);
  if (HD == 64 && !a->mask_bits) {
    // 256 queries per CTA, probabilities and output accumulator in tensor memory (attention5.cuh)
    static bool cfg5 = false;
    if (!cfg5) {
      PST3R_CHECK_CUDA(cudaFuncSetAttribute(attention5_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT5_DYN_BYTES));
      PST3R_CHECK_CUDA(cudaFuncSetAttribute(attention5_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT5_DYN_BYTES));
      cfg5 = true;
    }
    const dim3 grid2((a->Nq + 255) / 256, a->B * a->H, splits);
    static const char* trace_path = getenv("PST3R_ATT_TRACE");
    if (trace_path && a->Nk >= 4096) {
      // development aid: one traced launch, synchronous, timeline of CTA (0,0,0) written as text (tools/attn_trace.py reads it)
      constexpr int TR_N = 5 * 16 * 128;
      long long* dtr = nullptr;
      PST3R_CHECK_CUDA(cudaMalloc(&dtr, TR_N * sizeof(long long)));
      PST3R_CHECK_CUDA(cudaMemsetAsync(dtr, 0, TR_N * sizeof(long long), stream));
      p.trace = dtr;
      PST3R_CHECK_CUDA(launch_pdl(attention5_fwd_kernel<true>, grid2, dim3(AT5_THREADS), AT5_DYN_BYTES, stream, tmQ, tmK, tmV, p));
      PST3R_CHECK_CUDA(cudaStreamSynchronize(stream));
      std::vector<long long> h(TR_N);
      PST3R_CHECK_CUDA(cudaMemcpy(h.data(), dtr, TR_N * sizeof(long long), cudaMemcpyDeviceToHost));
      PST3R_CHECK_CUDA(cudaFree(dtr));
      if (FILE* f = fopen(trace_path, "w")) {
        for (int sl = 0; sl < 5; ++sl)
          for (int ev = 0; ev < 16; ++ev)
            for (int j = 0; j < 128; ++j)
              if (h[((sl * 16 + ev) << 7) + j]) fprintf(f, "%d %d %d %lld\n", sl, ev, j, h[((sl * 16 + ev) << 7) + j]);
        fclose(f);
      }
      p.trace = nullptr;
    } else {
      PST3R_CHECK_CUDA(launch_pdl(attention5_fwd_kernel<false>, grid2, dim3(AT5_THREADS), AT5_DYN_BYTES, stream, tmQ, tmK, tmV, p));
    }
  } else if (a->mask_bits) {
    auto kern = attention_fwd_kernel<HD, true>;
    static bool cfg = false;
    if (!cfg) { PST3R_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::DYN_BYTES)); cfg = true; }
    PST3R_CHECK_CUDA(launch_pdl(kern, grid, dim3(ATT_THREADS), C::DYN_BYTES, stream, tmQ, tmK, tmV, p));
  } else {
    auto kern = attention_fwd_kernel<HD, false>;
    static bool cfg = false;
    if (!cfg) { PST3R_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::DYN_BYTES)); cfg = true; }
    PST3R_CHECK_CUDA(launch_pdl(kern, grid, dim3(ATT_THREADS), C::DYN_BYTES, stream, tmQ, tmK, tmV, p));
  }
  PST3R_CHECK_CUDA(cudaGetLastError());
  if (splits > 1) {
    const long long rows = (long long)a->B * a->H * a->Nq;
    const int wpb = 8;
    const unsigned blocks = (unsigned)((rows + wpb - 1) / wpb);
    PST3R_CHECK_CUDA(launch_pdl(attention_combine_kernel<HD>, dim3(blocks), dim3(wpb * 32), 0, stream, p.ws_o, p.ws_ml, p.o,
                                p.o_sb, p.o_sn, a->B, a->H, a->Nq, splits, p.scale_log2));
  }
  return PST3R_OK;
}

}  // namespace pst3r

using namespace pst3r;

extern "C" int32_t pst3r_attention_auto_splits(int32_t B, int32_t H, int32_t Nq, int32_t Nk) {
  const int ctas = ((Nq + 255) / 256) * B * H;  // CTAs of the 256-query kernel (a slight over-split for the 128-query one)
  const int tiles = (Nk + ATT_BN - 1) / ATT_BN;
  const int sms = sm_budget();
  if (ctas >= sms || tiles <= 1) return 1;
  int s = sms / ctas;
  if (s > tiles) s = tiles;
  if (s > 16) s = 16;
  return s < 1 ? 1 : s;
}

extern "C" int64_t pst3r_attention_workspace_bytes(int32_t B, int32_t H, int32_t Nq, int32_t head_dim,
                                                   int32_t kv_splits) {
  if (kv_splits <= 1) return 0;
  const long long rows = (long long)B * H * Nq;
  return (long long)kv_splits * rows * (head_dim + 2) * 4;
}

extern "C" int pst3r_attention(const pst3r_attn_args* a, pst3r_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PST3R_CHECK_ARG(a && a->q && a->k && a->v && a->o, "attention: null pointer");
  PST3R_CHECK_ARG(a->head_dim == 64 || a->head_dim == 96, "attention: head_dim %d unsupported (64 or 96)",
                  a->head_dim);
  PST3R_CHECK_ARG(a->B > 0 && a->H > 0 && a->Nq > 0 && a->Nk > 0, "attention: bad shape");
  PST3R_CHECK_ARG((a->q_sn % 8) == 0 && (a->k_sn % 8) == 0 && (a->v_sn % 8) == 0 && (a->q_sh % 8) == 0 &&
                      (a->k_sh % 8) == 0 && (a->v_sh % 8) == 0 && (a->q_sb % 8) == 0 && (a->k_sb % 8) == 0 &&
                      (a->v_sb % 8) == 0,
                  "attention: q/k/v strides must be multiples of 8 elements");
  PST3R_CHECK_ARG((a->k_sb == 0) == (a->v_sb == 0), "attention: k and v must both be batch-shared or neither");
  PST3R_CHECK_ARG((a->o_sn % 8) == 0 && (a->o_sb % 8) == 0, "attention: o strides must be multiples of 8");
  if (a->mask_bits)
    PST3R_CHECK_ARG((a->mask_sq % 4) == 0 && (a->mask_sq == 0 || a->mask_sq * 32 >= ((a->Nk + 127) / 128) * 128) &&
                        (reinterpret_cast<uintptr_t>(a->mask_bits) % 16) == 0 && (a->mask_sb % 4) == 0,
                    "attention: mask rows must be 16-byte aligned and padded to 128 keys");
  int splits = a->kv_splits > 0 ? a->kv_splits : pst3r_attention_auto_splits(a->B, a->H, a->Nq, a->Nk);
  const int tiles = (a->Nk + ATT_BN - 1) / ATT_BN;
  if (splits > tiles) splits = tiles;
  // avoid empty trailing splits
  { const int tps = (tiles + splits - 1) / splits; splits = (tiles + tps - 1) / tps; }
  if (splits > 1) {
    const int64_t need = pst3r_attention_workspace_bytes(a->B, a->H, a->Nq, a->head_dim, splits);
    PST3R_CHECK_ARG(a->workspace && a->workspace_bytes >= need,
                    "attention: workspace too small (%lld < %lld bytes for %d splits)",
                    (long long)a->workspace_bytes, (long long)need, splits);
  }
  if (a->head_dim == 64 && a->Nq == 1 && !a->mask_bits && a->Nk <= 8192) {
    // a single query per (batch, head): CUDA-core kernel (DINOv2's CLS token)
    const size_t smem = (((size_t)a->Nk + 3) & ~(size_t)3) * 4 + (4 * 64 + 16) * 4;
    PST3R_CHECK_CUDA(launch_pdl(attention_q1_kernel, dim3(a->B * a->H), dim3(256), smem, stream,
                                reinterpret_cast<const bf16*>(a->q), (long long)a->q_sb, (long long)a->q_sh,
                                reinterpret_cast<const bf16*>(a->k), (long long)a->k_sb, (long long)a->k_sn, (long long)a->k_sh,
                                reinterpret_cast<const bf16*>(a->v), (long long)a->v_sb, (long long)a->v_sn, (long long)a->v_sh,
                                reinterpret_cast<bf16*>(a->o), (long long)a->o_sb, (int)a->H, (int)a->Nk,
                                a->scale * 1.4426950408889634f));
    return PST3R_OK;
  }
  if (a->head_dim == 64) return launch_attention<64>(a, splits, stream);
  return launch_attention<96>(a, splits, stream);
}
