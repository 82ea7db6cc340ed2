// tcgen05 / TMA bf16 GEMM with fused epilogues:  C[M,N] = epi(A[M,K] * B[N,K]^T)
//
// Replaces every nn.Linear-shaped contraction on the PanSt3R forward path (QKV / proj / MLP of the
// encoder, decoder, DINOv2, InputMixer; the v1 PixelShuffle MLP chain; the mask-logit einsum; the
// pointmap head).  Fused epilogues: bias, GELU/ReLU, LayerScale, residual add, 2-D RoPE on q/k columns,
// pixel_shuffle(2) store, 16x16 depth-to-space store, transposed (plane-major) fp32 store.
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0      : TMA producer   (A 128x64 + B BNx64 bf16 tiles, SWIZZLE_128B, STAGES-deep mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x n x 16, fp32 accum in TMEM)
//   warps 2..9  : epilogue, two warps per TMEM lane quadrant taking alternate 32-column chunks
//                 (tcgen05.ld 32 lanes x 32 cols -> registers -> math -> global)
//   TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the mainloop of tile i+1.
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"
#include "../../include/panst3r_b200.h"

namespace pst3r {

enum { KIND_BF16 = 0, KIND_F32 = 1, KIND_SPLIT = 2 };

struct GemmEpi {
  void* out;
  long long ldo;
  int out_kind;  // KIND_*: bf16, fp32, or split bf16 (hi at column c, lo at column c + out_lo_off)
  int act;
  const float* bias;
  const float* col_scale;
  const void* residual;  // res_kind: bf16, fp32 or split bf16 (lo res_lo_off elements after hi)
  long long ldr;
  int res_mod_rows;
  float alpha;
  int store_mode;
  long long rows_per_batch, batch_stride, ldt;
  int grid_h, grid_w;
  int d2s_patch, d2s_ch;
  const float2* rope_cs;
  const int* rope_pos;
  int rope_cols;
  int rope_maxpos;
  // implicit-GEMM 3x3 convolution over a pixel-major [V, Hc, Wc, C] map (A is a 4-D TMA map; OOB = zero padding)
  int conv_cblocks;  // 64-channel blocks per tap (0 = plain GEMM)
  int conv_w, conv_h, conv_tpr;  // map width / height, 128-pixel tiles per image row
  // strided-batched mode (batches > 1): problem b reads A/B through the third TMA coordinate and writes
  // out + b * out_bs with bias + b * bias_bs (PLAIN store only)
  int batches;
  long long out_bs, bias_bs;
  // LayerNorm folded into this GEMM: A holds the RAW rows x, B the gamma-scaled weights, and the epilogue applies
  //   y = rstd * (acc - mu * ln_colsum[col]) (+ bias', which already contains beta W^T)
  // with (mu, rstd) of each row from the per-32-column partial sums (sum x, sum x^2) its producer left behind.
  const float2* ln_stats;
  int ln_slots;
  const float* ln_colsum;
  float ln_inv_dim, ln_eps;
  // producer side: partial (sum, sum of squares) of every stored bf16 32-column chunk, stats_out[row][col / 32]
  float2* stats_out;
  int stats_slots;
  // reference-precision mode: operands held as bf16 pairs x = hi + lo (hi = bf16(x), lo = bf16(x - hi), ~16 mantissa
  // bits).  The K loop runs split_terms passes over the k-blocks, accumulating into the same TMEM tile:
  //   3: A_hi B_hi + A_lo B_hi + A_hi B_lo (lo x lo, 2^-16 relative, dropped);  2: A (plain bf16) x (B_hi + B_lo).
  // The hi / lo part is a dimension of the TMA maps, so the producer only changes one coordinate.
  int split_terms;
  long long out_lo_off;
  int res_kind;
  long long res_lo_off;
  // plane-major fp32 store (PST3R_STORE_TRANSPOSED) through shared memory + TMA tile stores: every epilogue warp stages
  // its 32 rows x 32 columns transposed ([column][row], conflict free) and one lane issues a 32 x 32 box store — 32 full
  // 128-byte lines per instruction instead of 32 per-thread scalar stores with 64-bit address arithmetic each (the
  // mask-logit einsum was issue bound in exactly that loop: 21 instructions per element, profiles/r02_mask_logit_store.md)
  int tma_store;
  // Accumulator promotion (PROMOTE kernels, split mode with long reductions).  The tensor core ADDS into its fp32
  // accumulator with truncation, a bias of ~2^-24 per tcgen05.mma that grows linearly with the reduction length
  // (measured: 1.2e-4 after 3 x 11264 / 16 = 2112 accumulations, profiles/r02_error_table.md).  The K loop is therefore cut
  // into chunks of `promote` k-iterations that alternate between the two TMEM accumulators; the epilogue warps, idle
  // during the main loop otherwise, fold every finished chunk into fp32 REGISTER accumulators (round-to-nearest adds).
  int promote;
};

constexpr uint32_t GEMM_OUT_STAGE_BYTES = 8 * 32 * 32 * 4;  // 8 epilogue warps x (32 x 32 fp32)

// TMA-store epilogue of one 32-column chunk of one warp (all 32 lanes take part).  row0: first of the warp's 32 rows.
__device__ __forceinline__ void epilogue_chunk_tma(const GemmEpi& ep, const CUtensorMap* tmOut, float* stage, float (&v)[32],
                                                   int row0, int col0, int lane) {
  if (ep.alpha != 1.0f) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= ep.alpha;
  }
  if (lane == 0) bulk_wait_read_all();  // the previous box of this warp has left the staging tile
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; ++i) stage[i * 32 + lane] = v[i];
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    const int b = row0 / (int)ep.rows_per_batch;
    tma_store_3d(tmOut, stage, row0 - b * (int)ep.rows_per_batch, col0, b);
    bulk_commit_group();
  }
}

__device__ __forceinline__ void split_pack8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    const float2 f = unpack_bf16x2(h[i]);
    l[i] = pack_bf16x2(v[2 * i] - f.x, v[2 * i + 1] - f.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// (mean, rstd) of row `row` from the partial sums of its producer (fixed summation order: deterministic)
__device__ __forceinline__ float2 ln_row_stats(const GemmEpi& ep, int row, int M) {
  if (!ep.ln_stats || row >= M) return make_float2(0.0f, 1.0f);
  const float4* s = reinterpret_cast<const float4*>(ep.ln_stats + (long long)row * ep.ln_slots);
  float s1 = 0.0f, s2 = 0.0f;
  for (int i = 0; i < (ep.ln_slots >> 1); ++i) {
    const float4 t = __ldg(s + i);
    s1 += t.x + t.z;
    s2 += t.y + t.w;
  }
  const float mu = s1 * ep.ln_inv_dim;
  const float var = fmaxf(s2 * ep.ln_inv_dim - mu * mu, 0.0f);
  return make_float2(mu, rsqrtf(var + ep.ln_eps));
}

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr uint32_t B_BYTES = BN * GEMM_BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr uint32_t TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16;
  static constexpr uint32_t DYN_BYTES = TOTAL + 1024;  // slack for manual 1024 B alignment
  // epilogue staging tiles of the TMA-store mode live BEHIND everything else and are only requested by launches that use
  // them: the common launches keep ~31 KB of the SM's shared memory free, which lets the CTAs of a dependent
  // (programmatic dependent launch) kernel without shared memory — LayerNorm, split-KV combine — become resident early
  static constexpr uint32_t OUT_OFFSET = (TOTAL + 127) & ~127u;
  static constexpr uint32_t DYN_BYTES_TMA = OUT_OFFSET + GEMM_OUT_STAGE_BYTES + 1024;
  static_assert(DYN_BYTES_TMA <= 232448, "shared memory budget");
};

// ---- epilogue for one 32-column chunk owned by one thread (one output row) --------------------
// SPLIT_IO = false is the all-bf16 hot path (plain bf16 / fp32 outputs, bf16 residual): it compiles to exactly the
// round-1 epilogue.  SPLIT_IO = true adds the split-bf16 / fp32 residual kinds and the split stores of the
// reference-precision head; keeping them out of the common instantiation keeps its register allocation and unrolling.
template <bool SPLIT_IO>
__device__ __forceinline__ void epilogue_chunk(const GemmEpi& ep, float (&v)[32], int row, int col0, int M, int N,
                                               int bidx = 0, float2 ln = make_float2(0.0f, 1.0f)) {
  if (row >= M || col0 >= N) return;
  const int ncols = min(32, N - col0);
  const bool full = (ncols == 32);

  if (ep.alpha != 1.0f) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= ep.alpha;
  }
  if (ep.ln_colsum) {  // folded LayerNorm (host guarantees N % 32 == 0)
    const float4* u4 = reinterpret_cast<const float4*>(ep.ln_colsum + col0);
    const float nm = -ln.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 u = __ldg(u4 + i);
      v[4 * i + 0] = ln.y * fmaf(nm, u.x, v[4 * i + 0]);
      v[4 * i + 1] = ln.y * fmaf(nm, u.y, v[4 * i + 1]);
      v[4 * i + 2] = ln.y * fmaf(nm, u.z, v[4 * i + 2]);
      v[4 * i + 3] = ln.y * fmaf(nm, u.w, v[4 * i + 3]);
    }
  }
  if (ep.bias) {
    const float* bias = ep.bias + bidx * ep.bias_bs;
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 b = __ldg(b4 + i);
        v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < ncols) v[i] += __ldg(bias + col0 + i);
    }
  }
  if (ep.rope_cs && col0 < ep.rope_cols) {
    // chunk == one 32-wide half of a 64-wide head: even chunk rotates by y, odd chunk by x
    const int half = (col0 >> 5) & 1;
    int p = __ldg(ep.rope_pos + 2 * (long long)row + half);
    p = max(0, min(p, ep.rope_maxpos - 1));
    const float4* cs = reinterpret_cast<const float4*>(ep.rope_cs + (long long)p * 16);  // 128 B per position
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 c = __ldg(cs + j);  // (cos, sin) of frequencies 2j, 2j+1
      const float a0 = v[2 * j], b0 = v[2 * j + 16], a1 = v[2 * j + 1], b1 = v[2 * j + 17];
      v[2 * j] = a0 * c.x - b0 * c.y;
      v[2 * j + 16] = b0 * c.x + a0 * c.y;
      v[2 * j + 1] = a1 * c.z - b1 * c.w;
      v[2 * j + 17] = b1 * c.z + a1 * c.w;
    }
  }
  if (ep.act == PST3R_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float2 g = gelu_erf2(make_float2(v[i], v[i + 1]));
      v[i] = g.x; v[i + 1] = g.y;
    }
  } else if (ep.act == PST3R_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
  if (ep.col_scale) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols) v[i] *= __ldg(ep.col_scale + col0 + i);
  }
  if (ep.residual) {
    const int rrow = ep.res_mod_rows > 0 ? row % ep.res_mod_rows : row;
    if (SPLIT_IO && ep.res_kind == KIND_F32) {
      const float* r = reinterpret_cast<const float*>(ep.residual) + (long long)rrow * ep.ldr + col0;
      if (full && ((ep.ldr & 3) == 0)) {
        const float4* r4 = reinterpret_cast<const float4*>(r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 u = __ldg(r4 + i);
          v[4 * i + 0] += u.x; v[4 * i + 1] += u.y; v[4 * i + 2] += u.z; v[4 * i + 3] += u.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < ncols) v[i] += r[i];
      }
    } else {
      const bf16* r = reinterpret_cast<const bf16*>(ep.residual) + (long long)rrow * ep.ldr + col0;
#pragma unroll
      for (int part = 0; part < (SPLIT_IO ? 2 : 1); ++part) {
        if (part == 1) {
          if (ep.res_kind != KIND_SPLIT) break;
          r += ep.res_lo_off;
        }
        if (full && ((ep.ldr & 7) == 0) && (!SPLIT_IO || (ep.res_lo_off & 7) == 0)) {
          const uint4* r4 = reinterpret_cast<const uint4*>(r);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = __ldg(r4 + i);
            float2 f;
            f = unpack_bf16x2(u.x); v[8 * i + 0] += f.x; v[8 * i + 1] += f.y;
            f = unpack_bf16x2(u.y); v[8 * i + 2] += f.x; v[8 * i + 3] += f.y;
            f = unpack_bf16x2(u.z); v[8 * i + 4] += f.x; v[8 * i + 5] += f.y;
            f = unpack_bf16x2(u.w); v[8 * i + 6] += f.x; v[8 * i + 7] += f.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) v[i] += __bfloat162float(r[i]);
        }
      }
    }
  }

  switch (ep.store_mode) {
    case PST3R_STORE_PLAIN: {
      long long roff = (long long)row * ep.ldo + bidx * ep.out_bs;
      if (ep.rows_per_batch > 0) {
        const long long bb = row / ep.rows_per_batch;
        roff = bb * ep.batch_stride + (row - bb * ep.rows_per_batch) * ep.ldo;
      }
      if (SPLIT_IO && ep.out_kind == KIND_SPLIT) {
        bf16* o = reinterpret_cast<bf16*>(ep.out) + roff + col0;
        if (full && ((ep.ldo & 7) == 0) && ((ep.batch_stride & 7) == 0) && ((ep.out_lo_off & 7) == 0) && ((ep.out_bs & 7) == 0)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 hi, lo;
            split_pack8(v + 8 * i, hi, lo);
            reinterpret_cast<uint4*>(o)[i] = hi;
            reinterpret_cast<uint4*>(o + ep.out_lo_off)[i] = lo;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) {
              const bf16 h = __float2bfloat16(v[i]);
              o[i] = h;
              o[i + ep.out_lo_off] = __float2bfloat16(v[i] - __bfloat162float(h));
            }
        }
      } else if (ep.out_kind == KIND_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + roff + col0;
        if (full && ((ep.ldo & 3) == 0) && ((ep.batch_stride & 3) == 0) && ((ep.out_bs & 3) == 0)) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) o[i] = v[i];
        }
      } else {
        bf16* o = reinterpret_cast<bf16*>(ep.out) + roff + col0;
        if (full && ((ep.ldo & 7) == 0) && ((ep.batch_stride & 7) == 0)) {
          float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
            u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
            u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
            u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
            reinterpret_cast<uint4*>(o)[i] = u;
            if (ep.stats_out) {  // statistics of the values as stored (bf16-rounded): what the consumer's A operand holds
              float2 f;
              f = unpack_bf16x2(u.x); s1 += f.x + f.y; s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
              f = unpack_bf16x2(u.y); s1 += f.x + f.y; s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
              f = unpack_bf16x2(u.z); s1 += f.x + f.y; s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
              f = unpack_bf16x2(u.w); s1 += f.x + f.y; s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
            }
          }
          if (ep.stats_out) ep.stats_out[(long long)row * ep.stats_slots + (col0 >> 5)] = make_float2(s1, s2);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) o[i] = __float2bfloat16(v[i]);
        }
      }
    } break;
    case PST3R_STORE_TRANSPOSED: {
      // lanes of a warp hold consecutive rows -> each per-column store is a coalesced 128 B line
      const long long b = row / ep.rows_per_batch;
      const long long r = row - b * ep.rows_per_batch;
      if (ep.out_kind == KIND_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + b * ep.batch_stride + (long long)col0 * ep.ldt + r;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < ncols) o[(long long)i * ep.ldt] = v[i];
      } else if (SPLIT_IO && ep.out_kind == KIND_SPLIT) {
        bf16* o = reinterpret_cast<bf16*>(ep.out) + b * ep.batch_stride + (long long)col0 * ep.ldt + r;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < ncols) {
            const bf16 h = __float2bfloat16(v[i]);
            o[(long long)i * ep.ldt] = h;
            o[(long long)i * ep.ldt + ep.out_lo_off] = __float2bfloat16(v[i] - __bfloat162float(h));
          }
      } else {
        bf16* o = reinterpret_cast<bf16*>(ep.out) + b * ep.batch_stride + (long long)col0 * ep.ldt + r;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < ncols) o[(long long)i * ep.ldt] = __float2bfloat16(v[i]);
      }
    } break;
    case PST3R_STORE_PIXSHUF2: {
      // F.pixel_shuffle(., 2): out[b, c, 2y+i, 2x+j] = in[b, 4c+2i+j, y, x]; we keep pixel-major (NHWC) output.
      const int gw = ep.grid_w, gh = ep.grid_h;
      const int x = row % gw;
      const int t = row / gw;
      const int y = t % gh;
      const int b = t / gh;
      const int c0 = col0 >> 2;  // 8 output channels per 32-column chunk
      bf16* base = reinterpret_cast<bf16*>(ep.out);
#pragma unroll
      for (int ij = 0; ij < 4; ++ij) {
        const int i = ij >> 1, j = ij & 1;
        const long long orow = ((long long)(b * 2 * gh + 2 * y + i)) * (2 * gw) + 2 * x + j;
        bf16* o = base + orow * ep.ldo + c0;
        if (SPLIT_IO && ep.out_kind == KIND_SPLIT) {
          float t8[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) t8[c] = v[4 * c + ij];
          if (full && ((ep.ldo & 7) == 0) && ((ep.out_lo_off & 7) == 0)) {
            uint4 hi, lo;
            split_pack8(t8, hi, lo);
            *reinterpret_cast<uint4*>(o) = hi;
            *reinterpret_cast<uint4*>(o + ep.out_lo_off) = lo;
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (4 * c + ij < ncols) {
                const bf16 h = __float2bfloat16(t8[c]);
                o[c] = h;
                o[c + ep.out_lo_off] = __float2bfloat16(t8[c] - __bfloat162float(h));
              }
          }
        } else if (full && ((ep.ldo & 7) == 0)) {
          uint4 u;
          u.x = pack_bf16x2(v[0 * 4 + ij], v[1 * 4 + ij]);
          u.y = pack_bf16x2(v[2 * 4 + ij], v[3 * 4 + ij]);
          u.z = pack_bf16x2(v[4 * 4 + ij], v[5 * 4 + ij]);
          u.w = pack_bf16x2(v[6 * 4 + ij], v[7 * 4 + ij]);
          *reinterpret_cast<uint4*>(o) = u;
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (4 * c + ij < ncols) o[c] = __float2bfloat16(v[4 * c + ij]);
        }
      }
    } break;
    case PST3R_STORE_D2S: {
      // Pointmap head: weight rows were permuted on the host so that col = (i*P + j)*C + c; the output
      // pixel (y*P+i, x*P+j) then receives C contiguous floats and a whole patch row i is one contiguous run.
      const int gw = ep.grid_w, gh = ep.grid_h, P = ep.d2s_patch, C = ep.d2s_ch;
      const int x = row % gw;
      const int t = row / gw;
      const int y = t % gh;
      const int b = t / gh;
      const int run = P * C;
      float* base = reinterpret_cast<float*>(ep.out);
      const long long Wfull = (long long)gw * P;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if (k < ncols) {
          const int col = col0 + k;
          const int i = col / run;
          const int rem = col - i * run;
          const long long o = (((long long)(b * gh + y) * P + i) * Wfull + (long long)x * P) * C + rem;
          base[o] = v[k];
        }
      }
    } break;
    default: break;
  }
}

// MODE 0: all-bf16 hot path; 1: split-bf16 operands / outputs (reference-precision head), short reductions;
// 2: split mode with accumulator promotion (GemmEpi::promote)
template <int BN, int STAGES, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const GemmEpi ep, const int M, const int N, const int K) {
  using L = GemmSmem<BN, STAGES>;
  constexpr bool PROMOTE = MODE == 2;
  constexpr bool SPLIT_IO = MODE != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m_blocks = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n_blocks = (N + BN - 1) / BN;
  const int tiles_per_batch = num_m_blocks * num_n_blocks;
  const int num_tiles = tiles_per_batch * ep.batches;
  const int num_k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  const int terms = (SPLIT_IO && ep.split_terms > 1) ? ep.split_terms : 1;
  const int num_k_iters = num_k_blocks * terms;  // split mode: one pass over the k-blocks per product term
  // PROMOTE: the reduction is cut into chunks that alternate between the two accumulators (see GemmEpi::promote)
  const int chunk_iters = PROMOTE ? ep.promote : num_k_iters;
  const int num_chunks = (num_k_iters + chunk_iters - 1) / chunk_iters;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
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
  pdl_launch_dependents();  // prologue done: the next kernel's prologue may overlap our main loop's tail
  pdl_wait();               // predecessor's global writes are visible from here on

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int bidx = tile / tiles_per_batch;
        const int trem = tile - bidx * tiles_per_batch;
        const int m_blk = trem % num_m_blocks;
        const int n_blk = trem / num_m_blocks;
        for (int term = 0; term < terms; ++term) {
        // hi / lo part of each operand for this product term: (hi,hi), (lo,hi), (hi,lo) | (A,hi), (A,lo)
        const int pa = (terms == 3 && term == 1) ? 1 : 0;
        const int pb = (terms > 1 && term == terms - 1) ? 1 : 0;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          uint8_t* a_dst = smem + s * L::STAGE_BYTES;
          uint8_t* b_dst = a_dst + L::A_BYTES;
          if (ep.conv_cblocks > 0) {
            // k-block -> (tap, channel block); m-tile -> (view, y, 128-pixel segment of the row)
            const int tap = kb / ep.conv_cblocks, cb = kb - tap * ep.conv_cblocks;
            const int xb = m_blk % ep.conv_tpr, rowidx = m_blk / ep.conv_tpr;
            if (terms > 1)
              tma_load_5d(a_dst, &tmA, &full_bar[s], cb * 64, xb * GEMM_BM + tap % 3 - 1, rowidx % ep.conv_h + tap / 3 - 1,
                          rowidx / ep.conv_h, pa);
            else
              tma_load_4d(a_dst, &tmA, &full_bar[s], cb * 64, xb * GEMM_BM + tap % 3 - 1, rowidx % ep.conv_h + tap / 3 - 1,
                          rowidx / ep.conv_h);
          } else if (terms > 1) {
            if (ep.batches > 1)
              tma_load_4d(a_dst, &tmA, &full_bar[s], kb * GEMM_BK, m_blk * GEMM_BM, pa, bidx);
            else
              tma_load_3d(a_dst, &tmA, &full_bar[s], kb * GEMM_BK, m_blk * GEMM_BM, pa);
          } else if (ep.batches > 1) {
            tma_load_3d(a_dst, &tmA, &full_bar[s], kb * GEMM_BK, m_blk * GEMM_BM, bidx);
          } else {
            tma_load_2d(a_dst, &tmA, &full_bar[s], kb * GEMM_BK, m_blk * GEMM_BM);
          }
          if (terms > 1) {
            if (ep.batches > 1)
              tma_load_4d(b_dst, &tmB, &full_bar[s], kb * GEMM_BK, n_blk * BN, pb, bidx);
            else
              tma_load_3d(b_dst, &tmB, &full_bar[s], kb * GEMM_BK, n_blk * BN, pb);
          } else if (ep.batches > 1) {
            tma_load_3d(b_dst, &tmB, &full_bar[s], kb * GEMM_BK, n_blk * BN, bidx);
          } else {
            tma_load_2d(b_dst, &tmB, &full_bar[s], kb * GEMM_BK, n_blk * BN);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int local = 0;  // accumulator hand-overs so far: one per (tile, chunk)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = (tile % tiles_per_batch) / num_m_blocks;
        const int n_rem = N - n_blk * BN;
        const int n_mma = n_rem >= BN ? BN : ((n_rem + 15) & ~15);
        const uint32_t idesc = make_idesc_bf16(GEMM_BM, n_mma, 0, 0);
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
            const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
            const uint32_t b_addr = a_addr + L::A_BYTES;
            const uint64_t a_desc = make_smem_desc_sw128(a_addr, 0, 1024);
            const uint64_t b_desc = make_smem_desc_sw128(b_addr, 0, 1024);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // +32 B per UMMA_K step inside the 128 B swizzle span (address field is in 16 B units)
              umma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, ((kb - k0) | k) != 0);
            }
            umma_commit(&empty_bar[s]);
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          umma_commit(&tfull_bar[acc]);
        }
      }
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int cgrp = (warp - 2) >> 2;  // which half of the 32-column chunks
    float* out_stage = reinterpret_cast<float*>(smem + L::OUT_OFFSET) + (warp - 2) * 1024;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int bidx = tile / tiles_per_batch;
      const int trem = tile - bidx * tiles_per_batch;
      const int m_blk = trem % num_m_blocks;
      const int n_blk = trem / num_m_blocks;
      int row = m_blk * GEMM_BM + quad * 32 + lane;
      int m_lim = M;
      const float2 ln = ln_row_stats(ep, row, M);  // before the accumulator wait: overlaps the main loop
      if (ep.conv_cblocks > 0) {
        const int xcol = (m_blk % ep.conv_tpr) * GEMM_BM + quad * 32 + lane;
        row = (m_blk / ep.conv_tpr) * ep.conv_w + xcol;
        m_lim = xcol < ep.conv_w ? 0x7fffffff : 0;  // segment tail beyond the image row: nothing to store
      }
      const int n_rem = N - n_blk * BN;
      if constexpr (PROMOTE) {
        // fold every finished K chunk into register accumulators; this warp owns the 32-column chunks cgrp, cgrp + 2, ...
        constexpr int NCH = (BN / 32 + 1) / 2;
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
            const int c = cgrp + 2 * j;
            if (c < BN / 32 && c * 32 < n_rem) {  // warp-uniform
              uint32_t r[32];
              tmem_ld32(t_base + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) accv[j][i] += __uint_as_float(r[i]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int c = cgrp + 2 * j;
          if (c < BN / 32 && c * 32 < n_rem) epilogue_chunk<true>(ep, accv[j], row, n_blk * BN + c * 32, m_lim, N, bidx, ln);
        }
      } else {
        const int acc = local & 1;
        const uint32_t acc_ph = (local >> 1) & 1;
        mbar_wait(&tfull_bar[acc], acc_ph);
        tc_fence_after();
        const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = cgrp; c < BN / 32; c += 2) {
          if (c * 32 >= n_rem) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld32(t_base + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (ep.tma_store)
            epilogue_chunk_tma(ep, &tmOut, out_stage, v, m_blk * GEMM_BM + quad * 32, n_blk * BN + c * 32, lane);
          else
            epilogue_chunk<SPLIT_IO>(ep, v, row, n_blk * BN + c * 32, m_lim, N, bidx, ln);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        ++local;
      }
    }
    if (ep.tma_store && lane == 0) bulk_wait_all();  // our box stores are performed before the grid counts as complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int STAGES, int MODE = 0>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const GemmEpi& ep, int M,
                       int N, int K, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  auto kern = gemm_bf16_tn_kernel<BN, STAGES, MODE>;
  static bool configured = false;
  if (!configured) {
    PST3R_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES_TMA));
    configured = true;
  }
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + BN - 1) / BN) * ep.batches;
  const int grid = tiles < sm_budget() ? tiles : sm_budget();
  const size_t dyn = ep.tma_store ? L::DYN_BYTES_TMA : L::DYN_BYTES;
  PST3R_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), dyn, stream, tmA, tmB, tmOut, ep, M, N, K));
  return PST3R_OK;
}

}  // namespace pst3r

#include "gemm2.cuh"
#include "gemm_splitk.cuh"

using namespace pst3r;

static bool tma_store_enabled() {  // PST3R_TMA_STORE=0: per-thread stores (A/B measurements)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PST3R_TMA_STORE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static bool promote_enabled() {  // PST3R_PROMOTE=0: plain accumulation in split mode (A/B measurements)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PST3R_PROMOTE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

struct ConvCfg { int V, H, W, C, cpad; };
struct BatchCfg { int batches; long long a_bs, b_bs, out_bs, bias_bs; };

static int gemm_run(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t M, int32_t N, int32_t K,
                    const pst3r_gemm_epilogue* e, pst3r_stream_t stream_, const ConvCfg* conv,
                    const BatchCfg* bat = nullptr) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PST3R_CHECK_ARG(A && B && e && e->out, "gemm: null pointer");
  PST3R_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  PST3R_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 (lda=%lld ldb=%lld)",
                  (long long)lda, (long long)ldb);
  PST3R_CHECK_ARG((conv || lda >= K) && ldb >= K, "gemm: lda/ldb smaller than K");
  if (e->store_mode == PST3R_STORE_TRANSPOSED)
    PST3R_CHECK_ARG(e->rows_per_batch > 0 && e->ldt > 0, "gemm: TRANSPOSED store needs rows_per_batch/ldt");
  if (e->store_mode == PST3R_STORE_PIXSHUF2)
    PST3R_CHECK_ARG(e->grid_h > 0 && e->grid_w > 0 && (N % 4) == 0 && e->out_kind != PST3R_KIND_F32 && (M % (e->grid_h * e->grid_w)) == 0,
                    "gemm: PIXSHUF2 store needs grid, N%%4==0, bf16 out");
  if (e->store_mode == PST3R_STORE_D2S)
    PST3R_CHECK_ARG(e->grid_h > 0 && e->grid_w > 0 && e->d2s_patch > 0 && e->d2s_ch > 0 && e->out_kind == PST3R_KIND_F32 &&
                        N == e->d2s_patch * e->d2s_patch * e->d2s_ch && (M % (e->grid_h * e->grid_w)) == 0,
                    "gemm: D2S store needs grid/patch/channels, N == P*P*C, fp32 out");
  if (e->rope_cs)
    PST3R_CHECK_ARG(e->rope_pos && e->rope_cols > 0 && (e->rope_cols % 64) == 0 && e->rope_maxpos > 0 &&
                        (reinterpret_cast<uintptr_t>(e->rope_cs) & 15) == 0,
                    "gemm: rope epilogue needs pos, rope_cols %% 64 == 0, maxpos, 16-byte aligned table");

  GemmEpi ep;
  ep.out = e->out; ep.ldo = e->ldo; ep.out_kind = e->out_kind; ep.act = e->act;
  ep.bias = e->bias; ep.col_scale = e->col_scale;
  ep.residual = e->residual; ep.ldr = e->ldr; ep.res_mod_rows = e->res_mod_rows;
  ep.alpha = e->alpha; ep.store_mode = e->store_mode;
  ep.rows_per_batch = e->rows_per_batch; ep.batch_stride = e->batch_stride; ep.ldt = e->ldt;
  ep.grid_h = e->grid_h; ep.grid_w = e->grid_w; ep.d2s_patch = e->d2s_patch; ep.d2s_ch = e->d2s_ch;
  ep.rope_cs = reinterpret_cast<const float2*>(e->rope_cs); ep.rope_pos = e->rope_pos;
  ep.rope_cols = e->rope_cols; ep.rope_maxpos = e->rope_maxpos;
  ep.conv_cblocks = 0; ep.conv_w = ep.conv_h = ep.conv_tpr = 0;
  ep.batches = 1; ep.out_bs = 0; ep.bias_bs = 0;
  ep.ln_stats = reinterpret_cast<const float2*>(e->ln_stats); ep.ln_slots = e->ln_slots; ep.ln_colsum = e->ln_colsum;
  ep.ln_inv_dim = K > 0 ? 1.0f / (float)K : 0.0f; ep.ln_eps = e->ln_eps;
  ep.stats_out = reinterpret_cast<float2*>(e->stats_out); ep.stats_slots = (N + 31) / 32;
  if (e->ln_stats || e->ln_colsum)
    PST3R_CHECK_ARG(e->ln_stats && e->ln_colsum && e->ln_slots == (K + 31) / 32 && (K % 64) == 0 && (N % 32) == 0 && !conv &&
                        (reinterpret_cast<uintptr_t>(e->ln_stats) & 15) == 0 && (reinterpret_cast<uintptr_t>(e->ln_colsum) & 15) == 0,
                    "gemm: folded LayerNorm needs stats [M][K/32] + colsum [N], K %% 64 == 0, N %% 32 == 0");
  if (e->stats_out)
    PST3R_CHECK_ARG(e->store_mode == PST3R_STORE_PLAIN && e->rows_per_batch == 0 && e->out_kind == PST3R_KIND_BF16 && (N % 32) == 0 &&
                        (e->ldo % 8) == 0 && !conv && (reinterpret_cast<uintptr_t>(e->stats_out) & 7) == 0,
                    "gemm: stats_out needs a plain bf16 store with N %% 32 == 0 and ldo %% 8 == 0");
  const int terms = e->split_terms;
  ep.split_terms = terms; ep.out_lo_off = e->out_lo_off; ep.res_kind = e->res_kind; ep.res_lo_off = e->res_lo_off;
  PST3R_CHECK_ARG(terms == 0 || terms == 2 || terms == 3, "gemm: split_terms must be 0, 2 or 3");
  PST3R_CHECK_ARG(e->out_kind >= 0 && e->out_kind <= 2 && e->res_kind >= 0 && e->res_kind <= 2, "gemm: bad out_kind / res_kind");
  if (terms)
    PST3R_CHECK_ARG((!conv || terms == 3) && !e->ln_stats && (e->b_lo_off % 8) == 0 && e->b_lo_off > 0 &&
                        (terms == 2 || ((e->a_lo_off % 8) == 0 && e->a_lo_off > 0)),
                    "gemm: split operands need lo offsets that are positive multiples of 8 (no conv / folded LayerNorm)");
  if (e->out_kind == PST3R_KIND_SPLIT)
    PST3R_CHECK_ARG(e->out_lo_off > 0 && e->store_mode != PST3R_STORE_D2S && !e->stats_out, "gemm: split output needs out_lo_off");
  if (e->residual && e->res_kind == PST3R_KIND_SPLIT) PST3R_CHECK_ARG(e->res_lo_off > 0, "gemm: split residual needs res_lo_off");
  const int nb = bat ? bat->batches : 1;
  if (nb > 1) {
    ep.batches = nb; ep.out_bs = bat->out_bs; ep.bias_bs = bat->bias_bs;
  }
  if (conv) {
    ep.conv_cblocks = conv->cpad / 64; ep.conv_w = conv->W; ep.conv_h = conv->H; ep.conv_tpr = (conv->W + GEMM_BM - 1) / GEMM_BM;
  }

  // Split mode with a long reduction: promote the accumulator every PROMOTE_ITERS k-iterations (GemmEpi::promote);
  // short reductions (<= 16 iterations: 64 truncating accumulations, 4e-6) keep the plain kernels and tile shapes
  constexpr int PROMOTE_ITERS = 8;
  const int k_iters_total = ((K + GEMM_BK - 1) / GEMM_BK) * (terms ? terms : 1);
  ep.promote = (terms && k_iters_total > 16 && promote_enabled()) ? PROMOTE_ITERS : 0;

  // Plane-major fp32 planes (the mask-logit einsum): TMA tile stores when the layout allows it
  CUtensorMap tmOut;
  memset(&tmOut, 0, sizeof(tmOut));
  ep.tma_store = 0;
  if (e->store_mode == PST3R_STORE_TRANSPOSED && e->out_kind == PST3R_KIND_F32 && tma_store_enabled() && !conv && nb == 1 &&
      !ep.promote &&
      !e->bias && !e->residual && !e->col_scale && !e->rope_cs && !e->ln_stats && e->act == PST3R_ACT_NONE &&
      (e->rows_per_batch % 32) == 0 && (M % e->rows_per_batch) == 0 && (e->ldt % 4) == 0 && e->ldt >= e->rows_per_batch &&
      (reinterpret_cast<uintptr_t>(e->out) & 15) == 0 &&
      (M == e->rows_per_batch || ((e->batch_stride % 4) == 0 && e->batch_stride >= (int64_t)N * e->ldt))) {
    const uint64_t nbat = (uint64_t)(M / e->rows_per_batch);
    uint64_t dO[3] = {(uint64_t)e->rows_per_batch, (uint64_t)N, nbat};
    uint64_t sO[3] = {4, (uint64_t)e->ldt * 4, (uint64_t)(nbat > 1 ? e->batch_stride : (int64_t)N * e->ldt) * 4};
    uint32_t bO[3] = {32, 32, 1};
    int r = encode_tmap(&tmOut, e->out, 4, 3, dO, sO, bO, /*swizzle128=*/false);
    if (r) return r;
    ep.tma_store = 1;
  }

  // Tile-width heuristic: the widest BN whose tile count still fills the machine.
  const int sms = sm_budget();
  const int mb = (M + GEMM_BM - 1) / GEMM_BM;
  // A CTA ingests (128 + BN) x 64 bf16 per k-block and is bound by its L2->smem rate long before the tensor pipe
  // (measured: 128x256 tiles run at ~65 % of the cuBLAS peak), so pick the BN that minimises
  // waves x bytes-per-k-block, i.e. the per-SM load time of the slowest SM.
  int BN = 64;
  long long best = -1;
  for (int cand = 64; cand <= (ep.promote ? 128 : 256); cand *= 2) {  // promotion keeps BN/2 accumulators per thread
    const long long tiles = (long long)mb * ((N + cand - 1) / cand) * nb;
    const long long waves = (tiles + sms - 1) / sms;
    const long long cost = waves * (128 + cand);
    if (best < 0 || cost < best || (cost == best && cand > BN)) { best = cost; BN = cand; }
    if (cand >= N) break;  // wider tiles would only add padding
  }

  // large plain GEMMs with N % 256 == 0 go to the 2-CTA kernel (256 x 256 tiles per SM pair)
  const bool use2 = !conv && nb == 1 && gemm2_enabled() && (N % G2_BN) == 0 &&
                    (long long)((M + 255) / 256) * (N / G2_BN) >= (sms / 2);
  // split mode: the hi / lo part is dimension 2 of both maps (a plain-bf16 A has a single part)
  const uint64_t a_parts = terms == 3 ? 2 : 1;
  const uint64_t a_part_stride = terms == 3 ? (uint64_t)e->a_lo_off * 2 : (uint64_t)lda * 2;
  if (use2) {
    CUtensorMap tA2, tB2;
    uint64_t dA[3] = {(uint64_t)K, (uint64_t)M, a_parts}, sA[3] = {2, (uint64_t)lda * 2, a_part_stride};
    uint64_t dB[3] = {(uint64_t)K, (uint64_t)N, 2}, sB[3] = {2, (uint64_t)ldb * 2, (uint64_t)e->b_lo_off * 2};
    uint32_t bA[3] = {GEMM_BK, GEMM_BM, 1}, bB[3] = {GEMM_BK, G2_BN / 2, 1};
    int r = encode_tmap(&tA2, A, 2, terms ? 3 : 2, dA, sA, bA);
    if (r) return r;
    r = encode_tmap(&tB2, B, 2, terms ? 3 : 2, dB, sB, bB);
    if (r) return r;
    return launch_gemm2(tA2, tB2, tmOut, ep, M, N, K, stream);
  }

  // Small all-bf16 problems whose 128 x 64 tiles fill at most half of the machine: split the reduction over a 2-CTA cluster
  // (gemm_splitk.cuh) — the M = 768 projections and the K = 3072 fc2 of the sequential memory build.
  {
    static const bool splitk_env = !(getenv("PST3R_SPLITK") && getenv("PST3R_SPLITK")[0] == '0');  // A/B measurements
    const bool splitk_on = splitk_env && split_k_enabled();
    const long long tiles64 = (long long)mb * ((N + GSK_BN - 1) / GSK_BN);
    if (splitk_on && !conv && nb == 1 && !terms && !ep.promote && !ep.tma_store && e->out_kind != PST3R_KIND_SPLIT &&
        !(e->residual && e->res_kind != PST3R_KIND_BF16) && 2 * tiles64 <= sms && (K + GEMM_BK - 1) / GEMM_BK >= 8) {
      CUtensorMap tA, tB;
      uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {2, (uint64_t)lda * 2};
      uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[2] = {2, (uint64_t)ldb * 2};
      uint32_t bA[2] = {GEMM_BK, GEMM_BM}, bB[2] = {GEMM_BK, (uint32_t)GSK_BN};
      int r = encode_tmap(&tA, A, 2, 2, dA, sA, bA);
      if (r) return r;
      r = encode_tmap(&tB, B, 2, 2, dB, sB, bB);
      if (r) return r;
      return launch_gemm_splitk(tA, tB, ep, M, N, K, stream);
    }
  }

  CUtensorMap tmA, tmB;
  if (terms && conv) {
    // split pixel-major map: pixel rows [hi(ld) | lo(ld)], the part is dimension 4 of the activation map
    uint64_t dA[5] = {(uint64_t)conv->C, (uint64_t)conv->W, (uint64_t)conv->H, (uint64_t)conv->V, 2};
    uint64_t sA[5] = {2, (uint64_t)lda * 2, (uint64_t)lda * conv->W * 2, (uint64_t)lda * conv->W * conv->H * 2, (uint64_t)e->a_lo_off * 2};
    uint32_t bA[5] = {GEMM_BK, GEMM_BM, 1, 1, 1};
    uint64_t dB[3] = {(uint64_t)K, (uint64_t)N, 2}, sB[3] = {2, (uint64_t)ldb * 2, (uint64_t)e->b_lo_off * 2};
    uint32_t bB[3] = {GEMM_BK, (uint32_t)BN, 1};
    int r = encode_tmap(&tmA, A, 2, 5, dA, sA, bA);
    if (r) return r;
    r = encode_tmap(&tmB, B, 2, 3, dB, sB, bB);
    if (r) return r;
  } else if (terms) {
    // (K, rows, part[, batch]); rows / K beyond the logical extents are zero filled, never the next part / problem
    uint64_t dA[4] = {(uint64_t)K, (uint64_t)M, a_parts, (uint64_t)nb};
    uint64_t sA[4] = {2, (uint64_t)lda * 2, a_part_stride, (uint64_t)(bat ? bat->a_bs : 0) * 2};
    uint64_t dB[4] = {(uint64_t)K, (uint64_t)N, 2, (uint64_t)nb};
    uint64_t sB[4] = {2, (uint64_t)ldb * 2, (uint64_t)e->b_lo_off * 2, (uint64_t)(bat ? bat->b_bs : 0) * 2};
    uint32_t bA[4] = {GEMM_BK, GEMM_BM, 1, 1}, bB[4] = {GEMM_BK, (uint32_t)BN, 1, 1};
    int r = encode_tmap(&tmA, A, 2, nb > 1 ? 4 : 3, dA, sA, bA);
    if (r) return r;
    r = encode_tmap(&tmB, B, 2, nb > 1 ? 4 : 3, dB, sB, bB);
    if (r) return r;
  } else if (conv) {
    // pixel-major map [V, H, W, C] (row pitch lda): box = 64 channels x 128 pixels of one image row; coordinates outside
    // the map (x = -1, W; y = -1, H; channels >= C) are zero-filled by TMA == the convolution's zero padding
    uint64_t dims[4] = {(uint64_t)conv->C, (uint64_t)conv->W, (uint64_t)conv->H, (uint64_t)conv->V};
    uint64_t str[4] = {2, (uint64_t)lda * 2, (uint64_t)lda * conv->W * 2, (uint64_t)lda * conv->W * conv->H * 2};
    uint32_t box[4] = {GEMM_BK, GEMM_BM, 1, 1};
    int r = encode_tmap(&tmA, A, 2, 4, dims, str, box);
    if (r) return r;
  } else if (nb > 1) {
    // rows beyond M of problem b are out of bounds of dimension 1 (zero filled), never rows of problem b + 1
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, (uint64_t)nb};
    uint64_t str[3] = {2, (uint64_t)lda * 2, (uint64_t)bat->a_bs * 2};
    uint32_t box[3] = {GEMM_BK, GEMM_BM, 1};
    int r = encode_tmap(&tmA, A, 2, 3, dims, str, box);
    if (r) return r;
  } else {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[2] = {2, (uint64_t)lda * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int r = encode_tmap(&tmA, A, 2, 2, dims, str, box);
    if (r) return r;
  }
  if (terms) {
    // B map encoded above
  } else if (nb > 1) {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)nb};
    uint64_t str[3] = {2, (uint64_t)ldb * 2, (uint64_t)bat->b_bs * 2};
    uint32_t box[3] = {GEMM_BK, (uint32_t)BN, 1};
    int r = encode_tmap(&tmB, B, 2, 3, dims, str, box);
    if (r) return r;
  } else {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[2] = {2, (uint64_t)ldb * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)BN};
    int r = encode_tmap(&tmB, B, 2, 2, dims, str, box);
    if (r) return r;
  }
  if (ep.promote)
    return BN == 128 ? launch_gemm<128, 6, 2>(tmA, tmB, tmOut, ep, M, N, K, stream)
                     : launch_gemm<64, 8, 2>(tmA, tmB, tmOut, ep, M, N, K, stream);
  if (terms || e->out_kind == PST3R_KIND_SPLIT || (e->residual && e->res_kind != PST3R_KIND_BF16)) {
    switch (BN) {
      case 256: return launch_gemm<256, 4, 1>(tmA, tmB, tmOut, ep, M, N, K, stream);
      case 128: return launch_gemm<128, 6, 1>(tmA, tmB, tmOut, ep, M, N, K, stream);
      default: return launch_gemm<64, 8, 1>(tmA, tmB, tmOut, ep, M, N, K, stream);
    }
  }
  switch (BN) {
    case 256: return launch_gemm<256, 4>(tmA, tmB, tmOut, ep, M, N, K, stream);
    case 128: return launch_gemm<128, 6>(tmA, tmB, tmOut, ep, M, N, K, stream);
    default: return launch_gemm<64, 8>(tmA, tmB, tmOut, ep, M, N, K, stream);
  }
}

extern "C" int pst3r_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t M, int32_t N,
                               int32_t K, const pst3r_gemm_epilogue* e, pst3r_stream_t stream) {
  return gemm_run(A, lda, B, ldb, M, N, K, e, stream, nullptr);
}

extern "C" int pst3r_gemm_bf16_batched(const void* A, int64_t lda, int64_t a_batch_stride, const void* B, int64_t ldb,
                                       int64_t b_batch_stride, int32_t M, int32_t N, int32_t K, int32_t batches,
                                       const pst3r_gemm_epilogue* e, int64_t out_batch_stride, int64_t bias_batch_stride,
                                       pst3r_stream_t stream) {
  PST3R_CHECK_ARG(e && batches > 0, "gemm_batched: bad args");
  PST3R_CHECK_ARG(e->store_mode == PST3R_STORE_PLAIN && e->rows_per_batch == 0 && !e->residual && !e->rope_cs &&
                      !e->col_scale, "gemm_batched: plain store, bias / activation epilogue only");
  PST3R_CHECK_ARG((a_batch_stride % 8) == 0 && (b_batch_stride % 8) == 0 && (bias_batch_stride % 4) == 0 &&
                      (out_batch_stride % (e->out_kind == PST3R_KIND_F32 ? 4 : 8)) == 0,
                  "gemm_batched: batch strides must keep 16-byte alignment");
  BatchCfg bc{batches, a_batch_stride, b_batch_stride, out_batch_stride, bias_batch_stride};
  return gemm_run(A, lda, B, ldb, M, N, K, e, stream, nullptr, batches > 1 ? &bc : nullptr);
}

extern "C" int pst3r_conv3x3_nhwc(const void* x, int64_t ldx, int32_t V, int32_t H, int32_t W, int32_t C, const void* w,
                                  int32_t cpad, int32_t O, const pst3r_gemm_epilogue* e, pst3r_stream_t stream) {
  PST3R_CHECK_ARG(x && w && e && V > 0 && H > 0 && W > 0 && C > 0 && O > 0, "conv3x3: bad args");
  PST3R_CHECK_ARG(cpad >= C && (cpad % 64) == 0 && (ldx % 8) == 0 && ldx >= C, "conv3x3: cpad must be a multiple of 64 >= C; ldx %% 8 == 0");
  PST3R_CHECK_ARG(e->store_mode == PST3R_STORE_PLAIN && e->rows_per_batch == 0 && !e->rope_cs, "conv3x3: plain store only");
  ConvCfg c{V, H, W, C, cpad};
  const int tpr = (W + GEMM_BM - 1) / GEMM_BM;
  const long long mv = (long long)V * H * tpr * GEMM_BM;  // virtual rows: 128-pixel row segments
  PST3R_CHECK_ARG(mv < 0x7fffffffLL, "conv3x3: map too large");
  // split mode: weight rows are [hi(9 cpad) | lo(9 cpad)], pixel rows [hi | lo] with the lo part a_lo_off elements later
  return gemm_run(x, ldx, w, (e->split_terms ? 18LL : 9LL) * cpad, (int32_t)mv, O, 9 * cpad, e, stream, &c);
}
