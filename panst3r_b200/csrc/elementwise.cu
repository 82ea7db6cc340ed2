// HBM-bound helper kernels of the PanSt3R forward path: LayerNorm (with fused residual add), 2-D RoPE,
// broadcast adds, casts, ViT patchify (im2col), DINOv2 preprocessing, centre-2x2 pooling for the attention
// mask, mask-bit packing, row L2 normalisation, layout transposes, LoftUp featuriser pieces.
// All are plain coalesced / vectorised CUDA-core kernels: none of this work is GEMM shaped.
#include "common.cuh"
#include "host_util.h"
#include "../../include/panst3r_b200.h"

namespace pst3r {

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, three cached passes (mean, centred variance, normalise).
// ------------------------------------------------------------------------------------------------
// Kind-aware element access (PST3R_KIND_*): bf16, fp32, or split bf16 (value = hi + lo, lo stored lo_off elements
// after hi).  The reference-precision head (fp32 policy, panst3r.py:236-245) keeps its activations split.
__device__ __forceinline__ float kload(const void* p, int kind, long long idx, long long lo_off) {
  if (kind == PST3R_KIND_F32) return reinterpret_cast<const float*>(p)[idx];
  const bf16* b = reinterpret_cast<const bf16*>(p);
  float v = __bfloat162float(b[idx]);
  if (kind == PST3R_KIND_SPLIT) v += __bfloat162float(b[idx + lo_off]);
  return v;
}
__device__ __forceinline__ void kstore(void* p, int kind, long long idx, long long lo_off, float v) {
  if (kind == PST3R_KIND_F32) {
    reinterpret_cast<float*>(p)[idx] = v;
    return;
  }
  bf16* b = reinterpret_cast<bf16*>(p);
  const bf16 h = __float2bfloat16(v);
  b[idx] = h;
  if (kind == PST3R_KIND_SPLIT) b[idx + lo_off] = __float2bfloat16(v - __bfloat162float(h));
}

// Generic LayerNorm: one warp per row, three passes, any kinds (split rows are [hi(dim) | lo(dim)]).
__global__ void layernorm_kernel(const void* __restrict__ x, int x_kind, long long ldx, const void* __restrict__ add,
                                 int add_kind, long long ld_add, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, void* __restrict__ y, int y_kind, long long ldy,
                                 void* __restrict__ sum_out, int sum_kind, long long ld_sum, int rows, int dim, int x_rpb,
                                 long long x_bs) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long xo = x_rpb > 0 ? (long long)(row / x_rpb) * x_bs + (long long)(row % x_rpb) * ldx : (long long)row * ldx;
  const long long ao = (long long)row * ld_add;
  float s = 0.0f;
  for (int i = lane; i < dim; i += 32) {
    float v = kload(x, x_kind, xo + i, dim);
    if (add) v += kload(add, add_kind, ao + i, dim);
    s += v;
  }
  const float mean = warp_sum(s) / dim;
  float ss = 0.0f;
  for (int i = lane; i < dim; i += 32) {
    float v = kload(x, x_kind, xo + i, dim);
    if (add) v += kload(add, add_kind, ao + i, dim);
    const float d = v - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) / dim + eps);
  for (int i = lane; i < dim; i += 32) {
    float v = kload(x, x_kind, xo + i, dim);
    if (add) v += kload(add, add_kind, ao + i, dim);
    if (sum_out) kstore(sum_out, sum_kind, (long long)row * ld_sum + i, dim, v);
    kstore(y, y_kind, (long long)row * ldy + i, dim, (v - mean) * rstd * gamma[i] + beta[i]);
  }
}

// Fast path: bf16 rows with dim % 8 == 0 and dim <= VPL*256.  One warp per row; the row lives in registers
// (VPL 16-byte vectors per lane), so HBM sees exactly one read and one write per element.
template <int VPL, bool Y_F32>
__global__ void __launch_bounds__(256)
layernorm_vec_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ add, long long ld_add,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps, void* __restrict__ y,
                     long long ldy, bf16* __restrict__ sum_out, long long ld_sum, int rows, int dim, int x_rpb,
                     long long x_bs, long long y_bs, long long p_bs, int add_mod) {
  // x_rpb > 0: row = (batch, local row); x is read at batch * x_bs + local * ldx.  Batched mode additionally
  // writes y at batch * y_bs + local * ldy (y_bs != 0), uses gamma/beta + batch * p_bs and, with add_mod, the
  // SAME `add` rows for every batch (add row = local row).
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = dim >> 3;
  const int bidx = x_rpb > 0 ? row / x_rpb : 0;
  const int lrow = x_rpb > 0 ? row - bidx * x_rpb : row;
  const long long xo = x_rpb > 0 ? (long long)bidx * x_bs + (long long)lrow * ldx : (long long)row * ldx;
  const long long yo = y_bs != 0 ? (long long)bidx * y_bs + (long long)lrow * ldy : (long long)row * ldy;
  gamma += bidx * p_bs;
  beta += bidx * p_bs;
  const uint4* xp = reinterpret_cast<const uint4*>(x + xo);
  const uint4* ap = add ? reinterpret_cast<const uint4*>(add + (long long)(add_mod ? lrow : row) * ld_add) : nullptr;
  float v[VPL][8];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      const uint4 u = __ldg(xp + vi);
      float2 f;
      f = unpack_bf16x2(u.x); v[i][0] = f.x; v[i][1] = f.y;
      f = unpack_bf16x2(u.y); v[i][2] = f.x; v[i][3] = f.y;
      f = unpack_bf16x2(u.z); v[i][4] = f.x; v[i][5] = f.y;
      f = unpack_bf16x2(u.w); v[i][6] = f.x; v[i][7] = f.y;
      if (ap) {
        const uint4 a = __ldg(ap + vi);
        f = unpack_bf16x2(a.x); v[i][0] += f.x; v[i][1] += f.y;
        f = unpack_bf16x2(a.y); v[i][2] += f.x; v[i][3] += f.y;
        f = unpack_bf16x2(a.z); v[i][4] += f.x; v[i][5] += f.y;
        f = unpack_bf16x2(a.w); v[i][6] += f.x; v[i][7] += f.y;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[i][k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[i][k] = 0.0f;
    }
  }
  const float mean = warp_sum(s) / dim;
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = v[i][k] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / dim + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      if (sum_out) {
        uint4 u;
        u.x = pack_bf16x2(v[i][0], v[i][1]); u.y = pack_bf16x2(v[i][2], v[i][3]);
        u.z = pack_bf16x2(v[i][4], v[i][5]); u.w = pack_bf16x2(v[i][6], v[i][7]);
        reinterpret_cast<uint4*>(sum_out + (long long)row * ld_sum)[vi] = u;
      }
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi + 1);
      float o[8];
      o[0] = (v[i][0] - mean) * rstd * g0.x + b0.x; o[1] = (v[i][1] - mean) * rstd * g0.y + b0.y;
      o[2] = (v[i][2] - mean) * rstd * g0.z + b0.z; o[3] = (v[i][3] - mean) * rstd * g0.w + b0.w;
      o[4] = (v[i][4] - mean) * rstd * g1.x + b1.x; o[5] = (v[i][5] - mean) * rstd * g1.y + b1.y;
      o[6] = (v[i][6] - mean) * rstd * g1.z + b1.z; o[7] = (v[i][7] - mean) * rstd * g1.w + b1.w;
      if (Y_F32) {
        float4* yp = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + yo) + 2 * vi;
        yp[0] = make_float4(o[0], o[1], o[2], o[3]);
        yp[1] = make_float4(o[4], o[5], o[6], o[7]);
      } else {
        uint4 u;
        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
        u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
        reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(y) + yo)[vi] = u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 2-D RoPE (curope semantics), in place.  One thread per (token, head, pair).
// ------------------------------------------------------------------------------------------------
__global__ void rope2d_kernel(bf16* __restrict__ t, long long s_b, long long s_n, long long s_h,
                              const int* __restrict__ pos, int B, int N, int H, int D, float log2_base, float fwd) {
  const int Q = D / 4;  // pairs per half
  const long long total = (long long)B * N * H * 2 * Q;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = idx % Q;
  long long r = idx / Q;
  const int half = r % 2; r /= 2;
  const int h = r % H; r /= H;
  const int n = r % N;
  const int b = r / N;
  const int pp = pos[((long long)b * N + n) * 2 + half];
  const float inv_freq = exp2f(-log2_base * (float)j / (float)Q);
  float sn, cs;
  sincosf((float)pp * inv_freq * fwd, &sn, &cs);
  bf16* base = t + (long long)b * s_b + (long long)n * s_n + (long long)h * s_h + half * (D / 2);
  const float u = __bfloat162float(base[j]);
  const float v = __bfloat162float(base[j + Q]);
  base[j] = __float2bfloat16(u * cs - v * sn);
  base[j + Q] = __float2bfloat16(v * cs + u * sn);
}

// ------------------------------------------------------------------------------------------------
__global__ void add_bcast_kernel(const bf16* __restrict__ a, long long lda, const bf16* __restrict__ b,
                                 long long ldb, int b_rows, bf16* __restrict__ out, long long ldo, int rows,
                                 int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = cols >> 1;
  if (idx >= (long long)rows * cv) return;
  const int c = (idx % cv) * 2;
  const int r = idx / cv;
  const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + (long long)r * lda + c));
  const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(b + (long long)(r % b_rows) * ldb + c));
  *reinterpret_cast<__nv_bfloat162*>(out + (long long)r * ldo + c) = __floats2bfloat162_rn(x.x + y.x, x.y + y.y);
}
__global__ void add_bcast_k_kernel(const void* __restrict__ a, int a_kind, long long lda, const void* __restrict__ b,
                                   int b_kind, long long ldb, int b_rows, void* __restrict__ out, int out_kind,
                                   long long ldo, int rows, int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int c = idx % cols;
  const int r = idx / cols;
  const float v = kload(a, a_kind, (long long)r * lda + c, cols) + kload(b, b_kind, (long long)(r % b_rows) * ldb + c, cols);
  kstore(out, out_kind, (long long)r * ldo + c, cols, v);
}
__global__ void convert_kernel(const void* __restrict__ x, int x_kind, long long ldx, void* __restrict__ y, int y_kind,
                               long long ldy, int rows, int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int c = idx % cols;
  const int r = idx / cols;
  kstore(y, y_kind, (long long)r * ldy + c, cols, kload(x, x_kind, (long long)r * ldx + c, cols));
}

// Masked row softmax of the reference-precision (unfused) attention: P[r][k] = softmax_k(S[r][k]) over the keys the
// bit mask leaves open (bit k&31 of word [q][k>>5] set => blocked; q = r % Q: the mask is shared by all heads,
// mask_transformer.py:272).  S fp32 [rows][>=Nk] already carries the 1/sqrt(hd) scale.  One block per row; S stays in L2.
__global__ void softmax_rows_kernel(const float* __restrict__ S, long long lds, int Nk, const uint32_t* __restrict__ bits,
                                    long long mask_sq, int Q, void* __restrict__ out, int out_kind, long long ldo,
                                    long long out_lo_off) {
  const int r = blockIdx.x;
  const float* srow = S + (long long)r * lds;
  const uint32_t* mrow = bits ? bits + (long long)(r % Q) * mask_sq : nullptr;
  __shared__ float red[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  auto blocked = [&](int k) { return mrow && ((mrow[k >> 5] >> (k & 31)) & 1u); };
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < Nk; k += blockDim.x)
    if (!blocked(k)) mx = fmaxf(mx, srow[k]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nwarps ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  mx = red[32];
  __syncthreads();
  float sum = 0.0f;
  for (int k = threadIdx.x; k < Nk; k += blockDim.x)
    if (!blocked(k)) sum += expf(srow[k] - mx);
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nwarps ? red[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  const float inv = red[32] > 0.0f ? 1.0f / red[32] : 0.0f;
  for (int k = threadIdx.x; k < Nk; k += blockDim.x) {
    const float pv = blocked(k) ? 0.0f : expf(srow[k] - mx) * inv;
    kstore(out, out_kind, (long long)r * ldo + k, out_lo_off, pv);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                     int rows, int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int c = idx % cols;
  const int r = idx / cols;
  y[(long long)r * ldy + c] = __float2bfloat16(x[(long long)r * ldx + c]);
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ y, long long ldy,
                                     int rows, int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int c = idx % cols;
  const int r = idx / cols;
  y[(long long)r * ldy + c] = __bfloat162float(x[(long long)r * ldx + c]);
}

// ------------------------------------------------------------------------------------------------
// Patchify: out[(b, y, x), c*P*P + i*P + j] = img[b, c, y*P + i, x*P + j]
// ------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const float* __restrict__ img, int B, int H, int W, int P, bf16* __restrict__ out,
                                long long ldo) {
  const int gh = H / P, gw = W / P;
  const int kdim = 3 * P * P;
  const long long total = (long long)B * gh * gw * ldo;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int col = idx % ldo;
  const long long t = idx / ldo;
  float v = 0.0f;
  if (col < kdim) {
    const int j = col % P;
    const int i = (col / P) % P;
    const int c = col / (P * P);
    const int x = t % gw;
    const int y = (t / gw) % gh;
    const int b = t / ((long long)gw * gh);
    v = img[(((long long)b * 3 + c) * H + (y * P + i)) * W + (x * P + j)];
  }
  out[idx] = __float2bfloat16(v);
}

// DINOv2: [-1,1] -> [0,1] -> ImageNet normalise -> bilinear (align_corners=False) to (Ho, Wo) -> im2col(P)
__global__ void dino_preprocess_patchify_kernel(const float* __restrict__ img, int B, int H, int W, int Ho, int Wo,
                                                int P, bf16* __restrict__ out, long long ldo) {
  const int gh = Ho / P, gw = Wo / P;
  const int kdim = 3 * P * P;
  const long long total = (long long)B * gh * gw * ldo;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int col = idx % ldo;
  const long long t = idx / ldo;
  float v = 0.0f;
  if (col < kdim) {
    const int j = col % P;
    const int i = (col / P) % P;
    const int c = col / (P * P);
    const int x = t % gw;
    const int y = (t / gw) % gh;
    const int b = t / ((long long)gw * gh);
    const int oy = y * P + i, ox = x * P + j;
    const float sy = fmaxf(((float)oy + 0.5f) * ((float)H / (float)Ho) - 0.5f, 0.0f);
    const float sx = fmaxf(((float)ox + 0.5f) * ((float)W / (float)Wo) - 0.5f, 0.0f);
    const int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float* pc = img + ((long long)b * 3 + c) * H * W;
    const float v00 = pc[(long long)y0 * W + x0], v01 = pc[(long long)y0 * W + x1];
    const float v10 = pc[(long long)y1 * W + x0], v11 = pc[(long long)y1 * W + x1];
    const float raw = (1.0f - ly) * ((1.0f - lx) * v00 + lx * v01) + ly * ((1.0f - lx) * v10 + lx * v11);
    const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
    const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
    v = ((raw * 0.5f + 0.5f) - mean) / stdv;
  }
  out[idx] = __float2bfloat16(v);
}

// ------------------------------------------------------------------------------------------------
// centre 2x2 mean of every 8x8 cell (pixel-major feature map)
// ------------------------------------------------------------------------------------------------
__global__ void center_pool8_kernel(const bf16* __restrict__ f, int B, int Hm, int Wm, int C, bf16* __restrict__ out) {
  const int gh = Hm / 8, gw = Wm / 8;
  const int cv = C >> 1;
  const long long total = (long long)B * gh * gw * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (idx % cv) * 2;
  long long t = idx / cv;
  const int X = t % gw; t /= gw;
  const int Y = t % gh;
  const int b = t / gh;
  const bf16* base = f + (long long)b * Hm * Wm * C;
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const long long pix = (long long)(8 * Y + 3 + dy) * Wm + (8 * X + 3 + dx);
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base + pix * C + c));
      acc.x += v.x; acc.y += v.y;
    }
  *reinterpret_cast<__nv_bfloat162*>(out + (((long long)b * gh + Y) * gw + X) * C + c) =
      __floats2bfloat162_rn(acc.x * 0.25f, acc.y * 0.25f);
}
// split-bf16 map: pixel rows are [hi(C) | lo(C)], in and out
__global__ void center_pool8_split_kernel(const bf16* __restrict__ f, int B, int Hm, int Wm, int C, bf16* __restrict__ out) {
  const int gh = Hm / 8, gw = Wm / 8;
  const long long total = (long long)B * gh * gw * C;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % C;
  long long t = idx / C;
  const int X = t % gw; t /= gw;
  const int Y = t % gh;
  const int b = t / gh;
  const bf16* base = f + (long long)b * Hm * Wm * 2 * C;
  float acc = 0.0f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const long long pix = (long long)(8 * Y + 3 + dy) * Wm + (8 * X + 3 + dx);
      acc += kload(base, PST3R_KIND_SPLIT, pix * 2 * C + c, C);
    }
  kstore(out, PST3R_KIND_SPLIT, (((long long)b * gh + Y) * gw + X) * 2 * C + c, C, acc * 0.25f);
}

// ------------------------------------------------------------------------------------------------
// mask bits: one block per query row
// ------------------------------------------------------------------------------------------------
__global__ void attn_mask_bits_kernel(const float* __restrict__ logits_t, long long ld, int Nk, uint32_t* __restrict__ bits,
                                      int words_per_row) {
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* row = logits_t + (long long)q * ld;
  uint32_t* brow = bits + (long long)q * words_per_row;
  int blocked = 0;
  for (int w = warp; w < words_per_row; w += nwarps) {
    const int k = w * 32 + lane;
    const bool blk = (k < Nk) && (row[k] < 0.0f);
    const uint32_t word = __ballot_sync(0xffffffffu, blk);
    if (lane == 0) brow[w] = word;
    blocked += __popc(word);
  }
  __shared__ int s_cnt[32];
  if (lane == 0) s_cnt[warp] = blocked;
  __syncthreads();
  int tot = 0;
  for (int i = 0; i < nwarps; ++i) tot += s_cnt[i];
  if (tot == Nk) {  // every key blocked -> attend everywhere (mask_transformer.py:172)
    for (int w = threadIdx.x; w < words_per_row; w += blockDim.x) brow[w] = 0u;
  }
}

__global__ void l2norm_rows_kernel(const float* __restrict__ x, long long ldx, void* __restrict__ y, int y_f32,
                                   long long ldy, int rows, int cols, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float ss = 0.0f;
  for (int i = lane; i < cols; i += 32) {
    const float v = x[(long long)row * ldx + i];
    ss += v * v;
  }
  const float inv = 1.0f / (sqrtf(warp_sum(ss)) + eps);
  for (int i = lane; i < cols; i += 32) {
    kstore(y, y_f32, (long long)row * ldy + i, cols, x[(long long)row * ldx + i] * inv);  // y_f32 is a PST3R_KIND_*
  }
}

// bf16 [B, HW, C] -> fp32 [B, C, HW]
__global__ void nhwc_to_nchw_f32_kernel(const bf16* __restrict__ x, int x_kind, int HW, int C, float* __restrict__ y) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const long long ldx = x_kind == PST3R_KIND_SPLIT ? 2 * C : C;
  const bf16* xb = x + (long long)b * HW * ldx;
  float* yb = y + (long long)b * HW * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? kload(xb, x_kind, (long long)p * ldx + c, C) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (p < HW && c < C) yb[(long long)c * HW + p] = tile[threadIdx.x][i];
  }
}

static inline unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace pst3r

using namespace pst3r;

extern "C" const char* pst3r_last_error(void) { return get_last_error(); }
extern "C" int pst3r_version(void) { return PST3R_ABI_VERSION; }
extern "C" int pst3r_check_device(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_last_error("no CUDA device");
    return PST3R_ERR_NODEVICE;
  }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    set_last_error("cudaGetDeviceProperties failed");
    return PST3R_ERR_NODEVICE;
  }
  if (p.major != 10) {
    set_last_error("device %s is sm_%d%d; this library contains sm_100a code only", p.name, p.major, p.minor);
    return PST3R_ERR_NODEVICE;
  }
  return PST3R_OK;
}

static int layernorm_run(const void* x, int32_t x_kind, int64_t ldx, const void* add, int32_t add_kind, int64_t ld_add,
                         const float* gamma, const float* beta, float eps, void* y, int32_t y_kind, int64_t ldy,
                         void* sum_out, int32_t sum_kind, int64_t ld_sum, int32_t rows, int32_t dim, int32_t x_rpb,
                         int64_t x_bs, int64_t y_bs, int64_t p_bs, int32_t add_mod, pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(x && gamma && beta && y && rows > 0 && dim > 0, "layernorm: bad args");
  auto kind_ok = [](int k) { return k >= 0 && k <= 2; };
  PST3R_CHECK_ARG(kind_ok(x_kind) && kind_ok(y_kind) && kind_ok(add_kind) && kind_ok(sum_kind), "layernorm: bad element kind");
  const int wpb = 8;
  const unsigned grid = blocks_for(rows, wpb);
  const bf16* a = reinterpret_cast<const bf16*>(add);
  bf16* so = reinterpret_cast<bf16*>(sum_out);
  const int y_f32 = y_kind == PST3R_KIND_F32;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec_ok = x_kind == PST3R_KIND_BF16 && y_kind != PST3R_KIND_SPLIT && (!add || add_kind == PST3R_KIND_BF16) &&
                      (!sum_out || sum_kind == PST3R_KIND_BF16) && (dim % 8) == 0 && dim <= 12 * 256 && (ldx % 8) == 0 &&
                      (x_bs % 8) == 0 && al16(x) &&
                      (!add || ((ld_add % 8) == 0 && al16(add))) && (!sum_out || ((ld_sum % 8) == 0 && al16(sum_out))) &&
                      al16(y) && (ldy % (y_f32 ? 4 : 8)) == 0 && al16(gamma) && al16(beta) &&
                      (y_bs % (y_f32 ? 4 : 8)) == 0 && (p_bs % 4) == 0;
  if (vec_ok) {
    const int vpl = (dim / 8 + 31) / 32;
    const bf16* xb = reinterpret_cast<const bf16*>(x);
#define PST3R_LN_CASE(V)                                                                                              \
  if (y_f32)                                                                                                          \
    PST3R_CHECK_CUDA(launch_pdl(layernorm_vec_kernel<V, true>, dim3(grid), dim3(wpb * 32), 0, s, xb, ldx, a, ld_add, gamma, \
                                beta, eps, y, ldy, so, ld_sum, rows, dim, x_rpb, x_bs, y_bs, p_bs, add_mod));       \
  else                                                                                                                \
    PST3R_CHECK_CUDA(launch_pdl(layernorm_vec_kernel<V, false>, dim3(grid), dim3(wpb * 32), 0, s, xb, ldx, a, ld_add, gamma, \
                                beta, eps, y, ldy, so, ld_sum, rows, dim, x_rpb, x_bs, y_bs, p_bs, add_mod));
    if (vpl <= 2) { PST3R_LN_CASE(2) }
    else if (vpl <= 4) { PST3R_LN_CASE(4) }
    else if (vpl <= 8) { PST3R_LN_CASE(8) }
    else { PST3R_LN_CASE(12) }
#undef PST3R_LN_CASE
    PST3R_CHECK_CUDA(cudaGetLastError());
    return PST3R_OK;
  }
  PST3R_CHECK_ARG(y_bs == 0 && p_bs == 0 && !add_mod, "layernorm_batched: needs bf16 rows, dim %% 8 == 0, 16-byte alignment");
  layernorm_kernel<<<grid, wpb * 32, 0, s>>>(x, x_kind, ldx, add, add_kind, ld_add, gamma, beta, eps, y, y_kind, ldy, sum_out,
                                              sum_kind, ld_sum, rows, dim, x_rpb, x_bs);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_layernorm(const void* x, int32_t x_kind, int64_t ldx, const void* add, int32_t add_kind, int64_t ld_add,
                               const float* gamma, const float* beta, float eps, void* y, int32_t y_kind, int64_t ldy,
                               void* sum_out, int32_t sum_kind, int64_t ld_sum, int32_t rows, int32_t dim, int32_t x_rpb,
                               int64_t x_bs, pst3r_stream_t s_) {
  return layernorm_run(x, x_kind, ldx, add, add_kind, ld_add, gamma, beta, eps, y, y_kind, ldy, sum_out, sum_kind, ld_sum, rows,
                       dim, x_rpb, x_bs, 0, 0, 0, s_);
}

extern "C" int pst3r_layernorm_batched(const void* x, int64_t ldx, int64_t x_batch_stride, const void* add, int64_t ld_add,
                                       const float* gamma, const float* beta, int64_t param_batch_stride, float eps,
                                       void* y, int64_t ldy, int64_t y_batch_stride, int32_t rows_per_batch,
                                       int32_t batches, int32_t dim, pst3r_stream_t s_) {
  PST3R_CHECK_ARG(rows_per_batch > 0 && batches > 0 && y_batch_stride != 0, "layernorm_batched: bad args");
  return layernorm_run(x, 0, ldx, add, 0, ld_add, gamma, beta, eps, y, 0, ldy, nullptr, 0, 0, rows_per_batch * batches, dim,
                       rows_per_batch, x_batch_stride, y_batch_stride, param_batch_stride, 1, s_);
}

extern "C" int pst3r_rope2d(void* tokens, int64_t s_b, int64_t s_n, int64_t s_h, const int32_t* pos, int32_t B,
                            int32_t N, int32_t H, int32_t D, float base, float fwd, pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(tokens && pos && B > 0 && N > 0 && H > 0 && D > 0 && (D % 4) == 0, "rope2d: bad args (D %% 4 != 0?)");
  const long long total = (long long)B * N * H * (D / 2);
  rope2d_kernel<<<blocks_for(total, 256), 256, 0, s>>>(reinterpret_cast<bf16*>(tokens), s_b, s_n, s_h, pos, B, N, H, D,
                                                       log2f(base), fwd);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_add_bcast(const void* a, int32_t a_kind, int64_t lda, const void* b, int32_t b_kind, int64_t ldb,
                               int32_t b_rows, void* out, int32_t out_kind, int64_t ldo, int32_t rows, int32_t cols,
                               pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(a && b && out && rows > 0 && cols > 0 && b_rows > 0, "add_bcast: bad args");
  if (a_kind == 0 && b_kind == 0 && out_kind == 0 && (cols % 2) == 0 && (lda % 2) == 0 && (ldb % 2) == 0 && (ldo % 2) == 0) {
    add_bcast_kernel<<<blocks_for((long long)rows * (cols / 2), 256), 256, 0, s>>>(
        reinterpret_cast<const bf16*>(a), lda, reinterpret_cast<const bf16*>(b), ldb, b_rows, reinterpret_cast<bf16*>(out),
        ldo, rows, cols);
  } else {
    PST3R_CHECK_ARG(a_kind >= 0 && a_kind <= 2 && b_kind >= 0 && b_kind <= 2 && out_kind >= 0 && out_kind <= 2, "add_bcast: bad kind");
    add_bcast_k_kernel<<<blocks_for((long long)rows * cols, 256), 256, 0, s>>>(a, a_kind, lda, b, b_kind, ldb, b_rows, out,
                                                                              out_kind, ldo, rows, cols);
  }
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_convert(const void* x, int32_t x_kind, int64_t ldx, void* y, int32_t y_kind, int64_t ldy, int32_t rows,
                             int32_t cols, pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(x && y && rows > 0 && cols > 0 && x_kind >= 0 && x_kind <= 2 && y_kind >= 0 && y_kind <= 2, "convert: bad args");
  convert_kernel<<<blocks_for((long long)rows * cols, 256), 256, 0, s>>>(x, x_kind, ldx, y, y_kind, ldy, rows, cols);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_softmax_rows(const float* S, int64_t lds, int32_t rows, int32_t Nk, const uint32_t* mask_bits,
                                  int64_t mask_sq, int32_t Q, void* out, int32_t out_kind, int64_t ldo, int64_t out_lo_off,
                                  pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(S && out && rows > 0 && Nk > 0 && lds >= Nk && out_kind >= 0 && out_kind <= 2 && Q > 0 &&
                      (out_kind == PST3R_KIND_SPLIT ? (out_lo_off >= Nk && ldo >= out_lo_off + Nk) : ldo >= Nk),
                  "softmax_rows: bad args");
  if (mask_bits) PST3R_CHECK_ARG(mask_sq * 32 >= Nk, "softmax_rows: mask rows shorter than Nk");
  softmax_rows_kernel<<<rows, 256, 0, s>>>(S, lds, Nk, mask_bits, mask_sq, Q, out, out_kind, ldo, out_lo_off);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_cast_f32_to_bf16(const float* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t cols,
                                      pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(x && y && rows > 0 && cols > 0, "cast: bad args");
  cast_f32_bf16_kernel<<<blocks_for((long long)rows * cols, 256), 256, 0, s>>>(x, ldx, reinterpret_cast<bf16*>(y), ldy, rows, cols);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}
extern "C" int pst3r_cast_bf16_to_f32(const void* x, int64_t ldx, float* y, int64_t ldy, int32_t rows, int32_t cols,
                                      pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(x && y && rows > 0 && cols > 0, "cast: bad args");
  cast_bf16_f32_kernel<<<blocks_for((long long)rows * cols, 256), 256, 0, s>>>(reinterpret_cast<const bf16*>(x), ldx, y, ldy, rows, cols);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_patchify(const float* img, int32_t B, int32_t H, int32_t W, int32_t P, void* out, int64_t ldo,
                              pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(img && out && B > 0 && P > 0 && (H % P) == 0 && (W % P) == 0 && ldo >= 3 * P * P, "patchify: bad args");
  const long long total = (long long)B * (H / P) * (W / P) * ldo;
  patchify_kernel<<<blocks_for(total, 256), 256, 0, s>>>(img, B, H, W, P, reinterpret_cast<bf16*>(out), ldo);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_dino_preprocess_patchify(const float* img, int32_t B, int32_t H, int32_t W, int32_t Ho, int32_t Wo,
                                              int32_t P, void* out, int64_t ldo, pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(img && out && B > 0 && P > 0 && (Ho % P) == 0 && (Wo % P) == 0 && ldo >= 3 * P * P,
                  "dino_preprocess_patchify: bad args");
  const long long total = (long long)B * (Ho / P) * (Wo / P) * ldo;
  dino_preprocess_patchify_kernel<<<blocks_for(total, 256), 256, 0, s>>>(img, B, H, W, Ho, Wo, P,
                                                                        reinterpret_cast<bf16*>(out), ldo);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_center_pool8(const void* feats, int32_t kind, int32_t B, int32_t Hm, int32_t Wm, int32_t C, void* out,
                                  pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(feats && out && B > 0 && (Hm % 8) == 0 && (Wm % 8) == 0 && (C % 2) == 0 &&
                      (kind == PST3R_KIND_BF16 || kind == PST3R_KIND_SPLIT), "center_pool8: bad args");
  if (kind == PST3R_KIND_SPLIT) {
    const long long total = (long long)B * (Hm / 8) * (Wm / 8) * C;
    center_pool8_split_kernel<<<blocks_for(total, 256), 256, 0, s>>>(reinterpret_cast<const bf16*>(feats), B, Hm, Wm, C,
                                                                    reinterpret_cast<bf16*>(out));
    PST3R_CHECK_CUDA(cudaGetLastError());
    return PST3R_OK;
  }
  const long long total = (long long)B * (Hm / 8) * (Wm / 8) * (C / 2);
  center_pool8_kernel<<<blocks_for(total, 256), 256, 0, s>>>(reinterpret_cast<const bf16*>(feats), B, Hm, Wm, C,
                                                            reinterpret_cast<bf16*>(out));
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_attn_mask_bits(const float* logits_t, int64_t ld, int32_t Q, int32_t Nk, uint32_t* bits,
                                    pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(logits_t && bits && Q > 0 && Nk > 0 && ld >= Nk, "attn_mask_bits: bad args");
  const int words = ((Nk + 127) / 128) * 4;
  attn_mask_bits_kernel<<<Q, 256, 0, s>>>(logits_t, ld, Nk, bits, words);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_l2norm_rows(const float* x, int64_t ldx, void* y, int32_t y_kind, int64_t ldy, int32_t rows,
                                 int32_t cols, float eps, pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(x && y && rows > 0 && cols > 0 && y_kind >= 0 && y_kind <= 2, "l2norm_rows: bad args");
  l2norm_rows_kernel<<<blocks_for(rows, 8), 256, 0, s>>>(x, ldx, y, y_kind, ldy, rows, cols, eps);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_nhwc_to_nchw_f32(const void* x, int32_t x_kind, int32_t B, int32_t HW, int32_t C, float* y,
                                      pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(x && y && B > 0 && HW > 0 && C > 0 && (x_kind == PST3R_KIND_BF16 || x_kind == PST3R_KIND_SPLIT),
                  "nhwc_to_nchw_f32: bad args");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  nhwc_to_nchw_f32_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const bf16*>(x), x_kind, HW, C, y);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}
