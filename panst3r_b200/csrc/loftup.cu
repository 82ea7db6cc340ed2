// LoftUp guidance path (reference src/panst3r/model/upscalers/loftup.py:9-79, 122-130, 152-157):
//   bilinear x0.5 of the image (exact 2x2 mean) -> batch-global per-channel MinMaxScaler -> ImplicitFeaturizer
//   (sin/cos Fourier features of [gy, gx, r, g, b] at n_freqs frequencies + the scaled image) -> GroupNorm(1, C)
//   -> [conv3x3 + GroupNorm(8) + ReLU] x 2  (the convolutions run as implicit GEMMs in gemm.cu).
// All reductions are two-level with a fixed summation order: results are deterministic run to run.
#include "common.cuh"
#include "host_util.h"
#include "../../include/panst3r_b200.h"

namespace pst3r {

// ---------------------------------------------------------------------------------------------------
// 2x2 mean + per-block channel min / max partials
// ---------------------------------------------------------------------------------------------------
__global__ void loftup_half_kernel(const float* __restrict__ img, int V, int H, int W, float* __restrict__ half,
                                   float* __restrict__ partial /* [gridDim.x][3][2] */) {
  const int Hh = H / 2, Wh = W / 2;
  const long long total = (long long)V * Hh * Wh;
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = idx % Wh;
    const int y = (idx / Wh) % Hh;
    const int v = idx / ((long long)Wh * Hh);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* p = img + (((long long)v * 3 + c) * H + 2 * y) * W + 2 * x;
      // F.interpolate(scale_factor=0.5, bilinear, align_corners=False): source 2x+0.5 -> weights (.5,.5) per axis
      const float top = 0.5f * p[0] + 0.5f * p[1];
      const float bot = 0.5f * p[W] + 0.5f * p[W + 1];
      const float val = 0.5f * top + 0.5f * bot;
      half[(((long long)v * 3 + c) * Hh + y) * Wh + x] = val;
      mn[c] = fminf(mn[c], val);
      mx[c] = fmaxf(mx[c], val);
    }
  }
  __shared__ float s_mn[3][32], s_mx[3][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float a = mn[c], b = mx[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (lane == 0) { s_mn[c][warp] = a; s_mx[c][warp] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float a = 3.4e38f, b = -3.4e38f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a = fminf(a, s_mn[threadIdx.x][w]); b = fmaxf(b, s_mx[threadIdx.x][w]); }
    partial[(blockIdx.x * 3 + threadIdx.x) * 2 + 0] = a;
    partial[(blockIdx.x * 3 + threadIdx.x) * 2 + 1] = b;
  }
}

__global__ void loftup_minmax_final_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ minmax /* [3][2] */) {
  if (threadIdx.x < 3) {
    float a = 3.4e38f, b = -3.4e38f;
    for (int i = 0; i < nblocks; ++i) {
      a = fminf(a, partial[(i * 3 + threadIdx.x) * 2 + 0]);
      b = fmaxf(b, partial[(i * 3 + threadIdx.x) * 2 + 1]);
    }
    minmax[threadIdx.x * 2 + 0] = a;
    minmax[threadIdx.x * 2 + 1] = b;
  }
}

// ---------------------------------------------------------------------------------------------------
// ImplicitFeaturizer channel c of pixel (v, y, x);  channel order: sin[f*5+m], cos[f*5+m], scaled rgb
// ---------------------------------------------------------------------------------------------------
struct FourierArgs {
  const float* half;    // [V,3,Hh,Wh]
  const float* minmax;  // [3][2]
  const float* gy;      // [Hh]  torch.linspace(-1, 1, Hh)
  const float* gx;      // [Wh]
  const float* freqs;   // [n_freqs] exp(linspace(-2, 10, n_freqs))
  const float* biases;  // flat parameter (2, 5, n_freqs), consumed as [2][n_freqs][5] (reshape, loftup.py:62-63)
  int V, Hh, Wh, n_freqs;
};

__device__ __forceinline__ void fourier_base(const FourierArgs& a, int v, int y, int x, float (&base)[5]) {
  base[0] = a.gy[y];
  base[1] = a.gx[x];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float val = a.half[(((long long)v * 3 + c) * a.Hh + y) * a.Wh + x];
    const float lo = a.minmax[2 * c], hi = a.minmax[2 * c + 1];
    base[2 + c] = (val - lo) / fmaxf(hi - lo, 1e-4f) - 0.5f;
  }
}

// per-block partial (sum, sum of squares) of all C feature channels over a pixel chunk of view blockIdx.y
__global__ void loftup_fourier_stats_kernel(FourierArgs a, double* __restrict__ partial /* [V][gridDim.x][2] */) {
  const int v = blockIdx.y;
  const int npix = a.Hh * a.Wh;
  const int nf = a.n_freqs;
  double s = 0.0, ss = 0.0;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    float base[5];
    fourier_base(a, v, pix / a.Wh, pix % a.Wh, base);
    float ls = 0.0f, lss = 0.0f;
    for (int f = 0; f < nf; ++f) {
      const float fr = a.freqs[f];
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const float t = base[m] * fr;
        const float sv = sinf(t + a.biases[f * 5 + m]);
        const float cv = cosf(t + a.biases[nf * 5 + f * 5 + m]);
        ls += sv + cv;
        lss += sv * sv + cv * cv;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { ls += base[2 + c]; lss += base[2 + c] * base[2 + c]; }
    s += ls;
    ss += lss;
  }
  __shared__ double sh[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0.0, a1 = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a0 += sh[0][w]; a1 += sh[1][w]; }
    partial[((long long)v * gridDim.x + blockIdx.x) * 2 + 0] = a0;
    partial[((long long)v * gridDim.x + blockIdx.x) * 2 + 1] = a1;
  }
}

// (mean, rstd) per group from per-block partials: stats[g] = {mean, rstd};  count = elements per group
__global__ void group_stats_final_kernel(const double* __restrict__ partial, int nparts, double count, float eps,
                                         float* __restrict__ stats /* [groups][2] */, int groups) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  double s = 0.0, ss = 0.0;
  for (int i = 0; i < nparts; ++i) {
    s += partial[((long long)g * nparts + i) * 2 + 0];
    ss += partial[((long long)g * nparts + i) * 2 + 1];
  }
  const double mean = s / count;
  const double var = fmax(ss / count - mean * mean, 0.0);
  stats[2 * g + 0] = (float)mean;
  stats[2 * g + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// features -> GroupNorm(1, C) -> bf16 NHWC [V, Hh, Wh, ldo] (channels [C, ldo) zero)
// split != 0: pixel rows are [hi(ldo) | lo(ldo)] (reference-precision head, PST3R_KIND_SPLIT)
__device__ __forceinline__ void put_bf16(bf16* o, int c, int lo_off, int split, float v) {
  const bf16 h = __float2bfloat16(v);
  o[c] = h;
  if (split) o[c + lo_off] = __float2bfloat16(v - __bfloat162float(h));
}

__global__ void loftup_fourier_write_kernel(FourierArgs a, const float* __restrict__ stats /* [V][2] */,
                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                            bf16* __restrict__ out, int ldo, int split) {
  const int v = blockIdx.y;
  const int npix = a.Hh * a.Wh;
  const int nf = a.n_freqs;
  const int C = 10 * nf + 3;
  const float mean = stats[2 * v], rstd = stats[2 * v + 1];
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    float base[5];
    fourier_base(a, v, pix / a.Wh, pix % a.Wh, base);
    bf16* o = out + ((long long)v * npix + pix) * (split ? 2 * ldo : ldo);
    for (int f = 0; f < nf; ++f) {
      const float fr = a.freqs[f];
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const float t = base[m] * fr;
        const int cs = f * 5 + m, cc = nf * 5 + f * 5 + m;
        const float sv = sinf(t + a.biases[cs]);
        const float cv = cosf(t + a.biases[cc]);
        put_bf16(o, cs, ldo, split, (sv - mean) * rstd * gamma[cs] + beta[cs]);
        put_bf16(o, cc, ldo, split, (cv - mean) * rstd * gamma[cc] + beta[cc]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int ch = 10 * nf + c;
      put_bf16(o, ch, ldo, split, (base[2 + c] - mean) * rstd * gamma[ch] + beta[ch]);
    }
    for (int ch = C; ch < ldo; ++ch) put_bf16(o, ch, ldo, split, 0.0f);
  }
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm on pixel-major bf16 maps [V, npix, C]: per-(view, group) stats, then normalise + affine (+ReLU) in place
// ---------------------------------------------------------------------------------------------------
__global__ void groupnorm_stats_kernel(const bf16* __restrict__ x, int npix, int C, int groups,
                                       double* __restrict__ partial /* [V*groups][gridDim.x][2] */, int split) {
  const int v = blockIdx.y;
  const int cpg = C / groups;
  // thread -> fixed channel pair; rows strided: coalesced 4-byte loads along channels
  const int cv = C >> 1;
  const int tc = threadIdx.x % cv;                // channel pair index (blockDim.x is a multiple of C/2)
  const int rows_per_iter = blockDim.x / cv;
  const int r0 = threadIdx.x / cv;
  const int chunk = (npix + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * chunk, p1 = min(npix, p0 + chunk);
  float s = 0.0f, ss = 0.0f;
  const int ld = split ? 2 * C : C;
  for (int p = p0 + r0; p < p1; p += rows_per_iter) {
    const bf16* px = x + ((long long)v * npix + p) * ld + 2 * tc;
    float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(px));
    if (split) {
      const float2 l = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(px + C));
      f.x += l.x; f.y += l.y;
    }
    s += f.x + f.y;
    ss += f.x * f.x + f.y * f.y;
  }
  extern __shared__ float sh[];  // [blockDim.x][2]
  sh[2 * threadIdx.x] = s;
  sh[2 * threadIdx.x + 1] = ss;
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    double a0 = 0.0, a1 = 0.0;
    for (int r = 0; r < rows_per_iter; ++r)
      for (int c = g * cpg / 2; c < (g + 1) * cpg / 2; ++c) {
        a0 += sh[2 * (r * cv + c)];
        a1 += sh[2 * (r * cv + c) + 1];
      }
    partial[(((long long)v * groups + g) * gridDim.x + blockIdx.x) * 2 + 0] = a0;
    partial[(((long long)v * groups + g) * gridDim.x + blockIdx.x) * 2 + 1] = a1;
  }
}

__global__ void groupnorm_apply_kernel(bf16* __restrict__ x, int npix, int C, int groups, const float* __restrict__ stats,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, int relu, long long total2,
                                       int split) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total2) return;
  const int cv = C >> 1;
  const int c = (idx % cv) * 2;
  const long long row = idx / cv;
  const int v = row / npix;
  const int g = c / (C / groups);
  const float mean = stats[2 * (v * groups + g)], rstd = stats[2 * (v * groups + g) + 1];
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(x + row * (split ? 2 * C : C) + c);
  float2 f = __bfloat1622float2(*p);
  if (split) {
    const float2 l = __bfloat1622float2(p[C / 2]);
    f.x += l.x; f.y += l.y;
  }
  f.x = (f.x - mean) * rstd * gamma[c] + beta[c];
  f.y = (f.y - mean) * rstd * gamma[c + 1] + beta[c + 1];
  if (relu) { f.x = fmaxf(f.x, 0.0f); f.y = fmaxf(f.y, 0.0f); }
  const __nv_bfloat162 h = __floats2bfloat162_rn(f.x, f.y);
  *p = h;
  if (split) {
    const float2 hf = __bfloat1622float2(h);
    p[C / 2] = __floats2bfloat162_rn(f.x - hf.x, f.y - hf.y);
  }
}

}  // namespace pst3r

using namespace pst3r;

extern "C" int64_t pst3r_loftup_workspace_bytes(int32_t V, int32_t C, int32_t groups) {
  // min/max partials (256 blocks x 3 x 2 floats) | minmax (8 floats) | double partials max(V*64, V*groups*64) x 2 | stats
  const long long parts = (long long)V * (groups > 1 ? groups : 1) * 64;
  return 256 * 6 * 4 + 32 + parts * 16 + (long long)V * (groups > 1 ? groups : 1) * 8 + 64;
}

extern "C" int pst3r_loftup_guidance(const float* img, int32_t V, int32_t H, int32_t W, float* half, float* minmax,
                                     void* workspace, pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(img && half && minmax && workspace && V > 0 && (H % 2) == 0 && (W % 2) == 0, "loftup_guidance: bad args");
  float* partial = reinterpret_cast<float*>(workspace);
  const int nblocks = 256;
  loftup_half_kernel<<<nblocks, 256, 0, s>>>(img, V, H, W, half, partial);
  PST3R_CHECK_CUDA(cudaGetLastError());
  loftup_minmax_final_kernel<<<1, 32, 0, s>>>(partial, nblocks, minmax);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_loftup_fourier_gn(const float* half, const float* minmax, const float* gy, const float* gx,
                                       const float* freqs, const float* biases, int32_t V, int32_t Hh, int32_t Wh,
                                       int32_t n_freqs, const float* gamma, const float* beta, float eps, void* out,
                                       int32_t out_kind, int64_t ldo, void* workspace, pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(half && minmax && gy && gx && freqs && biases && gamma && beta && out && workspace && V > 0 &&
                      ldo >= 10 * n_freqs + 3 && (out_kind == PST3R_KIND_BF16 || out_kind == PST3R_KIND_SPLIT),
                  "loftup_fourier_gn: bad args");
  FourierArgs a{half, minmax, gy, gx, freqs, biases, V, Hh, Wh, n_freqs};
  const int C = 10 * n_freqs + 3;
  const int nb = 64;
  double* partial = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + 256 * 6 * 4 + 32);
  float* stats = reinterpret_cast<float*>(partial + (long long)V * nb * 2);
  loftup_fourier_stats_kernel<<<dim3(nb, V), 256, 0, s>>>(a, partial);
  PST3R_CHECK_CUDA(cudaGetLastError());
  group_stats_final_kernel<<<(V + 63) / 64, 64, 0, s>>>(partial, nb, (double)C * Hh * Wh, eps, stats, V);
  PST3R_CHECK_CUDA(cudaGetLastError());
  loftup_fourier_write_kernel<<<dim3(nb * 4, V), 256, 0, s>>>(a, stats, gamma, beta, reinterpret_cast<bf16*>(out), (int)ldo,
                                                                   out_kind == PST3R_KIND_SPLIT);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_groupnorm_nhwc(void* x, int32_t kind, int32_t V, int32_t npix, int32_t C, int32_t groups, const float* gamma,
                                    const float* beta, float eps, int32_t relu, void* workspace, pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(x && gamma && beta && workspace && V > 0 && npix > 0 && C > 0 && groups > 0 && (C % (2 * groups)) == 0 &&
                      C / 2 <= 512 && groups <= 32 && (kind == PST3R_KIND_BF16 || kind == PST3R_KIND_SPLIT),
                  "groupnorm_nhwc: bad args");
  const int split = kind == PST3R_KIND_SPLIT;
  const int nb = 64;
  const int cv = C / 2;
  const int threads = (512 / cv) * cv;  // multiple of C/2, <= 512
  double* partial = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + 256 * 6 * 4 + 32);
  float* stats = reinterpret_cast<float*>(partial + (long long)V * groups * nb * 2);
  groupnorm_stats_kernel<<<dim3(nb, V), threads, threads * 2 * sizeof(float), s>>>(reinterpret_cast<const bf16*>(x), npix, C, groups, partial, split);
  PST3R_CHECK_CUDA(cudaGetLastError());
  group_stats_final_kernel<<<(V * groups + 63) / 64, 64, 0, s>>>(partial, nb, (double)(C / groups) * npix, eps, stats, V * groups);
  PST3R_CHECK_CUDA(cudaGetLastError());
  const long long total2 = (long long)V * npix * cv;
  groupnorm_apply_kernel<<<(unsigned)((total2 + 255) / 256), 256, 0, s>>>(reinterpret_cast<bf16*>(x), npix, C, groups, stats, gamma, beta, relu, total2, split);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}
