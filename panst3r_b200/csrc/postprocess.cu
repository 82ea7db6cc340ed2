// Panoptic post-processing front half on the GPU (reference src/panst3r/engine/postprocess.py:18-27, 38-45, 63-120):
//   class scores            scores / labels = max over classes of sigmoid(class logits)                  (:39)
//   fused argmax            sigmoid(mask logits) -> bilinear resize (align_corners=False) -> score-weighted
//                           argmax over the kept queries, per pixel, plus the two per-query pixel counts the
//                           reference's filtering loop needs (>= 0.5 area; >= mask_threshold area that wins)   (:24-25, :63, :77-85)
//   finalize                segment ids / confidences from the winner map and a per-query lookup table     (:103-105)
// The reference materialises the (V, Q, H, W) fp32 probability tensor (629 MB at 16 views of 512x384, several
// copies of it) and walks the queries from Python with .item() calls; here every mask-logit plane is read once
// per round (HBM-bound: V*Qkept*h*w*4 bytes) and only (V, H, W) maps + 2*Qkept counters leave the kernel.
#include <math_constants.h>

#include "common.cuh"
#include "host_util.h"
#include "../../include/panst3r_b200.h"

namespace pst3r {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// one warp per query row; ties keep the first class (torch.max semantics on the sigmoid values)
__global__ void class_scores_kernel(const float* __restrict__ logits, long long ldl, int Q, int K,
                                    float* __restrict__ scores, int* __restrict__ labels) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int lane = threadIdx.x & 31;
  float best = -1.0f;
  int bi = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    const float s = sigmoid_f(logits[(long long)q * ldl + k]);
    if (s > best) { best = s; bi = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) { scores[q] = best; labels[q] = bi; }
}

// PyTorch's source index for align_corners=False (area_pixel_compute_source_index + guard_index_and_lambda)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.0f ? 0.0f : s;
  i0 = min((int)s, in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  l1 = fminf(fmaxf(s - (float)i0, 0.0f), 1.0f);
}

constexpr int PA_TW = 32, PA_TH = 32;   // output tile
constexpr int PA_ROWS = 4;              // output rows per thread (256 threads: 8 x 32)
constexpr int PA_SMAX = 36;             // max source-tile extent per axis (scale <= 1, i.e. the mask is never finer than the image)

struct PanArgmaxParams {
  const float* masks;            // [V][Q][hm][wm]
  long long view_stride, query_stride;
  int hm, wm, V;
  const int* keep_idx;           // [nkeep] query indices, in reference (ascending) order
  const float* keep_scores;      // [nkeep]
  int nkeep;
  int H, W;
  float scale_h, scale_w;        // (float)hm / H, (float)wm / W
  float mask_threshold;
  int* ids;                      // [V][out_view_stride], row pitch out_row_stride; index into keep_idx
  float* win;                    // winning query's mask probability
  long long out_view_stride;
  int out_row_stride;
  int* area_half;                // [nkeep] += #pixels with probability >= 0.5
  int* area_won;                 // [nkeep] += #pixels won with probability >= mask_threshold
  // band launches (pst3r_panoptic_argmax_band): this launch covers output rows [y0, y0 + rows) and `masks` holds the
  // source rows [src_row0, src_row0 + src_rows) of every plane (plane pitch = query_stride); whole-map launches: 0, H, 0
  int y0, rows, src_row0;
};

constexpr int PA_PLANE = PA_SMAX * PA_SMAX;                 // floats of one staged source window
constexpr int PA_STAGES = 8;                                // source windows in flight per CTA

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One CTA = one 32 x 32 output tile of one view, walking the kept queries in order.  The walk is a chain of dependent steps
// (every query needs its source window in shared memory), so its speed is set by how many windows are in flight: the first
// version fetched ONE window ahead through registers and paid a full DRAM / L2 latency per query (1.0 ms for the 314 MB of
// 16 views x 100 kept planes, 0.13 ms for a 96-CTA band of the lazy path whatever its size).  Now PA_STAGES windows are in
// flight as cp.async copies (LDGSTS, no registers held), each thread applies the sigmoid in place to the elements it copied
// itself (no barrier between arrival and transform), and ONE barrier per query publishes the window and retires the
// stage that is refilled next.  The arithmetic per pixel is unchanged.
__global__ void __launch_bounds__(256) panoptic_argmax_kernel(const PanArgmaxParams p) {
  extern __shared__ int smem_i[];
  int* h_half = smem_i;                       // [nkeep]
  int* h_won = smem_i + p.nkeep;              // [nkeep]
  int* idx_s = smem_i + 2 * p.nkeep;          // [nkeep] query index of kept query k
  float* sc_s = reinterpret_cast<float*>(smem_i + 3 * p.nkeep);  // [nkeep] its class score
  int* goff_s = smem_i + 4 * p.nkeep;         // [PA_PLANE] global offset of window element i (the same for every query)
  float* tile = reinterpret_cast<float*>(goff_s + PA_PLANE);     // [PA_STAGES][PA_PLANE]
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int v = blockIdx.z;
  const int x = blockIdx.x * PA_TW + tx;
  const int y_base = p.y0 + blockIdx.y * PA_TH;
  const int y_end = min(p.y0 + p.rows, p.H);
  for (int i = tid; i < 2 * p.nkeep; i += 256) smem_i[i] = 0;
  for (int i = tid; i < p.nkeep; i += 256) {
    idx_s[i] = __ldg(p.keep_idx + i);
    sc_s[i] = __ldg(p.keep_scores + i);
  }

  // source window of this output tile
  int sx0, sx1, sy0, sy1, d0, d1;
  float dl;
  src_index(blockIdx.x * PA_TW, p.scale_w, p.wm, sx0, d1, dl);
  src_index(min(blockIdx.x * PA_TW + PA_TW - 1, p.W - 1), p.scale_w, p.wm, d0, sx1, dl);
  src_index(y_base, p.scale_h, p.hm, sy0, d1, dl);
  src_index(min(y_base + PA_TH - 1, y_end - 1), p.scale_h, p.hm, d0, sy1, dl);
  const int sw = sx1 - sx0 + 1, sh = sy1 - sy0 + 1;
  const int n_el = sh * sw;

  // this thread's pixels: column x, rows y_base + ty + 8 * i
  int xi0, xi1;
  float xl1;
  src_index(min(x, p.W - 1), p.scale_w, p.wm, xi0, xi1, xl1);
  xi0 -= sx0; xi1 -= sx0;
  const float xl0 = 1.0f - xl1;
  int yo0[PA_ROWS], yo1[PA_ROWS];
  float yl1[PA_ROWS];
  bool valid[PA_ROWS];
#pragma unroll
  for (int i = 0; i < PA_ROWS; ++i) {
    const int y = y_base + ty + 8 * i;
    valid[i] = (x < p.W) && (y < y_end);
    int a, b;
    src_index(min(y, y_end - 1), p.scale_h, p.hm, a, b, yl1[i]);
    yo0[i] = (a - sy0) * sw;
    yo1[i] = (b - sy0) * sw;
  }
  float best[PA_ROWS], bestv[PA_ROWS];
  int bestk[PA_ROWS];
#pragma unroll
  for (int i = 0; i < PA_ROWS; ++i) { best[i] = -CUDART_INF_F; bestv[i] = 0.0f; bestk[i] = 0; }

  // a thread's share of a window: elements tid, tid + 256, ... < n_el (324 of them at a 2x resize: one or two per thread;
  // real loops, not PA_EPT predicated copies of the body: the kernel is issue bound)
  for (int i = tid; i < n_el; i += 256) {
    const int r = i / sw;
    goff_s[i] = r * p.wm + (i - r * sw);
  }
  const float* vbase = p.masks + (long long)v * p.view_stride + (long long)(sy0 - p.src_row0) * p.wm + sx0;
  auto issue = [&](int k) {  // window of kept query k -> stage k % PA_STAGES (asynchronous)
    const float* src = vbase + (long long)idx_s[k] * p.query_stride;
    float* dst = tile + (k % PA_STAGES) * PA_PLANE;
#pragma unroll 1
    for (int i = tid; i < n_el; i += 256) cp_async_f32(dst + i, src + goff_s[i]);
  };
  __syncthreads();  // idx_s / sc_s / goff_s / zeroed counters visible
#pragma unroll 1
  for (int k = 0; k < PA_STAGES - 1; ++k) {
    if (k < p.nkeep) issue(k);
    cp_async_commit();
  }
#pragma unroll 1
  for (int k = 0; k < p.nkeep; ++k) {
    cp_async_wait<PA_STAGES - 2>();  // this thread's copies of window k have landed ...
    float* cur = tile + (k % PA_STAGES) * PA_PLANE;
#pragma unroll 1
    for (int i = tid; i < n_el; i += 256) cur[i] = sigmoid_f(cur[i]);  // ... and become probabilities in place
    __syncthreads();  // window k complete for everybody; everybody is done reading window k - 1
    if (k + PA_STAGES - 1 < p.nkeep) issue(k + PA_STAGES - 1);  // refills the stage window k - 1 occupied
    cp_async_commit();
    const float sc = sc_s[k];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < PA_ROWS; ++i) {
      const float t0 = xl0 * cur[yo0[i] + xi0] + xl1 * cur[yo0[i] + xi1];
      const float t1 = xl0 * cur[yo1[i] + xi0] + xl1 * cur[yo1[i] + xi1];
      const float val = (1.0f - yl1[i]) * t0 + yl1[i] * t1;
      const float pr = sc * val;
      if (pr > best[i]) { best[i] = pr; bestv[i] = val; bestk[i] = k; }  // strict: ties keep the first query
      cnt += (valid[i] && val >= 0.5f) ? 1 : 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (tx == 0 && cnt) atomicAdd(&h_half[k], cnt);
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < PA_ROWS; ++i) {
    if (!valid[i]) continue;
    const long long o = (long long)v * p.out_view_stride + (long long)(y_base + ty + 8 * i) * p.out_row_stride + x;
    p.ids[o] = bestk[i];
    p.win[o] = bestv[i];
    if (p.nkeep > 0 && bestv[i] >= p.mask_threshold) atomicAdd(&h_won[bestk[i]], 1);
  }
  __syncthreads();
  for (int i = tid; i < p.nkeep; i += 256) {
    if (h_half[i]) atomicAdd(&p.area_half[i], h_half[i]);
    if (h_won[i]) atomicAdd(&p.area_won[i], h_won[i]);
  }
}

__global__ void panoptic_finalize_kernel(const int* __restrict__ ids, const float* __restrict__ win,
                                         const int* __restrict__ lut, int nkeep, float thr, float void_conf,
                                         int* __restrict__ pan, float* __restrict__ conf, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = ids[i];
  const float w = win[i];
  const int seg = (nkeep > 0 && k < nkeep && w >= thr) ? lut[k] : 0;
  pan[i] = seg;
  conf[i] = seg ? w : void_conf;
}

}  // namespace pst3r

using namespace pst3r;

extern "C" int pst3r_class_scores(const float* logits, int64_t ldl, int32_t Q, int32_t K, float* scores, int32_t* labels,
                                  pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(logits && scores && labels && Q > 0 && K > 0 && ldl >= K, "class_scores: bad args");
  class_scores_kernel<<<(Q + 7) / 8, 256, 0, s>>>(logits, ldl, Q, K, scores, labels);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

// the kernel's source-row arithmetic on the host (same fp32 operations): first / last source row read for an output row
static void host_src_rows(int dst, float scale, int in_size, int* i0, int* i1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.0f ? 0.0f : s;
  *i0 = (int)s < in_size - 1 ? (int)s : in_size - 1;
  *i1 = *i0 + 1 < in_size - 1 ? *i0 + 1 : in_size - 1;
}

extern "C" int pst3r_panoptic_argmax_band(const float* masks, int64_t view_stride, int64_t query_stride, int32_t V, int32_t hm,
                                          int32_t wm, int32_t src_row0, int32_t src_rows, const int32_t* keep_idx,
                                          const float* keep_scores, int32_t nkeep, int32_t H, int32_t W, int32_t y0,
                                          int32_t rows, float mask_threshold, int32_t* ids, float* win,
                                          int64_t out_view_stride, int32_t out_row_stride, int32_t* area_half,
                                          int32_t* area_won, pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(masks && ids && win && V > 0 && hm > 0 && wm > 0 && H > 0 && W > 0 && nkeep >= 0,
                  "panoptic_argmax: bad args");
  PST3R_CHECK_ARG(nkeep == 0 || (keep_idx && keep_scores && area_half && area_won), "panoptic_argmax: null keep list / counters");
  PST3R_CHECK_ARG(hm <= H && wm <= W, "panoptic_argmax: the mask grid (%d x %d) must not be finer than the image (%d x %d)",
                  hm, wm, H, W);
  PST3R_CHECK_ARG(out_row_stride >= W && out_view_stride >= (int64_t)out_row_stride * H, "panoptic_argmax: bad output strides");
  PST3R_CHECK_ARG(nkeep <= 4096, "panoptic_argmax: at most 4096 kept queries");
  PST3R_CHECK_ARG(y0 >= 0 && rows > 0 && y0 + rows <= H && src_row0 >= 0 && src_rows > 0 && src_row0 + src_rows <= hm,
                  "panoptic_argmax: bad band (rows [%d, %d) of %d, source rows [%d, %d) of %d)", y0, y0 + rows, H, src_row0,
                  src_row0 + src_rows, hm);
  PST3R_CHECK_ARG(query_stride >= (int64_t)src_rows * wm, "panoptic_argmax: plane pitch smaller than the band");
  const float scale_h = (float)hm / (float)H, scale_w = (float)wm / (float)W;
  {
    int lo, hi, d;
    host_src_rows(y0, scale_h, hm, &lo, &d);
    host_src_rows(y0 + rows - 1, scale_h, hm, &d, &hi);
    PST3R_CHECK_ARG(lo >= src_row0 && hi < src_row0 + src_rows,
                    "panoptic_argmax: output rows [%d, %d) read source rows [%d, %d], the band holds [%d, %d)", y0, y0 + rows, lo,
                    hi, src_row0, src_row0 + src_rows);
  }
  PanArgmaxParams p;
  p.masks = masks; p.view_stride = view_stride; p.query_stride = query_stride;
  p.hm = hm; p.wm = wm; p.V = V;
  p.keep_idx = keep_idx; p.keep_scores = keep_scores; p.nkeep = nkeep;
  p.H = H; p.W = W;
  p.scale_h = scale_h; p.scale_w = scale_w;
  p.mask_threshold = mask_threshold;
  p.ids = ids; p.win = win; p.out_view_stride = out_view_stride; p.out_row_stride = out_row_stride;
  p.area_half = area_half; p.area_won = area_won;
  p.y0 = y0; p.rows = rows; p.src_row0 = src_row0;
  const size_t smem = (size_t)4 * nkeep * sizeof(int) + (size_t)(PA_STAGES + 1) * PA_PLANE * sizeof(float);
  static bool configured = false;
  if (!configured) {  // up to 4096 kept queries: 64 KB of per-query state + 45.6 KB of staged windows and offsets
    PST3R_CHECK_CUDA(cudaFuncSetAttribute(panoptic_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          4 * 4096 * (int)sizeof(int) + (PA_STAGES + 1) * PA_PLANE * (int)sizeof(float)));
    configured = true;
  }
  dim3 grid((W + PA_TW - 1) / PA_TW, (rows + PA_TH - 1) / PA_TH, V);
  panoptic_argmax_kernel<<<grid, 256, smem, s>>>(p);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}

extern "C" int pst3r_panoptic_argmax(const float* masks, int64_t view_stride, int64_t query_stride, int32_t V, int32_t hm,
                                     int32_t wm, const int32_t* keep_idx, const float* keep_scores, int32_t nkeep,
                                     int32_t H, int32_t W, float mask_threshold, int32_t* ids, float* win,
                                     int64_t out_view_stride, int32_t out_row_stride, int32_t* area_half,
                                     int32_t* area_won, pst3r_stream_t s_) {
  PST3R_CHECK_ARG(H > 0 && hm > 0, "panoptic_argmax: bad args");
  return pst3r_panoptic_argmax_band(masks, view_stride, query_stride, V, hm, wm, 0, hm, keep_idx, keep_scores, nkeep, H, W, 0, H,
                                    mask_threshold, ids, win, out_view_stride, out_row_stride, area_half, area_won, s_);
}

extern "C" int pst3r_panoptic_finalize(const int32_t* ids, const float* win, const int32_t* lut, int32_t nkeep,
                                       float mask_threshold, float void_confidence, int32_t* pan, float* conf, int64_t n,
                                       pst3r_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  PST3R_CHECK_ARG(ids && win && pan && conf && n > 0 && (nkeep == 0 || lut), "panoptic_finalize: bad args");
  panoptic_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ids, win, lut, nkeep, mask_threshold, void_confidence,
                                                                      pan, conf, n);
  PST3R_CHECK_CUDA(cudaGetLastError());
  return PST3R_OK;
}
