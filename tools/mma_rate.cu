// tcgen05.mma issue-rate microbenchmark (development tool, not part of the library).
//
// One CTA; one thread issues sequences of M = 128 bf16 MMAs (cta_group::1) in the shapes the attention kernel uses and
// measures clock64() from the first issue to the arrival of the closing tcgen05.commit.  Answers: how many cycles does an
// MMA of N = 64 / 128 / 256 really take, SS vs TS (A from tensor memory), K-major vs MN-major B, when consecutive
// instructions accumulate into the SAME tensor-memory accumulator and when they alternate between independent ones.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/mma_rate tools/mma_rate.cu -lcuda && tools/_bin/mma_rate
#include <string>
#include <cuda_fp16.h>

#include "../panst3r_b200/csrc/common.cuh"

using namespace pst3r;

constexpr int NCASE = 24;
constexpr uint32_t OFF_A = 0;            // 128 x 64 bf16, K-major SW128 (Q)
constexpr uint32_t OFF_B = 16384;        // 256 x 64 bf16, K-major SW128 (K, two 128-row atoms)
constexpr uint32_t OFF_V = 16384 + 32768;  // 128 keys x 64 bf16, MN-major SW128 (V)
constexpr uint32_t OFF_BAR = OFF_V + 16384;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 64 + 1024;

struct Result { long long total, issued; int n; };

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Result* res, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (uint32_t i = threadIdx.x; i < OFF_BAR / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  // zero tensor memory (A operands of the TS cases are read from it)
  {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0;
    const uint32_t la = tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    for (int c = 0; c < 512; c += 32) tmem_st32(la + c, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem + OFF_A), b_addr = smem_u32(smem + OFF_B), v_addr = smem_u32(smem + OFF_V);
    constexpr uint32_t id256 = make_idesc_bf16(128, 256, 0, 0);
    constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t id64k = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t id64mn = make_idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t id128mn = make_idesc_bf16(128, 128, 0, 1);
    constexpr uint32_t id32mn = make_idesc_bf16(128, 32, 0, 1);
    auto adesc = [&](int ks) { return make_smem_desc_sw128(a_addr + (ks & 3) * 32, 0, 1024); };
    auto bdesc = [&](int ks) { return make_smem_desc_sw128(b_addr + (ks & 3) * 32, 0, 1024); };
    auto vdesc = [&](int ks) { return make_smem_desc_sw128(v_addr + (ks & 7) * 2048, 16384, 1024); };
    uint32_t phase = 0;
    for (int c = 0; c < NCASE; ++c) {
      int n = 0;
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        switch (c) {
          case 0: for (int k = 0; k < 8; ++k) { umma_ss(tm, adesc(k), bdesc(k), id256, 1); ++n; } break;
          case 1: for (int k = 0; k < 8; ++k) { umma_ss(tm, adesc(k), bdesc(k), id128, 1); ++n; } break;
          case 2: for (int k = 0; k < 8; ++k) { umma_ss(tm + (k & 1) * 128, adesc(k), bdesc(k), id128, 1); ++n; } break;
          case 3: for (int k = 0; k < 8; ++k) { umma_ss(tm, adesc(k), bdesc(k), id64k, 1); ++n; } break;
          case 4: for (int k = 0; k < 8; ++k) { umma_ss(tm + (k & 1) * 64, adesc(k), bdesc(k), id64k, 1); ++n; } break;
          case 5: for (int k = 0; k < 8; ++k) { umma_ss(tm + (k & 3) * 64, adesc(k), bdesc(k), id64k, 1); ++n; } break;
          case 6: for (int k = 0; k < 8; ++k) { umma_ss(tm, adesc(k), vdesc(k), id64mn, 1); ++n; } break;
          case 7: for (int k = 0; k < 8; ++k) { umma_ts(tm + 256, tm + 384 + k * 8, vdesc(k), id64mn, 1); ++n; } break;
          case 8: for (int k = 0; k < 8; ++k) { umma_ts(tm + 256 + (k & 1) * 64, tm + 384 + k * 8, vdesc(k), id64mn, 1); ++n; } break;
          case 9: for (int k = 0; k < 8; ++k) { umma_ts(tm + (k & 3) * 64, tm + 384 + k * 8, vdesc(k), id64mn, 1); ++n; } break;
          case 10:  // TS with a K-major B operand
            for (int k = 0; k < 8; ++k) { umma_ts(tm + 256, tm + 384 + k * 8, bdesc(k), id64k, 1); ++n; } break;
          case 11:  // the attention kernel's order: 4 x S (N = 128), then 8 x PV (TS, N = 64)
            for (int k = 0; k < 4; ++k) { umma_ss(tm, adesc(k), bdesc(k), id128, k != 0); ++n; }
            for (int k = 0; k < 8; ++k) { umma_ts(tm + 256, tm + 384 + k * 8, vdesc(k), id64mn, k != 0); ++n; }
            break;
          case 12:  // the same twelve, interleaved PV PV S
            for (int k = 0; k < 4; ++k) {
              umma_ts(tm + 256, tm + 384 + (2 * k) * 8, vdesc(2 * k), id64mn, k != 0); ++n;
              umma_ts(tm + 256, tm + 384 + (2 * k + 1) * 8, vdesc(2 * k + 1), id64mn, 1); ++n;
              umma_ss(tm, adesc(k), bdesc(k), id128, k != 0); ++n;
            }
            break;
          case 13:  // as 11 but the two sub-tiles' work back to back (S_0, PV_0, S_1, PV_1 with separate accumulators)
            for (int t = 0; t < 2; ++t) {
              for (int k = 0; k < 4; ++k) { umma_ss(tm + t * 128, adesc(k), bdesc(k), id128, k != 0); ++n; }
              for (int k = 0; k < 8; ++k) { umma_ts(tm + 256 + t * 64, tm + 384 + t * 64 + k * 8, vdesc(k), id64mn, k != 0); ++n; }
            }
            break;
          case 14:  // PV as TS with N = 128 (two heads' worth of columns; would need hd = 128 or two V tiles side by side)
            for (int k = 0; k < 8; ++k) { umma_ts(tm, tm + 384 + k * 8, make_smem_desc_sw128(v_addr + (k & 3) * 2048, 8192, 1024), id128mn, 1); ++n; } break;
          case 15:  // PV as SS (P in shared memory, K-major A), N = 64 MN-major B
            for (int k = 0; k < 8; ++k) { umma_ss(tm + 256, adesc(k), vdesc(k), id64mn, 1); ++n; } break;
          case 16:  // S with accumulate = 0 on every instruction (no read of the accumulator)
            for (int k = 0; k < 8; ++k) { umma_ss(tm, adesc(k), bdesc(k), id128, 0); ++n; } break;
          case 17:  // PV TS, accumulate = 0
            for (int k = 0; k < 8; ++k) { umma_ts(tm + 256, tm + 384 + k * 8, vdesc(k), id64mn, 0); ++n; } break;
          case 18:  // PV TS N = 32 (half the columns), 2 accumulators
            for (int k = 0; k < 8; ++k) { umma_ts(tm + 256 + (k & 1) * 32, tm + 384 + k * 8, vdesc(k), id32mn, 1); ++n; } break;
          case 19:  // S N = 128 with the SAME descriptors every time (operands stay put)
            for (int k = 0; k < 8; ++k) { umma_ss(tm, adesc(0), bdesc(0), id128, 1); ++n; } break;
          case 20:  // PV TS same A columns / same V rows every time
            for (int k = 0; k < 8; ++k) { umma_ts(tm + 256, tm + 384, vdesc(0), id64mn, 1); ++n; } break;
          case 21:  // S N = 256 alternating accumulators
            for (int k = 0; k < 8; ++k) { umma_ss(tm + (k & 1) * 256, adesc(k), bdesc(k), id256, 1); ++n; } break;
          case 22:  // commit after every 4 S MMAs (as the kernel does per tile), same accumulator otherwise as case 1
            for (int k = 0; k < 8; ++k) { umma_ss(tm, adesc(k), bdesc(k), id128, 1); ++n; if ((k & 3) == 3) umma_commit(bar + 1); } break;
          default:  // 23: PV TS, 8 per commit
            for (int k = 0; k < 8; ++k) { umma_ts(tm + 256, tm + 384 + k * 8, vdesc(k), id64mn, 1); ++n; }
            umma_commit(bar + 1);
            break;
        }
      }
      const long long t1 = clock64();
      umma_commit(bar);
      mbar_wait(bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      res[c].total = t2 - t0;
      res[c].issued = t1 - t0;
      res[c].n = n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// Two issuing threads (lane 0 of warps 0 and 1), independent accumulators: is the ~80-cycle minimum per instruction a limit
// of the issuing thread or of the tensor core?
__global__ void __launch_bounds__(128, 1) mma_two_issuers_kernel(Result* res, int reps, int n_issuers, int ts) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (uint32_t i = threadIdx.x; i < OFF_BAR / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < n_issuers) {
    const uint32_t a_addr = smem_u32(smem + OFF_A), b_addr = smem_u32(smem + OFF_B), v_addr = smem_u32(smem + OFF_V);
    constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t id64mn = make_idesc_bf16(128, 64, 0, 1);
    int n = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < 8; ++k) {
        if (ts) umma_ts(tm + 256 + w * 64, tm + 384 + w * 64 + k * 8, make_smem_desc_sw128(v_addr + (k & 7) * 2048, 16384, 1024), id64mn, 1);
        else umma_ss(tm + w * 128, make_smem_desc_sw128(a_addr + (k & 3) * 32, 0, 1024), make_smem_desc_sw128(b_addr + (k & 3) * 32, 0, 1024), id128, 1);
        ++n;
      }
    const long long t1 = clock64();
    umma_commit(bar + w);
    mbar_wait(bar + w, 0);
    const long long t2 = clock64();
    res[w].total = t2 - t0; res[w].issued = t1 - t0; res[w].n = n;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// Softmax-sweep instruction mixes on register data: which pipe bounds the exp sweep?  `warps_per_smsp` warps per SM sub-partition
// each run `iters` x 32 elements of: 0 = MUFU.EX2 only; 1 = FFMA2 + EX2; 2 = FFMA2 + EX2 + FADD2 (row sum);
// 3 = the full sweep (+ F2FP bf16x2 pack); 4 = full sweep with a PRMT (truncating) pack instead of F2FP; 5 = F2FP only.
template <int mix>
__global__ void __launch_bounds__(512, 1) sweep_rate_kernel(long long* out, float* sink, int iters) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = -0.01f * (float)((threadIdx.x + i) & 63);
  float2 sA = make_float2(0.f, 0.f), sB = make_float2(0.f, 0.f);
  uint32_t acc = 0;
  const float2 c2 = make_float2(0.999f, 0.999f), nm2 = make_float2(-0.001f, -0.001f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float2 x01 = make_float2(x[i], x[i + 1]), x23 = make_float2(x[i + 2], x[i + 3]);
      if ((mix >= 1 && mix <= 4) || mix >= 6) { x01 = __ffma2_rn(x01, c2, nm2); x23 = __ffma2_rn(x23, c2, nm2); }
      float e0 = x01.x, e1 = x01.y, e2 = x23.x, e3 = x23.y;
      if (mix == 6 || mix == 7) {
        // packed half-precision exponentials: one MUFU instruction per PAIR
        uint32_t p01, p23;
        if (mix == 6) {
          asm("{ .reg .b32 t; cvt.rn.f16x2.f32 t, %2, %1; ex2.approx.f16x2 %0, t; }" : "=r"(p01) : "f"(e0), "f"(e1));
          asm("{ .reg .b32 t; cvt.rn.f16x2.f32 t, %2, %1; ex2.approx.f16x2 %0, t; }" : "=r"(p23) : "f"(e2), "f"(e3));
          acc ^= p01 ^ p23;
          e0 = __half2float(__ushort_as_half((unsigned short)(p01 & 0xffff))); e1 = __half2float(__ushort_as_half((unsigned short)(p01 >> 16)));
          e2 = __half2float(__ushort_as_half((unsigned short)(p23 & 0xffff))); e3 = __half2float(__ushort_as_half((unsigned short)(p23 >> 16)));
        } else {
          asm("{ .reg .b32 t; cvt.rn.bf16x2.f32 t, %2, %1; ex2.approx.ftz.bf16x2 %0, t; }" : "=r"(p01) : "f"(e0), "f"(e1));
          asm("{ .reg .b32 t; cvt.rn.bf16x2.f32 t, %2, %1; ex2.approx.ftz.bf16x2 %0, t; }" : "=r"(p23) : "f"(e2), "f"(e3));
          acc ^= p01 ^ p23;
          e0 = __uint_as_float(p01 << 16); e1 = __uint_as_float(p01 & 0xffff0000u);
          e2 = __uint_as_float(p23 << 16); e3 = __uint_as_float(p23 & 0xffff0000u);
        }
        sA = __fadd2_rn(sA, make_float2(e0, e1)); sB = __fadd2_rn(sB, make_float2(e2, e3));
      } else
      if (mix != 5) { e0 = ex2_approx(e0); e1 = ex2_approx(e1); e2 = ex2_approx(e2); e3 = ex2_approx(e3); }
      if (mix >= 2 && mix <= 4) { sA = __fadd2_rn(sA, make_float2(e0, e1)); sB = __fadd2_rn(sB, make_float2(e2, e3)); }
      if (mix == 3 || mix == 5) { acc ^= pack_bf16x2(e0, e1); acc ^= pack_bf16x2(e2, e3); }
      if (mix == 4) { acc ^= __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632); acc ^= __byte_perm(__float_as_uint(e2), __float_as_uint(e3), 0x7632); }
      x[i] = e0 - 1.5f; x[i + 1] = e1 - 1.5f; x[i + 2] = e2 - 1.5f; x[i + 3] = e3 - 1.5f;  // keeps the chain data dependent, 1 FADD each
    }
  }
  const long long t1 = clock64();
  float r = sA.x + sA.y + sB.x + sB.y + __uint_as_float(acc & 0x3fffffffu);
#pragma unroll
  for (int i = 0; i < 32; ++i) r += x[i];
  if (r == 123.456f) sink[0] = r;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

// The exp sweep (full mix, register data) on 2 warps per SM sub-partition WHILE one or two threads keep the tensor core busy with
// the attention kernel's MMA mix: does tensor-core activity slow the softmax warps down?
__global__ void __launch_bounds__(384, 1) sweep_with_mma_kernel(long long* out, float* sink, int iters, int n_issuers) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(bar + 4);
  for (uint32_t i = threadIdx.x; i < OFF_BAR / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_fence_init(); *stop = 0; }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  const int w = threadIdx.x >> 5;
  if (w < 2) {
    if ((threadIdx.x & 31) == 0 && w < n_issuers) {
      const uint32_t a_addr = smem_u32(smem + OFF_A), b_addr = smem_u32(smem + OFF_B), v_addr = smem_u32(smem + OFF_V);
      constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t id64mn = make_idesc_bf16(128, 64, 0, 1);
      long long n = 0;
      while (!*stop) {
        if (w == 0 || n_issuers == 1)
          for (int k = 0; k < 4; ++k) umma_ss(tm, make_smem_desc_sw128(a_addr + k * 32, 0, 1024), make_smem_desc_sw128(b_addr + k * 32, 0, 1024), id128, 1);
        if (w == 1 || n_issuers == 1)
          for (int k = 0; k < 8; ++k) umma_ts(tm + 256, tm + 384 + k * 8, make_smem_desc_sw128(v_addr + k * 2048, 16384, 1024), id64mn, 1);
        n += 12;
      }
      umma_commit(bar + w);
      mbar_wait(bar + w, 0);
      out[1 + w] = n;
    }
  } else if (w >= 4) {
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = -0.01f * (float)((threadIdx.x + i) & 63);
    float2 sA = make_float2(0.f, 0.f), sB = make_float2(0.f, 0.f);
    uint32_t acc = 0;
    const float2 c2 = make_float2(0.999f, 0.999f), nm2 = make_float2(-0.001f, -0.001f);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float2 x01 = __ffma2_rn(make_float2(x[i], x[i + 1]), c2, nm2), x23 = __ffma2_rn(make_float2(x[i + 2], x[i + 3]), c2, nm2);
        const float e0 = ex2_approx(x01.x), e1 = ex2_approx(x01.y), e2 = ex2_approx(x23.x), e3 = ex2_approx(x23.y);
        sA = __fadd2_rn(sA, make_float2(e0, e1)); sB = __fadd2_rn(sB, make_float2(e2, e3));
        acc ^= pack_bf16x2(e0, e1); acc ^= pack_bf16x2(e2, e3);
        x[i] = e0 - 1.5f; x[i + 1] = e1 - 1.5f; x[i + 2] = e2 - 1.5f; x[i + 3] = e3 - 1.5f;
      }
    }
    const long long t1 = clock64();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float r = sA.x + sA.y + sB.x + sB.y + __uint_as_float(acc & 0x3fffffffu);
#pragma unroll
    for (int i = 0; i < 32; ++i) r += x[i];
    if (r == 123.456f) sink[0] = r;
    if (threadIdx.x == 128) { out[0] = t1 - t0; *stop = 1; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  const char* names[NCASE] = {
      "SS  N=256 K-major B, one accumulator",
      "SS  N=128 K-major B, one accumulator",
      "SS  N=128 K-major B, 2 accumulators alternating",
      "SS  N=64  K-major B, one accumulator",
      "SS  N=64  K-major B, 2 accumulators",
      "SS  N=64  K-major B, 4 accumulators",
      "SS  N=64  MN-major B (V layout), one accumulator",
      "TS  N=64  MN-major B (the kernel's P V), one accumulator",
      "TS  N=64  MN-major B, 2 accumulators",
      "TS  N=64  MN-major B, 4 accumulators",
      "TS  N=64  K-major B, one accumulator",
      "kernel order: 4 x S(N=128) then 8 x PV(TS N=64)",
      "interleaved: (PV PV S) x 4",
      "both sub-tiles back to back: S0 PV0 S1 PV1",
      "TS  N=128 MN-major B, one accumulator",
      "SS  N=64  MN-major B with A from shared memory (P in smem)",
      "SS  N=128, accumulate = 0 every time",
      "TS  N=64, accumulate = 0 every time",
      "TS  N=32  MN-major B, 2 accumulators",
      "SS  N=128, identical descriptors every time",
      "TS  N=64, identical operands every time",
      "SS  N=256, 2 accumulators alternating",
      "SS  N=128, a tcgen05.commit after every 4",
      "TS  N=64, a tcgen05.commit after every 8"};
  Result* d;
  cudaMalloc(&d, NCASE * sizeof(Result));
  cudaMemset(d, 0, NCASE * sizeof(Result));
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  // The mbarrier "bar + 1" collects the intermediate commits of cases 22 / 23 and is never waited on.
  for (int pass = 0; pass < 2; ++pass) {
    mma_rate_kernel<<<1, 128, SMEM_BYTES>>>(d, 32);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  Result h[NCASE];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("| case | MMAs | clk per MMA (issue to completion) | clk per MMA until the last one was issued | floor 128*N/256 |\n|---|---|---|---|---|\n");
  const int floors[NCASE] = {128, 64, 64, 32, 32, 32, 32, 32, 32, 32, 32, 0, 0, 0, 64, 32, 64, 32, 16, 64, 32, 128, 64, 32};
  for (int c = 0; c < NCASE; ++c)
    printf("| %s | %d | %.1f | %.1f | %s |\n", names[c], h[c].n, (double)h[c].total / h[c].n, (double)h[c].issued / h[c].n,
           floors[c] ? std::to_string(floors[c]).c_str() : "42.7 (mean of 4 x 64 + 8 x 32)");
  cudaFuncSetAttribute(mma_two_issuers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  printf("\n| issuing threads | shape | clk per MMA of one thread | MMAs per 1000 clk, all threads |\n|---|---|---|---|\n");
  for (int ts = 0; ts < 2; ++ts)
    for (int ni = 1; ni <= 2; ++ni) {
      cudaMemset(d, 0, NCASE * sizeof(Result));
      mma_two_issuers_kernel<<<1, 128, SMEM_BYTES>>>(d, 32, ni, ts);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 2 * sizeof(Result), cudaMemcpyDeviceToHost);
      double rate = 0;
      for (int w = 0; w < ni; ++w) rate += 1000.0 * h[w].n / h[w].total;
      printf("| %d | %s | %.1f | %.1f |\n", ni, ts ? "TS N=64 MN-major B" : "SS N=128", (double)h[0].total / h[0].n, rate);
    }
  {
    long long* dt; float* sink;
    cudaMalloc(&dt, 8); cudaMalloc(&sink, 4);
    const char* mixes[8] = {"MUFU.EX2 only (+ 1 FADD per element to keep the data flowing)", "FFMA2 + EX2", "FFMA2 + EX2 + FADD2 row sums",
                            "full sweep: FFMA2 + EX2 + FADD2 + F2FP bf16x2 pack", "full sweep with PRMT (truncating) pack", "F2FP pack only, no EX2",
                            "FFMA2 + cvt.f16x2 + ex2.approx.ftz.f16x2 (one MUFU per pair) + unpack + FADD2 row sums of the rounded values",
                            "FFMA2 + cvt.bf16x2 + ex2.approx.ftz.bf16x2 (one MUFU per pair) + unpack + FADD2 row sums of the rounded values"};
    printf("\n| sweep mix | warps per SM sub-partition | clk per 32 elements per warp | clk per element-row of a sub-partition (x warps) |\n|---|---|---|---|\n");
    for (int mix = 0; mix < 8; ++mix)
      for (int wps = 1; wps <= 4; wps *= 2) {
        const int iters = 64;
        switch (mix) {
          case 0: sweep_rate_kernel<0><<<1, 128 * wps>>>(dt, sink, iters); break;
          case 1: sweep_rate_kernel<1><<<1, 128 * wps>>>(dt, sink, iters); break;
          case 2: sweep_rate_kernel<2><<<1, 128 * wps>>>(dt, sink, iters); break;
          case 3: sweep_rate_kernel<3><<<1, 128 * wps>>>(dt, sink, iters); break;
          case 4: sweep_rate_kernel<4><<<1, 128 * wps>>>(dt, sink, iters); break;
          case 5: sweep_rate_kernel<5><<<1, 128 * wps>>>(dt, sink, iters); break;
          case 6: sweep_rate_kernel<6><<<1, 128 * wps>>>(dt, sink, iters); break;
          default: sweep_rate_kernel<7><<<1, 128 * wps>>>(dt, sink, iters); break;
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        long long c; cudaMemcpy(&c, dt, 8, cudaMemcpyDeviceToHost);
        printf("| %s | %d | %.0f | %.1f per MUFU-or-pack group of 32 lanes |\n", mixes[mix], wps, (double)c / iters, (double)c / iters / 32.0 / wps);
      }
  }
  {
    long long* dt; float* sink;
    cudaMalloc(&dt, 64); cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(sweep_with_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    printf("\n| MMA-issuing threads running next to the sweep | clk per 32 elements per warp (2 sweeping warps per sub-partition) | MMAs issued meanwhile |\n|---|---|---|\n");
    for (int ni = 0; ni <= 2; ++ni) {
      const int iters = 256;
      cudaMemset(dt, 0, 64);
      sweep_with_mma_kernel<<<1, 384, SMEM_BYTES>>>(dt, sink, iters, ni);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      long long c[3]; cudaMemcpy(c, dt, 24, cudaMemcpyDeviceToHost);
      printf("| %d | %.0f | %lld |\n", ni, (double)c[0] / iters, c[1] + c[2]);
    }
  }
  return 0;
}
