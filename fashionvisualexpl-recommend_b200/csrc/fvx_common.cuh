// Shared device/host helpers for libfvx (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fvx.h"

#define FVX_HD __host__ __device__ __forceinline__
#define FVX_D __device__ __forceinline__

// ---- error plumbing (host) ---------------------------------------------------------
void fvx_set_error(const char* fmt, ...);
#define FVX_FAIL(code, ...)      \
  do {                           \
    fvx_set_error(__VA_ARGS__);  \
    return (code);               \
  } while (0)
#define FVX_CHECK_ARG(cond, ...) \
  do {                           \
    if (!(cond)) FVX_FAIL(-2, __VA_ARGS__); \
  } while (0)
#define FVX_CHECK_LAUNCH(name)                                                       \
  do {                                                                               \
    cudaError_t e__ = cudaPeekAtLastError();                                         \
    if (e__ != cudaSuccess) {                                                        \
      cudaGetLastError();                                                            \
      FVX_FAIL(-3, "%s: launch failed: %s", (name), cudaGetErrorString(e__));        \
    }                                                                                \
  } while (0)

static inline cudaStream_t fvx_cu(fvx_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
int fvx_num_sms();      // SM count of the CURRENT device (cached per device)
int fvx_cur_device();   // current CUDA device ordinal (0 on error)
// Per-device state of a kernel's opt-in dynamic shared memory limit: a process may drive several GPUs
// (one Engine per device), and cudaFuncSetAttribute applies to the current device's context only.
#define FVX_MAX_DEV 64
struct FvxSmemMark { size_t v[FVX_MAX_DEV]; };
// Raises the limit of `func` on the current device when `smem` exceeds what was configured there before.
int fvx_ensure_smem(const void* func, FvxSmemMark* mark, size_t smem, const char* who);

// ---- Adam constants (Keras defaults; BPRMF.py:52, VBPR.py:56) --------------------
#define FVX_BETA1 0.9f
#define FVX_BETA2 0.999f
#define FVX_EPS 1e-7f
#define FVX_CLIP_LO (-80.0f)
#define FVX_CLIP_HI (1e8f)
#define FVX_REPLAY_MAX 192  // deferred replay: iterations after which beta1^k*m is < 2e-9*m

// alpha_t = lr*sqrt(1-b2^t)/(1-b1^t), 1-based t, evaluated in double once per thread
FVX_HD float fvx_alpha(float lr, long long t) {
  double b1t = pow(0.9, (double)t), b2t = pow(0.999, (double)t);
  return (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
}

// ---- Philox4x32-10 ---------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};
FVX_HD void fvx_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  unsigned long long p = (unsigned long long)a * b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}
FVX_HD Philox4 fvx_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    fvx_mulhilo(0xD2511F53u, c0, hi0, lo0);
    fvx_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
#define FVX_STREAM_NEG 0u
#define FVX_STREAM_PERM 1u
#define FVX_MAX_ATTEMPTS 256

// sorted-row membership test: is x in col[lo, hi) (ascending)?
FVX_HD bool fvx_in_sorted(const int32_t* __restrict__ col, long long lo, long long hi, int32_t x) {
  const long long end = hi;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (col[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo < end && col[lo] == x;
}

// one uniform negative for global triple index g (oracle/sampler.py: philox_negatives)
FVX_HD int32_t fvx_draw_negative(const int32_t* __restrict__ col_sorted, long long lo, long long hi,
                                 unsigned long long g, uint32_t num_items, unsigned long long seed) {
  uint32_t cand = 0;
  for (uint32_t a = 0; a < FVX_MAX_ATTEMPTS; a += 4) {
    Philox4 r = fvx_philox((uint32_t)g, (uint32_t)(g >> 32), a >> 2, FVX_STREAM_NEG, (uint32_t)seed,
                           (uint32_t)(seed >> 32));
    uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cand = (uint32_t)(((unsigned long long)w[q] * num_items) >> 32);
      if (!fvx_in_sorted(col_sorted, lo, hi, (int32_t)cand)) return (int32_t)cand;
    }
  }
  return (int32_t)cand;  // bound reached: keep the last candidate (never happens on real data)
}

// ---- warp helpers ------------------------------------------------------------------
#ifdef __CUDACC__
FVX_D float fvx_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
FVX_D float fvx_sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
FVX_D float fvx_rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
FVX_D void fvx_red_add(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
#endif
