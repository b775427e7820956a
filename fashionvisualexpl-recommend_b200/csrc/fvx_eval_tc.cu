// Full-catalog top-k on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// predict_all (BPRMF.py:85, VBPR.py:95-97) is the contraction
//     S[u,i] = <[Gu|Tu|1|1][u,:], [Gi|theta|b_hi|b_lo][i,:]>,   b = Bi + vbias
// and Evaluator.store_recommendation (Evaluator.py:231-237) wants, per user, the k best
// non-train items.  The sweep below never writes S:
//
//   1. k_pack_users / k_pack_items   bf16 operands (K+d+3 padded to KP); the item bias rides in
//                                    two columns (bf16 hi + lo) so that it enters the score
//                                    almost exactly, the rounding bound in a third
//   2. k_topk_tc (persistent, warp-specialised, one CTA per SM)
//        warp 0    TMA producer : item tiles [128 x KP] -> 64B-swizzled smem ring; the two
//                                 user tiles [128 x KP] of the work unit
//        warp 1    UMMA issuer  : D[ut][128 x 128] (TMEM, fp32) = A_ut * B_tile^T for both user
//                                 tiles of the unit (the item tile is read from smem once for
//                                 256 users), 4 accumulators = 2 user tiles x double buffer
//        warps 2-9 epilogue     : one thread per user row (single owner): tcgen05.ld 32
//                                 columns at a time (next chunk in flight while this one is
//                                 tested), a 3-input max tree against the row's running
//                                 threshold; survivors are appended to the row's candidate
//                                 list; a warp-cooperative radix select tightens the threshold
//                                 when the list grows past k + slack; thresholds are shared
//                                 between the item splits of a row through global memory
//   3. k_rescore_select              exact fp32 re-scoring of the surviving candidates with the
//                                    same fvx_score_one() the fp32 path uses, train-item mask,
//                                    sort, top-k.
//
// Exactness.  With a = [Gu|Tu][u], b = [Gi|theta][i] rounded to bf16 (relative error 2^-9 each)
//   |s_bf16 - s_fp32| <= c * |a| * |b_i| + beta0,   c = 1.003 * 2^-8 + KP * 2^-21  (rounding + fp32
//   accumulation),  beta0 = (2^-17 + KP * 2^-21) * max_i |bias_i|                (hi+lo residual).
// The bound is PER ITEM and costs nothing: one more K column holds eps_u = c*|a_u| (rounded up to
// bf16) on the user side and |b_i| (rounded up) on the item side, so the UMMA itself delivers the
// upper bound s_ub = s_bf16 + eps_u * |b_i|; the lower bound is s_lb = s_ub - 2.001 * eps_u * |b_i|
// - beta0.  A row keeps every item with s_ub >= tau - beta0, tau = the (k + #train items)-th best
// s_lb seen so far (a lower bound of the k-th best true score over the non-train items), so the
// true top-k is always among the candidates and the output equals the fp32 kernel's bit for bit.
// A row whose list overflows is flagged and the caller re-runs it through the fp32 kernel.
#include <cuda_bf16.h>
#include <math_constants.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"
#include "fvx_tc.cuh"

#define TCK_BM 128          // users per tile  (UMMA M)
#define TCK_BN 128          // items per tile  (UMMA N)
#define TCK_KB 32           // bf16 elements per K block (64-byte swizzle rows)
#define TCK_CAP 512         // candidate slots per (user, split)
#define TCK_SLACK 192       // a list is compacted once it holds k + #train + TCK_SLACK entries
#define TCK_THREADS 320     // warp 0 producer, warp 1 UMMA, warps 2-9 epilogue
#define TCK_RS_MAX 2048     // candidates per user the final selection can take
#define KEY_PAD 0xFFFFFFFFFFFFFFFFull

// order-preserving float <-> uint (ascending uint == ascending float)
__device__ __forceinline__ uint32_t tck_mono(float s) {
  const uint32_t b = __float_as_uint(s);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float tck_unmono(uint32_t m) {
  return __uint_as_float((m & 0x80000000u) ? (m ^ 0x80000000u) : ~m);
}
// list key: ascending key == descending score, then ascending item id
__device__ __forceinline__ unsigned long long tck_key(float s, int32_t id) {
  return ((unsigned long long)(~tck_mono(s)) << 32) | (uint32_t)id;
}
__device__ __forceinline__ float tck_score_of_hi(uint32_t hi) { return tck_unmono(~hi); }
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// ---------------------------------------------------------------------------------
__device__ __forceinline__ __nv_bfloat16 bf16_up(float x) {   // smallest bf16 >= x (x >= 0)
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  if (__bfloat162float(h) < x) h = __ushort_as_bfloat16((unsigned short)(__bfloat16_as_ushort(h) + 1));
  return h;
}

__global__ void k_pack_users(FvxModel M, int u0, int u1, __nv_bfloat16* __restrict__ A, float* __restrict__ epsa,
                             uint32_t* __restrict__ thr_g, int KP, float c_rel) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int kd = M.K + M.d;
  for (int u = u0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); u < u1; u += warps) {
    const float* src = M.users.w + (size_t)u * M.users.stride;
    float sq = 0.0f;
    for (int c = lane; c < kd; c += 32) { const float v = src[c]; sq += v * v; }
    sq = fvx_warp_sum(sq);
    const __nv_bfloat16 eh = bf16_up(c_rel * sqrtf(sq) * 1.0001f);
    for (int c = lane; c < KP; c += 32) {
      __nv_bfloat16 o = __float2bfloat16_rn(0.0f);
      if (c < kd) o = __float2bfloat16_rn(src[c]);
      else if (c == kd || c == kd + 1) o = __float2bfloat16_rn(1.0f);   // multiplies bias_hi and bias_lo
      else if (c == kd + 2) o = eh;                                     // multiplies |b_i|
      A[(size_t)(u - u0) * KP + c] = o;
    }
    if (lane == 0) {
      epsa[u - u0] = __bfloat162float(eh);
      thr_g[u - u0] = tck_mono(-CUDART_INF_F);
    }
  }
}

// nb[i] = |[Gi|theta][i]| rounded up to bf16; stat[1] = max |bias| (uint bits of a non-negative float)
__global__ void k_pack_items(FvxModel M, const float* __restrict__ theta, __nv_bfloat16* __restrict__ Bm,
                             float* __restrict__ nb, uint32_t* __restrict__ stat, int KP) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int K = M.K, d = M.d, kd = K + d;
  float wmax = 0.0f, bmax = 0.0f;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < M.item_cnt; i += warps) {
    const float* row = M.items.w + (size_t)i * M.items.stride;
    const float* th = d > 0 ? theta + (size_t)i * M.de : nullptr;
    const float bias = row[K] + (d > 0 ? th[d] : 0.0f);
    const __nv_bfloat16 bh = __float2bfloat16_rn(bias);
    const __nv_bfloat16 bl = __float2bfloat16_rn(bias - __bfloat162float(bh));
    float sq = 0.0f;
    for (int c = lane; c < kd; c += 32) { const float v = c < K ? row[c] : th[c - K]; sq += v * v; }
    sq = fvx_warp_sum(sq);
    const __nv_bfloat16 nh = bf16_up(sqrtf(sq) * 1.0001f);
    for (int c = lane; c < KP; c += 32) {
      __nv_bfloat16 o = __float2bfloat16_rn(0.0f);
      if (c < kd) o = __float2bfloat16_rn(c < K ? row[c] : th[c - K]);
      else if (c == kd) o = bh;
      else if (c == kd + 1) o = bl;
      else if (c == kd + 2) o = nh;
      Bm[(size_t)i * KP + c] = o;
    }
    if (lane == 0) nb[i] = __bfloat162float(nh);
    wmax = fmaxf(wmax, sqrtf(sq));
    bmax = fmaxf(bmax, fabsf(bias));
  }
  if (lane == 0) {   // non-negative floats order like uints
    atomicMax(stat, __float_as_uint(wmax));
    atomicMax(stat + 1, __float_as_uint(bmax));
  }
}

// ---------------------------------------------------------------------------------
struct TckParams {
  int n_users;          // users in this call (rows of A)
  int item_cnt, item_lo;
  int nkb;              // K blocks of 32
  int stages;
  // work units: the first n_full user-tile pairs (a multiple of the grid size) sweep the whole
  // catalog each (one list per row, no threshold restart); the remaining pairs are cut into
  // `splits` item ranges so that the last wave still fills the machine
  int splits, tiles_per_split, n_item_tiles, n_pairs, n_full;
  int k;
  int u0;
  float beta_c;         // 2^-17 + KP * 2^-21: bias residual + fp32 accumulation, per unit of max|bias|
  const float* epsa;    // [n_users] eps_u as multiplied by the UMMA (bf16 value)
  const float* nb;      // [item_cnt] |b_i| as multiplied by the UMMA (bf16 value)
  const uint32_t* stat; // [0] max item norm, [1] max |bias|  (float bits)
  const int64_t* mask_row_ptr;
  unsigned long long* cand;   // [lists * CAP]
  int32_t* ccount;            // [lists]
  int32_t* flags;             // [n_users]
  uint32_t* thr_g;            // [n_users] best known row threshold on s_ub (tck_mono encoding)
};

// list index of (row, split): rows of the full pairs own one list, tail rows `splits` lists
__device__ __forceinline__ size_t tck_list(const TckParams& P, int row, int sp) {
  const int full_rows = P.n_full * 2 * TCK_BM;
  return row < full_rows ? (size_t)row : (size_t)full_rows + (size_t)(row - full_rows) * P.splits + sp;
}
// unit w of the global enumeration -> (pair, split, tile range)
__device__ __forceinline__ void tck_unit(const TckParams& P, int w, int& pair, int& sp, int& t0, int& t1) {
  if (w < P.n_full) { pair = w; sp = 0; t0 = 0; t1 = P.n_item_tiles; return; }
  const int x = w - P.n_full;
  pair = P.n_full + x / P.splits;
  sp = x - (x / P.splits) * P.splits;
  t0 = sp * P.tiles_per_split;
  t1 = min(P.n_item_tiles, t0 + P.tiles_per_split);
}

// warp-cooperative: tighten the threshold of lane `L`'s row and prune its candidate list
__device__ __forceinline__ void tck_compact_row(const TckParams& P, int L, int lane, int my_row, int split,
                                                float my_eps2, float beta0, int my_kk, int& cnt, float& thr) {
  const int row = __shfl_sync(0xffffffffu, my_row, L);
  const int n = __shfl_sync(0xffffffffu, cnt, L);
  const int kk = __shfl_sync(0xffffffffu, my_kk, L);
  const float eps2 = __shfl_sync(0xffffffffu, my_eps2, L);     // 2.001 * eps_u
  const float old_thr = __shfl_sync(0xffffffffu, thr, L);
  unsigned long long* buf = P.cand + tck_list(P, row, split) * TCK_CAP;
  // The train-item mask is NOT consulted here: the (k + #train items)-th best score over ALL
  // items is a lower bound of the k-th best over the non-train items.  k_rescore_select applies
  // the mask exactly.
  uint32_t e[TCK_CAP / 32];     // ~mono(s_ub): the list key (ascending = best first)
  uint32_t lbk[TCK_CAP / 32];   // ~mono(s_lb)
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
  for (int q = 0; q < TCK_CAP / 32; ++q) {
    const int idx = q * 32 + lane;
    e[q] = 0xFFFFFFFFu;
    lbk[q] = 0xFFFFFFFFu;
    if (idx < n) {
      const unsigned long long key = buf[idx];
      e[q] = (uint32_t)(key >> 32);
      const float lb = tck_score_of_hi(e[q]) - eps2 * P.nb[(uint32_t)key - (uint32_t)P.item_lo] - beta0;
      lbk[q] = ~tck_mono(lb);
      kmin = min(kmin, lbk[q]);
      kmax = max(kmax, lbk[q]);
    }
  }
  float new_thr = fmaxf(old_thr, tck_unmono(__ldcg(P.thr_g + row)));
  if (n >= kk) {
    // a 32-bit key T with  kk <= #(lb key <= T) <= kk + 16  (or exactly the kk-th best lower bound)
    uint32_t lo = __reduce_min_sync(0xffffffffu, kmin), hi = __reduce_max_sync(0xffffffffu, kmax);
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int c = 0;
#pragma unroll
      for (int q = 0; q < TCK_CAP / 32; ++q) c += (lbk[q] <= mid) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= kk) { hi = mid; if (c <= kk + 16) break; } else lo = mid + 1;
    }
    new_thr = fmaxf(new_thr, tck_score_of_hi(hi) - beta0);
  }
  const uint32_t cut = ~tck_mono(new_thr);          // keep keys <= cut  <=>  s_ub >= new_thr
  int out = 0;
#pragma unroll
  for (int q = 0; q < TCK_CAP / 32; ++q) {
    const int idx = q * 32 + lane;
    const bool keep = idx < n && e[q] <= cut;
    const unsigned long long full = keep ? buf[idx] : 0ull;
    __syncwarp();
    const uint32_t b = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[out + __popc(b & ((1u << lane) - 1u))] = full;
    out += __popc(b);
  }
  __syncwarp();
  if (out > TCK_CAP - 40) {            // too many items inside the margin: row is re-run in fp32
    if (lane == 0) P.flags[row] = 1;
    out = TCK_CAP - 40;
  }
  if (lane == L) {
    cnt = out;
    thr = new_thr;
    atomicMax(P.thr_g + row, tck_mono(new_thr));
  }
}

// one 32-column chunk of one row: group maxima first, then only the groups that can hold a survivor
__device__ __forceinline__ void tck_scan_chunk(const uint32_t (&v)[32], float thr, int ibase, int item_cnt,
                                               int item_lo, unsigned long long* buf, int& cnt) {
  float g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float a0 = max3(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1]), __uint_as_float(v[8 * q + 2]));
    const float a1 = max3(__uint_as_float(v[8 * q + 3]), __uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
    g[q] = max3(a0, a1, fmaxf(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
  }
  const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
  if (m >= thr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (g[q] >= thr) {
#pragma unroll
        for (int j = 8 * q; j < 8 * q + 8; ++j) {
          const float s = __uint_as_float(v[j]);
          if (s >= thr && ibase + j < item_cnt && cnt < TCK_CAP) {
            buf[cnt] = tck_key(s, item_lo + ibase + j);
            ++cnt;
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(TCK_THREADS, 1)
k_topk_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TckParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // swizzled TMA / UMMA tiles need their base aligned to the swizzle repeat: round up by hand
  uint8_t* smem = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = (uint32_t)P.nkb * TCK_BM * 64u;      // one user tile
  const uint32_t b_bytes = (uint32_t)P.nkb * TCK_BN * 64u;
  uint8_t* sA = smem;                                            // [2][a_bytes]
  uint8_t* sB = smem + 2 * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)P.stages * b_bytes);
  uint64_t* full_b = bars;                  // [stages]
  uint64_t* empty_b = bars + P.stages;      // [stages]
  uint64_t* a_full = bars + 2 * P.stages;   // [1]
  uint64_t* a_empty = a_full + 1;           // [1]
  uint64_t* t_full = a_empty + 1;           // [4]  accumulator = ut * 2 + buffer
  uint64_t* t_empty = t_full + 4;           // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 4);

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_units = P.n_full + (P.n_pairs - P.n_full) * P.splits;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, unit_i = 0;
      for (int w = blockIdx.x; w < n_units; w += gridDim.x, ++unit_i) {
        int pair, sp, t0, t1;
        tck_unit(P, w, pair, sp, t0, t1);
        mbar_wait(a_empty, (unit_i & 1) ^ 1);
        mbar_expect_tx(a_full, 2 * a_bytes);
        for (int ut = 0; ut < 2; ++ut)
          for (int kb = 0; kb < P.nkb; ++kb)
            tma_load_2d(sA + (size_t)ut * a_bytes + (size_t)kb * TCK_BM * 64, &tmA, a_full, kb * TCK_KB,
                        (pair * 2 + ut) * TCK_BM);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&empty_b[stage], phase ^ 1);
          mbar_expect_tx(&full_b[stage], b_bytes);
          uint8_t* dst = sB + (size_t)stage * b_bytes;
          for (int kb = 0; kb < P.nkb; ++kb)
            tma_load_2d(dst + (size_t)kb * TCK_BN * 64, &tmB, &full_b[stage], kb * TCK_KB, t * TCK_BN);
          if (++stage == (uint32_t)P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== UMMA issuer (one elected thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TCK_BM, TCK_BN, 0, 0);
      uint32_t stage = 0, phase = 0, unit_i = 0, buf = 0, buf_phase = 0;
      for (int w = blockIdx.x; w < n_units; w += gridDim.x, ++unit_i) {
        int pair, sp, t0, t1;
        tck_unit(P, w, pair, sp, t0, t1);
        mbar_wait(a_full, unit_i & 1);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&full_b[stage], phase);
          tc_fence_after();
          const uint32_t b0 = tc_smem_u32(sB + (size_t)stage * b_bytes);
          for (int ut = 0; ut < 2; ++ut) {
            const uint32_t acc = ut * 2 + buf;
            mbar_wait(&t_empty[acc], buf_phase ^ 1);
            tc_fence_after();
            const uint32_t a0 = tc_smem_u32(sA + (size_t)ut * a_bytes);
            const uint32_t d = tmem_base + acc * TCK_BN;
            for (int kb = 0; kb < P.nkb; ++kb) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t ad = umma_smem_desc(a0 + kb * TCK_BM * 64 + ks * 32, 16, 512, TC_SWZ_64B);
                const uint64_t bd = umma_smem_desc(b0 + kb * TCK_BN * 64 + ks * 32, 16, 512, TC_SWZ_64B);
                umma_f16(d, ad, bd, idesc, (kb | ks) ? 1u : 0u);
              }
            }
            umma_commit(&t_full[acc]);     // accumulator ready for its epilogue warps
          }
          umma_commit(&empty_b[stage]);    // smem stage may be refilled once these UMMAs retire
          if (++stage == (uint32_t)P.stages) { stage = 0; phase ^= 1; }
          if (++buf == 2) { buf = 0; buf_phase ^= 1; }
        }
        umma_commit(a_empty);              // user tiles may be overwritten
      }
    }
  } else {
    // ===== epilogue: warps 2-5 own user tile 0, warps 6-9 user tile 1; thread = user row =====
    const int ew = warp - 2;
    const int ut = ew >> 2;
    const int quad = warp & 3;            // TMEM lanes this warp may read: [32*quad, 32*quad+32)
    const int rit = quad * 32 + lane;     // row in tile
    const float beta0 = P.beta_c * __uint_as_float(P.stat[1]) + 1e-30f;
    uint32_t buf = 0, buf_phase = 0;
    for (int w = blockIdx.x; w < n_units; w += gridDim.x) {
      int pair, sp, t0, t1;
      tck_unit(P, w, pair, sp, t0, t1);
      const bool shared = w >= P.n_full;    // the row's other splits run elsewhere: exchange bounds
      const int row = (pair * 2 + ut) * TCK_BM + rit;
      const bool live = row < P.n_users;
      float eps2 = 0.0f;
      int kk = P.k;
      if (live) {
        eps2 = 2.001f * P.epsa[row];
        const int gu = P.u0 + row;
        kk = P.k + (int)(P.mask_row_ptr[gu + 1] - P.mask_row_ptr[gu]);
        if (kk > TCK_CAP - TCK_SLACK - 40) {   // cannot bound this row's list: exact fp32 sweep instead
          kk = TCK_CAP - TCK_SLACK - 40;
          P.flags[row] = 1;
        }
      }
      const int trig = kk + TCK_SLACK;
      float thr = live ? -CUDART_INF_F : CUDART_INF_F;
      int cnt = 0;
      unsigned long long* lbuf = P.cand + tck_list(P, live ? row : 0, sp) * TCK_CAP;
      for (int t = t0; t < t1; ++t) {
        const uint32_t acc = ut * 2 + buf;
        // bounds found by the row's other splits (L2), fetched now and applied after this tile
        uint32_t peer = 0u;
        const bool refresh = shared && live && ((t - t0) & 7) == 0;
        if (refresh) peer = __ldcg(P.thr_g + row);
        mbar_wait(&t_full[acc], buf_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TCK_BN;
        const int item0 = t * TCK_BN;
        uint32_t va[32], vb[32];
        tmem_ld_32x32(taddr, va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          // next chunk in flight while this one is tested
          if (c == 0) tmem_ld_32x32(taddr + 32, vb);
          if (c == 1) tmem_ld_32x32(taddr + 64, va);
          if (c == 2) tmem_ld_32x32(taddr + 96, vb);
          if ((c & 1) == 0) tck_scan_chunk(va, thr, item0 + c * 32, P.item_cnt, P.item_lo, lbuf, cnt);
          else tck_scan_chunk(vb, thr, item0 + c * 32, P.item_cnt, P.item_lo, lbuf, cnt);
          uint32_t need = __ballot_sync(0xffffffffu, cnt >= trig);
          while (need) {
            const int L = __ffs(need) - 1;
            need &= need - 1;
            tck_compact_row(P, L, lane, row, sp, eps2, beta0, kk, cnt, thr);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acc]);
        if (++buf == 2) { buf = 0; buf_phase ^= 1; }
        if (refresh) thr = fmaxf(thr, tck_unmono(peer));
      }
      // publish this unit's bound for the rows that hold at least kk candidates, then the length
      uint32_t need = __ballot_sync(0xffffffffu, live && cnt >= kk);
      while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        tck_compact_row(P, L, lane, row, sp, eps2, beta0, kk, cnt, thr);
      }
      if (live) P.ccount[tck_list(P, row, sp)] = cnt;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------
// exact re-scoring + final selection: one warp per user
#define RS_WARPS 2
__device__ __forceinline__ void rs_bitonic(unsigned long long* keys, int n, int lane) {
  for (int size = 2; size <= n; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = lane; i < (n >> 1); i += 32) {
        const int l = 2 * i - (i & (stride - 1)), r = l + stride;
        const bool up = (l & size) == 0;
        const unsigned long long a = keys[l], b = keys[r];
        if ((a > b) == up) { keys[l] = b; keys[r] = a; }
      }
      __syncwarp();
    }
}

// x_ui with the fp32 kernel's arithmetic (fvx_score_one: K latent terms, d visual terms, item
// bias, visual bias, one fmaf chain in index order) but 16-byte loads of the item row
__device__ __forceinline__ float rs_score(const float* __restrict__ urow, const float* __restrict__ irow,
                                          const float* __restrict__ th, int K, int d) {
  float s = 0.0f;
  int c = 0;
  for (; c + 4 <= K; c += 4) {
    const float4 x = *reinterpret_cast<const float4*>(irow + c);
    s = fmaf(urow[c], x.x, s);
    s = fmaf(urow[c + 1], x.y, s);
    s = fmaf(urow[c + 2], x.z, s);
    s = fmaf(urow[c + 3], x.w, s);
  }
  for (; c < K; ++c) s = fmaf(urow[c], irow[c], s);
  if (d > 0) {
    int n = 0;
    for (; n + 4 <= d; n += 4) {
      const float4 x = *reinterpret_cast<const float4*>(th + n);
      s = fmaf(urow[K + n], x.x, s);
      s = fmaf(urow[K + n + 1], x.y, s);
      s = fmaf(urow[K + n + 2], x.z, s);
      s = fmaf(urow[K + n + 3], x.w, s);
    }
    for (; n < d; ++n) s = fmaf(urow[K + n], th[n], s);
    s += irow[K];
    s += th[d];
  } else {
    s += irow[K];
  }
  return s;
}

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rescore_select(FvxModel M, const float* __restrict__ theta, TckParams P,
                 const int64_t* __restrict__ mask_row_ptr, const int32_t* __restrict__ mask_col, int k,
                 int32_t* __restrict__ out_ids, float* __restrict__ out_scores) {
  __shared__ unsigned long long rs_keys[RS_WARPS][TCK_RS_MAX];
  __shared__ float rs_user[RS_WARPS][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* kk = rs_keys[warp];
  float* us = rs_user[warp];
  const int full_rows = P.n_full * 2 * TCK_BM;
  const int kd = M.K + M.d;
  for (int r = blockIdx.x * RS_WARPS + warp; r < P.n_users; r += gridDim.x * RS_WARPS) {
    const int gu = P.u0 + r;
    const float* urow = M.users.w + (size_t)gu * M.users.stride;
    for (int c = lane; c < kd; c += 32) us[c] = urow[c];
    const long long mlo = mask_row_ptr[gu], mhi = mask_row_ptr[gu + 1];
    // survivors of the row's final bound on s_ub
    const uint32_t cut = ~P.thr_g[r];
    const int nlists = r < full_rows ? 1 : P.splits;
    int total = 0;
    for (int sp = 0; sp < nlists; ++sp) {
      const size_t li = tck_list(P, r, sp);
      const int n = P.ccount[li];
      const unsigned long long* src = P.cand + li * TCK_CAP;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const unsigned long long key = i < n ? src[i] : KEY_PAD;
        const bool keep = i < n && (uint32_t)(key >> 32) <= cut;
        const uint32_t b = __ballot_sync(0xffffffffu, keep);
        const int pos = total + __popc(b & ((1u << lane) - 1u));
        if (keep && pos < TCK_RS_MAX) kk[pos] = key;
        total += __popc(b);
      }
    }
    if (total > TCK_RS_MAX) {
      if (lane == 0) P.flags[r] = 1;
      total = TCK_RS_MAX;
    }
    __syncwarp();
    // exact fp32 scores (the fp32 kernel's arithmetic) and the train mask
    for (int i = lane; i < total; i += 32) {
      const int32_t gid = (int32_t)(kk[i] & 0xFFFFFFFFu);
      unsigned long long key = KEY_PAD;
      if (!fvx_in_sorted(mask_col, mlo, mhi, gid)) {
        const int32_t li = gid - M.item_lo;
        const float* th = M.d > 0 ? theta + (size_t)li * M.de : nullptr;
        key = tck_key(rs_score(us, M.items.w + (size_t)li * M.items.stride, th, M.K, M.d), gid);
      }
      kk[i] = key;
    }
    int npad = 32;
    while (npad < total) npad <<= 1;
    for (int i = total + lane; i < npad; i += 32) kk[i] = KEY_PAD;
    __syncwarp();
    rs_bitonic(kk, npad, lane);
    for (int i = lane; i < k; i += 32) {
      const bool ok = i < npad && kk[i] != KEY_PAD;
      out_ids[(size_t)r * k + i] = ok ? (int32_t)(kk[i] & 0xFFFFFFFFu) : -1;
      out_scores[(size_t)r * k + i] = ok ? tck_score_of_hi((uint32_t)(kk[i] >> 32)) : -CUDART_INF_F;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------
tc_encode_tiled_fn tc_get_encode_tiled() {
  static tc_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tc_encode_tiled_fn>(p);
  }
  return fn;
}

int tc_make_tensor_map_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                            uint32_t box_cols, uint32_t box_rows, int swizzle) {
  tc_encode_tiled_fn enc = tc_get_encode_tiled();
  if (!enc) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

extern "C" {

// geometry shared by the query and the launch
static void tck_geometry(const FvxModel* m, int n_users, int* KP, int* n_pairs, int* n_full, int* splits, int* grid,
                         long long* lists) {
  const int kd3 = m->K + m->d + 3;
  *KP = (kd3 + TCK_KB - 1) / TCK_KB * TCK_KB;
  *n_pairs = (n_users + 2 * TCK_BM - 1) / (2 * TCK_BM);
  const int n_item_tiles = (m->item_cnt + TCK_BN - 1) / TCK_BN;
  const int G = fvx_num_sms();
  *n_full = (*n_pairs / G) * G;
  const int tail = *n_pairs - *n_full;
  int s = 1;
  if (tail > 0) {
    // the tail pairs are cut into s item ranges so that their units still fill the machine: the
    // tail then takes ceil(tail*s/G) rounds of 1/s of a sweep; every extra split costs a threshold
    // restart (~3 % of a sweep here), so the smallest s within 2 % of the best is taken
    int smax = n_item_tiles / 8 < 16 ? n_item_tiles / 8 : 16;
    if (smax < 1) smax = 1;
    double best = 1e30;
    for (int c = 1; c <= smax; ++c) {
      const double t = (double)((tail * c + G - 1) / G) / c + 0.03 * c;
      if (t < best * 0.98) { best = t; s = c; }
    }
  }
  *splits = s;
  const long long units = (long long)*n_full + (long long)tail * s;
  *grid = (int)(units < G ? units : G);
  *lists = (long long)*n_full * 2 * TCK_BM + ((long long)n_users - (long long)*n_full * 2 * TCK_BM > 0
                                               ? ((long long)n_users - (long long)*n_full * 2 * TCK_BM) * s : 0);
}

int fvx_eval_ws_query(const FvxModel* model, int32_t n_users, FvxEvalWs* ws) {
  FVX_CHECK_ARG(model && ws && n_users > 0, "fvx_eval_ws_query: bad arguments");
  int KP, n_pairs, n_full, splits, grid;
  long long lists;
  tck_geometry(model, n_users, &KP, &n_pairs, &n_full, &splits, &grid, &lists);
  ws->KP = KP;
  ws->splits = splits;
  ws->cap = TCK_CAP;
  ws->u_cap = n_users;
  ws->i_cap = model->item_cnt;
  ws->lists = lists;
  return 0;
}

int fvx_score_topk_tc(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                      const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k, int32_t* out_ids,
                      float* out_scores, const FvxEvalWs* ws, fvx_stream_t stream) {
  FVX_CHECK_ARG(model && model->abi_version == FVX_ABI_VERSION && ws, "fvx_score_topk_tc: bad model / workspace");
  FVX_CHECK_ARG(model->d == 0 || theta_ext, "fvx_score_topk_tc: VBPR scoring needs theta_ext");
  FVX_CHECK_ARG(0 <= u0 && u0 < u1 && u1 <= model->num_users, "fvx_score_topk_tc: bad user range");
  FVX_CHECK_ARG(k >= 1 && k <= 128, "fvx_score_topk_tc: k=%d outside [1,128]", k);
  const int n_users = u1 - u0;
  int KP, n_pairs, n_full, splits, grid;
  long long lists;
  tck_geometry(model, n_users, &KP, &n_pairs, &n_full, &splits, &grid, &lists);
  FVX_CHECK_ARG(ws->KP == KP && ws->splits == splits && ws->cap == TCK_CAP && ws->u_cap >= n_users &&
                ws->i_cap >= model->item_cnt && ws->lists >= lists,
                "fvx_score_topk_tc: workspace does not match fvx_eval_ws_query");
  FVX_CHECK_ARG(ws->A && ws->Bm && ws->epsa && ws->nb && ws->stat && ws->cand && ws->ccount && ws->flags && ws->thr,
                "fvx_score_topk_tc: null workspace buffer");
  const int nkb = KP / TCK_KB;
  FVX_CHECK_ARG(KP <= 128 && model->K + model->d <= 128,
                "fvx_score_topk_tc: K+d+3=%d too large for the tensor-core sweep (use fvx_score_topk)",
                model->K + model->d + 3);
  cudaStream_t st = fvx_cu(stream);

  cudaMemsetAsync(ws->stat, 0, 8, st);
  cudaMemsetAsync(ws->flags, 0, sizeof(int32_t) * n_users, st);
  cudaMemsetAsync(ws->ccount, 0, sizeof(int32_t) * lists, st);
  const float c_rel = 1.003f * 0.00390625f + (float)KP * 4.76837158e-7f;
  int g = (n_users * 32 + 255) / 256;
  if (g > fvx_num_sms() * 8) g = fvx_num_sms() * 8;
  k_pack_users<<<g, 256, 0, st>>>(*model, u0, u1, reinterpret_cast<__nv_bfloat16*>(ws->A), ws->epsa,
                                  reinterpret_cast<uint32_t*>(ws->thr), KP, c_rel);
  g = fvx_num_sms() * 8;
  k_pack_items<<<g, 256, 0, st>>>(*model, theta_ext, reinterpret_cast<__nv_bfloat16*>(ws->Bm), ws->nb,
                                  reinterpret_cast<uint32_t*>(ws->stat), KP);
  FVX_CHECK_LAUNCH("k_pack");

  CUtensorMap tmA, tmB;
  int rc = tc_make_tensor_map_bf16(&tmA, ws->A, n_users, KP, (uint64_t)KP * 2, TCK_KB, TCK_BM, 2);
  if (rc == 0) rc = tc_make_tensor_map_bf16(&tmB, ws->Bm, model->item_cnt, KP, (uint64_t)KP * 2, TCK_KB, TCK_BN, 2);
  if (rc != 0) FVX_FAIL(-4, "fvx_score_topk_tc: cuTensorMapEncodeTiled failed (%d)", rc);

  TckParams P;
  P.n_users = n_users; P.item_cnt = model->item_cnt; P.item_lo = model->item_lo; P.nkb = nkb;
  P.n_pairs = n_pairs; P.n_full = n_full;
  P.n_item_tiles = (model->item_cnt + TCK_BN - 1) / TCK_BN;
  P.splits = splits;
  P.tiles_per_split = (P.n_item_tiles + P.splits - 1) / P.splits;
  P.k = k; P.u0 = u0; P.epsa = ws->epsa; P.nb = ws->nb; P.stat = reinterpret_cast<const uint32_t*>(ws->stat);
  P.beta_c = 7.6294e-6f + (float)KP * 4.76837158e-7f;
  P.mask_row_ptr = mask_row_ptr;
  P.cand = reinterpret_cast<unsigned long long*>(ws->cand); P.ccount = ws->ccount; P.flags = ws->flags;
  P.thr_g = reinterpret_cast<uint32_t*>(ws->thr);
  const size_t a_bytes = (size_t)nkb * TCK_BM * 64, b_bytes = (size_t)nkb * TCK_BN * 64;
  int stages = (int)((200 * 1024 - 2 * a_bytes) / b_bytes);
  if (stages > 6) stages = 6;
  FVX_CHECK_ARG(stages >= 2, "fvx_score_topk_tc: tile does not fit shared memory");
  P.stages = stages;
  const size_t smem = 2 * a_bytes + stages * b_bytes + (2 * stages + 10) * 8 + 16 + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_topk_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) FVX_FAIL(-3, "fvx_score_topk_tc: cannot set %zu B smem: %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  k_topk_tc<<<grid, TCK_THREADS, smem, st>>>(tmA, tmB, P);
  FVX_CHECK_LAUNCH("k_topk_tc");

  long long rg = ((long long)n_users + RS_WARPS - 1) / RS_WARPS;
  if (rg > (long long)fvx_num_sms() * 8) rg = (long long)fvx_num_sms() * 8;
  k_rescore_select<<<(int)rg, RS_WARPS * 32, 0, st>>>(*model, theta_ext, P, mask_row_ptr, mask_col, k, out_ids,
                                                      out_scores);
  FVX_CHECK_LAUNCH("k_rescore_select");
  // rows whose candidate lists could not bound the result: exact fp32 sweep, written in place (the
  // thresholds and the statistics are dead by now and serve as the list / its counter)
  return fvx_launch_topk_flagged(model, theta_ext, ws->flags, n_users, u0, mask_row_ptr, mask_col, k, out_ids,
                                 out_scores, reinterpret_cast<int32_t*>(ws->thr), reinterpret_cast<int32_t*>(ws->stat), st);
}

}  // extern "C"
