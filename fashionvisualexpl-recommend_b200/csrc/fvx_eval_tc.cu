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
//   2. k_topk_tc (persistent, warp-specialised, one CTA per SM).  A work unit is (n_ut user tiles of
//      128 rows, an item range); the unit sweeps its range TWICE:
//        warp 0    TMA producer : item tiles [128 x KP] -> 128B-swizzled smem ring (both sweeps); the
//                                 user tiles [128 x KP] of the unit once
//        warp 1    UMMA issuer  : D[ut][128 x 128] (TMEM, fp32) = A_ut * B_tile^T for every user tile
//                                 of the unit (the item tile is read from smem once for 256 users),
//                                 2 * n_ut accumulators = user tiles x double buffer
//        warps 2-9 epilogue     : one thread per user row (single owner), tcgen05.ld 32 columns at a
//                                 time with the next chunk in flight.
//          sweep A (bounds)     : nothing but a max tree - the maximum of every GROUP of columns
//                                 (gchunks 32-column chunks; at most ~512 groups per range) goes to a
//                                 per-CTA scratch [group][row].  The (k + #train)-th largest group
//                                 maximum, minus the rounding margin, is a lower bound tau of the
//                                 (k + #train)-th best true score of the range (distinct groups hold
//                                 distinct items); each thread finds it for its row by bisection over
//                                 its <= 512 group maxima (coalesced, L2-resident).
//          sweep B (candidates) : the same max tree against the now FIXED tau; the few survivors
//                                 (~1.2 (k + #train) per row) are appended to the row's list.
//      No running threshold, no list compaction, no cross-CTA exchange: the first version of this
//      kernel spent 3.4 of its 5.1 ms per 256 users on exactly that (round-2 ncu capture,
//      profiles/r2a_*), whatever the size of the catalog.
//   3. k_rescore_select              exact fp32 re-scoring of the surviving candidates with the
//                                    same fvx_score_one() the fp32 path uses, train-item mask,
//                                    sort, top-k.
//
// Exactness.  With a = [Gu|Tu][u], b = [Gi|theta][i] rounded to bf16 (relative error 2^-9 each)
//   |s_bf16 - s_fp32| <= c * |a| * |b_i| + beta0,   c = 1.003 * 2^-8 + KP * 2^-21  (rounding + fp32
//   accumulation),  beta0 = (2^-17 + KP * 2^-21) * max_i |bias_i|                (hi+lo residual).
// One more K column holds eps_u = c*|a_u| (rounded up to bf16) on the user side and |b_i| (rounded up) on
// the item side, so the UMMA itself delivers s_ub = s_bf16 + eps_u * |b_i| >= s_fp32 - beta0 (an upper bound up to
// the bias residual), and s_ub - 2.001 * eps_u * |b_i| - beta0 <= s_fp32.  A group's entry is max(s_ub over the
// group) - 2.001 * eps_u * (largest |b_i| of the group) - beta0: a lower bound of the true score of the group's
// best column.  If tau is a value that at least kk = k + #train group entries reach, kk distinct items have a true
// score >= tau, so every item of the true top-kk has s >= tau and s_ub >= tau - beta0: the row publishes tau (a
// statement about true scores: item splits and item shards combine theirs with MAX), the candidates sweep keeps
// {s_ub >= tau - beta0} with its own beta0, which loses nothing, and the output equals the fp32 kernel's bit for bit
// (oracle/tc_bound.py restates this arithmetic; tests/test_oracle_tc_bound.py checks it on the CPU, including
// models whose scores are dominated by the biases, where beta0 is what matters).  A row whose list overflows, or whose
// range has fewer than kk groups, is flagged and re-run through the fp32 kernel inside the same call.
#include <cuda_bf16.h>
#include <math_constants.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"
#include "fvx_tc.cuh"

#define TCK_BM 128          // users per tile  (UMMA M)
#define TCK_BN 128          // items per tile  (UMMA N); wide operands (K = 256) take 64 (TckParams.bn)
#define TCK_KB 64           // bf16 elements per K block: 128-byte rows, 128B swizzle (the 64-byte rows of the
                            // first version cost ~2000 cycles per 256 x 128 tile - TMA rows and operand fetch)
#define TCK_CAP 512         // candidate slots per (user, split)
#define TCK_GMAX 512        // groups per (unit, range) the bounds sweep keeps
#define TCK_THREADS 352     // warp 0 producer, warps 1 and 10 UMMA (one per user tile), warps 2-9 epilogue
#define TCK_RS_MAX 1024     // candidates per user the final selection can take (<= 4 splits x ~200)
#define TCK_MIN_SPLIT_ITEMS 16384 // an item split keeps at least this many items (>= 1.5 kk groups of one chunk)
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
// float <-> int32 whose SIGNED order is the float order (atomicMax on int, all-reduce MAX on int32 tensors)
__device__ __forceinline__ int32_t tck_enc(float s) {
  const int32_t b = __float_as_int(s);
  return b >= 0 ? b : (b ^ 0x7FFFFFFF);
}
__device__ __forceinline__ float tck_dec(int32_t e) { return __int_as_float(e >= 0 ? e : (e ^ 0x7FFFFFFF)); }
#define TCK_ENC_NEG_INF ((int32_t)0x807FFFFF)
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
                             int32_t* __restrict__ thr_g, int KP, float c_rel) {
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
    if (lane == 0) { epsa[u - u0] = __bfloat162float(eh); thr_g[u - u0] = TCK_ENC_NEG_INF; }
  }
}

// nb[i] = |[Gi|theta][i]| rounded up to bf16; stat[0] = max_i nb[i], stat[1] = max |bias| (uint bits of non-negative floats)
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
    wmax = fmaxf(wmax, __bfloat162float(nh));
    bmax = fmaxf(bmax, fabsf(bias));
  }
  if (lane == 0) {   // non-negative floats order like uints
    atomicMax(stat, __float_as_uint(wmax));
    atomicMax(stat + 1, __float_as_uint(bmax));
  }
}

// nbc[c] = max of nb over the 32 items of chunk c (the rounding margin of a column group uses the largest item
// norm IN the group, not the largest of the catalog: a few heavy items no longer loosen every row's bound)
__global__ void k_chunk_nbmax(const float* __restrict__ nb, int item_cnt, float* __restrict__ nbc) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n_chunks = (item_cnt + 31) >> 5;
  for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n_chunks; c += warps) {
    const int i = c * 32 + lane;
    float v = i < item_cnt ? nb[i] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) nbc[c] = v;
  }
}

// ---------------------------------------------------------------------------------
struct TckParams {
  int n_users;          // users in this call (rows of A)
  int item_cnt, item_lo;
  int nkb;              // K blocks of 64
  int nk16;             // UMMA K steps that hold data: ceil((K+d+3) / 16) (the zero padding behind them is skipped)
  int stages;
  int n_ut;             // user tiles per work unit: 2, or 1 when two tiles + the item ring do not fit smem
  int bn;               // items per tile: 128, or 64 for wide operands
  // work units: the first n_full user groups (a multiple of the grid size) sweep the whole catalog each;
  // the remaining groups are cut into `splits` item ranges so that the last wave still fills the machine
  int splits, tiles_per_split, n_item_tiles, n_groups, n_full;
  int k;
  int u0;
  int a_stride;         // the bounds sweep visits every a_stride-th tile of the range (1 or 2)
  int do_a, do_b;       // which sweeps this launch runs: bounds (tau -> thr_g), candidates (tau <- thr_g)
  float beta_c;         // 2^-17 + KP * 2^-21: bias residual + fp32 accumulation, per unit of max|bias|
  const float* epsa;    // [n_users] eps_u as multiplied by the UMMA (bf16 value)
  const uint32_t* stat; // [0] max_i |b_i| as multiplied by the UMMA, [1] max |bias|  (float bits)
  const float* nbc;     // [ceil(item_cnt / 32)] max |b_i| over each 32-item chunk
  const int64_t* mask_row_ptr;
  float* gmax;                // [grid][TCK_GMAX][n_ut * 128] group maxima of the unit in flight
  unsigned long long* cand;   // [lists * CAP]
  int32_t* ccount;            // [lists]
  int32_t* flags;             // [n_users]
  int32_t* thr_g;             // [n_users] the row's bound tau (tck_enc: ordered like the float under SIGNED compare), the
                              //   maximum over the item splits of the row - and, item-sharded, over the ranks
};

// list index of (row, split): rows of the full groups own one list, tail rows `splits` lists
__device__ __forceinline__ size_t tck_list(const TckParams& P, int row, int sp) {
  const int full_rows = P.n_full * P.n_ut * TCK_BM;
  return row < full_rows ? (size_t)row : (size_t)full_rows + (size_t)(row - full_rows) * P.splits + sp;
}
// unit w of the global enumeration -> (user group, split, tile range)
__device__ __forceinline__ void tck_unit(const TckParams& P, int w, int& grp, int& sp, int& t0, int& t1) {
  if (w < P.n_full) { grp = w; sp = 0; t0 = 0; t1 = P.n_item_tiles; return; }
  const int x = w - P.n_full;
  grp = P.n_full + x / P.splits;
  sp = x - (x / P.splits) * P.splits;
  t0 = sp * P.tiles_per_split;
  t1 = min(P.n_item_tiles, t0 + P.tiles_per_split);
}
// 32-column chunks per group of a bounds sweep over `chunks` chunks (a power of two, so that a group is a
// whole number of chunks of a tile or a whole number of tiles): at most TCK_GMAX groups
__device__ __forceinline__ int tck_gchunks(int chunks) {
  int g = 1;
  while ((chunks + g - 1) / g > TCK_GMAX) g <<= 1;
  return g;
}

// maximum of a 32-column chunk: 11 three-input maxima, four 8-column sub-maxima on the way
__device__ __forceinline__ float tck_chunk_max(const uint32_t (&v)[32], float (&g)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float a0 = max3(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1]), __uint_as_float(v[8 * q + 2]));
    const float a1 = max3(__uint_as_float(v[8 * q + 3]), __uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
    g[q] = max3(a0, a1, fmaxf(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
  }
  return fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
}

// candidates sweep, one chunk of one row per lane: the 8-column sub-maxima decide warp-wide which octets
// are looked at; inside an octet only the lanes that hold a survivor do anything
__device__ __forceinline__ void tck_scan_chunk(const uint32_t (&v)[32], float thr, int ibase, int ilimit, int item_lo,
                                               unsigned long long* buf, int& cnt) {
  float g[4];
  const float m = tck_chunk_max(v, g);
  if (!__any_sync(0xffffffffu, m >= thr)) return;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (!__any_sync(0xffffffffu, g[q] >= thr)) continue;
    // (a vote per column of a visited octet, to skip the append code nobody needs, was measured slower: 3.0 ->
    // 3.6 ms for the select half at configs[1])
    if (g[q] >= thr) {
#pragma unroll
      for (int j = 8 * q; j < 8 * q + 8; ++j) {
        const float s = __uint_as_float(v[j]);
        if (s >= thr && ibase + j < ilimit) {
          if (cnt < TCK_CAP) buf[cnt] = tck_key(s, item_lo + ibase + j);
          ++cnt;                              // counts past the capacity: the row is flagged at the end
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
  const uint32_t a_bytes = (uint32_t)P.nkb * TCK_BM * 128u;     // one user tile
  const uint32_t b_bytes = (uint32_t)P.nkb * (uint32_t)P.bn * 128u;
  uint8_t* sA = smem;                                            // [n_ut][a_bytes]
  uint8_t* sB = smem + (size_t)P.n_ut * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)P.stages * b_bytes);
  uint64_t* full_b = bars;                  // [stages]
  uint64_t* empty_b = bars + P.stages;      // [stages]
  uint64_t* a_full = bars + 2 * P.stages;   // [1]
  uint64_t* a_empty = a_full + 1;           // [1]
  uint64_t* t_full = a_empty + 1;           // [4]  accumulator = ut * 2 + buffer
  uint64_t* t_empty = t_full + 4;           // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 4);

  if (threadIdx.x == 0) {
    // every UMMA warp (one per user tile) releases an item stage / the user tiles with its own commit
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], P.n_ut); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, P.n_ut);
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
  const int n_units = P.n_full + (P.n_groups - P.n_full) * P.splits;
  // Every role walks the same sequence of item tiles per unit: the bounds sweep (every a_stride-th tile of
  // [t0, t1)), then the candidates sweep (every tile).  seq -> tile:
  //   seq <  na : t0 + seq * a_stride          (sweep A)
  //   seq >= na : t0 + (seq - na)              (sweep B)

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, unit_i = 0;
      for (int w = blockIdx.x; w < n_units; w += gridDim.x, ++unit_i) {
        int grp, sp, t0, t1;
        tck_unit(P, w, grp, sp, t0, t1);
        const int nt = P.do_b ? t1 - t0 : 0, na = P.do_a ? (t1 - t0 + P.a_stride - 1) / P.a_stride : 0;
        mbar_wait(a_empty, (unit_i & 1) ^ 1);
        mbar_expect_tx(a_full, (uint32_t)P.n_ut * a_bytes);
        for (int ut = 0; ut < P.n_ut; ++ut)
          for (int kb = 0; kb < P.nkb; ++kb)
            tma_load_2d(sA + (size_t)ut * a_bytes + (size_t)kb * TCK_BM * 128, &tmA, a_full, kb * TCK_KB,
                        (grp * P.n_ut + ut) * TCK_BM);
        for (int seq = 0; seq < na + nt; ++seq) {
          const int t = seq < na ? t0 + seq * P.a_stride : t0 + (seq - na);
          mbar_wait(&empty_b[stage], phase ^ 1);
          mbar_expect_tx(&full_b[stage], b_bytes);
          uint8_t* dst = sB + (size_t)stage * b_bytes;
          for (int kb = 0; kb < P.nkb; ++kb)
            tma_load_2d(dst + (size_t)kb * P.bn * 128, &tmB, &full_b[stage], kb * TCK_KB, t * P.bn);
          if (++stage == (uint32_t)P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===== UMMA issuers: warp 1 fills the accumulators of user tile 0, warp 10 those of user tile 1 =====
    // The ISSUE of a UMMA, not the tensor pipe, paced the first versions of this kernel: one thread issuing all
    // twelve UMMAs of a 256 x 128 tile pair needed ~190 cycles for each (descriptor arithmetic, R2UR moves and
    // branches at ~15 cycles apiece on a lone warp; profiles/r2_eval_*) against 64 on the tensor pipe, and the
    // epilogue warps slept on t_full.  Two measures: (1) one issuing warp PER USER TILE, on different schedulers,
    // each with its own double-buffered accumulator pair - an item stage is released by both warps' commits
    // (a warp per accumulator, four in all, was measured slower: 1.92 vs 1.59 ms); (2) the warp stays
    // convergent, the descriptors' low words (start address >> 4) advance by adds (+2 per 32-byte K step inside
    // a 128-byte K block, + rows * 8 per K block) and only the instruction itself is predicated on elect.sync
    // (umma_f16_lohi_elect) - inside an `if (lane == 0)` region the compiler wraps every UMMA into an ELECT /
    // BRA.U.ANY emulation loop.
    const int ut = warp == 1 ? 0 : 1;
    if (ut < P.n_ut) {
      const uint32_t idesc = umma_idesc_bf16(TCK_BM, P.bn, 0, 0);
      const int nk = P.nk16;
      const uint64_t d0 = umma_smem_desc(0, 16, 1024, TC_SWZ_128B);
      const uint32_t dlo = (uint32_t)d0, dhi = (uint32_t)(d0 >> 32);      // everything but the start address
      const uint32_t a_base = dlo + ((tc_smem_u32(sA) + (uint32_t)ut * a_bytes) >> 4);
      const uint32_t b_base = dlo + (tc_smem_u32(sB) >> 4), b_st = b_bytes >> 4;
      const uint32_t a_kb = (TCK_BM * 128u) >> 4, b_kb = ((uint32_t)P.bn * 128u) >> 4;
      uint32_t stage = 0, phase = 0, unit_i = 0, buf = 0, buf_phase = 0;
      for (int w = blockIdx.x; w < n_units; w += gridDim.x, ++unit_i) {
        int grp, sp, t0, t1;
        tck_unit(P, w, grp, sp, t0, t1);
        const int nt = P.do_b ? t1 - t0 : 0, na = P.do_a ? (t1 - t0 + P.a_stride - 1) / P.a_stride : 0;
        mbar_wait(a_full, unit_i & 1);
        for (int seq = 0; seq < na + nt; ++seq) {
          const uint32_t acc = ut * 2 + buf;
          mbar_wait(&full_b[stage], phase);
          mbar_wait(&t_empty[acc], buf_phase ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + acc * TCK_BN;
          uint32_t alo = a_base, blo = b_base + stage * b_st;
          for (int k = 0; k < nk; k += 4, alo += a_kb, blo += b_kb) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (k + ks < nk) umma_f16_lohi_elect(d, alo + 2u * ks, dhi, blo + 2u * ks, dhi, idesc, (k + ks) ? 1u : 0u);
          }
          umma_commit_elect(&t_full[acc]);       // accumulator ready for its epilogue warps
          umma_commit_elect(&empty_b[stage]);    // this warp is done with the item stage
          if (++stage == (uint32_t)P.stages) { stage = 0; phase ^= 1; }
          if (++buf == 2) { buf = 0; buf_phase ^= 1; }
        }
        umma_commit_elect(a_empty);              // this warp is done with the user tiles
      }
    }
  } else if (warp < 10 && (warp - 2) >> 2 < P.n_ut) {
    // ===== epilogue: warps 2-5 own user tile 0, warps 6-9 user tile 1; thread = user row =====
    const int ut = (warp - 2) >> 2;
    const int quad = warp & 3;            // TMEM lanes this warp may read: [32*quad, 32*quad+32)
    const int rit = quad * 32 + lane;     // row in tile
    const int rows_u = P.n_ut * TCK_BM;   // rows of a unit
    const float beta0 = P.beta_c * __uint_as_float(P.stat[1]) + 1e-30f;
    float* gm = P.gmax + (size_t)blockIdx.x * TCK_GMAX * rows_u + ut * TCK_BM + rit;   // [group][rows_u], this row
    uint32_t buf = 0, buf_phase = 0;
    for (int w = blockIdx.x; w < n_units; w += gridDim.x) {
      int grp, sp, t0, t1;
      tck_unit(P, w, grp, sp, t0, t1);
      const int nt = P.do_b ? t1 - t0 : 0, na = P.do_a ? (t1 - t0 + P.a_stride - 1) / P.a_stride : 0;
      const int row = (grp * P.n_ut + ut) * TCK_BM + rit;
      const bool live = row < P.n_users;
      int kk = P.k;
      float eps2 = 0.0f;                        // 2.001 * eps_u: times the item norm = the width of the rounding band
      if (live) {
        const int gu = P.u0 + row;
        kk = P.k + (int)(P.mask_row_ptr[gu + 1] - P.mask_row_ptr[gu]);
        eps2 = 2.001f * P.epsa[row];
      }
      // ---- sweep A: group maxima of s_ub -> gm[g * rows_u] ----
      const int nch = P.bn >> 5;                // 32-column chunks per tile
      const int gch = tck_gchunks(na * nch);    // chunks per group (power of two)
      const int gsh = 31 - __clz(gch);
      const int n_grp = (na * nch + gch - 1) >> gsh;
      // a group's entry is a LOWER bound of the true score of its best column: max s_ub - eps2 * (largest item norm
      // of the group) - beta0
      float run = -CUDART_INF_F, run_nb = 0.0f, lo = CUDART_INF_F, hi = -CUDART_INF_F;
      int cdone = 0;                            // chunks of the sweep consumed so far
      int n_fin = 0;                            // groups that hold at least one catalog column
      for (int seq = 0; seq < na; ++seq) {
        const int t = t0 + seq * P.a_stride;
        const uint32_t acc = ut * 2 + buf;
        mbar_wait(&t_full[acc], buf_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TCK_BN;
        const int ragged = (t + 1) * P.bn - P.item_cnt;       // > 0: the last tile holds columns past the catalog
        uint32_t va[32], vb[32];
        tmem_ld_32x32(taddr, va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {
            tmem_ld_wait();
            if (c + 1 < nch) {                  // next chunk in flight while this one is reduced
              if (c & 1) tmem_ld_32x32(taddr + (c + 1) * 32, va); else tmem_ld_32x32(taddr + (c + 1) * 32, vb);
            }
            uint32_t (&v)[32] = (c & 1) ? vb : va;
            if (ragged > 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (t * P.bn + c * 32 + j >= P.item_cnt) v[j] = 0xFF800000u;     // -inf
            }
            float g4[4];
            run = fmaxf(run, tck_chunk_max(v, g4));
            {
              const int ch = ((t * P.bn) >> 5) + c;                                // warp-uniform address
              if (ch < ((P.item_cnt + 31) >> 5)) run_nb = fmaxf(run_nb, __ldg(P.nbc + ch));
            }
            if ((++cdone & (gch - 1)) == 0) {
              const float lb = run - eps2 * run_nb - beta0;
              gm[(size_t)((cdone >> gsh) - 1) * rows_u] = lb;
              if (run > -CUDART_INF_F) { lo = fminf(lo, lb); hi = fmaxf(hi, lb); ++n_fin; }
              run = -CUDART_INF_F;
              run_nb = 0.0f;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acc]);
        if (++buf == 2) { buf = 0; buf_phase ^= 1; }
      }
      if ((cdone & (gch - 1)) != 0) {           // the last, shorter group
        const float lb = run - eps2 * run_nb - beta0;
        gm[(size_t)(cdone >> gsh) * rows_u] = lb;
        if (run > -CUDART_INF_F) { lo = fminf(lo, lb); hi = fmaxf(hi, lb); ++n_fin; }
      }
      // ---- the row's bound: a value at least kk group maxima reach (bisection; counts in [kk, kk + kk/4]
      //      stop it early - a looser value only lets a few more candidates through) ----
      float thr = CUDART_INF_F;                 // dead rows keep nothing
      if (live && P.do_a) {
        if (n_fin >= kk) {                      // (fewer groups than kk - a tiny range - bound nothing)
          float a = lo, b = hi;                 // invariant: count(>= a) >= kk (lo: the smallest finite maximum)
          for (int it = 0; it < 16 && a < b; ++it) {
            const float mid = 0.5f * a + 0.5f * b;
            if (!(mid > a) || !(mid < b)) break;
            int c = 0;
            for (int g = 0; g < n_grp; ++g) c += (__ldcg(gm + (size_t)g * rows_u) >= mid) ? 1 : 0;
            if (c >= kk) { a = mid; if (c <= kk + (kk >> 2)) break; } else b = mid;
          }
          // the row's bound is the best one any range (item split, item shard) finds: a statement about TRUE
          // scores (the k-th best unmasked one is >= it); whoever compares s_ub with it subtracts its own beta0
          atomicMax(P.thr_g + row, tck_enc(a));
        }
      }
      if (!P.do_b) continue;
      // a launch that runs both sweeps reads its own bound back (other splits may have raised it meanwhile); a
      // candidates-only launch finds the maximum over all splits and ranks.  -inf (no range could bound the row):
      // everything passes, the list overflows, the exact kernel takes the row.
      // (minus beta0: s_ub may lie below the true score by the bias residual of THIS shard's items)
      if (live) thr = tck_dec(__ldcg(P.thr_g + row)) - beta0;
      // ---- sweep B: candidates with s_ub >= thr ----
      int cnt = 0;
      unsigned long long* lbuf = P.cand + tck_list(P, live ? row : 0, sp) * TCK_CAP;
      for (int seq = 0; seq < nt; ++seq) {
        const int t = t0 + seq;
        const uint32_t acc = ut * 2 + buf;
        mbar_wait(&t_full[acc], buf_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TCK_BN;
        const int item0 = t * P.bn;
        uint32_t va[32], vb[32];
        tmem_ld_32x32(taddr, va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {
            tmem_ld_wait();
            if (c + 1 < nch) {                  // next chunk in flight while this one is tested
              if (c & 1) tmem_ld_32x32(taddr + (c + 1) * 32, va); else tmem_ld_32x32(taddr + (c + 1) * 32, vb);
            }
            if ((c & 1) == 0) tck_scan_chunk(va, thr, item0 + c * 32, P.item_cnt, P.item_lo, lbuf, cnt);
            else tck_scan_chunk(vb, thr, item0 + c * 32, P.item_cnt, P.item_lo, lbuf, cnt);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acc]);
        if (++buf == 2) { buf = 0; buf_phase ^= 1; }
      }
      if (live) {
        if (cnt > TCK_CAP) { P.flags[row] = 1; cnt = 0; }
        P.ccount[tck_list(P, row, sp)] = cnt;
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

// ---------------------------------------------------------------------------------
// exact re-scoring + final selection: one warp per user
#define RS_WARPS 4
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

// Shared memory of one warp of k_rescore_select (11 KB: five 4-warp blocks per SM).
#define RS_MASK 256         // train items of a user staged in shared memory (longer lists: searched in global memory)
struct RsWarpSmem {
  unsigned long long keys[TCK_RS_MAX];
  float user[448];                      // K + d <= 445 (KP <= 448)
  int32_t mask[RS_MASK];
};

__device__ __forceinline__ bool rs_in_mask(const int32_t* __restrict__ m, int n, int32_t x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int32_t v = m[mid];
    if (v == x) return true;
    if (v < x) lo = mid + 1; else hi = mid;
  }
  return false;
}

// rs_score with the row pieces fetched EIGHT 16-byte loads at a time ahead of their 32 fmaf: the chain is the same
// (index order over [Gi | theta], then the two biases - fvx_score_one bit for bit), but a candidate costs three
// L2 round trips at K + d = 84 instead of twenty-one (the loop above is load -> 4 fmaf -> load ...: its loads sit
// behind the chain in program order and ptxas does not hoist them across a loop of unknown trip count; the kernel
// spent 1.0 ms of the 3.4 ms sweep at configs[1] waiting for them, profiles/r2_eval_full.txt).  K % 4 == 0.
__device__ __forceinline__ float rs_score8(const float* __restrict__ us, const float* __restrict__ irow,
                                           const float* __restrict__ th, int K, int d) {
  const int kc = K >> 2, nc = kc + ((d + 3) >> 2), kd = K + d;
  const float tail_b = irow[K];
  const float tail_v = d > 0 ? th[d] : 0.0f;
  float s = 0.0f;
  for (int g0 = 0; g0 < nc; g0 += 8) {
    float4 x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = g0 + j;
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < nc) x[j] = *reinterpret_cast<const float4*>(g < kc ? irow + 4 * g : th + 4 * (g - kc));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 4 * (g0 + j);                 // index into [Gu | Tu]: the latent columns end on a multiple of 4
      if (c < kd) {
        const float4 u = *reinterpret_cast<const float4*>(us + c);
        s = fmaf(u.x, x[j].x, s);
        if (c + 1 < kd) s = fmaf(u.y, x[j].y, s);
        if (c + 2 < kd) s = fmaf(u.z, x[j].z, s);
        if (c + 3 < kd) s = fmaf(u.w, x[j].w, s);
      }
    }
  }
  s += tail_b;
  if (d > 0) s += tail_v;
  return s;
}

// One warp per user row, one lane per candidate.
__global__ void __launch_bounds__(RS_WARPS * 32, 5)
k_rescore_select(FvxModel M, const float* __restrict__ theta, TckParams P,
                 const int64_t* __restrict__ mask_row_ptr, const int32_t* __restrict__ mask_col, int k,
                 int32_t* __restrict__ out_ids, float* __restrict__ out_scores) {
  __shared__ __align__(16) RsWarpSmem rs_sm[RS_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RsWarpSmem& W = rs_sm[warp];
  unsigned long long* kk = W.keys;
  float* us = W.user;
  const int full_rows = P.n_full * P.n_ut * TCK_BM;
  const int K = M.K, d = M.d, kd = K + d;
  const bool vec = (K & 3) == 0;          // 16-byte pieces must not straddle the end of the latent columns
  for (int r = blockIdx.x * RS_WARPS + warp; r < P.n_users; r += gridDim.x * RS_WARPS) {
    const int gu = P.u0 + r;
    const float* urow = M.users.w + (size_t)gu * M.users.stride;
    for (int c = lane; c < kd; c += 32) us[c] = urow[c];
    const long long mlo = mask_row_ptr[gu], mhi = mask_row_ptr[gu + 1];
    const int mlen = (int)(mhi - mlo);
    const bool mask_sh = mlen <= RS_MASK;
    if (mask_sh)
      for (int c = lane; c < mlen; c += 32) W.mask[c] = mask_col[mlo + c];
    // the candidates of the row's lists (one per item split)
    const int nlists = r < full_rows ? 1 : P.splits;
    int total = 0;
    for (int sp = 0; sp < nlists; ++sp) {
      const size_t li = tck_list(P, r, sp);
      const int n = P.ccount[li];
      const unsigned long long* src = P.cand + li * TCK_CAP;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const unsigned long long key = i < n ? src[i] : KEY_PAD;
        const bool keep = i < n;
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
      const bool masked = mask_sh ? rs_in_mask(W.mask, mlen, gid) : fvx_in_sorted(mask_col, mlo, mhi, gid);
      if (!masked) {
        const int32_t li = gid - M.item_lo;
        const float* irow = M.items.w + (size_t)li * M.items.stride;
        const float* th = d > 0 ? theta + (size_t)li * M.de : nullptr;
        key = tck_key(vec ? rs_score8(us, irow, th, K, d) : rs_score(us, irow, th, K, d), gid);
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
struct TckGeom {
  int KP, nkb, n_ut, bn, stages, n_groups, n_full, splits, grid;
  long long lists, gmax_elems;
  size_t smem;
};
static int tck_geometry(const FvxModel* m, int n_users, TckGeom* g) {
  const int kd3 = m->K + m->d + 3;
  g->KP = (kd3 + TCK_KB - 1) / TCK_KB * TCK_KB;
  g->nkb = g->KP / TCK_KB;
  // shared memory: n_ut user tiles + a ring of item tiles (each nkb * 8 KB).  Two user tiles per unit read
  // every item tile once for 256 users; wide operands take one user tile.
  // shared memory: n_ut user tiles (nkb * 16 KB each) + a ring of item tiles (nkb * bn * 128 B each).  Two user
  // tiles per unit read every item tile once for 256 users; wide operands (K = 256: 5 K blocks) take one user
  // tile and 64-item tiles.
  const long long ut_bytes = (long long)g->nkb * TCK_BM * 128, budget = 224 * 1024;
  static const int pref[6][3] = {{2, 128, 3}, {1, 128, 3}, {2, 128, 2}, {1, 64, 3}, {1, 128, 2}, {1, 64, 2}};
  g->stages = 0;
  for (int i = 0; i < 6 && g->stages == 0; ++i) {
    const long long it_bytes = (long long)g->nkb * pref[i][1] * 128;
    const long long st = (budget - pref[i][0] * ut_bytes) / it_bytes;
    if (st >= pref[i][2]) { g->n_ut = pref[i][0]; g->bn = pref[i][1]; g->stages = st > 6 ? 6 : (int)st; }
  }
  if (g->stages == 0) return -1;
  g->smem = (size_t)g->n_ut * ut_bytes + (size_t)g->stages * g->nkb * g->bn * 128 + (2 * g->stages + 10) * 8 + 16 + 1024;
  const int rows_u = g->n_ut * TCK_BM;
  g->n_groups = (n_users + rows_u - 1) / rows_u;
  const int n_item_tiles = (m->item_cnt + g->bn - 1) / g->bn;
  const int min_split = TCK_MIN_SPLIT_ITEMS / g->bn;
  const int G = fvx_num_sms();
  g->n_full = (g->n_groups / G) * G;
  const int tail = g->n_groups - g->n_full;
  int s = 1;
  if (tail > 0) {
    // the tail groups are cut into s item ranges so that their units still fill the machine: the tail then
    // takes ceil(tail*s/G) rounds of 1/s of a sweep.  A range keeps >= TCK_MIN_SPLIT_ITEMS items (its bound
    // needs more groups than kk) and a row at most 4 lists; the smallest s within 2 % of the best is taken.
    int smax = n_item_tiles / min_split < 4 ? n_item_tiles / min_split : 4;
    if (smax < 1) smax = 1;
    double best = 1e30;
    for (int c = 1; c <= smax; ++c) {
      const double t = (double)((tail * c + G - 1) / G) / c + 0.002 * c;
      if (t < best * 0.98) { best = t; s = c; }
    }
  }
  g->splits = s;
  const long long units = (long long)g->n_full + (long long)tail * s;
  g->grid = (int)(units < G ? units : G);
  const long long full_rows = (long long)g->n_full * rows_u;
  g->lists = full_rows + ((long long)n_users - full_rows > 0 ? ((long long)n_users - full_rows) * s : 0);
  g->gmax_elems = (long long)G * TCK_GMAX * rows_u;
  return 0;
}

int fvx_eval_ws_query(const FvxModel* model, int32_t n_users, FvxEvalWs* ws) {
  FVX_CHECK_ARG(model && ws && n_users > 0, "fvx_eval_ws_query: bad arguments");
  TckGeom g;
  FVX_CHECK_ARG(tck_geometry(model, n_users, &g) == 0,
                "fvx_eval_ws_query: K+d+3=%d too wide for the tensor-core sweep (use fvx_score_topk)",
                model->K + model->d + 3);
  ws->KP = g.KP;
  ws->splits = g.splits;
  ws->cap = TCK_CAP;
  ws->u_cap = n_users;
  ws->i_cap = model->item_cnt;
  ws->lists = g.lists;
  ws->gmax_elems = g.gmax_elems;
  ws->n_ut = g.n_ut;
  return 0;
}

// what both launches share: argument checks, geometry, tensor maps, kernel parameters
struct TckLaunch {
  TckGeom g;
  TckParams P;
  CUtensorMap tmA, tmB;
  int n_users;
};
static int tck_setup(TckLaunch* L, const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                     const int64_t* mask_row_ptr, int32_t k, const FvxEvalWs* ws, const char* who) {
  FVX_CHECK_ARG(model && model->abi_version == FVX_ABI_VERSION && ws, "%s: bad model / workspace", who);
  FVX_CHECK_ARG(model->d == 0 || theta_ext, "%s: VBPR scoring needs theta_ext", who);
  FVX_CHECK_ARG(0 <= u0 && u0 < u1 && u1 <= model->num_users, "%s: bad user range", who);
  FVX_CHECK_ARG(k >= 1 && k <= 128, "%s: k=%d outside [1,128]", who, k);
  FVX_CHECK_ARG(mask_row_ptr != nullptr, "%s: null mask", who);
  const int n_users = u1 - u0;
  L->n_users = n_users;
  TckGeom& g = L->g;
  FVX_CHECK_ARG(tck_geometry(model, n_users, &g) == 0, "%s: K+d+3=%d too wide for the tensor-core sweep (use fvx_score_topk)",
                who, model->K + model->d + 3);
  FVX_CHECK_ARG(ws->KP == g.KP && ws->splits == g.splits && ws->cap == TCK_CAP && ws->u_cap >= n_users &&
                ws->i_cap >= model->item_cnt && ws->lists >= g.lists && ws->gmax_elems >= g.gmax_elems &&
                ws->n_ut == g.n_ut, "%s: workspace does not match fvx_eval_ws_query", who);
  FVX_CHECK_ARG(ws->A && ws->Bm && ws->epsa && ws->nb && ws->nbc && ws->stat && ws->cand && ws->ccount && ws->flags &&
                ws->thr && ws->gmax, "%s: null workspace buffer", who);
  const int KP = g.KP;
  int rc = tc_make_tensor_map_bf16(&L->tmA, ws->A, n_users, KP, (uint64_t)KP * 2, TCK_KB, TCK_BM, 3);
  if (rc == 0) rc = tc_make_tensor_map_bf16(&L->tmB, ws->Bm, model->item_cnt, KP, (uint64_t)KP * 2, TCK_KB, g.bn, 3);
  if (rc != 0) FVX_FAIL(-4, "%s: cuTensorMapEncodeTiled failed (%d)", who, rc);
  TckParams& P = L->P;
  P.n_users = n_users; P.item_cnt = model->item_cnt; P.item_lo = model->item_lo; P.nkb = g.nkb;
  P.nk16 = (model->K + model->d + 3 + 15) / 16;
  P.n_ut = g.n_ut; P.bn = g.bn; P.n_groups = g.n_groups; P.n_full = g.n_full;
  P.n_item_tiles = (model->item_cnt + g.bn - 1) / g.bn;
  P.splits = g.splits;
  P.tiles_per_split = (P.n_item_tiles + P.splits - 1) / P.splits;
  P.k = k; P.u0 = u0; P.epsa = ws->epsa; P.stat = reinterpret_cast<const uint32_t*>(ws->stat); P.nbc = ws->nbc;
  P.a_stride = ws->a_stride == 2 ? 2 : 1;
  P.do_a = P.do_b = 1;
  P.beta_c = 7.6294e-6f + (float)KP * 4.76837158e-7f;
  P.mask_row_ptr = mask_row_ptr;
  P.gmax = ws->gmax;
  P.cand = reinterpret_cast<unsigned long long*>(ws->cand); P.ccount = ws->ccount; P.flags = ws->flags;
  P.thr_g = reinterpret_cast<int32_t*>(ws->thr);
  P.stages = g.stages;
  static FvxSmemMark topk_tc_smem;
  return fvx_ensure_smem((const void*)k_topk_tc, &topk_tc_smem, g.smem, who);
}

// Bounds: operands packed, then one sweep of every work unit for the group maxima; every row's bound tau lands in
// ws->thr (int32, ordered like the float under signed compare: an item-sharded caller takes the MAXIMUM over the
// ranks before fvx_score_topk_tc_select - each shard's bound is a valid bound of the whole catalog's top).
int fvx_score_topk_tc_bounds(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                             const int64_t* mask_row_ptr, int32_t k, const FvxEvalWs* ws, fvx_stream_t stream) {
  TckLaunch L;
  if (int rc = tck_setup(&L, model, theta_ext, u0, u1, mask_row_ptr, k, ws, "fvx_score_topk_tc_bounds")) return rc;
  cudaStream_t st = fvx_cu(stream);
  const int n_users = L.n_users, KP = L.g.KP;
  cudaMemsetAsync(ws->stat, 0, 8, st);
  cudaMemsetAsync(ws->flags, 0, sizeof(int32_t) * n_users, st);
  const float c_rel = 1.003f * 0.00390625f + (float)KP * 4.76837158e-7f;
  int gr = (n_users * 32 + 255) / 256;
  if (gr > fvx_num_sms() * 8) gr = fvx_num_sms() * 8;
  k_pack_users<<<gr, 256, 0, st>>>(*model, u0, u1, reinterpret_cast<__nv_bfloat16*>(ws->A), ws->epsa,
                                   reinterpret_cast<int32_t*>(ws->thr), KP, c_rel);
  gr = fvx_num_sms() * 8;
  k_pack_items<<<gr, 256, 0, st>>>(*model, theta_ext, reinterpret_cast<__nv_bfloat16*>(ws->Bm), ws->nb,
                                   reinterpret_cast<uint32_t*>(ws->stat), KP);
  k_chunk_nbmax<<<gr, 256, 0, st>>>(ws->nb, model->item_cnt, ws->nbc);
  FVX_CHECK_LAUNCH("k_pack");
  L.P.do_b = 0;
  k_topk_tc<<<L.g.grid, TCK_THREADS, L.g.smem, st>>>(L.tmA, L.tmB, L.P);
  FVX_CHECK_LAUNCH("k_topk_tc (bounds)");
  return 0;
}

// Candidates: one sweep with the bounds in ws->thr, exact fp32 re-scoring of the survivors, top-k; rows whose lists
// overflow go through the exact kernel inside the call.
int fvx_score_topk_tc_select(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                             const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k, int32_t* out_ids,
                             float* out_scores, const FvxEvalWs* ws, fvx_stream_t stream) {
  TckLaunch L;
  if (int rc = tck_setup(&L, model, theta_ext, u0, u1, mask_row_ptr, k, ws, "fvx_score_topk_tc_select")) return rc;
  FVX_CHECK_ARG(out_ids && out_scores, "fvx_score_topk_tc_select: null output");
  cudaStream_t st = fvx_cu(stream);
  const int n_users = L.n_users;
  cudaMemsetAsync(ws->ccount, 0, sizeof(int32_t) * L.g.lists, st);
  L.P.do_a = 0;
  k_topk_tc<<<L.g.grid, TCK_THREADS, L.g.smem, st>>>(L.tmA, L.tmB, L.P);
  FVX_CHECK_LAUNCH("k_topk_tc (candidates)");
  long long rg = ((long long)n_users + RS_WARPS - 1) / RS_WARPS;
  if (rg > (long long)fvx_num_sms() * 8) rg = (long long)fvx_num_sms() * 8;
  k_rescore_select<<<(int)rg, RS_WARPS * 32, 0, st>>>(*model, theta_ext, L.P, mask_row_ptr, mask_col, k, out_ids,
                                                      out_scores);
  FVX_CHECK_LAUNCH("k_rescore_select");
  // rows whose candidate lists could not bound the result: exact fp32 sweep, written in place, each row's
  // threshold seeded with its proven bound (ws->ccount is the scratch for their list - the counts are dead by now
  // and cleared at the start of every select; the statistics serve as the list's counter)
  return fvx_launch_topk_flagged(model, theta_ext, ws->flags, n_users, u0, mask_row_ptr, mask_col, k, out_ids,
                                 out_scores, ws->ccount, reinterpret_cast<int32_t*>(ws->stat), st,
                                 reinterpret_cast<const int32_t*>(ws->thr));
}

int fvx_score_topk_tc(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                      const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k, int32_t* out_ids,
                      float* out_scores, const FvxEvalWs* ws, fvx_stream_t stream) {
  if (int rc = fvx_score_topk_tc_bounds(model, theta_ext, u0, u1, mask_row_ptr, k, ws, stream)) return rc;
  return fvx_score_topk_tc_select(model, theta_ext, u0, u1, mask_row_ptr, mask_col, k, out_ids, out_scores, ws, stream);
}

}  // extern "C"
