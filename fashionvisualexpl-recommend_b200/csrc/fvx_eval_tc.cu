// Full-catalog top-k on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// predict_all (BPRMF.py:85, VBPR.py:95-97) is the contraction
//     S[u,i] = <[Gu|Tu|1][u,:], [Gi|theta|Bi+vbias][i,:]>
// and Evaluator.store_recommendation (Evaluator.py:231-237) wants, per user, the k best
// non-train items.  The sweep below never writes S:
//
//   1. k_pack_users / k_pack_items   bf16 operands (K+d+1 padded to KP), fp32 row norms
//   2. k_topk_tc (persistent, warp-specialised)
//        warp 0  TMA producer : item tiles [256 x KP] -> 64B-swizzled smem ring
//        warp 1  UMMA issuer  : D[128 x 256] (TMEM, fp32) = A_users[128 x KP] * B_items^T
//        warps 2-5 epilogue   : tcgen05.ld 32 columns at a time, one thread per user
//                               row; a 3-input max tree tests the 32 scores against the
//                               row's running threshold; survivors go to the row's
//                               candidate list; a warp-cooperative radix select tightens
//                               the threshold when the list fills up
//   3. k_rescore_select              exact fp32 re-scoring of the candidates with the
//                                    same fvx_score_one() the fp32 path uses, train-item
//                                    mask, sort, top-k.
//
// Exactness: bf16 rounding of both operands bounds |s_bf16 - s_fp32| by
// eps_u = 1.01 * 2^-7 * |a_u| * max_i |b_i| (Cauchy-Schwarz).  A row keeps every item with
// s_bf16 >= tau - 2*eps_u, tau = the k-th best bf16 score among its non-train items so
// far, so the true top-k is always among the candidates and the output equals the fp32
// kernel's bit for bit.  If a row's list overflows (more than CAP items inside the
// margin) the row is flagged and the caller re-runs it through the fp32 kernel.
#include <cuda_bf16.h>
#include <math_constants.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"
#include "fvx_tc.cuh"

#define TCK_BM 128          // users per tile  (UMMA M)
#define TCK_BN 256          // items per tile  (UMMA N)
#define TCK_KB 32           // bf16 elements per K block (64-byte swizzle rows)
#define TCK_CAP 512         // candidate slots per (user, split)
#define TCK_THREADS 192
#define KEY_PAD 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ unsigned long long tck_key(float s, int32_t id) {
  const uint32_t b = __float_as_uint(s);
  const uint32_t mono = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)(~mono) << 32) | (uint32_t)id;
}
__device__ __forceinline__ float tck_score_of_hi(uint32_t hi) {
  const uint32_t mono = ~hi;
  const uint32_t b = (mono & 0x80000000u) ? (mono ^ 0x80000000u) : ~mono;
  return __uint_as_float(b);
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// ---------------------------------------------------------------------------------
__global__ void k_pack_users(FvxModel M, int u0, int u1, __nv_bfloat16* __restrict__ A, float* __restrict__ unorm,
                             int KP) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int kd = M.K + M.d;
  for (int u = u0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); u < u1; u += warps) {
    const float* src = M.users.w + (size_t)u * M.users.stride;
    float sq = 0.0f;
    for (int c = lane; c < KP; c += 32) {
      const float v = c < kd ? src[c] : (c == kd ? 1.0f : 0.0f);
      sq += v * v;
      A[(size_t)(u - u0) * KP + c] = __float2bfloat16_rn(v);
    }
    sq = fvx_warp_sum(sq);
    if (lane == 0) unorm[u - u0] = sqrtf(sq);
  }
}

__global__ void k_pack_items(FvxModel M, const float* __restrict__ theta, __nv_bfloat16* __restrict__ Bm,
                             uint32_t* __restrict__ bmax_bits, int KP) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int K = M.K, d = M.d, kd = K + d;
  float wmax = 0.0f;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < M.item_cnt; i += warps) {
    const float* row = M.items.w + (size_t)i * M.items.stride;
    const float* th = d > 0 ? theta + (size_t)i * M.de : nullptr;
    float sq = 0.0f;
    for (int c = lane; c < KP; c += 32) {
      float v = 0.0f;
      if (c < K) v = row[c];
      else if (c < kd) v = th[c - K];
      else if (c == kd) v = row[K] + (d > 0 ? th[d] : 0.0f);
      sq += v * v;
      Bm[(size_t)i * KP + c] = __float2bfloat16_rn(v);
    }
    sq = fvx_warp_sum(sq);
    wmax = fmaxf(wmax, sqrtf(sq));
  }
  if (lane == 0) atomicMax(bmax_bits, __float_as_uint(wmax));   // non-negative floats order like uints
}

// ---------------------------------------------------------------------------------
struct TckParams {
  int n_users;          // users in this call (rows of A)
  int item_cnt, item_lo;
  int nkb;              // K blocks of 32
  int stages;
  int splits, tiles_per_split, n_item_tiles, n_user_tiles;
  int k;
  int u0;
  const float* unorm;
  const uint32_t* bmax_bits;
  const int64_t* mask_row_ptr;
  const int32_t* mask_col;
  unsigned long long* cand;   // [n_users * splits * CAP]
  int32_t* ccount;            // [n_users * splits]
  int32_t* flags;             // [n_users]
};

// warp-cooperative: tighten the threshold of lane `L`'s row and prune its candidate list
__device__ __forceinline__ void tck_compact_row(const TckParams& P, int L, int lane, int my_row, int split,
                                                float my_margin, int& cnt, float& thr) {
  const int row = __shfl_sync(0xffffffffu, my_row, L);
  const int n = __shfl_sync(0xffffffffu, cnt, L);
  const float margin = __shfl_sync(0xffffffffu, my_margin, L);
  unsigned long long* buf = P.cand + ((size_t)row * P.splits + split) * TCK_CAP;
  const int gu = P.u0 + row;
  const long long mlo = P.mask_row_ptr[gu], mhi = P.mask_row_ptr[gu + 1];
  unsigned long long e[TCK_CAP / 32];
#pragma unroll
  for (int q = 0; q < TCK_CAP / 32; ++q) {
    const int idx = q * 32 + lane;
    unsigned long long key = idx < n ? buf[idx] : KEY_PAD;
    if (key != KEY_PAD && fvx_in_sorted(P.mask_col, mlo, mhi, (int32_t)(key & 0xFFFFFFFFu))) key = KEY_PAD;
    e[q] = key;
  }
  int valid = 0;
#pragma unroll
  for (int q = 0; q < TCK_CAP / 32; ++q) valid += (e[q] != KEY_PAD) ? 1 : 0;
  valid = __reduce_add_sync(0xffffffffu, valid);
  float new_thr = -CUDART_INF_F;
  if (valid >= P.k) {
    // smallest 32-bit score key T with #(hi32 <= T) >= k  == key of the k-th best score
    uint32_t lo = 0u, hi = 0xFFFFFFFEu;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int c = 0;
#pragma unroll
      for (int q = 0; q < TCK_CAP / 32; ++q) c += ((uint32_t)(e[q] >> 32) <= mid) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= P.k) hi = mid; else lo = mid + 1;
    }
    new_thr = tck_score_of_hi(lo) - margin;
  }
  int out = 0;
  __syncwarp();
#pragma unroll
  for (int q = 0; q < TCK_CAP / 32; ++q) {
    const bool keep = e[q] != KEY_PAD && tck_score_of_hi((uint32_t)(e[q] >> 32)) >= new_thr;
    const uint32_t b = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[out + __popc(b & ((1u << lane) - 1u))] = e[q];
    out += __popc(b);
  }
  __syncwarp();
  if (out > TCK_CAP - 64) {          // too many items inside the margin: row is re-run in fp32
    if (lane == 0) P.flags[row] = 1;
    out = TCK_CAP - 64;
  }
  if (lane == L) { cnt = out; thr = new_thr; }
}

__global__ void __launch_bounds__(TCK_THREADS, 1)
k_topk_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TckParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // swizzled TMA / UMMA tiles need their base aligned to the swizzle repeat: round up by hand
  uint8_t* smem = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = (uint32_t)P.nkb * TCK_BM * 64u;
  const uint32_t b_bytes = (uint32_t)P.nkb * TCK_BN * 64u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)P.stages * b_bytes);
  uint64_t* full_b = bars;                  // [stages]
  uint64_t* empty_b = bars + P.stages;      // [stages]
  uint64_t* a_full = bars + 2 * P.stages;   // [1]
  uint64_t* a_empty = a_full + 1;           // [1]
  uint64_t* t_full = a_empty + 1;           // [2]
  uint64_t* t_empty = t_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_units = P.n_user_tiles * P.splits;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, unit_i = 0;
      for (int w = blockIdx.x; w < n_units; w += gridDim.x, ++unit_i) {
        const int ut = w / P.splits, sp = w - ut * P.splits;
        mbar_wait(a_empty, (unit_i & 1) ^ 1);
        mbar_expect_tx(a_full, a_bytes);
        for (int kb = 0; kb < P.nkb; ++kb)
          tma_load_2d(sA + (size_t)kb * TCK_BM * 64, &tmA, a_full, kb * TCK_KB, ut * TCK_BM);
        const int t0 = sp * P.tiles_per_split;
        const int t1 = min(P.n_item_tiles, t0 + P.tiles_per_split);
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
      uint32_t stage = 0, phase = 0, unit_i = 0, acc = 0, acc_phase = 0;
      for (int w = blockIdx.x; w < n_units; w += gridDim.x, ++unit_i) {
        const int sp = w % P.splits;
        mbar_wait(a_full, unit_i & 1);
        const int t0 = sp * P.tiles_per_split;
        const int t1 = min(P.n_item_tiles, t0 + P.tiles_per_split);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&t_empty[acc], acc_phase ^ 1);
          mbar_wait(&full_b[stage], phase);
          tc_fence_after();
          const uint32_t a0 = tc_smem_u32(sA), b0 = tc_smem_u32(sB + (size_t)stage * b_bytes);
          const uint32_t d = tmem_base + acc * TCK_BN;
          for (int kb = 0; kb < P.nkb; ++kb) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t ad = umma_smem_desc(a0 + kb * TCK_BM * 64 + ks * 32, 16, 512, TC_SWZ_64B);
              const uint64_t bd = umma_smem_desc(b0 + kb * TCK_BN * 64 + ks * 32, 16, 512, TC_SWZ_64B);
              umma_f16(d, ad, bd, idesc, (kb | ks) ? 1u : 0u);
            }
          }
          umma_commit(&empty_b[stage]);    // smem stage may be refilled once these UMMAs retire
          umma_commit(&t_full[acc]);       // accumulator ready for the epilogue
          if (++stage == (uint32_t)P.stages) { stage = 0; phase ^= 1; }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        umma_commit(a_empty);              // user tile may be overwritten
      }
    }
  } else {
    // ===== epilogue: one thread per user row =====
    const int quad = warp & 3;
    const float bmax = __uint_as_float(*P.bmax_bits);
    uint32_t acc = 0, acc_phase = 0;
    for (int w = blockIdx.x; w < n_units; w += gridDim.x) {
      const int ut = w / P.splits, sp = w - ut * P.splits;
      const int row = ut * TCK_BM + quad * 32 + lane;
      const bool live = row < P.n_users;
      const float margin = live ? 2.0f * 1.01f * 0.0078125f * P.unorm[row] * bmax + 1e-30f : 0.0f;
      float thr = live ? -CUDART_INF_F : CUDART_INF_F;
      int cnt = 0;
      unsigned long long* buf = P.cand + ((size_t)(live ? row : 0) * P.splits + sp) * TCK_CAP;
      const int t0 = sp * P.tiles_per_split;
      const int t1 = min(P.n_item_tiles, t0 + P.tiles_per_split);
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&t_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TCK_BN;
        const int item0 = t * TCK_BN;
#pragma unroll 1
        for (int c = 0; c < TCK_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          float m = max3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
#pragma unroll
          for (int j = 3; j < 31; j += 2) m = max3(m, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
          m = fmaxf(m, __uint_as_float(v[31]));
          if (m >= thr) {
            const int ibase = item0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = __uint_as_float(v[j]);
              if (s >= thr && ibase + j < P.item_cnt && cnt < TCK_CAP) {
                buf[cnt] = tck_key(s, P.item_lo + ibase + j);
                ++cnt;
              }
            }
          }
          uint32_t need = __ballot_sync(0xffffffffu, cnt > TCK_CAP - 32);
          while (need) {
            const int L = __ffs(need) - 1;
            need &= need - 1;
            tck_compact_row(P, L, lane, row, sp, margin, cnt, thr);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      // final prune of every live row, then publish the list length
      uint32_t need = __ballot_sync(0xffffffffu, live);
      while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        tck_compact_row(P, L, lane, row, sp, margin, cnt, thr);
      }
      if (live) P.ccount[(size_t)row * P.splits + sp] = cnt;
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

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rescore_select(FvxModel M, const float* __restrict__ theta, int u0, int n_users, int splits,
                 const unsigned long long* __restrict__ cand, const int32_t* __restrict__ ccount,
                 const int64_t* __restrict__ mask_row_ptr, const int32_t* __restrict__ mask_col, int k, int npad,
                 int32_t* __restrict__ out_ids, float* __restrict__ out_scores) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* kk = reinterpret_cast<unsigned long long*>(rs_smem) + (size_t)warp * npad;
  for (int r = blockIdx.x * RS_WARPS + warp; r < n_users; r += gridDim.x * RS_WARPS) {
    const int gu = u0 + r;
    const float* urow = M.users.w + (size_t)gu * M.users.stride;
    const long long mlo = mask_row_ptr[gu], mhi = mask_row_ptr[gu + 1];
    int total = 0;
    for (int sp = 0; sp < splits; ++sp) {
      const int n = ccount[(size_t)r * splits + sp];
      const unsigned long long* src = cand + ((size_t)r * splits + sp) * TCK_CAP;
      for (int i = lane; i < n; i += 32) {
        const int32_t gid = (int32_t)(src[i] & 0xFFFFFFFFu);
        unsigned long long key = KEY_PAD;
        if (!fvx_in_sorted(mask_col, mlo, mhi, gid)) {
          const int32_t li = gid - M.item_lo;
          const float* th = M.d > 0 ? theta + (size_t)li * M.de : nullptr;
          const float s = fvx_score_one(urow, M.items.w + (size_t)li * M.items.stride, th, M.K, M.d);
          key = tck_key(s, gid);
        }
        kk[total + i] = key;
      }
      total += n;
    }
    for (int i = total + lane; i < npad; i += 32) kk[i] = KEY_PAD;
    __syncwarp();
    rs_bitonic(kk, npad, lane);
    for (int i = lane; i < k; i += 32) {
      const bool ok = kk[i] != KEY_PAD;
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

int fvx_eval_ws_query(const FvxModel* model, int32_t n_users, FvxEvalWs* ws) {
  FVX_CHECK_ARG(model && ws && n_users > 0, "fvx_eval_ws_query: bad arguments");
  const int kd1 = model->K + model->d + 1;
  const int KP = (kd1 + TCK_KB - 1) / TCK_KB * TCK_KB;
  const int n_user_tiles = (n_users + TCK_BM - 1) / TCK_BM;
  int splits = (2 * fvx_num_sms() + n_user_tiles - 1) / n_user_tiles;
  const int n_item_tiles = (model->item_cnt + TCK_BN - 1) / TCK_BN;
  if (splits > 4) splits = 4;
  if (splits > n_item_tiles) splits = n_item_tiles;
  if (splits < 1) splits = 1;
  ws->KP = KP;
  ws->splits = splits;
  ws->cap = TCK_CAP;
  ws->u_cap = n_users;
  ws->i_cap = model->item_cnt;
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
  FvxEvalWs q;
  fvx_eval_ws_query(model, n_users, &q);
  FVX_CHECK_ARG(ws->KP == q.KP && ws->splits == q.splits && ws->cap == TCK_CAP && ws->u_cap >= n_users &&
                ws->i_cap >= model->item_cnt, "fvx_score_topk_tc: workspace does not match fvx_eval_ws_query");
  FVX_CHECK_ARG(ws->A && ws->Bm && ws->unorm && ws->bmax && ws->cand && ws->ccount && ws->flags,
                "fvx_score_topk_tc: null workspace buffer");
  const int KP = q.KP, nkb = KP / TCK_KB;
  FVX_CHECK_ARG(KP <= 128, "fvx_score_topk_tc: K+d+1=%d too large for the tensor-core sweep (use fvx_score_topk)",
                model->K + model->d + 1);
  cudaStream_t st = fvx_cu(stream);

  cudaMemsetAsync(ws->bmax, 0, 4, st);
  cudaMemsetAsync(ws->flags, 0, sizeof(int32_t) * n_users, st);
  int g = (n_users * 32 + 255) / 256;
  if (g > fvx_num_sms() * 8) g = fvx_num_sms() * 8;
  k_pack_users<<<g, 256, 0, st>>>(*model, u0, u1, reinterpret_cast<__nv_bfloat16*>(ws->A), ws->unorm, KP);
  g = fvx_num_sms() * 8;
  k_pack_items<<<g, 256, 0, st>>>(*model, theta_ext, reinterpret_cast<__nv_bfloat16*>(ws->Bm),
                                  reinterpret_cast<uint32_t*>(ws->bmax), KP);
  FVX_CHECK_LAUNCH("k_pack");

  CUtensorMap tmA, tmB;
  int rc = tc_make_tensor_map_bf16(&tmA, ws->A, n_users, KP, (uint64_t)KP * 2, TCK_KB, TCK_BM, 2);
  if (rc == 0) rc = tc_make_tensor_map_bf16(&tmB, ws->Bm, model->item_cnt, KP, (uint64_t)KP * 2, TCK_KB, TCK_BN, 2);
  if (rc != 0) FVX_FAIL(-4, "fvx_score_topk_tc: cuTensorMapEncodeTiled failed (%d)", rc);

  TckParams P;
  P.n_users = n_users; P.item_cnt = model->item_cnt; P.item_lo = model->item_lo; P.nkb = nkb;
  P.n_user_tiles = (n_users + TCK_BM - 1) / TCK_BM;
  P.n_item_tiles = (model->item_cnt + TCK_BN - 1) / TCK_BN;
  P.splits = q.splits;
  P.tiles_per_split = (P.n_item_tiles + P.splits - 1) / P.splits;
  P.k = k; P.u0 = u0; P.unorm = ws->unorm; P.bmax_bits = reinterpret_cast<const uint32_t*>(ws->bmax);
  P.mask_row_ptr = mask_row_ptr; P.mask_col = mask_col;
  P.cand = reinterpret_cast<unsigned long long*>(ws->cand); P.ccount = ws->ccount; P.flags = ws->flags;
  const size_t a_bytes = (size_t)nkb * TCK_BM * 64, b_bytes = (size_t)nkb * TCK_BN * 64;
  int stages = (int)((220 * 1024 - a_bytes) / b_bytes);
  if (stages > 4) stages = 4;
  FVX_CHECK_ARG(stages >= 2, "fvx_score_topk_tc: tile does not fit shared memory");
  P.stages = stages;
  const size_t smem = a_bytes + stages * b_bytes + (2 * stages + 6) * 8 + 16 + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_topk_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) FVX_FAIL(-3, "fvx_score_topk_tc: cannot set %zu B smem: %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  int grid = P.n_user_tiles * P.splits;
  if (grid > fvx_num_sms()) grid = fvx_num_sms();
  k_topk_tc<<<grid, TCK_THREADS, smem, st>>>(tmA, tmB, P);
  FVX_CHECK_LAUNCH("k_topk_tc");

  int npad = 32;
  while (npad < q.splits * TCK_CAP) npad <<= 1;
  const size_t rs_smem = (size_t)RS_WARPS * npad * 8;
  static bool rs_conf = false;
  if (rs_smem > 48 * 1024 && !rs_conf) {
    cudaFuncSetAttribute(k_rescore_select, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    rs_conf = true;
  }
  long long rg = ((long long)n_users + RS_WARPS - 1) / RS_WARPS;
  if (rg > (long long)fvx_num_sms() * 8) rg = (long long)fvx_num_sms() * 8;
  k_rescore_select<<<(int)rg, RS_WARPS * 32, rs_smem, st>>>(*model, theta_ext, u0, n_users, q.splits, P.cand,
                                                            P.ccount, mask_row_ptr, mask_col, k, npad, out_ids,
                                                            out_scores);
  FVX_CHECK_LAUNCH("k_rescore_select");
  return 0;
}

}  // extern "C"
