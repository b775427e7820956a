// Full-catalog evaluation, exact fp32 CUDA-core version: dense predict_all (small
// sizes), explicit pair scores, and the fused score + train-mask + top-k sweep that
// never materialises the U x I matrix.  Replaces predict_all (BPRMF.py:85,
// VBPR.py:95-97) and the host loops of Evaluator.store_recommendation
// (Evaluator.py:231-237) and _eval_by_user's rank count (Evaluator.py:96-98).
#include <math_constants.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"

static FvxSmemMark g_topk_smem;   // k_score_topk: dynamic shared memory configured per device

// ---- order-preserving key: ascending key == descending score, then ascending id ----
__device__ __forceinline__ unsigned long long topk_key(float s, int32_t id) {
  const uint32_t b = __float_as_uint(s);
  const uint32_t mono = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)(~mono) << 32) | (uint32_t)id;
}
__device__ __forceinline__ float key_score(unsigned long long k) {
  const uint32_t mono = ~(uint32_t)(k >> 32);
  const uint32_t b = (mono & 0x80000000u) ? (mono ^ 0x80000000u) : ~mono;
  return __uint_as_float(b);
}
#define KEY_PAD 0xFFFFFFFFFFFFFFFFull

// bitonic sort (ascending) of n (power of two) keys in shared memory by one warp
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* keys, int n, int lane) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = lane; i < (n >> 1); i += 32) {
        const int l = 2 * i - (i & (stride - 1));
        const int r = l + stride;
        const bool up = (l & size) == 0;
        const unsigned long long a = keys[l], b = keys[r];
        if ((a > b) == up) { keys[l] = b; keys[r] = a; }
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------
__global__ void k_predict_all(FvxModel M, const float* __restrict__ theta, int u0, int u1, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.item_cnt) return;
  const float* irow = M.items.w + (size_t)i * M.items.stride;
  const float* th = M.d > 0 ? theta + (size_t)i * M.de : nullptr;
  for (int u = u0 + blockIdx.y; u < u1; u += gridDim.y) {
    const float* urow = M.users.w + (size_t)u * M.users.stride;
    out[(size_t)(u - u0) * M.item_cnt + i] = fvx_score_one(urow, irow, th, M.K, M.d);
  }
}

__global__ void k_score_pairs(FvxModel M, const float* __restrict__ theta, const int32_t* __restrict__ user,
                              const int32_t* __restrict__ item, long long n, float* __restrict__ out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x) {
    const int32_t li = item[t] - M.item_lo;
    float s = 0.0f;
    if (li >= 0 && li < M.item_cnt && user[t] >= 0 && user[t] < M.num_users) {
      const float* th = M.d > 0 ? theta + (size_t)li * M.de : nullptr;
      s = fvx_score_one(M.users.w + (size_t)user[t] * M.users.stride, M.items.w + (size_t)li * M.items.stride,
                        th, M.K, M.d);
    }
    out[t] = s;
  }
}

// ---------------------------------------------------------------------------------
// Sweep: a CTA owns TK_UB users; each thread scores one item of the current
// 256-item tile against all of them (user rows in shared memory), survivors of the
// per-user running threshold are appended to that user's candidate buffer, which a
// warp compacts to the best k (bitonic sort) whenever it is more than half full.
#define TK_UB 8
#define TK_TILE 256
#define TK_CAP 512
#define TK_MAXTHR 4

__global__ void __launch_bounds__(TK_TILE)
k_score_topk(FvxModel M, const float* __restrict__ theta, int u0, int u1,
             const int64_t* __restrict__ mask_row_ptr, const int32_t* __restrict__ mask_col, int k,
             int32_t* __restrict__ out_ids, float* __restrict__ out_scores, int n_thr,
             const float* __restrict__ thr_scores, int32_t* __restrict__ out_counts,
             const int32_t* __restrict__ ulist,     // ulist: user of local index j (nullptr: j itself)
             const int32_t* __restrict__ n_dev,     // list mode: number of valid list entries lives on the device
             int scatter_base,                      // >= 0: output row of list entry j is ulist[j] - scatter_base
             const int32_t* __restrict__ tau_enc) { // list mode, optional: a proven lower bound of the k-th best unmasked
                                                    // score of row ulist[j] - scatter_base (int32 whose signed order is the
                                                    // float's, fvx_score_topk_tc_bounds): the row's threshold starts there
  extern __shared__ __align__(16) unsigned char tk_smem[];
  const int Su = M.users.stride, Si = M.items.stride, K = M.K, d = M.d, de = M.de;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(tk_smem);          // [UB][CAP]
  float* us = reinterpret_cast<float*>(keys + TK_UB * TK_CAP);                         // [UB][Su]
  float* tau = us + TK_UB * Su;                                                         // [UB]
  float* thr = tau + TK_UB;                                                             // [UB][MAXTHR]
  int* cnt = reinterpret_cast<int*>(thr + TK_UB * TK_MAXTHR);                           // [UB]
  int* cge = cnt + TK_UB;                                                               // [UB][MAXTHR]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // device-counted lists (the rows fvx_score_topk_tc flags) are short: one user per CTA pass, so
  // that a handful of users spread over the grid instead of queueing in one CTA
  const bool spread = n_dev != nullptr;
  if (spread) { const int nd = *n_dev; u1 = nd < u1 ? nd : u1; }
  const int ub_step = spread ? 1 : TK_UB;
  for (int ub = u0 + blockIdx.x * ub_step; ub < u1; ub += gridDim.x * ub_step) {
    const int nu = spread ? 1 : ((u1 - ub < TK_UB) ? (u1 - ub) : TK_UB);
    __syncthreads();
    for (int e = tid; e < TK_UB * Su; e += TK_TILE) {
      const int u = e / Su, c = e - u * Su;
      us[e] = (u < nu) ? M.users.w[(size_t)(ulist ? ulist[ub + u] : ub + u) * Su + c] : 0.0f;
    }
    if (tid < TK_UB) {
      float t0 = -CUDART_INF_F;
      if (tau_enc && scatter_base >= 0 && tid < nu) {
        // without it every item passes until the first compaction and the list is sorted once per 256-item tile:
        // 4 ms per flagged row at 100 k items; the bound cuts the insertions to the few hundred items that reach it
        const int32_t e = tau_enc[ulist[ub + tid] - scatter_base];
        const float a = __int_as_float(e >= 0 ? e : (e ^ 0x7FFFFFFF));
        if (a > -CUDART_INF_F && a < CUDART_INF_F) t0 = nextafterf(a, -CUDART_INF_F);   // items with s >= a pass (s > tau)
      }
      tau[tid] = t0;
      cnt[tid] = 0;
    }
    if (tid < TK_UB * TK_MAXTHR) {
      const int u = tid / TK_MAXTHR, t = tid - u * TK_MAXTHR;
      thr[tid] = (u < nu && t < n_thr) ? thr_scores[(size_t)(ub + u - u0) * n_thr + t] : CUDART_NAN_F;
      cge[tid] = 0;
    }
    __syncthreads();
    int cl[TK_UB][TK_MAXTHR];
#pragma unroll
    for (int u = 0; u < TK_UB; ++u)
#pragma unroll
      for (int t = 0; t < TK_MAXTHR; ++t) cl[u][t] = 0;

    for (int base = 0; base < M.item_cnt; base += TK_TILE) {
      const int li = base + tid;
      if (li < M.item_cnt) {
        const float4* irow = reinterpret_cast<const float4*>(M.items.w + (size_t)li * Si);
        float acc[TK_UB];
#pragma unroll
        for (int u = 0; u < TK_UB; ++u) acc[u] = 0.0f;
        // K latent terms (4 at a time; K % 4 tail handled scalar), in index order
        int c = 0;
        for (; c + 4 <= K; c += 4) {
          const float4 x = irow[c >> 2];
#pragma unroll
          for (int u = 0; u < TK_UB; ++u) {
            const float4 a = *reinterpret_cast<const float4*>(us + u * Su + c);
            acc[u] = fmaf(a.x, x.x, acc[u]);
            acc[u] = fmaf(a.y, x.y, acc[u]);
            acc[u] = fmaf(a.z, x.z, acc[u]);
            acc[u] = fmaf(a.w, x.w, acc[u]);
          }
        }
        const float* irs = reinterpret_cast<const float*>(irow);
        for (; c < K; ++c) {
          const float x = irs[c];
#pragma unroll
          for (int u = 0; u < TK_UB; ++u) acc[u] = fmaf(us[u * Su + c], x, acc[u]);
        }
        const float beta = irs[K];
        if (d > 0) {
          const float* th = theta + (size_t)li * de;
          for (int n = 0; n < d; ++n) {
            const float x = th[n];
#pragma unroll
            for (int u = 0; u < TK_UB; ++u) acc[u] = fmaf(us[u * Su + K + n], x, acc[u]);
          }
          const float vb = th[d];
#pragma unroll
          for (int u = 0; u < TK_UB; ++u) { acc[u] += beta; acc[u] += vb; }
        } else {
#pragma unroll
          for (int u = 0; u < TK_UB; ++u) acc[u] += beta;
        }
        const int32_t gid = li + M.item_lo;
#pragma unroll
        for (int u = 0; u < TK_UB; ++u) {
          if (u >= nu) break;
          const float s = acc[u];
#pragma unroll
          for (int t = 0; t < TK_MAXTHR; ++t) cl[u][t] += (s >= thr[u * TK_MAXTHR + t]) ? 1 : 0;
          if (s > tau[u]) {
            const int gu = ulist ? ulist[ub + u] : ub + u;
            if (!fvx_in_sorted(mask_col, mask_row_ptr[gu], mask_row_ptr[gu + 1], gid)) {
              const int p = atomicAdd(&cnt[u], 1);
              if (p < TK_CAP) keys[u * TK_CAP + p] = topk_key(s, gid);
            }
          }
        }
      }
      __syncthreads();
      if (warp < nu && cnt[warp] > TK_CAP / 2) {
        unsigned long long* kk = keys + warp * TK_CAP;
        const int c0 = cnt[warp] < TK_CAP ? cnt[warp] : TK_CAP;
        for (int i = c0 + lane; i < TK_CAP; i += 32) kk[i] = KEY_PAD;
        __syncwarp();
        warp_bitonic_sort(kk, TK_CAP, lane);
        if (lane == 0) {
          const int keep = c0 < k ? c0 : k;
          cnt[warp] = keep;
          if (keep == k) tau[warp] = key_score(kk[k - 1]);
        }
      }
      __syncthreads();
    }
    // final ordering + output; masked items were never inserted
    if (warp < nu) {
      unsigned long long* kk = keys + warp * TK_CAP;
      const int c0 = cnt[warp] < TK_CAP ? cnt[warp] : TK_CAP;
      for (int i = c0 + lane; i < TK_CAP; i += 32) kk[i] = KEY_PAD;
      __syncwarp();
      warp_bitonic_sort(kk, TK_CAP, lane);
      const size_t o = (size_t)(scatter_base >= 0 ? ulist[ub + warp] - scatter_base : ub + warp - u0) * k;
      for (int i = lane; i < k; i += 32) {
        const bool ok = i < c0;
        out_ids[o + i] = ok ? (int32_t)(kk[i] & 0xFFFFFFFFu) : -1;
        out_scores[o + i] = ok ? key_score(kk[i]) : -CUDART_INF_F;
      }
    }
    // rank counts: all owned items >= threshold, minus the masked (train) ones
    if (n_thr > 0) {
#pragma unroll
      for (int u = 0; u < TK_UB; ++u)
#pragma unroll
        for (int t = 0; t < TK_MAXTHR; ++t) {
          int v = cl[u][t];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && v) atomicAdd(&cge[u * TK_MAXTHR + t], v);
        }
      __syncthreads();
      if (warp < nu) {
        const int gu = ulist ? ulist[ub + warp] : ub + warp;
        const int64_t a = mask_row_ptr[gu], b = mask_row_ptr[gu + 1];
        int sub[TK_MAXTHR] = {0, 0, 0, 0};
        for (int64_t e = a + lane; e < b; e += 32) {
          const int32_t li = mask_col[e] - M.item_lo;
          if (li < 0 || li >= M.item_cnt) continue;
          const float* th = d > 0 ? theta + (size_t)li * de : nullptr;
          const float s = fvx_score_one(us + warp * Su, M.items.w + (size_t)li * Si, th, K, d);
#pragma unroll
          for (int t = 0; t < TK_MAXTHR; ++t) sub[t] += (s >= thr[warp * TK_MAXTHR + t]) ? 1 : 0;
        }
#pragma unroll
        for (int t = 0; t < TK_MAXTHR; ++t) {
          int v = sub[t];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && t < n_thr) out_counts[(size_t)(ub + warp - u0) * n_thr + t] = cge[warp * TK_MAXTHR + t] - v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// Merge of R sorted per-shard lists: one warp sorts the R*k keys of a user.
#define MG_WARPS 4
__global__ void __launch_bounds__(MG_WARPS * 32)
k_topk_merge(const int32_t* __restrict__ ids, const float* __restrict__ scores, long long n_users, int R, int k,
             int npad, int32_t* __restrict__ out_ids, float* __restrict__ out_scores) {
  extern __shared__ __align__(16) unsigned char mg_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* kk = reinterpret_cast<unsigned long long*>(mg_smem) + (size_t)warp * npad;
  for (long long u = (long long)blockIdx.x * MG_WARPS + warp; u < n_users; u += (long long)gridDim.x * MG_WARPS) {
    const size_t base = (size_t)u * R * k;
    for (int i = lane; i < npad; i += 32) {
      unsigned long long key = KEY_PAD;
      if (i < R * k && ids[base + i] >= 0) key = topk_key(scores[base + i], ids[base + i]);
      kk[i] = key;
    }
    __syncwarp();
    warp_bitonic_sort(kk, npad, lane);
    for (int i = lane; i < k; i += 32) {
      const bool ok = kk[i] != KEY_PAD;
      out_ids[(size_t)u * k + i] = ok ? (int32_t)(kk[i] & 0xFFFFFFFFu) : -1;
      out_scores[(size_t)u * k + i] = ok ? key_score(kk[i]) : -CUDART_INF_F;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------
static int check_eval_model(const FvxModel* m, const float* theta, const char* who) {
  FVX_CHECK_ARG(m && m->abi_version == FVX_ABI_VERSION, "%s: bad model", who);
  FVX_CHECK_ARG(m->users.w && m->items.w, "%s: null tables", who);
  FVX_CHECK_ARG(m->d == 0 || theta != nullptr, "%s: VBPR scoring needs theta_ext (fvx_project)", who);
  return 0;
}

extern "C" {

int fvx_predict_all(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1, float* out,
                    fvx_stream_t stream) {
  if (int rc = check_eval_model(model, theta_ext, "fvx_predict_all")) return rc;
  FVX_CHECK_ARG(0 <= u0 && u0 <= u1 && u1 <= model->num_users && out, "fvx_predict_all: bad user range");
  if (u1 == u0) return 0;
  int gy = u1 - u0;
  if (gy > 4096) gy = 4096;
  dim3 grid((model->item_cnt + 255) / 256, gy);
  k_predict_all<<<grid, 256, 0, fvx_cu(stream)>>>(*model, theta_ext, u0, u1, out);
  FVX_CHECK_LAUNCH("k_predict_all");
  return 0;
}

int fvx_score_pairs(const FvxModel* model, const float* theta_ext, const int32_t* user, const int32_t* item,
                    int64_t n, float* out, fvx_stream_t stream) {
  if (int rc = check_eval_model(model, theta_ext, "fvx_score_pairs")) return rc;
  FVX_CHECK_ARG(user && item && out, "fvx_score_pairs: null pointer");
  if (n <= 0) return 0;
  long long g = (n + 255) / 256;
  if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
  k_score_pairs<<<(int)g, 256, 0, fvx_cu(stream)>>>(*model, theta_ext, user, item, n, out);
  FVX_CHECK_LAUNCH("k_score_pairs");
  return 0;
}

static int score_topk_impl(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                           const int32_t* ulist, const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k,
                           int32_t* out_ids, float* out_scores, int32_t n_thr, const float* thr_scores,
                           int32_t* out_counts, fvx_stream_t stream) {
  if (int rc = check_eval_model(model, theta_ext, "fvx_score_topk")) return rc;
  FVX_CHECK_ARG(0 <= u0 && u0 <= u1 && (ulist != nullptr || u1 <= model->num_users), "fvx_score_topk: bad user range");
  FVX_CHECK_ARG(k >= 1 && k <= 128, "fvx_score_topk: k=%d outside [1,128]", k);
  FVX_CHECK_ARG(mask_row_ptr && mask_col && out_ids && out_scores, "fvx_score_topk: null pointer");
  FVX_CHECK_ARG(n_thr >= 0 && n_thr <= TK_MAXTHR, "fvx_score_topk: n_thr=%d outside [0,%d]", n_thr, TK_MAXTHR);
  FVX_CHECK_ARG(n_thr == 0 || (thr_scores && out_counts), "fvx_score_topk: thresholds need thr_scores/out_counts");
  if (u1 == u0) return 0;
  const size_t smem = (size_t)TK_UB * TK_CAP * 8 + (size_t)TK_UB * model->users.stride * 4 +
                      (size_t)TK_UB * 4 * (2 + 2 * TK_MAXTHR);
  if (int r = fvx_ensure_smem((const void*)k_score_topk, &g_topk_smem, smem, "fvx_score_topk")) return r;
  long long g = ((long long)(u1 - u0) + TK_UB - 1) / TK_UB;
  if (g > (long long)fvx_num_sms() * 4) g = (long long)fvx_num_sms() * 4;
  k_score_topk<<<(int)g, TK_TILE, smem, fvx_cu(stream)>>>(*model, theta_ext, u0, u1, mask_row_ptr, mask_col, k,
                                                          out_ids, out_scores, n_thr, thr_scores, out_counts, ulist,
                                                          nullptr, -1, nullptr);
  FVX_CHECK_LAUNCH("k_score_topk");
  return 0;
}

}  // extern "C"

// users u0 + r with flags[r] != 0 -> list (order arbitrary), *count = its length
__global__ void k_flag_list(const int32_t* __restrict__ flags, int n, int u0, int32_t* __restrict__ list,
                            int32_t* __restrict__ count) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    if (flags[r] != 0) list[atomicAdd(count, 1)] = u0 + r;
}

// The exact sweep for the rows of [u0, u0+n) that carry a flag; results are written into the rows'
// own slots of out_ids / out_scores.  No host synchronisation: the list and its length stay on the device.
int fvx_launch_topk_flagged(const FvxModel* model, const float* theta_ext, const int32_t* flags, int n, int u0,
                            const int64_t* mask_row_ptr, const int32_t* mask_col, int k, int32_t* out_ids,
                            float* out_scores, int32_t* list_scratch, int32_t* count_scratch, cudaStream_t st,
                            const int32_t* tau_enc) {
  cudaMemsetAsync(count_scratch, 0, sizeof(int32_t), st);
  int g = (n + 255) / 256;
  if (g > fvx_num_sms() * 4) g = fvx_num_sms() * 4;
  k_flag_list<<<g, 256, 0, st>>>(flags, n, u0, list_scratch, count_scratch);
  const size_t smem = (size_t)TK_UB * TK_CAP * 8 + (size_t)TK_UB * model->users.stride * 4 +
                      (size_t)TK_UB * 4 * (2 + 2 * TK_MAXTHR);
  if (int r = fvx_ensure_smem((const void*)k_score_topk, &g_topk_smem, smem, "fvx_score_topk")) return r;
  k_score_topk<<<fvx_num_sms() * 2, TK_TILE, smem, st>>>(*model, theta_ext, 0, n, mask_row_ptr, mask_col, k, out_ids,
                                                         out_scores, 0, nullptr, nullptr, list_scratch, count_scratch,
                                                         u0, tau_enc);
  FVX_CHECK_LAUNCH("k_score_topk (flagged rows)");
  return 0;
}

extern "C" {

int fvx_score_topk(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                   const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k, int32_t* out_ids,
                   float* out_scores, int32_t n_thr, const float* thr_scores, int32_t* out_counts,
                   fvx_stream_t stream) {
  return score_topk_impl(model, theta_ext, u0, u1, nullptr, mask_row_ptr, mask_col, k, out_ids, out_scores, n_thr,
                         thr_scores, out_counts, stream);
}

int fvx_score_topk_users(const FvxModel* model, const float* theta_ext, const int32_t* users, int32_t n,
                         const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k, int32_t* out_ids,
                         float* out_scores, fvx_stream_t stream) {
  FVX_CHECK_ARG(users != nullptr && n >= 0, "fvx_score_topk_users: bad user list");
  return score_topk_impl(model, theta_ext, 0, n, users, mask_row_ptr, mask_col, k, out_ids, out_scores, 0, nullptr,
                         nullptr, stream);
}

int fvx_topk_merge(const int32_t* ids, const float* scores, int64_t n_users, int32_t R, int32_t k,
                   int32_t* out_ids, float* out_scores, fvx_stream_t stream) {
  FVX_CHECK_ARG(ids && scores && out_ids && out_scores, "fvx_topk_merge: null pointer");
  FVX_CHECK_ARG(R >= 1 && k >= 1 && (long long)R * k <= 2048, "fvx_topk_merge: R*k=%lld outside [1,2048]",
                (long long)R * k);
  if (n_users <= 0) return 0;
  int npad = 32;
  while (npad < R * k) npad <<= 1;
  const size_t smem = (size_t)MG_WARPS * npad * 8;
  static FvxSmemMark merge_smem;
  if (int r = fvx_ensure_smem((const void*)k_topk_merge, &merge_smem, smem, "fvx_topk_merge")) return r;
  long long g = (n_users + MG_WARPS - 1) / MG_WARPS;
  if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
  k_topk_merge<<<(int)g, MG_WARPS * 32, smem, fvx_cu(stream)>>>(ids, scores, n_users, R, k, npad, out_ids,
                                                                out_scores);
  FVX_CHECK_LAUNCH("k_topk_merge");
  return 0;
}

}  // extern "C"
