// Rank counts of held-out items over the full catalog (Evaluator.py:96-98: position = number of
// candidates that score >= the held-out item) as a register-tiled fp32 sweep on the CUDA cores.
//
// Evaluator.eval needs, per user and held-out item, ONLY that count - AUC, HR@k, nDCG@k, precision and
// recall all follow from it (recommender/Evaluator.py of this package) - so the sweep keeps no lists:
// a CTA holds 128 users x 128 items of scores in registers (8 x 8 per thread), compares each against
// the user's thresholds and adds up.  The score is the SAME fmaf chain as fvx_score_one (K latent
// terms, d visual terms, item bias, visual bias - every k step of an accumulator in index order, no
// split along k), so the counts equal those of fvx_score_topk bit for bit; the tile just amortises the
// operand loads that kernel issues per FMA (8 users per item row there, 4 FMA per shared-memory load;
// here 16 FMA per 16-byte load).
#include <math_constants.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"

#define RC_TU 128            // users per CTA tile
#define RC_TI 128            // items per CTA tile
#define RC_THREADS 256       // 16 (user dim) x 16 (item dim); a thread owns users tu + 16u, items ti + 16j
#define RC_KQB 24            // k quads (4 columns) per shared-memory block: 96 columns
#define RC_PITCH (RC_TU + 1) // float4 pitch of a k-quad row: quad q of row r at q * 129 + r - conflict-free for
                             // the fill (consecutive q) and for the compute reads (consecutive r)

// columns [k0, k0+4) of the concatenated operand [latent(0..K) | vis(0..d) | 0 ...]
__device__ __forceinline__ float4 rc_quad(const float* __restrict__ latent, const float* __restrict__ vis, int K, int d,
                                          int k0, bool aligned) {
  if (aligned) {                                   // K % 4 == 0 and both pointers 16-byte aligned
    if (k0 + 4 <= K) return *reinterpret_cast<const float4*>(latent + k0);
    const int n0 = k0 - K;
    if (n0 + 4 <= d) return *reinterpret_cast<const float4*>(vis + n0);
  }
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = k0 + e;
    v[e] = k < K ? latent[k] : (k < K + d ? vis[k - K] : 0.0f);
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

template <int NT>
__global__ void __launch_bounds__(RC_THREADS, 2)
k_rank_counts(FvxModel M, const float* __restrict__ theta, int u0, int u1, int n_thr,
              const float* __restrict__ thr_scores, int32_t* __restrict__ out_counts, int n_slices, int tiles_per_slice) {
  extern __shared__ __align__(16) unsigned char rc_smem[];
  float4* As = reinterpret_cast<float4*>(rc_smem);                 // [RC_KQB][RC_PITCH]
  float4* Bs = As + RC_KQB * RC_PITCH;                             // [RC_KQB][RC_PITCH]
  float* bias = reinterpret_cast<float*>(Bs + RC_KQB * RC_PITCH);  // [2][RC_TI]: item bias (-inf: no such item), visual bias
  float* thr_s = bias + 2 * RC_TI;                                 // [RC_TU][NT] thresholds of the unit's users (registers
                                                                   // are what limits the main loop: see the b prefetch)
  const int K = M.K, d = M.d, de = M.de, Su = M.users.stride, Si = M.items.stride;
  const int KQ = (K + d + 3) >> 2;
  const int nkb = (KQ + RC_KQB - 1) / RC_KQB;
  const bool aligned = (K & 3) == 0;
  const bool vis = d > 0;
  const int tid = threadIdx.x, ti = tid & 15, tu = tid >> 4;
  const int n_ublocks = (u1 - u0 + RC_TU - 1) / RC_TU;
  const int n_units = n_ublocks * n_slices;
  const int n_tiles = (M.item_cnt + RC_TI - 1) / RC_TI;

  for (int w = blockIdx.x; w < n_units; w += gridDim.x) {
    const int ub = w / n_slices, sl = w - ub * n_slices;
    const int ubase = u0 + ub * RC_TU;
    int cnt[8][NT];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int t = 0; t < NT; ++t) cnt[u][t] = 0;
    __syncthreads();                       // the previous unit has read its thresholds
    for (int i = tid; i < RC_TU * NT; i += RC_THREADS) {
      const int r = i / NT, t = i - r * NT;
      const int gu = ubase + r;
      thr_s[i] = (gu < u1 && t < n_thr) ? thr_scores[(size_t)(gu - u0) * n_thr + t] : CUDART_NAN_F;
    }
    auto fill_users = [&](int kb) {
      const int q0 = kb * RC_KQB, nq = (KQ - q0 < RC_KQB) ? (KQ - q0) : RC_KQB;
      for (int idx = tid; idx < RC_TU * nq; idx += RC_THREADS) {
        const int r = idx / nq, q = idx - r * nq;
        const int gu = ubase + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gu < u1) {
          const float* row = M.users.w + (size_t)gu * Su;
          v = rc_quad(row, row + K, K, d, 4 * (q0 + q), aligned);
        }
        As[q * RC_PITCH + r] = v;
      }
    };
    if (nkb == 1) fill_users(0);           // the user tile stays for the whole unit

    const int t_begin = sl * tiles_per_slice;
    int t_end = t_begin + tiles_per_slice;
    if (t_end > n_tiles) t_end = n_tiles;
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int ibase = tile * RC_TI;
      float acc[8][8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[u][j] = 0.0f;
      for (int kb = 0; kb < nkb; ++kb) {
        const int q0 = kb * RC_KQB, nq = (KQ - q0 < RC_KQB) ? (KQ - q0) : RC_KQB;
        __syncthreads();                   // the previous block / tile has been consumed
        if (nkb > 1) fill_users(kb);
        for (int idx = tid; idx < RC_TI * nq; idx += RC_THREADS) {
          const int r = idx / nq, q = idx - r * nq;
          const int li = ibase + r;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (li < M.item_cnt) {
            const float* row = M.items.w + (size_t)li * Si;
            v = rc_quad(row, vis ? theta + (size_t)li * de : row, K, d, 4 * (q0 + q), aligned);
          }
          Bs[q * RC_PITCH + r] = v;
        }
        if (kb == 0 && tid < RC_TI) {
          const int li = ibase + tid;
          const bool ok = li < M.item_cnt;
          bias[tid] = ok ? M.items.w[(size_t)li * Si + K] : -CUDART_INF_F;
          bias[RC_TI + tid] = (ok && vis) ? theta[(size_t)li * de + d] : 0.0f;
        }
        __syncthreads();
        for (int q = 0; q < nq; ++q) {
          // Registers bound this loop (64 accumulators at two CTAs per SM): the user operand is held four
          // users at a time, which leaves room to fetch the item operand of column block j + 1 BEFORE the
          // FMAs of block j issue - with one register quad for b the first FMA after every LDS.128 waited
          // out the shared-memory latency (27 % of the stall samples, profiles/r1_v8_rank_counts_full.txt).
#pragma unroll
          for (int uh = 0; uh < 2; ++uh) {
            float4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = As[q * RC_PITCH + tu + 16 * (4 * uh + u)];
            float4 b_next = Bs[q * RC_PITCH + ti];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = b_next;
              if (j + 1 < 8) b_next = Bs[q * RC_PITCH + ti + 16 * (j + 1)];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float s = acc[4 * uh + u][j];
                s = fmaf(a[u].x, b.x, s);
                s = fmaf(a[u].y, b.y, s);
                s = fmaf(a[u].z, b.z, s);
                s = fmaf(a[u].w, b.w, s);
                acc[4 * uh + u][j] = s;
              }
            }
          }
        }
      }
      // biases in fvx_score_one's order, then one compare per (score, threshold)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b1 = bias[ti + 16 * j], b2 = bias[RC_TI + ti + 16 * j];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float s = acc[u][j] + b1;
          if (vis) s += b2;
#pragma unroll
          for (int t = 0; t < NT; ++t) cnt[u][t] += (s >= thr_s[(tu + 16 * u) * NT + t]) ? 1 : 0;
        }
      }
    }
    // the 16 item-lanes of a user add up; one atomic per (user, threshold, unit)
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        int v = cnt[u][t];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int gu = ubase + tu + 16 * u;
        if (ti == 0 && gu < u1 && t < n_thr && v) atomicAdd(out_counts + (size_t)(gu - u0) * n_thr + t, v);
      }
    __syncthreads();                       // As is refilled by the next unit
  }
}

// counts -= the user's masked (train) items that reach the threshold: they are not candidates
// (Evaluator.py:40: all items minus the training items).  One warp per user.
__global__ void __launch_bounds__(256)
k_rank_unmask(FvxModel M, const float* __restrict__ theta, int u0, int u1, const int64_t* __restrict__ mask_row_ptr,
              const int32_t* __restrict__ mask_col, int n_thr, const float* __restrict__ thr_scores,
              int32_t* __restrict__ out_counts) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int u = u0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); u < u1; u += warps) {
    const float* urow = M.users.w + (size_t)u * M.users.stride;
    float thr[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) thr[t] = t < n_thr ? thr_scores[(size_t)(u - u0) * n_thr + t] : CUDART_NAN_F;
    int sub[4] = {0, 0, 0, 0};
    for (int64_t e = mask_row_ptr[u] + lane; e < mask_row_ptr[u + 1]; e += 32) {
      const int32_t li = mask_col[e] - M.item_lo;
      if (li < 0 || li >= M.item_cnt) continue;
      const float* th = M.d > 0 ? theta + (size_t)li * M.de : nullptr;
      const float s = fvx_score_one(urow, M.items.w + (size_t)li * M.items.stride, th, M.K, M.d);
#pragma unroll
      for (int t = 0; t < 4; ++t) sub[t] += (s >= thr[t]) ? 1 : 0;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int v = sub[t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && t < n_thr && v) atomicSub(out_counts + (size_t)(u - u0) * n_thr + t, v);
    }
  }
}

extern "C" int fvx_rank_counts(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                               const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t n_thr,
                               const float* thr_scores, int32_t* out_counts, fvx_stream_t stream) {
  FVX_CHECK_ARG(model && model->abi_version == FVX_ABI_VERSION, "fvx_rank_counts: bad model");
  FVX_CHECK_ARG(model->users.w && model->items.w, "fvx_rank_counts: null tables");
  FVX_CHECK_ARG(model->d == 0 || theta_ext != nullptr, "fvx_rank_counts: VBPR scoring needs theta_ext (fvx_project)");
  FVX_CHECK_ARG(0 <= u0 && u0 <= u1 && u1 <= model->num_users, "fvx_rank_counts: bad user range");
  FVX_CHECK_ARG(n_thr >= 1 && n_thr <= 4, "fvx_rank_counts: n_thr=%d outside [1,4]", n_thr);
  FVX_CHECK_ARG(mask_row_ptr && mask_col && thr_scores && out_counts, "fvx_rank_counts: null pointer");
  if (u1 == u0) return 0;
  cudaStream_t st = fvx_cu(stream);
  const int n = u1 - u0;
  cudaMemsetAsync(out_counts, 0, (size_t)n * n_thr * sizeof(int32_t), st);
  const int n_ublocks = (n + RC_TU - 1) / RC_TU;
  const int n_tiles = (model->item_cnt + RC_TI - 1) / RC_TI;
  // item slices: enough units for ~8 rounds of a full grid (2 CTAs per SM), at least 4 tiles per slice
  const int slots = fvx_num_sms() * 2;
  int n_slices = (8 * slots + n_ublocks - 1) / n_ublocks;
  if (n_slices > (n_tiles + 3) / 4) n_slices = (n_tiles + 3) / 4;
  if (n_slices < 1) n_slices = 1;
  const int tiles_per_slice = (n_tiles + n_slices - 1) / n_slices;
  n_slices = (n_tiles + tiles_per_slice - 1) / tiles_per_slice;
  const size_t smem = (size_t)2 * RC_KQB * RC_PITCH * sizeof(float4) + 2 * RC_TI * sizeof(float) +
                      (size_t)RC_TU * 4 * sizeof(float);
  static FvxSmemMark rc2_smem, rc4_smem;
  if (int r = fvx_ensure_smem((const void*)k_rank_counts<2>, &rc2_smem, smem, "fvx_rank_counts")) return r;
  if (int r = fvx_ensure_smem((const void*)k_rank_counts<4>, &rc4_smem, smem, "fvx_rank_counts")) return r;
  long long grid = (long long)n_ublocks * n_slices;
  if (grid > slots) grid = slots;
  if (n_thr <= 2)
    k_rank_counts<2><<<(int)grid, RC_THREADS, smem, st>>>(*model, theta_ext, u0, u1, n_thr, thr_scores, out_counts,
                                                           n_slices, tiles_per_slice);
  else
    k_rank_counts<4><<<(int)grid, RC_THREADS, smem, st>>>(*model, theta_ext, u0, u1, n_thr, thr_scores, out_counts,
                                                           n_slices, tiles_per_slice);
  FVX_CHECK_LAUNCH("k_rank_counts");
  long long g2 = ((long long)n * 32 + 255) / 256;
  if (g2 > (long long)fvx_num_sms() * 8) g2 = (long long)fvx_num_sms() * 8;
  k_rank_unmask<<<(int)g2, 256, 0, st>>>(*model, theta_ext, u0, u1, mask_row_ptr, mask_col, n_thr, thr_scores, out_counts);
  FVX_CHECK_LAUNCH("k_rank_unmask");
  return 0;
}
