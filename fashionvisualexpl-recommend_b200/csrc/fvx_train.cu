// The BPR optimiser step (single rank): replaces BPRMF.train_step / VBPR.train_step
// (src/recommender/models/BPRMF.py:87-125, VBPR.py:99-144) and the Keras-Adam update
// they call (math: SURVEY.md Appendix A; restated in oracle/bpr.py).
//
// Kernel sequence per step (one stream, no host synchronisation, graph-capturable):
//   k_prep        unique touched user / item rows -> lists, local row ids of the slots,
//                 DEFERRED Adam catch-up of the touched rows, bf16 planes of E_ext^T
//   projection    TH = F[rows] * E_ext                       (fvx_project[_tc].cu, VBPR only)
//   k_score_grad  x_uij, loss, gradient coefficients, scatter-add into g, W for dE
//   grad_E        gE_part = F[rows]^T * W                    (fvx_project[_tc].cu, VBPR only)
//   k_update      Adam on the touched rows (or whole tables: DENSE), dense Adam on E_ext,
//                 step += 1 and list reset by the last block
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"

// ---------------------------------------------------------------------------------
// DEFERRED Adam catch-up.  A row current to step `last` is brought to step `target`:
//   step last+1 is a FULL Adam step with the row's pending gradient g (the gradient of the step at
//     which the row was last touched stays in g until the row is needed again - the single-rank step
//     has no separate row-update kernel; g is zero when k_update already applied it, or the row was
//     never touched, and the step degenerates to a zero-gradient one);
//   steps last+2 .. target are the zero-gradient steps the reference's dense-semantics Adam takes on
//     rows without gradient:  m <- b1*m ; v <- b2*v ; w <- w - alpha_tau * m / (sqrt(v) + eps).
// The zero-gradient loop is truncated after FVX_REPLAY_MAX iterations (the remaining updates are
// below 2e-9 of the first one); m and v then take their closed-form decay.
// `nl` lanes (lane index li) cooperate on one row, 4 columns at a time.
__device__ __forceinline__ void replay_row(const FvxTable& T, int32_t r, int32_t target, float lr, int li, int nl) {
  const int32_t last = T.last[r];
  const int32_t gap = target - last;
  if (gap > 0) {
    const int nz = gap - 1;
    const int n = nz < FVX_REPLAY_MAX ? nz : FVX_REPLAY_MAX;
    const int rem = nz - n;
    // q1 = 1 - b1^tau, q2 = 1 - b2^tau for tau = last+1 (expm1f: no cancellation for small tau), then
    // q <- (1-b) + b*q.  Single precision throughout: the kernel was instruction-bound on the double
    // expm1 / sqrt / divide every lane evaluated per row (28.5 M warp instructions per launch); alpha
    // carries ~2 ulp, i.e. a relative error of 1e-7 on an update that is itself ~1e-2 of the weight.
    const float tau = (float)(last + 1);
    const float q1f = -expm1f(tau * -0.10536051565782628f);    // ln 0.9
    const float q2f = -expm1f(tau * -0.0010005003335835335f);  // ln 0.999
    const float a0 = lr * sqrtf(q2f) / q1f;                    // alpha of step last+1 (fvx_alpha)
    const float q1_0 = fmaf(FVX_BETA1, q1f, 1.0f - FVX_BETA1);
    const float q2_0 = fmaf(FVX_BETA2, q2f, 1.0f - FVX_BETA2);
    const float d1 = rem > 0 ? expf((float)rem * -0.10536051565782628f) : 1.0f;
    const float d2 = rem > 0 ? expf((float)rem * -0.0010005003335835335f) : 1.0f;
    float4* __restrict__ w = reinterpret_cast<float4*>(T.w + (size_t)r * T.stride);
    float4* __restrict__ m = reinterpret_cast<float4*>(T.m + (size_t)r * T.stride);
    float4* __restrict__ v = reinterpret_cast<float4*>(T.v + (size_t)r * T.stride);
    float4* __restrict__ g = reinterpret_cast<float4*>(T.g + (size_t)r * T.stride);
    auto advance = [&](float4& wc, float4& mc, float4& vc, const float4& gc) {
#define FVX_RP0(f)                                                        \
      mc.f = FVX_BETA1 * mc.f + (1.0f - FVX_BETA1) * gc.f;                \
      vc.f = FVX_BETA2 * vc.f + (1.0f - FVX_BETA2) * (gc.f * gc.f);       \
      wc.f -= a0 * mc.f * fvx_rcp_approx(fvx_sqrt_approx(vc.f) + FVX_EPS);
      FVX_RP0(x) FVX_RP0(y) FVX_RP0(z) FVX_RP0(w)
#undef FVX_RP0
      float q1 = q1_0, q2 = q2_0;
      for (int k = 0; k < n; ++k) {
        const float a = lr * fvx_sqrt_approx(q2) * fvx_rcp_approx(q1);
#define FVX_RP1(f)                                                        \
        mc.f *= FVX_BETA1;                                                \
        vc.f *= FVX_BETA2;                                                \
        wc.f -= a * mc.f * fvx_rcp_approx(fvx_sqrt_approx(vc.f) + FVX_EPS);
        FVX_RP1(x) FVX_RP1(y) FVX_RP1(z) FVX_RP1(w)
#undef FVX_RP1
        q1 = fmaf(FVX_BETA1, q1, 1.0f - FVX_BETA1);
        q2 = fmaf(FVX_BETA2, q2, 1.0f - FVX_BETA2);
      }
      mc.x *= d1; mc.y *= d1; mc.z *= d1; mc.w *= d1;
      vc.x *= d2; vc.y *= d2; vc.z *= d2; vc.w *= d2;
    };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int s4 = T.stride >> 2;
    // two column chunks per pass: the eight loads of a pass are in flight together (a row of K = 64 is
    // 17 - 21 float4: with 8 lanes per row two passes instead of three dependent ones)
    for (int c0 = 0; c0 < s4; c0 += 2 * nl) {
      const int ca = c0 + li, cb = c0 + nl + li;
      const bool ha = ca < s4, hb = cb < s4;
      float4 wa = z4, ma = z4, va = z4, ga = z4, wb = z4, mb = z4, vb = z4, gb = z4;
      if (ha) { wa = w[ca]; ma = m[ca]; va = v[ca]; ga = g[ca]; }
      if (hb) { wb = w[cb]; mb = m[cb]; vb = v[cb]; gb = g[cb]; }
      if (ha) {
        advance(wa, ma, va, ga);
        w[ca] = wa; m[ca] = ma; v[ca] = va;
        if (ga.x != 0.0f || ga.y != 0.0f || ga.z != 0.0f || ga.w != 0.0f) g[ca] = z4;
      }
      if (hb) {
        advance(wb, mb, vb, gb);
        w[cb] = wb; m[cb] = mb; v[cb] = vb;
        if (gb.x != 0.0f || gb.y != 0.0f || gb.z != 0.0f || gb.w != 0.0f) g[cb] = z4;
      }
    }
  }
  // every lane of the group has read T.last[r] before it is advanced
  __syncwarp(nl == 32 ? 0xffffffffu : (0xFFu << ((threadIdx.x & 31) & ~7)));
  if (li == 0 && gap > 0) T.last[r] = target;
}

// Claims row r of table T for step t.  The winning lanes of the warp append their rows to
// T.list with ONE atomic on the counter; returns the ballot of winners.
__device__ __forceinline__ uint32_t claim_rows(const FvxTable& T, int32_t r, bool want, int32_t t, int lane) {
  bool win = false;
  if (want) win = atomicMax(&T.mark[r], t) < t;
  const uint32_t b = __ballot_sync(0xffffffffu, win);
  if (b) {
    const int leader = __ffs(b) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(T.count, __popc(b));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (win) {
      const int idx = base + __popc(b & ((1u << lane) - 1u));
      if (idx < T.list_cap) T.list[idx] = r;
    }
  }
  return b;
}

// The two item rows of a triple claimed together: both atomicMax stamps in flight at once, then ONE
// warp-aggregated append for all winners of the warp.  upos[r] records where in the list row r went:
// every slot of the batch that refers to row r finds its projection / coefficient sum through it.
__device__ __forceinline__ void claim_pair_pos(const FvxTable& T, int32_t* __restrict__ upos, int32_t ri, int32_t rj,
                                               int32_t t, int lane) {
  bool wi = false, wj = false;
  // a row that already carries this step's stamp is lost without asking: the stamps of a popular row (tens of
  // thousands of slots of one batch on an item shard) would otherwise queue up as atomics on one address
  const int32_t mi = ri >= 0 ? __ldcg(&T.mark[ri]) : t, mj = rj >= 0 ? __ldcg(&T.mark[rj]) : t;
  if (mi != t) wi = atomicMax(&T.mark[ri], t) < t;
  if (mj != t) wj = atomicMax(&T.mark[rj], t) < t;      // rj == ri: the second stamp finds t and loses
  const uint32_t bi = __ballot_sync(0xffffffffu, wi), bj = __ballot_sync(0xffffffffu, wj);
  const int ni = __popc(bi), total = ni + __popc(bj);
  if (total) {
    int base = 0;
    if (lane == 0) base = atomicAdd(T.count, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t lt = (1u << lane) - 1u;
    if (wi) {
      const int idx = base + __popc(bi & lt);
      if (idx < T.list_cap) { T.list[idx] = ri; upos[ri] = idx; }
    }
    if (wj) {
      const int idx = base + ni + __popc(bj & lt);
      if (idx < T.list_cap) { T.list[idx] = rj; upos[rj] = idx; }
    }
  }
}

#define PREP_TPW 8
// blocks [0, nb_mark): PREP_TPW triples per warp pass; blocks beyond: bf16 planes of E_ext^T
// bf16 hi / lo planes of E_ext^T [NP, D] (rows >= de zero), blocks blk of nblk
__device__ __forceinline__ void split_E_planes(const FvxModel& M, int NP, int blk, int nblk) {
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(M.ET_hi);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(M.ET_lo);
  const int total = NP * M.D;
  for (int i = blk * blockDim.x + threadIdx.x; i < total; i += nblk * blockDim.x) {
    const int n = i / M.D, f = i - n * M.D;
    const float x = n < M.de ? M.E[(size_t)f * M.de + n] : 0.0f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

// What the projection needs from the batch, and nothing else: the local item row of every
// (triple, side) slot ([pos(B) | neg(B)]) and the planes of E_ext^T.  Launched on the main stream while
// k_prep (claims + deferred-Adam catch-up) runs beside the projection on the side stream.
__global__ void __launch_bounds__(256)
k_rows_et(FvxModel M, const int32_t* __restrict__ pos, const int32_t* __restrict__ neg, int B, int nb_rows, int NP) {
  if ((int)blockIdx.x >= nb_rows) {
    split_E_planes(M, NP, blockIdx.x - nb_rows, gridDim.x - nb_rows);
    return;
  }
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += nb_rows * blockDim.x) {
    int32_t li = pos[b] - M.item_lo, lj = neg[b] - M.item_lo;
    if (li < 0 || li >= M.item_cnt) li = -1;
    if (lj < 0 || lj >= M.item_cnt) lj = -1;
    M.rows[b] = li;
    M.rows[B + b] = lj;
  }
}

// Unique-row step: the slot rows AND the claims of the item rows in one pass, on the main stream ahead of
// the projection - items.list (the distinct catalog rows of the batch, in claim order) is the row list of
// the projection and of grad_E, upos[row] the list position every slot of that row refers to.  The
// deferred-Adam catch-up of the listed rows follows on the side stream (k_prep, items_listed = 1).
__global__ void __launch_bounds__(256)
k_uniq_rows(FvxModel M, const int32_t* __restrict__ pos, const int32_t* __restrict__ neg, int B, int nb_rows, int NP) {
  if ((int)blockIdx.x >= nb_rows) {
    split_E_planes(M, NP, blockIdx.x - nb_rows, gridDim.x - nb_rows);
    return;
  }
  const int32_t t = (int32_t)(*M.step) + 1;
  const int lane = threadIdx.x & 31;
  for (int b0 = blockIdx.x * blockDim.x; b0 < B; b0 += nb_rows * blockDim.x) {   // warp-uniform trip count
    const int b = b0 + threadIdx.x;
    int32_t li = -1, lj = -1;
    if (b < B) {
      li = pos[b] - M.item_lo;
      lj = neg[b] - M.item_lo;
      if (li < 0 || li >= M.item_cnt) li = -1;
      if (lj < 0 || lj >= M.item_cnt) lj = -1;
      M.rows[b] = li;
      M.rows[B + b] = lj;
    }
    claim_pair_pos(M.items, M.upos, li, lj, t, lane);
  }
}

__global__ void __launch_bounds__(256, 4)
k_prep(FvxModel M, const int32_t* __restrict__ user, const int32_t* __restrict__ pos,
       const int32_t* __restrict__ neg, int B, int nb_mark, int NP, int write_rows, int tpw, int items_listed,
       int part) {     // part (items_listed only): 0 = everything, 1 = the user rows only, 2 = uslot + the listed item rows only
  const int lane = threadIdx.x & 31;
  if ((int)blockIdx.x >= nb_mark) {
    split_E_planes(M, NP, blockIdx.x - nb_mark, gridDim.x - nb_mark);
    return;
  }
  const int32_t done = (int32_t)(*M.step);
  const int32_t t = done + 1;
  const bool deferred = M.adam_mode == FVX_ADAM_DEFERRED;
  if (items_listed && part != 1) {
    // list position of every slot's row (k_uniq_rows has finished): the scoring kernel reads it with its
    // other indices instead of chasing rows -> upos
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < 2 * B; s += nb_mark * blockDim.x) {
      const int32_t r = M.rows[s];
      M.uslot[s] = r >= 0 ? M.upos[r] : 0;
    }
  }
  // a warp takes PREP_TPW triples at a time (lanes 0..PREP_TPW-1 claim their three rows), then
  // replays the claimed rows one after the other with all 32 lanes on the row's columns
  const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (nb_mark * blockDim.x) >> 5;
  for (int b0 = warp_g * tpw; b0 < (part == 2 ? 0 : B); b0 += nwarps * tpw) {
    const int b = b0 + lane;
    const bool live = lane < tpw && b < B;
    int32_t u = -1, li_ = -1, lj = -1;
    bool ustart = false;
    if (live) {
      u = user[b];
      ustart = (b == 0 || user[b - 1] != u);
      if (!items_listed) {
        li_ = pos[b] - M.item_lo;
        lj = neg[b] - M.item_lo;
      }
      if (li_ < 0 || li_ >= M.item_cnt) li_ = -1;
      if (lj < 0 || lj >= M.item_cnt) lj = -1;
      if (write_rows) {
        M.rows[b] = li_;
        M.rows[B + b] = lj;
      }
    }
    // only the OWNER of a user claims and catches up the user's row (one rank: every user is owned)
    uint32_t wu = claim_rows(M.users, u, live && ustart && u >= M.user_lo && u < M.user_lo + M.user_cnt, t, lane);
    uint32_t wi = claim_rows(M.items, li_, li_ >= 0, t, lane);
    uint32_t wj = claim_rows(M.items, lj, lj >= 0, t, lane);
    if (deferred) {
      // the claimed rows of this pass (<= 3 * PREP_TPW), four at a time: 8 lanes per row
      const int grp = lane >> 3, li = lane & 7;
      const int nu_ = __popc(wu), ni_ = __popc(wi), total = nu_ + ni_ + __popc(wj);
      for (int base = 0; base < total; base += 4) {
        const int k = base + grp;
        // k-th claimed row: users first, then pos items, then neg items
        int32_t row = -1;
        int which = 0;
        {
          uint32_t msk = wu; int kk = k;
          if (kk >= nu_) { kk -= nu_; msk = wi; which = 1; if (kk >= ni_) { kk -= ni_; msk = wj; which = 2; } }
          int src = -1;
          if (k < total) { for (int q = 0; q < kk; ++q) msk &= msk - 1; src = __ffs(msk) - 1; }
          const int32_t vu = __shfl_sync(0xffffffffu, u, src < 0 ? 0 : src);
          const int32_t vi = __shfl_sync(0xffffffffu, li_, src < 0 ? 0 : src);
          const int32_t vj = __shfl_sync(0xffffffffu, lj, src < 0 ? 0 : src);
          if (src >= 0) row = which == 0 ? vu : (which == 1 ? vi : vj);
        }
        if (row >= 0) replay_row(which == 0 ? M.users : M.items, row, done, M.lr, li, 8);
      }
    }
  }
  if (items_listed && deferred && part != 1) {
    // the item rows were claimed by k_uniq_rows: catch up the listed rows, 8 lanes per row
    int n = *M.items.count;
    if (n > M.items.list_cap) n = M.items.list_cap;
    const int li = lane & 7;
    const int ngrp = (nb_mark * blockDim.x) >> 3;
    for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; e < n; e += ngrp)
      replay_row(M.items, M.items.list[e], done, M.lr, li, 8);
  }
}

// every row -> step (*step)  (fvx_adam_flush)
__global__ void k_catchup_all(FvxTable T, const int64_t* __restrict__ step, float lr, long long r0, long long r1) {
  const int32_t target = (int32_t)(*step);
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = r0 + (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < r1; r += warps)
    replay_row(T, (int32_t)r, target, lr, lane, 32);
}

// ---------------------------------------------------------------------------------
// Scores, loss and gradients of one batch: one warp per triple, every row of the triple
// in flight at once (user row, two item rows, two theta rows).  Gradients leave as
// red.global.add into the zero-initialised accumulators g (duplicates of a row inside a
// batch - the reference's sampler emits runs of one user, dataset.py:96-99 - are summed
// there, which is the dedup-sum TF's sparse Adam does before squaring).
#define SG_WARPS 8

// TH layout: row stride np, ks K-split partials ss floats apart (summed on read).
struct SgTheta {
  const float* p;
  int np, ks;
  long long ss;
  int chunks, nsm;      // unique-row step: ks is a cap, the split follows the device-side row count
  __device__ __forceinline__ float at(long long slot, int n) const {
    const float* q = p + slot * np + n;
    float v = q[0];
    for (int s = 1; s < ks; ++s) v += q[s * ss];
    return v;
  }
};

template <int MAXV>
__global__ void __launch_bounds__(SG_WARPS * 32)
k_score_grad(FvxModel M, const int32_t* __restrict__ user, int B, int loss_slot, SgTheta T, int wnp, int wpitch) {
  __shared__ double loss_sh[SG_WARPS];
  const int Su = M.users.stride, Si = M.items.stride, K = M.K, d = M.d, de = M.de;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float reg = M.reg, reg2 = 2.0f * M.reg;
  const bool vis = M.D > 0;
  const long long gw = (long long)blockIdx.x * SG_WARPS + warp;
  const long long nw = (long long)gridDim.x * SG_WARPS;
  double loss_acc = 0.0;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(M.W_hi);
  __nv_bfloat16* wl = reinterpret_cast<__nv_bfloat16*>(M.W_lo);

  for (long long b = gw; b < B; b += nw) {
    const int32_t u = user[b];
    const int32_t li = M.rows[b], lj = M.rows[B + b];
    if (li < 0 || lj < 0 || u < 0 || u >= M.num_users) {
      // item id outside the catalog: triple ignored (and no gradient to E either)
      if (vis) {
        const int nw_ = wnp > 0 ? wnp : de;
        for (int n = lane; n < nw_; n += 32) {
          if (wnp > 0) {
            const __nv_bfloat16 z = __float2bfloat16_rn(0.0f);
            wh[(size_t)b * wpitch + n] = z; wl[(size_t)b * wpitch + n] = z;
            wh[(size_t)(B + b) * wpitch + n] = z; wl[(size_t)(B + b) * wpitch + n] = z;
          } else {
            M.W[(size_t)b * de + n] = 0.0f; M.W[(size_t)(B + b) * de + n] = 0.0f;
          }
        }
      }
      continue;
    }
    const float* ur = M.users.w + (size_t)u * Su;
    const float* gi = M.items.w + (size_t)li * Si;
    const float* gj = M.items.w + (size_t)lj * Si;
    float a[MAXV], x[MAXV], y[MAXV], tu[MAXV], dt[MAXV];
#pragma unroll
    for (int q = 0; q < MAXV; ++q) {
      const int c = lane + 32 * q;
      const bool in = c < K;
      a[q] = in ? ur[c] : 0.0f;
      x[q] = in ? gi[c] : 0.0f;
      y[q] = in ? gj[c] : 0.0f;
      const bool iv = vis && c < d;
      tu[q] = iv ? ur[K + c] : 0.0f;
      dt[q] = iv ? T.at(b, c) - T.at(B + b, c) : 0.0f;
    }
    const float bi = gi[K], bj = gj[K];
    const float vb = vis ? T.at(b, d) - T.at(B + b, d) : 0.0f;
    float part = 0.0f, sq = 0.0f;
#pragma unroll
    for (int q = 0; q < MAXV; ++q) {
      part = fmaf(a[q], x[q] - y[q], part);
      part = fmaf(tu[q], dt[q], part);
      sq += a[q] * a[q] + x[q] * x[q] + y[q] * y[q] + tu[q] * tu[q];
    }
    const float xs = fvx_warp_sum(part) + (bi - bj) + vb;
    const float sqs = fvx_warp_sum(sq);
    const bool inside = (xs >= FVX_CLIP_LO) && (xs <= FVX_CLIP_HI);
    const float coef = inside ? -1.0f / (1.0f + expf(xs)) : 0.0f;  // d softplus(-x)/dx
    const float z = -fminf(fmaxf(xs, FVX_CLIP_LO), FVX_CLIP_HI);
    const float sp = z > 13.942385f ? z : (z < -13.942385f ? expf(z) : log1pf(expf(z)));
    if (lane == 0) loss_acc += (double)sp + (double)(reg * sqs) + (double)(reg * bi * bi) +
                               (double)(reg * bj * bj * M.bias_neg_scale);
    float* gu = M.users.g + (size_t)u * Su;
    float* ggi = M.items.g + (size_t)li * Si;
    float* ggj = M.items.g + (size_t)lj * Si;
#pragma unroll
    for (int q = 0; q < MAXV; ++q) {
      const int c = lane + 32 * q;
      if (c < K) {
        fvx_red_add(gu + c, coef * (x[q] - y[q]) + reg2 * a[q]);
        fvx_red_add(ggi + c, coef * a[q] + reg2 * x[q]);
        fvx_red_add(ggj + c, -coef * a[q] + reg2 * y[q]);
      }
      if (vis && c < d) fvx_red_add(gu + K + c, coef * dt[q] + reg2 * tu[q]);
    }
    if (lane == 0) {
      fvx_red_add(ggi + K, coef + reg2 * bi);
      fvx_red_add(ggj + K, -coef + (reg2 * M.bias_neg_scale) * bj);
    }
    if (vis) {
      const int nw_ = wnp > 0 ? wnp : de;
#pragma unroll
      for (int q = 0; q < MAXV; ++q) {
        const int n = lane + 32 * q;
        if (n < nw_) {
          const float wv = n < d ? coef * tu[q] : (n == d ? coef : 0.0f);
          if (wnp > 0) {   // bf16 hi/lo planes for the tensor-core backward
            const __nv_bfloat16 h = __float2bfloat16_rn(wv);
            const __nv_bfloat16 l = __float2bfloat16_rn(wv - __bfloat162float(h));
            wh[(size_t)b * wpitch + n] = h;
            wl[(size_t)b * wpitch + n] = l;
            wh[(size_t)(B + b) * wpitch + n] = __hneg(h);
            wl[(size_t)(B + b) * wpitch + n] = __hneg(l);
          } else {
            M.W[(size_t)b * de + n] = wv;
            M.W[(size_t)(B + b) * de + n] = -wv;
          }
        }
      }
    }
  }
  // one atomic per block
  if (lane == 0) loss_sh[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < SG_WARPS; ++w) s += loss_sh[w];
    if (s != 0.0) atomicAdd(M.loss + loss_slot, s);
  }
}

// The same work with 16-byte loads and 16-byte reductions (red.global.add.v4.f32, sm_90+): a lane owns
// four consecutive columns.  The scatter-add is what bounds this kernel (214 float reductions per
// triple at K=64, d=20 through the L2 atomic units); vectors cut the operation count by four.
// Requires K % 4 == 0 (user row: Tu starts 16-byte aligned).
__device__ __forceinline__ void red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 sg_at4(const SgTheta& T, long long slot, int c4) {
  const float4* q = reinterpret_cast<const float4*>(T.p + slot * T.np) + c4;
  float4 v = q[0];
  for (int s = 1; s < T.ks; ++s) {
    const float4 w = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q) + s * T.ss);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  return v;
}
// theta chunk c4 of the positive and the negative slot, minus each other.  All loads of the common
// splits (1 or 2 partials) are issued before the first use: written as a loop over the partials the
// compiler emits load -> add -> load -> add, four dependent memory round trips per triple (31 % of the
// kernel's stall samples sat on those adds in the round-1 v7 profile).
__device__ __forceinline__ float4 sg_diff4(const SgTheta& T, long long si, long long sj, int c4) {
  const float4* qi = reinterpret_cast<const float4*>(T.p + si * T.np) + c4;
  const float4* qj = reinterpret_cast<const float4*>(T.p + sj * T.np) + c4;
  if (T.ks <= 2) {
    const float4 a0 = qi[0], b0 = qj[0];
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = a1;
    if (T.ks == 2) {
      a1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(qi) + T.ss);
      b1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(qj) + T.ss);
    }
    return make_float4((a0.x + a1.x) - (b0.x + b1.x), (a0.y + a1.y) - (b0.y + b1.y),
                       (a0.z + a1.z) - (b0.z + b1.z), (a0.w + a1.w) - (b0.w + b1.w));
  }
  const float4 ti = sg_at4(T, si, c4), tj = sg_at4(T, sj, c4);
  return make_float4(ti.x - tj.x, ti.y - tj.y, ti.z - tj.z, ti.w - tj.w);
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
__device__ __forceinline__ float4 unpack_bf16x4(uint2 p) {
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&p.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&p.y);
  const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

// 16 lanes per triple (two triples per warp in flight: the kernel is bound by the latency of the
// index -> row -> theta load chain, not by bytes), chunk c of a row handled by lane c % 16.
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// DEDUP (unique-row step): theta of slot (b, side) is row upos[item row] of TH, and the backward
// coefficients are ADDED into W_sum[upos[item row]] (fp32; k_w_planes turns the sums into bf16 planes).
template <int MQ, int MD, bool DEDUP>
__global__ void __launch_bounds__(SG_WARPS * 32, (MQ + MD > 4 ? 2 : 4))   // the wide variants hold up to 100 row values per lane
k_score_grad_v4(FvxModel M, const int32_t* __restrict__ user, int B, int loss_slot, SgTheta T, int wnp, int wpitch) {
  __shared__ double loss_sh[SG_WARPS * 2];
  if (DEDUP) {
    int nv = *M.items.count;
    if (nv > M.items.list_cap) nv = M.items.list_cap;
    int streamk;
    T.ks = fvx_tc_split_dyn((nv + 127) / 128, T.chunks, T.nsm, T.ks, T.nsm, &streamk);
  }
  const int Su = M.users.stride, Si = M.items.stride, K = M.K, d = M.d, de = M.de;
  const int K4 = K >> 2, D4 = (d + 3) >> 2, D4b = (d + 4) >> 2;
  const int grp = threadIdx.x >> 4, sub = threadIdx.x & 15;
  const float reg = M.reg, reg2 = 2.0f * M.reg;
  const bool vis = M.D > 0;
  const long long gg0 = (long long)blockIdx.x * (SG_WARPS * 2) + grp;
  const long long ng = (long long)gridDim.x * (SG_WARPS * 2);
  const int nw4 = (wnp > 0 ? wnp : de) >> 2;
  double loss_acc = 0.0;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long Bpad = ((long long)B + 1) & ~1LL;       // both halves of a warp walk the loop together

  // the indices of the NEXT triple are loaded while this one is processed: one level less in the
  // dependent load chain (index -> rows -> reductions) that bounds the kernel
  int32_t u_n = -1, li_n = -1, lj_n = -1;
  int32_t pi_n = 0, pj_n = 0;            // DEDUP: list positions of the two item rows (k_prep: uslot)
  if (gg0 < B) { u_n = __ldg(user + gg0); li_n = __ldg(M.rows + gg0); lj_n = __ldg(M.rows + B + gg0); }
  if (DEDUP && gg0 < B) { pi_n = __ldg(M.uslot + gg0); pj_n = __ldg(M.uslot + B + gg0); }
  for (long long b = gg0; b < Bpad; b += ng) {
    const bool live = b < B;
    const int32_t u = u_n, li = li_n, lj = lj_n;
    const long long si = DEDUP ? (long long)pi_n : b, sj = DEDUP ? (long long)pj_n : (long long)B + b;
    {
      const long long bn = b + ng;
      u_n = li_n = lj_n = -1;
      if (bn < B) { u_n = __ldg(user + bn); li_n = __ldg(M.rows + bn); lj_n = __ldg(M.rows + B + bn); }
      if (DEDUP) {
        pi_n = pj_n = 0;
        if (bn < B) { pi_n = __ldg(M.uslot + bn); pj_n = __ldg(M.uslot + B + bn); }
      }
    }
    const bool dead = li < 0 || lj < 0 || u < 0 || u >= M.num_users;   // id outside the catalog: triple ignored
    float coef = 0.0f;
    float4 tu[MD];
#pragma unroll
    for (int q = 0; q < MD; ++q) tu[q] = z4;
    float4 a[MQ], x[MQ], y[MQ], dt[MD];
    float bi = 0.f, bj = 0.f, vb = 0.f;
    if (!dead) {
      const float4* ur = reinterpret_cast<const float4*>(M.users.w + (size_t)u * Su);
      const float4* gi = reinterpret_cast<const float4*>(M.items.w + (size_t)li * Si);
      const float4* gj = reinterpret_cast<const float4*>(M.items.w + (size_t)lj * Si);
#pragma unroll
      for (int q = 0; q < MQ; ++q) {
        const int c = sub + 16 * q;
        a[q] = c < K4 ? ur[c] : z4;
        x[q] = c < K4 ? gi[c] : z4;
        y[q] = c < K4 ? gj[c] : z4;
      }
#pragma unroll
      for (int q = 0; q < MD; ++q) {
        const int c = sub + 16 * q;
        dt[q] = z4;
        if (vis && c < D4b) {                // chunks of columns 0 .. d: latent terms and the visual bias
          if (c < D4) tu[q] = ur[K4 + c];
          dt[q] = sg_diff4(T, si, sj, c);
          if (c == (d >> 2)) {               // column d: theta_i[d] - theta_j[d] = the visual-bias difference
            const int e_ = d & 3;
            vb = e_ == 0 ? dt[q].x : (e_ == 1 ? dt[q].y : (e_ == 2 ? dt[q].z : dt[q].w));
          }
          const int n0 = 4 * c;      // columns >= d of the chunk are not latent terms
          if (n0 >= d) { tu[q].x = 0.f; dt[q].x = 0.f; }
          if (n0 + 1 >= d) { tu[q].y = 0.f; dt[q].y = 0.f; }
          if (n0 + 2 >= d) { tu[q].z = 0.f; dt[q].z = 0.f; }
          if (n0 + 3 >= d) { tu[q].w = 0.f; dt[q].w = 0.f; }
        }
      }
      bi = M.items.w[(size_t)li * Si + K];
      bj = M.items.w[(size_t)lj * Si + K];
    } else {
#pragma unroll
      for (int q = 0; q < MQ; ++q) { a[q] = z4; x[q] = z4; y[q] = z4; }
#pragma unroll
      for (int q = 0; q < MD; ++q) dt[q] = z4;
    }
    float part = 0.0f, sq = 0.0f;
#pragma unroll
    for (int q = 0; q < MQ; ++q) {
      const float4 df = make_float4(x[q].x - y[q].x, x[q].y - y[q].y, x[q].z - y[q].z, x[q].w - y[q].w);
      part += dot4(a[q], df);
      sq += dot4(a[q], a[q]) + dot4(x[q], x[q]) + dot4(y[q], y[q]);
    }
#pragma unroll
    for (int q = 0; q < MD; ++q) {
      part += dot4(tu[q], dt[q]);
      sq += dot4(tu[q], tu[q]);
    }
    // the lane that holds column d hands the visual-bias difference to its half-warp (0 elsewhere: a sum does it)
    const float xs = half_sum(part) + (bi - bj) + half_sum(vb);
    const float sqs = half_sum(sq);
    if (!dead) {
      const bool inside = (xs >= FVX_CLIP_LO) && (xs <= FVX_CLIP_HI);
      coef = inside ? -1.0f / (1.0f + expf(xs)) : 0.0f;  // d softplus(-x)/dx
      const float z = -fminf(fmaxf(xs, FVX_CLIP_LO), FVX_CLIP_HI);
      const float sp = z > 13.942385f ? z : (z < -13.942385f ? expf(z) : log1pf(expf(z)));
      if (sub == 0) loss_acc += (double)sp + (double)(reg * sqs) + (double)(reg * bi * bi) +
                                (double)(reg * bj * bj * M.bias_neg_scale);
      float* gu = M.users.g + (size_t)u * Su;
      float* ggi = M.items.g + (size_t)li * Si;
      float* ggj = M.items.g + (size_t)lj * Si;
#pragma unroll
      for (int q = 0; q < MQ; ++q) {
        const int c = sub + 16 * q;
        if (c < K4) {
          const float4 A = a[q], X = x[q], Y = y[q];
          red_add4(gu + 4 * c, make_float4(coef * (X.x - Y.x) + reg2 * A.x, coef * (X.y - Y.y) + reg2 * A.y,
                                           coef * (X.z - Y.z) + reg2 * A.z, coef * (X.w - Y.w) + reg2 * A.w));
          red_add4(ggi + 4 * c, make_float4(coef * A.x + reg2 * X.x, coef * A.y + reg2 * X.y,
                                            coef * A.z + reg2 * X.z, coef * A.w + reg2 * X.w));
          red_add4(ggj + 4 * c, make_float4(-coef * A.x + reg2 * Y.x, -coef * A.y + reg2 * Y.y,
                                            -coef * A.z + reg2 * Y.z, -coef * A.w + reg2 * Y.w));
        }
      }
#pragma unroll
      for (int q = 0; q < MD; ++q) {
        const int c = sub + 16 * q;
        if (vis && c < D4) {
          const float4 U4 = tu[q], Dt = dt[q];
          red_add4(gu + K + 4 * c, make_float4(coef * Dt.x + reg2 * U4.x, coef * Dt.y + reg2 * U4.y,
                                               coef * Dt.z + reg2 * U4.z, coef * Dt.w + reg2 * U4.w));
        }
      }
      if (sub == 0) {
        fvx_red_add(ggi + K, coef + reg2 * bi);
        fvx_red_add(ggj + K, -coef + (reg2 * M.bias_neg_scale) * bj);
      }
    }
    if (vis && live && !(DEDUP && dead)) {
      // W[slot, n] = +-coef * [Tu | 1 | 0...] (zero rows for an ignored triple)
#pragma unroll
      for (int q = 0; q < MD; ++q) {
        const int c = sub + 16 * q;
        if (c < nw4) {
          const int n0 = 4 * c;
          float4 wv = make_float4(coef * tu[q].x, coef * tu[q].y, coef * tu[q].z, coef * tu[q].w);
          if (n0 == d) wv.x = coef;
          if (n0 + 1 == d) wv.y = coef;
          if (n0 + 2 == d) wv.z = coef;
          if (n0 + 3 == d) wv.w = coef;
          if (DEDUP) {     // the slots of one catalog row are summed: one row of the backward operand per row
            if (n0 <= d) {
              red_add4(M.W_sum + (size_t)si * wnp + n0, wv);
              red_add4(M.W_sum + (size_t)sj * wnp + n0, make_float4(-wv.x, -wv.y, -wv.z, -wv.w));
            }
          } else if (wnp > 0) {   // bf16 hi/lo planes for the tensor-core backward
            const uint2 h = pack_bf16x4(wv);
            const float4 hf = unpack_bf16x4(h);
            const uint2 l = pack_bf16x4(make_float4(wv.x - hf.x, wv.y - hf.y, wv.z - hf.z, wv.w - hf.w));
            uint2* wh = reinterpret_cast<uint2*>(M.W_hi);
            uint2* wl = reinterpret_cast<uint2*>(M.W_lo);
            const size_t wp4 = (size_t)(wpitch >> 2);
            wh[(size_t)b * wp4 + c] = h;
            wl[(size_t)b * wp4 + c] = l;
            wh[(size_t)(B + b) * wp4 + c] = make_uint2(h.x ^ 0x80008000u, h.y ^ 0x80008000u);   // negation
            wl[(size_t)(B + b) * wp4 + c] = make_uint2(l.x ^ 0x80008000u, l.y ^ 0x80008000u);
          } else {
            reinterpret_cast<float4*>(M.W)[(size_t)b * nw4 + c] = wv;
            reinterpret_cast<float4*>(M.W)[(size_t)(B + b) * nw4 + c] = make_float4(-wv.x, -wv.y, -wv.z, -wv.w);
          }
        }
      }
    }
  }
  if (sub == 0) loss_sh[grp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < SG_WARPS * 2; ++w) s += loss_sh[w];
    if (s != 0.0) atomicAdd(M.loss + loss_slot, s);
  }
}

// Unique-row step: W_sum[r, 0:NP) (fp32 sums over the slots of listed row r) -> the bf16 hi | lo planes the
// tensor-core backward reads; the sums are zeroed for the next step, and the rows between the list length
// and the next multiple of the backward's 32-row stage are cleared (they hold an older step's planes).
__global__ void __launch_bounds__(256)
k_w_planes(FvxModel M, int NP, int wpitch) {
  int n = *M.items.count;
  if (n > M.items.list_cap) n = M.items.list_cap;
  long long npad = ((long long)n + 31) & ~31LL;
  if (npad > 2LL * M.max_batch) npad = 2LL * M.max_batch;
  const int np4 = NP >> 2;
  const size_t wp4 = (size_t)(wpitch >> 2);
  float4* ws = reinterpret_cast<float4*>(M.W_sum);
  uint2* wh = reinterpret_cast<uint2*>(M.W_hi);
  uint2* wl = reinterpret_cast<uint2*>(M.W_lo);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npad * np4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / np4;
    const int c = (int)(i - r * np4);
    float4 wv = z4;
    if (r < n) {
      wv = ws[i];
      if (wv.x != 0.f || wv.y != 0.f || wv.z != 0.f || wv.w != 0.f) ws[i] = z4;
    }
    const uint2 h = pack_bf16x4(wv);
    const float4 hf = unpack_bf16x4(h);
    const uint2 l = pack_bf16x4(make_float4(wv.x - hf.x, wv.y - hf.y, wv.z - hf.z, wv.w - hf.w));
    wh[(size_t)r * wp4 + c] = h;
    wl[(size_t)r * wp4 + c] = l;
  }
}

// ---------------------------------------------------------------------------------
// One launch for the whole parameter update.  Blocks [0, nb_u): user rows, [nb_u, nb_u+nb_i):
// item rows, the rest: E_ext.  The last block to finish advances the step counter.
struct UpdParams {
  int nb_u, nb_i, nb_e;
  int parts, gnp;       // gE_part: `parts` row-group partials with row stride gnp
  int loss_slot;
  int dense;            // DENSE: sweep the whole tables
  int finalize;         // the last block advances the step counter and resets the lists
  int32_t* sync;        // [1] blocks-done counter (zero between launches)
  const float* gE_src;  // partial gradients of E_ext (model->gE_part, or an all-reduced [D, de] buffer)
  const float* loss_pair;   // sharded step: n_tails x {hi, lo, overflow flag, -}: the loss shares of the ranks (all-reduced
                            //   into one tail, or one tail per rank)
  int n_tails;
  int skip_E;           // GradFashion: E is a composed scratch matrix - its Adam step happens in k_gf_adam
};

__device__ __forceinline__ void adam4(float4& w, float4& m, float4& v, const float4 g, float a) {
#define FVX_ADAM1(f)                                            \
  m.f = FVX_BETA1 * m.f + (1.0f - FVX_BETA1) * g.f;             \
  v.f = FVX_BETA2 * v.f + (1.0f - FVX_BETA2) * (g.f * g.f);     \
  w.f -= a * m.f / (sqrtf(v.f) + FVX_EPS);
  FVX_ADAM1(x) FVX_ADAM1(y) FVX_ADAM1(z) FVX_ADAM1(w)
#undef FVX_ADAM1
}

__device__ __forceinline__ void adam_table_part(const FvxTable& T, int blk, int nblk, float a, int32_t t, bool dense) {
  float4* __restrict__ W = reinterpret_cast<float4*>(T.w);
  float4* __restrict__ Mo = reinterpret_cast<float4*>(T.m);
  float4* __restrict__ V = reinterpret_cast<float4*>(T.v);
  float4* __restrict__ G = reinterpret_cast<float4*>(T.g);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (dense) {   // the reference's literal behaviour: every element of the table moves
    const long long n4 = (T.rows * T.stride) >> 2;
    for (long long i = (long long)blk * blockDim.x + threadIdx.x; i < n4; i += (long long)nblk * blockDim.x) {
      float4 g = G[i], m = Mo[i], v = V[i], w = W[i];
      adam4(w, m, v, g, a);
      Mo[i] = m; V[i] = v; W[i] = w;
      if (g.x != 0.0f || g.y != 0.0f || g.z != 0.0f || g.w != 0.0f) G[i] = z4;
    }
    return;
  }
  const int lane = threadIdx.x & 31;
  int n = *T.count;
  if (n > T.list_cap) n = T.list_cap;
  const int s4 = T.stride >> 2;
  const int warps = (nblk * blockDim.x) >> 5;
  const int w0 = (blk * blockDim.x + threadIdx.x) >> 5;
  for (int e = w0; e < n; e += warps) {
    const int32_t r = T.list[e];
    const size_t o = (size_t)r * s4;
    for (int c = lane; c < s4; c += 32) {
      float4 g = G[o + c], m = Mo[o + c], v = V[o + c], w = W[o + c];
      adam4(w, m, v, g, a);
      Mo[o + c] = m; V[o + c] = v; W[o + c] = w; G[o + c] = z4;
    }
    if (lane == 0) T.last[r] = t;
  }
}

__global__ void __launch_bounds__(256)
k_update(FvxModel M, UpdParams U) {
  const long long t = *M.step + 1;
  const float a = fvx_alpha(M.lr, t);
  const int b = blockIdx.x;
  if (b < U.nb_u) {
    adam_table_part(M.users, b, U.nb_u, a, (int32_t)t, U.dense);
  } else if (b < U.nb_u + U.nb_i) {
    adam_table_part(M.items, b - U.nb_u, U.nb_i, a, (int32_t)t, U.dense);
  } else {
    // dense Adam on E_ext [D,de]: gradient = sum of the row-group partials + 2*reg*E
    // (VBPR.py:129 puts reg*(|E|^2+|Bp|^2) into the loss); adds that loss term as well.
    const int n = U.skip_E ? 0 : M.D * M.de;
    const float reg = M.reg;
    float sq = 0.0f;
    for (int i = (b - U.nb_u - U.nb_i) * blockDim.x + threadIdx.x; i < n; i += U.nb_e * blockDim.x) {
      const int f = i / M.de, c = i - f * M.de;
      // the row-group partials are summed in a fixed order (0, 1, 2, ...) four loads at a time
      const float* gp = U.gE_src + (size_t)f * U.gnp + c;
      const size_t pstride = (size_t)M.D * U.gnp;
      float g = 0.0f;
      int p = 0;
      for (; p + 16 <= U.parts; p += 16) {      // sixteen loads in flight, summed in the same fixed order
        float t_[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) t_[q] = gp[(size_t)(p + q) * pstride];
#pragma unroll
        for (int q = 0; q < 16; ++q) g += t_[q];
      }
      for (; p + 4 <= U.parts; p += 4) {
        const float g0 = gp[(size_t)p * pstride], g1 = gp[(size_t)(p + 1) * pstride];
        const float g2 = gp[(size_t)(p + 2) * pstride], g3 = gp[(size_t)(p + 3) * pstride];
        g = (((g + g0) + g1) + g2) + g3;
      }
      for (; p < U.parts; ++p) g += gp[(size_t)p * pstride];
      const float e = M.E[i];
      sq += e * e;
      g += 2.0f * reg * e;
      const float m = FVX_BETA1 * M.mE[i] + (1.0f - FVX_BETA1) * g;
      const float v = FVX_BETA2 * M.vE[i] + (1.0f - FVX_BETA2) * (g * g);
      M.mE[i] = m;
      M.vE[i] = v;
      M.E[i] = e - a * m / (sqrtf(v) + FVX_EPS);
    }
    sq = fvx_warp_sum(sq);
    if ((threadIdx.x & 31) == 0 && sq != 0.0f && U.loss_slot >= 0) atomicAdd(M.loss + U.loss_slot, (double)(reg * sq));
    if (U.loss_pair && b == U.nb_u + U.nb_i && threadIdx.x == 0 && U.loss_slot >= 0) {
      // a batch with more runs than the exchange buffers hold poisons the loss: nothing fails silently
      double l = 0.0;
      bool over = false;
      for (int r = 0; r < U.n_tails; ++r) {
        l += (double)U.loss_pair[4 * r] + (double)U.loss_pair[4 * r + 1];
        over = over || U.loss_pair[4 * r + 2] != 0.0f;
      }
      atomicAdd(M.loss + U.loss_slot, over ? (double)__int_as_float(0x7fc00000) : l);
    }
  }
  // last block: the step is complete
  if (!U.finalize) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int prev = atomicAdd(U.sync, 1);
    if (prev == (int)gridDim.x - 1) {
      U.sync[0] = 0;
      U.sync[1] = 0;                     // owned-slot counter of the sharded step
      *M.step += 1;
      *M.users.count = 0;
      *M.items.count = 0;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------
// GradFashion (GradFashion.py:97-134): theta_ext = F * Eff with Eff = blockdiag(Ec, Ee) * E2, F = [Fc | Fe].
// k_gf_compose builds Eff into M.E ahead of the step; after grad_E, k_gf_grads turns the gradient of Eff (the
// row-group partials) into the gradients of Ec, Ee, E2 (+ 2 reg w, GradFashion.py:173-176) and k_gf_adam applies
// Keras-Adam to the three tensors and adds their L2 loss term.
__global__ void k_gf_compose(FvxModel M) {
  const int n = M.D * M.de, ecee = M.ec + M.ee;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int f = i / M.de, c = i - f * M.de;
    float s = 0.0f;
    if (c <= M.d) {
      if (f < M.Dc) { for (int r = 0; r < M.ec; ++r) s = fmaf(M.Ec[(size_t)f * M.ec + r], M.E2[(size_t)r * M.de + c], s); }
      else { for (int r = 0; r < M.ee; ++r) s = fmaf(M.Ee[(size_t)(f - M.Dc) * M.ee + r], M.E2[(size_t)(M.ec + r) * M.de + c], s); }
    }
    (void)ecee;
    M.E[i] = s;
  }
}
// scratch layout: gEff [D*de] | gEc [Dc*ec] | gEe [De*ee] | gE2 [(ec+ee)*de]
__global__ void k_gf_reduce(FvxModel M, int parts, int gnp) {
  const int n = M.D * M.de;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int f = i / M.de, c = i - f * M.de;
    float g = 0.0f;
    for (int p = 0; p < parts; ++p) g += M.gE_part[((size_t)p * M.D + f) * gnp + c];
    M.gf_scratch[i] = g;
  }
}
__global__ void k_gf_grads(FvxModel M) {
  const float* gEff = M.gf_scratch;
  float* gEc = M.gf_scratch + (size_t)M.D * M.de;
  float* gEe = gEc + (size_t)M.Dc * M.ec;
  float* gE2 = gEe + (size_t)M.De * M.ee;
  const float reg2 = 2.0f * M.reg;
  const int nEc = M.Dc * M.ec, nEe = M.De * M.ee, nE2 = (M.ec + M.ee) * M.de;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nEc + nEe + nE2; i += gridDim.x * blockDim.x) {
    if (i < nEc) {                     // dEc[f, r] = sum_c gEff[f, c] * E2[r, c]
      const int f = i / M.ec, r = i - f * M.ec;
      float g = 0.0f;
      for (int c = 0; c <= M.d; ++c) g = fmaf(gEff[(size_t)f * M.de + c], M.E2[(size_t)r * M.de + c], g);
      gEc[i] = g + reg2 * M.Ec[i];
    } else if (i < nEc + nEe) {        // dEe[f, r] = sum_c gEff[Dc + f, c] * E2[ec + r, c]
      const int j = i - nEc, f = j / M.ee, r = j - f * M.ee;
      float g = 0.0f;
      for (int c = 0; c <= M.d; ++c) g = fmaf(gEff[(size_t)(M.Dc + f) * M.de + c], M.E2[(size_t)(M.ec + r) * M.de + c], g);
      gEe[j] = g + reg2 * M.Ee[j];
    } else {                           // dE2[r, c] = sum_f Mblock[f, r] * gEff[f, c]
      const int j = i - nEc - nEe, r = j / M.de, c = j - r * M.de;
      float g = 0.0f;
      if (c <= M.d) {
        if (r < M.ec) { for (int f = 0; f < M.Dc; ++f) g = fmaf(M.Ec[(size_t)f * M.ec + r], gEff[(size_t)f * M.de + c], g); }
        else { for (int f = 0; f < M.De; ++f) g = fmaf(M.Ee[(size_t)f * M.ee + (r - M.ec)], gEff[(size_t)(M.Dc + f) * M.de + c], g); }
        g += reg2 * M.E2[j];
      }
      gE2[j] = g;
    }
  }
}
__global__ void k_gf_adam(FvxModel M, int loss_slot) {
  const long long t = *M.step + 1;
  const float a = fvx_alpha(M.lr, t);
  const float* gEc = M.gf_scratch + (size_t)M.D * M.de;
  const float* gEe = gEc + (size_t)M.Dc * M.ec;
  const float* gE2 = gEe + (size_t)M.De * M.ee;
  const int nEc = M.Dc * M.ec, nEe = M.De * M.ee, nE2 = (M.ec + M.ee) * M.de;
  float sq = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nEc + nEe + nE2; i += gridDim.x * blockDim.x) {
    float *w, *m, *v;
    float g;
    if (i < nEc) { w = M.Ec + i; m = M.mEc + i; v = M.vEc + i; g = gEc[i]; }
    else if (i < nEc + nEe) { const int j = i - nEc; w = M.Ee + j; m = M.mEe + j; v = M.vEe + j; g = gEe[j]; }
    else { const int j = i - nEc - nEe; w = M.E2 + j; m = M.mE2 + j; v = M.vE2 + j; g = gE2[j]; }
    const float w0 = *w;
    sq += w0 * w0;
    const float mn = FVX_BETA1 * *m + (1.0f - FVX_BETA1) * g;
    const float vn = FVX_BETA2 * *v + (1.0f - FVX_BETA2) * (g * g);
    *m = mn; *v = vn;
    *w = w0 - a * mn / (sqrtf(vn) + FVX_EPS);
  }
  sq = fvx_warp_sum(sq);
  if ((threadIdx.x & 31) == 0 && sq != 0.0f && loss_slot >= 0) atomicAdd(M.loss + loss_slot, (double)(M.reg * sq));
}
int fvx_launch_gf_compose(const FvxModel* m, cudaStream_t st) {
  k_gf_compose<<<(m->D * m->de + 255) / 256, 256, 0, st>>>(*m);
  FVX_CHECK_LAUNCH("k_gf_compose");
  return 0;
}
// gradient of the effective matrix (row-group partials, pitch gnp) -> Adam on Ec, Ee, E2; before the step counter moves
static int gf_backward(const FvxModel* m, int parts, int gnp, int loss_slot, cudaStream_t st) {
  k_gf_reduce<<<(m->D * m->de + 255) / 256, 256, 0, st>>>(*m, parts, gnp);
  const int n = m->Dc * m->ec + m->De * m->ee + (m->ec + m->ee) * m->de;
  k_gf_grads<<<(n + 127) / 128, 128, 0, st>>>(*m);
  k_gf_adam<<<(n + 255) / 256, 256, 0, st>>>(*m, loss_slot);
  FVX_CHECK_LAUNCH("k_gf_backward");
  return 0;
}

// ---------------------------------------------------------------------------------
static int check_model(const FvxModel* m, const char* who) {
  FVX_CHECK_ARG(m != nullptr, "%s: null model", who);
  FVX_CHECK_ARG(m->abi_version == FVX_ABI_VERSION, "%s: FvxModel.abi_version %d != %d", who, m->abi_version,
                FVX_ABI_VERSION);
  FVX_CHECK_ARG(m->num_users > 0 && m->num_items > 0 && m->item_cnt > 0 && m->K > 0, "%s: bad geometry", who);
  FVX_CHECK_ARG(m->item_lo >= 0 && m->item_lo + m->item_cnt <= m->num_items, "%s: bad item shard", who);
  FVX_CHECK_ARG(m->user_lo >= 0 && m->user_cnt >= 0 && m->user_lo + m->user_cnt <= m->num_users && m->users.rows >= m->num_users,
                "%s: bad user block [%d, +%d) of %d", who, m->user_lo, m->user_cnt, m->num_users);
  FVX_CHECK_ARG(m->users.stride % 4 == 0 && m->users.stride >= m->K + m->d, "%s: bad user stride", who);
  FVX_CHECK_ARG(m->items.stride % 4 == 0 && m->items.stride >= m->K + 1, "%s: bad item stride", who);
  FVX_CHECK_ARG(m->users.w && m->items.w && m->step, "%s: null table pointer", who);
  if (m->D > 0) {
    FVX_CHECK_ARG(m->d > 0 && m->de % 4 == 0 && m->de >= m->d + 1, "%s: bad de", who);
    FVX_CHECK_ARG(m->E && (m->F || (m->use_tensor_cores && m->F_pl)), "%s: VBPR needs E and F", who);
    if (m->two_stage)
      FVX_CHECK_ARG(m->Dc > 0 && m->De > 0 && m->Dc + m->De == m->D && m->ec > 0 && m->ee > 0 && m->Ec && m->Ee && m->E2 &&
                    m->mEc && m->vEc && m->mEe && m->vEe && m->mE2 && m->vE2 && m->gf_scratch,
                    "%s: GradFashion (two_stage) needs Ec, Ee, E2, their Adam moments and the scratch", who);
  }
  return 0;
}

static inline int warp_grid(long long rows, int block = 256, int per_sm = 8) {
  long long g = (rows * 32 + block - 1) / block;
  long long cap = (long long)fvx_num_sms() * per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

int fvx_launch_score_grad(const FvxModel* m, const int32_t* user, int B, int loss_slot, int th_ks, cudaStream_t st,
                          int dedup) {
  long long g = ((long long)B + SG_WARPS - 1) / SG_WARPS;
  long long cap = (long long)fvx_num_sms() * 8;
  if (g > cap) g = cap;
  SgTheta T;
  T.p = m->TH;
  const bool tc = m->D > 0 && m->use_tensor_cores;
  T.np = tc ? fvx_tc_np(m->de) : m->de;
  T.ks = tc ? th_ks : 1;
  T.ss = 2LL * B * T.np;
  T.chunks = m->D > 0 ? m->D / 64 : 1;
  T.nsm = fvx_num_sms();
  const int wnp = tc ? T.np : 0;
  const int wpitch = tc ? fvx_w_pitch(m) : 0;
  const int need = (m->K > m->d + 1 ? m->K : m->d + 1);
  FVX_CHECK_ARG(need <= 288 && (wnp == 0 || wnp <= 320), "fvx_bpr_step: K=%d / d=%d too large for the score kernel",
                m->K, m->d);
  const int wcols = wnp > 0 ? wnp : m->de;
  if (m->K % 4 == 0 && m->K <= 256 && ((m->d <= 252 && wcols <= 256) || (wnp > 0 && wcols <= 320))) {
    // vector path: a lane owns 4 columns, 16 lanes per triple (MD = 5: the 320 padded columns of embed_d = 256 on
    // the tensor-core path)
    const long long g2 = (g + 1) / 2 > 0 ? (g + 1) / 2 : 1;
    if (dedup) {
      FVX_CHECK_ARG(tc && m->upos && m->W_sum && m->uslot, "fvx_bpr_step: unique-row step without upos / W_sum / uslot");
      if (m->K <= 64 && wcols <= 64)
        k_score_grad_v4<1, 1, true><<<(int)g2, SG_WARPS * 32, 0, st>>>(*m, user, B, loss_slot, T, wnp, wpitch);
      else if (m->K <= 128 && wcols <= 128)
        k_score_grad_v4<2, 2, true><<<(int)g2, SG_WARPS * 32, 0, st>>>(*m, user, B, loss_slot, T, wnp, wpitch);
      else if (wcols <= 256)
        k_score_grad_v4<4, 4, true><<<(int)g2, SG_WARPS * 32, 0, st>>>(*m, user, B, loss_slot, T, wnp, wpitch);
      else
        k_score_grad_v4<4, 5, true><<<(int)g2, SG_WARPS * 32, 0, st>>>(*m, user, B, loss_slot, T, wnp, wpitch);
    } else if (m->K <= 64 && wcols <= 64)
      k_score_grad_v4<1, 1, false><<<(int)g2, SG_WARPS * 32, 0, st>>>(*m, user, B, loss_slot, T, wnp, wpitch);
    else if (m->K <= 128 && wcols <= 128)
      k_score_grad_v4<2, 2, false><<<(int)g2, SG_WARPS * 32, 0, st>>>(*m, user, B, loss_slot, T, wnp, wpitch);
    else if (wcols <= 256)
      k_score_grad_v4<4, 4, false><<<(int)g2, SG_WARPS * 32, 0, st>>>(*m, user, B, loss_slot, T, wnp, wpitch);
    else
      k_score_grad_v4<4, 5, false><<<(int)g2, SG_WARPS * 32, 0, st>>>(*m, user, B, loss_slot, T, wnp, wpitch);
  } else if (dedup) {
    FVX_FAIL(-2, "fvx_bpr_step: the unique-row step needs K %% 4 == 0");
  } else {
    // K % 4 != 0: the scalar kernel, a lane per column (up to 9 x 32 = 288 columns: configs[4] with embed_d = 256)
    k_score_grad<9><<<(int)g, SG_WARPS * 32, 0, st>>>(*m, user, B, loss_slot, T, wnp, wpitch);
  }
  FVX_CHECK_LAUNCH("k_score_grad");
  return 0;
}

int fvx_launch_prep(const FvxModel* m, const int32_t* user, const int32_t* pos, const int32_t* neg, int B,
                    cudaStream_t st, int what) {
  const bool tc = m->D > 0 && m->use_tensor_cores;
  const int NP = tc ? fvx_tc_np(m->de) : m->de;
  const int nb_e = tc ? 32 : 0;
  if (what == FVX_PREP_ROWS) {
    int nb_rows = (B + 255) / 256;
    if (nb_rows > fvx_num_sms() * 4) nb_rows = fvx_num_sms() * 4;
    k_rows_et<<<nb_rows + nb_e, 256, 0, st>>>(*m, pos, neg, B, nb_rows, NP);
    FVX_CHECK_LAUNCH("k_rows_et");
    return 0;
  }
  if (what == FVX_PREP_UNIQ) {
    FVX_CHECK_ARG(m->upos != nullptr, "fvx_bpr_step: unique-row step without upos");
    int nb_rows = (B + 255) / 256;
    if (nb_rows > fvx_num_sms() * 4) nb_rows = fvx_num_sms() * 4;
    k_uniq_rows<<<nb_rows + nb_e, 256, 0, st>>>(*m, pos, neg, B, nb_rows, NP);
    FVX_CHECK_LAUNCH("k_uniq_rows");
    return 0;
  }
  // item-sharded ranks own few of the slots they scan: a full warp of triples per pass there
  const int tpw = (m->item_cnt < m->num_items) ? 32 : PREP_TPW;
  int nb_mark = (B + 8 * tpw - 1) / (8 * tpw);     // 8 warps per block
  if (nb_mark > fvx_num_sms() * 8) nb_mark = fvx_num_sms() * 8;
  const bool all = what == FVX_PREP_ALL;
  const int listed = (what == FVX_PREP_CLAIMS_LISTED || what == FVX_PREP_USERS_ONLY || what == FVX_PREP_ITEMS_LISTED) ? 1 : 0;
  k_prep<<<nb_mark + (all ? nb_e : 0), 256, 0, st>>>(*m, user, pos, neg, B, nb_mark, NP, all ? 1 : 0, tpw, listed,
                                                     what == FVX_PREP_USERS_ONLY ? 1 : (what == FVX_PREP_ITEMS_LISTED ? 2 : 0));
  FVX_CHECK_LAUNCH("k_prep");
  return 0;
}

int fvx_launch_update(const FvxModel* m, int B, int parts, int gnp, const float* gE_src, int loss_slot,
                      cudaStream_t st, int what, const float* loss_pair, int n_tails) {
  UpdParams U;
  U.dense = m->adam_mode == FVX_ADAM_DENSE;
  if (U.dense) {
    U.nb_u = fvx_num_sms() * 4;
    U.nb_i = fvx_num_sms() * 4;
  } else {
    U.nb_u = warp_grid(B, 256, 2);
    U.nb_i = warp_grid(2LL * B, 256, 6);
  }
  U.nb_e = m->D > 0 ? (m->D * m->de + 255) / 256 : 0;
  if (U.nb_e > fvx_num_sms() * 4) U.nb_e = fvx_num_sms() * 4;
  // split launch: the table rows (no finalisation) beside grad_E, then E_ext + finalisation
  if (what == FVX_UPD_TABLES) U.nb_e = 0;
  if (what == FVX_UPD_E) { U.nb_u = 0; U.nb_i = 0; if (U.nb_e == 0) U.nb_e = 1; }
  U.finalize = what != FVX_UPD_TABLES;
  U.parts = parts; U.gnp = gnp; U.loss_slot = loss_slot; U.sync = m->sync; U.gE_src = gE_src; U.loss_pair = loss_pair; U.n_tails = n_tails; U.skip_E = m->two_stage ? 1 : 0;
  if (loss_pair && U.nb_e == 0) U.nb_e = 1;
  k_update<<<U.nb_u + U.nb_i + U.nb_e, 256, 0, st>>>(*m, U);
  FVX_CHECK_LAUNCH("k_update");
  return 0;
}

int fvx_launch_w_planes(const FvxModel* m, int B, cudaStream_t st) {
  const int NP = fvx_tc_np(m->de);
  long long g = (2LL * B * (NP >> 2) + 255) / 256;
  if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
  k_w_planes<<<(int)(g < 1 ? 1 : g), 256, 0, st>>>(*m, NP, fvx_w_pitch(m));
  FVX_CHECK_LAUNCH("k_w_planes");
  return 0;
}

int fvx_check_model(const FvxModel* m, const char* who) { return check_model(m, who); }

// Unique-row step: ON by default where the model carries upos / W_sum;
// FVX_STEP_DEDUP=0 or the hook below (tests compare the two paths in one process) turn it off.
static int g_dedup = -1;
extern "C" int fvx_debug_set_dedup(int on) {
  const int old = g_dedup;
  g_dedup = on ? 1 : 0;
  return old;
}
bool fvx_dedup_enabled() {
  if (g_dedup < 0) {
    const char* e = getenv("FVX_STEP_DEDUP");
    g_dedup = (e && atoi(e) == 0) ? 0 : 1;
  }
  return g_dedup == 1;
}

// Side stream of the two-stream step schedule: one per device, created on first use.
// FVX_STEP_OVERLAP=0 keeps every kernel of the step on the caller's stream.
struct SideStream {
  cudaStream_t s;
  cudaEvent_t fork, prep_done, score_done, upd_done;
};
static SideStream* side_stream() {
  static SideStream pool[FVX_MAX_DEV];
  static int state[FVX_MAX_DEV];   // 0: untried, 1: ready, -1: unavailable
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("FVX_STEP_OVERLAP");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled) return nullptr;
  const int dev = fvx_cur_device();
  if (state[dev] == 0) {
    SideStream& p = pool[dev];
    bool ok = cudaStreamCreateWithFlags(&p.s, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p.prep_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p.score_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p.upd_done, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) cudaGetLastError();
    state[dev] = ok ? 1 : -1;
  }
  return state[dev] == 1 ? &pool[dev] : nullptr;
}

// fork / join of the side stream for the other translation units (fvx_train_sharded.cu)
cudaStream_t fvx_side_begin(cudaStream_t main_stream) {
  SideStream* p = side_stream();
  if (!p) return nullptr;
  cudaEventRecord(p->fork, main_stream);
  cudaStreamWaitEvent(p->s, p->fork, 0);
  return p->s;
}
void fvx_side_join(cudaStream_t main_stream) {
  SideStream* p = side_stream();
  if (!p) return;
  cudaEventRecord(p->prep_done, p->s);
  cudaStreamWaitEvent(main_stream, p->prep_done, 0);
}

// Two-stream timeline of the unique-row step (diagnostics, declared in fvx.h): with tracing on, the step
// records a timing event after every kernel on the stream that kernel runs on; fvx_debug_trace_read
// returns their times in microseconds relative to the first one.
enum { TR_BEGIN = 0, TR_UNIQ, TR_FWD, TR_PREP0, TR_PREP1, TR_SCORE, TR_WPL, TR_GRADE, TR_UPD0, TR_UPD1, TR_END, TR_COUNT };
static cudaEvent_t g_trace_ev[TR_COUNT];
static int g_trace_on = 0;
extern "C" int fvx_debug_trace(int on) {
  if (on && !g_trace_ev[0])
    for (int i = 0; i < TR_COUNT; ++i)
      if (cudaEventCreate(&g_trace_ev[i]) != cudaSuccess) return -3;
  g_trace_on = on ? 1 : 0;
  return 0;
}
extern "C" int fvx_debug_trace_read(float* us_host) {   // [TR_COUNT]; synchronises
  if (!g_trace_ev[0]) return -2;
  if (cudaEventSynchronize(g_trace_ev[TR_END]) != cudaSuccess) return -3;
  for (int i = 0; i < TR_COUNT; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_trace_ev[TR_BEGIN], g_trace_ev[i]) != cudaSuccess) { cudaGetLastError(); ms = -1e-3f; }
    us_host[i] = ms * 1e3f;
  }
  return 0;
}
#define TRACE(i, stream) do { if (g_trace_on) cudaEventRecord(g_trace_ev[i], (stream)); } while (0)

// phases of one step, in launch order (FVX_N_PHASES entries; see fvx.h)
enum { PH_PREP = 0, PH_PROJECT, PH_SCORE_GRAD, PH_GRAD_E, PH_UPDATE, PH_COUNT };

// DEFERRED mode: the Adam step of the rows a batch touched is NOT applied at the end of the step.  The
// gradient stays in g and the row is brought up to date - pending step first, then the zero-gradient
// steps - when it is next needed (replay_row: the catch-up of the next step that touches it, or
// fvx_adam_flush).  One read-modify-write of (w, m, v, g) per touched row instead of two, and one
// kernel less per step; DENSE / LAZY keep the row update (k_update, tables part).
bool fvx_merged_update(const FvxModel* m) { return m->adam_mode == FVX_ADAM_DEFERRED; }
static int step_update(const FvxModel* m, int B, int parts, int gnp, int loss_slot, cudaStream_t st, int what) {
  if (m->two_stage && what != FVX_UPD_TABLES) {
    if (int rc = gf_backward(m, parts, gnp, loss_slot, st)) return rc;
  }
  if (fvx_merged_update(m)) {
    if (what == FVX_UPD_TABLES) return 0;
    what = FVX_UPD_E;
  }
  return fvx_launch_update(m, B, parts, gnp, m->gE_part, loss_slot, st, what);
}

// use_side: 1 = two-stream schedule where it pays (VBPR), 0 = everything on `st` (timed entry point, and the
// small batches of the graph-replayed loop, whose kernels are too short for a fork / join to pay)
static int bpr_step_impl(const FvxModel* model, const int32_t* user, const int32_t* pos, const int32_t* neg,
                         int32_t B, int32_t loss_slot, cudaStream_t st, cudaEvent_t* ev, int use_side) {
  if (int rc = check_model(model, "fvx_bpr_step")) return rc;
  const FvxModel& M = *model;
  FVX_CHECK_ARG(user && pos && neg, "fvx_bpr_step: null batch pointer");
  FVX_CHECK_ARG(B >= 1 && B <= M.max_batch, "fvx_bpr_step: B=%d outside [1, max_batch=%d]", B, M.max_batch);
  FVX_CHECK_ARG(M.item_lo == 0 && M.item_cnt == M.num_items,
                "fvx_bpr_step: model is item-sharded; use the sharded entry points");
  FVX_CHECK_ARG(loss_slot >= 0 && loss_slot < M.loss_slots, "fvx_bpr_step: loss_slot out of range");
  FVX_CHECK_ARG(M.users.list_cap >= B && M.items.list_cap >= 2 * B, "fvx_bpr_step: touched-row lists too small");
  FVX_CHECK_ARG(M.rows != nullptr && M.loss != nullptr && M.sync != nullptr, "fvx_bpr_step: null scratch");
  const bool vis = M.D > 0;
  const bool tc = vis && M.use_tensor_cores;
  const int NP = tc ? fvx_tc_np(M.de) : M.de;
  int th_ks = 1;
  if (vis) FVX_CHECK_ARG(M.TH && M.gE_part && M.ge_parts > 0 && (tc || M.W), "fvx_bpr_step: VBPR scratch missing");
  if (tc) {
    FVX_CHECK_ARG(M.F_pl && M.ET_hi && M.ET_lo && M.W_hi && M.W_lo,
                  "fvx_bpr_step: use_tensor_cores=1 needs the bf16 planes (F_pl, ET_*, W_*)");
    th_ks = fvx_tc_ksplit(&M, 2LL * B);
    while (th_ks > 1 && (long long)th_ks * 2 * B * NP > M.th_cap) th_ks >>= 1;
    FVX_CHECK_ARG((long long)th_ks * 2 * B * NP <= M.th_cap, "fvx_bpr_step: TH scratch too small");
  } else if (vis) {
    FVX_CHECK_ARG(M.F != nullptr, "fvx_bpr_step: fp32 projection needs F");
    FVX_CHECK_ARG(2LL * B * M.de <= M.th_cap, "fvx_bpr_step: TH scratch too small");
  }
#define PHASE(i) do { if (ev) cudaEventRecord(ev[i], st); } while (0)
  if (M.two_stage) {
    if (int rc = fvx_launch_gf_compose(&M, st)) return rc;   // the effective projection matrix of this step
  }

  // Two-stream schedule (VBPR): the claims + deferred-Adam catch-up touch only the embedding tables and run
  // BESIDE the projection; in DENSE / LAZY mode the Adam update of the touched rows runs beside grad_E.  The
  // timed entry point keeps everything on one stream so that each phase is measured alone.
  SideStream* side = (vis && !ev && use_side) ? side_stream() : nullptr;

  // Unique-row step: k_uniq_rows lists the distinct catalog rows of the batch (items.list / upos) ahead of
  // the projection; the projection and grad_E run over that list (its length stays on the device), the
  // scoring kernel reads theta through uslot and sums the backward coefficients per listed row, k_w_planes
  // turns the sums into the bf16 planes grad_E reads.  At B = 65 536 on a 100 k catalog 2B slots are
  // ~58 k distinct rows: both contractions shrink by more than half.
  const bool dedup = tc && M.upos && M.W_sum && M.uslot && M.K % 4 == 0 && M.K <= 256 && M.d <= 256 && NP <= 320 &&
                     fvx_dedup_enabled();
  if (dedup) {
    int ks_cap = 8;
    while (ks_cap > 1 && (long long)ks_cap * 2 * B * NP > M.th_cap) ks_cap >>= 1;
    const int32_t* cnt = M.items.count;
    PHASE(PH_PREP);
    TRACE(TR_BEGIN, st);
    if (int rc = fvx_launch_prep(&M, user, pos, neg, B, st, FVX_PREP_UNIQ)) return rc;
    TRACE(TR_UNIQ, st);
    // Submission order matters: a kernel submitted first fills the SMs first.  The bandwidth-bound
    // projection kernels (one CTA per SM) go in AHEAD of the latency-bound side kernels, which then
    // run in the registers the projection CTAs leave; the other way round the side kernel's blocks
    // occupy every SM and the projection starts only when they retire.
    if (side) cudaEventRecord(side->fork, st);
    else if (int rc = fvx_launch_prep(&M, user, pos, neg, B, st, FVX_PREP_CLAIMS_LISTED)) return rc;
    PHASE(PH_PROJECT);
    if (int rc = fvx_launch_project_tc(&M, M.items.list, 0, 2 * B, ks_cap, M.TH, st, cnt, 1)) return rc;
    TRACE(TR_FWD, st);
    if (side) {
      cudaStreamWaitEvent(side->s, side->fork, 0);
      TRACE(TR_PREP0, side->s);
      if (int rc = fvx_launch_prep(&M, user, pos, neg, B, side->s, FVX_PREP_CLAIMS_LISTED)) return rc;
      TRACE(TR_PREP1, side->s);
      cudaEventRecord(side->prep_done, side->s);
      cudaStreamWaitEvent(st, side->prep_done, 0);
    }
    PHASE(PH_SCORE_GRAD);
    if (int rc = fvx_launch_score_grad(&M, user, B, loss_slot, ks_cap, st, 1)) return rc;
    TRACE(TR_SCORE, st);
    PHASE(PH_GRAD_E);
    if (int rc = fvx_launch_w_planes(&M, B, st)) return rc;
    TRACE(TR_WPL, st);
    // (DENSE / LAZY) the row update may not start before grad_E's CTAs are resident: released by the scoring
    // kernel it fills every SM during k_w_planes and grad_E starts only when its whole grid has retired.  It
    // is released by k_w_planes instead; grad_E, next on the main stream, wins that race.
    if (side) cudaEventRecord(side->score_done, st);
    int parts = 0;
    if (int rc = fvx_launch_grad_E_tc(&M, M.items.list, 2 * B, &parts, st, cnt)) return rc;
    TRACE(TR_GRADE, st);
    PHASE(PH_UPDATE);
    if (side && !fvx_merged_update(&M)) {
      cudaStreamWaitEvent(side->s, side->score_done, 0);
      TRACE(TR_UPD0, side->s);
      if (int rc = step_update(&M, B, 0, NP, loss_slot, side->s, FVX_UPD_TABLES)) return rc;
      TRACE(TR_UPD1, side->s);
      cudaEventRecord(side->upd_done, side->s);
      cudaStreamWaitEvent(st, side->upd_done, 0);      // join: the step is complete on `st`
      if (int rc = step_update(&M, B, parts, NP, loss_slot, st, FVX_UPD_E)) return rc;
    } else {
      if (int rc = step_update(&M, B, parts, NP, loss_slot, st, FVX_UPD_ALL)) return rc;
    }
    TRACE(TR_END, st);
    PHASE(PH_COUNT);
    return 0;
  }
  if (side) {
    // one projection per (triple, side) slot: fp32 CUDA-core kernels, or tensor cores without the row lists
    if (int rc = fvx_launch_prep(&M, user, pos, neg, B, st, FVX_PREP_ROWS)) return rc;
    cudaEventRecord(side->fork, st);
    cudaStreamWaitEvent(side->s, side->fork, 0);
    if (int rc = fvx_launch_prep(&M, user, pos, neg, B, side->s, FVX_PREP_CLAIMS)) return rc;
    cudaEventRecord(side->prep_done, side->s);
    if (tc) {
      if (int rc = fvx_launch_project_tc(&M, M.rows, 0, 2 * B, th_ks, M.TH, st)) return rc;
    } else {
      if (int rc = fvx_launch_project(&M, M.rows, 2 * B, M.TH, st)) return rc;
    }
    cudaStreamWaitEvent(st, side->prep_done, 0);
    if (int rc = fvx_launch_score_grad(&M, user, B, loss_slot, th_ks, st)) return rc;
    cudaEventRecord(side->score_done, st);
    cudaStreamWaitEvent(side->s, side->score_done, 0);
    if (int rc = step_update(&M, B, 0, NP, loss_slot, side->s, FVX_UPD_TABLES)) return rc;
    cudaEventRecord(side->upd_done, side->s);
    int parts = 0;
    if (tc) {
      if (int rc = fvx_launch_grad_E_tc(&M, M.rows, 2 * B, &parts, st)) return rc;
    } else {
      if (int rc = fvx_launch_grad_E(&M, M.rows, 2 * B, &parts, st)) return rc;
    }
    cudaStreamWaitEvent(st, side->upd_done, 0);      // join: the step is complete on `st`
    if (int rc = step_update(&M, B, parts, NP, loss_slot, st, FVX_UPD_E)) return rc;
    return 0;
  }

  PHASE(PH_PREP);
  if (int rc = fvx_launch_prep(&M, user, pos, neg, B, st)) return rc;
  PHASE(PH_PROJECT);
  int parts = 0;
  if (tc) {
    if (int rc = fvx_launch_project_tc(&M, M.rows, 0, 2 * B, th_ks, M.TH, st)) return rc;
  } else if (vis) {
    if (int rc = fvx_launch_project(&M, M.rows, 2 * B, M.TH, st)) return rc;
  }
  PHASE(PH_SCORE_GRAD);
  if (int rc = fvx_launch_score_grad(&M, user, B, loss_slot, th_ks, st)) return rc;
  PHASE(PH_GRAD_E);
  if (tc) {
    if (int rc = fvx_launch_grad_E_tc(&M, M.rows, 2 * B, &parts, st)) return rc;
  } else if (vis) {
    if (int rc = fvx_launch_grad_E(&M, M.rows, 2 * B, &parts, st)) return rc;
  }
  PHASE(PH_UPDATE);
  if (int rc = step_update(&M, B, parts, NP, loss_slot, st, FVX_UPD_ALL)) return rc;
  PHASE(PH_COUNT);
#undef PHASE
  return 0;
}

// ---------------------------------------------------------------------------------
// Small-batch regime (the reference's default batch is 256, train_rec.py:23): a step is a handful of
// microsecond kernels and the host's launch cost bounds it.  fvx_bpr_steps runs n consecutive steps on
// batches that lie back to back in epoch-long index arrays: a one-block kernel copies batch `cursor` into the
// model's staging area and advances the cursor, the step's kernels read the staging area - so the launch
// sequence of GRAPH_STEPS steps is the same whatever the batch and is captured ONCE into a CUDA graph
// (per device, keyed by the model struct, the array pointers and B), then replayed.
#define GRAPH_STEPS 8
__global__ void k_fetch_batch(const int32_t* __restrict__ user, const int32_t* __restrict__ pos,
                              const int32_t* __restrict__ neg, int B, int32_t* __restrict__ stage,
                              long long* __restrict__ cursor) {
  const long long c = *cursor;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
    stage[i] = user[c * B + i];
    stage[B + i] = pos[c * B + i];
    stage[2 * B + i] = neg[c * B + i];
  }
}
__global__ void k_cursor_set(long long* cursor, long long v, long long add) { *cursor = add ? *cursor + add : v; }

// the cursor lives behind the three index arrays of the staging area, 8-byte aligned
static inline long long* stage_cursor(const FvxModel* m) {
  return reinterpret_cast<long long*>(m->batch_stage + ((3 * (size_t)m->max_batch + 1) & ~(size_t)1));
}

struct StepGraph {
  FvxModel model;
  const int32_t *user, *pos, *neg;
  int B, loss_slot;
  cudaGraphExec_t exec;
};
static StepGraph g_graphs[FVX_MAX_DEV][4];
static int g_graph_next[FVX_MAX_DEV];

static int one_staged_step(const FvxModel* m, const int32_t* user, const int32_t* pos, const int32_t* neg, int B,
                           int loss_slot, cudaStream_t st) {
  long long* cursor = stage_cursor(m);
  int g = (B + 255) / 256;
  if (g > fvx_num_sms()) g = fvx_num_sms();
  k_fetch_batch<<<g, 256, 0, st>>>(user, pos, neg, B, m->batch_stage, cursor);
  FVX_CHECK_LAUNCH("k_fetch_batch");
  if (int rc = bpr_step_impl(m, m->batch_stage, m->batch_stage + B, m->batch_stage + 2 * B, B, loss_slot, st, nullptr,
                             B >= 16384 ? 1 : 0))
    return rc;
  k_cursor_set<<<1, 1, 0, st>>>(cursor, 0, 1);
  FVX_CHECK_LAUNCH("k_cursor_set");
  return 0;
}

extern "C" {

int fvx_bpr_step(const FvxModel* model, const int32_t* user, const int32_t* pos, const int32_t* neg,
                 int32_t B, int32_t loss_slot, fvx_stream_t stream) {
  return bpr_step_impl(model, user, pos, neg, B, loss_slot, fvx_cu(stream), nullptr, 1);
}

int fvx_bpr_steps(const FvxModel* model, const int32_t* user, const int32_t* pos, const int32_t* neg,
                  int64_t first, int32_t n_steps, int32_t B, int32_t loss_slot, fvx_stream_t stream) {
  if (int rc = check_model(model, "fvx_bpr_steps")) return rc;
  FVX_CHECK_ARG(user && pos && neg && first >= 0 && n_steps >= 0, "fvx_bpr_steps: bad arguments");
  FVX_CHECK_ARG(model->batch_stage != nullptr, "fvx_bpr_steps: FvxModel.batch_stage is not allocated");
  FVX_CHECK_ARG(B >= 1 && B <= model->max_batch, "fvx_bpr_steps: B=%d outside [1, max_batch=%d]", B, model->max_batch);
  cudaStream_t st = fvx_cu(stream);
  long long* cursor = stage_cursor(model);
  k_cursor_set<<<1, 1, 0, st>>>(cursor, (long long)first, 0);
  FVX_CHECK_LAUNCH("k_cursor_set");
  int left = n_steps;
  if (left >= GRAPH_STEPS) {
    const int dev = fvx_cur_device();
    StepGraph* G = nullptr;
    for (int i = 0; i < 4; ++i) {
      StepGraph& c = g_graphs[dev][i];
      if (c.exec && c.user == user && c.pos == pos && c.neg == neg && c.B == B && c.loss_slot == loss_slot &&
          memcmp(&c.model, model, sizeof(FvxModel)) == 0) { G = &c; break; }
    }
    if (!G) {
      StepGraph& c = g_graphs[dev][g_graph_next[dev]++ & 3];
      if (c.exec) { cudaGraphExecDestroy(c.exec); c.exec = nullptr; }
      if (B >= 16384) side_stream();       // per-device helpers are created before the capture starts
      // captured on a stream of our own (the caller's may be the legacy default stream, which cannot be
      // captured) and launched on the caller's
      static cudaStream_t cap[FVX_MAX_DEV];
      if (!cap[dev] && cudaStreamCreateWithFlags(&cap[dev], cudaStreamNonBlocking) != cudaSuccess)
        FVX_FAIL(-3, "fvx_bpr_steps: cannot create the capture stream: %s", cudaGetErrorString(cudaGetLastError()));
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(cap[dev], cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        FVX_FAIL(-3, "fvx_bpr_steps: cannot begin the capture: %s", cudaGetErrorString(cudaGetLastError()));
      int rc = 0;
      for (int s = 0; s < GRAPH_STEPS && rc == 0; ++s) rc = one_staged_step(model, user, pos, neg, B, loss_slot, cap[dev]);
      cudaError_t ce = cudaStreamEndCapture(cap[dev], &graph);
      if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess || !graph) FVX_FAIL(-3, "fvx_bpr_steps: capture failed: %s", cudaGetErrorString(ce));
      ce = cudaGraphInstantiate(&c.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { c.exec = nullptr; FVX_FAIL(-3, "fvx_bpr_steps: instantiate failed: %s", cudaGetErrorString(ce)); }
      c.model = *model; c.user = user; c.pos = pos; c.neg = neg; c.B = B; c.loss_slot = loss_slot;
      G = &c;
    }
    for (; left >= GRAPH_STEPS; left -= GRAPH_STEPS)
      if (cudaGraphLaunch(G->exec, st) != cudaSuccess)
        FVX_FAIL(-3, "fvx_bpr_steps: graph launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  for (; left > 0; --left)
    if (int rc = one_staged_step(model, user, pos, neg, B, loss_slot, st)) return rc;
  return 0;
}

int fvx_bpr_step_timed(const FvxModel* model, const int32_t* user, const int32_t* pos, const int32_t* neg,
                       int32_t B, int32_t loss_slot, float* phase_ms_host, fvx_stream_t stream) {
  FVX_CHECK_ARG(phase_ms_host != nullptr, "fvx_bpr_step_timed: null output");
  cudaEvent_t ev[PH_COUNT + 1];
  for (int i = 0; i <= PH_COUNT; ++i)
    if (cudaEventCreate(&ev[i]) != cudaSuccess) FVX_FAIL(-3, "fvx_bpr_step_timed: cudaEventCreate failed");
  int rc = bpr_step_impl(model, user, pos, neg, B, loss_slot, fvx_cu(stream), ev, 0);
  if (rc == 0) {
    cudaError_t e = cudaEventSynchronize(ev[PH_COUNT]);
    if (e != cudaSuccess) {
      fvx_set_error("fvx_bpr_step_timed: %s", cudaGetErrorString(e));
      rc = -3;
    } else {
      for (int i = 0; i < PH_COUNT; ++i) cudaEventElapsedTime(&phase_ms_host[i], ev[i], ev[i + 1]);
    }
  }
  for (int i = 0; i <= PH_COUNT; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

int fvx_adam_flush(const FvxModel* model, fvx_stream_t stream) {
  if (int rc = check_model(model, "fvx_adam_flush")) return rc;
  const FvxModel& M = *model;
  if (M.adam_mode != FVX_ADAM_DEFERRED) return 0;
  cudaStream_t st = fvx_cu(stream);
  // users: the rows this rank owns (every row on one GPU); items: the owned shard
  k_catchup_all<<<warp_grid(M.user_cnt), 256, 0, st>>>(M.users, M.step, M.lr, M.user_lo, (long long)M.user_lo + M.user_cnt);
  k_catchup_all<<<warp_grid(M.items.rows), 256, 0, st>>>(M.items, M.step, M.lr, 0, M.items.rows);
  FVX_CHECK_LAUNCH("k_catchup_all");
  return 0;
}

}  // extern "C"
