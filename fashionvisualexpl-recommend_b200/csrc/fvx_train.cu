// The BPR optimiser step (single rank): replaces BPRMF.train_step / VBPR.train_step
// (src/recommender/models/BPRMF.py:87-125, VBPR.py:99-144) and the Keras-Adam update
// they call (math: SURVEY.md Appendix A; restated in oracle/bpr.py).
//
// Kernel sequence per step (all on one stream, no host sync):
//   k_mark        unique touched user / item rows -> lists; local row ids for the GEMMs
//   k_catchup x2  DEFERRED Adam: replay skipped zero-gradient steps of touched rows
//   projection    TH = F[rows] * E_ext                      (fvx_project.cu, VBPR only)
//   k_score_grad  x_uij, loss, gradient coefficients, scatter-add into g, W for dE
//   grad_E        gE_part = F[rows]^T * W                   (fvx_project.cu, VBPR only)
//   k_adam_rows x2 / k_adam_sweep x2   Adam on touched rows / on whole tables (DENSE)
//   k_adam_E      dense Adam on E_ext (+ its L2 term)       (VBPR only)
//   k_finish      step += 1, reset lists
#include <cuda_bf16.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"

// ---------------------------------------------------------------------------------
__device__ __forceinline__ void touch_row(const FvxTable& T, int32_t r, int32_t t) {
  const int32_t old = atomicMax(&T.mark[r], t);
  if (old < t) {
    const int32_t idx = atomicAdd(T.count, 1);
    if (idx < T.list_cap) T.list[idx] = r;
  }
}

__global__ void k_mark(FvxModel M, const int32_t* __restrict__ user, const int32_t* __restrict__ pos,
                       const int32_t* __restrict__ neg, int B) {
  const int32_t t = (int32_t)(*M.step) + 1;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const int32_t u = user[b];
    if (b == 0 || user[b - 1] != u) touch_row(M.users, u, t);
    int32_t li = pos[b] - M.item_lo, lj = neg[b] - M.item_lo;
    if (li < 0 || li >= M.item_cnt) li = -1;
    if (lj < 0 || lj >= M.item_cnt) lj = -1;
    M.rows[b] = li;
    M.rows[B + b] = lj;
    if (li >= 0) touch_row(M.items, li, t);
    if (lj >= 0) touch_row(M.items, lj, t);
  }
}

// ---------------------------------------------------------------------------------
// DEFERRED Adam catch-up: a row last brought up to step `last` has, under the
// reference's dense-semantics Adam, taken zero-gradient steps last+1 .. target:
//   m <- b1*m ; v <- b2*v ; w <- w - alpha_tau * m / (sqrt(v) + eps).
// The loop is truncated after FVX_REPLAY_MAX iterations (the remaining updates are
// below 2e-9 of the first one); m and v then take their closed-form decay.
__device__ __forceinline__ void replay_row(const FvxTable& T, int32_t r, int32_t target, float lr, int lane) {
  const int32_t last = T.last[r];
  const int32_t gap = target - last;
  if (gap <= 0) return;
  if (last > 0) {  // rows never updated have m = v = 0: nothing moves
    const int n = gap < FVX_REPLAY_MAX ? gap : FVX_REPLAY_MAX;
    const int rem = gap - n;
    const float p1_0 = (float)pow(0.9, (double)(last + 1));
    const float p2_0 = (float)pow(0.999, (double)(last + 1));
    const float d1 = rem > 0 ? (float)pow(0.9, (double)rem) : 1.0f;
    const float d2 = rem > 0 ? (float)pow(0.999, (double)rem) : 1.0f;
    float* __restrict__ w = T.w + (size_t)r * T.stride;
    float* __restrict__ m = T.m + (size_t)r * T.stride;
    float* __restrict__ v = T.v + (size_t)r * T.stride;
    for (int c = lane; c < T.stride; c += 32) {
      float wc = w[c], mc = m[c], vc = v[c];
      float p1 = p1_0, p2 = p2_0;
      for (int k = 0; k < n; ++k) {
        const float a = lr * fvx_sqrt_approx(1.0f - p2) * fvx_rcp_approx(1.0f - p1);
        mc *= FVX_BETA1;
        vc *= FVX_BETA2;
        wc -= a * mc * fvx_rcp_approx(fvx_sqrt_approx(vc) + FVX_EPS);
        p1 *= FVX_BETA1;
        p2 *= FVX_BETA2;
      }
      w[c] = wc;
      m[c] = mc * d1;
      v[c] = vc * d2;
    }
  }
  __syncwarp();
  if (lane == 0) T.last[r] = target;
}

// rows in T.list -> step (*step); one warp per row
__global__ void k_catchup_list(FvxTable T, const int64_t* __restrict__ step, float lr) {
  const int32_t target = (int32_t)(*step);
  const int lane = threadIdx.x & 31;
  int n = *T.count;
  if (n > T.list_cap) n = T.list_cap;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n; e += warps)
    replay_row(T, T.list[e], target, lr, lane);
}

// every row -> step (*step)  (fvx_adam_flush)
__global__ void k_catchup_all(FvxTable T, const int64_t* __restrict__ step, float lr) {
  const int32_t target = (int32_t)(*step);
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < T.rows; r += warps)
    replay_row(T, (int32_t)r, target, lr, lane);
}

// ---------------------------------------------------------------------------------
// Scores, loss and gradients of one batch.  One warp walks TPW consecutive triples so
// that the run of equal users the reference's sampler produces (dataset.py:96-99) is
// reduced in shared memory and leaves as ONE atomic row update.
#define SG_TPW 8
#define SG_WARPS 8

// TH layout: row stride th_np, th_ks K-split partials th_ss floats apart (summed here).
struct SgTheta {
  const float* p;
  int np, ks;
  long long ss;
  __device__ __forceinline__ float at(long long slot, int n) const {
    const float* q = p + slot * np + n;
    float v = q[0];
    for (int s = 1; s < ks; ++s) v += q[s * ss];
    return v;
  }
};

__global__ void __launch_bounds__(SG_WARPS * 32)
k_score_grad(FvxModel M, const int32_t* __restrict__ user, int B, int loss_slot, SgTheta T, int wnp) {
  extern __shared__ float sg_smem[];
  const int Su = M.users.stride, Si = M.items.stride, K = M.K, d = M.d, de = M.de;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* urow = sg_smem + (size_t)warp * 2 * Su;  // current user's row
  float* uacc = urow + Su;                        // its gradient accumulator
  const float reg = M.reg, reg2 = 2.0f * M.reg;
  const bool vis = M.D > 0;
  const long long gw = (long long)blockIdx.x * SG_WARPS + warp;
  const long long nw = (long long)gridDim.x * SG_WARPS;
  double loss_acc = 0.0;

  for (long long b0 = gw * SG_TPW; b0 < B; b0 += nw * SG_TPW) {
    int32_t cur_u = -1;
    const int bend = (int)((b0 + SG_TPW < B) ? b0 + SG_TPW : B);
    for (int b = (int)b0; b < bend; ++b) {
      const int32_t u = user[b];
      if (u != cur_u) {
        if (cur_u >= 0) {
          float* g = M.users.g + (size_t)cur_u * Su;
          for (int c = lane; c < Su; c += 32) fvx_red_add(g + c, uacc[c]);
        }
        const float* src = M.users.w + (size_t)u * Su;
        for (int c = lane; c < Su; c += 32) { urow[c] = src[c]; uacc[c] = 0.0f; }
        cur_u = u;
        __syncwarp();
      }
      const int32_t li = M.rows[b], lj = M.rows[B + b];
      if (li < 0 || lj < 0) {  // item id outside the catalog: triple ignored (no gradient to E either)
        if (vis && wnp > 0) {
          for (int n = lane; n < wnp; n += 32) {
            const __nv_bfloat16 z = __float2bfloat16_rn(0.0f);
            reinterpret_cast<__nv_bfloat16*>(M.W_hi)[(size_t)b * wnp + n] = z;
            reinterpret_cast<__nv_bfloat16*>(M.W_lo)[(size_t)b * wnp + n] = z;
            reinterpret_cast<__nv_bfloat16*>(M.W_hi)[(size_t)(B + b) * wnp + n] = z;
            reinterpret_cast<__nv_bfloat16*>(M.W_lo)[(size_t)(B + b) * wnp + n] = z;
          }
        } else if (vis) {
          for (int n = lane; n < de; n += 32) { M.W[(size_t)b * de + n] = 0.0f; M.W[(size_t)(B + b) * de + n] = 0.0f; }
        }
        continue;
      }
      const float* gi = M.items.w + (size_t)li * Si;
      const float* gj = M.items.w + (size_t)lj * Si;
      float part = 0.0f, sq = 0.0f;
      for (int c = lane; c < K; c += 32) {
        const float a = urow[c], x = gi[c], y = gj[c];
        part = fmaf(a, x - y, part);
        sq += a * a + x * x + y * y;
      }
      const float bi = gi[K], bj = gj[K];
      float vb = 0.0f;
      float dth0 = 0.0f;   // (theta_i - theta_j)[lane]; further columns are re-read below when d > 32
      if (vis) {
        for (int n = lane; n < d; n += 32) {
          const float a = urow[K + n];
          const float dt = T.at(b, n) - T.at(B + b, n);
          if (n < 32) dth0 = dt;
          part = fmaf(a, dt, part);
          sq += a * a;
        }
        vb = T.at(b, d) - T.at(B + b, d);
      }
      const float x = fvx_warp_sum(part) + (bi - bj) + vb;
      const float sqs = fvx_warp_sum(sq);
      const bool inside = (x >= FVX_CLIP_LO) && (x <= FVX_CLIP_HI);
      const float coef = inside ? -1.0f / (1.0f + expf(x)) : 0.0f;  // d softplus(-x)/dx
      const float z = -fminf(fmaxf(x, FVX_CLIP_LO), FVX_CLIP_HI);
      const float sp = z > 13.942385f ? z : (z < -13.942385f ? expf(z) : log1pf(expf(z)));
      if (lane == 0) loss_acc += (double)sp + (double)(reg * sqs) + (double)(reg * bi * bi) +
                                 (double)(reg * bj * bj / 10.0f);
      float* ggi = M.items.g + (size_t)li * Si;
      float* ggj = M.items.g + (size_t)lj * Si;
      for (int c = lane; c < K; c += 32) {
        const float a = urow[c], xg = gi[c], yg = gj[c];
        uacc[c] += coef * (xg - yg) + reg2 * a;
        fvx_red_add(ggi + c, coef * a + reg2 * xg);
        fvx_red_add(ggj + c, -coef * a + reg2 * yg);
      }
      if (lane == 0) {
        fvx_red_add(ggi + K, coef + reg2 * bi);
        fvx_red_add(ggj + K, -coef + (reg2 / 10.0f) * bj);
      }
      if (vis) {
        const int nw = wnp > 0 ? wnp : de;
        for (int n = lane; n < nw; n += 32) {
          float wv = 0.0f;
          if (n < d) {
            const float a = urow[K + n];
            const float dt = n < 32 ? dth0 : T.at(b, n) - T.at(B + b, n);
            uacc[K + n] += coef * dt + reg2 * a;
            wv = coef * a;
          } else if (n == d) {
            wv = coef;
          }
          if (wnp > 0) {   // bf16 hi/lo planes for the tensor-core backward
            const __nv_bfloat16 h = __float2bfloat16_rn(wv);
            const __nv_bfloat16 l = __float2bfloat16_rn(wv - __bfloat162float(h));
            __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(M.W_hi);
            __nv_bfloat16* wl = reinterpret_cast<__nv_bfloat16*>(M.W_lo);
            wh[(size_t)b * wnp + n] = h;
            wl[(size_t)b * wnp + n] = l;
            wh[(size_t)(B + b) * wnp + n] = __hneg(h);
            wl[(size_t)(B + b) * wnp + n] = __hneg(l);
          } else {
            M.W[(size_t)b * de + n] = wv;
            M.W[(size_t)(B + b) * de + n] = -wv;
          }
        }
      }
      __syncwarp();
    }
    if (cur_u >= 0) {
      float* g = M.users.g + (size_t)cur_u * Su;
      for (int c = lane; c < Su; c += 32) fvx_red_add(g + c, uacc[c]);
    }
    __syncwarp();
  }
  if (lane == 0 && loss_acc != 0.0) atomicAdd(M.loss + loss_slot, loss_acc);
}

// ---------------------------------------------------------------------------------
// Adam on the touched rows (DEFERRED / LAZY): one warp per row of T.list.
__global__ void k_adam_rows(FvxTable T, const int64_t* __restrict__ step, float lr) {
  const long long t = *step + 1;
  const float a = fvx_alpha(lr, t);
  const int lane = threadIdx.x & 31;
  int n = *T.count;
  if (n > T.list_cap) n = T.list_cap;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n; e += warps) {
    const int32_t r = T.list[e];
    const size_t o = (size_t)r * T.stride;
    for (int c = lane; c < T.stride; c += 32) {
      const float g = T.g[o + c];
      const float m = FVX_BETA1 * T.m[o + c] + (1.0f - FVX_BETA1) * g;
      const float v = FVX_BETA2 * T.v[o + c] + (1.0f - FVX_BETA2) * (g * g);
      T.m[o + c] = m;
      T.v[o + c] = v;
      T.w[o + c] -= a * m / (sqrtf(v) + FVX_EPS);
      T.g[o + c] = 0.0f;
    }
    if (lane == 0) T.last[r] = (int32_t)t;
  }
}

// Adam on every element of the table (DENSE: the reference's literal behaviour).
__global__ void k_adam_sweep(FvxTable T, const int64_t* __restrict__ step, float lr) {
  const long long t = *step + 1;
  const float a = fvx_alpha(lr, t);
  const long long n4 = (T.rows * T.stride) >> 2;
  float4* __restrict__ W = reinterpret_cast<float4*>(T.w);
  float4* __restrict__ Mo = reinterpret_cast<float4*>(T.m);
  float4* __restrict__ V = reinterpret_cast<float4*>(T.v);
  float4* __restrict__ G = reinterpret_cast<float4*>(T.g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 g = G[i], m = Mo[i], v = V[i], w = W[i];
#define FVX_ADAM1(f)                                            \
  m.f = FVX_BETA1 * m.f + (1.0f - FVX_BETA1) * g.f;             \
  v.f = FVX_BETA2 * v.f + (1.0f - FVX_BETA2) * (g.f * g.f);     \
  w.f -= a * m.f / (sqrtf(v.f) + FVX_EPS);
    FVX_ADAM1(x) FVX_ADAM1(y) FVX_ADAM1(z) FVX_ADAM1(w)
#undef FVX_ADAM1
    Mo[i] = m; V[i] = v; W[i] = w;
    if (g.x != 0.0f || g.y != 0.0f || g.z != 0.0f || g.w != 0.0f) G[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Dense Adam on E_ext [D,de]: gradient = sum of the per-group partials + 2*reg*E
// (VBPR.py:129 puts reg*(|E|^2+|Bp|^2) into the loss); adds that loss term as well.
// gE_part rows have stride gnp floats (de on the fp32 path, NP on the tensor-core path).
__global__ void k_adam_E(FvxModel M, int parts, int loss_slot, const float* __restrict__ extra_grad, int gnp) {
  const long long t = *M.step + 1;
  const float a = fvx_alpha(M.lr, t);
  const int n = M.D * M.de;
  const float reg = M.reg;
  float sq = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float g = extra_grad ? extra_grad[i] : 0.0f;
    const int f = i / M.de, c = i - f * M.de;
    for (int p = 0; p < parts; ++p) g += M.gE_part[((size_t)p * M.D + f) * gnp + c];
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
  if ((threadIdx.x & 31) == 0 && sq != 0.0f && loss_slot >= 0) atomicAdd(M.loss + loss_slot, (double)(reg * sq));
}

__global__ void k_finish(FvxModel M) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *M.step += 1;
    *M.users.count = 0;
    *M.items.count = 0;
  }
}

// ---------------------------------------------------------------------------------
static int check_model(const FvxModel* m, const char* who) {
  FVX_CHECK_ARG(m != nullptr, "%s: null model", who);
  FVX_CHECK_ARG(m->abi_version == FVX_ABI_VERSION, "%s: FvxModel.abi_version %d != %d", who, m->abi_version,
                FVX_ABI_VERSION);
  FVX_CHECK_ARG(m->num_users > 0 && m->num_items > 0 && m->item_cnt > 0 && m->K > 0, "%s: bad geometry", who);
  FVX_CHECK_ARG(m->item_lo >= 0 && m->item_lo + m->item_cnt <= m->num_items, "%s: bad item shard", who);
  FVX_CHECK_ARG(m->users.stride % 4 == 0 && m->users.stride >= m->K + m->d, "%s: bad user stride", who);
  FVX_CHECK_ARG(m->items.stride % 4 == 0 && m->items.stride >= m->K + 1, "%s: bad item stride", who);
  FVX_CHECK_ARG(m->users.w && m->items.w && m->step, "%s: null table pointer", who);
  if (m->D > 0) {
    FVX_CHECK_ARG(m->d > 0 && m->de % 4 == 0 && m->de >= m->d + 1, "%s: bad de", who);
    FVX_CHECK_ARG(m->E && (m->F || (m->use_tensor_cores && m->F_pl)), "%s: VBPR needs E and F", who);
  }
  return 0;
}

static inline int warp_grid(long long rows, int block = 256, int per_sm = 8) {
  long long g = (rows * 32 + block - 1) / block;
  long long cap = (long long)fvx_num_sms() * per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

int fvx_launch_score_grad(const FvxModel* m, const int32_t* user, int B, int loss_slot, int th_ks, cudaStream_t st) {
  const size_t smem = (size_t)SG_WARPS * 2 * m->users.stride * sizeof(float);
  FVX_CHECK_ARG(smem <= 48 * 1024, "fvx_bpr_step: K+d too large for the score kernel (%zu B smem)", smem);
  long long groups = ((long long)B + SG_TPW - 1) / SG_TPW;
  long long g = (groups + SG_WARPS - 1) / SG_WARPS;
  long long cap = (long long)fvx_num_sms() * 8;
  if (g > cap) g = cap;
  SgTheta T;
  T.p = m->TH;
  const bool tc = m->D > 0 && m->use_tensor_cores;
  T.np = tc ? fvx_tc_np(m->de) : m->de;
  T.ks = tc ? th_ks : 1;
  T.ss = 2LL * B * T.np;
  k_score_grad<<<(int)g, SG_WARPS * 32, smem, st>>>(*m, user, B, loss_slot, T, tc ? T.np : 0);
  FVX_CHECK_LAUNCH("k_score_grad");
  return 0;
}

// phases of one step, in launch order (FVX_N_PHASES entries; see fvx.h)
enum { PH_MARK = 0, PH_CATCHUP, PH_PROJECT, PH_SCORE_GRAD, PH_GRAD_E, PH_ADAM_ROWS, PH_ADAM_E, PH_FINISH, PH_COUNT };

static int bpr_step_impl(const FvxModel* model, const int32_t* user, const int32_t* pos, const int32_t* neg,
                         int32_t B, int32_t loss_slot, cudaStream_t st, cudaEvent_t* ev) {
  if (int rc = check_model(model, "fvx_bpr_step")) return rc;
  const FvxModel& M = *model;
  FVX_CHECK_ARG(user && pos && neg, "fvx_bpr_step: null batch pointer");
  FVX_CHECK_ARG(B >= 1 && B <= M.max_batch, "fvx_bpr_step: B=%d outside [1, max_batch=%d]", B, M.max_batch);
  FVX_CHECK_ARG(M.item_lo == 0 && M.item_cnt == M.num_items,
                "fvx_bpr_step: model is item-sharded; use the sharded entry points");
  FVX_CHECK_ARG(loss_slot >= 0 && loss_slot < M.loss_slots, "fvx_bpr_step: loss_slot out of range");
  FVX_CHECK_ARG(M.users.list_cap >= B && M.items.list_cap >= 2 * B, "fvx_bpr_step: touched-row lists too small");
  FVX_CHECK_ARG(M.rows != nullptr && M.loss != nullptr, "fvx_bpr_step: null scratch");
  const bool vis = M.D > 0;
  const bool tc = vis && M.use_tensor_cores;
  int th_ks = 1;
  if (vis) FVX_CHECK_ARG(M.TH && M.gE_part && M.ge_parts > 0 && (tc || M.W), "fvx_bpr_step: VBPR scratch missing");
  if (tc) {
    FVX_CHECK_ARG(M.F_pl && M.ET_hi && M.ET_lo && M.W_hi && M.W_lo,
                  "fvx_bpr_step: use_tensor_cores=1 needs the bf16 planes (F_pl, ET_*, W_*)");
    th_ks = fvx_tc_ksplit(&M, 2LL * B);
    while (th_ks > 1 && (long long)th_ks * 2 * B * fvx_tc_np(M.de) > M.th_cap) th_ks >>= 1;
    FVX_CHECK_ARG((long long)th_ks * 2 * B * fvx_tc_np(M.de) <= M.th_cap, "fvx_bpr_step: TH scratch too small");
  } else if (vis) {
    FVX_CHECK_ARG(M.F != nullptr, "fvx_bpr_step: fp32 projection needs F");
    FVX_CHECK_ARG(2LL * B * M.de <= M.th_cap, "fvx_bpr_step: TH scratch too small");
  }
#define PHASE(i) do { if (ev) cudaEventRecord(ev[i], st); } while (0)

  PHASE(PH_MARK);
  k_mark<<<warp_grid((B + 31) / 32), 256, 0, st>>>(M, user, pos, neg, B);
  FVX_CHECK_LAUNCH("k_mark");
  PHASE(PH_CATCHUP);
  if (M.adam_mode == FVX_ADAM_DEFERRED) {
    k_catchup_list<<<warp_grid(B), 256, 0, st>>>(M.users, M.step, M.lr);
    k_catchup_list<<<warp_grid(2LL * B), 256, 0, st>>>(M.items, M.step, M.lr);
    FVX_CHECK_LAUNCH("k_catchup_list");
  }
  PHASE(PH_PROJECT);
  if (tc) {
    if (int rc = fvx_launch_split_E(&M, st)) return rc;
    if (int rc = fvx_launch_project_tc(&M, M.rows, 0, 2 * B, th_ks, M.TH, st)) return rc;
  } else if (vis) {
    if (int rc = fvx_launch_project(&M, M.rows, 2 * B, M.TH, st)) return rc;
  }
  PHASE(PH_SCORE_GRAD);
  if (int rc = fvx_launch_score_grad(&M, user, B, loss_slot, th_ks, st)) return rc;
  PHASE(PH_GRAD_E);
  int parts = 0;
  if (tc) {
    if (int rc = fvx_launch_grad_E_tc(&M, M.rows, 2 * B, &parts, st)) return rc;
  } else if (vis) {
    if (int rc = fvx_launch_grad_E(&M, M.rows, 2 * B, &parts, st)) return rc;
  }
  PHASE(PH_ADAM_ROWS);
  if (M.adam_mode == FVX_ADAM_DENSE) {
    k_adam_sweep<<<fvx_num_sms() * 8, 256, 0, st>>>(M.users, M.step, M.lr);
    k_adam_sweep<<<fvx_num_sms() * 8, 256, 0, st>>>(M.items, M.step, M.lr);
    FVX_CHECK_LAUNCH("k_adam_sweep");
  } else {
    k_adam_rows<<<warp_grid(B), 256, 0, st>>>(M.users, M.step, M.lr);
    k_adam_rows<<<warp_grid(2LL * B), 256, 0, st>>>(M.items, M.step, M.lr);
    FVX_CHECK_LAUNCH("k_adam_rows");
  }
  PHASE(PH_ADAM_E);
  if (vis) {
    const int n = M.D * M.de;
    int g = (n + 255) / 256;
    if (g > fvx_num_sms() * 4) g = fvx_num_sms() * 4;
    k_adam_E<<<g, 256, 0, st>>>(M, parts, loss_slot, nullptr, tc ? fvx_tc_np(M.de) : M.de);
    FVX_CHECK_LAUNCH("k_adam_E");
  }
  PHASE(PH_FINISH);
  k_finish<<<1, 32, 0, st>>>(M);
  FVX_CHECK_LAUNCH("k_finish");
  PHASE(PH_COUNT);
#undef PHASE
  return 0;
}

extern "C" {

int fvx_bpr_step(const FvxModel* model, const int32_t* user, const int32_t* pos, const int32_t* neg,
                 int32_t B, int32_t loss_slot, fvx_stream_t stream) {
  return bpr_step_impl(model, user, pos, neg, B, loss_slot, fvx_cu(stream), nullptr);
}

int fvx_bpr_step_timed(const FvxModel* model, const int32_t* user, const int32_t* pos, const int32_t* neg,
                       int32_t B, int32_t loss_slot, float* phase_ms_host, fvx_stream_t stream) {
  FVX_CHECK_ARG(phase_ms_host != nullptr, "fvx_bpr_step_timed: null output");
  cudaEvent_t ev[PH_COUNT + 1];
  for (int i = 0; i <= PH_COUNT; ++i)
    if (cudaEventCreate(&ev[i]) != cudaSuccess) FVX_FAIL(-3, "fvx_bpr_step_timed: cudaEventCreate failed");
  int rc = bpr_step_impl(model, user, pos, neg, B, loss_slot, fvx_cu(stream), ev);
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
  k_catchup_all<<<warp_grid(M.users.rows), 256, 0, st>>>(M.users, M.step, M.lr);
  k_catchup_all<<<warp_grid(M.items.rows), 256, 0, st>>>(M.items, M.step, M.lr);
  FVX_CHECK_LAUNCH("k_catchup_all");
  return 0;
}

}  // extern "C"
