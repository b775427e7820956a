// The BPR optimiser step with the item catalog row-sharded over R ranks (one process per GPU).
//
// x_uij = s_ui - s_uj is linear in the item-side quantities, s_ui = Bi[i] + <Gu[u],Gi[i]> +
// <Tu[u], F[i]E> + F[i]Bp (BPRMF.py:74, VBPR.py:82-84), so each rank scores the (triple, side)
// slots whose item it owns and one small all-reduce assembles every x.  Every rank sees the same
// batch; user tables and E are replicated, Gi / Bi / F and their Adam state live on the owner.
//
//   phase A (fvx_bpr_step_sharded_a)  prep (touched rows, catch-up, E planes) -> projection of the
//                                     owned slots -> partial scores S[2B] (0 for foreign slots)
//        -- host: all-reduce(S) --
//   phase B (fvx_bpr_step_sharded_b)  x, loss and gradient coefficients from S; gradients of the
//                                     owned item rows; this rank's share of the user-row gradients
//                                     into the packed run buffer RU; W of the owned slots; grad_E
//                                     of the owned slots reduced to dE[D, de]
//        -- host: all-reduce(RU), all-reduce(dE) --
//   phase C (fvx_bpr_step_sharded_c)  RU rows -> user gradient accumulators; Adam on users (every
//                                     rank, identical), owned items, E (identical)
//
// RU is indexed by RUN: the reference's sampler emits runs of one user (dataset.py:96-99), run_id[b]
// = number of positions <= b where user[b] != user[b-1], minus one.  The layout depends only on the
// batch, so it is the same on every rank and the all-reduce needs no index exchange.  Ownership of
// the per-triple terms that are not tied to an item (softplus loss, user-side L2): the rank that
// owns the POSITIVE item.
#include <cuda_bf16.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"

#define SS_WARPS 8

struct SsTheta {
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

// The slots whose item this rank owns, compacted (order arbitrary): crow[j] = local item row,
// cslot[j] = slot, cpos[slot] = j (or -1); *count = number of owned slots.  The projection and
// grad_E kernels then touch only owned rows (2B/R of them), indexed by j.
__global__ void k_compact_owned(FvxModel M, int B, int32_t* __restrict__ count) {
  int32_t* crow = M.cmap;
  int32_t* cslot = M.cmap + 2 * (size_t)M.max_batch;
  int32_t* cpos = M.cmap + 4 * (size_t)M.max_batch;
  const int lane = threadIdx.x & 31;
  const long long n = 2LL * B, npad = (n + 31) & ~31LL;
  for (long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x; slot < npad;
       slot += (long long)gridDim.x * blockDim.x) {
    const int32_t li = slot < n ? M.rows[slot] : -1;
    const uint32_t b = __ballot_sync(0xffffffffu, li >= 0);
    int base = 0;
    if (b) {
      const int leader = __ffs(b) - 1;
      if (lane == leader) base = atomicAdd(count, __popc(b));
      base = __shfl_sync(0xffffffffu, base, leader);
    }
    if (li >= 0) {
      const int j = base + __popc(b & ((1u << lane) - 1u));
      crow[j] = li;
      cslot[j] = (int32_t)slot;
      cpos[slot] = j;
    } else if (slot < n) {
      cpos[slot] = -1;
    }
  }
}

// phase A: 16 lanes per owned slot, 16-byte loads (K % 4 == 0; the scalar kernel below otherwise)
__device__ __forceinline__ float ss_half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 ss_theta4(const SsTheta& T, long long slot, int c4) {
  const float4* q = reinterpret_cast<const float4*>(T.p + slot * T.np) + c4;
  float4 v = q[0];
  for (int s = 1; s < T.ks; ++s) {
    const float4 w = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q) + s * T.ss);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  return v;
}

// uniq != 0 (unique-row step): theta of a slot is row uslot[slot] of TH (one projection per DISTINCT owned row)
__global__ void __launch_bounds__(SS_WARPS * 32)
k_partial_scores_v4(FvxModel M, const int32_t* __restrict__ user, int B, SsTheta T, float* __restrict__ S,
                    const int32_t* __restrict__ count, int uniq) {
  const int Su = M.users.stride, Si = M.items.stride, K = M.K, d = M.d;
  if (uniq) {
    int nv = *M.items.count, sk;
    if (nv > M.items.list_cap) nv = M.items.list_cap;
    T.ks = fvx_tc_split_dyn((nv + 127) / 128, T.chunks, T.nsm, T.ks, T.nsm, &sk);
  }
  const int K4 = K >> 2, D4 = (d + 3) >> 2;
  const int grp = threadIdx.x >> 4, sub = threadIdx.x & 15;
  const bool vis = M.D > 0;
  const int32_t* crow = M.cmap;
  const int32_t* cslot = M.cmap + 2 * (size_t)M.max_batch;
  const long long ng = (long long)gridDim.x * (SS_WARPS * 2);
  const long long n_owned = *count, npad = (n_owned + 1) & ~1LL;   // both halves of a warp iterate together
  for (long long j = (long long)blockIdx.x * (SS_WARPS * 2) + grp; j < npad; j += ng) {
    const bool live = j < n_owned;
    float part = 0.0f, tail = 0.0f;
    int32_t slot = 0;
    if (live) {
      const int32_t li = crow[j];
      slot = cslot[j];
      const int b = slot < B ? slot : slot - B;
      const int32_t u = user[b];
      const float4* ur = reinterpret_cast<const float4*>(M.users.w + (size_t)u * Su);
      const float4* gi = reinterpret_cast<const float4*>(M.items.w + (size_t)li * Si);
      for (int c = sub; c < K4; c += 16) {
        const float4 a = ur[c], x = gi[c];
        part = fmaf(a.x, x.x, fmaf(a.y, x.y, fmaf(a.z, x.z, fmaf(a.w, x.w, part))));
      }
      const long long tj = uniq ? (long long)M.uslot[slot] : j;
      if (vis)
        for (int c = sub; c < D4; c += 16) {
          const float4 tu = ur[K4 + c], th = ss_theta4(T, tj, c);
          const int n0 = 4 * c;
          part = fmaf(tu.x, th.x, part);
          if (n0 + 1 < d) part = fmaf(tu.y, th.y, part);
          if (n0 + 2 < d) part = fmaf(tu.z, th.z, part);
          if (n0 + 3 < d) part = fmaf(tu.w, th.w, part);
        }
      if (sub == 0) tail = M.items.w[(size_t)li * Si + K] + (vis ? T.at(tj, d) : 0.0f);
    }
    const float s = ss_half_sum(part);
    if (live && sub == 0) S[slot] = s + tail;
  }
}

// phase A: one warp per owned slot
__global__ void __launch_bounds__(SS_WARPS * 32)
k_partial_scores(FvxModel M, const int32_t* __restrict__ user, int B, SsTheta T, float* __restrict__ S,
                 const int32_t* __restrict__ count) {
  const int Su = M.users.stride, Si = M.items.stride, K = M.K, d = M.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vis = M.D > 0;
  const int32_t* crow = M.cmap;
  const int32_t* cslot = M.cmap + 2 * (size_t)M.max_batch;
  const long long nw = (long long)gridDim.x * SS_WARPS;
  const long long n_owned = *count;
  for (long long j = (long long)blockIdx.x * SS_WARPS + warp; j < n_owned; j += nw) {
    const int32_t li = crow[j], slot = cslot[j];
    const int b = slot < B ? slot : slot - B;
    const int32_t u = user[b];
    const float* ur = M.users.w + (size_t)u * Su;
    const float* gi = M.items.w + (size_t)li * Si;
    float part = 0.0f;
    for (int c = lane; c < K; c += 32) part = fmaf(ur[c], gi[c], part);
    if (vis)
      for (int n = lane; n < d; n += 32) part = fmaf(ur[K + n], T.at(j, n), part);
    const float s = fvx_warp_sum(part) + gi[K] + (vis ? T.at(j, d) : 0.0f);
    if (lane == 0) S[slot] = s;
  }
}

// phase B: one warp per triple; each side only if its item is owned
__global__ void __launch_bounds__(SS_WARPS * 32)
k_grads_sharded(FvxModel M, const int32_t* __restrict__ user, int B, int loss_slot, SsTheta T, int wnp, int wpitch,
                const float* __restrict__ S, const int32_t* __restrict__ run_id, float* __restrict__ RU,
                long long ru_rows) {
  __shared__ double loss_sh[SS_WARPS];
  const int Su = M.users.stride, Si = M.items.stride, K = M.K, d = M.d, de = M.de;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float reg = M.reg, reg2 = 2.0f * M.reg;
  const bool vis = M.D > 0;
  const long long nw = (long long)gridDim.x * SS_WARPS;
  double loss_acc = 0.0;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(M.W_hi);
  __nv_bfloat16* wl = reinterpret_cast<__nv_bfloat16*>(M.W_lo);
  const int32_t* crow = M.cmap;
  const int32_t* cpos = M.cmap + 4 * (size_t)M.max_batch;
  for (long long b = (long long)blockIdx.x * SS_WARPS + warp; b < B; b += nw) {
    const int32_t u = user[b];
    const float xs = S[b] - S[B + b];
    const bool inside = (xs >= FVX_CLIP_LO) && (xs <= FVX_CLIP_HI);
    const float coef = inside ? -1.0f / (1.0f + expf(xs)) : 0.0f;
    const float* ur = M.users.w + (size_t)u * Su;
    if (run_id[b] >= ru_rows) {       // more runs than the caller sized RU for: reported through sync[2]
      if (lane == 0) M.sync[2] = 1;
      continue;
    }
    float* ru = RU + (size_t)run_id[b] * Su;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const long long slot = side ? B + b : b;
      const long long j = cpos[slot];           // position in the compact list; < 0: foreign slot
      if (j < 0) continue;
      const int32_t li = crow[j];
      const float cs = side ? -coef : coef;
      const int nw_ = wnp > 0 ? wnp : de;
      const float* gi = M.items.w + (size_t)li * Si;
      float* gg = M.items.g + (size_t)li * Si;
      float sq = 0.0f;
      for (int c = lane; c < K; c += 32) {
        const float a = ur[c], x = gi[c];
        fvx_red_add(gg + c, cs * a + reg2 * x);
        // the user row's data term from this side; its L2 term once per triple (positive side)
        fvx_red_add(ru + c, cs * x + (side == 0 ? reg2 * a : 0.0f));
        sq += x * x + (side == 0 ? a * a : 0.0f);
      }
      const float bi = gi[K];
      if (lane == 0) fvx_red_add(gg + K, cs + (side == 0 ? reg2 : reg2 / 10.0f) * bi);
      if (vis) {
        for (int n = lane; n < nw_; n += 32) {
          float wv = 0.0f;
          if (n < d) {
            const float tu = ur[K + n];
            fvx_red_add(ru + K + n, cs * T.at(j, n) + (side == 0 ? reg2 * tu : 0.0f));
            if (side == 0) sq += tu * tu;
            wv = cs * tu;
          } else if (n == d) {
            wv = cs;
          }
          if (wnp > 0) {
            const __nv_bfloat16 h = __float2bfloat16_rn(wv);
            wh[j * wpitch + n] = h;
            wl[j * wpitch + n] = __float2bfloat16_rn(wv - __bfloat162float(h));
          } else {
            M.W[j * de + n] = wv;
          }
        }
      }
      const float sqs = fvx_warp_sum(sq);
      if (lane == 0) {
        loss_acc += (double)(reg * sqs) + (double)(reg * bi * bi / (side == 0 ? 1.0f : 10.0f));
        if (side == 0) {
          const float z = -fminf(fmaxf(xs, FVX_CLIP_LO), FVX_CLIP_HI);
          loss_acc += (double)(z > 13.942385f ? z : (z < -13.942385f ? expf(z) : log1pf(expf(z))));
        }
      }
    }
  }
  if (lane == 0) loss_sh[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < SS_WARPS; ++w) s += loss_sh[w];
    if (s != 0.0) atomicAdd(M.loss + loss_slot, s);
  }
}


// phase B, scalable form: one warp per OWNED slot (the compact list of phase A), 16-byte loads and
// reductions.  The work of a rank is 2B/R slots however many triples the global batch holds; the
// per-triple terms that are not tied to an item (softplus loss, user-side L2) belong to the slot
// of the POSITIVE item.  Requires K % 4 == 0.
__device__ __forceinline__ void ss_red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ss_at4(const SsTheta& T, long long slot, int c4) {
  const float4* q = reinterpret_cast<const float4*>(T.p + slot * T.np) + c4;
  float4 v = q[0];
  for (int s = 1; s < T.ks; ++s) {
    const float4 w = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q) + s * T.ss);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  return v;
}
__device__ __forceinline__ uint2 ss_pack_bf16x4(float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
__device__ __forceinline__ float4 ss_unpack_bf16x4(uint2 p) {
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&p.x);
  const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&p.y);
  const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

__global__ void __launch_bounds__(SS_WARPS * 32)
k_grads_owned(FvxModel M, const int32_t* __restrict__ user, int B, int loss_slot, SsTheta T, int wnp, int wpitch,
              const float* __restrict__ S, const int32_t* __restrict__ run_id, float* __restrict__ RU,
              long long ru_rows, const int32_t* __restrict__ count, int uniq) {
  __shared__ double loss_sh[SS_WARPS * 2];
  const int Su = M.users.stride, Si = M.items.stride, K = M.K, d = M.d, de = M.de;
  const int K4 = K >> 2, D4 = (d + 3) >> 2;
  const int grp = threadIdx.x >> 4, sub = threadIdx.x & 15;     // 16 lanes per slot
  if (uniq) {
    int nv = *M.items.count, sk;
    if (nv > M.items.list_cap) nv = M.items.list_cap;
    T.ks = fvx_tc_split_dyn((nv + 127) / 128, T.chunks, T.nsm, T.ks, T.nsm, &sk);
  }
  const float reg = M.reg, reg2 = 2.0f * M.reg;
  const bool vis = M.D > 0;
  const long long ng = (long long)gridDim.x * (SS_WARPS * 2);
  const long long n_owned = *count, npad = (n_owned + 1) & ~1LL;
  const int nw4 = (wnp > 0 ? wnp : de) >> 2;
  const int32_t* crow = M.cmap;
  const int32_t* cslot = M.cmap + 2 * (size_t)M.max_batch;
  double loss_acc = 0.0;
  if (blockIdx.x == 0 && wnp > 0 && !uniq) {   // (unique-row step: k_w_planes writes the planes and clears the tail)
    // the last 32-row tile of the backward reads W rows past the owned ones: they must be zero
    const long long cap = 2LL * M.max_batch;
    for (long long i = threadIdx.x; i < 32LL * (wpitch >> 2); i += blockDim.x) {
      const long long r = n_owned + i / (wpitch >> 2);
      if (r < cap) reinterpret_cast<uint2*>(M.W_hi)[(size_t)r * (wpitch >> 2) + i % (wpitch >> 2)] = make_uint2(0u, 0u);
      if (r < cap && wpitch == wnp) reinterpret_cast<uint2*>(M.W_lo)[(size_t)r * (wpitch >> 2) + i % (wpitch >> 2)] = make_uint2(0u, 0u);
    }
  }
  for (long long j = (long long)blockIdx.x * (SS_WARPS * 2) + grp; j < npad; j += ng) {
    bool live = j < n_owned;
    float sq = 0.0f, xs = 0.0f, bi = 0.0f;
    int side = 0;
    if (live) {
      const int32_t li = crow[j], slot = cslot[j];
      side = slot >= B ? 1 : 0;
      const int b = slot - side * B;
      const int32_t u = user[b];
      const int32_t run = run_id[b];
      xs = S[b] - S[B + b];
      const bool inside = (xs >= FVX_CLIP_LO) && (xs <= FVX_CLIP_HI);
      const float coef = inside ? -1.0f / (1.0f + expf(xs)) : 0.0f;
      if (run >= ru_rows) {           // more runs than the caller sized RU for: reported through sync[2]
        if (sub == 0) M.sync[2] = 1;
        live = false;
      } else {
        const float cs = side ? -coef : coef;
        const float ul2 = side ? 0.0f : reg2;       // the user row's L2 term once per triple
        const float4* ur = reinterpret_cast<const float4*>(M.users.w + (size_t)u * Su);
        const float4* gi = reinterpret_cast<const float4*>(M.items.w + (size_t)li * Si);
        float* gg = M.items.g + (size_t)li * Si;
        float* ru = RU + (size_t)run * Su;
        const long long tj = uniq ? (long long)M.uslot[slot] : j;   // row of TH / of the coefficient sums
        for (int c = sub; c < K4; c += 16) {
          const float4 a = ur[c], x = gi[c];
          ss_red_add4(gg + 4 * c, make_float4(cs * a.x + reg2 * x.x, cs * a.y + reg2 * x.y, cs * a.z + reg2 * x.z,
                                              cs * a.w + reg2 * x.w));
          ss_red_add4(ru + 4 * c, make_float4(cs * x.x + ul2 * a.x, cs * x.y + ul2 * a.y, cs * x.z + ul2 * a.z,
                                              cs * x.w + ul2 * a.w));
          sq += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
          if (!side) sq += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
        bi = M.items.w[(size_t)li * Si + K];
        if (sub == 0) fvx_red_add(gg + K, cs + (side == 0 ? reg2 : reg2 / 10.0f) * bi);
        if (vis) {
          for (int c = sub; c < nw4; c += 16) {
            const int n0 = 4 * c;
            float4 tu = make_float4(0.f, 0.f, 0.f, 0.f), wv = tu;
            if (c < D4) {
              tu = ur[K4 + c];
              if (n0 + 1 >= d) tu.y = 0.f;
              if (n0 + 2 >= d) tu.z = 0.f;
              if (n0 + 3 >= d) tu.w = 0.f;
              float4 th = ss_at4(T, tj, c);
              if (n0 + 1 >= d) th.y = 0.f;
              if (n0 + 2 >= d) th.z = 0.f;
              if (n0 + 3 >= d) th.w = 0.f;
              ss_red_add4(ru + K + 4 * c, make_float4(cs * th.x + ul2 * tu.x, cs * th.y + ul2 * tu.y,
                                                      cs * th.z + ul2 * tu.z, cs * th.w + ul2 * tu.w));
              if (!side) sq += tu.x * tu.x + tu.y * tu.y + tu.z * tu.z + tu.w * tu.w;
              wv = make_float4(cs * tu.x, cs * tu.y, cs * tu.z, cs * tu.w);
            }
            if (n0 == d) wv.x = cs;
            if (n0 + 1 == d) wv.y = cs;
            if (n0 + 2 == d) wv.z = cs;
            if (n0 + 3 == d) wv.w = cs;
            if (uniq) {            // the slots of one catalog row are summed: one backward row per distinct row
              if (n0 <= d) ss_red_add4(M.W_sum + (size_t)tj * wnp + n0, wv);
            } else if (wnp > 0) {
              const uint2 h = ss_pack_bf16x4(wv);
              const float4 hf = ss_unpack_bf16x4(h);
              const uint2 l = ss_pack_bf16x4(make_float4(wv.x - hf.x, wv.y - hf.y, wv.z - hf.z, wv.w - hf.w));
              reinterpret_cast<uint2*>(M.W_hi)[(size_t)j * (wpitch >> 2) + c] = h;
              reinterpret_cast<uint2*>(M.W_lo)[(size_t)j * (wpitch >> 2) + c] = l;
            } else {
              reinterpret_cast<float4*>(M.W)[(size_t)j * nw4 + c] = wv;
            }
          }
        }
      }
    }
    const float sqs = ss_half_sum(sq);
    if (live && sub == 0) {
      loss_acc += (double)(reg * sqs) + (double)(reg * bi * bi / (side == 0 ? 1.0f : 10.0f));
      if (side == 0) {
        const float z = -fminf(fmaxf(xs, FVX_CLIP_LO), FVX_CLIP_HI);
        loss_acc += (double)(z > 13.942385f ? z : (z < -13.942385f ? expf(z) : log1pf(expf(z))));
      }
    }
  }
  if (sub == 0) loss_sh[grp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < SS_WARPS * 2; ++w) s += loss_sh[w];
    if (s != 0.0) atomicAdd(M.loss + loss_slot, s);
  }
}

// phase C: one warp per run start adds the all-reduced run gradient into the user's accumulator
__global__ void k_scatter_runs(FvxModel M, const int32_t* __restrict__ user, int B,
                               const int32_t* __restrict__ run_id, const float* __restrict__ RU, long long ru_rows) {
  // a warp looks at 32 triples at once, then adds the rows of the run starts among them (a global
  // batch of R*B triples holds ~R*B/6 runs: most triples are not a run start)
  const int lane = threadIdx.x & 31;
  const int Su = M.users.stride, S4 = Su >> 2;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long npass = ((long long)B + 31) >> 5;
  for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < npass; p += nw) {
    const long long b = p * 32 + lane;
    int32_t u = -1, run = 0;
    bool start = false;
    if (b < B) {
      u = user[b];
      run = run_id[b];
      start = (b == 0 || user[b - 1] != u) && run < ru_rows;
    }
    uint32_t msk = __ballot_sync(0xffffffffu, start);
    while (msk) {
      const int src = __ffs(msk) - 1;
      msk &= msk - 1;
      const int32_t uu = __shfl_sync(0xffffffffu, u, src);
      const int32_t rr = __shfl_sync(0xffffffffu, run, src);
      const float4* s4 = reinterpret_cast<const float4*>(RU + (size_t)rr * Su);
      float* g = M.users.g + (size_t)uu * Su;
      for (int c = lane; c < S4; c += 32) ss_red_add4(g + 4 * c, s4[c]);   // a user may own several runs
    }
  }
}

// run_id[b] = (number of positions <= b where user[b] != user[b-1]) - 1, i.e. the index of the run of
// equal users triple b belongs to.  Two launches: run starts per 4096-element block, then every block
// adds up the counts of the blocks before it (<= 256 values) and scans its own elements.
#define RI_THREADS 256
#define RI_PER 16
__global__ void __launch_bounds__(RI_THREADS)
k_run_count(const int32_t* __restrict__ user, long long n, int32_t* __restrict__ part) {
  __shared__ int sh[RI_THREADS / 32];
  const long long base = (long long)blockIdx.x * RI_THREADS * RI_PER;
  int c = 0;
  for (int e = 0; e < RI_PER; ++e) {
    const long long b = base + (long long)e * RI_THREADS + threadIdx.x;      // coalesced
    if (b < n) c += (b == 0 || user[b] != user[b - 1]) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < RI_THREADS / 32; ++w) t += sh[w];
    part[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(RI_THREADS)
k_run_write(const int32_t* __restrict__ user, long long n, const int32_t* __restrict__ part,
            int32_t* __restrict__ run_id) {
  __shared__ int sh[RI_THREADS / 32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    int t = 0;
    for (int q = 0; q < (int)blockIdx.x; ++q) t += part[q];
    carry = t;
  }
  __syncthreads();
  const long long base = (long long)blockIdx.x * RI_THREADS * RI_PER;
  for (int e = 0; e < RI_PER; ++e) {                      // 256 consecutive elements per round
    const long long b = base + (long long)e * RI_THREADS + threadIdx.x;
    const int f = (b < n && (b == 0 || user[b] != user[b - 1])) ? 1 : 0;
    int x = f;                                            // inclusive scan of the round
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) sh[warp] = x;
    __syncthreads();
    int off = carry;
    for (int w = 0; w < warp; ++w) off += sh[w];
    if (b < n) run_id[b] = off + x - 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = carry;
      for (int w = 0; w < RI_THREADS / 32; ++w) t += sh[w];
      carry = t;
    }
    __syncthreads();
  }
}

extern "C" int fvx_run_ids(const int32_t* user, int64_t n, int32_t* run_id, int32_t* scratch, fvx_stream_t stream) {
  FVX_CHECK_ARG(user && run_id && scratch && n >= 0, "fvx_run_ids: bad arguments");
  if (n == 0) return 0;
  const long long nb = (n + RI_THREADS * RI_PER - 1) / (RI_THREADS * RI_PER);
  FVX_CHECK_ARG(nb <= 4096, "fvx_run_ids: n=%lld too large", (long long)n);
  cudaStream_t st = fvx_cu(stream);
  k_run_count<<<(int)nb, RI_THREADS, 0, st>>>(user, n, scratch);
  k_run_write<<<(int)nb, RI_THREADS, 0, st>>>(user, n, scratch, run_id);
  FVX_CHECK_LAUNCH("k_run_ids");
  return 0;
}

static int sharded_common(const FvxModel* m, const int32_t* user, int B, const char* who) {
  if (int rc = fvx_check_model(m, who)) return rc;
  FVX_CHECK_ARG(user != nullptr && B >= 1 && B <= m->max_batch, "%s: bad batch", who);
  FVX_CHECK_ARG(m->users.list_cap >= B && m->items.list_cap >= 2 * B, "%s: touched-row lists too small", who);
  FVX_CHECK_ARG(m->rows && m->loss && m->sync && m->cmap, "%s: null scratch (rows / loss / sync / cmap)", who);
  if (m->D > 0) {
    FVX_CHECK_ARG(m->TH && m->gE_part && m->ge_parts > 0, "%s: VBPR scratch missing", who);
    if (m->use_tensor_cores) FVX_CHECK_ARG(m->F_pl && m->ET_hi && m->ET_lo && m->W_hi && m->W_lo, "%s: bf16 planes missing", who);
    else FVX_CHECK_ARG(m->F && m->W, "%s: fp32 projection needs F and W", who);
  }
  return 0;
}

static int sharded_ks(const FvxModel* m, int B) {
  if (!(m->D > 0 && m->use_tensor_cores)) return 1;
  const int NP = fvx_tc_np(m->de);
  int ks = fvx_tc_ksplit(m, 2LL * B);
  while (ks > 1 && (long long)ks * 2 * B * NP > m->th_cap) ks >>= 1;
  return ks;
}

static SsTheta make_theta(const FvxModel* m, int B, int ks) {
  SsTheta T;
  const bool tc = m->D > 0 && m->use_tensor_cores;
  T.p = m->TH;
  T.np = tc ? fvx_tc_np(m->de) : m->de;
  T.ks = tc ? ks : 1;
  T.ss = 2LL * B * T.np;
  T.chunks = m->D > 0 ? m->D / 64 : 1;
  T.nsm = fvx_num_sms();
  return T;
}

// Unique-row step on a shard (see fvx_train.cu): with R ranks a rank owns I/R catalog rows but 2B_global/R
// slots - at 8 ranks ~131 k slots over 12.5 k rows - so projecting each DISTINCT owned row once shrinks both
// contractions by the duplication factor.  Needs the tensor-core path, K % 4 == 0 and the scratch
// (upos, W_sum, uslot); FVX_STEP_DEDUP=0 turns it off.
static bool sharded_uniq(const FvxModel* m) {
  return m->D > 0 && m->use_tensor_cores && m->upos && m->W_sum && m->uslot && m->K % 4 == 0 &&
         fvx_tc_np(m->de) <= 256 && fvx_dedup_enabled();
}
static int uniq_ks_cap(const FvxModel* m, int B) {
  int ks_cap = 8;
  while (ks_cap > 1 && (long long)ks_cap * 2 * B * fvx_tc_np(m->de) > m->th_cap) ks_cap >>= 1;
  return ks_cap;
}

static inline int ss_grid(long long warps_needed) {
  long long g = (warps_needed + SS_WARPS - 1) / SS_WARPS;
  const long long cap = (long long)fvx_num_sms() * 8;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

extern "C" {

int fvx_bpr_step_sharded_a(const FvxModel* model, const int32_t* user, const int32_t* pos, const int32_t* neg,
                           int32_t B, float* S, fvx_stream_t stream) {
  if (int rc = sharded_common(model, user, B, "fvx_bpr_step_sharded_a")) return rc;
  FVX_CHECK_ARG(pos && neg && S, "fvx_bpr_step_sharded_a: null pointer");
  const FvxModel& M = *model;
  cudaStream_t st = fvx_cu(stream);
  // The projection needs only the slot rows and the planes of E_ext^T: the claims and the
  // deferred-Adam catch-up of the touched rows run beside it on the side stream (joined before the
  // partial scores read the tables).
  const bool uniq = sharded_uniq(&M);
  cudaStream_t side = nullptr;
  if (uniq) {
    // slot rows + claims of the owned rows (list positions in upos) ahead of the projection; user claims,
    // catch-up and the slots' list positions (uslot) beside it
    if (int rc = fvx_launch_prep(&M, user, pos, neg, B, st, FVX_PREP_UNIQ)) return rc;
    side = fvx_side_begin(st);
    if (int rc = fvx_launch_prep(&M, user, pos, neg, B, side ? side : st, FVX_PREP_CLAIMS_LISTED)) return rc;
  } else {
    side = M.D > 0 ? fvx_side_begin(st) : nullptr;
    if (side) {
      if (int rc = fvx_launch_prep(&M, user, pos, neg, B, side, FVX_PREP_CLAIMS)) return rc;
      if (int rc = fvx_launch_prep(&M, user, pos, neg, B, st, FVX_PREP_ROWS)) return rc;
    } else {
      if (int rc = fvx_launch_prep(&M, user, pos, neg, B, st)) return rc;
    }
  }
  // compact list of the owned slots; foreign entries of crow stay -1 (the fp32 kernels skip them)
  int32_t* count = M.sync + 1;
  cudaMemsetAsync(M.cmap, 0xFF, sizeof(int32_t) * 2 * (size_t)M.max_batch, st);
  cudaMemsetAsync(S, 0, sizeof(float) * 2 * (size_t)B, st);
  {
    long long g = (2LL * B + 255) / 256;
    if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
    k_compact_owned<<<(int)g, 256, 0, st>>>(M, B, count);
    FVX_CHECK_LAUNCH("k_compact_owned");
  }
  const int ks = uniq ? uniq_ks_cap(&M, B) : sharded_ks(&M, B);
  if (uniq) {
    FVX_CHECK_ARG(2LL * B * fvx_tc_np(M.de) <= M.th_cap, "fvx_bpr_step_sharded_a: TH scratch too small");
    if (int rc = fvx_launch_project_tc(&M, M.items.list, 0, 2 * B, ks, M.TH, st, M.items.count, 1)) return rc;
  } else if (M.D > 0) {
    if (M.use_tensor_cores) {
      FVX_CHECK_ARG((long long)ks * 2 * B * fvx_tc_np(M.de) <= M.th_cap, "fvx_bpr_step_sharded_a: TH scratch too small");
      // (the W rows past the owned ones, read by the last backward tile, are zeroed by phase B)
      if (int rc = fvx_launch_project_tc(&M, M.cmap, 0, 2 * B, ks, M.TH, st, count)) return rc;
    } else {
      FVX_CHECK_ARG(2LL * B * M.de <= M.th_cap, "fvx_bpr_step_sharded_a: TH scratch too small");
      if (int rc = fvx_launch_project(&M, M.cmap, 2 * B, M.TH, st)) return rc;
    }
  }
  if (side) fvx_side_join(st);
  if (M.K % 4 == 0)
    k_partial_scores_v4<<<ss_grid(B), SS_WARPS * 32, 0, st>>>(M, user, B, make_theta(&M, B, ks), S, count, uniq ? 1 : 0);
  else
    k_partial_scores<<<ss_grid(2LL * B), SS_WARPS * 32, 0, st>>>(M, user, B, make_theta(&M, B, ks), S, count);
  FVX_CHECK_LAUNCH("k_partial_scores");
  return 0;
}

int fvx_bpr_step_sharded_b1(const FvxModel* model, const int32_t* user, int32_t B, const float* S,
                            const int32_t* run_id, float* RU, int64_t ru_rows, int32_t loss_slot,
                            fvx_stream_t stream) {
  if (int rc = sharded_common(model, user, B, "fvx_bpr_step_sharded_b")) return rc;
  FVX_CHECK_ARG(S && run_id && RU && ru_rows >= 1, "fvx_bpr_step_sharded_b: null pointer");
  FVX_CHECK_ARG(loss_slot >= 0 && loss_slot < model->loss_slots, "fvx_bpr_step_sharded_b: loss_slot out of range");
  const FvxModel& M = *model;
  cudaStream_t st = fvx_cu(stream);
  if (cudaMemsetAsync(RU, 0, sizeof(float) * ru_rows * M.users.stride, st) != cudaSuccess)
    FVX_FAIL(-3, "fvx_bpr_step_sharded_b: memset failed");
  const bool uniq = sharded_uniq(&M);
  const int ks = uniq ? uniq_ks_cap(&M, B) : sharded_ks(&M, B);
  const bool tc = M.D > 0 && M.use_tensor_cores;
  if (M.K % 4 == 0) {
    // one warp per owned slot: the work does not grow with the number of ranks
    k_grads_owned<<<ss_grid(B), SS_WARPS * 32, 0, st>>>(M, user, B, loss_slot, make_theta(&M, B, ks),
                                                              tc ? fvx_tc_np(M.de) : 0, tc ? fvx_w_pitch(&M) : 0, S,
                                                              run_id, RU, (long long)ru_rows, M.sync + 1, uniq ? 1 : 0);
    FVX_CHECK_LAUNCH("k_grads_owned");
    if (uniq) {
      if (int rc = fvx_launch_w_planes(&M, B, st)) return rc;
    }
  } else {
    if (tc) {   // W rows past the owned ones must read as zero in the last backward tile
      const int np = fvx_tc_np(M.de), pitch = fvx_w_pitch(&M);
      cudaMemsetAsync(M.W_hi, 0, sizeof(uint16_t) * 2 * (size_t)B * pitch, st);
      if (pitch == np) cudaMemsetAsync(M.W_lo, 0, sizeof(uint16_t) * 2 * (size_t)B * np, st);
    }
    k_grads_sharded<<<ss_grid(B), SS_WARPS * 32, 0, st>>>(M, user, B, loss_slot, make_theta(&M, B, ks),
                                                          tc ? fvx_tc_np(M.de) : 0, tc ? fvx_w_pitch(&M) : 0, S, run_id,
                                                          RU, (long long)ru_rows);
    FVX_CHECK_LAUNCH("k_grads_sharded");
  }
  return 0;
}

int fvx_bpr_step_sharded_b2(const FvxModel* model, int32_t B, float* dE, fvx_stream_t stream) {
  FVX_CHECK_ARG(model != nullptr, "fvx_bpr_step_sharded_b: null model");
  const FvxModel& M = *model;
  if (M.D == 0) return 0;
  FVX_CHECK_ARG(dE != nullptr && B >= 1 && B <= M.max_batch, "fvx_bpr_step_sharded_b: VBPR needs the dE buffer");
  cudaStream_t st = fvx_cu(stream);
  const bool tc = M.use_tensor_cores;
  int parts = 0;
  if (tc && sharded_uniq(&M)) {
    if (int rc = fvx_launch_grad_E_tc(&M, M.items.list, 2 * B, &parts, st, M.items.count)) return rc;
  } else if (tc) {
    if (int rc = fvx_launch_grad_E_tc(&M, M.cmap, 2 * B, &parts, st, M.sync + 1)) return rc;
  } else {
    if (int rc = fvx_launch_grad_E(&M, M.cmap, 2 * B, &parts, st)) return rc;
  }
  return fvx_launch_reduce_gE(&M, parts, tc ? fvx_tc_np(M.de) : M.de, dE, st);
}

int fvx_bpr_step_sharded_b(const FvxModel* model, const int32_t* user, int32_t B, const float* S,
                           const int32_t* run_id, float* RU, int64_t ru_rows, float* dE, int32_t loss_slot,
                           fvx_stream_t stream) {
  FVX_CHECK_ARG(model != nullptr && (model->D == 0 || dE != nullptr), "fvx_bpr_step_sharded_b: VBPR needs the dE buffer");
  if (int rc = fvx_bpr_step_sharded_b1(model, user, B, S, run_id, RU, ru_rows, loss_slot, stream)) return rc;
  return fvx_bpr_step_sharded_b2(model, B, dE, stream);
}

int fvx_bpr_step_sharded_c(const FvxModel* model, const int32_t* user, int32_t B, const int32_t* run_id,
                           const float* RU, int64_t ru_rows, const float* dE, int32_t loss_slot,
                           fvx_stream_t stream) {
  if (int rc = sharded_common(model, user, B, "fvx_bpr_step_sharded_c")) return rc;
  FVX_CHECK_ARG(run_id && RU, "fvx_bpr_step_sharded_c: null pointer");
  const FvxModel& M = *model;
  FVX_CHECK_ARG(M.D == 0 || dE != nullptr, "fvx_bpr_step_sharded_c: VBPR needs the reduced dE");
  cudaStream_t st = fvx_cu(stream);
  long long g = ((long long)B * 32 + 255) / 256;
  if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
  k_scatter_runs<<<(int)g, 256, 0, st>>>(M, user, B, run_id, RU, (long long)ru_rows);
  FVX_CHECK_LAUNCH("k_scatter_runs");
  // loss_slot < 0: the E term of the loss (VBPR.py:129) is not added (ranks other than 0, so that
  // the per-rank losses sum to the batch loss)
  // DEFERRED: the touched rows keep their gradient and take the step when they are next needed (replay_row)
  return fvx_launch_update(&M, B, M.D > 0 ? 1 : 0, M.de, dE, loss_slot, st,
                           fvx_merged_update(&M) ? FVX_UPD_E : FVX_UPD_ALL);
}

}  // extern "C"
