// The BPR optimiser step with the model sharded over R ranks (one process per GPU): include/fvx.h,
// "the same step with the model sharded over R ranks".
//
// x_uij = s_ui - s_uj is linear in the item-side quantities, s_ui = Bi[i] + <Gu[u],Gi[i]> + <Tu[u], F[i]E> +
// F[i]Bp (BPRMF.py:74, VBPR.py:82-84), so each rank scores the (triple, side) slots whose item it owns and one
// small all-reduce assembles every x.  Items (Gi / Bi / F + Adam state) live on their owner; a USER's Adam state
// lives on the owner of the user, who brings the row up to date when a batch touches it and publishes it to the
// other ranks for that step (WU, indexed by run of equal users); E is replicated.
//
//   piece 1  run ids; slot rows + claims of the owned item rows (unique-row step), E planes         [main]
//   piece 2  claims + catch-up of the OWNED users of the batch, catch-up of the listed item rows;   [side]
//            fresh rows of the owned users -> WU                      -- all-reduce(WU) [side] --
//   piece 3  compact list of the owned slots; projection of the distinct owned rows                 [main]
//   piece 4  partial scores of the owned slots (user rows from WU) -> S   -- all-reduce(S) --       [main]
//   piece 5  gradients of the owned slots: item rows, user shares -> RU, coefficient sums -> planes [main]
//                                                                     -- all-reduce(RU) [side] --
//   piece 6  grad_E over the distinct owned rows -> dE (+ loss share) -- all-reduce(dE) --          [main]
//   piece 7  RU rows of the OWNED users -> their gradient accumulators                              [side]
//   piece 8  Adam on E (identical on every rank), loss, step += 1                                   [main]
//
// RU / WU are indexed by RUN: the reference's sampler emits runs of one user (dataset.py:96-99), run_id[b]
// = number of positions <= b where user[b] != user[b-1], minus one.  The layout depends only on the
// batch, so it is the same on every rank and the all-reduces need no index exchange.  Ownership of
// the per-triple terms that are not tied to an item (softplus loss, user-side L2): the rank that
// owns the POSITIVE item.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "fvx_comm.cuh"
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
                    const int32_t* __restrict__ count, int uniq, const float* __restrict__ WU,
                    const int32_t* __restrict__ run_id, long long ru_rows) {
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
      // the user's row: published by its owner for this step (WU, by run), or the local table (one rank)
      const float* urow = M.users.w + (size_t)user[b] * Su;
      if (WU) {
        const int32_t run = run_id[b];
        urow = WU + (size_t)(run < ru_rows ? run : 0) * Su;       // (a run past the buffer: the step is poisoned anyway)
      }
      const float4* ur = reinterpret_cast<const float4*>(urow);
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

#define GH_HOT 4             // hot rows per block
#define GH_W 592             // floats per hot row: item row (<= 260) + coefficient sums (<= 320), 9.3 KB for four
__global__ void __launch_bounds__(SS_WARPS * 32, 4)
k_grads_owned(FvxModel M, const int32_t* __restrict__ user, int B, int loss_slot, SsTheta T, int wnp, int wpitch,
              const float* __restrict__ S, const int32_t* __restrict__ run_id, float* __restrict__ RU,
              long long ru_rows, const int32_t* __restrict__ count, int uniq, const float* __restrict__ WU,
              double* __restrict__ loss_out) {
  __shared__ double loss_sh[SS_WARPS * 2];
  // Hot rows.  With a popularity law the most popular item of a shard draws a large share of the shard's slots
  // (Zipf(1.0), 8 ranks, 524 k triples: 37 k of the owner's 131 k slots), and every one of them is a chain of
  // red.add onto the SAME accumulator row - L2 serialises them, and the rank that owns the item holds up every
  // exchange of the step (57 ... 256 us across the ranks, profiles/r2_sharded_emul8_grads.txt).  Each block
  // therefore looks at a sample of 64 of its own slots, takes the rows that appear at least three times as its
  // hot rows (at most GH_HOT), sums their contributions in shared memory and adds them to the global
  // accumulators once at the end: a few hundred reductions per address instead of tens of thousands.
  __shared__ int32_t hot_row[GH_HOT], hot_tj[GH_HOT];
  __shared__ float hot_acc[GH_HOT][GH_W];
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
  const int hot_w = Si + ((uniq && vis) ? wnp : 0);
  const bool hot_on = hot_w <= GH_W;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    // this block's first four passes of the slot loop: 64 slots, two per lane
    int32_t a[2];
    int c[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long j = (long long)blockIdx.x * (SS_WARPS * 2) + (lane & 15) + (long long)(2 * (lane >> 4) + h) * ng;
      a[h] = (hot_on && j < n_owned) ? crow[j] : -2 - 2 * lane - h;           // (no slot: a value nobody shares)
    }
    for (int it = 0; it < 32; ++it) {
      const int32_t b0 = __shfl_sync(0xffffffffu, a[0], it), b1 = __shfl_sync(0xffffffffu, a[1], it);
      c[0] += (a[0] == b0) + (a[0] == b1);
      c[1] += (a[1] == b0) + (a[1] == b1);
    }
    for (int h = 0; h < GH_HOT; ++h) {
      // the most frequent row still in the sample (count in the high word, row in the low one)
      long long best = -1;
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (c[q] >= 3 && a[q] >= 0) { const long long key = ((long long)c[q] << 32) | (uint32_t)a[q]; best = key > best ? key : best; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const long long y = __shfl_xor_sync(0xffffffffu, best, o); best = y > best ? y : best; }
      const int32_t row = best >= 0 ? (int32_t)(best & 0xFFFFFFFFLL) : -1;
      if (lane == 0) {
        hot_row[h] = row;
        hot_tj[h] = (row >= 0 && uniq && vis) ? M.upos[row] : 0;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) if (a[q] == row) c[q] = 0;
    }
  }
  for (int i = threadIdx.x; i < GH_HOT * GH_W; i += blockDim.x) (&hot_acc[0][0])[i] = 0.0f;
  __syncthreads();
  int32_t hr[GH_HOT];
#pragma unroll
  for (int h = 0; h < GH_HOT; ++h) hr[h] = hot_row[h];
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
        const float4* ur = reinterpret_cast<const float4*>(WU ? WU + (size_t)run * Su : M.users.w + (size_t)u * Su);
        const float4* gi = reinterpret_cast<const float4*>(M.items.w + (size_t)li * Si);
        float* gg = M.items.g + (size_t)li * Si;
        float* ru = RU + (size_t)run * Su;
        const long long tj = uniq ? (long long)M.uslot[slot] : j;   // row of TH / of the coefficient sums
        int hot = -1;
#pragma unroll
        for (int h = 0; h < GH_HOT; ++h) if (li == hr[h]) hot = h;
        float* hacc = hot >= 0 ? hot_acc[hot] : nullptr;
        for (int c = sub; c < K4; c += 16) {
          const float4 a = ur[c], x = gi[c];
          const float4 gv = make_float4(cs * a.x + reg2 * x.x, cs * a.y + reg2 * x.y, cs * a.z + reg2 * x.z,
                                        cs * a.w + reg2 * x.w);
          if (hacc) {
            atomicAdd(hacc + 4 * c, gv.x); atomicAdd(hacc + 4 * c + 1, gv.y);
            atomicAdd(hacc + 4 * c + 2, gv.z); atomicAdd(hacc + 4 * c + 3, gv.w);
          } else {
            ss_red_add4(gg + 4 * c, gv);
          }
          ss_red_add4(ru + 4 * c, make_float4(cs * x.x + ul2 * a.x, cs * x.y + ul2 * a.y, cs * x.z + ul2 * a.z,
                                              cs * x.w + ul2 * a.w));
          sq += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
          if (!side) sq += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
        bi = M.items.w[(size_t)li * Si + K];
        if (sub == 0) {
          const float gb = cs + (side == 0 ? reg2 : reg2 * M.bias_neg_scale) * bi;
          if (hacc) atomicAdd(hacc + K, gb); else fvx_red_add(gg + K, gb);
        }
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
              if (n0 <= d) {
                if (hacc) {
                  float* hw = hacc + Si + n0;
                  atomicAdd(hw, wv.x); atomicAdd(hw + 1, wv.y); atomicAdd(hw + 2, wv.z); atomicAdd(hw + 3, wv.w);
                } else {
                  ss_red_add4(M.W_sum + (size_t)tj * wnp + n0, wv);
                }
              }
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
      loss_acc += (double)(reg * sqs) + (double)(reg * bi * bi * (side == 0 ? 1.0f : M.bias_neg_scale));
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
    if (s != 0.0) atomicAdd(loss_out, s);
  }
  // the block's sums of its hot rows -> the global accumulators
  for (int h = 0; h < GH_HOT; ++h) {
    const int32_t row = hot_row[h];
    if (row < 0) continue;
    float* gg = M.items.g + (size_t)row * Si;
    float* ws = (uniq && vis) ? M.W_sum + (size_t)hot_tj[h] * wnp : nullptr;
    for (int i = threadIdx.x; i < hot_w; i += blockDim.x) {
      const float v = hot_acc[h][i];
      if (v != 0.0f) fvx_red_add(i < Si ? gg + i : ws + (i - Si), v);
    }
  }
}

// one warp per run start of an OWNED user adds the all-reduced run gradient into the user's accumulator
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
      start = (b == 0 || user[b - 1] != u) && run < ru_rows && u >= M.user_lo && u < M.user_lo + M.user_cnt;
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

// The owner of a user publishes the user's up-to-date row for this step: WU[run] = w[u] for the run starts of
// OWNED users (k_prep has caught the row up) - they all lie in the owner's segment of WU, which is then gathered.
__global__ void k_pack_wu(FvxModel M, const int32_t* __restrict__ user, int B, const int32_t* __restrict__ run_id,
                          float* __restrict__ WU, long long ru_rows) {
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
      start = (b == 0 || user[b - 1] != u) && u >= M.user_lo && u < M.user_lo + M.user_cnt;
      if (start && run >= ru_rows) { M.sync[2] = 1; start = false; }    // more runs than the buffers hold
    }
    uint32_t msk = __ballot_sync(0xffffffffu, start);
    while (msk) {
      const int src = __ffs(msk) - 1;
      msk &= msk - 1;
      const int32_t uu = __shfl_sync(0xffffffffu, u, src);
      const int32_t rr = __shfl_sync(0xffffffffu, run, src);
      const float4* s4 = reinterpret_cast<const float4*>(M.users.w + (size_t)uu * Su);
      float4* d4 = reinterpret_cast<float4*>(WU + (size_t)rr * Su);
      for (int c = lane; c < S4; c += 32) d4[c] = s4[c];
    }
  }
}

// dE[D, de] = sum of the row-group partials; behind it this rank's loss share as two floats (hi + lo) and the
// run-overflow flag, so that one all-reduce carries all three
__global__ void k_dE_pack(const float* __restrict__ part, int parts, int D, int gnp, int de, float* __restrict__ out,
                          double* __restrict__ loss_part, int32_t* __restrict__ sync) {
  const int n = D * de;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int f = i / de, c = i - f * de;
    const float* gp = part + (size_t)f * gnp + c;
    const size_t ps = (size_t)D * gnp;
    float g = 0.0f;
    int p = 0;
    for (; p + 8 <= parts; p += 8) {             // eight loads in flight, summed in a fixed order
      float t_[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t_[q] = gp[(size_t)(p + q) * ps];
#pragma unroll
      for (int q = 0; q < 8; ++q) g += t_[q];
    }
    for (; p < parts; ++p) g += gp[(size_t)p * ps];
    out[i] = g;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double l = *loss_part;
    const float hi = (float)l;
    out[n] = hi;
    out[n + 1] = (float)(l - (double)hi);
    out[n + 2] = sync[2] ? 1.0f : 0.0f;
    out[n + 3] = 0.0f;
    sync[2] = 0;
    *loss_part = 0.0;
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

// Run SLOTS: the same runs, but numbered per OWNER of the run's user: slot = owner * cap + (index of the run among
// the runs of that owner, in batch order).  The rows of the exchanged buffers (WU, RU) are then grouped in one
// contiguous segment of `cap` rows per owner, so the fresh user rows travel as an all-gather and the gradient
// shares as a reduce-scatter - half the bytes of the all-reduces a run-ordered layout needs.  A run whose index
// reaches cap gets slot 0x7fffffff (the kernels skip it and flag the step).  Up to 8 owners: the per-owner counts
// of a 256-element round are scanned as eight 16-bit fields of two 64-bit words.
#define RS_OWNERS 8
#define RS_PER 4          // rounds of 256 elements per block: 1024 elements, so that a batch of 131 k triples fills the machine
__device__ __forceinline__ int rs_owner(int32_t u, int per, int R) {
  int o = per > 0 ? u / per : 0;
  return o < 0 ? 0 : (o >= R ? R - 1 : o);
}
__global__ void __launch_bounds__(RI_THREADS)
k_slot_count(const int32_t* __restrict__ user, long long n, int per, int R, int32_t* __restrict__ part) {
  __shared__ int sh[RS_OWNERS];
  if (threadIdx.x < RS_OWNERS) sh[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * RI_THREADS * RS_PER;
  for (int e = 0; e < RS_PER; ++e) {
    const long long b = base + (long long)e * RI_THREADS + threadIdx.x;
    if (b < n && (b == 0 || user[b] != user[b - 1])) atomicAdd(&sh[rs_owner(user[b], per, R)], 1);
  }
  __syncthreads();
  if (threadIdx.x < RS_OWNERS) part[blockIdx.x * RS_OWNERS + threadIdx.x] = sh[threadIdx.x];
}
__global__ void __launch_bounds__(RI_THREADS)
k_slot_write(const int32_t* __restrict__ user, long long n, int per, int R, int cap, const int32_t* __restrict__ part,
             int32_t* __restrict__ run_id, int32_t* __restrict__ totals) {
  __shared__ unsigned long long sh[RI_THREADS / 32][2];
  __shared__ int carry[RS_OWNERS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < RS_OWNERS) carry[threadIdx.x] = 0;
  __syncthreads();
  {                                   // runs of each owner in the blocks before this one: 32 threads per owner
    const int o = threadIdx.x >> 5;
    int t = 0;
    for (int q = lane; q < (int)blockIdx.x; q += 32) t += part[q * RS_OWNERS + o];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    if (lane == 0) carry[o] = t;
  }
  __syncthreads();
  const long long base = (long long)blockIdx.x * RI_THREADS * RS_PER;
  for (int e = 0; e < RS_PER; ++e) {                      // 256 consecutive elements per round
    const long long b = base + (long long)e * RI_THREADS + threadIdx.x;
    int o = 0;
    unsigned long long x0 = 0ull, x1 = 0ull;               // this element's contribution: a start of owner o
    if (b < n) {
      o = rs_owner(user[b], per, R);
      if (b == 0 || user[b] != user[b - 1]) {
        if (o < 4) x0 = 1ull << (16 * o); else x1 = 1ull << (16 * (o - 4));
      }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {                     // inclusive scan of the round's eight counters
      const unsigned long long y0 = __shfl_up_sync(0xffffffffu, x0, d), y1 = __shfl_up_sync(0xffffffffu, x1, d);
      if (lane >= d) { x0 += y0; x1 += y1; }
    }
    if (lane == 31) { sh[warp][0] = x0; sh[warp][1] = x1; }
    __syncthreads();
    unsigned long long off0 = 0ull, off1 = 0ull, tot0 = 0ull, tot1 = 0ull;
    for (int w = 0; w < RI_THREADS / 32; ++w) {
      if (w < warp) { off0 += sh[w][0]; off1 += sh[w][1]; }
      tot0 += sh[w][0]; tot1 += sh[w][1];
    }
    if (b < n) {
      const unsigned long long w_ = o < 4 ? (x0 + off0) >> (16 * o) : (x1 + off1) >> (16 * (o - 4));
      const int idx = carry[o] + (int)(w_ & 0xFFFFull) - 1;   // runs of owner o that start at or before b, minus one
      run_id[b] = idx < cap ? o * cap + idx : 0x7fffffff;
    }
    __syncthreads();
    if (threadIdx.x < RS_OWNERS) {
      const unsigned long long t_ = threadIdx.x < 4 ? tot0 >> (16 * threadIdx.x) : tot1 >> (16 * (threadIdx.x - 4));
      carry[threadIdx.x] += (int)(t_ & 0xFFFFull);
    }
    __syncthreads();
  }
  // runs of every owner in the whole batch (capped at cap): what the peer-to-peer exchange moves per segment
  if (totals && blockIdx.x == gridDim.x - 1 && threadIdx.x < RS_OWNERS) totals[threadIdx.x] = min(carry[threadIdx.x], cap);
}

static int run_slots_impl(const int32_t* user, int64_t n, int32_t users_per_owner, int32_t owners, int32_t cap,
                          int32_t* run_slot, int32_t* scratch, int32_t* totals, fvx_stream_t stream);
extern "C" int fvx_run_slots(const int32_t* user, int64_t n, int32_t users_per_owner, int32_t owners, int32_t cap,
                             int32_t* run_slot, int32_t* scratch, fvx_stream_t stream) {
  return run_slots_impl(user, n, users_per_owner, owners, cap, run_slot, scratch, nullptr, stream);
}
static int run_slots_impl(const int32_t* user, int64_t n, int32_t users_per_owner, int32_t owners, int32_t cap,
                          int32_t* run_slot, int32_t* scratch, int32_t* totals, fvx_stream_t stream) {
  FVX_CHECK_ARG(user && run_slot && scratch && n >= 0, "fvx_run_slots: bad arguments");
  FVX_CHECK_ARG(owners >= 1 && owners <= RS_OWNERS && users_per_owner >= 1 && cap >= 1, "fvx_run_slots: owners=%d outside [1, %d]",
                owners, RS_OWNERS);
  if (n == 0) return 0;
  const long long nb = (n + RI_THREADS * RS_PER - 1) / (RI_THREADS * RS_PER);
  FVX_CHECK_ARG(nb <= 8192, "fvx_run_slots: n=%lld too large", (long long)n);
  cudaStream_t st = fvx_cu(stream);
  k_slot_count<<<(int)nb, RI_THREADS, 0, st>>>(user, n, users_per_owner, owners, scratch);
  k_slot_write<<<(int)nb, RI_THREADS, 0, st>>>(user, n, users_per_owner, owners, cap, scratch, run_slot, totals);
  FVX_CHECK_LAUNCH("k_run_slots");
  return 0;
}

// ---- peer-to-peer exchange (FvxComm arena: every rank's exchange buffers are mapped into every other rank) ----
// The NCCL collectives of the step are latency-bound at these sizes and slow down beside the tensor-core kernels
// (8 GPUs: all-gather of the user rows 220 us, all-reduce of S 128 us, reduce-scatter of the gradient shares 273 us
// per step - profiles/r2_sharded_*).  Here the kernels that PRODUCE the data store it straight into the consumers'
// memory over NVLink, and a one-warp barrier kernel separates producers from consumers.
struct PeerPtrs {
  uint8_t* p[FVX_COMM_MAX_RANKS];      // p[r]: rank r's arena as mapped here
  uint8_t* self;                       // this rank's arena
  int rank, world;
};
template <typename T>
__device__ __forceinline__ T* peer_of(const PeerPtrs& X, int r, T* local) {
  return reinterpret_cast<T*>(X.p[r] + (reinterpret_cast<uint8_t*>(local) - X.self));
}

// Cross-GPU barrier: every rank writes `epoch` into its slot of every peer's flag array, then waits until all slots
// of its own array have reached it.  One warp; the kernel boundary before it has made the producer kernel's peer
// stores visible (plus the fence below).  A peer that never arrives (a crashed rank, a protocol error) traps this
// rank after ~30 s instead of hanging the box; ranks that are merely late (host work between steps) are waited for.
__global__ void k_xbar(PeerPtrs X, uint32_t* flags, uint32_t epoch) {
  const int lane = threadIdx.x;
  __threadfence_system();
  if (lane < X.world) {
    volatile uint32_t* dst = peer_of(X, lane, flags) + X.rank;
    *dst = epoch;
  }
  __threadfence_system();
  if (lane < X.world) {
    volatile uint32_t* mine = flags + lane;
    const long long t0 = clock64();
    while ((int32_t)(*mine - epoch) < 0) {
      if (clock64() - t0 > 60000000000LL) __trap();
    }
  }
  __threadfence_system();
}

// owner of a catalog item under shard_bounds(): the first `rem` shards hold base + 1 rows
__device__ __forceinline__ int item_owner(int32_t i, int base, int rem) {
  const int cut = rem * (base + 1);
  return i < cut ? i / (base + 1) : rem + (i - cut) / (base > 0 ? base : 1);
}

// k_pack_wu, peer-to-peer: the fresh row of an owned run goes into EVERY rank's WU (its own included); the user of
// the run is kept for the gradient scatter (run_user[index in this rank's segment])
__global__ void k_pack_wu_p2p(FvxModel M, const int32_t* __restrict__ user, int B, const int32_t* __restrict__ run_id,
                              float* __restrict__ WU, long long ru_rows, int32_t* __restrict__ run_user, int cap, PeerPtrs X) {
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
      start = (b == 0 || user[b - 1] != u) && u >= M.user_lo && u < M.user_lo + M.user_cnt;
      if (start && run >= ru_rows) { M.sync[2] = 1; start = false; }    // more runs than the buffers hold
    }
    uint32_t msk = __ballot_sync(0xffffffffu, start);
    while (msk) {
      const int src = __ffs(msk) - 1;
      msk &= msk - 1;
      const int32_t uu = __shfl_sync(0xffffffffu, u, src);
      const int32_t rr = __shfl_sync(0xffffffffu, run, src);
      const float4* s4 = reinterpret_cast<const float4*>(M.users.w + (size_t)uu * Su);
      if (lane == 0) run_user[rr - X.rank * cap] = uu;
      for (int c = lane; c < S4; c += 32) {
        const float4 v = s4[c];
        for (int r = 0; r < X.world; ++r) reinterpret_cast<float4*>(peer_of(X, r, WU) + (size_t)rr * Su)[c] = v;
      }
    }
  }
}

// After k_partial_scores_v4 (which fills this rank's S): the partial score of an owned slot also goes to the one
// other rank that needs it - the owner of the triple's OTHER item (x_b = S[b] - S[B+b] is formed by the owners of
// the two sides only).
__global__ void k_send_scores(FvxModel M, const int32_t* __restrict__ pos, const int32_t* __restrict__ neg, int B,
                              float* __restrict__ S, const int32_t* __restrict__ count, int ibase, int irem, PeerPtrs X) {
  const int32_t* cslot = M.cmap + 2 * (size_t)M.max_batch;
  const long long n_owned = *count;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_owned; j += (long long)gridDim.x * blockDim.x) {
    const int32_t slot = cslot[j];
    const int32_t other = slot < B ? neg[slot] : pos[slot - B];
    if (other < 0 || other >= M.num_items) continue;
    const int o = item_owner(other, ibase, irem);
    if (o != X.rank) peer_of(X, o, S)[slot] = S[slot];
  }
}

// The gradient shares of the runs of owner o (segment o of this rank's RU) go into owner o's RUin[this rank]
__global__ void k_push_ru(const float* __restrict__ RU, float* __restrict__ RUin, const int32_t* __restrict__ run_counts,
                          int cap, int S4, PeerPtrs X) {
  for (int o = 0; o < X.world; ++o) {
    const int n = run_counts[o];
    const float4* src = reinterpret_cast<const float4*>(RU) + (size_t)o * cap * S4;
    float4* dst = reinterpret_cast<float4*>(peer_of(X, o, RUin)) + (size_t)X.rank * cap * S4;
    const long long total = (long long)n * S4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
      dst[i] = src[i];
  }
}

// The owner sums the shares of every source rank (fixed order: deterministic) and adds the run into the user's
// accumulator (a user may own two runs of a batch)
__global__ void k_scatter_runs_p2p(FvxModel M, const float* __restrict__ RUin, const int32_t* __restrict__ run_user,
                                   const int32_t* __restrict__ run_counts, int cap, PeerPtrs X) {
  const int lane = threadIdx.x & 31;
  const int Su = M.users.stride, S4 = Su >> 2;
  const int n = run_counts[X.rank];
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += nw) {
    const int32_t u = run_user[k];
    float* g = M.users.g + (size_t)u * Su;
    for (int c = lane; c < S4; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < X.world; ++r) {
        const float4 v = reinterpret_cast<const float4*>(RUin + ((size_t)r * cap + k) * Su)[c];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      ss_red_add4(g + 4 * c, acc);
    }
  }
}

// k_dE_pack, peer-to-peer: this rank's dE and tail go into slot `rank` of every rank's dEall / tails
__global__ void k_dE_pack_p2p(const float* __restrict__ part, int parts, int D, int gnp, int de, float* __restrict__ dEall,
                              float* __restrict__ tails, double* __restrict__ loss_part, int32_t* __restrict__ sync, PeerPtrs X) {
  const int n = D * de;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int f = i / de, c = i - f * de;
    const float* gp = part + (size_t)f * gnp + c;
    const size_t ps = (size_t)D * gnp;
    float g = 0.0f;
    int p = 0;
    for (; p + 8 <= parts; p += 8) {
      float t_[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t_[q] = gp[(size_t)(p + q) * ps];
#pragma unroll
      for (int q = 0; q < 8; ++q) g += t_[q];
    }
    for (; p < parts; ++p) g += gp[(size_t)p * ps];
    for (int r = 0; r < X.world; ++r) peer_of(X, r, dEall)[(size_t)X.rank * n + i] = g;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double l = *loss_part;
    const float hi = (float)l;
    const float lo = (float)(l - (double)hi), fl = sync[2] ? 1.0f : 0.0f;
    for (int r = 0; r < X.world; ++r) {
      float* t = peer_of(X, r, tails) + 4 * X.rank;
      t[0] = hi; t[1] = lo; t[2] = fl; t[3] = 0.0f;
    }
    sync[2] = 0;
    *loss_part = 0.0;
  }
}

static int sharded_common(const FvxModel* m, const FvxShardWs* ws, const int32_t* user, int B, const char* who) {
  if (int rc = fvx_check_model(m, who)) return rc;
  FVX_CHECK_ARG(user != nullptr && B >= 1 && B <= m->max_batch, "%s: bad batch", who);
  FVX_CHECK_ARG(m->users.list_cap >= B && m->items.list_cap >= 2 * B, "%s: touched-row lists too small", who);
  FVX_CHECK_ARG(m->rows && m->loss && m->sync && m->cmap, "%s: null scratch (rows / loss / sync / cmap)", who);
  FVX_CHECK_ARG(m->K % 4 == 0, "%s: the sharded step needs embed_k %% 4 == 0 (got %d)", who, m->K);
  FVX_CHECK_ARG(!m->two_stage, "%s: GradFashion (two_stage) runs on one GPU", who);
  FVX_CHECK_ARG(ws && ws->S && ws->run_id && ws->run_scratch && ws->WU && ws->RU && ws->dE && ws->loss_part &&
                ws->run_cap >= 1 && ws->owners >= 1 && ws->max_runs == ws->owners * ws->run_cap && ws->users_per_owner >= 1,
                "%s: incomplete FvxShardWs", who);
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

static SsTheta make_theta(const FvxModel* m, int B, int ks, int sm_reserve = 0) {
  SsTheta T;
  const bool tc = m->D > 0 && m->use_tensor_cores;
  T.p = m->TH;
  T.np = tc ? fvx_tc_np(m->de) : m->de;
  T.ks = tc ? ks : 1;
  T.ss = 2LL * B * T.np;
  T.chunks = m->D > 0 ? m->D / 64 : 1;
  T.nsm = fvx_num_sms() - sm_reserve;      // what the projection was launched with (its K split follows it)
  if (T.nsm < 8) T.nsm = 8;
  return T;
}

// Unique-row step on a shard (see fvx_train.cu): with R ranks a rank owns I/R catalog rows but 2B_global/R
// slots, so projecting each DISTINCT owned row once shrinks both contractions by the duplication factor.
// Needs the tensor-core path and the scratch (upos, W_sum, uslot); FVX_STEP_DEDUP=0 turns it off.
static bool sharded_uniq(const FvxModel* m) {
  return m->D > 0 && m->use_tensor_cores && m->upos && m->W_sum && m->uslot &&
         fvx_tc_np(m->de) <= 320 && fvx_dedup_enabled();
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
static inline int scan_grid(int B) {
  long long g = ((long long)B * 32 + 255) / 256 / 32 + 1;       // a warp looks at 32 triples per pass
  if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
  return (int)g;
}

// ---- the pieces of the step (file header) ---------------------------------------------------------------
struct ShCtx {
  const FvxModel* m;
  const FvxShardWs* ws;
  const int32_t *user, *pos, *neg;
  int B, loss_slot;
  bool uniq;
  int ks;
  int sm_reserve;       // SMs the tensor-core kernels leave to the NCCL kernels that travel beside them
};

// piece 1 in two halves: (a) what the projection needs - slot rows, claims of the owned item rows, E planes;
// (b) what only the user side needs - run ids, the cleared exchange buffers - which the full step runs on the side stream
static int sh_p1a(const ShCtx& c, cudaStream_t st) {
  return fvx_launch_prep(c.m, c.user, c.pos, c.neg, c.B, st, c.uniq ? FVX_PREP_UNIQ : FVX_PREP_ROWS);
}
static int sh_p1b(const ShCtx& c, cudaStream_t st) {
  // (WU needs no clearing: every row a kernel reads was written by the run's owner and gathered)
  return run_slots_impl(c.user, c.B, c.ws->users_per_owner, c.ws->owners, c.ws->run_cap, c.ws->run_id, c.ws->run_scratch,
                        c.ws->run_counts, st);
}
static int sh_clear_ru(const ShCtx& c, cudaStream_t st) {
  const FvxModel& M = *c.m;
  if (cudaMemsetAsync(c.ws->RU, 0, sizeof(float) * (size_t)c.ws->max_runs * M.users.stride, st) != cudaSuccess)
    FVX_FAIL(-3, "fvx_bpr_step_sharded: memset failed");
  return 0;
}
// piece 2 in two halves (unique-row step): (u) the owned users are caught up and published - the exchange of WU can
// start; (i) the listed item rows are caught up, which nothing waits for before the partial scores
static int sh_p2u(const ShCtx& c, cudaStream_t st) {
  const FvxModel& M = *c.m;
  if (int rc = fvx_launch_prep(&M, c.user, c.pos, c.neg, c.B, st, c.uniq ? FVX_PREP_USERS_ONLY : FVX_PREP_CLAIMS))
    return rc;
  k_pack_wu<<<scan_grid(c.B), 256, 0, st>>>(M, c.user, c.B, c.ws->run_id, c.ws->WU, (long long)c.ws->max_runs);
  FVX_CHECK_LAUNCH("k_pack_wu");
  return 0;
}
static int sh_p2i(const ShCtx& c, cudaStream_t st) {
  if (!c.uniq) return 0;
  return fvx_launch_prep(c.m, c.user, c.pos, c.neg, c.B, st, FVX_PREP_ITEMS_LISTED);
}
static int sh_p3(const ShCtx& c, cudaStream_t st) {
  const FvxModel& M = *c.m;
  const int B = c.B;
  // compact list of the owned slots; foreign entries of crow stay -1 (the fp32 kernels skip them)
  int32_t* count = M.sync + 1;
  cudaMemsetAsync(M.cmap, 0xFF, sizeof(int32_t) * 2 * (size_t)M.max_batch, st);
  cudaMemsetAsync(c.ws->S, 0, sizeof(float) * 2 * (size_t)B, st);
  {
    long long g = (2LL * B + 255) / 256;
    if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
    k_compact_owned<<<(int)g, 256, 0, st>>>(M, B, count);
    FVX_CHECK_LAUNCH("k_compact_owned");
  }
  if (c.uniq) {
    FVX_CHECK_ARG(2LL * B * fvx_tc_np(M.de) <= M.th_cap, "fvx_bpr_step_sharded: TH scratch too small");
    return fvx_launch_project_tc(&M, M.items.list, 0, 2 * B, c.ks, M.TH, st, M.items.count, 1, c.sm_reserve);
  }
  if (M.D > 0) {
    if (M.use_tensor_cores) {
      FVX_CHECK_ARG((long long)c.ks * 2 * B * fvx_tc_np(M.de) <= M.th_cap, "fvx_bpr_step_sharded: TH scratch too small");
      // (the W rows past the owned ones, read by the last backward tile, are zeroed by piece 5)
      return fvx_launch_project_tc(&M, M.cmap, 0, 2 * B, c.ks, M.TH, st, count, 0, c.sm_reserve);
    }
    FVX_CHECK_ARG(2LL * B * M.de <= M.th_cap, "fvx_bpr_step_sharded: TH scratch too small");
    return fvx_launch_project(&M, M.cmap, 2 * B, M.TH, st);
  }
  return 0;
}
static int sh_p4(const ShCtx& c, cudaStream_t st) {
  const FvxModel& M = *c.m;
  k_partial_scores_v4<<<ss_grid(c.B), SS_WARPS * 32, 0, st>>>(M, c.user, c.B, make_theta(&M, c.B, c.ks, c.sm_reserve), c.ws->S, M.sync + 1,
                                                             c.uniq ? 1 : 0, c.ws->WU, c.ws->run_id,
                                                             (long long)c.ws->max_runs);
  FVX_CHECK_LAUNCH("k_partial_scores");
  return 0;
}
static int sh_p5(const ShCtx& c, cudaStream_t st) {
  const FvxModel& M = *c.m;
  const bool tc = M.D > 0 && M.use_tensor_cores;
  // one half-warp per owned slot: the work does not grow with the number of ranks
  k_grads_owned<<<ss_grid(c.B), SS_WARPS * 32, 0, st>>>(M, c.user, c.B, c.loss_slot, make_theta(&M, c.B, c.ks, c.sm_reserve),
                                                       tc ? fvx_tc_np(M.de) : 0, tc ? fvx_w_pitch(&M) : 0, c.ws->S,
                                                       c.ws->run_id, c.ws->RU, (long long)c.ws->max_runs, M.sync + 1,
                                                       c.uniq ? 1 : 0, c.ws->WU, c.ws->loss_part);
  FVX_CHECK_LAUNCH("k_grads_owned");
  if (c.uniq) return fvx_launch_w_planes(&M, c.B, st);
  return 0;
}
static int sh_p6(const ShCtx& c, cudaStream_t st) {
  const FvxModel& M = *c.m;
  int parts = 0;
  const bool tc = M.D > 0 && M.use_tensor_cores;
  if (M.D > 0) {
    if (c.uniq) {
      if (int rc = fvx_launch_grad_E_tc(&M, M.items.list, 2 * c.B, &parts, st, M.items.count, c.sm_reserve)) return rc;
    } else if (tc) {
      if (int rc = fvx_launch_grad_E_tc(&M, M.cmap, 2 * c.B, &parts, st, M.sync + 1, c.sm_reserve)) return rc;
    } else {
      if (int rc = fvx_launch_grad_E(&M, M.cmap, 2 * c.B, &parts, st)) return rc;
    }
  }
  const int n = M.D * M.de;
  k_dE_pack<<<n > 0 ? (n + 255) / 256 : 1, 256, 0, st>>>(M.gE_part, parts, M.D, tc ? fvx_tc_np(M.de) : M.de, M.de, c.ws->dE,
                                                        c.ws->loss_part, M.sync);
  FVX_CHECK_LAUNCH("k_dE_pack");
  return 0;
}
static int sh_p7(const ShCtx& c, cudaStream_t st) {
  const FvxModel& M = *c.m;
  k_scatter_runs<<<scan_grid(c.B), 256, 0, st>>>(M, c.user, c.B, c.ws->run_id, c.ws->RU, (long long)c.ws->max_runs);
  FVX_CHECK_LAUNCH("k_scatter_runs");
  return 0;
}
static int sh_p8(const ShCtx& c, cudaStream_t st) {
  const FvxModel& M = *c.m;
  // DEFERRED: the touched rows keep their gradient and take the step when they are next needed (replay_row);
  // the reduced dE, the loss (sum of the ranks' shares, NaN after a run overflow) and the step counter
  return fvx_launch_update(&M, c.B, M.D > 0 ? 1 : 0, M.de, c.ws->dE, c.loss_slot, st,
                           fvx_merged_update(&M) ? FVX_UPD_E : FVX_UPD_ALL, c.ws->dE + (size_t)M.D * M.de, 1);
}
// peer-to-peer: dE arrives as one partial per rank (dEall[R][D*de], summed in rank order by the update), the loss
// shares as one tail per rank
static int sh_p8_p2p(const ShCtx& c, int world, cudaStream_t st) {
  const FvxModel& M = *c.m;
  return fvx_launch_update(&M, c.B, M.D > 0 ? world : 0, M.de, c.ws->dEall, c.loss_slot, st,
                           fvx_merged_update(&M) ? FVX_UPD_E : FVX_UPD_ALL, c.ws->tails, world);
}

static int make_ctx(ShCtx* c, const FvxModel* m, const FvxShardWs* ws, const int32_t* user, const int32_t* pos,
                    const int32_t* neg, int B, int loss_slot, const char* who) {
  if (int rc = sharded_common(m, ws, user, B, who)) return rc;
  FVX_CHECK_ARG(pos && neg, "%s: null batch pointer", who);
  FVX_CHECK_ARG(loss_slot >= 0 && loss_slot < m->loss_slots, "%s: loss_slot out of range", who);
  c->m = m; c->ws = ws; c->user = user; c->pos = pos; c->neg = neg; c->B = B; c->loss_slot = loss_slot;
  c->uniq = sharded_uniq(m);
  c->ks = c->uniq ? uniq_ks_cap(m, B) : sharded_ks(m, B);
  c->sm_reserve = 0;
  return 0;
}

// Timeline of fvx_bpr_step_sharded (diagnostics, declared in fvx.h): with tracing on, the step records a timing
// event after every piece / collective on the stream it runs on.
enum { ST_BEGIN = 0, ST_P1, ST_P2, ST_AR_WU, ST_P3, ST_P4, ST_AR_S, ST_P5, ST_AR_RU, ST_P7, ST_P6, ST_AR_DE, ST_END, ST_COUNT };
static cudaEvent_t g_st_ev[ST_COUNT];
static int g_st_on = 0;
#define STRACE(i, stream) do { if (g_st_on) cudaEventRecord(g_st_ev[i], (stream)); } while (0)

extern "C" {

int fvx_debug_trace_sharded(int on) {
  if (on && !g_st_ev[0])
    for (int i = 0; i < ST_COUNT; ++i)
      if (cudaEventCreate(&g_st_ev[i]) != cudaSuccess) return -3;
  g_st_on = on ? 1 : 0;
  return 0;
}
int fvx_debug_trace_sharded_read(float* us_host) {   // [13]; synchronises
  if (!g_st_ev[0]) return -2;
  if (cudaEventSynchronize(g_st_ev[ST_END]) != cudaSuccess) return -3;
  for (int i = 0; i < ST_COUNT; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_st_ev[ST_BEGIN], g_st_ev[i]) != cudaSuccess) { cudaGetLastError(); ms = -1e-3f; }
    us_host[i] = ms * 1e3f;
  }
  return 0;
}

int fvx_bpr_step_sharded_phase(const FvxModel* model, const FvxShardWs* ws, const int32_t* user, const int32_t* pos,
                               const int32_t* neg, int32_t B, int32_t loss_slot, int32_t phase, fvx_stream_t stream) {
  ShCtx c;
  if (int rc = make_ctx(&c, model, ws, user, pos, neg, B, loss_slot, "fvx_bpr_step_sharded_phase")) return rc;
  cudaStream_t st = fvx_cu(stream);
  switch (phase) {
    case 0:
      if (int rc = sh_p1a(c, st)) return rc;
      if (int rc = sh_p1b(c, st)) return rc;
      if (int rc = sh_p2u(c, st)) return rc;
      if (int rc = sh_clear_ru(c, st)) return rc;
      return sh_p2i(c, st);
    case 1:
      if (int rc = sh_p3(c, st)) return rc;
      return sh_p4(c, st);
    case 2:
      if (int rc = sh_p5(c, st)) return rc;
      return sh_p6(c, st);
    case 3:
      if (int rc = sh_p7(c, st)) return rc;
      return sh_p8(c, st);
    default:
      FVX_FAIL(-2, "fvx_bpr_step_sharded_phase: phase %d outside [0, 3]", phase);
  }
}

// The step with the peer-to-peer exchange (file header; needs the communicator's arena): same pieces, the four
// collectives replaced by peer stores inside the producing kernels + one-warp barrier kernels.
static int step_sharded_p2p(ShCtx& c, FvxComm* comm, cudaStream_t st) {
  const FvxModel& M = *c.m;
  const FvxShardWs* ws = c.ws;
  cudaStream_t sd = comm->side;
  PeerPtrs X;
  for (int r = 0; r < FVX_COMM_MAX_RANKS; ++r) X.p[r] = r < comm->world ? comm->peer[r] : nullptr;
  X.self = comm->arena; X.rank = comm->rank; X.world = comm->world;
  const uint8_t *lo = comm->arena, *hi = comm->arena + comm->arena_bytes;
  auto in_arena = [&](const void* p) { return (const uint8_t*)p >= lo && (const uint8_t*)p < hi; };
  FVX_CHECK_ARG(in_arena(ws->WU) && in_arena(ws->S) && in_arena(ws->RUin) && in_arena(ws->dEall) && in_arena(ws->tails) &&
                in_arena(ws->flags) && ws->run_user && ws->run_counts,
                "fvx_bpr_step_sharded: peer-to-peer mode needs WU, S, RUin, dEall, tails, flags inside the communicator's arena");
  const uint32_t epoch = ++comm->epoch;                 // one value per step; one flag array per barrier of the step
  uint32_t* flags = ws->flags;
  const int base = M.num_items / comm->world, rem = M.num_items % comm->world;
  const int cap = ws->run_cap, S4 = M.users.stride >> 2;
  STRACE(ST_BEGIN, st);
  if (int rc = sh_p1a(c, st)) return rc;
  if (int rc = sh_p1b(c, st)) return rc;
  STRACE(ST_P1, st);
  if (int rc = fvx_launch_prep(&M, c.user, c.pos, c.neg, c.B, st, c.uniq ? FVX_PREP_USERS_ONLY : FVX_PREP_CLAIMS)) return rc;
  cudaMemsetAsync(ws->S, 0, sizeof(float) * 2 * (size_t)c.B, st);
  k_pack_wu_p2p<<<scan_grid(c.B), 256, 0, st>>>(M, c.user, c.B, ws->run_id, ws->WU, (long long)ws->max_runs, ws->run_user, cap, X);
  FVX_CHECK_LAUNCH("k_pack_wu_p2p");
  STRACE(ST_P2, st);
  cudaEventRecord(comm->ev[0], st);
  cudaStreamWaitEvent(sd, comm->ev[0], 0);
  // side, beside the projection: barrier 0 (every rank's fresh user rows have landed), cleared RU, item catch-up
  k_xbar<<<1, 32, 0, sd>>>(X, flags + 0 * FVX_COMM_MAX_RANKS, epoch);
  STRACE(ST_AR_WU, sd);
  if (int rc = sh_clear_ru(c, sd)) return rc;
  if (int rc = sh_p2i(c, sd)) return rc;
  cudaEventRecord(comm->ev[1], sd);
  {
    // piece 3 without its own clearing of S (cleared above, before this rank's barrier signal: peers store into it)
    int32_t* count = M.sync + 1;
    cudaMemsetAsync(M.cmap, 0xFF, sizeof(int32_t) * 2 * (size_t)M.max_batch, st);
    long long g = (2LL * c.B + 255) / 256;
    if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
    k_compact_owned<<<(int)g, 256, 0, st>>>(M, c.B, count);
    FVX_CHECK_LAUNCH("k_compact_owned");
    if (c.uniq) {
      FVX_CHECK_ARG(2LL * c.B * fvx_tc_np(M.de) <= M.th_cap, "fvx_bpr_step_sharded: TH scratch too small");
      if (int rc = fvx_launch_project_tc(&M, M.items.list, 0, 2 * c.B, c.ks, M.TH, st, M.items.count, 1, c.sm_reserve)) return rc;
    } else if (M.D > 0) {
      if (M.use_tensor_cores) {
        if (int rc = fvx_launch_project_tc(&M, M.cmap, 0, 2 * c.B, c.ks, M.TH, st, count, 0, c.sm_reserve)) return rc;
      } else {
        if (int rc = fvx_launch_project(&M, M.cmap, 2 * c.B, M.TH, st)) return rc;
      }
    }
  }
  STRACE(ST_P3, st);
  cudaStreamWaitEvent(st, comm->ev[1], 0);
  if (int rc = sh_p4(c, st)) return rc;
  {
    long long g = (2LL * c.B / comm->world + 255) / 256 + 1;
    if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
    k_send_scores<<<(int)g, 256, 0, st>>>(M, c.pos, c.neg, c.B, ws->S, M.sync + 1, base, rem, X);
    FVX_CHECK_LAUNCH("k_send_scores");
  }
  STRACE(ST_P4, st);
  k_xbar<<<1, 32, 0, st>>>(X, flags + 1 * FVX_COMM_MAX_RANKS, epoch);       // barrier 1: both sides of every owned triple
  STRACE(ST_AR_S, st);
  if (int rc = sh_p5(c, st)) return rc;
  STRACE(ST_P5, st);
  cudaEventRecord(comm->ev[2], st);
  cudaStreamWaitEvent(sd, comm->ev[2], 0);
  // side, beside grad_E: the gradient shares go to the owners of their runs, who add them up
  {
    long long g = ((long long)ws->max_runs * S4 + 255) / 256;
    if (g > (long long)fvx_num_sms() * 4) g = (long long)fvx_num_sms() * 4;
    k_push_ru<<<(int)g, 256, 0, sd>>>(ws->RU, ws->RUin, ws->run_counts, cap, S4, X);
    FVX_CHECK_LAUNCH("k_push_ru");
    k_xbar<<<1, 32, 0, sd>>>(X, flags + 2 * FVX_COMM_MAX_RANKS, epoch);     // barrier 2: every source's shares have landed
    STRACE(ST_AR_RU, sd);
    long long g2 = ((long long)cap * 32 + 255) / 256;
    if (g2 > (long long)fvx_num_sms() * 8) g2 = (long long)fvx_num_sms() * 8;
    k_scatter_runs_p2p<<<(int)g2, 256, 0, sd>>>(M, ws->RUin, ws->run_user, ws->run_counts, cap, X);
    FVX_CHECK_LAUNCH("k_scatter_runs_p2p");
  }
  STRACE(ST_P7, sd);
  cudaEventRecord(comm->ev[3], sd);
  {
    int parts = 0;
    const bool tc = M.D > 0 && M.use_tensor_cores;
    if (M.D > 0) {
      if (c.uniq) {
        if (int rc = fvx_launch_grad_E_tc(&M, M.items.list, 2 * c.B, &parts, st, M.items.count, c.sm_reserve)) return rc;
      } else if (tc) {
        if (int rc = fvx_launch_grad_E_tc(&M, M.cmap, 2 * c.B, &parts, st, M.sync + 1, c.sm_reserve)) return rc;
      } else {
        if (int rc = fvx_launch_grad_E(&M, M.cmap, 2 * c.B, &parts, st)) return rc;
      }
    }
    const int n = M.D * M.de;
    k_dE_pack_p2p<<<n > 0 ? (n + 255) / 256 : 1, 256, 0, st>>>(M.gE_part, parts, M.D, tc ? fvx_tc_np(M.de) : M.de, M.de,
                                                              ws->dEall, ws->tails, ws->loss_part, M.sync, X);
    FVX_CHECK_LAUNCH("k_dE_pack_p2p");
  }
  STRACE(ST_P6, st);
  k_xbar<<<1, 32, 0, st>>>(X, flags + 3 * FVX_COMM_MAX_RANKS, epoch);       // barrier 3: every rank's dE has landed
  STRACE(ST_AR_DE, st);
  cudaStreamWaitEvent(st, comm->ev[3], 0);
  if (int rc = sh_p8_p2p(c, comm->world, st)) return rc;
  STRACE(ST_END, st);
  return 0;
}

int fvx_bpr_step_sharded(const FvxModel* model, const FvxShardWs* ws, FvxComm* comm, const int32_t* user,
                         const int32_t* pos, const int32_t* neg, int32_t B, int32_t loss_slot, fvx_stream_t stream) {
  ShCtx c;
  if (int rc = make_ctx(&c, model, ws, user, pos, neg, B, loss_slot, "fvx_bpr_step_sharded")) return rc;
  FVX_CHECK_ARG(comm != nullptr, "fvx_bpr_step_sharded: null communicator");
  if (ws->p2p) {
    FVX_CHECK_ARG(comm->arena != nullptr, "fvx_bpr_step_sharded: FvxShardWs.p2p needs fvx_comm_arena");
    FVX_CHECK_ARG(ws->owners == comm->world, "fvx_bpr_step_sharded: FvxShardWs.owners %d != communicator size %d", ws->owners,
                  comm->world);
    c.sm_reserve = 0;
    return step_sharded_p2p(c, comm, fvx_cu(stream));
  }
  const FvxModel& M = *model;
  cudaStream_t st = fvx_cu(stream), sd = comm->side;
  {
    static int serial = -1;            // FVX_SH_SERIAL=1: every piece and collective on the caller's stream (measurements)
    if (serial < 0) { const char* e_ = getenv("FVX_SH_SERIAL"); serial = (e_ && atoi(e_) == 1) ? 1 : 0; }
    if (serial) sd = st;
  }
  const size_t seg = (size_t)ws->run_cap * M.users.stride;      // one owner's segment of WU / RU
  FVX_CHECK_ARG(ws->owners == comm->world, "fvx_bpr_step_sharded: FvxShardWs.owners %d != communicator size %d", ws->owners,
                comm->world);
  {
    // the persistent tensor-core kernels leave a few SMs to the NCCL kernels that run beside them (FVX_COMM_SMS,
    // default 8): without them an all-reduce waits for the last CTA of the projection / grad_E to retire
    static int reserve = -1;
    if (reserve < 0) { const char* e_ = getenv("FVX_COMM_SMS"); reserve = e_ ? atoi(e_) : 8; if (reserve < 0 || reserve > 64) reserve = 8; }
    c.sm_reserve = comm->world > 1 ? reserve : 0;
  }
  STRACE(ST_BEGIN, st);
  // main: item claims, then the little the user side needs before its exchange can start - run slots, catch-up of
  // the OWNED users of the batch, their rows packed into this rank's segment of WU (~30 us; beside the projection
  // the same kernels took 100 us and the projection 60 us longer)
  if (int rc = sh_p1a(c, st)) return rc;
  if (int rc = sh_p1b(c, st)) return rc;
  STRACE(ST_P1, st);
  if (int rc = sh_p2u(c, st)) return rc;
  STRACE(ST_P2, st);
  cudaEventRecord(comm->ev[0], st);
  cudaStreamWaitEvent(sd, comm->ev[0], 0);
  // side, beside the projection: the fresh user rows travel; the listed item rows are caught up
  if (int rc = fvx_comm_allgather(comm, 1, ws->WU, seg, sd)) return rc;
  STRACE(ST_AR_WU, sd);
  if (int rc = sh_clear_ru(c, sd)) return rc;
  if (int rc = sh_p2i(c, sd)) return rc;
  cudaEventRecord(comm->ev[1], sd);
  if (int rc = sh_p3(c, st)) return rc;
  STRACE(ST_P3, st);
  cudaStreamWaitEvent(st, comm->ev[1], 0);
  if (int rc = sh_p4(c, st)) return rc;
  STRACE(ST_P4, st);
  if (int rc = fvx_comm_allreduce(comm, 0, ws->S, 2 * (size_t)B, st)) return rc;
  STRACE(ST_AR_S, st);
  if (int rc = sh_p5(c, st)) return rc;
  STRACE(ST_P5, st);
  cudaEventRecord(comm->ev[2], st);
  cudaStreamWaitEvent(sd, comm->ev[2], 0);
  // side: the user-row gradient shares travel, and the owners take theirs, while grad_E runs
  if (int rc = fvx_comm_reducescatter(comm, 1, ws->RU, seg, sd)) return rc;
  STRACE(ST_AR_RU, sd);
  if (int rc = sh_p7(c, sd)) return rc;
  STRACE(ST_P7, sd);
  cudaEventRecord(comm->ev[3], sd);
  if (int rc = sh_p6(c, st)) return rc;
  STRACE(ST_P6, st);
  if (int rc = fvx_comm_allreduce(comm, 0, ws->dE, (size_t)M.D * M.de + 4, st)) return rc;
  STRACE(ST_AR_DE, st);
  cudaStreamWaitEvent(st, comm->ev[3], 0);
  if (int rc = sh_p8(c, st)) return rc;
  STRACE(ST_END, st);
  return 0;
}

}  // extern "C"
