// Single-pass VBPR step: projection, scoring and the gradient of E read every feature row ONCE.
//
// Replaces the projection -> score -> grad_E kernel triple of fvx_train.cu for
// VBPR.train_step (src/recommender/models/VBPR.py:99-144: two VBPR.call forward passes
// `matmul(feature_i, E)`, the BPR loss, tape.gradient w.r.t. E / Bp and the sparse gradients).
// The two-kernel path reads the 8 KB feature row of every (triple, side) slot twice (forward and
// backward); here a thread-block CLUSTER of 8 CTAs keeps a tile of rows resident in distributed
// shared memory between the two contractions:
//
//   * tile = TR/2 triples = TR rows (positives, then negatives).  CTA c of the cluster owns the
//     feature slice [c*SL, (c+1)*SL), SL = D/8, of every row of the tile: TR x SL x 4 B per stage
//     (bf16 hi|lo planes, 128B-swizzled, written by 16-byte cp.async row gathers).
//   * forward : tcgen05.mma  part_c[TR x 32] = F_tile[:, slice_c] * E_ext[slice_c, :]   (K-major A)
//     hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM.  Each CTA pushes the partial rows of
//     triple t to the CTA that OWNS the triple (st.shared::cluster + remote mbarrier arrive).
//   * score   : the owner sums the 8 partials (fixed order), computes x_uij, the loss, the
//     coefficient and the user / item gradients (red.global.add.v4 into g), and pushes the rows
//     W = +-c*[theta_u | 1] (bf16 hi|lo, 64B-swizzled) into the W tile of all 8 CTAs.
//   * backward: tcgen05.mma  gE[slice_c, :] += F_tile[:, slice_c]^T * W   (the SAME shared-memory
//     tile read as the MN-major operand); the accumulator stays in TMEM over all tiles of the
//     cluster and leaves once as partial `cluster id` of gE_part.
//
// A stage is held from the start of its gather until the backward MMAs have consumed it; stages
// alternate so that the gather of one tile overlaps the forward -> score -> backward chain of the
// other.  Roles: warps 0-3 gather, warp 4 issues UMMAs (forward and backward, whichever is ready),
// warps 5-8 drain TMEM / exchange / score (one warp per triple).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"
#include "fvx_tc.cuh"

#define FS_CL 8         // CTAs per cluster = feature slices per row
#define FS_THREADS 288  // warps 0-3 producers, warp 4 UMMA issuer, warps 5-8 epilogue / scoring
#define FS_PROD 128
#define FS_NP 32        // padded width of E_ext (d + 1 <= 32)
#define FS_NACC 4       // independent forward accumulators (one per UMMA K step of a chunk)
#define FS_TMEM_BWD 256 // first TMEM column of the backward accumulators

// ---- cluster-scope PTX -------------------------------------------------------------------
TC_D uint32_t cl_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
TC_D uint32_t cl_clusterid() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
TC_D uint32_t cl_nclusters() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank`
TC_D uint32_t cl_map(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
TC_D void cl_st_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
// release at cluster scope: the stores of this thread issued before are visible to the waiter
TC_D void cl_arrive(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
TC_D bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(tc_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
TC_D void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cl(bar, parity)) return;
  const long long t0 = clock64();
#pragma unroll 1
  while (!mbar_try_wait_cl(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
TC_D void cl_sync() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
TC_D void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
TC_D void fs_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
TC_D void fs_cp_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
TC_D void fs_red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
TC_D float fs_dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
TC_D uint2 fs_pack_bf16x4(float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
TC_D float4 fs_unpack_bf16x4(uint2 p) {
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&p.x);
  const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&p.y);
  const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

struct FusedParams {
  const uint8_t* Fpl;   // [item_cnt][D/64][2][64] bf16 planes
  const int32_t* user;  // [B]
  int B;
  int n_tiles;          // ceil(B / (TR/2))
  int loss_slot;
  float* gE_out;        // [clusters][D][FS_NP]
  long long* trace;     // optional [tiles of cluster 0][16] clock64 stamps of CTA 0 (profiling aid), else nullptr
};

// Shared-memory map of one CTA (offsets from the 1024-byte aligned base).
template <int TR, int NST, int SL>
struct FusedSmem {
  static constexpr int TPC = TR / 16;              // triples scored per CTA and tile
  static constexpr int NCH = SL / 64;              // 64-feature chunks of the slice
  static constexpr int NMB = SL / 128;             // 128-feature UMMA M blocks of the backward
  static constexpr uint32_t A_SUB = TR * 128u;     // one (chunk, plane): [TR rows x 128 B]
  static constexpr uint32_t A_PLANE = NCH * A_SUB;
  static constexpr uint32_t A_STAGE = 2 * A_PLANE;
  static constexpr uint32_t E_SUB = FS_NP * 128u;  // one (chunk, plane) of E_ext^T: [32 x 128 B]
  static constexpr uint32_t E_PLANE = NCH * E_SUB;
  static constexpr uint32_t W_PLANE = TR * 64u;    // [TR rows x 32 bf16]
  static constexpr uint32_t W_STAGE = 2 * W_PLANE;
  static constexpr uint32_t XB_STAGE = FS_CL * 2 * TPC * 128u;  // [src CTA][own row][32 fp32]
  static constexpr uint32_t OFF_A = 0;
  static constexpr uint32_t OFF_E = OFF_A + NST * A_STAGE;
  static constexpr uint32_t OFF_W = OFF_E + 2 * E_PLANE;
  static constexpr uint32_t OFF_XB = OFF_W + NST * W_STAGE;
  static constexpr uint32_t OFF_STG = OFF_XB + NST * XB_STAGE;   // [4 warps][4 x 64 B]
  static constexpr uint32_t OFF_BAR = OFF_STG + 4 * 256u;
  static constexpr int N_BARS = 4 * NST + 2 + 2 + 2;             // full, free, xb, w | t_full, t_empty | e_full, final
  static constexpr uint32_t OFF_SLOT = OFF_BAR + N_BARS * 8u;
  static constexpr uint32_t BYTES = OFF_SLOT + 16u;
  // the forward UMMA is issued with M = 128 on a TR-row tile (rows >= TR are ignored garbage):
  // the last sub-tile reads 16 KB from its base and must stay inside the allocation
  static_assert(TR == 64 || TR == 48, "tile rows");
  static_assert(A_SUB % 1024 == 0 && W_PLANE % 512 == 0, "swizzle atoms");
  static_assert(2 * E_PLANE + NST * W_STAGE >= 16384, "forward M=128 over-read leaves the allocation");
};

#define FS_STAMP(it_, ev_) do { if (tracing) P.trace[(it_) * 16 + (ev_)] = clock64(); } while (0)

// ------------------------------------------------------------------------------------------
template <int TR, int NST, int SL>
__global__ void __launch_bounds__(FS_THREADS, 1)
k_step_fused(const __grid_constant__ CUtensorMap tmE_hi, const __grid_constant__ CUtensorMap tmE_lo,
             const FvxModel M, const FusedParams P) {
  using L = FusedSmem<TR, NST, SL>;
  constexpr int TPC = L::TPC, NCH = L::NCH, NMB = L::NMB, HT = TR / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = tc_smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cl_ctarank();
  const int cid = (int)cl_clusterid(), ncl = (int)cl_nclusters();

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* full_b = bars;                 // [NST] the gathered tile has landed        (128 cp.async arrivals)
  uint64_t* free_b = bars + NST;           // [NST] backward UMMAs have consumed it       (1 commit)
  uint64_t* xb_full = bars + 2 * NST;      // [NST] all partial rows of my triples are in (TR remote arrivals)
  uint64_t* w_full = bars + 3 * NST;       // [NST] the W tile is complete                (2*TR remote arrivals)
  uint64_t* t_full = bars + 4 * NST;       // [2]   forward accumulator ready             (1 commit)
  uint64_t* t_empty = t_full + 2;          // [2]   forward accumulator drained           (2 warps)
  uint64_t* e_full = t_empty + 2;          // [1]   E_ext^T slice loaded                  (TMA)
  uint64_t* final_b = e_full + 1;          // [1]   all UMMAs of the CTA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_SLOT);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_b[s], FS_PROD);
      mbar_init(&free_b[s], 1);
      mbar_init(&xb_full[s], TR);
      mbar_init(&w_full[s], 2 * TR);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 2); }
    mbar_init(e_full, 1);
    mbar_init(final_b, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmE_hi); tma_prefetch_desc(&tmE_lo);
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cl_sync();                               // every CTA's barriers exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int B = P.B;
  const bool tracing = P.trace != nullptr && cid == 0 && crank == 0 && lane == 0;
  const int cnt = cid < P.n_tiles ? (P.n_tiles - cid + ncl - 1) / ncl : 0;   // tiles of this cluster

  if (warp < 4) {
    // ===== producers: 16-byte cp.async pieces; SL/4 consecutive pieces (one row slice) per group =====
    constexpr int PR = SL / 4;             // 16-byte pieces per row slice (both planes)
    constexpr int RPP = FS_PROD / PR;      // rows per pass of the 128 threads
    constexpr int IT = TR / RPP;
    const int tid = threadIdx.x;
    const int e = tid % PR, sub = tid / PR;
    const int chunk = e >> 4, plane = (e >> 3) & 1, c16 = e & 7;
    const size_t row_bytes = (size_t)M.D * 4;
    const size_t slice_off = (size_t)crank * SL * 4 + (size_t)e * 16;
    const uint32_t doff = (uint32_t)plane * L::A_PLANE + (uint32_t)chunk * L::A_SUB;
    if (tid == 0 && cnt > 0) {
      // E_ext^T slice: [plane][chunk][32 x 128 B], resident for the whole kernel
      mbar_expect_tx(e_full, 2 * L::E_PLANE);
      for (int c = 0; c < NCH; ++c) {
        tma_load_2d(smem + L::OFF_E + c * L::E_SUB, &tmE_hi, e_full, (int)crank * SL + c * 64, 0);
        tma_load_2d(smem + L::OFF_E + L::E_PLANE + c * L::E_SUB, &tmE_lo, e_full, (int)crank * SL + c * 64, 0);
      }
    }
    for (int it = 0; it < cnt; ++it) {
      const int tile = cid + it * ncl;
      const uint32_t s = it % NST, ph = (it / NST) & 1;
      const uint8_t* src[IT];
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        const int r = j * RPP + sub;
        const int b = tile * HT + (r < HT ? r : r - HT);
        int32_t item = 0;
        if (b < B) item = __ldg(M.rows + (r < HT ? 0 : B) + b);
        if (item < 0) item = 0;            // ignored triple: any valid row, its W rows are zero
        src[j] = P.Fpl + (size_t)item * row_bytes + slice_off;
      }
      if (lane == 0) mbar_wait(&free_b[s], ph ^ 1);
      __syncwarp();
      if (warp == 0) FS_STAMP(it, 0);
      const uint32_t sA = sbase + L::OFF_A + s * L::A_STAGE + doff;
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        const int r = j * RPP + sub;
        fs_cp16(sA + r * 128 + ((c16 ^ (r & 7)) << 4), src[j]);
      }
      fs_cp_arrive(&full_b[s]);
      if (warp == 0) FS_STAMP(it, 1);
    }
  } else if (warp == 4) {
    // ===== UMMA issuer: forward of tile nf and backward of tile nb, whichever is ready =====
    if (lane == 0 && cnt > 0) {
      mbar_wait(e_full, 0);
      const uint32_t idesc_f = umma_idesc_bf16(128, FS_NP, 0, 0);
      const uint32_t idesc_b = umma_idesc_bf16(128, FS_NP, 1, 1);
      const uint32_t e_hi = sbase + L::OFF_E, e_lo = e_hi + L::E_PLANE;
      int nf = 0, nb = 0;
      long long t0 = clock64();
      while (nb < cnt) {
        bool did = false;
        if (nb < nf) {
          const uint32_t s = nb % NST, ph = (nb / NST) & 1;
          if (mbar_try_wait_cl(&w_full[s], ph)) {
            FS_STAMP(nb, 4);
            fence_proxy_async_all();       // W rows were written through the generic proxy (remote stores)
            tc_fence_after();
            const uint32_t a_hi = sbase + L::OFF_A + s * L::A_STAGE, a_lo = a_hi + L::A_PLANE;
            const uint32_t w_hi = sbase + L::OFF_W + s * L::W_STAGE, w_lo = w_hi + L::W_PLANE;
#pragma unroll
            for (int k = 0; k < TR / 16; ++k) {
#pragma unroll
              for (int pass = 0; pass < 3; ++pass) {
                const uint64_t dw = umma_smem_desc((pass == 2 ? w_lo : w_hi) + k * 16 * 64, L::W_PLANE, 512, TC_SWZ_64B);
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb) {
                  const uint32_t aoff = (uint32_t)(2 * mb) * L::A_SUB + k * 16 * 128;
                  const uint64_t da = umma_smem_desc((pass == 1 ? a_lo : a_hi) + aoff, L::A_SUB, 1024, TC_SWZ_128B);
                  const bool first = nb == 0 && pass == 0 && k < 2;
                  umma_f16(tmem_base + FS_TMEM_BWD + (uint32_t)(mb * 2 + (k & 1)) * FS_NP, da, dw, idesc_b,
                           first ? 0u : 1u);
                }
              }
            }
            umma_commit(&free_b[s]);
            FS_STAMP(nb, 5);
            ++nb;
            did = true;
          }
        }
        if (nf < cnt) {
          const uint32_t s = nf % NST, ph = (nf / NST) & 1, acc = nf & 1, aph = (nf >> 1) & 1;
          if (mbar_try_wait(&t_empty[acc], aph ^ 1) && mbar_try_wait(&full_b[s], ph)) {
            FS_STAMP(nf, 2);
            fence_proxy_async_smem();      // cp.async (generic proxy) writes -> UMMA (async proxy) reads
            tc_fence_after();
            const uint32_t a_hi = sbase + L::OFF_A + s * L::A_STAGE, a_lo = a_hi + L::A_PLANE;
            const uint32_t d0 = tmem_base + acc * (FS_NACC * FS_NP);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
#pragma unroll
              for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = umma_smem_desc((pass == 1 ? a_lo : a_hi) + c * L::A_SUB + k * 32, 16, 1024, TC_SWZ_128B);
                  const uint64_t db = umma_smem_desc((pass == 2 ? e_lo : e_hi) + c * L::E_SUB + k * 32, 16, 1024, TC_SWZ_128B);
                  umma_f16(d0 + (uint32_t)k * FS_NP, da, db, idesc_f, (c == 0 && pass == 0) ? 0u : 1u);
                }
              }
            }
            umma_commit(&t_full[acc]);
            FS_STAMP(nf, 3);
            ++nf;
            did = true;
          }
        }
        if (did) t0 = clock64();
        else if (clock64() - t0 > 4000000000LL) __trap();
      }
      umma_commit(final_b);
    }
  } else {
    // ===== epilogue warps: TMEM drain + exchange (2 warps), scoring (one warp per triple) =====
    const int widx = warp - 5;             // triple of the CTA this warp scores
    const int quad = warp & 3;             // TMEM lane quarter this warp may read
    const int Su = M.users.stride, Si = M.items.stride, K = M.K, d = M.d;
    const int K4 = K >> 2, D4 = (d + 3) >> 2;
    const float reg = M.reg, reg2 = 2.0f * M.reg;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int cv = lane - 16;              // lanes 16..23: 4-column chunk of the visual part
    double loss_acc = 0.0;
    uint8_t* stg = smem + L::OFF_STG + widx * 256;

    for (int it = 0; it < cnt; ++it) {
      const int tile = cid + it * ncl;
      const uint32_t s = it % NST, ph = (it / NST) & 1, acc = it & 1, aph = (it >> 1) & 1;

      // (A) rows of my triple: independent of the projection, in flight while it runs
      const int lt = (int)crank * TPC + widx;          // triple within the tile
      const int b = tile * HT + lt;
      int32_t u = -1, li = -1, lj = -1;
      if (widx < TPC && b < B) { u = __ldg(P.user + b); li = __ldg(M.rows + b); lj = __ldg(M.rows + B + b); }
      const bool dead = li < 0 || lj < 0 || u < 0 || u >= M.num_users;
      float4 a = z4, x = z4, y = z4, tu = z4;
      float bi = 0.f, bj = 0.f;
      if (!dead) {
        const float4* ur = reinterpret_cast<const float4*>(M.users.w + (size_t)u * Su);
        const float4* gi = reinterpret_cast<const float4*>(M.items.w + (size_t)li * Si);
        const float4* gj = reinterpret_cast<const float4*>(M.items.w + (size_t)lj * Si);
        if (lane < K4) { a = ur[lane]; x = gi[lane]; y = gj[lane]; }
        if (cv >= 0 && cv < D4) {
          tu = ur[K4 + cv];
          const int n0 = 4 * cv;           // columns >= d of the chunk are not latent terms
          if (n0 + 1 >= d) tu.y = 0.f;
          if (n0 + 2 >= d) tu.z = 0.f;
          if (n0 + 3 >= d) tu.w = 0.f;
        }
        bi = __ldg(M.items.w + (size_t)li * Si + K);
        bj = __ldg(M.items.w + (size_t)lj * Si + K);
      }

      // (B) forward partial rows -> the CTA that owns the row's triple
      if (quad < 2) {
        mbar_wait(&t_full[acc], aph);
        if (warp == 5) FS_STAMP(it, 6);
        tc_fence_after();
        float sum[32];
#pragma unroll
        for (int q = 0; q < FS_NACC; ++q) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (FS_NACC * FS_NP) + q * FS_NP, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] = q == 0 ? __uint_as_float(v[j]) : sum[j] + __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acc]);
        const int r = quad * 32 + lane;
        if (r < TR) {
          const int side = r >= HT ? 1 : 0;
          const int t_ = r - side * HT;
          const uint32_t owner = (uint32_t)(t_ / TPC);
          const int lrow = side * TPC + (t_ - (int)owner * TPC);
          const uint32_t dst = cl_map(sbase + L::OFF_XB + s * L::XB_STAGE + ((uint32_t)crank * 2 * TPC + lrow) * 128u, owner);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            cl_st_v4(dst + j * 4, __float_as_uint(sum[j]), __float_as_uint(sum[j + 1]), __float_as_uint(sum[j + 2]),
                     __float_as_uint(sum[j + 3]));
          cl_arrive(cl_map(tc_smem_u32(&xb_full[s]), owner));
        }
        if (warp == 5) FS_STAMP(it, 7);
      }

      // (C) score my triple
      if (widx < TPC) {
        if (lane == 0) mbar_wait_cl(&xb_full[s], ph);
        __syncwarp();
        if (warp == 5) FS_STAMP(it, 8);
        float4 dt = z4;
        float vb_l = 0.f;
        if (cv >= 0 && cv < FS_NP / 4) {
          const float4* xp = reinterpret_cast<const float4*>(smem + L::OFF_XB + s * L::XB_STAGE) + cv;
          float4 tp = z4, tn = z4;
#pragma unroll
          for (int p = 0; p < FS_CL; ++p) {
            const float4 vp = xp[(p * 2 * TPC + widx) * 8], vn = xp[(p * 2 * TPC + TPC + widx) * 8];
            tp.x += vp.x; tp.y += vp.y; tp.z += vp.z; tp.w += vp.w;
            tn.x += vn.x; tn.y += vn.y; tn.z += vn.z; tn.w += vn.w;
          }
          dt = make_float4(tp.x - tn.x, tp.y - tn.y, tp.z - tn.z, tp.w - tn.w);
          const int n0 = 4 * cv;
          if (n0 == d) vb_l = dt.x;        // column d: visual bias F[i]*Bp - F[j]*Bp
          if (n0 + 1 == d) vb_l = dt.y;
          if (n0 + 2 == d) vb_l = dt.z;
          if (n0 + 3 == d) vb_l = dt.w;
          if (n0 >= d) dt.x = 0.f;
          if (n0 + 1 >= d) dt.y = 0.f;
          if (n0 + 2 >= d) dt.z = 0.f;
          if (n0 + 3 >= d) dt.w = 0.f;
        }
        const float vb = __shfl_sync(0xffffffffu, vb_l, 16 + (d >> 2));
        float part, sq;
        if (lane < 16) {
          const float4 df = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
          part = fs_dot4(a, df);
          sq = fs_dot4(a, a) + fs_dot4(x, x) + fs_dot4(y, y);
        } else {
          part = fs_dot4(tu, dt);
          sq = fs_dot4(tu, tu);
        }
        const float xs = fvx_warp_sum(part) + (bi - bj) + vb;
        const float sqs = fvx_warp_sum(sq);
        float coef = 0.0f;
        if (!dead) {
          const bool inside = (xs >= FVX_CLIP_LO) && (xs <= FVX_CLIP_HI);
          coef = inside ? -1.0f / (1.0f + expf(xs)) : 0.0f;  // d softplus(-x)/dx
          const float z = -fminf(fmaxf(xs, FVX_CLIP_LO), FVX_CLIP_HI);
          const float sp = z > 13.942385f ? z : (z < -13.942385f ? expf(z) : log1pf(expf(z)));
          if (lane == 0) loss_acc += (double)sp + (double)(reg * sqs) + (double)(reg * bi * bi) +
                                     (double)(reg * bj * bj / 10.0f);
          float* gu = M.users.g + (size_t)u * Su;
          float* ggi = M.items.g + (size_t)li * Si;
          float* ggj = M.items.g + (size_t)lj * Si;
          if (lane < K4) {
            fs_red_add4(gu + 4 * lane, make_float4(coef * (x.x - y.x) + reg2 * a.x, coef * (x.y - y.y) + reg2 * a.y,
                                                   coef * (x.z - y.z) + reg2 * a.z, coef * (x.w - y.w) + reg2 * a.w));
            fs_red_add4(ggi + 4 * lane, make_float4(coef * a.x + reg2 * x.x, coef * a.y + reg2 * x.y,
                                                    coef * a.z + reg2 * x.z, coef * a.w + reg2 * x.w));
            fs_red_add4(ggj + 4 * lane, make_float4(-coef * a.x + reg2 * y.x, -coef * a.y + reg2 * y.y,
                                                    -coef * a.z + reg2 * y.z, -coef * a.w + reg2 * y.w));
          }
          if (cv >= 0 && cv < D4)
            fs_red_add4(gu + K + 4 * cv, make_float4(coef * dt.x + reg2 * tu.x, coef * dt.y + reg2 * tu.y,
                                                     coef * dt.z + reg2 * tu.z, coef * dt.w + reg2 * tu.w));
          if (lane == 0) {
            fvx_red_add(ggi + K, coef + reg2 * bi);
            fvx_red_add(ggj + K, -coef + (reg2 / 10.0f) * bj);
          }
        }
        // W rows of the triple: +coef*[Tu | 1 | 0..] (positive slot), the negation (negative slot);
        // staged as [pos hi | pos lo | neg hi | neg lo] x 64 B, then pushed to all 8 CTAs
        if (cv >= 0 && cv < FS_NP / 4) {
          const int n0 = 4 * cv;
          float4 wv = make_float4(coef * tu.x, coef * tu.y, coef * tu.z, coef * tu.w);
          if (n0 == d) wv.x = coef;
          if (n0 + 1 == d) wv.y = coef;
          if (n0 + 2 == d) wv.z = coef;
          if (n0 + 3 == d) wv.w = coef;
          const uint2 h = fs_pack_bf16x4(wv);
          const float4 hf = fs_unpack_bf16x4(h);
          const uint2 l = fs_pack_bf16x4(make_float4(wv.x - hf.x, wv.y - hf.y, wv.z - hf.z, wv.w - hf.w));
          uint2* sg = reinterpret_cast<uint2*>(stg);
          sg[cv] = h;
          sg[8 + cv] = l;
          sg[16 + cv] = make_uint2(h.x ^ 0x80008000u, h.y ^ 0x80008000u);
          sg[24 + cv] = make_uint2(l.x ^ 0x80008000u, l.y ^ 0x80008000u);
        }
        __syncwarp();
        if (warp == 5) FS_STAMP(it, 9);
        {
          const uint32_t dest = (uint32_t)lane >> 2, piece = (uint32_t)lane & 3;
          const uint32_t wb = cl_map(sbase + L::OFF_W + s * L::W_STAGE, dest);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + q * 64 + piece * 16);
            const uint32_t row = (uint32_t)((q >> 1) * HT + lt);
            cl_st_v4(wb + (uint32_t)(q & 1) * L::W_PLANE + row * 64u + ((piece ^ ((row >> 1) & 3u)) << 4), v.x, v.y, v.z, v.w);
          }
          fence_proxy_async_all();
          cl_arrive(cl_map(tc_smem_u32(&w_full[s]), dest));
        }
        if (warp == 5) FS_STAMP(it, 10);
        __syncwarp();                      // the staging buffer is rewritten by the next tile
      }
    }
    if (lane == 0 && loss_acc != 0.0) atomicAdd(M.loss + P.loss_slot, loss_acc);

    // ---- the cluster's partial of gE_ext: features of my slice, all FS_NP columns ----
    if (cnt > 0) {
      mbar_wait(final_b, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int mb = 0; mb < NMB; ++mb) {
      const int f = (int)crank * SL + mb * 128 + quad * 32 + lane;
      float* dst = P.gE_out + ((size_t)cid * M.D + f) * FS_NP;
      float sum[32];
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t v[32];
        if (cnt > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + FS_TMEM_BWD + (mb * 2 + kp) * FS_NP, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[j] = kp == 0 ? __uint_as_float(v[j]) : sum[j] + __uint_as_float(v[j]);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
    }
  }
  // no CTA leaves while a peer may still write into its shared memory
  tc_fence_before();
  __syncthreads();
  cl_sync();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static long long* g_fused_trace = nullptr;
static long long g_fused_trace_cap = 0;
extern "C" int fvx_debug_fused_trace(long long* buf, long long n) {   // profiling aid, not part of fvx.h
  g_fused_trace = buf;
  g_fused_trace_cap = n;
  return 0;
}

// ------------------------------------------------------------------------------------------
bool fvx_fused_eligible(const FvxModel* m) {
  if (!(m->D > 0 && m->use_tensor_cores >= 2)) return false;
  if (m->item_lo != 0 || m->item_cnt != m->num_items) return false;
  if (fvx_tc_np(m->de) != FS_NP) return false;
  if (m->D != 2048 && m->D != 1024) return false;
  if (m->K % 4 != 0 || m->K > 64 || m->d < 1) return false;
  return true;
}

template <int TR, int NST, int SL>
static int launch_fused(const FvxModel* m, const int32_t* user, int B, int loss_slot, int* parts_out, cudaStream_t st) {
  using L = FusedSmem<TR, NST, SL>;
  const size_t smem = L::BYTES + 1024;
  auto kern = k_step_fused<TR, NST, SL>;
  static int max_clusters = -1;
  if (max_clusters < 0) {
    cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) FVX_FAIL(-3, "k_step_fused: cannot set %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    cudaLaunchConfig_t q = {};
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = FS_CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.gridDim = dim3(FS_CL * 32, 1, 1);
    q.blockDim = dim3(FS_THREADS, 1, 1);
    q.dynamicSmemBytes = smem;
    q.attrs = qa; q.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, (const void*)kern, &q);
    if (e != cudaSuccess || n < 1) {
      cudaGetLastError();
      FVX_FAIL(-3, "k_step_fused: no resident cluster of %d CTAs (%s)", FS_CL, cudaGetErrorString(e));
    }
    max_clusters = n;
  }
  FusedParams P;
  P.Fpl = reinterpret_cast<const uint8_t*>(m->F_pl);
  P.user = user; P.B = B; P.loss_slot = loss_slot; P.gE_out = m->gE_part;
  P.n_tiles = (B + TR / 2 - 1) / (TR / 2);
  P.trace = nullptr;
  int ncl = max_clusters < P.n_tiles ? max_clusters : P.n_tiles;
  if (ncl > m->ge_parts) ncl = m->ge_parts;
  FVX_CHECK_ARG(ncl >= 1, "k_step_fused: gE_part has no room");
  *parts_out = ncl;
  if (g_fused_trace && (long long)((P.n_tiles + ncl - 1) / ncl) * 16 <= g_fused_trace_cap) P.trace = g_fused_trace;
  CUtensorMap e_hi, e_lo;
  int rc = 0;
  const uint64_t pitch = (uint64_t)m->D * 2;
  rc |= tc_make_tensor_map_bf16(&e_hi, m->ET_hi, FS_NP, m->D, pitch, 64, FS_NP, 3);
  rc |= tc_make_tensor_map_bf16(&e_lo, m->ET_lo, FS_NP, m->D, pitch, 64, FS_NP, 3);
  if (rc != 0) FVX_FAIL(-4, "k_step_fused: cuTensorMapEncodeTiled failed");
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = FS_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(FS_CL * ncl, 1, 1);
  cfg.blockDim = dim3(FS_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, e_hi, e_lo, *m, P);
  if (e != cudaSuccess) {
    cudaGetLastError();
    FVX_FAIL(-3, "k_step_fused: launch failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

// tile shape: FVX_FUSED_TILE=48 selects 48-row tiles in 3 stages (default: 64 rows, 2 stages)
int fvx_launch_step_fused(const FvxModel* m, const int32_t* user, int B, int loss_slot, int* parts_out,
                          cudaStream_t st) {
  FVX_CHECK_ARG(fvx_fused_eligible(m), "fused step: model not eligible");
  FVX_CHECK_ARG(m->F_pl && m->ET_hi && m->ET_lo && m->gE_part, "fused step: buffers missing");
  static int tile = 0;
  if (tile == 0) {
    const char* s = getenv("FVX_FUSED_TILE");
    tile = (s && atoi(s) == 48) ? 48 : 64;
  }
  if (m->D == 2048) {
    if (tile == 48) return launch_fused<48, 3, 256>(m, user, B, loss_slot, parts_out, st);
    return launch_fused<64, 2, 256>(m, user, B, loss_slot, parts_out, st);
  }
  if (tile == 48) return launch_fused<48, 4, 128>(m, user, B, loss_slot, parts_out, st);
  return launch_fused<64, 4, 128>(m, user, B, loss_slot, parts_out, st);
}
