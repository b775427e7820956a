// The VBPR visual projection and its gradient on the 5th-generation tensor cores
// (tcgen05 + TMEM).  The gathered F rows arrive by 16-byte cp.async (TMA tile::gather4 saturates at 1.67 TB/s,
// profiles/r2_gather_bw.txt); the small operands (E_ext^T, W) by TMA tiles.
//
//   forward : TH[r,:]  = F[rows[r],:] * E_ext          (VBPR.py:83-84: matmul(feature_i, E / Bp))
//   backward: gE_ext   = sum_r F[rows[r],:]^T * W[r,:]  (tape.gradient w.r.t. E, Bp; VBPR.py:141)
//
// fp32-faithful on bf16 tensor cores: F lives in HBM as two bf16 planes (hi = bf16(F),
// lo = bf16(F - hi): 4 bytes per element like fp32, relative error ~2^-17) and the small
// operand (E_ext or W) is split the same way; each product is issued as three MMAs
// (hi*hi + lo*hi + hi*lo) accumulating in fp32 in TMEM.
//
// forward  (k_proj_fwd_tc): work unit = (128-row tile, K split).  A = gathered rows
//   [128 x 64 features] per stage (K-major, 128B swizzle), B = E_ext^T [NP x 64] (K-major),
//   D[128 x NP] in TMEM (double buffered).  The epilogue writes one fp32 partial per K
//   split; the consumer sums them.
// backward (k_grad_E_tc): CTA = (row group, 512-feature group).  The same gathered tile
//   [32 rows x 64 features] x 8 is now read as the MN-major operand A = F^T (M = features),
//   B = W tile [32 rows x NP] (MN-major), D[512 features x NP] stays in TMEM over all row
//   tiles of the group and leaves once as a per-row-group partial of gE_ext.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"
#include "fvx_tc.cuh"

#define PT_BM 128      // rows per forward tile (UMMA M)
#define PT_KC 64       // features per shared-memory chunk: 128 bytes of bf16 = one 128B-swizzle span
#define PT_PROD 128    // producer threads (warps 0-3, cp.async row gathers)
#define PT_THREADS 288 // warps 0-3: producers, warp 4: UMMA issuer, warps 5-8: epilogue
#define GE_RT 32       // rows per backward stage (2 UMMA K steps)
#define GE_FG 512      // features per backward CTA (4 UMMA M blocks)

// Feature planes in HBM ("F_pl"): per item row, per 64-feature chunk, 128 bytes of bf16 hi followed
// by 128 bytes of bf16 lo: [item][D/64][2][64].  One (row, chunk) is 256 contiguous bytes, a whole
// row 4*D bytes - the same footprint as fp32.  Rows are gathered with 16-byte cp.async (measured
// on B200: 5.3-5.8 TB/s for random 4 KB rows; TMA tile::gather4 saturates at 1.67 TB/s and 1-D
// bulk copies of <= 512 B at 1.9 TB/s - scripts/ubench/gather_bw.cu), each thread writing its
// 16 bytes at the 128B-swizzled position the UMMA descriptors expect.
__device__ __forceinline__ void pt_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void pt_cp_arrive(uint64_t* bar) {   // arrives when this thread's copies have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

struct ProjFwdParams {
  const uint8_t* Fpl;   // interleaved planes
  const int32_t* rows;  // nullptr: identity (catalog row row0 + r)
  long long nrows;      // rows the buffers are laid out for
  const int32_t* nrows_dev;   // optional: the number of valid rows lives on the device (<= nrows)
  int row0;
  int D;
  int NP;               // padded output width of THIS launch (multiple of 32): a column slice of the ld columns
  int ld;               // row stride of `out` in floats (the padded width of the whole of E_ext)
  int ksplit;           // K splits
  int chunks;           // 64-feature chunks per split
  int n_tiles;          // 128-row tiles
  int stages;
  int nacc;             // independent accumulators per buffer (one per UMMA K step of a chunk)
  int cat;              // 1: B = [E_hi | E_lo] as ONE operand of N = 2*NP - two UMMAs per K step
                        //    (A_hi, A_lo) instead of three; the epilogue adds the two column halves
  int dyn_ks;           // 1: the K split follows the DEVICE-side row count (fvx_tc_ksplit_rule on *nrows_dev,
                        //    at most `ksplit`): the unique-row step learns its row count on the device
  int chunks_total;     // D / 64
  int nsm;
  float* out;           // [ksplit][nrows][ld], already advanced to the first column of the slice
};

// Work of one CTA of the forward kernel, walked identically by its three roles.
//   units    : (tile, K split) pairs w = blockIdx.x, blockIdx.x + gridDim.x, ... - whole splits
//   stream-K : the flattened (tile, chunk) space is cut into gridDim.x EQUAL contiguous ranges; a CTA's
//              range covers whole tiles plus a head and a tail fragment.  Ranges are at least one tile long
//              (tiles >= CTAs), so a tile is cut at most once: the fragment that starts at chunk 0 writes
//              partial 0, the other one partial 1, and a tile that is not cut gets a zero partial 1.
struct FwdWork {
  int streamk, ksplit, chunks, C, n_units, w, step;
  long long pos, hi;
  __device__ __forceinline__ bool next(int& tile, int& cb, int& ce, int& slot, bool& whole) {
    if (!streamk) {
      if (w >= n_units) return false;
      tile = w / ksplit;
      slot = w - tile * ksplit;
      cb = slot * chunks; ce = cb + chunks; whole = false;
      w += step;
      return true;
    }
    if (pos >= hi) return false;
    tile = (int)(pos / C);
    cb = (int)(pos - (long long)tile * C);
    const long long rem = hi - pos;
    ce = rem < (long long)(C - cb) ? cb + (int)rem : C;
    slot = cb == 0 ? 0 : 1;
    whole = cb == 0 && ce == C;
    pos += ce - cb;
    return true;
  }
  __device__ __forceinline__ int peek_tile() const {     // tile of the fragment next() would return, -1: none
    if (!streamk) return w < n_units ? w / ksplit : -1;
    return pos < hi ? (int)(pos / C) : -1;
  }
};

// ---------------------------------------------------------------------------------
// min-blocks 2 is a REGISTER cap (112 per thread), not an occupancy claim - shared memory admits one CTA per
// SM: the claims / catch-up kernel of the step runs beside this one on the side stream and can only use the
// registers this kernel leaves (148 per thread left room for ONE 256-thread block per SM).
__global__ void __launch_bounds__(PT_THREADS, 2)
k_proj_fwd_tc(const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
              const ProjFwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = PT_BM * 128u;              // one plane of one stage
  const uint32_t b_bytes = (uint32_t)P.NP * 128u;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * stage_bytes);
  uint64_t* full_b = bars;
  uint64_t* empty_b = bars + P.stages;
  uint64_t* t_full = bars + 2 * P.stages;   // [2]
  uint64_t* t_empty = t_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full_b[s], PT_PROD + 1); mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
  }
  // A UMMA that accumulates into the columns its predecessor wrote waits for it; each K step of
  // a chunk therefore owns its own accumulator and the epilogue adds them.
  const uint32_t NW = P.cat ? 2u * P.NP : (uint32_t)P.NP;   // accumulator width in TMEM columns
  uint32_t tcols = 32;
  while (tcols < 2u * P.nacc * NW) tcols <<= 1;
  if (warp == 4) tmem_alloc(tmem_slot, tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long nvalid = P.nrows;
  if (P.nrows_dev) { const long long v = *P.nrows_dev; nvalid = v < nvalid ? v : nvalid; }
  FwdWork work0;
  {
    const long long tiles = (nvalid + PT_BM - 1) / PT_BM;
    int ksplit = P.ksplit, chunks = P.chunks, streamk = 0;
    if (P.dyn_ks) {
      ksplit = fvx_tc_split_dyn(tiles, P.chunks_total, P.nsm, P.ksplit, (int)gridDim.x, &streamk);
      chunks = P.chunks_total / (streamk ? 1 : ksplit);
    }
    work0.streamk = streamk; work0.ksplit = ksplit; work0.chunks = chunks; work0.C = P.chunks_total;
    work0.n_units = (int)tiles * ksplit; work0.w = blockIdx.x; work0.step = gridDim.x;
    const long long total = tiles * P.chunks_total;
    work0.pos = total * blockIdx.x / gridDim.x;
    work0.hi = total * (blockIdx.x + 1) / gridDim.x;
  }

  if (warp < 4) {
    // ===== producers: 16 lanes copy the 256 bytes (hi | lo) of one (row, chunk); a round of the
    //       128 threads covers 8 rows, 16 rounds fill the stage =====
    const int tid = threadIdx.x;
    const int sub = tid >> 4;                 // row within a round
    const int e = tid & 15;                   // 16-byte element of the 256 bytes
    const int plane = e >> 3, c16 = e & 7;
    const size_t row_bytes = (size_t)P.D * 4;
    uint32_t stage = 0, phase = 0;
    // The row indices of a unit are loaded one unit AHEAD: an index load issued behind a full
    // pipeline of cp.async requests takes microseconds, and a producer that waits for it lets the
    // stages run dry (36 % of the producers' samples in the round-1 v4 profile).
    int32_t nxt[16];
    auto load_idx = [&](int tile) {
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const long long r = (long long)tile * PT_BM + it * 8 + sub;
        int32_t item = 0;
        if (r < nvalid) item = P.rows ? __ldg(P.rows + r) : (int32_t)(P.row0 + r);
        nxt[it] = item;
      }
    };
    FwdWork work = work0;
    if (work.peek_tile() >= 0) load_idx(work.peek_tile());
    int tile, cb, ce, slot;
    bool whole;
    while (work.next(tile, cb, ce, slot, whole)) {
      const uint8_t* src[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        long long item = nxt[it];
        if (item < 0) item = 0;               // slot of another rank's item: any valid row, result unused
        src[it] = P.Fpl + (size_t)item * row_bytes + e * 16;
      }
      if (work.peek_tile() >= 0) load_idx(work.peek_tile());
      for (int chunk = cb; chunk < ce; ++chunk) {
        // one poller per warp: 128 threads spinning on the barrier word starve the arrive that flips it
        if (lane == 0) mbar_wait(&empty_b[stage], phase ^ 1);
        __syncwarp();
        const uint32_t sA = tc_smem_u32(smem + (size_t)stage * stage_bytes) + plane * a_bytes;
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int r = it * 8 + sub;
          pt_cp16(sA + r * 128 + ((c16 ^ (r & 7)) << 4), src[it] + (size_t)chunk * 256);
        }
        pt_cp_arrive(&full_b[stage]);
        if (tid == 0) {
          uint8_t* sB_hi = smem + (size_t)stage * stage_bytes + 2 * a_bytes;
          mbar_expect_tx(&full_b[stage], 2 * b_bytes);
          tma_load_2d(sB_hi, &tmB_hi, &full_b[stage], chunk * PT_KC, 0);
          tma_load_2d(sB_hi + b_bytes, &tmB_lo, &full_b[stage], chunk * PT_KC, 0);
        }
        if (++stage == (uint32_t)P.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    // ===== UMMA issuer: the whole warp walks the loop, one elected lane issues =====
    // (issued from an `if (lane == 0)` region a UMMA costs the lone thread ~150-190 cycles - descriptors
    // rebuilt per instruction plus the compiler's ELECT / BRA.U.ANY emulation loop around it - and the eight
    // N = 64 UMMAs of a 64-feature chunk then take as long as the chunk's 32 KB take to arrive from HBM; see
    // fvx_eval_tc.cu and scripts/ubench/umma_chain.cu.  Convergent code, descriptor low words advanced by adds.)
    {
      const uint32_t idesc = umma_idesc_bf16(PT_BM, (int)NW, 0, 0);
      const uint64_t dsc0 = umma_smem_desc(0, 16, 1024, TC_SWZ_128B);
      const uint32_t dlo = (uint32_t)dsc0, dhi = (uint32_t)(dsc0 >> 32);
      const uint32_t s0 = tc_smem_u32(smem) >> 4, st16 = stage_bytes >> 4, a16 = a_bytes >> 4, b16 = b_bytes >> 4;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      FwdWork work = work0;
      int tile, cb, ce, slot;
      bool whole;
      while (work.next(tile, cb, ce, slot, whole)) {
        mbar_wait(&t_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * P.nacc * NW;
        for (int c = 0; c < ce - cb; ++c) {
          mbar_wait(&full_b[stage], phase);
          fence_proxy_async_smem();           // cp.async (generic proxy) writes -> UMMA (async proxy) reads
          tc_fence_after();
          const uint32_t a_hi = dlo + s0 + stage * st16;
          const uint32_t a_lo = a_hi + a16, b_hi = a_lo + a16, b_lo = b_hi + b16;
          // pass-major, K-step-minor: consecutive UMMAs target different accumulators
          if (P.cat) {
            // the hi and lo tiles of E_ext^T are adjacent in shared memory: one descriptor of 2*NP rows
            // reads both, so A_hi and A_lo are fetched once each
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
              for (int k = 0; k < PT_KC / 16; ++k) {
                const uint32_t d = d0 + (uint32_t)(k % P.nacc) * NW;
                const bool first = c == 0 && pass == 0 && k < P.nacc;
                umma_f16_lohi_elect(d, (pass == 1 ? a_lo : a_hi) + 2u * k, dhi, b_hi + 2u * k, dhi, idesc, first ? 0u : 1u);
              }
            }
          } else {
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
              for (int k = 0; k < PT_KC / 16; ++k) {
                const uint32_t d = d0 + (uint32_t)(k % P.nacc) * P.NP;
                const bool first = c == 0 && pass == 0 && k < P.nacc;
                umma_f16_lohi_elect(d, (pass == 1 ? a_lo : a_hi) + 2u * k, dhi, (pass == 2 ? b_lo : b_hi) + 2u * k, dhi,
                                    idesc, first ? 0u : 1u);
              }
            }
          }
          umma_commit_elect(&empty_b[stage]);
          if (++stage == (uint32_t)P.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_elect(&t_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> fp32 partial rows =====
    const int quad = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    FwdWork work = work0;
    int tile, cb, ce, slot;
    bool whole;
    while (work.next(tile, cb, ce, slot, whole)) {
      const long long row = (long long)tile * PT_BM + quad * 32 + lane;
      mbar_wait(&t_full[acc], acc_phase);
      tc_fence_after();
      float* dst = P.out + ((size_t)slot * P.nrows + (size_t)(row < nvalid ? row : 0)) * P.ld;
      if (whole && row < nvalid) {            // stream-K: an uncut tile has no second fragment - its partial 1 is zero
        float4* z = reinterpret_cast<float4*>(dst + (size_t)P.nrows * P.ld);
        for (int n4 = 0; n4 < (P.NP >> 2); ++n4) z[n4] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int n0 = 0; n0 < P.NP; n0 += 32) {
        float sum[32];
        const int nparts = P.cat ? 2 * P.nacc : P.nacc;   // cat: columns [0,NP) and [NP,2NP) of every accumulator
        for (int a = 0; a < nparts; ++a) {
          uint32_t v[32];
          const uint32_t col = P.cat ? (uint32_t)(acc * P.nacc + (a >> 1)) * NW + (uint32_t)(a & 1) * P.NP
                                     : (uint32_t)(acc * P.nacc + a) * P.NP;
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + col + n0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] = a == 0 ? __uint_as_float(v[j]) : sum[j] + __uint_as_float(v[j]);
        }
        if (row < nvalid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + n0 + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tcols);
  }
}

// ---------------------------------------------------------------------------------
struct GradEParams {
  const uint8_t* Fpl;
  const int32_t* rows;
  long long nrows;
  const int32_t* nrows_dev;   // optional: the number of valid rows lives on the device (<= nrows)
  int n_groups;         // row groups of the launch
  int D, NP;            // NP: padded width of THIS launch (a column slice of the ld columns of W / of gE_part)
  int ld;               // row stride of `out` in floats
  int fgs;              // features per CTA (multiple of 128)
  int nfg;              // feature groups = D / fgs
  int rows_per_group;   // multiple of GE_RT (recomputed on the device when nrows_dev is set)
  int stages;
  int w_atoms;          // 64-column sub-tiles of the W tile (1 for NP <= 64)
  int w_sw;             // swizzle of the W tile: TC_SWZ_64B (NP == 32) or TC_SWZ_128B
  int cat;              // 1 (NP == 32, interleaved W planes): the W tile is [32 rows x (hi 32 | lo 32)], ONE operand
                        //    of N = 64 - two UMMAs per K step (F_hi^T, F_lo^T) instead of three
  float* out;           // [row groups][D][ld], already advanced to the first column of the slice
};

__global__ void __launch_bounds__(PT_THREADS, 1)
k_grad_E_tc(const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
            const GradEParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = P.fgs / PT_KC;                          // 64-feature chunks per stage
  const uint32_t chunk_bytes = GE_RT * 128u;                 // [32 rows x 128 B]
  const uint32_t a_bytes = (uint32_t)nchunk * chunk_bytes;   // one plane
  const uint32_t wrow_bytes = P.cat ? 128u : (P.NP <= 64 ? (uint32_t)P.NP * 2u : 128u);
  const uint32_t watom_bytes = GE_RT * wrow_bytes;
  const uint32_t w_bytes = (uint32_t)P.w_atoms * watom_bytes;  // one plane (cat: both planes, one tile)
  const uint32_t stage_bytes = 2 * a_bytes + (P.cat ? 1u : 2u) * w_bytes;
  const uint32_t NW = P.cat ? 2u * P.NP : (uint32_t)P.NP;      // accumulator width in TMEM columns
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * stage_bytes);
  uint64_t* full_b = bars;
  uint64_t* empty_b = bars + P.stages;
  uint64_t* t_full = bars + 2 * P.stages;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full_b[s], PT_PROD + 1); mbar_init(&empty_b[s], 1); }
    mbar_init(t_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
  }
  const int nmb = P.fgs / 128;
  uint32_t tcols = 32;
  while (tcols < (uint32_t)nmb * NW) tcols <<= 1;
  if (warp == 4) tmem_alloc(tmem_slot, tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int rg = blockIdx.x / P.nfg, fg = blockIdx.x - rg * P.nfg;
  long long nvalid = P.nrows;
  long long rpg = P.rows_per_group;
  if (P.nrows_dev) {
    const long long v = *P.nrows_dev;
    nvalid = v < nvalid ? v : nvalid;
    rpg = (nvalid + P.n_groups - 1) / P.n_groups;
    rpg = (rpg + GE_RT - 1) / GE_RT * GE_RT;
    if (rpg < GE_RT) rpg = GE_RT;
  }
  const long long r_begin = (long long)rg * rpg;
  long long r_end = r_begin + rpg;
  if (r_end > nvalid) r_end = nvalid;
  const int n_tiles = r_end > r_begin ? (int)((r_end - r_begin + GE_RT - 1) / GE_RT) : 0;

  if (warp < 4) {
    // ===== producers: a warp copies the (fgs*4)-byte slice of one row per pass, 8 rows per stage =====
    const int tid = threadIdx.x;
    const size_t row_bytes = (size_t)P.D * 4;
    const int per_row = P.fgs / 4;            // 16-byte elements of one row slice (both planes)
    uint32_t stage = 0, phase = 0;
    // row indices one tile ahead (see the forward kernel: an index load queued behind the cp.async
    // traffic is slow, and here it used to sit in front of EVERY 32-row tile)
    int32_t nxt[GE_RT / 4];
    auto load_idx = [&](int t) {
      const long long r0 = r_begin + (long long)t * GE_RT;
#pragma unroll
      for (int j = 0; j < GE_RT / 4; ++j) {
        const long long r = r0 + j * 4 + warp;
        int32_t item = 0;
        if (r < nvalid) item = P.rows ? __ldg(P.rows + r) : (int32_t)r;
        nxt[j] = item;
      }
    };
    if (n_tiles > 0) load_idx(0);
    for (int t = 0; t < n_tiles; ++t) {
      const long long r0 = r_begin + (long long)t * GE_RT;
      const uint8_t* src[GE_RT / 4];
#pragma unroll
      for (int j = 0; j < GE_RT / 4; ++j) {
        long long item = nxt[j];
        if (item < 0) item = 0;
        src[j] = P.Fpl + (size_t)item * row_bytes + (size_t)fg * P.fgs * 4;
      }
      if (t + 1 < n_tiles) load_idx(t + 1);
      if (lane == 0) mbar_wait(&empty_b[stage], phase ^ 1);   // one poller per warp
      __syncwarp();
      const uint32_t sA = tc_smem_u32(smem + (size_t)stage * stage_bytes);
      for (int e = lane; e < per_row; e += 32) {
        const int c = e >> 4, plane = (e >> 3) & 1, c16 = e & 7;
        const uint32_t doff = (uint32_t)plane * a_bytes + (uint32_t)c * chunk_bytes;
#pragma unroll
        for (int j = 0; j < GE_RT / 4; ++j) {
          const int r = j * 4 + warp;
          pt_cp16(sA + doff + r * 128 + ((c16 ^ (r & 7)) << 4), src[j] + (size_t)e * 16);
        }
      }
      pt_cp_arrive(&full_b[stage]);
      if (tid == 0) {
        uint8_t* sW_hi = smem + (size_t)stage * stage_bytes + 2 * a_bytes;
        mbar_expect_tx(&full_b[stage], (P.cat ? 1u : 2u) * w_bytes);
        if (P.cat) {
          tma_load_2d(sW_hi, &tmW_hi, &full_b[stage], 0, (int)r0);   // [32 rows x 128 B]: hi | lo of every row
        } else
        for (int a = 0; a < P.w_atoms; ++a) {
          tma_load_2d(sW_hi + (size_t)a * watom_bytes, &tmW_hi, &full_b[stage], a * 64, (int)r0);
          tma_load_2d(sW_hi + w_bytes + (size_t)a * watom_bytes, &tmW_lo, &full_b[stage], a * 64, (int)r0);
        }
      }
      if (++stage == (uint32_t)P.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 4) {
    // ===== UMMA issuer: D[mb][128 features x NP] += F^T[128 x 16 rows] * W[16 rows x NP] =====
    // (the whole warp walks the loop, one elected lane issues; descriptor low words advanced by adds - see the
    // forward kernel)
    {
      const uint32_t idesc = umma_idesc_bf16(128, (int)NW, 1, 1);
      const uint32_t w_sbo = 8u * wrow_bytes;
      const uint64_t da0 = umma_smem_desc(0, chunk_bytes, 1024, TC_SWZ_128B);
      const uint64_t dw0 = umma_smem_desc(0, watom_bytes, w_sbo, P.cat ? TC_SWZ_128B : P.w_sw);
      const uint32_t alo0 = (uint32_t)da0, ahi = (uint32_t)(da0 >> 32), wlo0 = (uint32_t)dw0, whi = (uint32_t)(dw0 >> 32);
      const uint32_t s0 = tc_smem_u32(smem) >> 4, st16 = stage_bytes >> 4, a16 = a_bytes >> 4, w16 = w_bytes >> 4;
      const uint32_t mb16 = (2u * chunk_bytes) >> 4, k16a = (16u * 128u) >> 4, k16w = (16u * wrow_bytes) >> 4;
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(&full_b[stage], phase);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint32_t a_hi = s0 + stage * st16;
        const uint32_t a_lo = a_hi + a16, w_hi = a_lo + a16, w_lo = w_hi + w16;
        // (K step, pass)-major, M-block-minor: consecutive UMMAs target different accumulators
        if (P.cat) {
#pragma unroll
          for (int k = 0; k < GE_RT / 16; ++k) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
              const uint32_t dw = wlo0 + w_hi + k * k16w;
              uint32_t da = alo0 + (pass == 1 ? a_lo : a_hi) + k * k16a;
              for (int mb = 0; mb < nmb; ++mb, da += mb16)
                umma_f16_lohi_elect(tmem_base + mb * NW, da, ahi, dw, whi, idesc, (t | k | pass) ? 1u : 0u);
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < GE_RT / 16; ++k) {
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t dw = wlo0 + (pass == 2 ? w_lo : w_hi) + k * k16w;
              uint32_t da = alo0 + (pass == 1 ? a_lo : a_hi) + k * k16a;
              for (int mb = 0; mb < nmb; ++mb, da += mb16)
                umma_f16_lohi_elect(tmem_base + mb * P.NP, da, ahi, dw, whi, idesc, (t | k | pass) ? 1u : 0u);
            }
          }
        }
        umma_commit_elect(&empty_b[stage]);
        if (++stage == (uint32_t)P.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_elect(t_full);
    }
  } else {
    // ===== epilogue: the CTA's partial of gE_ext =====
    const int quad = warp & 3;
    if (n_tiles > 0) {
      mbar_wait(t_full, 0);
      tc_fence_after();
    }
    for (int mb = 0; mb < nmb; ++mb) {
      const int f = fg * P.fgs + mb * 128 + quad * 32 + lane;
      float* dst = P.out + ((size_t)rg * P.D + f) * P.ld;
      for (int n0 = 0; n0 < P.NP; n0 += 32) {
        uint32_t v[32];
        if (n_tiles > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + mb * NW + n0, v);
          tmem_ld_wait();
          if (P.cat) {                      // columns [NP, 2NP): the products with W_lo
            uint32_t v2[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + mb * NW + P.NP + n0, v2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + n0 + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                          __uint_as_float(v[j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tcols);
  }
}

// fp32 rows [n, D] -> interleaved bf16 planes [n][D/64][2][64]
__global__ void k_split_planes(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n_rows, int D) {
  const long long total = n_rows * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D;
    const int f = (int)(i - r * D);
    const float x = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const size_t o = (size_t)r * 2 * D + (size_t)(f >> 6) * 128 + (f & 63);
    dst[o] = h;
    dst[o + 64] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

// ---------------------------------------------------------------------------------
// E_ext [D, de] fp32 -> ET_hi / ET_lo [NP, D] bf16 (row n = column n of E_ext, rows >= de zero)
__global__ void k_split_E(const float* __restrict__ E, int D, int de, int NP, __nv_bfloat16* __restrict__ hi,
                          __nv_bfloat16* __restrict__ lo) {
  const int total = NP * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / D, f = i - n * D;
    const float x = n < de ? E[(size_t)f * de + n] : 0.0f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

// ---------------------------------------------------------------------------------
int fvx_tc_np(int de) { return de <= 32 ? 32 : (de + 63) / 64 * 64; }

// One launch of either tensor-core kernel covers at most 256 columns of E_ext (UMMA N, the TMA box, the TMEM
// columns of the accumulators); a wider model (BASELINE configs[4] with embed_d = 256: 257 columns, NP = 320)
// is cut into balanced COLUMN SLICES of a multiple of 64 columns, one launch each.  Every launch re-reads the
// gathered rows, which costs little here: with three hi/lo passes of N >= 128 the tensor pipe, not HBM, paces
// a 64-feature stage, so two launches of half the width take about as long as one launch of the whole would.
#define TC_MAX_SLICES 4
static int tc_slices(int NP, int* c0, int* w) {
  if (NP <= 256) { c0[0] = 0; w[0] = NP; return 1; }
  const int n = (NP + 255) / 256;
  const int per = (NP / n + 63) / 64 * 64;
  int k = 0;
  for (int c = 0; c < NP && k < TC_MAX_SLICES; c += per, ++k) { c0[k] = c; w[k] = NP - c < per ? NP - c : per; }
  return k;
}

int fvx_tc_ksplit(const FvxModel* m, long long nrows) {
  return fvx_tc_ksplit_rule((nrows + PT_BM - 1) / PT_BM, m->D / PT_KC, fvx_num_sms(), 1 << 30);
}


int fvx_launch_split_E(const FvxModel* m, cudaStream_t st) {
  FVX_CHECK_ARG(m->ET_hi && m->ET_lo, "tensor-core projection: ET planes missing");
  const int NP = fvx_tc_np(m->de);
  k_split_E<<<(NP * m->D + 255) / 256, 256, 0, st>>>(m->E, m->D, m->de, NP, reinterpret_cast<__nv_bfloat16*>(m->ET_hi),
                                                     reinterpret_cast<__nv_bfloat16*>(m->ET_lo));
  FVX_CHECK_LAUNCH("k_split_E");
  return 0;
}

// out: [ksplit][nrows][NP] fp32 partials (ksplit from fvx_tc_ksplit)
int fvx_launch_project_tc(const FvxModel* m, const int32_t* rows, int row0, int64_t nrows, int ksplit, float* out,
                          cudaStream_t st, const int32_t* nrows_dev, int dyn_ks, int sm_reserve) {
  FVX_CHECK_ARG(m->F_pl && m->ET_hi && m->ET_lo, "tensor-core projection: bf16 planes missing");
  FVX_CHECK_ARG(m->D % PT_KC == 0, "tensor-core projection: D=%d must be a multiple of %d", m->D, PT_KC);
  const int NPT = fvx_tc_np(m->de);
  FVX_CHECK_ARG(NPT <= 256 * TC_MAX_SLICES, "tensor-core projection: d+1=%d too wide", m->de);
  if (nrows <= 0) return 0;
  const int chunks_total = m->D / PT_KC;
  FVX_CHECK_ARG(ksplit >= 1 && (dyn_ks || chunks_total % ksplit == 0), "tensor-core projection: bad K split %d", ksplit);
  FVX_CHECK_ARG(!dyn_ks || nrows_dev != nullptr, "tensor-core projection: a device-side K split needs nrows_dev");
  int sl_c0[TC_MAX_SLICES], sl_w[TC_MAX_SLICES];
  const int n_slices = tc_slices(NPT, sl_c0, sl_w);
  for (int sl = 0; sl < n_slices; ++sl) {
    const int NP = sl_w[sl], c0 = sl_c0[sl];
    CUtensorMap b_hi, b_lo;
    int rc = 0;
    const uint64_t pitch = (uint64_t)m->D * 2;
    rc |= tc_make_tensor_map_bf16(&b_hi, m->ET_hi + (size_t)c0 * m->D, NP, m->D, pitch, PT_KC, NP, 3);
    rc |= tc_make_tensor_map_bf16(&b_lo, m->ET_lo + (size_t)c0 * m->D, NP, m->D, pitch, PT_KC, NP, 3);
    if (rc != 0) FVX_FAIL(-4, "tensor-core projection: cuTensorMapEncodeTiled failed");
    ProjFwdParams P;
    P.Fpl = reinterpret_cast<const uint8_t*>(m->F_pl); P.D = m->D;
    P.rows = rows; P.row0 = row0; P.nrows = nrows; P.nrows_dev = nrows_dev; P.NP = NP; P.ld = NPT; P.ksplit = ksplit;
    P.chunks = dyn_ks ? chunks_total : chunks_total / ksplit;
    P.dyn_ks = dyn_ks; P.chunks_total = chunks_total; P.nsm = fvx_num_sms();
    P.n_tiles = (int)((nrows + PT_BM - 1) / PT_BM);
    P.out = out + c0;
    P.cat = NP <= 64 ? 1 : 0;                       // N = 2*NP <= 128 and 2 buffers x nacc x 2*NP <= 512 columns
    {
      const int nw = P.cat ? 2 * NP : NP;
      P.nacc = 256 / nw < 4 ? (256 / nw < 1 ? 1 : 256 / nw) : 4;
    }
    const size_t stage_bytes = 2 * PT_BM * 128 + 2 * (size_t)NP * 128;
    int stages = (int)((220 * 1024) / stage_bytes);
    if (stages > 6) stages = 6;
    {
      static int cap = -1;                 // FVX_FWD_STAGES: cap on the pipeline depth (measurements)
      if (cap < 0) { const char* e = getenv("FVX_FWD_STAGES"); cap = e ? atoi(e) : 0; }
      if (cap >= 2 && stages > cap) stages = cap;
      // the step's forward kernel shares the memory system with the catch-up kernel on the side stream:
      // six 32 KB stages per SM keep ~28 MB of requests queued and the latency-bound neighbour waits behind
      // them (its end moved from 146 to 138 us of the step with four stages; the projection itself did not slow)
      if (cap < 2 && dyn_ks && stages > 4) stages = 4;
    }
    FVX_CHECK_ARG(stages >= 2, "tensor-core projection: tile does not fit shared memory");
    P.stages = stages;
    const size_t smem = stages * stage_bytes + (2 * stages + 4) * 8 + 16 + 1024;
    static FvxSmemMark fwd_smem;
    if (int r = fvx_ensure_smem((const void*)k_proj_fwd_tc, &fwd_smem, smem, "k_proj_fwd_tc")) return r;
    long long grid = (long long)P.n_tiles * ksplit;
    int sms = fvx_num_sms() - sm_reserve;
    if (sms < 8) sms = 8;
    if (grid > sms || dyn_ks) grid = sms;
    P.nsm = sms;
    k_proj_fwd_tc<<<(int)grid, PT_THREADS, smem, st>>>(b_hi, b_lo, P);
    FVX_CHECK_LAUNCH("k_proj_fwd_tc");
  }
  return 0;
}

// gE_part[p][D][NP] = partial sums; W planes [nrows][NP] bf16.  *parts_out = row groups written.
int fvx_launch_grad_E_tc(const FvxModel* m, const int32_t* rows, int64_t nrows, int* parts_out, cudaStream_t st,
                         const int32_t* nrows_dev, int sm_reserve) {
  FVX_CHECK_ARG(m->F_pl && m->W_hi && m->W_lo && m->gE_part, "tensor-core grad_E: buffers missing");
  FVX_CHECK_ARG(m->D % 128 == 0, "tensor-core grad_E: D=%d must be a multiple of 128", m->D);
  const int NPT = fvx_tc_np(m->de);
  FVX_CHECK_ARG(NPT <= 256 * TC_MAX_SLICES, "tensor-core grad_E: d+1=%d too wide", m->de);
  int sl_c0[TC_MAX_SLICES], sl_w[TC_MAX_SLICES];
  const int n_slices = tc_slices(NPT, sl_c0, sl_w);
  // the row groups (= partials the E update sums) are the same for every slice: the widest slice decides
  int fgs_all = GE_FG;
  for (int sl = 0; sl < n_slices; ++sl) {
    int fgs = GE_FG;
    while (fgs > 128 && (m->D % fgs != 0 || (fgs / 128) * sl_w[sl] > 512)) fgs >>= 1;
    FVX_CHECK_ARG(m->D % fgs == 0 && (fgs / 128) * sl_w[sl] <= 512, "tensor-core grad_E: d+1=%d too wide", m->de);
    if (fgs < fgs_all) fgs_all = fgs;
  }
  const int fgs = fgs_all;
  const int nfg = m->D / fgs;
  int max_rg = (fvx_num_sms() - sm_reserve) / nfg;
  if (max_rg < 1) max_rg = 1;
  if (max_rg > m->ge_parts) max_rg = m->ge_parts;
  long long rpg = (nrows + max_rg - 1) / max_rg;
  rpg = (rpg + GE_RT - 1) / GE_RT * GE_RT;
  if (rpg < GE_RT) rpg = GE_RT;
  const int parts = nrows > 0 ? (int)((nrows + rpg - 1) / rpg) : 0;
  *parts_out = parts;
  if (parts == 0) return 0;
  const int wpitch = fvx_w_pitch(m);
  for (int sl = 0; sl < n_slices; ++sl) {
    const int NP = sl_w[sl], c0 = sl_c0[sl];
    CUtensorMap w_hi, w_lo;
    int rc = 0;
    const int cat = (NPT == 32 && wpitch == 2 * NP) ? 1 : 0;
    const int wbox = NP <= 64 ? NP : 64;
    const int wsw = NP == 32 ? 2 : 3;
    if (cat) {
      // one map over the interleaved rows [nrows, hi 32 | lo 32]: 128-byte rows, 128B swizzle
      rc |= tc_make_tensor_map_bf16(&w_hi, m->W_hi, nrows, 2 * NP, (uint64_t)wpitch * 2, 2 * NP, GE_RT, 3);
      w_lo = w_hi;
    } else {
      rc |= tc_make_tensor_map_bf16(&w_hi, m->W_hi + c0, nrows, NP, (uint64_t)wpitch * 2, wbox, GE_RT, wsw);
      rc |= tc_make_tensor_map_bf16(&w_lo, m->W_lo + c0, nrows, NP, (uint64_t)wpitch * 2, wbox, GE_RT, wsw);
    }
    if (rc != 0) FVX_FAIL(-4, "tensor-core grad_E: cuTensorMapEncodeTiled failed");
    GradEParams P;
    P.Fpl = reinterpret_cast<const uint8_t*>(m->F_pl);
    P.rows = rows; P.nrows = nrows; P.nrows_dev = nrows_dev; P.n_groups = parts; P.D = m->D; P.NP = NP; P.ld = NPT;
    P.fgs = fgs; P.nfg = nfg; P.rows_per_group = (int)rpg;
    P.w_atoms = NP <= 64 ? 1 : NP / 64;
    P.w_sw = NP == 32 ? TC_SWZ_64B : TC_SWZ_128B;
    P.cat = cat;
    P.out = m->gE_part + c0;
    const size_t wrow = cat ? 128 : (NP <= 64 ? (size_t)NP * 2 : 128);
    const size_t stage_bytes = 2 * (size_t)(fgs / PT_KC) * GE_RT * 128 + (cat ? 1 : 2) * (size_t)P.w_atoms * GE_RT * wrow;
    int stages = (int)((220 * 1024) / stage_bytes);
    if (stages > 4) stages = 4;
    {
      static int cap = -1;                 // FVX_GE_STAGES: cap on the pipeline depth (measurements)
      if (cap < 0) { const char* e = getenv("FVX_GE_STAGES"); cap = e ? atoi(e) : 0; }
      if (cap >= 2 && stages > cap) stages = cap;
    }
    FVX_CHECK_ARG(stages >= 2, "tensor-core grad_E: tile does not fit shared memory");
    P.stages = stages;
    const size_t smem = stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
    static FvxSmemMark ge_smem;
    if (int r = fvx_ensure_smem((const void*)k_grad_E_tc, &ge_smem, smem, "k_grad_E_tc")) return r;
    k_grad_E_tc<<<parts * nfg, PT_THREADS, smem, st>>>(w_hi, w_lo, P);
    FVX_CHECK_LAUNCH("k_grad_E_tc");
  }
  return 0;
}

// out[r, 0:de] = sum over the K-split partials part[s][r][0:de]
__global__ void k_reduce_partials(const float* __restrict__ part, long long nrows, int NP, int ks, int de,
                                  float* __restrict__ out) {
  const long long total = nrows * de;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / de;
    const int c = (int)(i - r * de);
    float v = 0.0f;
    for (int s = 0; s < ks; ++s) v += part[((size_t)s * nrows + r) * NP + c];
    out[i] = v;
  }
}

int fvx_launch_reduce_partials(const float* part, long long nrows, int NP, int ks, int de, float* out,
                               cudaStream_t st) {
  if (nrows <= 0) return 0;
  long long g = (nrows * de + 255) / 256;
  if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
  k_reduce_partials<<<(int)g, 256, 0, st>>>(part, nrows, NP, ks, de, out);
  FVX_CHECK_LAUNCH("k_reduce_partials");
  return 0;
}

// W fp32 [nrows, de] -> bf16 planes [nrows, NP] (rows with rows[r] < 0 and columns >= de: zero)
__global__ void k_split_W(const float* __restrict__ W, const int32_t* __restrict__ rows, long long nrows, int de,
                          int NP, int pitch, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = nrows * NP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / NP;
    const int c = (int)(i - r * NP);
    float x = 0.0f;
    if (c < de && (rows == nullptr || rows[r] >= 0)) x = W[r * de + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[r * pitch + c] = h;
    lo[r * pitch + c] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

int fvx_launch_split_W(const FvxModel* m, const float* W, const int32_t* rows, long long nrows, cudaStream_t st) {
  if (nrows <= 0) return 0;
  const int NP = fvx_tc_np(m->de);
  long long g = (nrows * NP + 255) / 256;
  if (g > (long long)fvx_num_sms() * 8) g = (long long)fvx_num_sms() * 8;
  k_split_W<<<(int)g, 256, 0, st>>>(W, rows, nrows, m->de, NP, fvx_w_pitch(m), reinterpret_cast<__nv_bfloat16*>(m->W_hi),
                                    reinterpret_cast<__nv_bfloat16*>(m->W_lo));
  FVX_CHECK_LAUNCH("k_split_W");
  return 0;
}

// out[D, de] = sum over the row-group partials gE_part[p][D][gnp]
__global__ void k_reduce_gE(const float* __restrict__ part, int parts, int D, int gnp, int de, float* __restrict__ out) {
  const int n = D * de;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int f = i / de, c = i - f * de;
    float g = 0.0f;
    for (int p = 0; p < parts; ++p) g += part[((size_t)p * D + f) * gnp + c];
    out[i] = g;
  }
}

int fvx_launch_reduce_gE(const FvxModel* m, int parts, int gnp, float* out, cudaStream_t st) {
  k_reduce_gE<<<(m->D * m->de + 255) / 256, 256, 0, st>>>(m->gE_part, parts, m->D, gnp, m->de, out);
  FVX_CHECK_LAUNCH("k_reduce_gE");
  return 0;
}

int fvx_launch_split_planes(const float* src, uint16_t* dst, long long n_rows, int D, cudaStream_t st) {
  FVX_CHECK_ARG(D % PT_KC == 0, "feature planes: D=%d must be a multiple of %d", D, PT_KC);
  if (n_rows <= 0) return 0;
  long long g = (n_rows * D + 255) / 256;
  if (g > (long long)fvx_num_sms() * 16) g = (long long)fvx_num_sms() * 16;
  k_split_planes<<<(int)g, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n_rows, D);
  FVX_CHECK_LAUNCH("k_split_planes");
  return 0;
}
