// The VBPR visual projection and its gradient, fp32 CUDA-core version.
//   forward : TH[r,:]  = F[rows[r],:] * E_ext          (VBPR.py:83-84: matmul(feature_i, E / Bp))
//   backward: gE_ext   = sum_r F[rows[r],:]^T * W[r,:]  (tape.gradient w.r.t. E, Bp; VBPR.py:141)
// E_ext[D,de] holds E in columns [0,d) and Bp in column d.  This is the exact-fp32
// path (and the on-device check for the tcgen05 path in fvx_project_tc.cu).
#include <cuda_bf16.h>

#include "fvx_common.cuh"
#include "fvx_kernels.cuh"

#define PJ_TM 64
#define PJ_TK 32
#define PJ_LD 36

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int NCOL>
__global__ void __launch_bounds__(256)
k_project(const float* __restrict__ F, const int32_t* __restrict__ rows, long long nrows, int D,
          const float* __restrict__ E, int de, float* __restrict__ out) {
  constexpr int NC = 16 * NCOL;
  __shared__ __align__(16) float As[2][PJ_TM][PJ_LD];
  __shared__ __align__(16) float Bs[2][NC][PJ_LD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long row0 = (long long)blockIdx.x * PJ_TM;
  const int col0 = blockIdx.y * NC;

  const float* arow[2];
  int a_r[2], a_k[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int q = tid + 256 * s;
    a_r[s] = q >> 3;
    a_k[s] = (q & 7) * 4;
    const long long r = row0 + a_r[s];
    long long item = -1;
    if (r < nrows) item = rows ? (long long)rows[r] : r;
    arow[s] = item >= 0 ? F + (size_t)item * D : nullptr;
  }
  auto load_tiles = [&](int buf, int k0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      float* dst = &As[buf][a_r[s]][a_k[s]];
      if (arow[s] != nullptr && k0 + a_k[s] < D) cp_async16(dst, arow[s] + k0 + a_k[s]);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int e = tid; e < PJ_TK * NC; e += 256) {
      const int k = e / NC, n = e - k * NC;
      float v = 0.0f;
      if (k0 + k < D && col0 + n < de) v = E[(size_t)(k0 + k) * de + col0 + n];
      Bs[buf][n][k] = v;
    }
    cp_async_commit();
  };

  float acc[4][NCOL];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NCOL; ++j) acc[i][j] = 0.0f;

  const int nchunks = (D + PJ_TK - 1) / PJ_TK;
  load_tiles(0, 0);
  cp_async_wait_all();
  __syncthreads();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) load_tiles(buf ^ 1, (ch + 1) * PJ_TK);
#pragma unroll
    for (int k4 = 0; k4 < PJ_TK / 4; ++k4) {
      float4 a[4], b[NCOL];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&As[buf][ty + 16 * i][k4 * 4]);
#pragma unroll
      for (int j = 0; j < NCOL; ++j) b[j] = *reinterpret_cast<const float4*>(&Bs[buf][tx + 16 * j][k4 * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
    cp_async_wait_all();
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = row0 + ty + 16 * i;
    if (r >= nrows) continue;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
      const int c = col0 + tx + 16 * j;
      if (c < de) out[(size_t)r * de + c] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------
#define GE_COLS 512
#define GE_NB 24
#define GE_RT 32

__global__ void __launch_bounds__(256)
k_grad_E(const float* __restrict__ F, const int32_t* __restrict__ rows, long long nrows, int D,
         const float* __restrict__ W, int de, float* __restrict__ gE_part, long long rows_per_part) {
  __shared__ __align__(16) float Ws[GE_RT][GE_NB];
  __shared__ long long its[GE_RT];
  const int tid = threadIdx.x;
  const int ka = blockIdx.x * GE_COLS + tid, kb = ka + 256;
  const int p = blockIdx.y, n0 = blockIdx.z * GE_NB;
  const long long r_begin = (long long)p * rows_per_part;
  const long long r_end = (r_begin + rows_per_part < nrows) ? r_begin + rows_per_part : nrows;
  float acc_a[GE_NB], acc_b[GE_NB];
#pragma unroll
  for (int n = 0; n < GE_NB; ++n) { acc_a[n] = 0.0f; acc_b[n] = 0.0f; }

  for (long long rt = r_begin; rt < r_end; rt += GE_RT) {
    __syncthreads();
    for (int e = tid; e < GE_RT * GE_NB; e += 256) {
      const int rr = e / GE_NB, n = e - rr * GE_NB;
      const long long r = rt + rr;
      // slots whose item this rank does not own carry no W row (and no F row)
      const bool live = r < r_end && n0 + n < de && (rows == nullptr || rows[r] >= 0);
      Ws[rr][n] = live ? W[(size_t)r * de + n0 + n] : 0.0f;
    }
    if (tid < GE_RT) {
      const long long r = rt + tid;
      long long it = -1;
      if (r < r_end) it = rows ? (long long)rows[r] : r;
      its[tid] = it;
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < GE_RT / 8; ++g) {
      float fa[8], fb[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const long long it = its[g * 8 + q];
        fa[q] = (it >= 0 && ka < D) ? __ldcs(F + (size_t)it * D + ka) : 0.0f;
        fb[q] = (it >= 0 && kb < D) ? __ldcs(F + (size_t)it * D + kb) : 0.0f;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int n4 = 0; n4 < GE_NB / 4; ++n4) {
          const float4 w = *reinterpret_cast<const float4*>(&Ws[g * 8 + q][n4 * 4]);
          acc_a[n4 * 4 + 0] = fmaf(fa[q], w.x, acc_a[n4 * 4 + 0]);
          acc_a[n4 * 4 + 1] = fmaf(fa[q], w.y, acc_a[n4 * 4 + 1]);
          acc_a[n4 * 4 + 2] = fmaf(fa[q], w.z, acc_a[n4 * 4 + 2]);
          acc_a[n4 * 4 + 3] = fmaf(fa[q], w.w, acc_a[n4 * 4 + 3]);
          acc_b[n4 * 4 + 0] = fmaf(fb[q], w.x, acc_b[n4 * 4 + 0]);
          acc_b[n4 * 4 + 1] = fmaf(fb[q], w.y, acc_b[n4 * 4 + 1]);
          acc_b[n4 * 4 + 2] = fmaf(fb[q], w.z, acc_b[n4 * 4 + 2]);
          acc_b[n4 * 4 + 3] = fmaf(fb[q], w.w, acc_b[n4 * 4 + 3]);
        }
      }
    }
  }
  float* dst = gE_part + (size_t)p * D * de;
#pragma unroll
  for (int n = 0; n < GE_NB; ++n) {
    if (n0 + n < de) {
      if (ka < D) dst[(size_t)ka * de + n0 + n] = acc_a[n];
      if (kb < D) dst[(size_t)kb * de + n0 + n] = acc_b[n];
    }
  }
}

// ---------------------------------------------------------------------------------
int fvx_launch_project(const FvxModel* m, const int32_t* rows, int64_t nrows, float* out, cudaStream_t st) {
  FVX_CHECK_ARG(m->D % 4 == 0, "projection: D=%d must be a multiple of 4", m->D);
  FVX_CHECK_ARG(out != nullptr, "projection: null output");
  if (nrows <= 0) return 0;
  const int de = m->de;
  const long long gx = (nrows + PJ_TM - 1) / PJ_TM;
  FVX_CHECK_ARG(gx < 2147483647LL, "projection: too many rows");
  if (de <= 32) {
    k_project<2><<<dim3((unsigned)gx, 1), 256, 0, st>>>(m->F, rows, nrows, m->D, m->E, de, out);
  } else if (de <= 48) {
    k_project<3><<<dim3((unsigned)gx, 1), 256, 0, st>>>(m->F, rows, nrows, m->D, m->E, de, out);
  } else {
    k_project<5><<<dim3((unsigned)gx, (de + 79) / 80), 256, 0, st>>>(m->F, rows, nrows, m->D, m->E, de, out);
  }
  FVX_CHECK_LAUNCH("k_project");
  return 0;
}

int fvx_launch_grad_E(const FvxModel* m, const int32_t* rows, int64_t nrows, int* parts_out, cudaStream_t st) {
  FVX_CHECK_ARG(m->gE_part && m->ge_parts > 0 && m->W, "grad_E: scratch missing");
  long long want = (nrows + 63) / 64;
  if (want < 1) want = 1;
  if (want > m->ge_parts) want = m->ge_parts;
  long long rpp = (nrows + want - 1) / want;
  rpp = (rpp + GE_RT - 1) / GE_RT * GE_RT;
  const int parts = (int)((nrows + rpp - 1) / rpp);
  dim3 grid((m->D + GE_COLS - 1) / GE_COLS, parts, (m->de + GE_NB - 1) / GE_NB);
  k_grad_E<<<grid, 256, 0, st>>>(m->F, rows, nrows, m->D, m->W, m->de, m->gE_part, rpp);
  FVX_CHECK_LAUNCH("k_grad_E");
  *parts_out = parts;
  return 0;
}

extern "C" {

// rows == nullptr: catalog rows row0 .. row0+nrows.  TC path: blocks that fit the TH scratch.
static int project_any(const FvxModel* m, const int32_t* rows, int row0, int64_t nrows, float* out, cudaStream_t st) {
  if (!m->use_tensor_cores) {
    FVX_CHECK_ARG(m->F != nullptr, "projection: the fp32 path needs F");
    FVX_CHECK_ARG(rows != nullptr || row0 == 0, "projection: internal (row0)");
    return fvx_launch_project(m, rows, nrows, out, st);
  }
  FVX_CHECK_ARG(m->TH && m->th_cap > 0, "projection: TH scratch missing");
  if (int rc = fvx_launch_split_E(m, st)) return rc;
  const int NP = fvx_tc_np(m->de);
  const int64_t blk = m->th_cap / NP / 128 * 128;
  FVX_CHECK_ARG(blk >= 128, "projection: TH scratch too small");
  for (int64_t r0 = 0; r0 < nrows; r0 += blk) {
    const int64_t n = nrows - r0 < blk ? nrows - r0 : blk;
    int ks = fvx_tc_ksplit(m, n);
    while (ks > 1 && (int64_t)ks * n * NP > m->th_cap) ks >>= 1;
    if (int rc = fvx_launch_project_tc(m, rows ? rows + r0 : nullptr, row0 + (int)r0, n, ks, m->TH, st)) return rc;
    if (int rc = fvx_launch_reduce_partials(m->TH, n, NP, ks, m->de, out + (size_t)r0 * m->de, st)) return rc;
  }
  return 0;
}

int fvx_project(const FvxModel* model, float* theta_ext, fvx_stream_t stream) {
  FVX_CHECK_ARG(model && model->abi_version == FVX_ABI_VERSION, "fvx_project: bad model");
  FVX_CHECK_ARG(model->D > 0 && model->E && theta_ext, "fvx_project: model has no visual part");
  if (model->two_stage) {                       // GradFashion: E is composed from Ec, Ee, E2 first
    if (int rc = fvx_launch_gf_compose(model, fvx_cu(stream))) return rc;
  }
  return project_any(model, nullptr, 0, model->item_cnt, theta_ext, fvx_cu(stream));
}

int fvx_tc_width(int32_t de) { return fvx_tc_np(de); }

int fvx_project_rows(const FvxModel* model, const int32_t* rows, int64_t nrows, float* out, fvx_stream_t stream) {
  FVX_CHECK_ARG(model && model->abi_version == FVX_ABI_VERSION, "fvx_project_rows: bad model");
  FVX_CHECK_ARG(model->D > 0 && model->E && rows && out, "fvx_project_rows: bad arguments");
  FVX_CHECK_ARG(nrows >= 0 && nrows <= 2LL * model->max_batch, "fvx_project_rows: nrows outside [0, 2*max_batch]");
  if (model->two_stage) {
    if (int rc = fvx_launch_gf_compose(model, fvx_cu(stream))) return rc;
  }
  return project_any(model, rows, 0, nrows, out, fvx_cu(stream));
}

int fvx_grad_e_rows(const FvxModel* model, const int32_t* rows, int64_t nrows, const float* W, float* out,
                    fvx_stream_t stream) {
  FVX_CHECK_ARG(model && model->abi_version == FVX_ABI_VERSION, "fvx_grad_e_rows: bad model");
  FVX_CHECK_ARG(model->D > 0 && rows && W && out && model->gE_part, "fvx_grad_e_rows: bad arguments");
  FVX_CHECK_ARG(nrows >= 0 && nrows <= 2LL * model->max_batch, "fvx_grad_e_rows: nrows outside [0, 2*max_batch]");
  cudaStream_t st = fvx_cu(stream);
  int parts = 0;
  if (model->use_tensor_cores) {
    FVX_CHECK_ARG(model->W_hi && model->W_lo, "fvx_grad_e_rows: W planes missing");
    if (int rc = fvx_launch_split_W(model, W, rows, nrows, st)) return rc;
    if (int rc = fvx_launch_grad_E_tc(model, rows, nrows, &parts, st)) return rc;
    return fvx_launch_reduce_gE(model, parts, fvx_tc_np(model->de), out, st);
  }
  FVX_CHECK_ARG(model->F != nullptr && model->W != nullptr, "fvx_grad_e_rows: the fp32 path needs F and W scratch");
  if (cudaMemcpyAsync(model->W, W, sizeof(float) * nrows * model->de, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    FVX_FAIL(-3, "fvx_grad_e_rows: copy failed");
  if (int rc = fvx_launch_grad_E(model, rows, nrows, &parts, st)) return rc;
  return fvx_launch_reduce_gE(model, parts, model->de, out, st);
}

int fvx_split_planes(const float* src, uint16_t* dst, int64_t n_rows, int32_t D, fvx_stream_t stream) {
  FVX_CHECK_ARG(src && dst && D > 0, "fvx_split_planes: bad arguments");
  return fvx_launch_split_planes(src, dst, n_rows, D, fvx_cu(stream));
}

}  // extern "C"
