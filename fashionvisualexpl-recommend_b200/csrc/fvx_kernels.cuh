// Internal launchers shared between translation units of libfvx (not part of the C ABI).
#pragma once
#include "fvx_common.cuh"

// TH[r, :] = F[rows[r], :] * E_ext for r in [0, nrows); rows == nullptr means the
// identity (whole owned catalog); rows[r] < 0 yields a zero output row.
int fvx_launch_project(const FvxModel* m, const int32_t* rows, int64_t nrows, float* out, cudaStream_t st);

// gE_part[p] = sum over the rows of group p of F[rows[r], :]^T * W[r, :]; *parts_out groups written.
int fvx_launch_grad_E(const FvxModel* m, const int32_t* rows, int64_t nrows, int* parts_out, cudaStream_t st);

int fvx_launch_score_grad(const FvxModel* m, const int32_t* user, int B, int loss_slot, cudaStream_t st);

// x_ui for one (user row, item row, theta row): the single definition every scoring
// kernel uses, so that a score compared with itself compares equal
// (BPRMF.py:74 / VBPR.py:82-84; accumulation order: K latent terms, d visual terms,
// item bias, visual bias).
template <typename UPtr, typename IPtr, typename TPtr>
__device__ __forceinline__ float fvx_score_one(UPtr urow, IPtr irow, TPtr th, int K, int d) {
  float s = 0.0f;
  for (int c = 0; c < K; ++c) s = fmaf(urow[c], irow[c], s);
  if (d > 0) {
    for (int n = 0; n < d; ++n) s = fmaf(urow[K + n], th[n], s);
    s += irow[K];
    s += th[d];
  } else {
    s += irow[K];
  }
  return s;
}
