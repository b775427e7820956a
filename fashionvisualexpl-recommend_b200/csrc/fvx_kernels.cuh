// Internal launchers shared between translation units of libfvx (not part of the C ABI).
#pragma once
#include "fvx_common.cuh"

// TH[r, :] = F[rows[r], :] * E_ext for r in [0, nrows); rows == nullptr means the
// identity (whole owned catalog); rows[r] < 0 yields a zero output row.
int fvx_launch_project(const FvxModel* m, const int32_t* rows, int64_t nrows, float* out, cudaStream_t st);

// gE_part[p] = sum over the rows of group p of F[rows[r], :]^T * W[r, :]; *parts_out groups written.
int fvx_launch_grad_E(const FvxModel* m, const int32_t* rows, int64_t nrows, int* parts_out, cudaStream_t st);

// tensor-core versions (fvx_project_tc.cu).  The forward writes [ksplit][nrows][NP] fp32 K-split
// partials (NP = fvx_tc_np(de), ksplit = fvx_tc_ksplit(...)); the backward reads the bf16 planes
// m->W_hi / m->W_lo [nrows][NP] and writes gE_part[parts][D][NP].
int fvx_tc_np(int de);
// Row pitch (bf16 elements) of the W planes: 2*NP when the two planes are interleaved row by row in
// one allocation (W_lo == W_hi + NP: row r = [hi 0..NP | lo 0..NP], what the engine allocates - the
// backward then reads [W_hi | W_lo] as ONE operand of N = 2*NP), else NP (two separate planes).
static inline int fvx_w_pitch(const FvxModel* m) {
  const int np = fvx_tc_np(m->de);
  return (m->W_lo == m->W_hi + np) ? 2 * np : np;
}
int fvx_tc_ksplit(const FvxModel* m, long long nrows);
// K split of the forward projection for `tiles` 128-row tiles of `chunks` 64-feature chunks: aim at >= 6 work
// units per SM so that the last wave is short; a split keeps >= 4 chunks.  Host and device evaluate the
// same rule (the unique-row step knows its row count only on the device).
FVX_HD int fvx_tc_ksplit_rule(long long tiles, int chunks, int nsm, int ks_cap) {
  int ks = 1;
  while (tiles * ks < 6LL * nsm && ks * 2 <= chunks / 4 && chunks % (ks * 2) == 0 && ks * 2 <= ks_cap) ks *= 2;
  return ks;
}
// The split the unique-row step derives on the device from its row count (forward kernel AND scoring kernel
// evaluate it): the plain rule, or - when that leaves the last wave of units more than 4 % short and there
// is at least one tile per CTA - stream-K with two partials per row (returns 2, *streamk = 1).
FVX_HD int fvx_tc_split_dyn(long long tiles, int chunks, int nsm, int ks_cap, int ctas, int* streamk) {
  const int ks = fvx_tc_ksplit_rule(tiles, chunks, nsm, ks_cap);
  const long long units = tiles * ks;
  const long long rounds = (units + ctas - 1) / ctas;
  *streamk = 0;
  if (ks_cap >= 2 && tiles >= ctas && rounds * ctas * 25 > units * 26) { *streamk = 1; return 2; }
  return ks;
}
int fvx_launch_split_E(const FvxModel* m, cudaStream_t st);
int fvx_launch_split_planes(const float* src, uint16_t* dst, long long n_rows, int D, cudaStream_t st);
// dyn_ks = 1: `ksplit` is only the cap; the kernel derives the split from *nrows_dev (fvx_tc_ksplit_rule) and
// lays the partials out [split][nrows][NP] with the HOST-side nrows as the row capacity.
// sm_reserve: SMs the launch leaves free (the sharded step keeps a few for the NCCL kernels that travel beside
// the tensor-core kernels: a persistent one-CTA-per-SM grid would make them wait for its last CTA)
int fvx_launch_project_tc(const FvxModel* m, const int32_t* rows, int row0, int64_t nrows, int ksplit, float* out,
                          cudaStream_t st, const int32_t* nrows_dev = nullptr, int dyn_ks = 0, int sm_reserve = 0);
int fvx_launch_reduce_partials(const float* part, long long nrows, int NP, int ks, int de, float* out,
                               cudaStream_t st);
int fvx_launch_split_W(const FvxModel* m, const float* W, const int32_t* rows, long long nrows, cudaStream_t st);
int fvx_launch_reduce_gE(const FvxModel* m, int parts, int gnp, float* out, cudaStream_t st);
int fvx_launch_grad_E_tc(const FvxModel* m, const int32_t* rows, int64_t nrows, int* parts_out, cudaStream_t st,
                         const int32_t* nrows_dev = nullptr, int sm_reserve = 0);

// Side stream of the step (one per device, FVX_STEP_OVERLAP=0 disables it): begin() makes it wait for the
// work queued on `main_stream` and returns it (nullptr: unavailable); join() makes `main_stream` wait for it.
cudaStream_t fvx_side_begin(cudaStream_t main_stream);
void fvx_side_join(cudaStream_t main_stream);

// pieces of the optimiser step shared with the item-sharded path (fvx_train_sharded.cu)
int fvx_check_model(const FvxModel* m, const char* who);
// what: ALL = one launch (rows, claims + catch-up, E planes); ROWS = only what the projection needs
// (slot rows + E planes); CLAIMS = claims + deferred-Adam catch-up without the row / plane writes
// UNIQ = slot rows + E planes + claims of the item rows with their list positions (unique-row step);
// CLAIMS_LISTED = user claims + catch-up, item catch-up driven by the list UNIQ built
// USERS_ONLY / ITEMS_LISTED: the two halves of CLAIMS_LISTED (the sharded step publishes the user rows before
// the item rows are caught up)
enum { FVX_PREP_ALL = 0, FVX_PREP_ROWS = 1, FVX_PREP_CLAIMS = 2, FVX_PREP_UNIQ = 3, FVX_PREP_CLAIMS_LISTED = 4,
       FVX_PREP_USERS_ONLY = 5, FVX_PREP_ITEMS_LISTED = 6 };
int fvx_launch_prep(const FvxModel* m, const int32_t* user, const int32_t* pos, const int32_t* neg, int B,
                    cudaStream_t st, int what = FVX_PREP_ALL);
// what: ALL = tables + E_ext + finalisation; TABLES = touched rows only; E = E_ext + finalisation
enum { FVX_UPD_ALL = 0, FVX_UPD_TABLES = 1, FVX_UPD_E = 2 };
int fvx_launch_update(const FvxModel* m, int B, int parts, int gnp, const float* gE_src, int loss_slot,
                      cudaStream_t st, int what = FVX_UPD_ALL, const float* loss_pair = nullptr, int n_tails = 1);
// dedup = 1 (unique-row step): theta rows / coefficient sums are addressed through upos, th_ks is the CAP of
// the K split (the kernel derives the split from the list length like the projection does)
int fvx_launch_score_grad(const FvxModel* m, const int32_t* user, int B, int loss_slot, int th_ks, cudaStream_t st,
                          int dedup = 0);
// switches of the step (environment: FVX_STEP_DEDUP; test hook fvx_debug_set_dedup)
bool fvx_dedup_enabled();
bool fvx_merged_update(const FvxModel* m);   // DEFERRED mode without the row-update kernel
// GradFashion: E = blockdiag(Ec, Ee) * E2 (the effective [D, de] projection matrix; ahead of anything that projects)
int fvx_launch_gf_compose(const FvxModel* m, cudaStream_t st);
// W_sum -> bf16 planes of the listed rows (unique-row step)
int fvx_launch_w_planes(const FvxModel* m, int B, cudaStream_t st);

// exact fp32 top-k for the flagged rows of [u0, u0+n), in place, no host synchronisation (fvx_eval.cu)
int fvx_launch_topk_flagged(const FvxModel* model, const float* theta_ext, const int32_t* flags, int n, int u0,
                            const int64_t* mask_row_ptr, const int32_t* mask_col, int k, int32_t* out_ids,
                            float* out_scores, int32_t* list_scratch, int32_t* count_scratch, cudaStream_t st,
                            const int32_t* tau_enc = nullptr);

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
