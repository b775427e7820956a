/* fvx.h - C ABI of libfvx.so: the B200 (sm_100a) BPR train step and full-catalog
 * top-k evaluation of BPRMF / VBPR.
 *
 * The reference (peternara/FashionVisualExpl-recommend) has no FFI layer: its hot
 * path is a Python class protocol over TensorFlow ops.  Each entry point below
 * replaces the TensorFlow/NumPy work of the reference call site cited next to it;
 * the Python mirror classes in fashionvisualexpl-recommend_b200/ bind them with
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types;
 *   - every pointer in FvxModel / FvxTable / arguments is a DEVICE pointer unless
 *     its name ends in _host; memory is owned by the caller (torch tensors) and
 *     only borrowed for the duration of the stream-ordered call;
 *   - every call is asynchronous on `stream` (a cudaStream_t) and performs no
 *     host synchronisation, so it can be captured into a CUDA graph;
 *   - return value 0 = success, negative = error; fvx_last_error() gives the text
 *     (thread-local).  There is no CPU fallback.
 *   - the optimiser step counter lives on the device (FvxModel.step) so that a
 *     captured graph of steps can be replayed.
 */
#ifndef FVX_H_
#define FVX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fvx_stream_t; /* cudaStream_t */

#define FVX_ABI_VERSION 5

/* Adam semantics.  The reference's Keras Adam moves
 * EVERY row of an embedding table on every step (rows without gradient keep
 * moving on momentum).  DENSE reproduces that literally with a whole-table sweep;
 * DEFERRED gives the same result (to fp32 rounding) lazily: a row is brought up to
 * date only when it is next needed (the next step that touches it, or fvx_adam_flush).
 * Up to date means: first the Adam step of the batch that last touched it - that
 * batch's gradient stays in `g` until then, the step has no separate row-update
 * kernel - then the zero-gradient steps it skipped since.  `w` of a touched row is
 * therefore STALE between steps: call fvx_adam_flush before reading the tables.
 * LAZY skips untouched rows (NOT reference semantics - a labelled fast mode). */
enum { FVX_ADAM_DENSE = 0, FVX_ADAM_DEFERRED = 1, FVX_ADAM_LAZY = 2 };

/* One embedding table with its optimiser state, row-major, row stride `stride`
 * floats (a multiple of 4 so rows are 16-byte aligned).
 *   user table: cols [0,K) = Gu (BPRMF.py:49), cols [K,K+d) = Tu (VBPR.py:46-48)
 *   item table: cols [0,K) = Gi (BPRMF.py:50), col K = Bi (BPRMF.py:48)          */
typedef struct FvxTable {
  float* w;        /* [rows, stride] parameters                                   */
  float* m;        /* [rows, stride] Adam first moment                            */
  float* v;        /* [rows, stride] Adam second moment                           */
  float* g;        /* [rows, stride] gradient accumulator; DENSE / LAZY: all-zero between
                      steps; DEFERRED: the pending gradient of step last+1 (zero after a flush) */
  int32_t* last;   /* [rows] step up to which the row is current (DEFERRED)       */
  int32_t* mark;   /* [rows] step at which the row was last marked as touched     */
  int32_t* list;   /* [list_cap] rows touched by the step in flight               */
  int32_t* count;  /* [1] number of valid entries in list                         */
  int64_t rows;
  int32_t stride;
  int32_t list_cap;
} FvxTable;

typedef struct FvxModel {
  int32_t abi_version;  /* FVX_ABI_VERSION */
  int32_t num_users;    /* U */
  int32_t num_items;    /* I: size of the whole catalog                            */
  int32_t item_lo;      /* first catalog row owned by this rank (0 on one GPU)      */
  int32_t item_cnt;     /* catalog rows owned by this rank (I on one GPU)           */
  int32_t K;            /* embed_k                                                  */
  int32_t d;            /* embed_d (0 for BPRMF)                                    */
  int32_t D;            /* CNN feature dimension (0 for BPRMF)                      */
  int32_t de;           /* row stride of E_ext: round_up4(d+1)                      */
  int32_t adam_mode;    /* FVX_ADAM_*                                               */
  float lr, reg;        /* Adam(lr) BPRMF.py:52; reg BPRMF.py:38                    */
  FvxTable users;       /* [U, round_up4(K+d)]                                      */
  FvxTable items;       /* [item_cnt, round_up4(K+1)], local row = item - item_lo   */
  /* dense visual parameters, VBPR.py:44-54: E_ext[D,de], cols [0,d) = E, col d = Bp */
  float *E, *mE, *vE;
  float* gE_part;       /* [ge_parts, D, max(de, NP)] per-row-group partial gradients */
  int32_t ge_parts;
  int32_t _pad0;
  const float* F;       /* [item_cnt, D] fp32 features, already /max|F|
                           (visual_loader_mixin.py:30); may be NULL when
                           use_tensor_cores = 1 (only the planes are read then)      */
  const uint16_t* F_pl; /* [item_cnt, D/64, 2, 64] bf16 planes of F (tensor-core path):
                           per 64-feature chunk 64 x hi = bf16(F) then 64 x lo =
                           bf16(F - hi); 4*D bytes per row (fvx_split_planes)       */
  uint16_t* ET_hi;      /* [NP, D] bf16 planes of E_ext^T (NP = fvx_tc_width(de)),  */
  uint16_t* ET_lo;      /*   scratch refreshed by every call that projects          */
  uint16_t* W_hi;       /* [2*max_batch, NP] bf16 planes of the backward coefficients (tensor-core path), row pitch NP; or */
  uint16_t* W_lo;       /* interleaved: W_lo == W_hi + NP, row pitch 2*NP ([hi NP | lo NP] per row) - the faster layout */
  int64_t* step;        /* [1] number of optimiser steps applied so far             */
  double* loss;         /* [loss_slots] per-step loss accumulators                  */
  int32_t loss_slots;
  int32_t _pad1;
  /* per-step scratch, sized for max_batch triples */
  float* TH;            /* F[i]*E_ext for every (triple, side) slot: [2*max_batch, de]
                           (fp32 path) or [ksplit, 2*max_batch, NP] K-split partials
                           (tensor-core path)                                       */
  int64_t th_cap;       /* floats available at TH                                   */
  float* W;             /* [2*max_batch, de] backward coefficients (fp32 path)      */
  int32_t* rows;        /* [2*max_batch] local item row of each (triple, side) slot,
                           -1 when the item belongs to another rank                 */
  int32_t* sync;        /* [4] zero-initialised inter-block counters of the step kernels */
  int32_t* cmap;        /* [6*max_batch] sharded step only: compact list of the slots this rank
                           owns - local item row [2B] | slot [2B] | position of a slot [2B] */
  int32_t max_batch;
  int32_t use_tensor_cores; /* 0: fp32 SIMT projection; 1: tcgen05 (needs the planes) */
  /* unique-row step (tensor-core path, one rank; all NULL: every slot is projected on its own).
   * A batch of B triples touches far fewer than 2B distinct catalog rows (popular positives repeat,
   * and 2B draws from I items collide): the step projects each distinct row ONCE - the touched-row
   * list items.list doubles as the row list of the projection and of grad_E - and the backward
   * coefficients of all slots of a row are summed before the contraction.                      */
  int32_t* upos;        /* [item_cnt] position of the row in items.list (valid while
                           items.mark[row] == step + 1)                                      */
  float* W_sum;         /* [2*max_batch, NP] fp32 sums of the backward coefficients per listed
                           row; all-zero between steps                                        */
  int32_t* uslot;       /* [2*max_batch] list position of the row of each (triple, side) slot  */
  int32_t user_lo;      /* first user row OWNED by this rank (0 on one GPU)                                   */
  int32_t user_cnt;     /* user rows owned by this rank (num_users on one GPU).  The user tables are allocated
                           and indexed globally on every rank, but only the owner keeps a user's Adam state
                           (m, v, g, last) and its authoritative row; other ranks' copies of w are stale until
                           the rows are gathered (evaluation, checkpoint)                                      */
  /* GradFashion (two_stage != 0; src/recommender/models/GradFashion.py:97-134): the item's visual feature is a
   * learned projection of two descriptors, v_i = [Fc[i] Ec | Fe[i] Ee], and theta_i = v_i E, vbias_i = v_i Bp.
   * With F = [Fc | Fe] (D = Dc + De) that is theta_ext = F * (blockdiag(Ec, Ee) * E2): every step composes the
   * effective [D, de] matrix into `E` (scratch here), runs the VBPR kernels on it, and carries the gradient of the
   * effective matrix back to Ec, Ee and E2 = [E | Bp] ([ec + ee, de]), which are the trained tensors. */
  int32_t two_stage;
  int32_t Dc, De, ec, ee;     /* feature dims of the colour / edge descriptors, embed_color, embed_edges */
  float bias_neg_scale;       /* weight of the NEGATIVE item's bias in the L2 term: 0.1 (BPRMF.py:110, VBPR.py:125:
                                 beta_neg / 10) or 1.0 (GradFashion.py:171-172)                              */
  float *Ec, *mEc, *vEc;      /* [Dc, ec] + Adam moments */
  float *Ee, *mEe, *vEe;      /* [De, ee] */
  float *E2, *mE2, *vE2;      /* [ec + ee, de]: cols [0, d) = E, col d = Bp */
  float* gf_scratch;          /* [D*de + Dc*ec + De*ee + (ec+ee)*de] gradients of one step */
  int32_t* batch_stage; /* [3*max_batch + 4] staging area of fvx_bpr_steps: the (user | pos | neg) indices of
                           the batch in flight and the 64-bit batch cursor (may be NULL otherwise)     */
} FvxModel;

/* ---- library ------------------------------------------------------------------ */
int fvx_abi_version(void);
const char* fvx_last_error(void);
/* size of the structs as compiled, so a binding can assert its own layout */
int fvx_sizeof_model(void);
int fvx_sizeof_table(void);
int fvx_sizeof_shard_ws(void);
int fvx_sizeof_eval_ws(void);

/* ---- data: replaces DataLoader.all_triple_batches (dataset.py:83-114) ---------- */

/* (user, pos) pairs of one epoch in the reference's enumeration order: users in
 * `perm` order, each user's positives in file order (dataset.py:96-99).
 * offs[p] = exclusive prefix sum of the train-list lengths in perm order. */
int fvx_enumerate_epoch(const int64_t* row_ptr, const int32_t* col_file, const int32_t* perm,
                        const int64_t* offs, int32_t num_users, int32_t* out_user,
                        int32_t* out_pos, fvx_stream_t stream);

/* User order of one epoch, replacing random.shuffle (dataset.py:95): perm[p] = user at
 * position p, computed per position by a 4-round Feistel network keyed by Philox(seed,
 * epoch) with cycle walking (no sort, no host synchronisation).  lens (optional) receives
 * the train-list length of perm[p]; its inclusive prefix sum is offs_incl below. */
int fvx_epoch_perm(int32_t* perm, int64_t* lens, const int64_t* row_ptr, int32_t num_users,
                   uint64_t seed, uint32_t epoch, fvx_stream_t stream);

/* fvx_enumerate_epoch + fvx_sample_negatives in one pass (offs_incl = INCLUSIVE prefix sum of
 * the train-list lengths in perm order): triple n of the epoch uses counter offset+n. */
int fvx_epoch_triples(const int64_t* row_ptr, const int32_t* col_file, const int32_t* col_sorted,
                      const int32_t* perm, const int64_t* offs_incl, int32_t num_users,
                      int32_t num_items, uint64_t seed, uint64_t offset, int32_t* out_user,
                      int32_t* out_pos, int32_t* out_neg, fvx_stream_t stream);

/* Uniform negatives with rejection against the user's TRAIN items
 * (dataset.py:100-103); counter-based: triple n uses counter offset+n. */
int fvx_sample_negatives(const int64_t* row_ptr, const int32_t* col_sorted, const int32_t* user,
                         int32_t* neg, int64_t n, int32_t num_items, uint64_t seed,
                         uint64_t offset, fvx_stream_t stream);

/* ---- training: replaces train_step (BPRMF.py:87-125, VBPR.py:99-144) ----------- */

/* One optimiser step on a batch of B triples (global item ids).  Adds the batch
 * loss (BPRMF.py:104-115 / VBPR.py:117-130) into model->loss[loss_slot]. */
int fvx_bpr_step(const FvxModel* model, const int32_t* user, const int32_t* pos,
                 const int32_t* neg, int32_t B, int32_t loss_slot, fvx_stream_t stream);

/* n_steps consecutive optimiser steps on batches that lie back to back in epoch-long index arrays (what
 * DataLoader.next_triple_batch yields as consecutive slices, dataset.py:116-122): step s trains on triples
 * [(first + s) * B, (first + s + 1) * B).  The launch sequence of 8 steps is captured ONCE into a CUDA graph
 * (per device, keyed by the model struct, the array pointers, B and loss_slot) and replayed: at the
 * reference's default batch of 256 (train_rec.py:23) a step is a handful of microsecond kernels and the
 * host's launch cost would bound it.  Needs FvxModel.batch_stage.  Losses add up in loss[loss_slot]. */
int fvx_bpr_steps(const FvxModel* model, const int32_t* user, const int32_t* pos, const int32_t* neg,
                  int64_t first, int32_t n_steps, int32_t B, int32_t loss_slot, fvx_stream_t stream);

/* Profiling variant of fvx_bpr_step: same work, CUDA events between the phases, and a
 * host synchronisation at the end (so it cannot be graph-captured).  phase_ms_host
 * receives FVX_N_PHASES durations in launch order: prep (touched-row lists + deferred-Adam
 * catch-up + E planes), projection, score+grad, grad_E, update (Adam rows + Adam E + step). */
#define FVX_N_PHASES 5
int fvx_bpr_step_timed(const FvxModel* model, const int32_t* user, const int32_t* pos,
                       const int32_t* neg, int32_t B, int32_t loss_slot, float* phase_ms_host,
                       fvx_stream_t stream);

/* ---- the same step with the model sharded over R ranks (one process per GPU) --------------------------
 * Items (Gi, Bi, F and their Adam state) are row-sharded in contiguous blocks (item_lo, item_cnt); USERS are
 * owned in contiguous blocks too (user_lo, user_cnt): the owner keeps a user's Adam state and applies its
 * updates; E is replicated.  Every rank sees the same batch (global ids).  x_uij = s_ui - s_uj is linear in
 * the item-side terms, so a rank scores the (triple, side) slots whose item it owns, and the ranks exchange
 * four buffers per step:
 *   WU [owners * run_cap, users.stride]  the up-to-date rows of the batch's users, one row per RUN of equal users
 *                                (the reference's sampler emits runs of one user, dataset.py:96-99), grouped in one
 *                                segment of run_cap rows per OWNER of the user: all-gather (each owner publishes
 *                                its segment)
 *   S  [2B]                      partial scores of the slots [pos(B) | neg(B)], zero where a rank owns nothing:
 *                                all-reduce (sum)
 *   RU [owners * run_cap, users.stride]  user-row gradient shares in the same layout: reduce-scatter (each owner
 *                                receives the sum of its segment and adds its runs into its accumulators)
 *   dE [D*de + 4]                dense gradient of E_ext, then the loss share (two floats, hi + lo) and the
 *                                run-overflow flag: all-reduce (sum)
 * Work per rank: 2B/R slots, B/R user rows - nothing grows with the number of ranks except the scans of the
 * batch indices.  No item row ever leaves its owner.  run_id[b] = row of the run triple b belongs to (fvx_run_slots:
 * owner * run_cap + index of the run among that owner's runs, identical on every rank); a batch in which one owner
 * has more than run_cap runs poisons the step's loss with NaN on every rank (nothing fails silently). */
typedef struct FvxShardWs {
  float* S;
  int32_t* run_id;       /* [max_batch] */
  int32_t* run_scratch;  /* [8 * (max_batch / 1024 + 2)] */
  float* WU;
  float* RU;
  float* dE;
  double* loss_part;     /* [1] this rank's loss share of the step in flight */
  int32_t max_runs;      /* rows of WU / RU = owners * run_cap */
  int32_t run_cap;       /* rows per owner segment */
  int32_t owners;        /* number of ranks (<= 8) */
  int32_t users_per_owner;   /* block size of the user ownership: owner(u) = u / users_per_owner */
  /* peer-to-peer exchange (p2p != 0; fvx_comm_arena): WU, S and the four buffers below lie inside the communicator's
   * peer-mapped arena at the same offsets on every rank; the kernels that produce the data store it straight into
   * their consumers' copies over NVLink and one-warp barrier kernels replace the collectives */
  int32_t p2p;
  int32_t _pad;
  float* RUin;           /* [owners, run_cap, users.stride] gradient shares of THIS rank's runs, one block per source */
  float* dEall;          /* [owners, D*de] one dE per rank */
  float* tails;          /* [owners, 4] loss share (hi, lo) and run-overflow flag of every rank */
  uint32_t* flags;       /* [4, 8] barrier flags, zero-initialised */
  int32_t* run_user;     /* [run_cap] (local) user of each run of this rank's segment */
  int32_t* run_counts;   /* [8] (local) runs per owner in the batch in flight */
} FvxShardWs;

/* Communicators of the sharded step: two NCCL communicators over the same ranks, one for the collectives on
 * the caller's stream (S, dE) and one for those that travel on a side stream beside the tensor-core kernels
 * (WU beside the projection, RU beside grad_E).  libnccl.so.2 is resolved at run time (the copy the process
 * already loaded, e.g. torch's, else the system one); libfvx does not link against it.
 *   rank 0: fvx_comm_unique_id(id)  ->  broadcast the 256 bytes (torch.distributed, MPI, a file ...)
 *   every rank: fvx_comm_create(id, rank, world, &comm)   [collective]                                   */
typedef struct FvxComm FvxComm;
#define FVX_COMM_ID_BYTES 256
int fvx_comm_unique_id(uint8_t* id_host);
int fvx_comm_create(const uint8_t* id_host, int32_t rank, int32_t world, FvxComm** out);
int fvx_comm_destroy(FvxComm* comm);
/* Peer-mapped arena [collective]: every rank allocates `bytes` (zero-filled) and maps every other rank's allocation
 * (CUDA IPC; NVLink peer access) - kernels of the sharded step then store into their peers' exchange buffers
 * directly.  *local receives this rank's base address; lay the buffers out at the same offsets on every rank.
 * A second call replaces the arena (the buffers of the first become invalid on every rank). */
int fvx_comm_arena(FvxComm* comm, int64_t bytes, void** local);
/* all-reduce (sum) of n floats in place on the caller's stream, e.g. to assemble per-rank statistics */
int fvx_comm_all_reduce_f32(FvxComm* comm, float* buf, int64_t n, fvx_stream_t stream);

/* run_id[b] = index of the run of equal users triple b belongs to, counted from 0 in batch order; two small
 * launches; scratch: >= n / 4096 + 2 int32. */
int fvx_run_ids(const int32_t* user, int64_t n, int32_t* run_id, int32_t* scratch, fvx_stream_t stream);
/* run_slot[b] = owner * cap + (index of b's run among the runs of that owner), owner = user / users_per_owner
 * (clamped to owners - 1); 0x7fffffff when the index reaches cap.  owners <= 8; scratch: >= 8 * (n / 1024 + 2) int32. */
int fvx_run_slots(const int32_t* user, int64_t n, int32_t users_per_owner, int32_t owners, int32_t cap,
                  int32_t* run_slot, int32_t* scratch, fvx_stream_t stream);

/* One optimiser step of the sharded model: ONE call per rank per step; the four exchanges (WU: fresh user rows,
 * S: partial scores, RU: user-gradient shares, dE) are issued inside, WU and RU on a side stream.  ws->p2p = 0:
 * NCCL collectives (all-gather, all-reduce, reduce-scatter, all-reduce); ws->p2p = 1: the producer kernels store
 * into their peers' buffers of the arena (fvx_comm_arena: ws->WU, S, RUin, dEall, tails, flags point into it) and
 * one-warp barrier kernels order the steps.  The batch loss lands in model->loss[loss_slot] on EVERY rank; a batch
 * with more runs of equal users per owner than ws->run_cap makes it NaN on every rank. */
int fvx_bpr_step_sharded(const FvxModel* model, const FvxShardWs* ws, FvxComm* comm, const int32_t* user,
                         const int32_t* pos, const int32_t* neg, int32_t B, int32_t loss_slot, fvx_stream_t stream);

/* The same step cut at its collectives, everything on `stream`, for callers that perform the sums themselves
 * (tests that emulate R ranks inside one process; other transports):
 *   phase 0  run slots, rows of the owned slots, catch-up of the owned users, their fresh rows -> own segment of WU
 *            -- every rank receives every owner's segment of WU --
 *   phase 1  projection of the distinct owned rows, partial scores -> S
 *            -- sum S --
 *   phase 2  gradients of the owned slots (item rows, RU, backward coefficients), dE, loss share -> dE
 *            -- sum RU (at least each owner's segment), sum dE --
 *   phase 3  RU rows of the owned users -> their accumulators; Adam on E; loss; step += 1               */
int fvx_bpr_step_sharded_phase(const FvxModel* model, const FvxShardWs* ws, const int32_t* user, const int32_t* pos,
                               const int32_t* neg, int32_t B, int32_t loss_slot, int32_t phase, fvx_stream_t stream);

/* DEFERRED mode: bring every row of both tables up to the current step - pending
 * gradient step, then the skipped zero-gradient steps (call before reading
 * parameters: evaluation, checkpoint, predict_all).  Idempotent. */
int fvx_adam_flush(const FvxModel* model, fvx_stream_t stream);

/* ---- evaluation: replaces predict_all + the evaluator's host loops ------------- */

/* theta_ext[rows, de] = F[rows] * E_ext  (VBPR.py:95-97: matmul(F, E), matmul(F, Bp)) */
int fvx_project(const FvxModel* model, float* theta_ext, fvx_stream_t stream);

/* Dense scores for users [u0,u1) over the owned catalog rows, row-major
 * [u1-u0, item_cnt] (BPRMF.py:85 / VBPR.py:95-97).  Small sizes only. */
int fvx_predict_all(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                    float* out, fvx_stream_t stream);

/* Masked top-k for users [u0,u1) over the owned catalog rows without materialising
 * the score matrix: items in the mask CSR (the user's train items, Evaluator.py:235;
 * global ids, ascending per user) are excluded; ties go to the smaller item id.
 * out_ids [u1-u0, k] are GLOBAL item ids (-1 / -inf when fewer than k candidates).
 * Rank counts (Evaluator.py:96-98): with n_thr > 0, thr_scores[(u-u0)*n_thr + t]
 * (NaN = unused) are scores of held-out items and out_counts[(u-u0)*n_thr + t]
 * receives the number of owned, non-masked items whose score is >= that value
 * (the held-out item itself included when owned).  k <= 128, n_thr <= 4. */
int fvx_score_topk(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                   const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k,
                   int32_t* out_ids, float* out_scores, int32_t n_thr, const float* thr_scores,
                   int32_t* out_counts, fvx_stream_t stream);

/* Rank counts alone (Evaluator.py:96-98) - all Evaluator.eval needs: AUC, HR@k, nDCG@k, precision and
 * recall follow from the position of each held-out item.  out_counts[(u-u0)*n_thr + t] = number of owned,
 * non-masked items whose score is >= thr_scores[(u-u0)*n_thr + t] (NaN = unused), the same value
 * fvx_score_topk returns (same fmaf chain, bit for bit), from a register-tiled sweep that keeps no
 * candidate lists: 128 users x 128 items per CTA.  1 <= n_thr <= 4. */
int fvx_rank_counts(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                    const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t n_thr,
                    const float* thr_scores, int32_t* out_counts, fvx_stream_t stream);

/* fvx_score_topk for an explicit list of n users (each in [0, num_users)): row j of the outputs
 * belongs to users[j].  The exact path for the rows fvx_score_topk_tc flags. */
int fvx_score_topk_users(const FvxModel* model, const float* theta_ext, const int32_t* users, int32_t n,
                         const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k,
                         int32_t* out_ids, float* out_scores, fvx_stream_t stream);

/* Tensor-core variant of fvx_score_topk (tcgen05 + TMA; top-k only, no rank counts): a bf16 sweep over
 * the catalog finds, per user, a lower bound of the (k + #train)-th best score from the maxima of column
 * groups, a second sweep collects every item whose bf16 score (plus its rounding bound) reaches it, and the
 * candidates are re-scored in fp32 with the same arithmetic as fvx_score_topk, so both return identical ids
 * and scores.  The caller owns the workspace: fill KP / splits / cap / n_ut / lists / gmax_elems with
 * fvx_eval_ws_query() and allocate A [u_cap*KP] bf16, Bm [i_cap*KP] bf16, epsa [u_cap] f32, nb [i_cap] f32,
 * stat [2] f32, cand [lists*cap] u64, ccount [lists] i32, flags [u_cap] i32, thr [u_cap] u32,
 * gmax [gmax_elems] f32, nbc [i_cap/32 + 1] f32.  a_stride (0 / 1: every tile, 2: every other tile) thins the bounds sweep.
 * Rows whose candidate list overflows are recomputed by the exact fp32 kernel inside the same call;
 * flags[u-u0] != 0 tells which (diagnostics only).  Needs K+d+3 <= 448. */
typedef struct FvxEvalWs {
  uint16_t* A;
  uint16_t* Bm;
  float* epsa;
  float* nb;
  float* stat;
  uint64_t* cand;
  int32_t* ccount;
  int32_t* flags;
  uint32_t* thr;
  float* gmax;
  float* nbc;
  int64_t lists;
  int64_t gmax_elems;
  int32_t u_cap, i_cap, KP, splits, cap, n_ut, a_stride, _pad;
} FvxEvalWs;
int fvx_eval_ws_query(const FvxModel* model, int32_t n_users, FvxEvalWs* ws);
int fvx_score_topk_tc(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                      const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k,
                      int32_t* out_ids, float* out_scores, const FvxEvalWs* ws, fvx_stream_t stream);
/* The two halves of fvx_score_topk_tc, for an item-sharded catalog: after _bounds, ws->thr [n_users] holds every
 * row's bound as an int32 whose signed order is the float's; a shard's bound is a valid lower bound of the
 * (k + #train)-th best score of the WHOLE catalog, so the ranks take the element-wise MAXIMUM of their ws->thr
 * (one all-reduce of 4 bytes per user) and _select then keeps, per shard, only what can reach the global top:
 * candidate lists, re-scoring and the top-k exchange shrink with the number of shards. */
int fvx_score_topk_tc_bounds(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                             const int64_t* mask_row_ptr, int32_t k, const FvxEvalWs* ws, fvx_stream_t stream);
int fvx_score_topk_tc_select(const FvxModel* model, const float* theta_ext, int32_t u0, int32_t u1,
                             const int64_t* mask_row_ptr, const int32_t* mask_col, int32_t k, int32_t* out_ids,
                             float* out_scores, const FvxEvalWs* ws, fvx_stream_t stream);

/* Scores of explicit (user, item) pairs with owned items (0 for others):
 * BPRMF.call / VBPR.call x_ui (BPRMF.py:69-74, VBPR.py:73-84). */
int fvx_score_pairs(const FvxModel* model, const float* theta_ext, const int32_t* user,
                    const int32_t* item, int64_t n, float* out, fvx_stream_t stream);

/* Merge R per-shard top-k lists per user ([n_users, R, k] ids/scores, each list
 * sorted descending) into one [n_users, k] list; ties -> smaller item id. */
int fvx_topk_merge(const int32_t* ids, const float* scores, int64_t n_users, int32_t R, int32_t k,
                   int32_t* out_ids, float* out_scores, fvx_stream_t stream);

/* ---- the projection as a building block (VBPR.py:83-84 and its gradient) ------- */
/* Padded width of the tensor-core operands for a given de = round_up4(d+1). */
int fvx_tc_width(int32_t de);
/* out[r, 0:de] = F[rows[r], :] * E_ext for r in [0, nrows); rows[r] < 0 gives an
 * unspecified row.  nrows <= 2*max_batch.  Uses tcgen05 when use_tensor_cores = 1. */
int fvx_project_rows(const FvxModel* model, const int32_t* rows, int64_t nrows, float* out,
                     fvx_stream_t stream);
/* out[D, de] = sum_r F[rows[r], :]^T * W[r, :], W fp32 [nrows, de] (rows[r] < 0: skipped).
 * nrows <= 2*max_batch.  Uses tcgen05 when use_tensor_cores = 1. */
int fvx_grad_e_rows(const FvxModel* model, const int32_t* rows, int64_t nrows, const float* W,
                    float* out, fvx_stream_t stream);

/* ---- feature planes for the tensor-core path ----------------------------------- */
/* hi = bf16(F), lo = bf16(F - float(hi)) : 4 bytes/element like fp32, ~2^-17 relative.
 * src fp32 [n_rows, D] -> dst [n_rows, D/64, 2, 64] (the FvxModel.F_pl layout); D % 64 == 0. */
int fvx_split_planes(const float* src, uint16_t* dst, int64_t n_rows, int32_t D, fvx_stream_t stream);

/* ---- diagnostics (tests and profiling scripts; no product code calls them) ------------------------ */
/* Unique-row step on (1) / off (0: one projection per slot) for this process; returns the previous value
 * (-1: the FVX_STEP_DEDUP environment default was still in force). */
int fvx_debug_set_dedup(int on);
/* Two-stream timeline of fvx_bpr_step: with tracing on the step records a timing event after every kernel;
 * fvx_debug_trace_read copies 11 times in microseconds (begin, uniq, fwd, prep0, prep1, score, w_planes,
 * grad_E, upd0, upd1, end) relative to the first one and synchronises. */
int fvx_debug_trace(int on);
int fvx_debug_trace_read(float* us_host);
/* The same for fvx_bpr_step_sharded: 13 times (begin, piece 1, piece 2 [side], all-reduce WU [side], piece 3, piece 4,
 * all-reduce S, piece 5, all-reduce RU [side], piece 7 [side], piece 6, all-reduce dE, end). */
int fvx_debug_trace_sharded(int on);
int fvx_debug_trace_sharded_read(float* us_host);

#ifdef __cplusplus
}
#endif
#endif /* FVX_H_ */
