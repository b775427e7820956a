// Epoch enumeration, permutation keys and on-device negative sampling.
// Replaces the host loop of DataLoader.all_triple_batches (src/dataset/dataset.py:83-114).
#include "fvx_common.cuh"

// One warp per user slot p of the permutation: copies that user's train items (file
// order, dataset.py:99) to out_pos[offs[p] ...] and writes the user id beside them.
__global__ void k_enumerate_epoch(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_file,
                                  const int32_t* __restrict__ perm, const int64_t* __restrict__ offs,
                                  int num_users, int32_t* __restrict__ out_user, int32_t* __restrict__ out_pos) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < num_users; p += warps_per_grid) {
    const int32_t u = perm[p];
    const int64_t a = row_ptr[u], n = row_ptr[u + 1] - a, o = offs[p];
    for (int64_t t = lane; t < n; t += 32) {
      out_user[o + t] = u;
      out_pos[o + t] = col_file[a + t];
    }
  }
}

// Epoch permutation without a sort: a 4-round Feistel network over 2h bits (2^(2h) >= U) whose
// round function is one Philox word, restricted to [0, U) by cycle walking.  Replaces
// random.shuffle (dataset.py:95); restated in oracle/sampler.py: feistel_user_permutation.
__global__ void k_epoch_perm(int32_t* __restrict__ perm, int64_t* __restrict__ lens,
                             const int64_t* __restrict__ row_ptr, int num_users, int h, unsigned long long seed,
                             uint32_t epoch) {
  const uint32_t mask = (1u << h) - 1u;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < num_users; p += gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)p;
    do {
      uint32_t L = x >> h, R = x & mask;
#pragma unroll 1
      for (uint32_t r = 0; r < 4; ++r) {
        const uint32_t f = fvx_philox(R, r, epoch, FVX_STREAM_PERM, (uint32_t)seed, (uint32_t)(seed >> 32)).x & mask;
        const uint32_t nl = R;
        R = L ^ f;
        L = nl;
      }
      x = (L << h) | R;
    } while (x >= (uint32_t)num_users);
    perm[p] = (int32_t)x;
    if (lens) lens[p] = row_ptr[x + 1] - row_ptr[x];
  }
}

// enumerate + negatives in one pass: triple o+t of the epoch uses Philox counter offset+o+t
__global__ void k_epoch_triples(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_file,
                                const int32_t* __restrict__ col_sorted, const int32_t* __restrict__ perm,
                                const int64_t* __restrict__ offs_incl, int num_users, uint32_t num_items,
                                unsigned long long seed, unsigned long long offset, int32_t* __restrict__ out_user,
                                int32_t* __restrict__ out_pos, int32_t* __restrict__ out_neg) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < num_users; p += warps_per_grid) {
    const int32_t u = perm[p];
    const int64_t a = row_ptr[u], b = row_ptr[u + 1], n = b - a, o = offs_incl[p] - n;
    for (int64_t t = lane; t < n; t += 32) {
      out_user[o + t] = u;
      out_pos[o + t] = col_file[a + t];
      out_neg[o + t] = fvx_draw_negative(col_sorted, a, b, offset + (unsigned long long)(o + t), num_items, seed);
    }
  }
}

__global__ void k_sample_negatives(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_sorted,
                                   const int32_t* __restrict__ user, int32_t* __restrict__ neg, long long n,
                                   uint32_t num_items, unsigned long long seed, unsigned long long offset) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x) {
    const int32_t u = user[t];
    neg[t] = fvx_draw_negative(col_sorted, row_ptr[u], row_ptr[u + 1], offset + (unsigned long long)t,
                               num_items, seed);
  }
}

static inline int grid_for(long long n, int block, int per_sm = 8) {
  long long g = (n + block - 1) / block;
  long long cap = (long long)fvx_num_sms() * per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

extern "C" {

int fvx_enumerate_epoch(const int64_t* row_ptr, const int32_t* col_file, const int32_t* perm,
                        const int64_t* offs, int32_t num_users, int32_t* out_user, int32_t* out_pos,
                        fvx_stream_t stream) {
  FVX_CHECK_ARG(row_ptr && col_file && perm && offs && out_user && out_pos, "fvx_enumerate_epoch: null pointer");
  if (num_users <= 0) return 0;
  k_enumerate_epoch<<<grid_for((long long)num_users * 32, 256), 256, 0, fvx_cu(stream)>>>(
      row_ptr, col_file, perm, offs, num_users, out_user, out_pos);
  FVX_CHECK_LAUNCH("k_enumerate_epoch");
  return 0;
}

int fvx_epoch_perm(int32_t* perm, int64_t* lens, const int64_t* row_ptr, int32_t num_users, uint64_t seed,
                   uint32_t epoch, fvx_stream_t stream) {
  FVX_CHECK_ARG(perm && (lens == nullptr || row_ptr != nullptr), "fvx_epoch_perm: null pointer");
  if (num_users <= 0) return 0;
  int bits = 1;
  while (bits < 31 && (1LL << bits) < (long long)num_users) ++bits;
  const int h = (bits + 1) / 2;
  k_epoch_perm<<<grid_for(num_users, 256), 256, 0, fvx_cu(stream)>>>(perm, lens, row_ptr, num_users, h, seed, epoch);
  FVX_CHECK_LAUNCH("k_epoch_perm");
  return 0;
}

int fvx_epoch_triples(const int64_t* row_ptr, const int32_t* col_file, const int32_t* col_sorted,
                      const int32_t* perm, const int64_t* offs_incl, int32_t num_users, int32_t num_items,
                      uint64_t seed, uint64_t offset, int32_t* out_user, int32_t* out_pos, int32_t* out_neg,
                      fvx_stream_t stream) {
  FVX_CHECK_ARG(row_ptr && col_file && col_sorted && perm && offs_incl && out_user && out_pos && out_neg,
                "fvx_epoch_triples: null pointer");
  FVX_CHECK_ARG(num_items > 0, "fvx_epoch_triples: num_items must be positive");
  if (num_users <= 0) return 0;
  k_epoch_triples<<<grid_for((long long)num_users * 32, 256), 256, 0, fvx_cu(stream)>>>(
      row_ptr, col_file, col_sorted, perm, offs_incl, num_users, (uint32_t)num_items, seed, offset, out_user,
      out_pos, out_neg);
  FVX_CHECK_LAUNCH("k_epoch_triples");
  return 0;
}

int fvx_sample_negatives(const int64_t* row_ptr, const int32_t* col_sorted, const int32_t* user, int32_t* neg,
                         int64_t n, int32_t num_items, uint64_t seed, uint64_t offset, fvx_stream_t stream) {
  FVX_CHECK_ARG(row_ptr && col_sorted && user && neg, "fvx_sample_negatives: null pointer");
  FVX_CHECK_ARG(num_items > 0, "fvx_sample_negatives: num_items must be positive");
  if (n <= 0) return 0;
  k_sample_negatives<<<grid_for(n, 256), 256, 0, fvx_cu(stream)>>>(row_ptr, col_sorted, user, neg, n,
                                                                  (uint32_t)num_items, seed, offset);
  FVX_CHECK_LAUNCH("k_sample_negatives");
  return 0;
}

}  // extern "C"
