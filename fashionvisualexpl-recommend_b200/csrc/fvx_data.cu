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

__global__ void k_perm_keys(uint32_t* __restrict__ keys, int num_users, unsigned long long seed, uint32_t epoch) {
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < num_users; u += gridDim.x * blockDim.x)
    keys[u] = fvx_philox((uint32_t)u, 0u, epoch, FVX_STREAM_PERM, (uint32_t)seed, (uint32_t)(seed >> 32)).x;
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

int fvx_perm_keys(uint32_t* keys, int32_t num_users, uint64_t seed, uint32_t epoch, fvx_stream_t stream) {
  FVX_CHECK_ARG(keys, "fvx_perm_keys: null pointer");
  if (num_users <= 0) return 0;
  k_perm_keys<<<grid_for(num_users, 256), 256, 0, fvx_cu(stream)>>>(keys, num_users, seed, epoch);
  FVX_CHECK_LAUNCH("k_perm_keys");
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
