// Micro-benchmark: tcgen05.mma (kind::f16, bf16 in, fp32 accumulate, both operands in shared memory, 128B swizzle)
// issue patterns on B200: how long does a UMMA take when it accumulates into the columns its predecessor wrote,
// and what does interleaving independent accumulators buy?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../fashionvisualexpl-recommend_b200/csrc -o umma_chain umma_chain.cu
// One CTA per SM; thread 0 issues `chains` accumulator chains of `nk` K steps each, `rounds` times, either chain
// after chain (order 0) or K-step-interleaved across the chains (order 1); prints cycles per UMMA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fvx_tc.cuh"

__global__ void __launch_bounds__(128, 1) k_umma(long long* cycles, int N, int nk, int chains, int order, int rounds,
                                                 int fresh) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 * 128 * 2 + 256 * 128 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = tc_smem_u32(smem), b0 = a0 + 128 * 128 * 2;      // A: 2 K blocks of [128 x 128 B]; B: [256 x 128 B] x 2
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < rounds; ++r) {
      if (order == 0) {
        for (int c = 0; c < chains; ++c)
          for (int k = 0; k < nk; ++k) {
            const uint32_t off = (uint32_t)((k >> 2) & 1) * 128 * 128 + (k & 3) * 32;
            const uint32_t offb = (uint32_t)((k >> 2) & 1) * N * 128 + (k & 3) * 32;
            umma_f16(tm + c * N, umma_smem_desc(a0 + off, 16, 1024, TC_SWZ_128B), umma_smem_desc(b0 + offb, 16, 1024, TC_SWZ_128B),
                     idesc, (k || !fresh) ? 1u : 0u);
          }
      } else {
        for (int k = 0; k < nk; ++k)
          for (int c = 0; c < chains; ++c) {
            const uint32_t off = (uint32_t)((k >> 2) & 1) * 128 * 128 + (k & 3) * 32;
            const uint32_t offb = (uint32_t)((k >> 2) & 1) * N * 128 + (k & 3) * 32;
            umma_f16(tm + c * N, umma_smem_desc(a0 + off, 16, 1024, TC_SWZ_128B), umma_smem_desc(b0 + offb, 16, 1024, TC_SWZ_128B),
                     idesc, (k || !fresh) ? 1u : 0u);
          }
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  const size_t smem = 128 * 128 * 2 + 256 * 128 * 2 + 2048;
  cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Cfg { int N, nk, chains, order, fresh; };
  const Cfg cfgs[] = {
      {128, 6, 1, 0, 1}, {128, 6, 2, 0, 1}, {128, 6, 4, 0, 1}, {128, 6, 2, 1, 1}, {128, 6, 3, 1, 1}, {128, 6, 4, 1, 1},
      {128, 32, 1, 0, 1}, {128, 32, 4, 0, 1}, {128, 32, 4, 1, 1}, {256, 6, 1, 0, 1}, {256, 6, 2, 0, 1}, {256, 6, 2, 1, 1},
      {256, 32, 1, 0, 1}, {256, 32, 2, 1, 1}, {64, 6, 1, 0, 1}, {64, 6, 8, 0, 1}, {64, 6, 4, 1, 1}, {64, 6, 8, 1, 1},
      {64, 32, 8, 1, 1}, {32, 8, 8, 1, 1}, {128, 6, 4, 1, 0}};
  for (const Cfg& c : cfgs) {
    const int rounds = 200;
    k_umma<<<148, 128, smem>>>(cyc, c.N, c.nk, c.chains, c.order, rounds, c.fresh);
    long long h[148];
    cudaError_t e = cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
    const double per = m / ((double)rounds * c.chains * c.nk);
    printf("N=%3d nk=%2d chains=%d %-11s fresh=%d : %7.1f cycles/UMMA  (ideal %d)  %5.1f %% of the tensor rate\n", c.N, c.nk, c.chains,
           c.order ? "interleaved" : "chain-major", c.fresh, per, c.N / 2, 100.0 * (c.N / 2) / per);
  }
  return 0;
}
