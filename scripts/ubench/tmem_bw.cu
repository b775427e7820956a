// Micro-benchmark: tcgen05.ld throughput per SM (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
// One CTA per SM allocates all 512 TMEM columns; W warps (a multiple of 4: warp w reads lanes 32*(w%4)..)
// each loop over the columns with tcgen05.ld.32x32b.xN, `depth` loads in flight before a wait.
// Prints bytes per clock per SM for every (warps, shape, depth).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N> __device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v);
template <> __device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void ld<64>(uint32_t taddr, uint32_t* v) {
  ld<32>(taddr, v); ld<32>(taddr + 32, v + 32);     // two x32 back to back (no wait between)
}

template <int N, int DEPTH>
__global__ void __launch_bounds__(512, 1) k_tmem(long long* cycles, unsigned* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t v[DEPTH][N];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c0 = 0; c0 < 512; c0 += N * DEPTH) {
#pragma unroll
      for (int q = 0; q < DEPTH; ++q) ld<N>(base + c0 + q * N, v[q]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < DEPTH; ++q)
#pragma unroll
        for (int j = 0; j < N; j += 8) acc ^= v[q][j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int N, int DEPTH> void run(int warps, const char* name) {
  long long* cyc; unsigned* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  k_tmem<N, DEPTH><<<148, warps * 32>>>(cyc, sink, iters);
  k_tmem<N, DEPTH><<<148, warps * 32>>>(cyc, sink, iters);
  long long h[148];
  cudaError_t e = cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
  const double bytes = (double)warps * iters * 512.0 * 32 * 4;     // every warp reads 32 lanes x 512 columns per iteration
  printf("warps=%2d shape=%-4s depth=%d : %.1f B/clk/SM  (%.0f cycles per 4 KB warp-load round)\n", warps, name, DEPTH,
         bytes / mean, mean / (iters * 512.0 / (N * DEPTH)));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<16, 1>(w, "x16"); run<32, 1>(w, "x32"); run<32, 2>(w, "x32"); run<64, 1>(w, "2x32"); run<16, 4>(w, "x16");
  }
  return 0;
}
