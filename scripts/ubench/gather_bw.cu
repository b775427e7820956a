// Micro-benchmark: how fast can one B200 stream RANDOMLY GATHERED 4 KB rows (one bf16 plane of
// a 2048-d feature row) from HBM into shared memory, by mechanism?  Informs the producer design
// of fvx_project_tc.cu.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../../fashionvisualexpl-recommend_b200/csrc/fvx_tc.cuh"

#define STAGES 8
#define STAGE_BYTES 16384

static tc_encode_tiled_fn get_enc() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return (tc_encode_tiled_fn)p;
}
static CUtensorMap make_map(const void* base, uint64_t rows, uint64_t cols, uint32_t bc, uint32_t br, int sw, int promo) {
  CUtensorMap m; cuuint64_t dims[2] = {cols, rows}; cuuint64_t str[1] = {cols * 2}; cuuint32_t box[2] = {bc, br};
  cuuint32_t es[2] = {1, 1};
  CUresult r = get_enc()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
      promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

__device__ __forceinline__ void cpasync16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpasync_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc_smem_u32(bar)) : "memory");
}

// mode 0 tiled identity | 1 gather4 x32 lanes | 2 gather4 one lane | 5 cp.async 4rows x128B per warp-instr
// | 6 cp.async 1 row x 512 B per warp-instr | 7 bulk1d 512 B pieces | 8 bulk1d 2 KB pieces | 9 bulk1d 4 KB rows
__global__ void __launch_bounds__(192, 1)
k_gather(const __grid_constant__ CUtensorMap tm, const __nv_bfloat16* F, const int* rows, int nrows, int D, int mode,
         unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_b = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_b = full_b + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool cpa = (mode == 5 || mode == 6);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_b[s], cpa ? 128 : 1); mbar_init(&empty_b[s], 1); }
    mbar_fence_init();
  }
  __syncthreads();
  const int row_bytes = D * 2;
  // work: nrows rows x row_bytes, in stages of 16 KB.  stage geometry by mode:
  //   modes 0,1,2,5: 128 rows x 128 B ; mode 6,7: 32 rows x 512 B ; mode 8: 8 rows x 2 KB ; mode 9: 4 rows x 4 KB
  int srows = 128, sbytes = 128;
  if (mode == 6 || mode == 7) { srows = 32; sbytes = 512; }
  if (mode == 8) { srows = 8; sbytes = 2048; }
  if (mode == 9) { srows = 4; sbytes = 4096; }
  const int pieces = row_bytes / sbytes;            // stages per row tile
  const int n_tiles = nrows / srows;
  if (warp < 4) {
    // ---- producers
    uint32_t stage = 0, phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int p = 0; p < pieces; ++p) {
        if (mode <= 2 || mode >= 7) {
          if (warp == 0) {
            if (lane == 0) { mbar_wait(&empty_b[stage], phase ^ 1); mbar_expect_tx(&full_b[stage], STAGE_BYTES); }
            __syncwarp();
            uint8_t* dst = smem + stage * STAGE_BYTES;
            if (mode == 0) { if (lane == 0) tma_load_2d(dst, &tm, &full_b[stage], p * 64, t * 128); }
            else if (mode == 1) {
              const int* r = rows + t * 128 + lane * 4;
              tma_gather4(dst + lane * 512, &tm, &full_b[stage], p * 64, r[0], r[1], r[2], r[3]);
            } else if (mode == 2) {
              if (lane == 0) for (int j = 0; j < 32; ++j) {
                const int* r = rows + t * 128 + j * 4;
                tma_gather4(dst + j * 512, &tm, &full_b[stage], p * 64, r[0], r[1], r[2], r[3]);
              }
            } else {  // bulk1d: srows pieces of sbytes
              if (lane < srows) {
                const int r = rows[t * srows + lane];
                bulk1d(dst + lane * sbytes, reinterpret_cast<const uint8_t*>(F) + (size_t)r * row_bytes + (size_t)p * sbytes,
                       sbytes, &full_b[stage]);
              }
            }
          }
        } else {
          // cp.async by 4 warps (128 threads x 16 B = 2 KB per instruction round; 8 rounds per stage)
          if (threadIdx.x == 0) mbar_wait(&empty_b[stage], phase ^ 1);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const uint32_t dst0 = tc_smem_u32(smem + stage * STAGE_BYTES);
          const int tid = threadIdx.x;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int e = it * 128 + tid;              // 16-byte element of the stage
            const int rr = e / (sbytes / 16), cc = e % (sbytes / 16);
            const int r = rows[t * srows + rr];
            cpasync16(dst0 + e * 16, reinterpret_cast<const uint8_t*>(F) + (size_t)r * row_bytes + (size_t)p * sbytes + cc * 16);
          }
          cpasync_arrive_noinc(&full_b[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    // ---- consumer: touch one word, release
    uint32_t stage = 0, phase = 0; unsigned acc = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
      for (int p = 0; p < pieces; ++p) {
        mbar_wait(&full_b[stage], phase);
        acc += reinterpret_cast<const unsigned*>(smem + stage * STAGE_BYTES)[lane * 37];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_b[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    if (acc == 0x12345678u) sink[0] = acc;
  }
}


// fwd-like: stage = 128 rows x 256 B (interleaved hi|lo planes), NST stages, each of the 128 producer threads keeps
// 16 source pointers and issues 16 cp.async per stage; consumer waits `delay` cycles before releasing the stage.
template <int NST>
__global__ void __launch_bounds__(192, 1)
k_fwdlike(const uint8_t* F, const int* rows, int nrows, int row_bytes, int chunks_per_unit, int delay, int swz, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  const int SB = 32768;
  uint64_t* full_b = reinterpret_cast<uint64_t*>(smem + NST * SB);
  uint64_t* empty_b = full_b + NST;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < NST; ++s) { mbar_init(&full_b[s], 128); mbar_init(&empty_b[s], 1); } mbar_fence_init(); }
  __syncthreads();
  const int chunks_total = row_bytes / 256;
  const int ksplit = chunks_total / chunks_per_unit;
  const int n_units = (nrows / 128) * ksplit;
  if (warp < 4) {
    const int tid = threadIdx.x, sub = tid >> 4, e = tid & 15, plane = e >> 3, c16 = e & 7;
    uint32_t stage = 0, phase = 0;
    for (int w = blockIdx.x; w < n_units; w += gridDim.x) {
      const int tile = w / ksplit, ks = w - tile * ksplit;
      const uint8_t* src[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) src[it] = F + (size_t)rows[tile * 128 + it * 8 + sub] * row_bytes + e * 16;
      for (int c = 0; c < chunks_per_unit; ++c) {
        const int chunk = ks * chunks_per_unit + c;
        mbar_wait(&empty_b[stage], phase ^ 1);
        const uint32_t sA = tc_smem_u32(smem + stage * SB) + plane * 16384;
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int r = it * 8 + sub;
          cpasync16(sA + r * 128 + (((swz ? (c16 ^ (r & 7)) : c16)) << 4), src[it] + (size_t)chunk * 256);
        }
        cpasync_arrive_noinc(&full_b[stage]);
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    uint32_t stage = 0, phase = 0; unsigned acc = 0;
    for (int w = blockIdx.x; w < n_units; w += gridDim.x)
      for (int c = 0; c < chunks_per_unit; ++c) {
        mbar_wait(&full_b[stage], phase);
        acc += reinterpret_cast<const unsigned*>(smem + stage * SB)[lane * 37];
        if (delay) { const long long t0 = clock64(); while (clock64() - t0 < delay) {} }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_b[stage]);
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    if (acc == 0x12345678u) sink[0] = acc;
  }
}

template <int NST>
static void run_fwdlike(const char* name, const uint8_t* F, const int* rows, int nrows, int D, int cpu, int delay, int swz, unsigned* sink) {
  const size_t smem = NST * 32768 + 2 * NST * 8 + 1024;
  cudaFuncSetAttribute(k_fwdlike<NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) k_fwdlike<NST><<<148, 192, smem>>>(F, rows, nrows, D * 4, cpu, delay, swz, sink);
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) k_fwdlike<NST><<<148, 192, smem>>>(F, rows, nrows, D * 4, cpu, delay, swz, sink);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-60s %8.1f us  %8.1f GB/s  %s\n", name, ms / 5 * 1e3, (double)nrows * D * 4 / (ms / 5) / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char** argv) {
  const int I = 100000, D = 2048, nrows = 32768;
  __nv_bfloat16* F; int* rows; unsigned* sink;
  cudaMalloc(&F, (size_t)I * D * 2); cudaMemset(F, 0, (size_t)I * D * 2);
  cudaMalloc(&rows, nrows * 4); cudaMalloc(&sink, 4);
  std::vector<int> h(nrows); srand(1);
  for (int i = 0; i < nrows; ++i) h[i] = (int)(((unsigned)rand() * 2654435761u) % I);
  cudaMemcpy(rows, h.data(), nrows * 4, cudaMemcpyHostToDevice);
  const size_t smem = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;
  cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Cfg { const char* name; int mode, br, sw, promo; } cfgs[] = {
      {"tiled identity 128x128B sw128", 0, 128, 1, 1},
      {"gather4 x32 lanes sw128 promo128", 1, 1, 1, 1},
      {"gather4 x32 lanes sw128 promo256", 1, 1, 1, 2},
      {"gather4 x32 lanes sw128 promo none", 1, 1, 1, 0},
      {"gather4 x32 lanes no swizzle", 1, 1, 0, 1},
      {"gather4 one lane sw128", 2, 1, 1, 1},
      {"cp.async 16B: 4 rows x 128 B per warp-instr", 5, 1, 1, 1},
      {"cp.async 16B: 1 row x 512 B per warp-instr", 6, 1, 1, 1},
      {"bulk1d 512 B pieces", 7, 1, 1, 1},
      {"bulk1d 2 KB pieces", 8, 1, 1, 1},
      {"bulk1d 4 KB rows", 9, 1, 1, 1},
  };
  {
    // fwd-like study on a full 8 KB-per-row plane pair (F2: 100000 x 8192 B)
    uint8_t* F2; cudaMalloc(&F2, (size_t)I * D * 4); cudaMemset(F2, 0, (size_t)I * D * 4);
    const uint8_t* f = F2;
    run_fwdlike<5>("fwdlike 5x32KB, 8 chunks/unit (ksplit 4), no delay", f, rows, nrows, D, 8, 0, 1, sink);
    run_fwdlike<5>("fwdlike 5x32KB, 32 chunks/unit (ksplit 1), no delay", f, rows, nrows, D, 32, 0, 1, sink);
    run_fwdlike<5>("fwdlike 5x32KB, 8 chunks/unit, delay 500 cyc", f, rows, nrows, D, 8, 500, 1, sink);
    run_fwdlike<5>("fwdlike 5x32KB, 8 chunks/unit, delay 1000 cyc", f, rows, nrows, D, 8, 1000, 1, sink);
    run_fwdlike<5>("fwdlike 5x32KB, 8 chunks/unit, no swizzle", f, rows, nrows, D, 8, 0, 0, sink);
    run_fwdlike<6>("fwdlike 6x32KB, 8 chunks/unit, no delay", f, rows, nrows, D, 8, 0, 1, sink);
    run_fwdlike<3>("fwdlike 3x32KB, 8 chunks/unit, no delay", f, rows, nrows, D, 8, 0, 1, sink);
    cudaFree(F2);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (auto& c : cfgs) {
    CUtensorMap tm = make_map(F, I, D, 64, c.br, c.sw, c.promo);
    for (int grid : {148, 296}) {
      if (grid == 296 && c.mode != 1 && c.mode != 6 && c.mode != 9) continue;
      const int g = grid == 296 ? 148 : grid;   // 1 CTA/SM by launch bounds + smem; 296 = two launches' worth of rows
      (void)g;
      for (int w = 0; w < 2; ++w) k_gather<<<148, 192, smem>>>(tm, F, rows, nrows, D, c.mode, sink);
      cudaEventRecord(e0);
      const int iters = 5;
      for (int w = 0; w < iters; ++w) k_gather<<<148, 192, smem>>>(tm, F, rows, nrows, D, c.mode, sink);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%-48s %8.1f us  %8.1f GB/s  %s\n", c.name, ms / iters * 1e3, (double)nrows * D * 2 / (ms / iters) / 1e6,
             err == cudaSuccess ? "" : cudaGetErrorString(err));
      break;
    }
  }
  return 0;
}
