// FvxComm (include/fvx.h): internal layout, shared between fvx_comm.cu and fvx_train_sharded.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define FVX_COMM_EVENTS 8
#define FVX_COMM_MAX_RANKS 8
struct FvxComm {
  void* nccl[2];          // [0]: collectives on the caller's stream (S, dE); [1]: on the side stream (WU, RU)
  int rank, world;
  cudaStream_t side;      // the side stream of the sharded step
  cudaEvent_t ev[FVX_COMM_EVENTS];
  // peer-mapped arena (fvx_comm_arena): the same allocation on every rank, each rank's copy mapped into every
  // other rank's address space (CUDA IPC over NVLink): kernels store straight into their peers' buffers
  uint8_t* arena;                         // this rank's copy
  uint8_t* peer[FVX_COMM_MAX_RANKS];      // peer[r]: rank r's copy as seen from here (peer[rank] == arena)
  size_t arena_bytes;
  uint32_t epoch;                         // cross-GPU barriers passed so far (same on every rank)
};
// all-reduce (sum, fp32, in place) on communicator `which` (0 / 1), enqueued on `st`
int fvx_comm_allreduce(FvxComm* c, int which, float* buf, size_t n, cudaStream_t st);
// in place over `world` equal segments of `seg` floats at `buf` (rank r's segment: buf + r * seg)
int fvx_comm_allgather(FvxComm* c, int which, float* buf, size_t seg, cudaStream_t st);
int fvx_comm_reducescatter(FvxComm* c, int which, float* buf, size_t seg, cudaStream_t st);
