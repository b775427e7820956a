// FvxComm (include/fvx.h): internal layout, shared between fvx_comm.cu and fvx_train_sharded.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#define FVX_COMM_EVENTS 6
struct FvxComm {
  void* nccl[2];          // [0]: collectives on the caller's stream (S, dE); [1]: on the side stream (WU, RU)
  int rank, world;
  cudaStream_t side;      // the side stream of the sharded step
  cudaEvent_t ev[FVX_COMM_EVENTS];
};
// all-reduce (sum, fp32, in place) on communicator `which` (0 / 1), enqueued on `st`
int fvx_comm_allreduce(FvxComm* c, int which, float* buf, size_t n, cudaStream_t st);
// in place over `world` equal segments of `seg` floats at `buf` (rank r's segment: buf + r * seg)
int fvx_comm_allgather(FvxComm* c, int which, float* buf, size_t seg, cudaStream_t st);
int fvx_comm_reducescatter(FvxComm* c, int which, float* buf, size_t seg, cudaStream_t st);
