// libfvx: library-level entry points (version, error text, struct sizes).
#include <stdarg.h>
#include <string.h>

#include "fvx_common.cuh"

static thread_local char g_err[512] = "";

void fvx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fvx_cur_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= FVX_MAX_DEV) dev = 0;
  return dev;
}

int fvx_num_sms() {
  static int sms[FVX_MAX_DEV];
  const int dev = fvx_cur_device();
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;  // B200
    sms[dev] = n;
  }
  return sms[dev];
}

int fvx_ensure_smem(const void* func, FvxSmemMark* mark, size_t smem, const char* who) {
  const int dev = fvx_cur_device();
  if (smem <= 48 * 1024 || smem <= mark->v[dev]) return 0;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) FVX_FAIL(-3, "%s: cannot set %zu B of dynamic shared memory: %s", who, smem, cudaGetErrorString(e));
  mark->v[dev] = smem;
  return 0;
}

extern "C" {
int fvx_abi_version(void) { return FVX_ABI_VERSION; }
const char* fvx_last_error(void) { return g_err; }
int fvx_sizeof_model(void) { return (int)sizeof(FvxModel); }
int fvx_sizeof_table(void) { return (int)sizeof(FvxTable); }
int fvx_sizeof_shard_ws(void) { return (int)sizeof(FvxShardWs); }
int fvx_sizeof_eval_ws(void) { return (int)sizeof(FvxEvalWs); }
}
