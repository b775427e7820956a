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

int fvx_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;  // B200
  }
  return sms;
}

extern "C" {
int fvx_abi_version(void) { return FVX_ABI_VERSION; }
const char* fvx_last_error(void) { return g_err; }
int fvx_sizeof_model(void) { return (int)sizeof(FvxModel); }
int fvx_sizeof_table(void) { return (int)sizeof(FvxTable); }
}
