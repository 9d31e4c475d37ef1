// Error reporting and device facts for libsie_b200.
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

static thread_local char g_err[512] = "";

void sie_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int sie_abi_version(void) { return 1; }
extern "C" const char* sie_last_error(void) { return g_err; }

extern "C" int sie_device_info(int* sm_count, int* max_smem_optin, size_t* l2_bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    sie_set_error("sie_device_info: no CUDA device");
    return SIE_ERR_LAUNCH;
  }
  int v = 0;
  if (sm_count) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); *sm_count = v; }
  if (max_smem_optin) { cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); *max_smem_optin = v; }
  if (l2_bytes) { cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev); *l2_bytes = (size_t)v; }
  return SIE_OK;
}
