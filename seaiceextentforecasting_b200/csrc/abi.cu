// Error reporting and device facts for libsie_b200.
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void sie_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Device facts are read once per device and cached: the entry points only enqueue kernels, they do not query the
// driver or set function attributes on every call (a sweep step enqueues ~60 kernels).
static SieDevice g_dev[SIE_MAX_DEVICES];
static std::mutex g_dev_mutex;

const SieDevice* sie_device(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= SIE_MAX_DEVICES) {
    (void)cudaGetLastError();
    sie_set_error("no CUDA device (or device ordinal >= %d)", SIE_MAX_DEVICES);
    return nullptr;
  }
  SieDevice* d = &g_dev[dev];
  if (!d->ready) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (!d->ready) {
      int v = 0;
      d->ordinal = dev;
      cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); d->sm_count = v;
      cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); d->max_smem_optin = v;
      cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev); d->smem_per_sm = v;
      cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev); d->l2_bytes = (size_t)v;
      for (int i = 0; i < SIE_ATTR_SLOTS; ++i) d->attr_smem[i] = -1;
      d->ready = 1;
    }
  }
  return d;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when the requested size exceeds what this kernel (slot) was
// last given on this device.
int sie_ensure_smem(const SieDevice* d, int slot, const void* func, size_t bytes) {
  SieDevice* m = const_cast<SieDevice*>(d);
  if ((long long)bytes <= (long long)m->attr_smem[slot]) return SIE_OK;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if ((long long)bytes <= (long long)m->attr_smem[slot]) return SIE_OK;
  if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    sie_set_error("cudaFuncSetAttribute(%zu B of dynamic shared memory) failed: %s", bytes,
                  cudaGetErrorString(cudaGetLastError()));
    return SIE_ERR_LAUNCH;
  }
  m->attr_smem[slot] = (long long)bytes;
  return SIE_OK;
}

// Tensor map (TMA descriptor) of a batch of square FP64 matrices M [B][ld][ld]: box = box_rows x box_cols elements, no
// swizzle.  The driver entry point is resolved once through the runtime (no link against libcuda).  `out` = 128 bytes.
int sie_tensor_map_f64_3d(void* out, const double* base, int B, int ld, int box_cols, int box_rows) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    else
      (void)cudaGetLastError();
  });
  if (!encode) {
    sie_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SIE_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)ld, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(double), (cuuint64_t)ld * (cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult rc = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                             const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    sie_set_error("cuTensorMapEncodeTiled failed (CUresult %d; B=%d ld=%d box=%dx%d)", (int)rc, B, ld, box_rows, box_cols);
    return SIE_ERR_LAUNCH;
  }
  return SIE_OK;
}

extern "C" int sie_abi_version(void) { return 2; }
extern "C" const char* sie_last_error(void) { return g_err; }

extern "C" int sie_device_info(int* sm_count, int* max_smem_optin, size_t* l2_bytes) {
  const SieDevice* d = sie_device();
  if (!d) return SIE_ERR_LAUNCH;
  if (sm_count) *sm_count = d->sm_count;
  if (max_smem_optin) *max_smem_optin = d->max_smem_optin;
  if (l2_bytes) *l2_bytes = d->l2_bytes;
  return SIE_OK;
}
