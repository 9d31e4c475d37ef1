// Cost of one numpy-order row sum (len elements) per thread out of shared memory / global memory, and per 8-lane group.
#include <cstdio>
#include "../../seaiceextentforecasting_b200/csrc/common.cuh"
void sie_set_error(const char*, ...) {}
constexpr int NT = 512;
__global__ void __launch_bounds__(NT, 1) probe(const double* g, double* out, long long* cyc, int len, int nbb) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, j = tid & 7;
  unsigned gmask = 0xffu << (lane & 24);
  if (len < 0) gmask = 0;   // keep it a runtime value
  if (nbb == -7) gmask = 0xffffffffu;
  for (int i = tid; i < 16 * 288; i += NT) sm[i] = g[i];
  __syncthreads();
  double acc = 0.0;
  long long t0, t1;
  // (b) 8-lane group per item, global memory (L2 resident)
  {
    const double* base = g + (size_t)(tid >> 3) * 1024;
    int nanc = 0;
    t0 = clock64();
    for (int rep = 0; rep < 8; ++rep) {
      double s = sie_pw_tree<6>([&](int lo, int ln) -> double {
        const double r = sie_pw_leaf8_contig(base + rep * 64 * 1024 + lo + j, ln, j, gmask);
        if (r == r) return r;
        return sie_pw_leaf8([&](int i) { return base[i]; }, lo, ln, j, gmask, nanc);
      }, len + (rep & 1));
      acc += s;
    }
    t1 = clock64();
    if (tid == 0) cyc[1] = (t1 - t0) / 8;
  }
  __syncthreads();
  // (d) 8-lane group per item, global memory, FULL-warp shuffle mask (all groups converged here)
  {
    const double* base = g + (size_t)(tid >> 3) * 1024;
    int nanc = 0;
    t0 = clock64();
    for (int rep = 0; rep < 8; ++rep) {
      double s = sie_pw_tree<6>([&](int lo, int ln) -> double {
        const double r = sie_pw_leaf8_contig(base + rep * 64 * 1024 + lo + j, ln, j, 0xffffffffu);
        if (r == r) return r;
        return sie_pw_leaf8([&](int i) { return base[i]; }, lo, ln, j, 0xffffffffu, nanc);
      }, len + (rep & 1));
      acc += s;
    }
    t1 = clock64();
    if (tid == 0) cyc[3] = (t1 - t0) / 8;
  }
  out[tid] = acc;
}
int main() {
  const size_t n = (size_t)4 * 512 * 1024 + 4096;
  double* g; cudaMalloc(&g, n * 8);
  double* h = new double[n]; for (size_t i = 0; i < n; ++i) h[i] = 1e-3 * (double)((i * 7919) % 1000) - 0.5;
  cudaMemcpy(g, h, n * 8, cudaMemcpyHostToDevice);
  double* out; cudaMalloc(&out, NT * 8);
  long long* cyc; cudaMalloc(&cyc, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 288 * 8);
  for (int len : {60, 120, 186}) {
    for (int it = 0; it < 2; ++it) probe<<<1, NT, 16 * 288 * 8>>>(g, out, cyc, len, len - 12);
    cudaDeviceSynchronize();
    long long hc[8]; cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost);
    printf("len %3d: thread/item smem %6lld | 8-lane group global(L2), group mask %6lld | same, full mask %6lld | thread/item global %6lld  (%s)\n", len, hc[0], hc[1], hc[3], hc[2], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
