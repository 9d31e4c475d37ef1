// Micro-latency probes for the domain-growth kernel's building blocks (B200). Build: nvcc -arch=sm_100a -O3 lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
#define NT 512
__global__ void fill(double* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = (double)((i * 2654435761ull + 12345ull) % n);
}
__global__ void probe(double* out, long long* cyc, const double* gsrc, int n, const double* big, size_t nbig) {
  __shared__ double sh[4096];
  __shared__ unsigned long long slots[32];
  const int tid = threadIdx.x;
  for (int i = tid; i < 4096; i += NT) sh[i] = 1.0 + i * 1e-9;
  __syncthreads();
  long long t0, t1;
  double acc = out[tid];
  // 0: dependent DADD chain (64)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) acc = __dadd_rn(acc, 1.000001);
  t1 = clock64(); if (tid == 0) cyc[0] = t1 - t0;
  __syncthreads();
  // 1: all warps: 8 independent accumulators x 64 adds (throughput)
  double r[8]; for (int q = 0; q < 8; ++q) r[q] = acc + q;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], 1.000001);
  t1 = clock64(); if (tid == 0) cyc[1] = t1 - t0;
  for (int q = 0; q < 8; ++q) acc += r[q];
  __syncthreads();
  // 2: FP64 division chain (16)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; ++i) acc = acc / (1.0 + acc * 1e-3);
  t1 = clock64(); if (tid == 0) cyc[2] = t1 - t0;
  __syncthreads();
  // 3: shuffle chain (32 x 64-bit)
  unsigned long long u = (unsigned long long)__double_as_longlong(acc);
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) { unsigned long long w = __shfl_xor_sync(0xffffffffu, u, 1 + (i & 15)); u = (w > u) ? w : u + 1; }
  t1 = clock64(); if (tid == 0) cyc[3] = t1 - t0;
  // 4: __syncthreads x 16
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; ++i) __syncthreads();
  t1 = clock64(); if (tid == 0) cyc[4] = t1 - t0;
  // 5: dependent smem load chain (32)
  int idx = tid;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) idx = (int)(sh[idx & 4095] * 7.0) + idx;
  t1 = clock64(); if (tid == 0) cyc[5] = t1 - t0;
  __syncthreads();
  // 6: dependent global gather chain (16), L2-resident small array
  int gi = tid;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; ++i) gi = (int)__ldg(gsrc + (gi % n));
  t1 = clock64(); if (tid == 0) cyc[6] = t1 - t0;
  // 7: int -> double conversion + DSETP chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) { double d = (double)(gi + i); if (d > acc) acc = d * 0.5; }
  t1 = clock64(); if (tid == 0) cyc[7] = t1 - t0;
  // 8: block argmax as in area.cu (5 shuffle levels of 2 x u64, slot write, barrier, 16-slot scan)
  t0 = clock64();
  for (int rep = 0; rep < 8; ++rep) {
    unsigned long long hi = u + tid + rep, lo = ~u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long wh = __shfl_xor_sync(0xffffffffu, hi, o), wl = __shfl_xor_sync(0xffffffffu, lo, o);
      bool tb = (wh > hi) || (wh == hi && wl > lo);
      hi = tb ? wh : hi; lo = tb ? wl : lo;
    }
    if ((tid & 31) == 0) { slots[2 * (tid >> 5)] = hi; slots[2 * (tid >> 5) + 1] = lo; }
    __syncthreads();
    unsigned long long bh = slots[0], bl = slots[1];
#pragma unroll
    for (int w = 1; w < 16; ++w) { unsigned long long wh = slots[2 * w], wl = slots[2 * w + 1]; bool tb = (wh > bh) || (wh == bh && wl > bl); bh = tb ? wh : bh; bl = tb ? wl : bl; }
    u += bh ^ bl;
    __syncthreads();
  }
  t1 = clock64(); if (tid == 0) cyc[8] = t1 - t0;
  // 9: dependent DRAM gather chain (16) over a huge array (big = 4 GB), one thread per warp distinct lines
  {
    size_t gj = (size_t)tid * 1000003ull + blockIdx.x * 7777777ull;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) gj = (size_t)__ldg(big + (gj % nbig)) + tid * 131ull;
    t1 = clock64(); if (tid == 0) cyc[9] = t1 - t0;
    gi += (int)gj;
  }
  out[tid] = acc + (double)u + idx + gi;
}
int main() {
  double *out, *g; long long* cyc;
  cudaMalloc(&out, NT * 8); cudaMemset(out, 0, NT * 8);
  const int n = 1 << 20;
  cudaMalloc(&g, n * 8);
  double* h = new double[n]; for (int i = 0; i < n; ++i) h[i] = (double)((i * 7919LL + 13) % n);
  cudaMemcpy(g, h, n * 8, cudaMemcpyHostToDevice);
  const size_t nbig = (size_t)1 << 29;   // 4 GB
  double* big; cudaMalloc(&big, nbig * 8);
  fill<<<4096, 256>>>(big, nbig);
  cudaMalloc(&cyc, 16 * 8); cudaMemset(cyc, 0, 128);
  for (int it = 0; it < 2; ++it) probe<<<1, NT>>>(out, cyc, g, n, big, nbig);
  cudaDeviceSynchronize();
  long long hc[16]; cudaMemcpy(hc, cyc, 128, cudaMemcpyDeviceToHost);
  const char* names[] = {"dadd chain x64", "dadd 8acc x64 (512 thr)", "ddiv chain x16", "shfl64+sel chain x32", "syncthreads x16", "smem dep chain x32", "global L2 gather chain x16", "i2d+dsetp chain x32", "block argmax x8", "DRAM gather chain x16 (512 thr)"};
  for (int i = 0; i < 10; ++i) printf("%-28s %8lld cycles\n", names[i], hc[i]);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
