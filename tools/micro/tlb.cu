// In-situ latency probe: G CTAs, each chasing dependent random loads inside its own window of `win` bytes, windows
// spread over `span` bytes.  Reports cycles per dependent load (thread 0 of CTA 0) for several (G, threads, span).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill(unsigned long long* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = i * 2654435761ull + 12345ull;
}
__global__ void chase(const unsigned long long* buf, size_t stride_elems, size_t win_elems, int iters, long long* out, int mlp) {
  const unsigned long long* w = buf + (size_t)blockIdx.x * stride_elems;
  size_t idx[4];
  for (int m = 0; m < 4; ++m) idx[m] = (threadIdx.x * 7919ull + m * 104729ull + blockIdx.x * 31ull) % win_elems;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    for (int m = 0; m < 4; ++m) if (m < mlp) idx[m] = (size_t)(__ldg(w + idx[m]) % win_elems);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters + (idx[0] + idx[1] + idx[2] + idx[3] == 1 ? 1 : 0);
}
int main() {
  const size_t total = (size_t)7 << 30;  // 7 GB
  unsigned long long* buf; cudaMalloc(&buf, total);
  fill<<<4096, 256>>>(buf, total / 8);
  long long* out; cudaMalloc(&out, 1024 * 8);
  cudaDeviceSynchronize();
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  struct Cfg { int G, T; size_t win_mb, stride_mb; int mlp; } cfgs[] = {
    {1, 32, 40, 48, 1}, {1, 512, 40, 48, 1}, {148, 32, 40, 48, 1}, {148, 512, 40, 48, 1}, {148, 512, 40, 48, 4},
    {148, 512, 1, 1, 1}, {148, 512, 1, 48, 1}, {148, 64, 40, 48, 1}, {148, 512, 8, 48, 1}, {148, 512, 2, 48, 1}, {36, 512, 40, 48, 1}};
  for (auto& c : cfgs) {
    size_t stride = c.stride_mb * (1 << 20) / 8, win = c.win_mb * (1 << 20) / 8;
    for (int rep = 0; rep < 2; ++rep) chase<<<c.G, c.T>>>(buf, stride, win, 64, out, c.mlp);
    cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, out, c.G * 8, cudaMemcpyDeviceToHost);
    long long mx = 0, sum = 0; for (int i = 0; i < c.G; ++i) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
    printf("G=%3d T=%3d win=%zuMB stride=%zuMB mlp=%d : avg %lld max %lld cycles/trip\n", c.G, c.T, c.win_mb, c.stride_mb, c.mlp, sum / c.G, mx);
  }
  return 0;
}
