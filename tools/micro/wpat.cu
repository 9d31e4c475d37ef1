// Store-pattern probe: how fast can 148 SMs write a batch of pitch-strided matrices tile by tile, as a function of the
// contiguous run length of a tile row?  (decides the tile shape of the stored-R correlation epilogue)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o wpat wpat.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

// grid-stride over tiles; tile = ROWS rows x W doubles; matrix = [nmat][ldn][ldn]; contiguous chunk of tiles per CTA if contig
__global__ void k_tiles(double* R, int nmat, int ldn, int ROWS, int W, int contig) {
  const long long tpr = ldn / W, tpm = (long long)(ldn / ROWS) * tpr, total = tpm * nmat;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  long long per = (total + gridDim.x - 1) / gridDim.x;
  long long lo = contig ? blockIdx.x * per : blockIdx.x, hi = contig ? min(total, lo + per) : total, step = contig ? 1 : gridDim.x;
  for (long long t = lo; t < hi; t += step) {
    const long long m = t / tpm, r = t % tpm;
    const long long rb = r / tpr, cb = r % tpr;
    double* base = R + (m * ldn + rb * ROWS) * ldn + cb * W;
    for (int row = warp; row < ROWS; row += nw)
      for (int c = lane * 2; c < W; c += 64)
        *reinterpret_cast<double2*>(base + (long long)row * ldn + c) = make_double2(1.0, 2.0);
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// same, but each row run is one bulk async store from shared memory (TMA engine)
__global__ void k_tiles_bulk(double* R, int nmat, int ldn, int ROWS, int W, int contig) {
  extern __shared__ __align__(128) double buf[];
  for (int i = threadIdx.x; i < ROWS * W; i += blockDim.x) buf[i] = 3.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const long long tpr = ldn / W, tpm = (long long)(ldn / ROWS) * tpr, total = tpm * nmat;
  long long per = (total + gridDim.x - 1) / gridDim.x;
  long long lo = contig ? blockIdx.x * per : blockIdx.x, hi = contig ? min(total, lo + per) : total, step = contig ? 1 : gridDim.x;
  for (long long t = lo; t < hi; t += step) {
    const long long m = t / tpm, r = t % tpm;
    const long long rb = r / tpr, cb = r % tpr;
    double* base = R + (m * ldn + rb * ROWS) * ldn + cb * W;
    for (int row = threadIdx.x; row < ROWS; row += 32)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (long long)row * ldn),
                   "r"(smem_u32(buf + row * W)), "r"(W * 8) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int ldn = 2304, nmat = 48;                 // 48 x 42.5 MB = 2 GB
  double* R;
  size_t bytes = (size_t)nmat * ldn * ldn * 8;
  cudaMalloc(&R, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int Ws[] = {64, 128, 256, 768, 2304};
  for (int bulk = 0; bulk < 2; ++bulk)
    for (int contig = 0; contig < 2; ++contig)
      for (int ROWS : {128, 64})
        for (int W : Ws) {
          if (bulk && (size_t)ROWS * W * 8 > 200 * 1024) continue;
          float best = 1e9;
          for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (bulk) {
              cudaFuncSetAttribute(k_tiles_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * W * 8);
              k_tiles_bulk<<<148, 256, ROWS * W * 8>>>(R, nmat, ldn, ROWS, W, contig);
            } else {
              k_tiles<<<296, 256>>>(R, nmat, ldn, ROWS, W, contig);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
          }
          printf("%s contig=%d tile %3d rows x %4d B runs: %.3f ms  %.0f GB/s  (%s)\n", bulk ? "bulk" : "stg ", contig, ROWS, W * 8, best,
                 bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
  // how many storing warps per SM are needed?  (ROWS=128, 512-B runs, contiguous chunks)
  for (int thr : {32, 64, 128, 256, 512, 1024}) {
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      k_tiles<<<148, thr>>>(R, nmat, ldn, 128, 64, 1);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("stg 148 CTAs x %4d threads: %.3f ms %.0f GB/s\n", thr, best, bytes / best / 1e6);
  }
  return 0;
}
