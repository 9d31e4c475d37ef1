// K6: area-weighted node time series, covariance links, strength and strength map.
// Reference: Network.intra_links, ComplexNetworks.py:283-326.
//
// anomaly[A][t] = nansum over the grid of data*scale restricted to V[A]; numpy reduces axes (0,1) of the
// (X,Y,T) temporary in row-major cell order, so the series is a *sequential* sum over member cells in
// ascending flat-cell order with the product rounded before each add.  The kernel reproduces that order
// (no atomics, no tree): one CTA per (network, area), thread t owns time step t; the area's member list (list
// order, from K4/K5) is rank-sorted into ascending cell order in shared memory and walked 8 cells at a time with
// all loads in flight before the ordered adds.  HBM traffic: 8*(member cells)*T + 4*(member cells) bytes per area.
#include "common.cuh"

namespace {

constexpr int NS_LIST = 2048;     // member cells sorted on chip; larger areas scan the label row instead
constexpr int AREA_SLOTS = 64;    // grid.x: CTA x walks areas x, x+64, ... of its network (n_areas is only known on the device)

__global__ void __launch_bounds__(64) k_node_series(const double* __restrict__ dt, const double* __restrict__ scale,
                                                    const int32_t* __restrict__ job_T,
                                                    const int32_t* __restrict__ n_areas,
                                                    const int32_t* __restrict__ area_cells,
                                                    const int32_t* __restrict__ area_start,
                                                    const int32_t* __restrict__ label, int C, int Tstride,
                                                    int MA, double* __restrict__ anomaly) {
  __shared__ int32_t raw[NS_LIST], srt[NS_LIST];
  const int b = blockIdx.y;
  const int nA = n_areas[b];
  const int T = job_T[b];
  const int lane = threadIdx.x & 31;
  const int32_t* lab = label + (size_t)b * C;
  const double* d = dt + (size_t)b * C * Tstride;
  for (int a = blockIdx.x; a < nA; a += gridDim.x) {
    double* out = anomaly + ((size_t)b * MA + a) * Tstride;
    const int st = area_start[(size_t)b * (MA + 1) + a];
    const int n = area_start[(size_t)b * (MA + 1) + a + 1] - st;
    if (n <= NS_LIST) {
      // member cells in ascending flat-cell order: rank sort of the area's list (cells are distinct)
      __syncthreads();                                   // previous area's list no longer read
      for (int i = threadIdx.x; i < n; i += blockDim.x) raw[i] = area_cells[(size_t)b * C + st + i];
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ci = raw[i];
        int r = 0;
        for (int q = 0; q < n; ++q) r += (raw[q] < ci);
        srt[r] = ci;
      }
      __syncthreads();
      for (int t0 = 0; t0 < Tstride; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        double acc = 0.0;
        if (t < T) {
          for (int i0 = 0; i0 < n; i0 += 8) {            // 8 member cells in flight, added in order
            double pr[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int cc = srt[min(i0 + u, n - 1)];
              pr[u] = __dmul_rn(d[(size_t)cc * Tstride + t], scale[cc]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (i0 + u < n) {
                double x = pr[u];
                if (x != x) x = 0.0;                      // np.nansum
                acc = __dadd_rn(acc, x);
              }
            }
          }
        }
        if (t < Tstride) out[t] = (t < T) ? acc : 0.0;
      }
      continue;
    }
    for (int t0 = 0; t0 < Tstride; t0 += blockDim.x) {
      const int t = t0 + threadIdx.x;
      double acc = 0.0;
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        unsigned m = __ballot_sync(0xffffffffu, c < C && lab[c] == a);
        while (m) {
          const int cc = c0 + __ffs(m) - 1;
          m &= m - 1;
          if (t < T) {
            double pr = __dmul_rn(d[(size_t)cc * Tstride + t], scale[cc]);
            if (pr != pr) pr = 0.0;                       // np.nansum
            acc = __dadd_rn(acc, pr);
          }
        }
      }
      if (t < Tstride) out[t] = (t < T) ? acc : 0.0;
    }
  }
}

// links[a][a2] = population covariance of the two node series (pearsonr*sdA*sdA2 in the reference, :309-316),
// 0 on the diagonal; strength[a] = sum |links[a][:]| (:318-323).  One CTA per (network, area a).
__global__ void __launch_bounds__(128) k_links(const double* __restrict__ anomaly, const int32_t* __restrict__ job_T,
                                               const int32_t* __restrict__ n_areas, int Tstride, int MA,
                                               double* __restrict__ links, double* __restrict__ strength) {
  extern __shared__ double sh[];     // [Tstride] centred series of area a
  __shared__ double red[4];
  const int b = blockIdx.y;
  const int nA = n_areas[b];
  const int T = job_T[b];
  const double* base = anomaly + (size_t)b * MA * Tstride;
  for (int a = blockIdx.x; a < nA; a += gridDim.x) {
  __syncthreads();                     // previous area's sh / red no longer read
  const double* xa = base + (size_t)a * Tstride;
  double ma = 0.0;
  for (int t = 0; t < T; ++t) ma += xa[t];
  ma /= (double)T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) sh[t] = xa[t] - ma;
  __syncthreads();
  double sabs = 0.0;
  for (int a2 = threadIdx.x; a2 < nA; a2 += blockDim.x) {
    double v = 0.0;
    if (a2 != a) {
      const double* xb = base + (size_t)a2 * Tstride;
      double mb = 0.0;
      for (int t = 0; t < T; ++t) mb += xb[t];
      mb /= (double)T;
      double s = 0.0;
      for (int t = 0; t < T; ++t) s += sh[t] * (xb[t] - mb);
      v = s / (double)T;
    }
    links[((size_t)b * MA + a) * MA + a2] = v;
    if (v == v) sabs += fabs(v);      // nansum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sabs += __shfl_xor_sync(0xffffffffu, sabs, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sabs;
  __syncthreads();
  if (threadIdx.x == 0) strength[(size_t)b * MA + a] = (red[0] + red[1]) + (red[2] + red[3]);
  }
}

__global__ void k_strengthmap(const int32_t* __restrict__ label, const double* __restrict__ strength, int B, int C,
                              int MA, double* __restrict__ smap) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C) return;
  const int b = (int)(i / C);
  const int l = label[i];
  smap[i] = (l >= 0) ? strength[(size_t)b * MA + l] : sie_nan();
}

}  // namespace

extern "C" int sie_intra_links(const double* dt, const double* scale, const int32_t* job_T,
                               const int32_t* area_cells, const int32_t* area_start, const int32_t* n_areas,
                               const int32_t* label, int B, int C, int Tstride, int max_areas, double* anomaly,
                               double* links, double* strength, double* strengthmap, void* stream) {
  SIE_CHECK_ARG(dt && scale && job_T && area_cells && area_start && n_areas && label && anomaly && links && strength &&
                    strengthmap, "null pointer");
  SIE_CHECK_ARG(B > 0 && C > 0 && Tstride > 0 && max_areas > 0, "non-positive size");
  SIE_CHECK_ARG(B <= 65535, "at most 65535 jobs per call");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(max_areas < AREA_SLOTS ? max_areas : AREA_SLOTS, B);
  k_node_series<<<grid, 64, 0, st>>>(dt, scale, job_T, n_areas, area_cells, area_start, label, C, Tstride, max_areas,
                                     anomaly);
  SIE_CHECK_LAUNCH();
  k_links<<<grid, 128, sizeof(double) * (size_t)Tstride, st>>>(anomaly, job_T, n_areas, Tstride, max_areas, links,
                                                              strength);
  SIE_CHECK_LAUNCH();
  const long long total = (long long)B * C;
  k_strengthmap<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(label, strength, B, C, max_areas, strengthmap);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}
