// K1: per-cell OLS detrend, node mask, node compaction and unit-norm rows.
// Reference: detrend() north/June1st.py:179-194; node mask + np.corrcoef row centring
// ComplexNetworks.py:32-34,:37; NaN sentinel cell ComplexNetworks.py:50-51.
//
// HBM-bound streaming: a CTA stages 128 consecutive series (one contiguous run) in shared memory with coalesced
// loads, one thread per cell fits and removes its line from there, coalesced stores write the residuals.
// Algorithmic bytes: 16*C*T per window (read raw, write residuals) + 8*N*Tp for the compacted z rows.
#include "common.cuh"

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int DT_CELLS = 128;     // cells (= threads) per CTA; their series are staged in shared memory

// One thread per (job, cell).  The CTA's 128 series are one contiguous run of the field: it is copied to shared
// memory with coalesced loads (row stride padded to an odd number of doubles, so the per-thread walks are
// bank-conflict free), each thread fits and removes its own line with plain sequential sums (no shuffles, ~20
// instructions per sample instead of a warp per 42-sample series), and the residuals go back with coalesced stores.
__global__ void __launch_bounds__(DT_CELLS) k_detrend_cells(
    const double* __restrict__ fields, const int32_t* __restrict__ job_field,
    const int32_t* __restrict__ job_T, int C, int Tstride, int ld, int do_detrend,
    double* __restrict__ dt, double* __restrict__ trend, int32_t* __restrict__ cell_flag,
    int32_t* __restrict__ first_nan_cell) {
  extern __shared__ double sm[];          // [DT_CELLS][ld]
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * DT_CELLS;
  const int nc = min(DT_CELLS, C - c0);
  const int T = job_T[b];
  const double* src = fields + ((size_t)job_field[b] * C + c0) * Tstride;
  double* dst = dt + ((size_t)b * C + c0) * Tstride;
  const int tid = threadIdx.x;
  // only the job's window (the first T samples of every series) is staged: runs of T doubles out of rows of Tstride,
  // flat index -> (cell, t) by a multiply-high with the rounded-up reciprocal of T (exact for idx < 2^16 * T), 8 loads in
  // flight per thread before the first shared-memory store (a load -> store loop pays one memory round trip per iteration)
  const int total = nc * T;
  const unsigned rcpT = 0xffffffffu / (unsigned)T + 1u;
  for (int base = tid; base < total; base += 8 * DT_CELLS) {
    double v[8];
    int so[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * DT_CELLS;
      const int cell = (int)__umulhi((unsigned)idx, rcpT), t = idx - cell * T;
      so[u] = cell * ld + t;
      v[u] = idx < total ? src[cell * Tstride + t] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (base + u * DT_CELLS < total) sm[so[u]] = v[u];
  }
  __syncthreads();
  if (tid < nc) {
    double* y = sm + tid * ld;
    const int c = c0 + tid;
    // pass 1: NaN census, sum of y
    double sy = 0.0;
    int n_nan = 0;
    for (int t = 0; t < T; ++t) {
      const double v = y[t];
      if (v != v) ++n_nan; else sy += v;
    }
    int flag;
    if (n_nan > 0) {
      // any NaN: linregress propagates NaN through the whole row (all-NaN cells are skipped and stay NaN)
      atomicMin(first_nan_cell + b, c);
      flag = 0;
      if (trend) { trend[((size_t)b * C + c) * 2] = sie_nan(); trend[((size_t)b * C + c) * 2 + 1] = sie_nan(); }
      if (do_detrend) {
        for (int t = 0; t < T; ++t) y[t] = sie_nan();
      } else if (n_nan < T) {
        // pass-through mode keeps a partially-NaN series visible: it is a node whose correlations are NaN
        // (np.nanmax ignores the NaNs, np.corrcoef does not) -- ComplexNetworks.py:32-34
        double mx = -INFINITY;
        for (int t = 0; t < T; ++t) { const double v = y[t]; if (v == v) mx = fmax(mx, v); }
        flag = (fabs(mx) > 0.0) ? 1 : 0;
      }
    } else {
      double mx = -INFINITY;
      if (do_detrend) {
        const double ym = sy / (double)T;
        const double xm = 0.5 * (double)(T - 1);
        double sxy = 0.0, sxx = 0.0;
        for (int t = 0; t < T; ++t) {
          const double dx = (double)t - xm;
          sxy += dx * (y[t] - ym);
          sxx += dx * dx;
        }
        sxy /= (double)T;
        sxx /= (double)T;
        const double slope = sxy / sxx;
        const double icpt = ym - slope * xm;
        for (int t = 0; t < T; ++t) {
          // y - ((slope*t) + intercept): product rounded before the add, like the numpy expression
          const double line = __dadd_rn(__dmul_rn(slope, (double)t), icpt);
          const double r = __dsub_rn(y[t], line);
          y[t] = r;
          mx = fmax(mx, r);
        }
        if (trend) {
          trend[((size_t)b * C + c) * 2] = slope;
          trend[((size_t)b * C + c) * 2 + 1] = icpt;
        }
      } else {
        for (int t = 0; t < T; ++t) mx = fmax(mx, y[t]);
      }
      flag = (fabs(mx) > 0.0) ? 1 : 0;   // |nanmax| > 0
    }
    cell_flag[(size_t)b * C + c] = flag;
  }
  if (!do_detrend) return;
  __syncthreads();
  // residuals back: the window only (samples T.. of dt are never consumed: K1's z rows, K6 and the host read [0, T))
  for (int base = tid; base < total; base += 8 * DT_CELLS) {
    double v[8];
    int go[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * DT_CELLS;
      const int cell = (int)__umulhi((unsigned)idx, rcpT), t = idx - cell * T;
      go[u] = cell * Tstride + t;
      v[u] = idx < total ? sm[cell * ld + t] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (base + u * DT_CELLS < total) dst[go[u]] = v[u];
  }
}

// One CTA per job: order-preserving compaction of the node flags (ascending flat cell id, :37).
__global__ void __launch_bounds__(1024) k_compact_nodes(int C, int ldn, int32_t* __restrict__ cell_node,
                                                        int32_t* __restrict__ node_cell,
                                                        int32_t* __restrict__ n_nodes,
                                                        int32_t* __restrict__ first_nan_cell,
                                                        int32_t* __restrict__ status) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int b = blockIdx.x;
  int32_t* cn = cell_node + (size_t)b * C;
  int32_t* nc = node_cell + (size_t)b * ldn;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    const int f = (c < C) ? cn[c] : 0;
    const unsigned m = __ballot_sync(0xffffffffu, f != 0);
    const int pre = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int off = base_s;
    for (int i = 0; i < w; ++i) off += warp_tot[i];
    if (c < C) {
      if (f) {
        int idx = off + pre;
        cn[c] = idx;
        if (idx < ldn) nc[idx] = c;
      } else {
        cn[c] = -1;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += warp_tot[i];
      base_s += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    n_nodes[b] = base_s;
    status[b] = (base_s > ldn) ? SIE_JOB_CAPACITY : SIE_JOB_OK;
    if (first_nan_cell[b] >= C) first_nan_cell[b] = -1;
  }
}

// Centre and scale every node series to unit norm (np.corrcoef's (x-mean)/sqrt(sum (x-mean)^2)), pad with 0.
// One warp per ZR_NODES consecutive nodes: the (short) series of all of them are in flight before the first reduction
// (a warp per node had 2 loads in flight and 10 dependent shuffles: latency-bound at 25 % of the HBM roof); the
// arithmetic per node - lane t holds samples t and t+32, xor-butterfly sums - is unchanged.
constexpr int ZR_NODES = 4;
__global__ void __launch_bounds__(256) k_zrows(const double* __restrict__ dt, const int32_t* __restrict__ job_T,
                                               const int32_t* __restrict__ node_cell,
                                               const int32_t* __restrict__ n_nodes, int B, int C, int Tstride,
                                               int Tp, int ldn, double* __restrict__ z) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int per_job = ldn / ZR_NODES;                       // ldn is a multiple of 128
  if (warp >= (long long)B * per_job) return;
  const int b = (int)(warp / per_job);
  const int n0 = (int)(warp - (long long)b * per_job) * ZR_NODES;
  const int N = min(n_nodes[b], ldn);
  const int T = job_T[b];
  double* zr0 = z + ((size_t)b * ldn + n0) * Tp;
  if (T <= 64) {
    // the whole series sits in two registers per lane: one pass over memory
    double v0[ZR_NODES], v1[ZR_NODES];
    int cell[ZR_NODES];
#pragma unroll
    for (int u = 0; u < ZR_NODES; ++u) cell[u] = (n0 + u < N) ? node_cell[(size_t)b * ldn + n0 + u] : -1;
#pragma unroll
    for (int u = 0; u < ZR_NODES; ++u) {
      const double* src = dt + ((size_t)b * C + max(cell[u], 0)) * Tstride;
      v0[u] = (cell[u] >= 0 && lane < T) ? src[lane] : 0.0;
      v1[u] = (cell[u] >= 0 && lane + 32 < T) ? src[lane + 32] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < ZR_NODES; ++u) {
      double* zr = zr0 + (size_t)u * Tp;
      if (cell[u] < 0) {                                    // rows past N are zero so padded tiles are harmless
        for (int t = lane; t < Tp; t += 32) zr[t] = 0.0;
        continue;
      }
      const double mean = warp_sum((0.0 + v0[u]) + v1[u]) / (double)T;
      const double d0 = lane < T ? v0[u] - mean : 0.0, d1 = lane + 32 < T ? v1[u] - mean : 0.0;
      const double inv = 1.0 / sqrt(warp_sum(fma(d1, d1, fma(d0, d0, 0.0))));   // q += d*d contracts to an FMA in the loop form
      if (lane < Tp) zr[lane] = lane < T ? d0 * inv : 0.0;
      if (lane + 32 < Tp) zr[lane + 32] = lane + 32 < T ? d1 * inv : 0.0;
      for (int t = lane + 64; t < Tp; t += 32) zr[t] = 0.0;
    }
    return;
  }
  for (int u = 0; u < ZR_NODES; ++u) {
    const int n = n0 + u;
    double* zr = zr0 + (size_t)u * Tp;
    if (n >= N) {
      for (int t = lane; t < Tp; t += 32) zr[t] = 0.0;
      continue;
    }
    const double* src = dt + ((size_t)b * C + node_cell[(size_t)b * ldn + n]) * Tstride;
    double s = 0.0;
    for (int t = lane; t < T; t += 32) s += src[t];
    const double mean = warp_sum(s) / (double)T;
    double q = 0.0;
    for (int t = lane; t < T; t += 32) { double d = src[t] - mean; q += d * d; }
    const double inv = 1.0 / sqrt(warp_sum(q));
    for (int t = lane; t < Tp; t += 32) zr[t] = (t < T) ? (src[t] - mean) * inv : 0.0;
  }
}

}  // namespace

extern "C" int sie_detrend_zscore(const double* fields, const int32_t* job_field, const int32_t* job_T,
                                  int B, int C, int Tstride, int Tp, int do_detrend, double* dt,
                                  double* trend, double* z, int32_t* node_cell, int32_t* cell_node,
                                  int32_t* n_nodes, int32_t* first_nan_cell, int32_t* status, int ldn,
                                  void* stream) {
  SIE_CHECK_ARG(fields && job_field && job_T && dt && z && node_cell && cell_node && n_nodes &&
                    first_nan_cell && status, "null pointer");
  SIE_CHECK_ARG(B > 0 && C > 0 && Tstride > 0 && ldn > 0, "non-positive size");
  SIE_CHECK_ARG(Tp >= Tstride && (Tp % 4) == 0, "Tp must be >= Tstride and a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(first_nan_cell, 0x7f, sizeof(int32_t) * (size_t)B, st);
  SIE_CHECK_ARG(B <= 65535, "at most 65535 jobs per call");
  const int ld = Tp | 1;                                     // odd row stride (doubles) >= Tp >= Tstride
  const size_t smem = sizeof(double) * (size_t)DT_CELLS * ld;
  SIE_CHECK_ARG(smem <= 200 * 1024, "window too long for the shared-memory staging");
  const SieDevice* dev = sie_device();
  if (!dev) return SIE_ERR_LAUNCH;
  if (int rc = sie_ensure_smem(dev, SIE_K_DETREND, (const void*)k_detrend_cells, smem)) return rc;
  k_detrend_cells<<<dim3((unsigned)((C + DT_CELLS - 1) / DT_CELLS), (unsigned)B), DT_CELLS, smem, st>>>(
      fields, job_field, job_T, C, Tstride, ld, do_detrend, dt, trend, cell_node, first_nan_cell);
  SIE_CHECK_LAUNCH();
  k_compact_nodes<<<B, 1024, 0, st>>>(C, ldn, cell_node, node_cell, n_nodes, first_nan_cell, status);
  SIE_CHECK_LAUNCH();
  SIE_CHECK_ARG((ldn % ZR_NODES) == 0, "ldn must be a multiple of 4");
  const long long zwarps = (long long)B * (ldn / ZR_NODES);
  const int wpb = 8;
  k_zrows<<<(unsigned)((zwarps + wpb - 1) / wpb), wpb * 32, 0, st>>>(dt, job_T, node_cell, n_nodes, B, C,
                                                                      Tstride, Tp, ldn, z);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}
