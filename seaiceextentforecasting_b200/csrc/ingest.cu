// Ingest ("next" rows of SURVEY.md 8(f), the steps immediately before the hot path): NSIDC binary decode + daily ->
// monthly mean, polar-hole fill, and the linear regrid of the 25 km field onto the analysis grid.
// Reference: readNSIDC, north/September1st.py:72-139 (decode :93-104 / :121-127, `>1 -> NaN` :128, polar hole
// :129-136, `griddata(..., 'linear')` :137-138).
//
// All three are HBM-bound streaming kernels; the arithmetic follows numpy / scipy operation by operation so the
// results are bit-identical where the reference's order is defined:
//  * monthly = np.nanmean(daily, 2): numpy's pairwise sum over the (contiguous) day axis, divided by the count;
//  * phole = np.nanmean(monthly[ring]): pairwise sum over the ring cells in raster order, NaN skipped;
//  * griddata 'linear' = scipy LinearNDInterpolator: out = 0; out += c_j * v_j over the 3 vertices of the Delaunay
//    simplex (separate multiply and add), NaN outside the hull.  The triangulation and the barycentric weights are
//    built once on the host (seaiceextentforecasting_b200/ingest.py) and applied here as a 3-nnz-per-row SpMV.
#include "common.cuh"

namespace {

// one thread per cell: v_f = byte/250 for each file, numpy pairwise mean over the files (F <= 128: one leaf)
__global__ void k_nsidc_monthly(const uint8_t* __restrict__ files, int F, size_t file_stride, int header, int C,
                                double* __restrict__ monthly) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const uint8_t* p = files + header + c;
  auto val = [&](int f) { return (double)p[(size_t)f * file_stride] / 250.0; };
  double res;
  if (F == 1) {
    res = val(0);                                   // single monthly file: no mean (:121-127)
  } else {
    if (F < 8) {
      res = 0.0;
      for (int f = 0; f < F; ++f) res = __dadd_rn(res, val(f));
    } else {
      double r[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) r[q] = val(q);
      const int nfull = F & ~7;
      for (int i = 8; i < nfull; i += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], val(i + q));
      }
      res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                      __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
      for (int i = nfull; i < F; ++i) res = __dadd_rn(res, val(i));
    }
    res = res / (double)F;
  }
  if (res > 1.0) res = sie_nan();                   // monthly[monthly>1] = np.nan (:128)
  monthly[c] = res;
}

// single CTA: ordered compaction of the ring cells, numpy pairwise nanmean over them, then the fill
__global__ void __launch_bounds__(256) k_polar_hole(const double* __restrict__ monthly, const double* __restrict__ lat,
                                                    double hole, int C, double* __restrict__ filled,
                                                    double* __restrict__ phole_out, double* __restrict__ ring) {
  __shared__ int s_cnt[8];
  __shared__ int s_base;
  __shared__ double s_phole;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < C; c0 += 256) {
    const int c = c0 + tid;
    const bool in = c < C && lat[c] > hole - 0.5 && lat[c] < hole;
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_cnt[w];
    if (in) ring[off + __popc(bal & ((1u << lane) - 1u))] = monthly[c];
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_cnt[w]; s_base += t; }
    __syncthreads();
  }
  const int n = s_base;
  if (tid < 8) {
    int nanc = 0;
    double sum = 0.0;
    if (n > 0) sum = sie_pw_sum8([&](int i) { return ring[i]; }, n, tid, 0xffu, nanc);
    nanc += __shfl_xor_sync(0xffu, nanc, 1);
    nanc += __shfl_xor_sync(0xffu, nanc, 2);
    nanc += __shfl_xor_sync(0xffu, nanc, 4);
    if (tid == 0) { s_phole = (n - nanc > 0) ? sum / (double)(n - nanc) : sie_nan(); *phole_out = s_phole; }
  }
  __syncthreads();
  const double ph = s_phole;
  for (int c = tid; c < C; c += 256) filled[c] = (lat[c] >= hole - 0.5) ? ph : monthly[c];
}

__global__ void k_regrid(const double* __restrict__ src, int F, int C, const int32_t* __restrict__ vert,
                         const double* __restrict__ bary, int Ct, double* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)F * Ct) return;
  const int f = (int)(idx / Ct), t = (int)(idx - (long long)f * Ct);
  const int v0 = vert[3 * t];
  double out = sie_nan();                           // outside the convex hull (fill_value of griddata)
  if (v0 >= 0) {
    const double* s = src + (size_t)f * C;
    out = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) out = __dadd_rn(out, __dmul_rn(bary[3 * t + k], s[vert[3 * t + k]]));
  }
  dst[idx] = out;
}

}  // namespace

extern "C" int sie_nsidc_monthly(const uint8_t* files, int n_files, size_t file_stride, int header_bytes, int C,
                                 double* monthly, void* stream) {
  SIE_CHECK_ARG(files && monthly, "null pointer");
  SIE_CHECK_ARG(n_files >= 1 && n_files <= 128 && C > 0 && header_bytes >= 0, "1 <= n_files <= 128, C > 0");
  SIE_CHECK_ARG(file_stride >= (size_t)header_bytes + (size_t)C, "file_stride shorter than header + C bytes");
  k_nsidc_monthly<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(files, n_files, file_stride, header_bytes, C,
                                                                     monthly);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}

extern "C" int sie_polar_hole_fill(const double* monthly, const double* lat, double hole, int C, double* filled,
                                   double* phole, double* scratch, void* stream) {
  SIE_CHECK_ARG(monthly && lat && filled && phole && scratch, "null pointer");
  SIE_CHECK_ARG(C > 0, "C > 0");
  k_polar_hole<<<1, 256, 0, (cudaStream_t)stream>>>(monthly, lat, hole, C, filled, phole, scratch);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}

extern "C" int sie_regrid_linear(const double* src, int F, int C, const int32_t* vert, const double* bary, int Ct,
                                 double* dst, void* stream) {
  SIE_CHECK_ARG(src && vert && bary && dst, "null pointer");
  SIE_CHECK_ARG(F > 0 && C > 0 && Ct > 0, "non-positive size");
  const long long total = (long long)F * Ct;
  k_regrid<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, F, C, vert, bary, Ct, dst);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}
