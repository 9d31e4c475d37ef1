// Shared helpers for libsie_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/sie_b200.h"

void sie_set_error(const char* fmt, ...);

// Cached facts of one device + the dynamic-shared-memory limit each kernel was last given there (abi.cu).
#define SIE_MAX_DEVICES 64
#define SIE_ATTR_SLOTS 16
enum SieAttrSlot { SIE_K_DETREND = 0, SIE_K_CORR_TILES, SIE_K_CORR_ROWS_ST, SIE_K_CORR_ROWS_MIR, SIE_K_CORR_ROWS_TAU, SIE_K_AREA_ON32,
                   SIE_K_AREA_OFF32, SIE_K_AREA_ON16, SIE_K_AREA_ON16_2, SIE_K_AREA_ON32_Z, SIE_K_AREA_OFF32_Z, SIE_K_GP,
                   SIE_K_LINKS, SIE_K_CORR_TMA };
struct SieDevice {
  volatile int ready;
  int ordinal, sm_count, max_smem_optin, smem_per_sm;
  size_t l2_bytes;
  long long attr_smem[SIE_ATTR_SLOTS];
};
const SieDevice* sie_device(void);                     // current device, nullptr (error set) if there is none
int sie_ensure_smem(const SieDevice* d, int slot, const void* func, size_t bytes);
int sie_tensor_map_f64_3d(void* out128, const double* base, int B, int ld, int box_cols, int box_rows);   // abi.cu

#define SIE_CHECK_ARG(cond, msg)                 \
  do {                                           \
    if (!(cond)) {                               \
      sie_set_error("%s: %s", __func__, msg);    \
      return SIE_ERR_ARG;                        \
    }                                            \
  } while (0)

#define SIE_CHECK_LAUNCH()                                                   \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      sie_set_error("%s: CUDA error %s", __func__, cudaGetErrorString(e__)); \
      return SIE_ERR_LAUNCH;                                                 \
    }                                                                        \
  } while (0)

__device__ __forceinline__ double sie_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// -------------------------------------------------------------------------------------------------
// numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE), the order in
// which `np.sum` / `np.nanmean` add a contiguous 1-D double array:
//   n < 8     : left-to-right from 0.0
//   n <= 128  : 8 accumulators r[j]=a[j]; r[j]+=a[i+j] for i=8,16,..<n-n%8;
//               ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); then the n%8 tail left-to-right
//   n > 128   : n2=(n/2)-(n/2)%8;  pairwise(a,n2)+pairwise(a+n2,n-n2)
// The area-growth and merge decisions of ComplexNetworks.py:113/:250/:252 are float comparisons of such
// sums, so the device evaluates them in exactly this order.
//
// `sie_pw_leaf8`: an aligned group of 8 lanes (lane j = accumulator j) sums one leaf (n <= 128) of a
// sequence whose i-th element is get(i); NaN entries count as 0 and are tallied in `nan_cnt` (nanmean; the
// per-lane tallies must be summed over the group by the caller).  All 8 lanes return the same value.
// -------------------------------------------------------------------------------------------------
// Compiler fence over a register array: everything that defines v[] (the gathers) is ordered before, everything
// that consumes it (the adds) after, so a leaf's gathers are all in flight before the first add instead of being
// interleaved load-use-load-use (which serialises the memory round trips).
#ifndef SIE_PW_SMALL_NQ
#define SIE_PW_SMALL_NQ 8       // unrolled leaf variants: <= 8 * SIE_PW_SMALL_NQ elements and <= 128 (0: only the latter); 8 measured best of 4..12
#endif
template <int NQ> __device__ __forceinline__ void sie_fence_regs(double (&v)[NQ]);
template <> __device__ __forceinline__ void sie_fence_regs<1>(double (&v)[1]) { asm volatile("" : "+d"(v[0])); }
template <> __device__ __forceinline__ void sie_fence_regs<4>(double (&v)[4]) {
  asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]));
}
template <> __device__ __forceinline__ void sie_fence_regs<8>(double (&v)[8]) {
  asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]), "+d"(v[5]), "+d"(v[6]), "+d"(v[7]));
}
template <> __device__ __forceinline__ void sie_fence_regs<6>(double (&v)[6]) {
  asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]), "+d"(v[5]));
}
template <> __device__ __forceinline__ void sie_fence_regs<10>(double (&v)[10]) {
  asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]), "+d"(v[5]), "+d"(v[6]), "+d"(v[7]),
                    "+d"(v[8]), "+d"(v[9]));
}
template <> __device__ __forceinline__ void sie_fence_regs<12>(double (&v)[12]) {
  asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]), "+d"(v[5]), "+d"(v[6]), "+d"(v[7]),
                    "+d"(v[8]), "+d"(v[9]), "+d"(v[10]), "+d"(v[11]));
}
template <> __device__ __forceinline__ void sie_fence_regs<16>(double (&v)[16]) {
  asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]), "+d"(v[5]), "+d"(v[6]), "+d"(v[7]),
                    "+d"(v[8]), "+d"(v[9]), "+d"(v[10]), "+d"(v[11]), "+d"(v[12]), "+d"(v[13]), "+d"(v[14]), "+d"(v[15]));
}

// res + tail[0] + tail[1] + ... (left to right), tail element t held by lane t of the 8-lane group: the <= 7 shuffles are
// issued back to back, then the adds run as one dependent chain (a runtime loop would pay shuffle latency per step).
__device__ __forceinline__ double sie_add_tail8(double res, double tv, int nt, unsigned gmask) {
  double t[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) t[i] = __shfl_sync(gmask, tv, i, 8);
#pragma unroll
  for (int i = 0; i < 7; ++i) if (i < nt) res = __dadd_rn(res, t[i]);
  return res;
}

// Lane j's part of one leaf: accumulator j over the `ngrp` full groups of 8 (sequential, numpy's order) and tail
// element j.  The NQ + 1 gathers are unconditional (indices clamped to the last element; the extras are masked
// out of the sums) so they are issued back to back.  get(i) must be valid for lo <= i < lo + n.
template <int NQ, typename Get>
__device__ __forceinline__ void sie_pw_lane8(Get get, int lo, int n, int ngrp, int ntail, int j, double& acc, double& tv,
                                             int& nan_cnt) {
  double v[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) v[k] = get(lo + min(8 * k + j, n - 1));
  double t[1];
  t[0] = get(lo + min(8 * ngrp + j, n - 1));
  sie_fence_regs<NQ>(v);
  sie_fence_regs<1>(t);
  tv = t[0];
  if (j >= ntail) tv = 0.0;
  if (tv != tv) { tv = 0.0; ++nan_cnt; }
  double r = 0.0;
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    if (k < ngrp) {
      double x = v[k];
      if (x != x) { x = 0.0; ++nan_cnt; }
      r = (k == 0) ? x : __dadd_rn(r, x);
    }
  }
  acc = r;
}
template <typename Get>
__device__ __forceinline__ void sie_pw_lane8_any(Get get, int lo, int n, int ngrp, int ntail, int j, double& acc,
                                                 double& tv, int& nan_cnt) {
  // two unrolled variants only (<= 32 elements, <= 128): every call site inlines them, and the domain-growth kernel's
  // instruction footprint matters (two CTAs in different phases share an SM's instruction cache)
#if SIE_PW_SMALL_NQ > 0
  if (ngrp <= SIE_PW_SMALL_NQ) sie_pw_lane8<SIE_PW_SMALL_NQ>(get, lo, n, ngrp, ntail, j, acc, tv, nan_cnt);
  else
#endif
    sie_pw_lane8<16>(get, lo, n, ngrp, ntail, j, acc, tv, nan_cnt);
}

template <typename Get>
__device__ __forceinline__ double sie_pw_leaf8(Get get, int lo, int n, int j, unsigned gmask, int& nan_cnt) {
  // All gathers of the leaf (<= 16 per lane + one tail element) are issued before the first add, so a leaf costs
  // one memory round trip; the adds then run in numpy's order.
  const int nfull = (n < 8) ? 0 : n - (n & 7);
  const int ngrp = nfull >> 3;                       // <= 16 full groups of 8
  const int ntail = n - nfull;                       // < 8 (or all of a short list)
  double r, tv;
  sie_pw_lane8_any(get, lo, n, ngrp, ntail, j, r, tv, nan_cnt);
  double res = 0.0;                                  // n < 8: left-to-right from 0.0
  if (ngrp > 0) {
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) : xor-butterfly inside the 8-lane group (IEEE add commutes)
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
    res = r;
  }
  res = sie_add_tail8(res, tv, ntail, gmask);        // tail, left to right
  return res;
}

// One leaf (n <= 128) whose elements are CONTIGUOUS in memory: q = address of element `lo + j` (lane j).  Loads use
// immediate offsets (no index arithmetic), NaN is not screened: a NaN element makes the result NaN and the caller
// redoes the leaf with the screening path.  Reads up to 135 elements past q[0] (masked out of the sum): the
// caller guarantees that memory is readable.
template <int NQ>
__device__ __forceinline__ double sie_pw_leaf8_contig_n(const double* q, int ngrp, int nt, int j, unsigned gmask) {
  double v[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) v[k] = q[8 * k];
  double t[1];
  t[0] = q[8 * ngrp];
  sie_fence_regs<NQ>(v);
  sie_fence_regs<1>(t);
  double r = v[0];
#pragma unroll
  for (int k = 1; k < NQ; ++k) if (k < ngrp) r = __dadd_rn(r, v[k]);
  double res = 0.0;
  if (ngrp > 0) {
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
    res = r;
  }
  const double tv = (j < nt) ? t[0] : 0.0;
  res = sie_add_tail8(res, tv, nt, gmask);
  return res;
}
__device__ __forceinline__ double sie_pw_leaf8_contig(const double* q, int n, int j, unsigned gmask) {
  const int ngrp = (n < 8) ? 0 : (n >> 3), nt = n - 8 * ngrp;
#if SIE_PW_SMALL_NQ > 0
  if (ngrp <= SIE_PW_SMALL_NQ) return sie_pw_leaf8_contig_n<SIE_PW_SMALL_NQ>(q, ngrp, nt, j, gmask);
#endif
  return sie_pw_leaf8_contig_n<16>(q, ngrp, nt, j, gmask);
}

// One leaf (n <= 128) that STRADDLES two contiguous runs: elements below the switch come from run a, the rest from run b.
// qa = address of this lane's element of the leaf if it lay in run a, qb = the same for run b; ks = number of this lane's
// full-group elements (k = 0, 1, ...) that lie in run a.  Per element one pointer select + an immediate-offset load
// (the generic getter path costs ~4x the instructions); no NaN screening, the caller redoes a NaN result.  Reads up to
// 135 elements past the end of run b (masked out of the sum).
template <int NQ>
__device__ __forceinline__ double sie_pw_leaf8_two_n(const double* qa, const double* qb, int ks, int ngrp, int nt, int j,
                                                     unsigned gmask) {
  double v[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) v[k] = ((k < ks) ? qa : qb)[8 * k];
  double t[1];
  t[0] = ((ngrp < ks) ? qa : qb)[8 * ngrp];
  sie_fence_regs<NQ>(v);
  sie_fence_regs<1>(t);
  double r = v[0];
#pragma unroll
  for (int k = 1; k < NQ; ++k) if (k < ngrp) r = __dadd_rn(r, v[k]);
  double res = 0.0;
  if (ngrp > 0) {
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
    res = r;
  }
  const double tv = (j < nt) ? t[0] : 0.0;
  res = sie_add_tail8(res, tv, nt, gmask);
  return res;
}
__device__ __forceinline__ double sie_pw_leaf8_two(const double* qa, const double* qb, int ks, int n, int j, unsigned gmask) {
  const int ngrp = (n < 8) ? 0 : (n >> 3), nt = n - 8 * ngrp;
#if SIE_PW_SMALL_NQ > 0
  if (ngrp <= SIE_PW_SMALL_NQ) return sie_pw_leaf8_two_n<SIE_PW_SMALL_NQ>(qa, qb, ks, ngrp, nt, j, gmask);
#endif
  return sie_pw_leaf8_two_n<16>(qa, qb, ks, ngrp, nt, j, gmask);
}

// numpy's pairwise recursion pairwise(lo,n) = pairwise(lo,n2) + pairwise(lo+n2,n-n2), n2 = n/2 - (n/2)%8, over leaves of
// <= 128 elements, walked left to right without a memory stack: the path from the root is a bit mask (bit d = "right
// child at depth d+1"), node bounds are recomputed from the root (<= 16 integer steps), and the pending left-sibling
// sums live in registers.  leaf(lo, len) returns the leaf's sum (uniform over the 8-lane group).  n < 2^22.
// MAXD = tree depth supported: n <= 128 * 2^MAXD.
template <int MAXD, typename Leaf>
__device__ __forceinline__ double sie_pw_tree(Leaf leaf, int n) {
  // (n <= 128 runs the loop once with depth 0: ONE inlined copy of the leaf per call site keeps the code small)
  double val[MAXD];
  unsigned path = 0u;
  int depth = 0, lo = 0, len = n;
  while (len > 128) { int n2 = len / 2; n2 -= n2 % 8; len = n2; ++depth; }
  double ret;
  while (true) {
    ret = leaf(lo, len);
    while (depth > 0 && ((path >> (depth - 1)) & 1u)) {          // finished a right child: add the left sibling
      double left = 0.0;
#pragma unroll
      for (int d = 0; d < MAXD; ++d) if (d == depth - 1) left = val[d];
      ret = __dadd_rn(left, ret);
      path &= ~(1u << (depth - 1));
      --depth;
    }
    if (depth == 0) break;
#pragma unroll
    for (int d = 0; d < MAXD; ++d) if (d == depth - 1) val[d] = ret;   // finished a left child
    path |= 1u << (depth - 1);
    lo = 0; len = n;
    for (int d = 0; d < depth; ++d) {
      int n2 = len / 2; n2 -= n2 % 8;
      if ((path >> d) & 1u) { lo += n2; len -= n2; } else { len = n2; }
    }
    while (len > 128) { int n2 = len / 2; n2 -= n2 % 8; len = n2; ++depth; }
  }
  return ret;
}

// Full pairwise sum of a sequence of length n by one 8-lane group (generic element getter).
template <int MAXD = 16, typename Get>
__device__ __forceinline__ double sie_pw_sum8(Get get, int n, int j, unsigned gmask, int& nan_cnt) {
  return sie_pw_tree<MAXD>([&](int lo, int len) { return sie_pw_leaf8(get, lo, len, j, gmask, nan_cnt); }, n);
}
