// Shared helpers for libsie_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/sie_b200.h"

void sie_set_error(const char* fmt, ...);

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
template <typename Get>
__device__ __forceinline__ double sie_pw_leaf8(Get get, int lo, int n, int j, unsigned gmask, int& nan_cnt) {
  // All gathers of the leaf (<= 16 per lane + one tail element) are issued before the first add, so a leaf costs
  // one memory round trip; the adds then run in numpy's order.
  const int nfull = (n < 8) ? 0 : n - (n & 7);
  const int ngrp = nfull >> 3;                       // <= 16 full groups of 8
  const int ntail = n - nfull;                       // < 8 (or all of a short list)
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = (k < ngrp) ? get(lo + 8 * k + j) : 0.0;
  double tv = (j < ntail) ? get(lo + nfull + j) : 0.0;
  if (tv != tv) { tv = 0.0; ++nan_cnt; }
  double res;
  if (ngrp == 0) {
    res = 0.0;                                       // n < 8: left-to-right from 0.0
  } else {
    double r = v[0];
    if (r != r) { r = 0.0; ++nan_cnt; }
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      if (k < ngrp) {
        double x = v[k];
        if (x != x) { x = 0.0; ++nan_cnt; }
        r = __dadd_rn(r, x);
      }
    }
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) : xor-butterfly inside the 8-lane group (IEEE add commutes)
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
    res = r;
  }
  for (int t = 0; t < ntail; ++t) res = __dadd_rn(res, __shfl_sync(gmask, tv, t, 8));   // tail, left to right
  return res;
}

// Full pairwise sum of a sequence of length n by one 8-lane group; leaves are visited left to right and
// combined along the recursion tree with an explicit stack (depth <= 24 covers n < 2^31).
template <typename Get>
__device__ __forceinline__ double sie_pw_sum8(Get get, int n, int j, unsigned gmask, int& nan_cnt) {
  if (n <= 128) return sie_pw_leaf8(get, 0, n, j, gmask, nan_cnt);
  // iterative post-order walk of pairwise(lo,n) = pairwise(lo,n2) + pairwise(lo+n2,n-n2)
  int st_lo[24], st_n[24];
  double st_val[24];
  unsigned char st_state[24];
  int sp = 0;
  st_lo[0] = 0; st_n[0] = n; st_state[0] = 0; sp = 1;
  double ret = 0.0;
  while (sp > 0) {
    int t = sp - 1;
    if (st_state[t] == 0) {
      if (st_n[t] <= 128) {
        ret = sie_pw_leaf8(get, st_lo[t], st_n[t], j, gmask, nan_cnt);
        --sp;
      } else {
        int n2 = st_n[t] / 2; n2 -= n2 % 8;
        st_state[t] = 1;
        st_lo[sp] = st_lo[t]; st_n[sp] = n2; st_state[sp] = 0; ++sp;
      }
    } else if (st_state[t] == 1) {
      st_val[t] = ret;
      int n2 = st_n[t] / 2; n2 -= n2 % 8;
      st_state[t] = 2;
      st_lo[sp] = st_lo[t] + n2; st_n[sp] = st_n[t] - n2; st_state[sp] = 0; ++sp;
    } else {
      ret = __dadd_rn(st_val[t], ret);
      --sp;
    }
  }
  return ret;
}

// One THREAD sums a whole sequence in numpy's pairwise order: the 8 accumulators of a leaf live in registers, so a
// warp evaluates 32 independent sequences at once (used where many rows are summed side by side and consecutive
// threads read consecutive addresses).  Bit-identical to sie_pw_sum8 / np.sum.
template <typename Get>
__device__ __forceinline__ double sie_pw_leaf_thread(Get get, int lo, int n, int& nan_cnt) {
  auto val = [&](int i) {
    double x = get(i);
    if (x != x) { x = 0.0; ++nan_cnt; }
    return x;
  };
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, val(lo + i));
    return res;
  }
  double r[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) r[q] = val(lo + q);
  const int nfull = n - (n & 7);
#pragma unroll 2
  for (int i = 8; i < nfull; i += 8) {
    double x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = get(lo + i + q);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double y = x[q];
      if (y != y) { y = 0.0; ++nan_cnt; }
      r[q] = __dadd_rn(r[q], y);
    }
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (int i = nfull; i < n; ++i) res = __dadd_rn(res, val(lo + i));
  return res;
}

template <typename Get>
__device__ __forceinline__ double sie_pw_sum_thread(Get get, int n, int& nan_cnt) {
  if (n <= 128) return sie_pw_leaf_thread(get, 0, n, nan_cnt);
  int st_lo[24], st_n[24];
  double st_val[24];
  unsigned char st_state[24];
  int sp = 1;
  st_lo[0] = 0; st_n[0] = n; st_state[0] = 0;
  double ret = 0.0;
  while (sp > 0) {
    const int t = sp - 1;
    if (st_state[t] == 0) {
      if (st_n[t] <= 128) {
        ret = sie_pw_leaf_thread(get, st_lo[t], st_n[t], nan_cnt);
        --sp;
      } else {
        int n2 = st_n[t] / 2; n2 -= n2 % 8;
        st_state[t] = 1;
        st_lo[sp] = st_lo[t]; st_n[sp] = n2; st_state[sp] = 0; ++sp;
      }
    } else if (st_state[t] == 1) {
      st_val[t] = ret;
      int n2 = st_n[t] / 2; n2 -= n2 % 8;
      st_state[t] = 2;
      st_lo[sp] = st_lo[t] + n2; st_n[sp] = st_n[t] - n2; st_state[sp] = 0; ++sp;
    } else {
      ret = __dadd_rn(st_val[t], ret);
      --sp;
    }
  }
  return ret;
}
