// K4 + K5: tau-thresholded domain growth (step 1) and largest-first merging (step 2).
// Reference: Network.area_level, ComplexNetworks.py:49-278 (semantics restated in SURVEY.md App. A).
//
// The algorithm is sequential per network (every decision consumes cells later decisions could have used),
// so one persistent CTA owns one network and B networks run side by side.  Inside a step the work is
// parallel: every frontier cell's mean correlation with the area (step 1) and every row of the
// hypothetical merged area (step 2) is a numpy-order pairwise sum evaluated by an 8-lane group (lane j =
// numpy's accumulator j), so the float comparisons `mean > tau` / first-max pick the same cell as the
// reference when fed the same R.  Candidate order (direction-major, then position in the area list,
// duplicates kept) only matters for ties; it is carried as a per-cell key instead of materialising the
// reference's candidate list.  Integer state lives in shared memory (global scratch for big grids).
//
// Latency engineering (the kernel is bound by its chain of dependent steps, not by bandwidth):
//  * step 1 keeps, per frontier cell, numpy's 8 accumulators + the <8-element tail of its correlation list in
//    shared memory.  While the area has <= 128 cells (one pairwise leaf) adding a cell appends ONE correlation per
//    frontier cell (one gather, one L2 round trip per growth step) and the mean is re-assembled from the
//    accumulators in numpy's order, bit-identical to summing the whole list again.  Larger areas fall back to
//    re-summing gathered lists.
//  * step 2 keeps a dense list-order copy D[p][q] = R[best_p][best_q] of the current largest area, extended when
//    a neighbour is merged, so the O(n^2) "mean of row means" statistic of every hypothetical merge streams
//    contiguous rows instead of re-gathering the best x best block for every candidate and round.
// No roofline fraction is claimed for this kernel; see DESIGN.md.
#include "common.cuh"

namespace {

constexpr int NT = 512;
constexpr int NG = NT / 8;          // 8-lane groups per CTA
constexpr int FCAP = NT;           // frontier slots with incremental state: one thread per slot
constexpr int DCAP_MAX = 1024;     // rows/cols of the dense best-area sub-matrix
constexpr int MAXCH = 32;          // neighbour areas evaluated per chunk of a merge round
constexpr uint32_t NOKEY = 0xffffffffu;
constexpr unsigned long long NOKEY64 = ~0ull;

struct AreaScratch {   // per-job global scratch (byte offsets computed on host and device the same way)
  int32_t* s1_cells;   // [C] step-1 member cells, area after area
  int32_t* ibuf;       // [6*C] fallback for the shared-memory integer arrays
  int32_t* nbuf;       // [C] neighbour-area node lists (step 2)
  double* rowmean;     // [RM] row means of hypothetical areas
  double* dmat;        // [dcap*dcap] dense list-order copy of R restricted to the current best area (step 2)
  int32_t* a_start;    // [MA] per step-1 area (segment)
  int32_t* a_len;      // [MA]
  int32_t* seg_next;   // [MA]
  int32_t* tail;       // [MA]
  int32_t* size;       // [MA] cells in the (merged) area headed by this key; 0 when absorbed
  int32_t* fin;        // [MA]
  int32_t* nlist;      // [MA] neighbour keys discovered this round
  int32_t* noff;       // [MA+1] offsets of neighbour node lists in nbuf
  unsigned long long* okey;  // [MA]
  double* stat;        // [MA]
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline size_t rm_cap(int C) { return (size_t)4 * C + 1024; }
__host__ __device__ inline int d_cap(int C) { return C < DCAP_MAX ? C : DCAP_MAX; }
__host__ __device__ inline size_t scratch_per_job(int C, int MA) {
  size_t s = 0;
  s += align_up(sizeof(int32_t) * (size_t)C, 256);          // s1_cells
  s += align_up(sizeof(int32_t) * (size_t)6 * C, 256);      // ibuf
  s += align_up(sizeof(int32_t) * (size_t)C, 256);          // nbuf
  s += align_up(sizeof(double) * rm_cap(C), 256);           // rowmean
  s += align_up(sizeof(double) * (size_t)d_cap(C) * d_cap(C), 256);   // dmat
  s += 8 * align_up(sizeof(int32_t) * (size_t)(MA + 1), 256);
  s += 2 * align_up(sizeof(double) * (size_t)MA, 256);
  return s;
}
__device__ inline AreaScratch carve(unsigned char* p, int C, int MA) {
  AreaScratch s;
  auto take = [&](size_t bytes) { unsigned char* r = p; p += align_up(bytes, 256); return r; };
  s.s1_cells = (int32_t*)take(sizeof(int32_t) * (size_t)C);
  s.ibuf = (int32_t*)take(sizeof(int32_t) * (size_t)6 * C);
  s.nbuf = (int32_t*)take(sizeof(int32_t) * (size_t)C);
  s.rowmean = (double*)take(sizeof(double) * rm_cap(C));
  s.dmat = (double*)take(sizeof(double) * (size_t)d_cap(C) * d_cap(C));
  s.a_start = (int32_t*)take(sizeof(int32_t) * (size_t)(MA + 1));
  s.a_len = (int32_t*)take(sizeof(int32_t) * (size_t)(MA + 1));
  s.seg_next = (int32_t*)take(sizeof(int32_t) * (size_t)(MA + 1));
  s.tail = (int32_t*)take(sizeof(int32_t) * (size_t)(MA + 1));
  s.size = (int32_t*)take(sizeof(int32_t) * (size_t)(MA + 1));
  s.fin = (int32_t*)take(sizeof(int32_t) * (size_t)(MA + 1));
  s.nlist = (int32_t*)take(sizeof(int32_t) * (size_t)(MA + 1));
  s.noff = (int32_t*)take(sizeof(int32_t) * (size_t)(MA + 1));
  s.okey = (unsigned long long*)take(sizeof(double) * (size_t)MA);
  s.stat = (double*)take(sizeof(double) * (size_t)MA);
  return s;
}

// Argmax record, compared branch-free as a 128-bit unsigned number: `hi` is an order-preserving image of the
// mean (0 = no candidate), `lo` the tie-break priority (larger wins: the bitwise complement of the reference's
// candidate-list position, so the earliest candidate wins ties) with the candidate index in its low bits.
struct Pick {
  unsigned long long hi, lo;
};
__device__ __forceinline__ unsigned long long ord_of(double x) {      // monotone map double -> u64, never 0
  x = __dadd_rn(x, 0.0);                                               // -0.0 -> +0.0 (they compare equal)
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_to_double(unsigned long long o) {
  const unsigned long long u = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ Pick pick_max(const Pick& a, const Pick& b) {
  const bool tb = (b.hi > a.hi) || (b.hi == a.hi && b.lo > a.lo);
  Pick r;
  r.hi = tb ? b.hi : a.hi;
  r.lo = tb ? b.lo : a.lo;
  return r;
}
__device__ __forceinline__ Pick pick_none() { Pick r; r.hi = 0ull; r.lo = 0ull; return r; }
// CTA-wide argmax; every thread returns the same winner.  `slots` is shared scratch of NT/32 records.
// `SyncBefore = false` is for callers that already placed a barrier between the previous call's reads and this one.
template <bool SyncBefore = true>
__device__ __forceinline__ Pick block_pick(Pick v, Pick* slots) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Pick w;
    w.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
    w.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
    v = pick_max(v, w);
  }
  if (SyncBefore) __syncthreads();     // slots may still be read from the previous call
  if ((threadIdx.x & 31) == 0) slots[threadIdx.x >> 5] = v;
  __syncthreads();
  Pick r = slots[0];
#pragma unroll
  for (int w = 1; w < NT / 32; ++w) r = pick_max(r, slots[w]);
  return r;
}

__global__ void __launch_bounds__(NT, 1)
k_area_level(const double* __restrict__ Rall, const double* __restrict__ stencil_all,
             const int32_t* __restrict__ node_cell_all, const int32_t* __restrict__ cell_node_all,
             const int32_t* __restrict__ n_nodes, const double* __restrict__ tau_all,
             const int32_t* __restrict__ first_nan_cell, int X, int Y, int ldn, int latlon, int MA,
             int32_t* __restrict__ area_cells_all, int32_t* __restrict__ area_start_all,
             int32_t* __restrict__ area_key_all, int32_t* __restrict__ n_areas_all,
             int32_t* __restrict__ label_all, int32_t* __restrict__ status_all,
             unsigned char* __restrict__ scratch_all, size_t scratch_stride, int use_smem,
             unsigned long long* __restrict__ work_all) {
  extern __shared__ __align__(16) int32_t smem_i[];
  __shared__ Pick slots[NT / 32];
  __shared__ int sh_i[8];
  __shared__ int sh_koff[MAXCH + 1], sh_uoff[MAXCH + 1];
  __shared__ unsigned long long sh_work;
  __shared__ unsigned long long ph[12];   // per-phase SM cycles (thread 0's view), reported through work[4..15]
  long long tick_last = 0;
#define TICK(i)                                                        \
  do {                                                                 \
    if (tid == 0) {                                                    \
      const long long t__ = clock64();                                 \
      ph[i] += (unsigned long long)(t__ - tick_last);                  \
      tick_last = t__;                                                 \
    }                                                                  \
  } while (0)
  unsigned long long wk = 0;   // correlations consumed (algorithmic gathers), tallied by lane 0 of each group

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid >> 3, j = tid & 7;
  const unsigned gmask = 0xffu << (lane & 24);
  const int C = X * Y;
  const int N = min(n_nodes[b], ldn);
  const double tau = tau_all[b];
  const double* R = Rall + (size_t)b * ldn * ldn;
  const double* sten = stencil_all + (size_t)b * ldn * 4;
  const int32_t* cnode_g = cell_node_all + (size_t)b * C;
  int32_t* out_cells = area_cells_all + (size_t)b * C;
  int32_t* out_start = area_start_all + (size_t)b * (MA + 1);
  int32_t* out_key = area_key_all + (size_t)b * MA;
  int32_t* out_label = label_all + (size_t)b * C;

  AreaScratch S = carve(scratch_all + (size_t)b * scratch_stride, C, MA);
  // dynamic shared memory: [frontier-slot state: facc 8*FCAP f64 | ftail 8*FCAP f64 | fnan FCAP i32][6*C i32]
  double* facc = reinterpret_cast<double*>(smem_i);     // [8][FCAP] numpy accumulator j of slot s at j*FCAP+s
  double* ftail = facc + 8 * FCAP;                       // [8][FCAP] tail element t of slot s at t*FCAP+s
  int32_t* fnan = reinterpret_cast<int32_t*>(ftail + 8 * FCAP);   // [FCAP] NaN entries seen by slot s
  int32_t* ib = use_smem ? fnan + FCAP : S.ibuf;
  int32_t* lab = ib;                      // [C] area key of each cell, -1 = unassigned
  uint32_t* fkey = (uint32_t*)(ib + C);   // [C] frontier key (step 1)
  int32_t* flist = ib + 2 * C;            // [C] frontier cells (step 1)
  int32_t* hn = ib + 3 * C;               // [C] node list of the current area (step 1) / best area (step 2)
  int32_t* hc = ib + 4 * C;               // [C] cell list of the best area (step 2)
  int32_t* cnl = ib + 5 * C;              // [C] local copy of cell -> node
  int32_t* knl = (use_smem >= 2) ? ib + 6 * C : S.nbuf;   // [C] node lists of the neighbour areas (step 2)
  const int32_t* cnode = cnl;
  const int dcap = d_cap(C);
  double* D = S.dmat;

  for (int c = tid; c < C; c += NT) { lab[c] = -1; fkey[c] = NOKEY; out_label[c] = -1; cnl[c] = cnode_g[c]; }
  if (tid < 12) ph[tid] = 0ull;
  if (tid == 0) { n_areas_all[b] = 0; out_start[0] = 0; sh_work = 0ull; if (work_all) work_all[SIE_AREA_WORK * b] = 0ull; }
  __syncthreads();
  if (first_nan_cell[b] < 0) {            // :50-51 IndexError in the reference
    if (tid == 0) status_all[b] = SIE_JOB_NO_NAN_CELL;
    return;
  }
  if (status_all[b] == SIE_JOB_CAPACITY) return;   // K1 already flagged this job

  // ---- step-1 helpers --------------------------------------------------------------------------------
  // gen_area_neighbours :80-94 for one member cell `cc` at list position `p` (no lat-lon wrap here): warp 0,
  // lane d < 4 = direction.  New frontier cells are appended at flist[cnt..]; returns the new count.
  auto frontier_add = [&](int cc, int p, int cnt) -> int {
    bool isnew = false;
    int f = -1;
    if (lane < 4) {
      const int d = lane;
      const int ci = cc / Y, cj = cc - ci * Y;
      const int a = ci + (d == 0 ? -1 : (d == 1 ? 1 : 0));
      const int q = cj + (d == 2 ? -1 : (d == 3 ? 1 : 0));
      if (a >= 0 && a < X && q >= 0 && q < Y) {
        f = a * Y + q;
        const int fn = cnode[f];
        if (lab[f] >= 0 || fn < 0 || fn >= N) f = -1;
      }
      if (f >= 0) {
        const uint32_t key = ((uint32_t)d << 28) | (uint32_t)p;
        const uint32_t old = fkey[f];
        isnew = (old == NOKEY);
        if (key < old) fkey[f] = key;
      }
    }
    const unsigned mnew = __ballot_sync(0xffffffffu, isnew);
    if (isnew) flist[cnt + __popc(mnew & ((1u << lane) - 1u))] = f;
    __syncwarp();
    return cnt + __popc(mnew);
  };
  // numpy pairwise state of slot s (its frontier cell vs the first n <= 128 member cells): 8-lane group, lane j
  // builds accumulator j = a[j] + a[8+j] + ... over the full groups of 8 and fetches tail element j.
  auto init_slot = [&](int s, int n) {
    const double* row = R + (size_t)cnode[flist[s]] * ldn;
    const int ngrp = n >> 3, nt = n & 7;
    double v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = (q < ngrp) ? __ldg(row + hn[8 * q + j]) : 0.0;
    double tv = (j < nt) ? __ldg(row + hn[8 * ngrp + j]) : 0.0;
    int nanc = 0;
    if (tv != tv) { tv = 0.0; ++nanc; }
    double r = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      if (q < ngrp) {
        double x = v[q];
        if (x != x) { x = 0.0; ++nanc; }
        r = (q == 0) ? x : __dadd_rn(r, x);
      }
    }
    facc[j * FCAP + s] = r;
    ftail[j * FCAP + s] = tv;
    nanc += __shfl_xor_sync(gmask, nanc, 1);
    nanc += __shfl_xor_sync(gmask, nanc, 2);
    nanc += __shfl_xor_sync(gmask, nanc, 4);
    if (j == 0) fnan[s] = nanc;
  };
  // append element index n_old (value v) to slot s
  auto fold_slot = [&](int s, double v, int n_old) {
    if (v != v) { v = 0.0; fnan[s] += 1; }
    const int t = n_old & 7;
    if (t == 7) {                         // a group of 8 is complete: it joins the accumulators
      if (n_old == 7) {
#pragma unroll
        for (int q = 0; q < 7; ++q) facc[q * FCAP + s] = ftail[q * FCAP + s];
        facc[7 * FCAP + s] = v;
      } else {
#pragma unroll
        for (int q = 0; q < 7; ++q) facc[q * FCAP + s] = __dadd_rn(facc[q * FCAP + s], ftail[q * FCAP + s]);
        facc[7 * FCAP + s] = __dadd_rn(facc[7 * FCAP + s], v);
      }
    } else {
      ftail[t * FCAP + s] = v;
    }
  };
  // np.nanmean of slot s's list of n <= 128 correlations, in numpy's pairwise order
  auto eval_slot = [&](int s, int n) -> double {
    const int nt = n & 7;
    double res = 0.0;
    if (n >= 8) {
      const double a0 = facc[s], a1 = facc[FCAP + s], a2 = facc[2 * FCAP + s], a3 = facc[3 * FCAP + s];
      const double a4 = facc[4 * FCAP + s], a5 = facc[5 * FCAP + s], a6 = facc[6 * FCAP + s], a7 = facc[7 * FCAP + s];
      res = __dadd_rn(__dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3)), __dadd_rn(__dadd_rn(a4, a5), __dadd_rn(a6, a7)));
    }
    for (int t = 0; t < nt; ++t) res = __dadd_rn(res, ftail[t * FCAP + s]);
    return res / (double)(n - fnan[s]);
  };

  // =============================================================== step 1 (:154-196)
  const long long clk0 = clock64();
  tick_last = clk0;
  unsigned long long n_steps = 0, n_rounds = 0, n_slow = 0;
  int nA = 0;        // areas created so far (uniform across the CTA)
  int base = 0;      // cells assigned so far
  bool overflow = false;
  int c0 = 0;
  while (c0 < C && !overflow) {
    // --- find the first cell >= c0 (raster order) that seeds an area
    const int c = c0 + tid;
    int dir = -1;
    if (c < C) {
      const int n = cnode[c];
      if (n >= 0 && n < N && lab[c] < 0) {
        const int ci = c / Y, cj = c - ci * Y;
        double mx = -INFINITY;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int a = ci + (d == 0 ? -1 : (d == 1 ? 1 : 0));
          const int q = cj + (d == 2 ? -1 : (d == 3 ? 1 : 0));
          double v = sten[n * 4 + d];     // NaN when off-grid / not a node; lat-lon wrap already applied
          // gen_cell_neighbours :53-78: an in-bounds neighbour that is already taken becomes the NaN
          // sentinel; the wrapped neighbour is not checked against `unavail`
          if (a >= 0 && a < X && q >= 0 && q < Y && lab[a * Y + q] >= 0) v = sie_nan();
          if (v == v && v > mx) { mx = v; dir = d; }   // strict '>' keeps the first maximum (:175-181)
        }
        if (!(mx > tau)) dir = -1;
      }
    }
    Pick cand = pick_none();
    if (dir >= 0) { cand.hi = 1ull; cand.lo = ~(unsigned long long)(unsigned)c; }   // earliest cell wins
    const Pick win = block_pick(cand, slots);
    TICK(0);                              // seed search
    if (win.hi == 0ull) { c0 += NT; continue; }
    const int seed = (int)(unsigned)(~win.lo);
    if (tid == seed - c0) sh_i[0] = dir;
    __syncthreads();
    const int sd = sh_i[0];
    const int si = seed / Y, sj = seed - si * Y;
    int ni = si + (sd == 0 ? -1 : (sd == 1 ? 1 : 0));
    int nj = sj + (sd == 2 ? -1 : (sd == 3 ? 1 : 0));
    if (nj < 0) nj = Y - 1;               // only reachable with latlon (stencil is NaN otherwise)
    if (nj >= Y) nj = 0;
    const int nbr = ni * Y + nj;
    c0 = seed + 1;
    if (lab[nbr] >= 0) continue;          // :182 wrapped neighbour already taken -> no area, move on
    if (nA >= MA) { overflow = true; break; }

    // --- new area k = nA : [seed, nbr], then expand (:120-152)
    const int k = nA;
    int n = 2, nf = 0;
    __syncthreads();
    if (tid < 32) {                       // warp 0 owns the integer bookkeeping
      if (lane == 0) {
        lab[seed] = k; lab[nbr] = k;
        hn[0] = cnode[seed]; hn[1] = cnode[nbr];
        S.s1_cells[base] = seed; S.s1_cells[base + 1] = nbr;
      }
      __syncwarp();
      int cnt = frontier_add(seed, 0, 0);
      cnt = frontier_add(nbr, 1, cnt);
      if (lane == 0) sh_i[1] = cnt;
    }
    __syncthreads();
    nf = sh_i[1];
    bool fast = (nf <= FCAP);             // incremental accumulators valid (uniform across the CTA)
    if (fast) {
      for (int s = g; s < nf; s += NG) init_slot(s, n);
      __syncthreads();
    }
    TICK(3);                              // area creation + first slot states
    while (nf > 0) {
      Pick loc = pick_none();
      if (!fast) ++n_slow;
      if (fast) {
        TICK(11);
        if (tid < nf) {
          const double mean = eval_slot(tid, n);
          if (mean == mean) { loc.hi = ord_of(mean); loc.lo = ((unsigned long long)(~fkey[flist[tid]]) << 32) | (unsigned)tid; }
        }
        if (tid == 0) wk += (unsigned long long)n * (unsigned long long)nf;
      } else {
        for (int q = g; q < nf; q += NG) {
          const int f = flist[q];
          const double* row = R + (size_t)cnode[f] * ldn;
          int nanc = 0;
          const double sum = sie_pw_sum8([&](int i) { return __ldg(row + hn[i]); }, n, j, gmask, nanc);
          if (j == 0) wk += (unsigned long long)n;
          nanc += __shfl_xor_sync(gmask, nanc, 1);
          nanc += __shfl_xor_sync(gmask, nanc, 2);
          nanc += __shfl_xor_sync(gmask, nanc, 4);
          const double mean = sum / (double)(n - nanc);   // np.nanmean: NaN -> 0, divide by the non-NaN count
          if (mean == mean) {
            Pick cur; cur.hi = ord_of(mean); cur.lo = ((unsigned long long)(~fkey[f]) << 32) | (unsigned)q;
            loc = pick_max(loc, cur);
          }
        }
      }
      TICK(10);                           // evaluate
      const Pick win2 = block_pick<false>(loc, slots);   // the barrier closing the previous step already ran
      TICK(1);                            // argmax
      if (win2.hi == 0ull || !(ord_to_double(win2.hi) > tau)) break;   // :134 (nanmax of all-NaN is NaN -> stop)
      const int widx = (int)(unsigned)(win2.lo & 0xffffffffull);
      const int last = nf - 1;
      if (tid < 32) {
        const int m = flist[widx];
        __syncwarp();
        if (lane == 0) {
          lab[m] = k;
          hn[n] = cnode[m];
          S.s1_cells[base + n] = m;
          fkey[m] = NOKEY;
          flist[widx] = flist[last];
          sh_i[4] = m;
        }
        __syncwarp();
        const int cnt = frontier_add(m, n, last);
        if (lane == 0) sh_i[1] = cnt;
      } else if (tid < 64 && fast && widx != last) {   // slot state follows the cell moved into the hole
        const int l2 = lane;
        if (l2 < 8) facc[l2 * FCAP + widx] = facc[l2 * FCAP + last];
        else if (l2 < 16) ftail[(l2 - 8) * FCAP + widx] = ftail[(l2 - 8) * FCAP + last];
        else if (l2 == 16) fnan[widx] = fnan[last];
      }
      __syncthreads();
      TICK(2);                            // frontier update
      nf = sh_i[1];
      if (fast) {
        if (n + 1 > 128 || nf > FCAP) {
          fast = false;                   // beyond one pairwise leaf: re-sum gathered lists from here on
        } else {
          const int mnode = hn[n];
          // new frontier cells (slots last..nf-1): build their state over the n+1 member cells; use the highest
          // groups, which rarely own an old slot as well
          for (int s = last + (NG - 1 - g); s < nf; s += NG) init_slot(s, n + 1);
          // old frontier cells: append the correlation with the new member
          if (tid < last) fold_slot(tid, __ldg(R + (size_t)cnode[flist[tid]] * ldn + mnode), n);
        }
      }
      ++n;
      ++n_steps;
      __syncthreads();
      TICK(3);                            // one gather per frontier cell + new-slot states
    }
    __syncthreads();
    for (int q = tid; q < nf; q += NT) fkey[flist[q]] = NOKEY;
    if (tid == 0) {
      S.a_start[k] = base; S.a_len[k] = n; S.seg_next[k] = -1; S.tail[k] = k; S.size[k] = n; S.fin[k] = 0;
      S.okey[k] = NOKEY64;
    }
    base += n;
    nA = k + 1;
    __syncthreads();
  }
  if (overflow) {
    if (tid == 0) status_all[b] = SIE_JOB_CAPACITY;
    return;
  }

  // =============================================================== step 2 (:200-265)
  const long long clk1 = clock64();
  tick_last = clk1;
  // `taken` is now "belongs to a finalised area"; lab[] keeps tracking the current owner key.
  int cur_best = -1, nb = 0;     // best area whose lists are materialised in hn/hc
  const size_t RM = rm_cap(C);
  bool d_ok = false;             // dense block valid for the first nb member cells
  // Dense list-order copy of R restricted to the best area, stored by diagonals: element (p, q), q > p, lives at
  // D[(q-p-1)*dcap + p], so thread p walking its row p+1, p+2, ... reads addresses consecutive with its
  // neighbours' (coalesced) in the row-mean pass.  extend_D adds the pairs with from <= q < to.
  auto extend_D = [&](int from, int to) {
    d_ok = (to <= dcap) && (from == 0 || d_ok);
    if (!d_ok) return;
    for (int i0 = 0; i0 < to - 1; i0 += 8) {            // 8 diagonals per pass: 8 gathers in flight per thread
      const int plo = max(0, from - 8 - i0);
      for (int pp = plo + tid; pp < to - 1 - i0; pp += NT) {
        const double* row = R + (size_t)hn[pp] * ldn;
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = pp + 1 + i0 + u;
          v[u] = (q < to && q >= from) ? __ldg(row + hn[q]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = pp + 1 + i0 + u;
          if (q < to && q >= from) D[(size_t)(i0 + u) * dcap + pp] = v[u];
        }
      }
    }
    __syncthreads();
  };
  while (true) {
    // --- largest not-yet-final area, first key on ties (:207-212)
    Pick loc = pick_none();
    for (int k = tid; k < nA; k += NT) {
      const int sz = S.size[k];
      if (sz > 0) {             // still a key of V
        Pick cur; cur.hi = ord_of(S.fin[k] ? 0.0 : (double)sz); cur.lo = ~(unsigned long long)(unsigned)k;
        loc = pick_max(loc, cur);
      }
    }
    const Pick bw = block_pick(loc, slots);
    if (bw.hi == 0ull || ord_to_double(bw.hi) == 0.0) break;   // no areas at all (:212) or all finalised
    const int best = (int)(unsigned)(~bw.lo);
    ++n_rounds;
    if (best != cur_best) {                    // materialise V[best] in list order
      int off = 0;
      for (int s = best; s >= 0; s = S.seg_next[s]) {
        const int st = S.a_start[s], ln = S.a_len[s];
        for (int i = tid; i < ln; i += NT) { const int c = S.s1_cells[st + i]; hc[off + i] = c; hn[off + i] = cnode[c]; }
        off += ln;
      }
      nb = off;
      cur_best = best;
      __syncthreads();
      extend_D(0, nb);
    }
    if (tid == 0) sh_i[2] = 0;
    __syncthreads();
    TICK(4);                              // pick the largest area, materialise its lists / dense block
    // --- neighbouring areas in discovery order (:217-223): key = (position of X in V[best], dict key, dir)
    for (int p = tid; p < nb; p += NT) {
      const int cc = hc[p];
      const int ci = cc / Y, cj = cc - ci * Y;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int a = ci + (d == 0 ? -1 : (d == 1 ? 1 : 0));
        int q = cj + (d == 2 ? -1 : (d == 3 ? 1 : 0));
        if (a < 0 || a >= X) continue;
        bool wrapped = false;
        if (q < 0) { if (!latlon) continue; q = Y - 1; wrapped = true; }
        if (q >= Y) { if (!latlon) continue; q = 0; wrapped = true; }
        const int kk = lab[a * Y + q];
        if (kk < 0 || kk == best) continue;
        if (!wrapped && S.fin[kk]) continue;          // finalised cells are `unavail` (sentinel), :54-77
        const unsigned long long key = ((unsigned long long)p << 32) | ((unsigned long long)kk << 2) | (unsigned)d;
        const unsigned long long old = atomicMin(&S.okey[kk], key);
        if (old == NOKEY64) S.nlist[atomicAdd(&sh_i[2], 1)] = kk;
      }
    }
    __syncthreads();
    const int nn = sh_i[2];
    TICK(5);                              // neighbour discovery
    // --- hypothetical merges (:224-253), neighbours processed in chunks that fit the row-mean buffer
    int q0 = 0;
    while (q0 < nn) {
      // chunk = neighbours q0..q1-1 ; every thread computes the same partition
      int q1 = q0;
      size_t rows = 0, cells = 0;
      while (q1 < nn && q1 - q0 < MAXCH) {
        const int sz = S.size[S.nlist[q1]];
        if (q1 > q0 && (rows + nb + sz > RM || cells + sz > (size_t)C)) break;
        rows += nb + sz; cells += sz; ++q1;
      }
      // materialise the neighbours' node lists; sh_koff = node-list offsets, sh_uoff = row-unit offsets
      if (tid == 0) {
        int off = 0, uo = 0;
        for (int q = q0; q < q1; ++q) {
          const int sz = S.size[S.nlist[q]];
          sh_koff[q - q0] = off; sh_uoff[q - q0] = uo;
          off += sz; uo += nb + sz;
        }
        sh_koff[q1 - q0] = off; sh_uoff[q1 - q0] = uo;
      }
      __syncthreads();
      for (int q = q0; q < q1; ++q) {
        int off = sh_koff[q - q0];
        for (int s = S.nlist[q]; s >= 0; s = S.seg_next[s]) {
          const int st = S.a_start[s], ln = S.a_len[s];
          for (int i = tid; i < ln; i += NT) knl[off + i] = cnode[S.s1_cells[st + i]];
          off += ln;
        }
      }
      __syncthreads();
      TICK(6);                            // neighbour node lists
      // row means r_p = nanmean(R[hyp_p, hyp_q], q > p): ONE THREAD per (neighbour, row), consecutive threads on
      // consecutive rows so the best x best part streams coalesced from the diagonal-major dense block
      {
        const int nch = q1 - q0;
        const int total = sh_uoff[nch];
        for (int u = tid; u < total; u += NT) {
          int qc = 0;
          while (u >= sh_uoff[qc + 1]) ++qc;
          const int p = u - sh_uoff[qc];
          const int nk = sh_koff[qc + 1] - sh_koff[qc];
          const int n = nb + nk;
          const int32_t* kn = knl + sh_koff[qc];
          const int len = n - 1 - p;
          int nanc = 0;
          double sum = 0.0;
          if (len > 0) {
            wk += (unsigned long long)len;
            if (p < nb) {
              const double* row = R + (size_t)hn[p] * ldn;
              const int nbb = nb - 1 - p;              // elements of the row inside the best area
              if (d_ok) {
                const double* dcol = D + p;
                sum = sie_pw_sum_thread(
                    [&](int i) { return (i < nbb) ? dcol[(size_t)i * dcap] : __ldg(row + kn[i - nbb]); }, len, nanc);
              } else {
                sum = sie_pw_sum_thread(
                    [&](int i) { return __ldg(row + ((i < nbb) ? hn[p + 1 + i] : kn[i - nbb])); }, len, nanc);
              }
            } else {
              const double* row = R + (size_t)kn[p - nb] * ldn;
              const int32_t* kq = kn + (p - nb) + 1;
              sum = sie_pw_sum_thread([&](int i) { return __ldg(row + kq[i]); }, len, nanc);
            }
          }
          S.rowmean[u] = (len - nanc > 0) ? sum / (double)(len - nanc) : sie_nan();   // nanmean([]) = nan
        }
      }
      __syncthreads();
      TICK(7);                            // row means of the hypothetical areas
      // stat_k = nanmean(r_0..r_{n-1})  (:253) -- one 8-lane group per neighbour
      for (int q = q0 + g; q < q1; q += NG) {
        const int n = sh_uoff[q - q0 + 1] - sh_uoff[q - q0];
        const double* rm = S.rowmean + sh_uoff[q - q0];
        int nanc = 0;
        const double sum = sie_pw_sum8([&](int i) { return rm[i]; }, n, j, gmask, nanc);
        nanc += __shfl_xor_sync(gmask, nanc, 1);
        nanc += __shfl_xor_sync(gmask, nanc, 2);
        nanc += __shfl_xor_sync(gmask, nanc, 4);
        if (j == 0) S.stat[S.nlist[q]] = (n - nanc > 0) ? sum / (double)(n - nanc) : sie_nan();
      }
      __syncthreads();
      TICK(8);                            // statistic per neighbour
      q0 = q1;
    }
    // --- max(Anei_Rs.items(), key=itemgetter(1)) (:255): first in discovery order wins ties; a NaN in
    //     first position is never displaced (list comparison semantics)
    Pick loc2 = pick_none(), f1 = pick_none();
    for (int q = tid; q < nn; q += NT) {
      const int kk = S.nlist[q];
      const double st = S.stat[kk];
      const unsigned long long ok = S.okey[kk];
      Pick c1; c1.hi = 1ull; c1.lo = ~ok;               // first discovered neighbour = smallest discovery key
      f1 = pick_max(f1, c1);
      if (st == st) {
        Pick cur; cur.hi = ord_of(st); cur.lo = ~ok;
        loc2 = pick_max(loc2, cur);
      }
    }
    const Pick firstn = block_pick(f1, slots);
    const Pick win = block_pick(loc2, slots);
    bool merge = false;
    int kk = -1;
    if (nn > 0 && win.hi != 0ull) {
      const int first_idx = (int)((~firstn.lo >> 2) & 0x3fffffffull);
      const double first_stat = S.stat[first_idx];
      if (first_stat == first_stat && ord_to_double(win.hi) > tau) { merge = true; kk = (int)((~win.lo >> 2) & 0x3fffffffull); }
    }
    __syncthreads();
    // reset discovery keys
    for (int q = tid; q < nn; q += NT) S.okey[S.nlist[q]] = NOKEY64;
    if (merge) {
      // V[best] += V.pop(kk)  (:259-261): append kk's cells to the materialised lists and relabel
      int off = nb;
      for (int s = kk; s >= 0; s = S.seg_next[s]) {
        const int st = S.a_start[s], ln = S.a_len[s];
        for (int i = tid; i < ln; i += NT) {
          const int c = S.s1_cells[st + i];
          hc[off + i] = c; hn[off + i] = cnode[c]; lab[c] = best;
        }
        off += ln;
      }
      __syncthreads();
      if (tid == 0) {
        S.seg_next[S.tail[best]] = kk;
        S.tail[best] = S.tail[kk];
        S.size[best] += S.size[kk];
        S.size[kk] = 0;
      }
      extend_D(nb, off);
      nb = off;
    } else {
      if (tid == 0) S.fin[best] = 1;           // :262-265 all cells of V[best] become unavailable
    }
    __syncthreads();
    TICK(9);                              // winner, merge / finalise, dense block extension
  }

  // =============================================================== output in dict order (ascending key)
  if (wk) atomicAdd(&sh_work, wk);
  __syncthreads();
  if (tid == 0 && work_all) {
    const long long clk2 = clock64();
    work_all[SIE_AREA_WORK * b] = sh_work;                                  // correlations consumed
    work_all[SIE_AREA_WORK * b + 1] = (unsigned long long)(clk1 - clk0);    // SM cycles in step 1
    work_all[SIE_AREA_WORK * b + 2] = (unsigned long long)(clk2 - clk1);    // SM cycles in step 2
    work_all[SIE_AREA_WORK * b + 3] = (n_steps << 32) | n_rounds;           // growth steps, merge rounds
    for (int i = 0; i < 11; ++i) work_all[SIE_AREA_WORK * b + 4 + i] = ph[i];
    work_all[SIE_AREA_WORK * b + 15] = ph[11];
  }
  if (tid == 0) {
    int cnt = 0, off = 0;
    for (int k = 0; k < nA; ++k) {
      if (S.size[k] > 0) {
        out_key[cnt] = k;
        out_start[cnt] = off;
        S.nlist[cnt] = k;
        off += S.size[k];
        ++cnt;
      }
    }
    out_start[cnt] = off;
    n_areas_all[b] = cnt;
    status_all[b] = (cnt < 2) ? SIE_JOB_FEW_AREAS : SIE_JOB_OK;   // :212 / :278 ValueError
    sh_i[3] = cnt;
  }
  __syncthreads();
  const int cnt = sh_i[3];
  for (int a = 0; a < cnt; ++a) {
    int off = out_start[a];
    for (int s = S.nlist[a]; s >= 0; s = S.seg_next[s]) {
      const int st = S.a_start[s], ln = S.a_len[s];
      for (int i = tid; i < ln; i += NT) {
        const int c = S.s1_cells[st + i];
        out_cells[off + i] = c;
        out_label[c] = a;
      }
      off += ln;
    }
  }
}

}  // namespace

extern "C" size_t sie_area_level_scratch_bytes(int B, int C) {
  // per-area arrays are sized for the worst case max_areas = C/2 + 1
  return (size_t)B * scratch_per_job(C, C / 2 + 1);
}

extern "C" int sie_area_level(const double* R, const double* stencil, const int32_t* node_cell,
                              const int32_t* cell_node, const int32_t* n_nodes, const double* tau,
                              const int32_t* first_nan_cell, int B, int X, int Y, int ldn, int latlon,
                              int max_areas, int32_t* area_cells, int32_t* area_start, int32_t* area_key,
                              int32_t* n_areas, int32_t* label, int32_t* status, void* scratch,
                              size_t scratch_bytes, uint64_t* work, void* stream) {
  SIE_CHECK_ARG(R && stencil && node_cell && cell_node && n_nodes && tau && first_nan_cell && area_cells &&
                    area_start && area_key && n_areas && label && status && scratch, "null pointer");
  SIE_CHECK_ARG(B > 0 && X > 0 && Y > 0 && ldn > 0 && max_areas > 0, "non-positive size");
  const int C = X * Y;
  SIE_CHECK_ARG(max_areas <= C / 2 + 1, "max_areas cannot exceed C/2+1");
  SIE_CHECK_ARG((long long)C < (1ll << 28), "grid too large for the frontier key encoding");
  const size_t per_job = scratch_per_job(C, max_areas);
  SIE_CHECK_ARG(scratch_bytes >= per_job * (size_t)B, "scratch too small");
  int dev = 0, max_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t slot_bytes = sizeof(double) * 16 * FCAP + sizeof(int32_t) * FCAP;
  // shared-memory tiers: 2 = all seven integer arrays, 1 = six (neighbour lists stay in global), 0 = slot state only
  size_t smem = slot_bytes + sizeof(int32_t) * (size_t)7 * C;
  int use_smem = 2;
  if (smem + 4096 > (size_t)max_optin) { smem = slot_bytes + sizeof(int32_t) * (size_t)6 * C; use_smem = 1; }
  if (smem + 4096 > (size_t)max_optin) { smem = slot_bytes; use_smem = 0; }
  cudaFuncSetAttribute(k_area_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_area_level<<<B, NT, smem, (cudaStream_t)stream>>>(
      R, stencil, node_cell, cell_node, n_nodes, tau, first_nan_cell, X, Y, ldn, latlon, max_areas, area_cells,
      area_start, area_key, n_areas, label, status, (unsigned char*)scratch, per_job, use_smem,
      (unsigned long long*)work);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}
