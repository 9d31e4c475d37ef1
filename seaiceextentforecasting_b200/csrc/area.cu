// K4 + K5: tau-thresholded domain growth (step 1) and largest-first merging (step 2).
// Reference: Network.area_level, ComplexNetworks.py:49-278 (semantics restated in SURVEY.md App. A).
//
// The algorithm is sequential per network (every decision consumes cells later decisions could have used),
// so one persistent CTA owns one network and B networks run side by side.  Every growth / merge decision is a
// float comparison of a numpy-order pairwise sum, so the sums are evaluated in numpy's order (8 strided
// accumulators per <=128-element leaf, recursive halving above) and the comparisons `mean > tau` / first-max
// pick the same cell as the reference when fed the same R.  Candidate order (direction-major, then position in
// the area list, duplicates kept) only matters for ties; it is carried as a per-cell key instead of
// materialising the reference's candidate list.
//
// The kernel is bound by its chain of dependent steps (one step = one cell added / one merge decided), so the
// design minimises the critical path of a step rather than bytes:
//  * step 1, areas of <= 128 cells (one pairwise leaf): every frontier cell owns a slot = one thread.  The slot
//    keeps numpy's 8 accumulators, the <8-element tail and the running sum in shared memory; adding a member
//    appends ONE correlation per slot (one gather, issued right after the winner is known) and one add.  A
//    dedicated warp updates the integer state and builds the state of the <=3 new frontier cells (one 8-lane
//    group each, all gathers in flight together) while the slot gathers are outstanding.  The CTA-wide argmax is
//    two levels of redux.sync on an order-preserving integer image of (mean, tie-break key): ~300 cycles.
//    Larger areas re-sum gathered lists (8-lane groups, lane j = accumulator j).
//  * step 2 keeps a dense row-major list-order copy D[p][q] = R[best_p][best_q] of the current largest area,
//    extended when a neighbour is merged, so the O(n^2) "mean of row means" statistic of every hypothetical
//    merge streams contiguous rows (8 lanes read 64 contiguous bytes) instead of re-gathering the best x best
//    block for every candidate and round.
// Integer state and the per-area tables live in shared memory; arrays that do not fit (big grids) fall back to
// global scratch one by one.  No roofline fraction is claimed for this kernel; see DESIGN.md.
//
// Throughput (the metric is forecasts/s over many networks, not the latency of one): the grid is PERSISTENT -- CTAs
// pop jobs from an atomic queue in index order (the host puts the long-window, i.e. slowest, networks first: LPT), so
// an SM is never idle while jobs remain and the global scratch is per CTA, not per job.  Three instantiations:
//   <256 threads, 16-bit indices>  small grids (57x57 SIC, 26x90 SST): ~105 KB of shared memory and 128 registers,
//                                  so TWO latency-bound chains share an SM and overlap each other's stalls;
//   <512 threads, 16-bit indices>  mid-size grids (81x81): every array still on chip, one CTA per SM;
//   <512 threads, 32-bit indices>  anything larger (25 km): arrays evicted to global scratch as needed.
#include <type_traits>
#include "common.cuh"

// Gathers from R are single-use: with two CTAs per SM only ~10 KB of L1 is left, so by default they are cached in L2
// only (ld.global.cg) and L1 keeps the dense step-2 scratch.  -DSIE_AREA_LDG restores read-only-path loads (A/B timing).
#ifdef SIE_AREA_LDG
#define SIE_RLOAD(p) __ldg(p)
#else
#define SIE_RLOAD(p) __ldcg(p)
#endif

namespace {

constexpr int LEAF = 128;          // numpy's pairwise block size
constexpr int DCAP_MAX = 1024;     // rows/cols of the dense best-area sub-matrix
constexpr int MAXCH = 32;          // neighbour areas evaluated per chunk of a merge round
constexpr unsigned long long NOKEY64 = ~0ull;
constexpr unsigned FULL = 0xffffffffu;

// NT threads; FCAP frontier slots with incremental state (slot s is owned by thread s); the last warp updates the
// integer state and initialises new slots.  IT = type of the per-cell index arrays and the per-area tables.
template <int NT_, typename IT_>
struct AreaCfg {
  static constexpr int NT = NT_;
  static constexpr int NW = NT_ / 32;
  static constexpr int NG = NT_ / 8;                  // 8-lane groups per CTA
  static constexpr int FCAP = NT_ >= 512 ? 256 : 192;
  static constexpr int BKW = NW - 1;
  static constexpr int CTAS = NT_ >= 512 ? 1 : 2;     // co-resident CTAs per SM the launch bounds allow
  using IT = IT_;
  using FK = typename std::conditional<sizeof(IT_) == 2, uint16_t, uint32_t>::type;   // frontier key
  static constexpr FK NOKEY = (FK)~(FK)0;
  static constexpr int DSHIFT = sizeof(IT_) == 2 ? 14 : 28;   // key = direction << DSHIFT | list position
  static_assert(FCAP <= NT - 32, "slot threads and the bookkeeping warp must be distinct");
};

// which arrays live in global scratch instead of shared memory (bit set = global)
enum : int { PL_LAB = 1, PL_FKEY = 2, PL_FLIST = 4, PL_HN = 8, PL_HC = 16, PL_CNL = 32, PL_KNL = 64, PL_AREA = 128 };

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline size_t rm_cap(int C) { return (size_t)4 * C + 1024 + 256; }     // + read-ahead pad
__host__ __device__ inline int d_cap(int C) { return C < DCAP_MAX ? C : DCAP_MAX; }
__host__ __device__ inline size_t area_tab_bytes(int MA, size_t isz) {       // 12 integer tables + okey + stat
  return 12 * align_up(isz * (size_t)(MA + 1), 16) + 2 * align_up(sizeof(double) * (size_t)MA, 16);
}
__host__ __device__ inline size_t self_cap(int C) { return (size_t)2 * C + 64; }           // allocation adds a read-ahead pad
__host__ __device__ inline size_t slot_bytes(int fcap) { return sizeof(double) * (18 * (size_t)fcap) + sizeof(int32_t) * 2 * (size_t)fcap; }
constexpr size_t QUEUE_BYTES = 256;                         // head of the scratch: the job-queue counter
__host__ __device__ inline size_t scratch_per_cta(int C, int MA) {
  size_t s = 0;
  s += align_up(sizeof(int32_t) * (size_t)C, 256);          // s1_cells
  s += align_up(sizeof(int32_t) * (size_t)7 * C, 256);      // fallback for the shared-memory integer arrays
  s += align_up(sizeof(double) * rm_cap(C), 256);           // rowmean
  s += align_up(sizeof(double) * ((size_t)d_cap(C) * d_cap(C) + 128), 256);   // dmat (+ read-ahead pad)
  s += align_up(area_tab_bytes(MA, sizeof(int32_t)), 256);  // fallback for the per-area tables
  s += align_up(sizeof(double) * (self_cap(C) + 256), 256);   // cached self row means of candidate areas (+ read-ahead pad)
  return s;
}

template <typename IT>
struct AreaTabs {       // per step-1 area (segment); merged areas are chains of segments
  IT *a_start, *a_len, *seg_next, *tail, *size, *fin, *nlist;
  IT *xoff, *xep, *xrows;          // step 2: column offset of the area's cross block in D, epoch it belongs to, best rows filled
  IT *self_off, *self_sz;          // step 2: cached self row means (offset into selfbuf; valid when self_sz == size)
  unsigned long long* okey;
  double* stat;
};
template <typename IT>
__device__ inline AreaTabs<IT> carve_tabs(unsigned char* p, int MA) {
  AreaTabs<IT> t;
  auto take = [&](size_t bytes) { unsigned char* r = p; p += align_up(bytes, 16); return r; };
  t.a_start = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.a_len = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.seg_next = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.tail = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.size = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.fin = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.nlist = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.xoff = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.xep = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.xrows = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.self_off = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.self_sz = (IT*)take(sizeof(IT) * (size_t)(MA + 1));
  t.okey = (unsigned long long*)take(sizeof(double) * (size_t)MA);
  t.stat = (double*)take(sizeof(double) * (size_t)MA);
  return t;
}

// Argmax record.  `ord` is an order-preserving image of the value (0 = no candidate), `pri` the tie-break
// priority (larger wins: the bitwise complement of the reference's candidate-list position, so the earliest
// candidate wins ties); a, b are payload.
struct Rec {
  unsigned long long ord, pri;
  uint32_t a, b;
};
__device__ __forceinline__ Rec rec_none() { Rec r; r.ord = 0ull; r.pri = 0ull; r.a = 0u; r.b = 0u; return r; }
__device__ __forceinline__ unsigned long long ord_of(double x) {      // monotone map double -> u64, never 0
  x = __dadd_rn(x, 0.0);                                               // -0.0 -> +0.0 (they compare equal)
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_to_double(unsigned long long o) {
  const unsigned long long u = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ Rec rec_max(const Rec& x, const Rec& y) {
  const bool ty = (y.ord > x.ord) || (y.ord == x.ord && y.pri > x.pri);
  return ty ? y : x;
}
// Warp argmax by lexicographic redux.sync over 32-bit words (PW = 32-bit words of `pri` that matter: 1 -> only
// the high word).  Every lane returns the winner.
template <int PW>
__device__ __forceinline__ Rec warp_arg(const Rec& r) {
  const uint32_t w0 = (uint32_t)(r.ord >> 32), w1 = (uint32_t)r.ord;
  const uint32_t w2 = (uint32_t)(r.pri >> 32), w3 = (uint32_t)r.pri;
  const uint32_t m0 = __reduce_max_sync(FULL, w0);
  bool ok = (w0 == m0);
  {   // common case: the high word alone decides (one lane holds the maximum) -> one redux + one ballot
    const unsigned b0 = __ballot_sync(FULL, ok);
    if ((b0 & (b0 - 1u)) == 0u) {
      const int src0 = __ffs(b0) - 1;
      Rec o;
      o.ord = ((unsigned long long)m0 << 32) | __shfl_sync(FULL, w1, src0);
      o.pri = ((unsigned long long)__shfl_sync(FULL, w2, src0) << 32) | __shfl_sync(FULL, w3, src0);
      o.a = __shfl_sync(FULL, r.a, src0);
      o.b = __shfl_sync(FULL, r.b, src0);
      return o;
    }
  }
  const uint32_t m1 = __reduce_max_sync(FULL, ok ? w1 : 0u);
  ok = ok && (w1 == m1);
  const uint32_t m2 = __reduce_max_sync(FULL, ok ? w2 : 0u);
  ok = ok && (w2 == m2);
  uint32_t m3 = 0u;
  if (PW > 1) {
    m3 = __reduce_max_sync(FULL, ok ? w3 : 0u);
    ok = ok && (w3 == m3);
  }
  const int src = __ffs(__ballot_sync(FULL, ok)) - 1;      // never empty: the maximal lane survives every stage
  Rec o;
  o.ord = ((unsigned long long)m0 << 32) | m1;
  o.pri = ((unsigned long long)m2 << 32) | m3;
  o.a = __shfl_sync(FULL, r.a, src);
  o.b = __shfl_sync(FULL, r.b, src);
  return o;
}
struct RecSlot { unsigned long long ord, pri; uint32_t a, b, pad0, pad1; };   // 32 bytes
// CTA-wide argmax; every thread returns the same winner.  `slots` is shared scratch [2][NW]; `par` alternates so
// the next call never overwrites records a slow warp is still reading (one barrier per call).
template <int PW, int NW>
__device__ __forceinline__ Rec block_arg(const Rec& v, RecSlot* slots, int& par) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Rec w = warp_arg<PW>(v);
  RecSlot* s = slots + par * NW;
  if (lane == 0) { s[warp].ord = w.ord; s[warp].pri = w.pri; s[warp].a = w.a; s[warp].b = w.b; }
  __syncthreads();
  Rec t = rec_none();
  if (lane < NW) { t.ord = s[lane].ord; t.pri = s[lane].pri; t.a = s[lane].a; t.b = s[lane].b; }
  par ^= 1;
  return warp_arg<PW>(t);
}

// numpy pairwise sum (8-lane group, lane j) of the sequence a[0..na) ++ b[0..n-na) held in (global) memory: leaves that
// lie inside one run are read with immediate-offset loads, others (and leaves holding a NaN) element by element with
// NaN screening.  ONE out-of-line copy serves the row means and the statistics of every merge round: the kernel's hot
// instruction footprint, not its arithmetic, limits two co-resident CTAs (profiles/r02_*).  Reads up to 135 doubles past
// the end of a run (masked out of the sum): the scratch buffers are padded accordingly.
#ifdef SIE_AREA_INLINE_RUNS
#define SIE_RUNS_ATTR __forceinline__
#else
#define SIE_RUNS_ATTR __noinline__
#endif
#ifdef SIE_AREA_INLINE_GATHER
#define SIE_GATHER_ATTR __forceinline__
#else
#define SIE_GATHER_ATTR __noinline__
#endif
template <int MAXD>
__device__ SIE_RUNS_ATTR double sie_pw_sum_runs(const double* __restrict__ a, int na, const double* __restrict__ b, int n,
                                               int j, unsigned gmask, int* nan_out) {
  int nanc = 0;
  const double sum = sie_pw_tree<MAXD>([&](int lo, int ln) -> double {
    if (lo + ln <= na || lo >= na) {
      const double r = sie_pw_leaf8_contig((lo >= na ? b + (lo - na) : a + lo) + j, ln, j, gmask);
      if (r == r) return r;
    } else {     // the leaf straddles the two runs: lane j's first ks full-group elements lie in run a
      const int rem = na - lo - j;
      const double r = sie_pw_leaf8_two(a + lo + j, b + (lo + j - na), rem > 0 ? (rem + 7) >> 3 : 0, ln, j, gmask);
      if (r == r) return r;
    }
    return sie_pw_leaf8([&](int i) { return i < na ? a[i] : b[i - na]; }, lo, ln, j, gmask, nanc);
  }, n);
  nanc += __shfl_xor_sync(gmask, nanc, 1);
  nanc += __shfl_xor_sync(gmask, nanc, 2);
  nanc += __shfl_xor_sync(gmask, nanc, 4);
  *nan_out = nanc;
  return sum;
}

struct RCtx { const double* R; int ldn, Tp, kT; };    // where correlations come from: R (stride ldn) or z rows (stride Tp)
__device__ double sie_zcorr(const double* __restrict__ za, const double* __restrict__ zc, int kT, bool diag);
// W32: ldn^2 < 2^31 (16-bit index variants), so the element offset min*ldn + max is 32-bit arithmetic: one IMAD + two
// min/max instead of 64-bit selects and multiplies at every gather site (hot-path instructions AND code footprint)
template <bool ZR, bool W32>
__device__ __forceinline__ double sie_rat(const RCtx& cx, int a, int c) {
  if constexpr (ZR) return sie_zcorr(cx.R + (size_t)a * cx.Tp, cx.R + (size_t)c * cx.Tp, cx.kT, a == c);
  else if constexpr (W32) return SIE_RLOAD(cx.R + (unsigned)(min(a, c) * cx.ldn + max(a, c)));
  else return (c >= a) ? SIE_RLOAD(cx.R + (size_t)a * cx.ldn + c) : SIE_RLOAD(cx.R + (size_t)c * cx.ldn + a);   // R[min][max]
}
// numpy pairwise sum of the correlations of node `rown` with the nodes ia[0..na) ++ ib[0..n-na) (index lists in shared
// or global memory), NaN entries screened and counted: one out-of-line copy for the re-summing growth steps, the
// neighbours' own row means and the gather path of the merge rounds.
template <int MAXD, typename IT, bool ZR>
__device__ SIE_GATHER_ATTR double sie_pw_sum_gather(RCtx cx, int rown, const IT* ia, int na, const IT* ib, int n, int j,
                                                 unsigned gmask, int* nan_out) {
  int nanc = 0;
  const double sum = sie_pw_sum8<MAXD>([&](int i) { return sie_rat<ZR, sizeof(IT) == 2>(cx, rown, (int)(i < na ? ia[i] : ib[i - na])); },
                                       n, j, gmask, nanc);
  nanc += __shfl_xor_sync(gmask, nanc, 1);
  nanc += __shfl_xor_sync(gmask, nanc, 2);
  nanc += __shfl_xor_sync(gmask, nanc, 4);
  *nan_out = nanc;
  return sum;
}

// Correlation of nodes a and c recomputed from their unit-norm rows (R not stored: `ZR` instantiations): the same
// dot product the correlation kernel accumulates (sequential in k), clipped, NaN on the diagonal.
__device__ __noinline__ double sie_zcorr(const double* __restrict__ za, const double* __restrict__ zc, int kT, bool diag) {
  if (diag) return sie_nan();
  double acc = 0.0;
  for (int k = 0; k < kT; k += 4) {
    const double2 a0 = *reinterpret_cast<const double2*>(za + k), a1 = *reinterpret_cast<const double2*>(za + k + 2);
    const double2 c0 = *reinterpret_cast<const double2*>(zc + k), c1 = *reinterpret_cast<const double2*>(zc + k + 2);
    acc = fma(a0.x, c0.x, acc);
    acc = fma(a0.y, c0.y, acc);
    acc = fma(a1.x, c1.x, acc);
    acc = fma(a1.y, c1.y, acc);
  }
  if (!(fabs(acc) <= 1.0)) acc = acc > 1.0 ? 1.0 : (acc < -1.0 ? -1.0 : acc);   // np.clip keeps NaN
  return acc;
}

template <class CFG, bool ONCHIP, bool ZR>
__global__ void __launch_bounds__(CFG::NT, CFG::CTAS)
k_area_level(const double* __restrict__ Rall, const int32_t* __restrict__ job_T, int Tp,
             const double* __restrict__ stencil_all,
             const int32_t* __restrict__ node_cell_all, const int32_t* __restrict__ cell_node_all,
             const int32_t* __restrict__ n_nodes, const double* __restrict__ tau_all,
             const int32_t* __restrict__ first_nan_cell, int B, int X, int Y, int ldn, int latlon, int MA,
             int32_t* __restrict__ area_cells_all, int32_t* __restrict__ area_start_all,
             int32_t* __restrict__ area_key_all, int32_t* __restrict__ n_areas_all,
             int32_t* __restrict__ label_all, int32_t* __restrict__ status_all,
             int* __restrict__ queue, unsigned char* __restrict__ scratch_all, size_t scratch_stride, int place_arg,
             unsigned long long* __restrict__ work_all) {
  constexpr int NT = CFG::NT, NW = CFG::NW, NG = CFG::NG, FCAP = CFG::FCAP, BKW = CFG::BKW;
  using IT = typename CFG::IT;
  using FK = typename CFG::FK;
  constexpr FK NOKEY = CFG::NOKEY;
  constexpr int MAXD = ONCHIP ? 6 : 16;        // pairwise tree depth: lists are <= C cells, C < 8192 when ONCHIP
  const int place = ONCHIP ? 0 : place_arg;   // ONCHIP: every array in shared memory (pointers known to be shared)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ RecSlot rslots[2 * NW];
  __shared__ int sh_i[8];
  __shared__ int sh_koff[MAXCH + 1], sh_uoff[MAXCH + 1];
  __shared__ int sh_goff[MAXCH + 1], sh_soff[MAXCH + 1], sh_xo[MAXCH], sh_r0[MAXCH], sh_so[MAXCH];
  __shared__ int sh_x[4];
  __shared__ int sh_job;
  __shared__ unsigned long long sh_work;
  __shared__ unsigned long long ph[16];   // per-phase SM cycles (thread 0's view), reported through work[4..15]
  long long tick_last = 0;
#ifdef SIE_AREA_PHASE_TIMERS
#define TICK(i)                                                        \
  do {                                                                 \
    if (tid == 0) {                                                    \
      const long long t__ = clock64();                                 \
      ph[i] += (unsigned long long)(t__ - tick_last);                  \
      tick_last = t__;                                                 \
    }                                                                  \
  } while (0)
#else
#define TICK(i) do { } while (0)
#endif
  long long btick = 0;
#ifdef SIE_AREA_PHASE_TIMERS
#define BTICK(i)                                                       \
  do {                                                                 \
    const long long t__ = clock64();                                   \
    bph[i] += (unsigned long long)(t__ - btick);                       \
    btick = t__;                                                       \
  } while (0)
#else
#define BTICK(i) do { } while (0)
#endif

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid >> 3, j = tid & 7;
  const unsigned gmask = 0xffu << (lane & 24);
  const int C = X * Y;
  int par = 0;

  // ---- global scratch of this CTA (reused by every job it pops)
  unsigned char* gp = scratch_all + (size_t)blockIdx.x * scratch_stride;
  auto gtake = [&](size_t bytes) { unsigned char* r = gp; gp += align_up(bytes, 256); return r; };
  int32_t* s1_cells = (int32_t*)gtake(sizeof(int32_t) * (size_t)C);   // step-1 member cells, area after area
  IT* ibuf = (IT*)gtake(sizeof(int32_t) * (size_t)7 * C);
  double* rowmean = (double*)gtake(sizeof(double) * rm_cap(C));
  const int dcap = d_cap(C);
  double* D = (double*)gtake(sizeof(double) * ((size_t)dcap * dcap + 128));
  unsigned char* gtabs = gtake(area_tab_bytes(MA, sizeof(int32_t)));
  double* selfbuf = (double*)gtake(sizeof(double) * (self_cap(C) + 256));
  // ---- shared memory: slot state | per-area tables | integer arrays (whatever `place` keeps on chip)
  unsigned char* sp = smem_raw;
  auto stake = [&](size_t bytes) { unsigned char* r = sp; sp += align_up(bytes, 16); return r; };
  double* facc = (double*)stake(sizeof(double) * 8 * FCAP);    // [8][FCAP] numpy accumulator q of slot s at q*FCAP+s
  double* ftail = (double*)stake(sizeof(double) * 8 * FCAP);   // [8][FCAP] tail element t of slot s at t*FCAP+s
  double* sres = (double*)stake(sizeof(double) * FCAP);        // running pairwise sum of the slot's list
  double* smean = (double*)stake(sizeof(double) * FCAP);       // its nanmean
  int32_t* snan = (int32_t*)stake(sizeof(int32_t) * FCAP);     // NaN entries seen
  int32_t* srow = (int32_t*)stake(sizeof(int32_t) * FCAP);     // node of the slot's cell, -1 = dead slot
  AreaTabs<IT> S = carve_tabs<IT>((place & PL_AREA) ? gtabs : stake(area_tab_bytes(MA, sizeof(IT))), MA);
  auto iarr = [&](int bit, int idx) -> IT* {
    return (place & bit) ? ibuf + (size_t)idx * C : (IT*)stake(sizeof(IT) * (size_t)C);
  };
  IT* lab = iarr(PL_LAB, 0);                           // [C] area key of each cell, -1 = unassigned
  FK* fkey = (FK*)iarr(PL_FKEY, 1);                    // [C] frontier key (step 1)
  IT* flist = iarr(PL_FLIST, 2);                       // [C] frontier cell of each slot (step 1), -1 = dead
  IT* hn = iarr(PL_HN, 3);                             // [C] node list of the current area (step 1) / best area (step 2)
  IT* hc = iarr(PL_HC, 4);                             // [C] cell list of the best area (step 2)
  IT* cnl = iarr(PL_CNL, 5);                           // [C] local copy of cell -> node
  IT* knl = iarr(PL_KNL, 6);                           // [C] node lists of the neighbour areas (step 2)
  const IT* cnode = cnl;

  // =============================================================== persistent loop over the job queue
  for (;;) {
  __syncthreads();                        // the previous job's state is no longer read
  if (tid == 0) sh_job = atomicAdd(queue, 1);
  __syncthreads();
  const int b = sh_job;
  if (b >= B) break;
  unsigned long long wk = 0;   // correlations consumed (algorithmic gathers)
#ifdef SIE_AREA_PHASE_TIMERS
  unsigned long long rbucket[6] = {0, 0, 0, 0, 0, 0};
#endif
  unsigned long long bph[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // bookkeeping-warp phase cycles (its lane 0), work[16..23]
  const int N = min(n_nodes[b], ldn);
  const double tau = tau_all[b];
  // ZR: `Rall` holds the unit-norm rows z [B][ldn][Tp] and every correlation is recomputed from them
  const double* R = ZR ? Rall + (size_t)b * ldn * Tp : Rall + (size_t)b * ldn * ldn;
  const int kT = ZR ? ((job_T[b] + 3) & ~3) : 0;
  const RCtx rcx = {R, ldn, Tp, kT};
  struct RowH { const double* p; int n; };
  auto rrow = [&](int a) -> RowH { RowH h; h.n = a; h.p = R + (size_t)a * (ZR ? Tp : ldn); return h; };
  auto rat = [&](const RowH& h, int c) -> double {
    if constexpr (ZR) return sie_zcorr(h.p, R + (size_t)c * Tp, kT, h.n == c);
    else if constexpr (sizeof(IT) == 2) return SIE_RLOAD(R + (unsigned)(min(c, h.n) * ldn + max(c, h.n)));   // 32-bit offsets
    else return (c >= h.n) ? SIE_RLOAD(h.p + c) : SIE_RLOAD(R + (size_t)c * ldn + h.n);   // upper triangle stored: R[min][max]
  };
  const double* sten = stencil_all + (size_t)b * ldn * 4;
  const int32_t* cnode_g = cell_node_all + (size_t)b * C;
  int32_t* out_cells = area_cells_all + (size_t)b * C;
  int32_t* out_start = area_start_all + (size_t)b * (MA + 1);
  int32_t* out_key = area_key_all + (size_t)b * MA;
  int32_t* out_label = label_all + (size_t)b * C;

  for (int c = tid; c < C; c += NT) { lab[c] = -1; fkey[c] = NOKEY; out_label[c] = -1; cnl[c] = (IT)cnode_g[c]; }
  if (tid < 16) ph[tid] = 0ull;
  if (tid == 0) { n_areas_all[b] = 0; out_start[0] = 0; sh_work = 0ull; if (work_all) work_all[SIE_AREA_WORK * b] = 0ull; }
  __syncthreads();
  if (first_nan_cell[b] < 0) {            // :50-51 IndexError in the reference
    if (tid == 0) status_all[b] = SIE_JOB_NO_NAN_CELL;
    continue;
  }
  if (status_all[b] == SIE_JOB_CAPACITY) continue;   // K1 already flagged this job

  // ---- step-1 helpers (bookkeeping warp only) -----------------------------------------------------------
  // State of slot s (frontier cell with node `fnode` vs the first n <= 128 member nodes hn[0..n)) by one 8-lane
  // group: lane j builds numpy's accumulator j = a[j] + a[8+j] + ... over the full groups of 8, keeps tail element
  // j, and the group assembles the pairwise sum in numpy's order.  All gathers are issued before the first add.
  auto init_slot = [&](int s, int fnode, int n) {
    const RowH row = rrow(fnode);
    const int ngrp = n >> 3, nt = n & 7;
    double r, tv;
    int nanc = 0;
    sie_pw_lane8_any([&](int i) { return rat(row, hn[i]); }, 0, n, ngrp, nt, j, r, tv, nanc);
    facc[j * FCAP + s] = r;
    ftail[j * FCAP + s] = tv;
    double res = 0.0;
    if (ngrp > 0) {     // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) : xor-butterfly inside the group (IEEE add commutes)
      res = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
      res = __dadd_rn(res, __shfl_xor_sync(gmask, res, 2));
      res = __dadd_rn(res, __shfl_xor_sync(gmask, res, 4));
    }
    res = sie_add_tail8(res, tv, nt, gmask);           // tail, left to right
    nanc += __shfl_xor_sync(gmask, nanc, 1);
    nanc += __shfl_xor_sync(gmask, nanc, 2);
    nanc += __shfl_xor_sync(gmask, nanc, 4);
    if (j == 0) {
      sres[s] = res;
      snan[s] = nanc;
      srow[s] = fnode;
      smean[s] = res / (double)(n - nanc);
    }
  };
  // gen_area_neighbours :80-94 for the member cell `m` at list position `p` (no lat-lon wrap here): lane d < 4 =
  // direction.  New frontier cells take the free slot `hole` first, then slots nf, nf+1, ...; with `do_init` their
  // incremental state over the `nmem` members is built.  Returns the new slot high-water mark; `fits` = all slots
  // are below FCAP.
  auto bk_add = [&](int m, int p, int nmem, int hole, int nf, bool do_init, bool& fits) -> int {
    bool isnew = false;
    int f = -1;
    if (lane < 4) {
      const int d = lane;
      const int ci = m / Y, cj = m - ci * Y;
      const int a = ci + (d == 0 ? -1 : (d == 1 ? 1 : 0));
      const int q = cj + (d == 2 ? -1 : (d == 3 ? 1 : 0));
      if (a >= 0 && a < X && q >= 0 && q < Y) {
        f = a * Y + q;
        const int fn = cnode[f];
        if (lab[f] >= 0 || fn < 0 || fn >= N) f = -1;
      }
      if (f >= 0) {
        const FK key = (FK)(((uint32_t)d << CFG::DSHIFT) | (uint32_t)p);
        const FK old = fkey[f];
        isnew = (old == NOKEY);
        if (key < old) fkey[f] = key;
      }
    }
    const unsigned mnew = __ballot_sync(FULL, isnew);
    const int cntnew = __popc(mnew);
    BTICK(1);                             // membership writes + neighbour detection
    int myslot = -1;
    if (isnew) {
      const int r = __popc(mnew & ((1u << lane) - 1u));
      myslot = (hole >= 0) ? (r == 0 ? hole : nf + r - 1) : nf + r;
      flist[myslot] = (IT)f;
    }
    const int nf_new = nf + cntnew - ((hole >= 0 && cntnew > 0) ? 1 : 0);
    if (hole >= 0 && cntnew == 0 && lane == 0) { flist[hole] = -1; if (hole < FCAP) srow[hole] = -1; }
    fits = (nf_new <= FCAP);
    __syncwarp();
    if (do_init && fits && cntnew > 0) {
      const int gi = lane >> 3;
      unsigned mm = mnew;
      for (int r = 0; r < gi; ++r) mm &= mm - 1u;          // drop the gi lowest set bits
      const int src = mm ? (__ffs(mm) - 1) : 0;
      const int cell = __shfl_sync(FULL, f, src);
      const int slot = __shfl_sync(FULL, myslot, src);
      BTICK(2);                           // slot assignment
      if (mm) init_slot(slot, cnode[cell], nmem);
      __syncwarp();
      BTICK(3);                           // new-slot state (gathers + numpy-order sums)
      bph[5] += 1;
    }
    __syncwarp();
    return nf_new;
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
    Rec cand = rec_none();
    if (dir >= 0) { cand.ord = 1ull; cand.pri = (unsigned long long)(~(uint32_t)c) << 32; cand.a = (uint32_t)c; cand.b = (uint32_t)dir; }
    const Rec sw = block_arg<1, NW>(cand, rslots, par);   // earliest cell wins
    TICK(0);                              // seed search
    if (sw.ord == 0ull) { c0 += NT; continue; }
    const int seed = (int)sw.a;
    const int sd = (int)sw.b;
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
    bool fast = true;                     // incremental slot state valid (uniform across the CTA)
    __syncthreads();                      // every thread has read lab[nbr] before it changes
    if (warp == BKW) {
      if (lane == 0) {
        lab[seed] = (IT)k; lab[nbr] = (IT)k;
        hn[0] = cnode[seed]; hn[1] = cnode[nbr];
        s1_cells[base] = seed; s1_cells[base + 1] = nbr;
      }
      __syncwarp();
      bool fits = true, fits2 = true;
      int cnt = bk_add(seed, 0, 2, -1, 0, true, fits);
      cnt = bk_add(nbr, 1, 2, -1, cnt, fits, fits2);
      if (lane == 0) { sh_i[1] = cnt; sh_i[2] = (fits && fits2) ? 1 : 0; }
    }
    __syncthreads();
    nf = sh_i[1];
    fast = sh_i[2] != 0;
    TICK(2);                              // area creation + first slot states
    while (nf > 0) {
      // --- evaluate every frontier cell's mean correlation with the area (:96-118) and pick the maximum
      Rec loc = rec_none();
      if (fast) {
        int cell = -1;
        if (tid < nf) cell = flist[tid];
        if (cell >= 0) {
          const double mean = smean[tid];
          if (mean == mean) { loc.ord = ord_of(mean); loc.pri = (unsigned long long)(~(uint32_t)fkey[cell]) << 32; loc.a = (uint32_t)tid; loc.b = (uint32_t)cell; }
        }
        const unsigned live = __ballot_sync(FULL, cell >= 0);
        if (lane == 0) wk += (unsigned long long)n * (unsigned long long)__popc(live);
      } else {
        ++n_slow;
        for (int q = g; q < nf; q += NG) {
          const int f = flist[q];
          if (f < 0) continue;
          int nanc = 0;
          const double sum = sie_pw_sum_gather<MAXD, IT, ZR>(rcx, cnode[f], hn, n, hn, n, j, gmask, &nanc);
          if (j == 0) wk += (unsigned long long)n;
          const double mean = sum / (double)(n - nanc);   // np.nanmean: NaN -> 0, divide by the non-NaN count
          if (mean == mean) {
            Rec cur; cur.ord = ord_of(mean); cur.pri = (unsigned long long)(~(uint32_t)fkey[f]) << 32; cur.a = (uint32_t)q; cur.b = (uint32_t)f;
            loc = rec_max(loc, cur);
          }
        }
      }
      TICK(10);                           // evaluate
      const Rec win = block_arg<1, NW>(loc, rslots, par);
      TICK(1);                            // argmax
      if (win.ord == 0ull || !(ord_to_double(win.ord) > tau)) break;   // :134 (nanmax of all-NaN is NaN -> stop)
      const int widx = (int)win.a;
      const int m = (int)win.b;
      const int mnode = cnode[m];
      const bool fast_upd = fast && (n + 1 <= LEAF);
      // --- slot owners: the one new correlation of this step, issued before anything else
      bool own = false;
      double v = 0.0;
      if (fast_upd && tid < nf && tid != widx) {
        const int rn = srow[tid];
        if (rn >= 0) { own = true; v = rat(rrow(rn), mnode); }
      }
      // --- bookkeeping warp: membership, frontier keys, new frontier cells and their slot state
      if (warp == BKW) {
        BTICK(0);                         // everything outside the update (evaluate, argmax, waiting)
        if (lane == 0) {
          lab[m] = (IT)k;
          hn[n] = (IT)mnode;
          s1_cells[base + n] = m;
          fkey[m] = NOKEY;
        }
        __syncwarp();
        bool fits = true;
        const int cnt = bk_add(m, n, n + 1, widx, nf, fast_upd, fits);
        if (lane == 0) { sh_i[1] = cnt; sh_i[2] = (fast_upd && fits) ? 1 : 0; }
        BTICK(4);
      }
      if (own) {                          // append element index n (value v) to the slot's list
        int nn = snan[tid];
        if (v != v) { v = 0.0; ++nn; snan[tid] = nn; }
        const int t = n & 7;
        double res;
        if (t == 7) {                     // a group of 8 is complete: it joins the accumulators
          double a[8];
          if (n == 7) {
#pragma unroll
            for (int q = 0; q < 7; ++q) a[q] = ftail[q * FCAP + tid];
            a[7] = v;
          } else {
#pragma unroll
            for (int q = 0; q < 7; ++q) a[q] = __dadd_rn(facc[q * FCAP + tid], ftail[q * FCAP + tid]);
            a[7] = __dadd_rn(facc[7 * FCAP + tid], v);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) facc[q * FCAP + tid] = a[q];
          res = __dadd_rn(__dadd_rn(__dadd_rn(a[0], a[1]), __dadd_rn(a[2], a[3])),
                          __dadd_rn(__dadd_rn(a[4], a[5]), __dadd_rn(a[6], a[7])));
        } else {
          ftail[t * FCAP + tid] = v;
          res = __dadd_rn(sres[tid], v);
        }
        sres[tid] = res;
        smean[tid] = res / (double)(n + 1 - nn);
        TICK(11);                         // owner path: gather latency + fold (thread 0's slot)
      }
      __syncthreads();
      nf = sh_i[1];
      fast = sh_i[2] != 0;
      ++n;
      ++n_steps;
      TICK(3);                            // frontier update + one gather per frontier cell + new-slot states
    }
    __syncthreads();
    for (int q = tid; q < nf; q += NT) { const int f = flist[q]; if (f >= 0) fkey[f] = NOKEY; }
    if (tid == 0) {
      S.a_start[k] = (IT)base; S.a_len[k] = (IT)n; S.seg_next[k] = -1; S.tail[k] = (IT)k; S.size[k] = (IT)n; S.fin[k] = 0;
      S.okey[k] = NOKEY64;
      S.xep[k] = 0; S.self_sz[k] = 0;
    }
    base += n;
    nA = k + 1;
    __syncthreads();
  }
  if (overflow) {
    if (tid == 0) status_all[b] = SIE_JOB_CAPACITY;
    continue;
  }

  // =============================================================== step 2 (:200-265)
  const long long clk1 = clock64();
  tick_last = clk1;
  // `taken` is now "belongs to a finalised area"; lab[] keeps tracking the current owner key.
  int cur_best = -1, nb = 0;     // best area whose lists are materialised in hn/hc
  const size_t RM = rm_cap(C);
  const int ld = dcap;           // row stride of D
  bool d_ok = false;             // dense block valid for the first nb member cells
  int epoch = 0, xtop = ld;      // cross-block column allocator (downwards from ld), restarted when the best area changes
  int self_top = 0;              // bump allocator of selfbuf
  // D: dense list-order copy of R for the current best area.  Row p = best member p.  Columns [0, nb): element
  // (p, q), q > p, = R[best_p][best_q].  Columns [xoff_k, xoff_k + |k|): the cross block R[best_p][k_q] of neighbour
  // area k, kept while the best area only grows, so a hypothetical merged row p is TWO contiguous runs of row p.
  // One 8-lane group per row; extend_D adds the pairs with from <= q < to.
  auto extend_D = [&](int from, int to) {
    d_ok = (to <= dcap) && (from == 0 || d_ok);
    if (!d_ok) return;
    for (int p = g; p < to - 1; p += NG) {
      const RowH row = rrow(hn[p]);
      double* drow = D + (size_t)p * ld;
      for (int q0 = max(from, p + 1); q0 < to; q0 += 64) {   // 8 gathers in flight per lane
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = min(q0 + 8 * u + j, to - 1);
          v[u] = rat(row, hn[q]);
        }
        sie_fence_regs<8>(v);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = q0 + 8 * u + j;
          if (q < to) drow[q] = v[u];
        }
      }
    }
    __syncthreads();
  };
  while (true) {
    // --- largest not-yet-final area, first key on ties (:207-212)
    Rec loc = rec_none();
    for (int k = tid; k < nA; k += NT) {
      const int sz = S.size[k];
      if (sz > 0) {             // still a key of V
        Rec cur; cur.ord = 1ull + (unsigned long long)(S.fin[k] ? 0 : sz); cur.pri = (unsigned long long)(~(uint32_t)k) << 32;
        cur.a = (uint32_t)k; cur.b = 0u;
        loc = rec_max(loc, cur);
      }
    }
    const Rec bw = block_arg<1, NW>(loc, rslots, par);
    if (bw.ord <= 1ull) break;   // no areas at all (:212) or all finalised
    const int best = (int)bw.a;
    ++n_rounds;
#ifdef SIE_AREA_PHASE_TIMERS
    const long long round_t0 = clock64();
    const int round_nb0 = S.size[best];
#endif
    if (best != cur_best) {                    // materialise V[best] in list order
      int off = 0;
      for (int s = best; s >= 0; s = S.seg_next[s]) {
        const int st = S.a_start[s], ln = S.a_len[s];
        for (int i = tid; i < ln; i += NT) { const int c = s1_cells[st + i]; hc[off + i] = (IT)c; hn[off + i] = cnode[c]; }
        off += ln;
      }
      nb = off;
      cur_best = best;
      ++epoch; xtop = ld;                      // every cross block belongs to the previous best area
      __syncthreads();
      extend_D(0, nb);
    }
    if (tid == 0) sh_i[3] = 0;
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
        if (old == NOKEY64) S.nlist[atomicAdd(&sh_i[3], 1)] = (IT)kk;
      }
    }
    __syncthreads();
    const int nn = sh_i[3];
    // --- dense path bookkeeping: every neighbour gets a column range of D for its cross block and a range of
    //     selfbuf for its own (best-independent) row means; if either does not fit the round uses the gather path
    if (tid == 0) {
      int ok = d_ok ? 1 : 0, top = xtop, stop = self_top;
      for (int q = 0; q < nn && ok; ++q) {
        const int kk = S.nlist[q], nk = S.size[kk];
        if (S.xep[kk] != epoch) {
          if (top - nk < nb) { ok = 0; break; }
          top -= nk; S.xoff[kk] = (IT)top; S.xep[kk] = (IT)epoch; S.xrows[kk] = 0;
        }
        if (S.self_sz[kk] != nk && S.self_sz[kk] != -nk) {
          if ((size_t)stop + nk > self_cap(C)) { ok = 0; break; }
          S.self_off[kk] = (IT)stop; stop += nk; S.self_sz[kk] = (IT)-nk;     // allocated, not yet computed
        }
      }
      sh_x[0] = top; sh_x[1] = ok; sh_x[2] = stop;
    }
    // correlations the reference consumes this round: n(n-1)/2 per neighbour (work counter only)
    for (int q = tid; q < nn; q += NT) { const unsigned long long n = (unsigned long long)(nb + S.size[S.nlist[q]]); wk += n * (n - 1) / 2; }
    __syncthreads();
    xtop = sh_x[0];
    self_top = sh_x[2];
    const bool dense = sh_x[1] != 0;
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
      const int nch = q1 - q0;
      // materialise the neighbours' node lists; sh_koff = node-list offsets, sh_uoff = row-unit offsets (gather path),
      // sh_goff / sh_soff = cross-block rows / self rows still to be filled (dense path)
      if (tid == 0) {
        int off = 0, uo = 0, go = 0, so = 0;
        for (int q = q0; q < q1; ++q) {
          const int kk = S.nlist[q], sz = S.size[kk];
          const int c = q - q0;
          sh_koff[c] = off; sh_uoff[c] = uo; sh_goff[c] = go; sh_soff[c] = so;
          off += sz; uo += nb + sz;
          if (dense) {
            sh_xo[c] = S.xoff[kk]; sh_r0[c] = S.xrows[kk]; sh_so[c] = S.self_off[kk];
            go += nb - S.xrows[kk];
            if (S.self_sz[kk] != sz) so += sz;
            S.xrows[kk] = (IT)nb;
          }
        }
        sh_koff[nch] = off; sh_uoff[nch] = uo; sh_goff[nch] = go; sh_soff[nch] = so;
      }
      __syncthreads();
      for (int q = q0; q < q1; ++q) {
        // dense rounds only need a neighbour's node list while its cross block / self means are being filled
        if (dense && sh_goff[q - q0 + 1] == sh_goff[q - q0] && sh_soff[q - q0 + 1] == sh_soff[q - q0]) continue;
        int off = sh_koff[q - q0];
        for (int s = S.nlist[q]; s >= 0; s = S.seg_next[s]) {
          const int st = S.a_start[s], ln = S.a_len[s];
          for (int i = tid; i < ln; i += NT) knl[off + i] = cnode[s1_cells[st + i]];
          off += ln;
        }
      }
      __syncthreads();
      TICK(6);                            // neighbour node lists
      if (dense) {
        // ---- (a) fill what is missing: cross-block rows [r0, nb) of every neighbour, one 8-lane group per row ...
        const int G = sh_goff[nch];
        for (int u = g; u < G; u += NG) {
          int qc = 0;
          while (u >= sh_goff[qc + 1]) ++qc;
          const int p = sh_r0[qc] + (u - sh_goff[qc]);
          const int nk = sh_koff[qc + 1] - sh_koff[qc];
          const IT* kn = knl + sh_koff[qc];
          const RowH row = rrow(hn[p]);
          double* dst = D + (size_t)p * ld + sh_xo[qc];
          for (int qq = 0; qq < nk; qq += 64) {
            double v[8];
#pragma unroll
            for (int u2 = 0; u2 < 8; ++u2) v[u2] = rat(row, kn[min(qq + 8 * u2 + j, nk - 1)]);
            sie_fence_regs<8>(v);
#pragma unroll
            for (int u2 = 0; u2 < 8; ++u2) { const int q = qq + 8 * u2 + j; if (q < nk) dst[q] = v[u2]; }
          }
        }
        // ---- ... and the neighbours' own row means r_{nb+p'} = nanmean(R[k_p', k_q'], q' > p'), which do not depend
        //      on the best area: computed once per area, reused every round it is a neighbour
        const int SU = sh_soff[nch];
        for (int u = g; u < SU; u += NG) {
          int qc = 0;
          while (u >= sh_soff[qc + 1]) ++qc;
          const int pp = u - sh_soff[qc];
          const int nk = sh_koff[qc + 1] - sh_koff[qc];
          const IT* kn = knl + sh_koff[qc];
          const int len = nk - 1 - pp;
          int nanc = 0;
          double sum = 0.0;
          if (len > 0) {
            sum = sie_pw_sum_gather<MAXD, IT, ZR>(rcx, kn[pp], kn + pp + 1, len, kn, len, j, gmask, &nanc);
          }
          if (j == 0) selfbuf[sh_so[qc] + pp] = (len - nanc > 0) ? sum / (double)(len - nanc) : sie_nan();   // nanmean([]) = nan
        }
        __syncthreads();
        if (tid == 0) for (int q = q0; q < q1; ++q) { const int kk = S.nlist[q]; S.self_sz[kk] = S.size[kk]; }
        TICK(9);                          // cross blocks + self row means (first sight of a neighbour)
        // ---- (b) row means of the best rows: r_p = nanmean(D[p][p+1..nb) ++ D[p][xoff..xoff+nk)), two contiguous runs
        const int total = nch * nb;
        for (int u = g; u < total; u += NG) {
          const int qc = u / nb;
          const int p = u - qc * nb;
          const int nk = sh_koff[qc + 1] - sh_koff[qc];
          const int nbb = nb - 1 - p;
          const int len = nbb + nk;
          const double* base1 = D + (size_t)p * ld + (p + 1);
          const int gap = sh_xo[qc] - nb;           // element i >= nbb lives at base1[i + gap]
          int nanc = 0;
          const double sum = sie_pw_sum_runs<MAXD>(base1, nbb, base1 + nbb + gap, len, j, gmask, &nanc);
          if (j == 0) rowmean[u] = (len - nanc > 0) ? sum / (double)(len - nanc) : sie_nan();
        }
        __syncthreads();
        TICK(7);                          // row means of the hypothetical areas
        // ---- (c) stat_k = nanmean(r_0..r_{n-1})  (:253): best rows from rowmean, the neighbour's own from selfbuf
        for (int q = q0 + g; q < q1; q += NG) {
          const int qc = q - q0;
          const int nk = sh_koff[qc + 1] - sh_koff[qc];
          const int n = nb + nk;
          int nanc = 0;
          const double sum = sie_pw_sum_runs<MAXD>(rowmean + (size_t)qc * nb, nb, selfbuf + sh_so[qc], n, j, gmask, &nanc);
          if (j == 0) S.stat[S.nlist[q]] = (n - nanc > 0) ? sum / (double)(n - nanc) : sie_nan();
        }
        __syncthreads();
        TICK(8);                          // statistic per neighbour
      } else {
        // ---- gather path (best area or neighbours too large for the dense block): one 8-lane group per
        //      (neighbour, row), everything gathered from R
        const int total = sh_uoff[nch];
        for (int u = g; u < total; u += NG) {
          int qc = 0;
          while (u >= sh_uoff[qc + 1]) ++qc;
          const int p = u - sh_uoff[qc];
          const int nk = sh_koff[qc + 1] - sh_koff[qc];
          const int n = nb + nk;
          const IT* kn = knl + sh_koff[qc];
          const int len = n - 1 - p;
          int nanc = 0;
          double sum = 0.0;
          if (len > 0) {
            // row p of the hypothetical area best ++ k: a best row continues into k's cells, a row of k stays in k
            const bool in_best = p < nb;
            const int nbb = in_best ? nb - 1 - p : 0;  // elements of the row inside the best area
            const IT* kq = in_best ? kn : kn + (p - nb) + 1;
            sum = sie_pw_sum_gather<MAXD, IT, ZR>(rcx, in_best ? hn[p] : kn[p - nb], hn + (p + 1), nbb, kq, len, j, gmask,
                                                  &nanc);
          }
          if (j == 0) rowmean[u] = (len - nanc > 0) ? sum / (double)(len - nanc) : sie_nan();   // nanmean([]) = nan
        }
        __syncthreads();
        TICK(7);                          // row means of the hypothetical areas
        // stat_k = nanmean(r_0..r_{n-1})  (:253) -- one 8-lane group per neighbour
        for (int q = q0 + g; q < q1; q += NG) {
          const int n = sh_uoff[q - q0 + 1] - sh_uoff[q - q0];
          const double* rm = rowmean + sh_uoff[q - q0];
          int nanc = 0;
          const double sum = sie_pw_sum_runs<MAXD>(rm, n, rm, n, j, gmask, &nanc);
          if (j == 0) S.stat[S.nlist[q]] = (n - nanc > 0) ? sum / (double)(n - nanc) : sie_nan();
        }
        __syncthreads();
        TICK(8);                          // statistic per neighbour
      }
      q0 = q1;
    }
    // --- max(Anei_Rs.items(), key=itemgetter(1)) (:255): first in discovery order wins ties; a NaN in
    //     first position is never displaced (list comparison semantics)
    Rec loc2 = rec_none(), f1 = rec_none();
    for (int q = tid; q < nn; q += NT) {
      const int kk = S.nlist[q];
      const double st = S.stat[kk];
      const unsigned long long ok = S.okey[kk];
      Rec c1; c1.ord = 1ull; c1.pri = ~ok; c1.a = (uint32_t)kk; c1.b = 0u;   // first discovered = smallest discovery key
      f1 = rec_max(f1, c1);
      if (st == st) {
        Rec cur; cur.ord = ord_of(st); cur.pri = ~ok; cur.a = (uint32_t)kk; cur.b = 0u;
        loc2 = rec_max(loc2, cur);
      }
    }
    const Rec firstn = block_arg<2, NW>(f1, rslots, par);
    const Rec win = block_arg<2, NW>(loc2, rslots, par);
    bool merge = false;
    int kk = -1;
    if (nn > 0 && win.ord != 0ull) {
      const double first_stat = S.stat[(int)firstn.a];
      if (first_stat == first_stat && ord_to_double(win.ord) > tau) { merge = true; kk = (int)win.a; }
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
          const int c = s1_cells[st + i];
          hc[off + i] = (IT)c; hn[off + i] = cnode[c]; lab[c] = (IT)best;
        }
        off += ln;
      }
      __syncthreads();
      if (tid == 0) {
        S.seg_next[S.tail[best]] = (IT)kk;
        S.tail[best] = S.tail[kk];
        S.size[best] = (IT)(S.size[best] + S.size[kk]);
        S.size[kk] = 0;
      }
      if (off > xtop) { ++epoch; xtop = ld; }  // the best columns would run into the cross blocks: drop them
      extend_D(nb, off);
      nb = off;
    } else {
      if (tid == 0) S.fin[best] = 1;           // :262-265 all cells of V[best] become unavailable
    }
    __syncthreads();
    TICK(12);                             // winner, merge / finalise, dense block extension
#ifdef SIE_AREA_PHASE_TIMERS
    if (tid == 0) {    // rounds bucketed by the size of the best area: (count << 44) | cycles, work[26..31]
      const int bk = round_nb0 <= 4 ? 0 : round_nb0 <= 8 ? 1 : round_nb0 <= 16 ? 2 : round_nb0 <= 32 ? 3 : round_nb0 <= 64 ? 4 : 5;
      rbucket[bk] += (1ull << 44) + (unsigned long long)(clock64() - round_t0);
    }
#endif
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
    work_all[SIE_AREA_WORK * b + 24] = ph[11];
    work_all[SIE_AREA_WORK * b + 25] = ph[12];
#ifdef SIE_AREA_PHASE_TIMERS
    for (int i = 0; i < 6; ++i) work_all[SIE_AREA_WORK * b + 26 + i] = rbucket[i];
#endif
    work_all[SIE_AREA_WORK * b + 15] = n_slow;
  }
  if (tid == BKW * 32 && work_all)
    for (int i = 0; i < 8; ++i) work_all[SIE_AREA_WORK * b + 16 + i] = bph[i];
  if (tid == 0) {
    int cnt = 0, off = 0;
    for (int k = 0; k < nA; ++k) {
      if (S.size[k] > 0) {
        out_key[cnt] = k;
        out_start[cnt] = off;
        S.nlist[cnt] = (IT)k;
        off += S.size[k];
        ++cnt;
      }
    }
    out_start[cnt] = off;
    n_areas_all[b] = cnt;
    status_all[b] = (cnt < 2) ? SIE_JOB_FEW_AREAS : SIE_JOB_OK;   // :212 / :278 ValueError
    sh_i[4] = cnt;
  }
  __syncthreads();
  const int cnt = sh_i[4];
  for (int a = 0; a < cnt; ++a) {
    int off = out_start[a];
    for (int s = S.nlist[a]; s >= 0; s = S.seg_next[s]) {
      const int st = S.a_start[s], ln = S.a_len[s];
      for (int i = tid; i < ln; i += NT) {
        const int c = s1_cells[st + i];
        out_cells[off + i] = c;
        out_label[c] = a;
      }
      off += ln;
    }
  }
  }   // persistent loop over the job queue
}

template <class CFG, bool ONCHIP, bool ZR>
int launch_area(const SieDevice* dev, int slot, int grid, size_t smem, cudaStream_t st, const double* R,
                const int32_t* job_T, int Tp, const double* stencil, const int32_t* node_cell, const int32_t* cell_node, const int32_t* n_nodes,
                const double* tau, const int32_t* first_nan_cell, int B, int X, int Y, int ldn, int latlon,
                int max_areas, int32_t* area_cells, int32_t* area_start, int32_t* area_key, int32_t* n_areas,
                int32_t* label, int32_t* status, int* queue, unsigned char* scratch, size_t per_cta, int place,
                unsigned long long* work) {
  auto kern = k_area_level<CFG, ONCHIP, ZR>;
  if (int rc = sie_ensure_smem(dev, slot, (const void*)kern, smem)) return rc;
  kern<<<grid, CFG::NT, smem, st>>>(R, job_T, Tp, stencil, node_cell, cell_node, n_nodes, tau, first_nan_cell, B, X, Y, ldn, latlon,
                                    max_areas, area_cells, area_start, area_key, n_areas, label, status, queue, scratch,
                                    per_cta, place, work);
  return SIE_OK;
}

// persistent grid: CTAs per SM of the variant this (grid, capacity) selects, and its dynamic shared memory
struct AreaPlan { int variant, ctas_per_sm, place; size_t smem; };   // variant 0: <256,i16>, 1: <512,i16>, 2: <512,i32>
AreaPlan plan_area(const SieDevice* dev, int B, int C, int ldn, int max_areas, bool force32 = false) {
  AreaPlan pl = {2, 1, 0, 0};
  const size_t budget = (size_t)dev->max_smem_optin - 4096;    // static shared memory + alignment slack
  if (!force32 && C < 8192 && ldn <= 32767 && max_areas <= 32766) {
    const size_t i16 = 7 * align_up(sizeof(int16_t) * (size_t)C, 16) + area_tab_bytes(max_areas, sizeof(int16_t)) + 64;
    const size_t s256 = align_up(slot_bytes(AreaCfg<256, int16_t>::FCAP), 16) + i16;
    const size_t s512 = align_up(slot_bytes(AreaCfg<512, int16_t>::FCAP), 16) + i16;
    // two 256-thread CTAs per SM against one 512-thread CTA: a job takes 1.3x longer with half the threads and 1.64x
    // when it shares its SM (tools/prof_area.py), so the pair wins whenever it saves a wave of the persistent grid:
    // waves x job time, in units of the 512-thread job
    const int sms = dev->sm_count;
    const double t512 = (double)((B + sms - 1) / sms);
    const double t256 = (double)((B + 2 * sms - 1) / (2 * sms)) * (B > sms ? 1.64 : 1.3);
    if (t256 < t512 && 2 * (s256 + 4096 + 1024) <= (size_t)dev->smem_per_sm) {
      pl.variant = 0; pl.ctas_per_sm = 2; pl.smem = s256; return pl;
    }
    if (s512 <= budget) { pl.variant = 1; pl.smem = s512; return pl; }
  }
  // 32-bit indices.  Shared-memory placement: slot state always on chip; then evict arrays to global scratch, least
  // latency-critical first, until the rest fits.
  const size_t ibytes = align_up(sizeof(int32_t) * (size_t)C, 16);
  const size_t tabs = area_tab_bytes(max_areas, sizeof(int32_t));
  size_t smem = align_up(slot_bytes(AreaCfg<512, int32_t>::FCAP), 16) + 64 + tabs + 7 * ibytes;
  const int order[8] = {PL_KNL, PL_AREA, PL_HC, PL_FLIST, PL_HN, PL_FKEY, PL_CNL, PL_LAB};
  for (int i = 0; i < 8 && smem > budget; ++i) {
    pl.place |= order[i];
    smem -= (order[i] == PL_AREA) ? tabs : ibytes;
  }
  pl.smem = smem;
  return pl;
}

}  // namespace

extern "C" size_t sie_area_level_scratch_bytes(int B, int C) {
  // one scratch block per CTA of the persistent grid (at most two CTAs per SM), not per job; per-area arrays are sized
  // for the worst case max_areas = C/2 + 1
  const SieDevice* dev = sie_device();
  const int sms = (dev && dev->sm_count > 0) ? dev->sm_count : 148;
  const int g = B < 2 * sms ? (B > 0 ? B : 1) : 2 * sms;
  return QUEUE_BYTES + (size_t)g * scratch_per_cta(C, C / 2 + 1);
}

extern "C" int sie_area_level(const double* R, const double* z, const int32_t* job_T, int Tp, const double* stencil,
                              const int32_t* node_cell, const int32_t* cell_node, const int32_t* n_nodes,
                              const double* tau, const int32_t* first_nan_cell, int B, int X, int Y, int ldn, int latlon,
                              int max_areas, int32_t* area_cells, int32_t* area_start, int32_t* area_key,
                              int32_t* n_areas, int32_t* label, int32_t* status, void* scratch,
                              size_t scratch_bytes, uint64_t* work, void* stream) {
  SIE_CHECK_ARG(stencil && node_cell && cell_node && n_nodes && tau && first_nan_cell && area_cells &&
                    area_start && area_key && n_areas && label && status && scratch, "null pointer");
  SIE_CHECK_ARG(R || (z && job_T && Tp >= 4 && (Tp % 4) == 0), "without R the unit-norm rows z, job_T and Tp are needed");
  SIE_CHECK_ARG(B > 0 && X > 0 && Y > 0 && ldn > 0 && max_areas > 0, "non-positive size");
  const int C = X * Y;
  SIE_CHECK_ARG(max_areas <= C / 2 + 1, "max_areas cannot exceed C/2+1");
  SIE_CHECK_ARG((long long)C < (1ll << 22), "grid too large (frontier key encoding / pairwise tree depth)");
  const SieDevice* dev = sie_device();
  if (!dev) return SIE_ERR_LAUNCH;
  const bool zr = (R == nullptr);      // correlations recomputed from z (the matrix is never materialised)
  const AreaPlan pl = plan_area(dev, B, C, ldn, max_areas, zr);
  SIE_CHECK_ARG(pl.smem <= (size_t)dev->max_smem_optin - 4096, "shared memory budget");
  const size_t per_cta = scratch_per_cta(C, max_areas);
  SIE_CHECK_ARG(scratch_bytes >= QUEUE_BYTES + per_cta, "scratch too small");
  long long grid = (long long)pl.ctas_per_sm * dev->sm_count;
  if (grid > B) grid = B;
  const long long fit = (long long)((scratch_bytes - QUEUE_BYTES) / per_cta);
  if (grid > fit) grid = fit;
  cudaStream_t st = (cudaStream_t)stream;
  int* queue = reinterpret_cast<int*>(scratch);
  unsigned char* blocks = reinterpret_cast<unsigned char*>(scratch) + QUEUE_BYTES;
  if (cudaMemsetAsync(queue, 0, sizeof(int), st) != cudaSuccess) SIE_CHECK_LAUNCH();
  int rc;
#define SIE_AREA_ARGS                                                                                                   \
  job_T, Tp, stencil, node_cell, cell_node, n_nodes, tau, first_nan_cell, B, X, Y, ldn, latlon, max_areas, area_cells,   \
      area_start, area_key, n_areas, label, status, queue, blocks, per_cta, pl.place, (unsigned long long*)work
  const bool on32 = (pl.place == 0 && C < 8192);
  if (zr && on32)
    rc = launch_area<AreaCfg<512, int32_t>, true, true>(dev, SIE_K_AREA_ON32_Z, (int)grid, pl.smem, st, z, SIE_AREA_ARGS);
  else if (zr)
    rc = launch_area<AreaCfg<512, int32_t>, false, true>(dev, SIE_K_AREA_OFF32_Z, (int)grid, pl.smem, st, z, SIE_AREA_ARGS);
  else if (pl.variant == 0)
    rc = launch_area<AreaCfg<256, int16_t>, true, false>(dev, SIE_K_AREA_ON16_2, (int)grid, pl.smem, st, R, SIE_AREA_ARGS);
  else if (pl.variant == 1)
    rc = launch_area<AreaCfg<512, int16_t>, true, false>(dev, SIE_K_AREA_ON16, (int)grid, pl.smem, st, R, SIE_AREA_ARGS);
  else if (on32)
    rc = launch_area<AreaCfg<512, int32_t>, true, false>(dev, SIE_K_AREA_ON32, (int)grid, pl.smem, st, R, SIE_AREA_ARGS);
  else
    rc = launch_area<AreaCfg<512, int32_t>, false, false>(dev, SIE_K_AREA_OFF32, (int)grid, pl.smem, st, R, SIE_AREA_ARGS);
#undef SIE_AREA_ARGS
  if (rc) return rc;
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}
