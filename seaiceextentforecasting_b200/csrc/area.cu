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
// Latency/gather bound: no roofline fraction is claimed for this kernel; see DESIGN.md.
#include "common.cuh"

namespace {

constexpr int NT = 512;
constexpr int NG = NT / 8;          // 8-lane groups per CTA
constexpr uint32_t NOKEY = 0xffffffffu;
constexpr unsigned long long NOKEY64 = ~0ull;

struct AreaScratch {   // per-job global scratch (byte offsets computed on host and device the same way)
  int32_t* s1_cells;   // [C] step-1 member cells, area after area
  int32_t* ibuf;       // [5*C] fallback for the shared-memory integer arrays
  int32_t* nbuf;       // [C] neighbour-area node lists (step 2)
  double* rowmean;     // [RM] row means of hypothetical areas
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
__host__ __device__ inline size_t scratch_per_job(int C, int MA) {
  size_t s = 0;
  s += align_up(sizeof(int32_t) * (size_t)C, 256);          // s1_cells
  s += align_up(sizeof(int32_t) * (size_t)5 * C, 256);      // ibuf
  s += align_up(sizeof(int32_t) * (size_t)C, 256);          // nbuf
  s += align_up(sizeof(double) * rm_cap(C), 256);           // rowmean
  s += 8 * align_up(sizeof(int32_t) * (size_t)(MA + 1), 256);
  s += 2 * align_up(sizeof(double) * (size_t)MA, 256);
  return s;
}
__device__ inline AreaScratch carve(unsigned char* p, int C, int MA) {
  AreaScratch s;
  auto take = [&](size_t bytes) { unsigned char* r = p; p += align_up(bytes, 256); return r; };
  s.s1_cells = (int32_t*)take(sizeof(int32_t) * (size_t)C);
  s.ibuf = (int32_t*)take(sizeof(int32_t) * (size_t)5 * C);
  s.nbuf = (int32_t*)take(sizeof(int32_t) * (size_t)C);
  s.rowmean = (double*)take(sizeof(double) * rm_cap(C));
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

struct Best {   // argmax record: larger mean wins, ties -> smaller key (earlier in the reference's list)
  double mean;
  unsigned long long key;
  int idx;
};
__device__ __forceinline__ bool better(const Best& a, const Best& b) {   // is a better than b
  if (a.idx < 0) return false;
  if (b.idx < 0) return true;
  return a.mean > b.mean || (a.mean == b.mean && a.key < b.key);
}
__device__ __forceinline__ Best shfl_best(const Best& v, int o) {
  Best r;
  r.mean = __shfl_xor_sync(0xffffffffu, v.mean, o);
  r.key = __shfl_xor_sync(0xffffffffu, v.key, o);
  r.idx = __shfl_xor_sync(0xffffffffu, v.idx, o);
  return r;
}
// CTA-wide argmax; every thread returns the same winner.  `slots` is shared scratch of NT/32 records.
__device__ __forceinline__ Best block_best(Best v, Best* slots) {
#pragma unroll
  for (int o = 8; o < 32; o <<= 1) {   // lanes inside an 8-group already agree
    Best w = shfl_best(v, o);
    if (better(w, v)) v = w;
  }
  __syncthreads();                     // slots may still be read from the previous call
  if ((threadIdx.x & 31) == 0) slots[threadIdx.x >> 5] = v;
  __syncthreads();
  Best r = slots[0];
  for (int w = 1; w < NT / 32; ++w)
    if (better(slots[w], r)) r = slots[w];
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
  __shared__ Best slots[NT / 32];
  __shared__ int sh_i[8];
  __shared__ unsigned long long sh_work;
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
  const int32_t* cnode = cell_node_all + (size_t)b * C;
  int32_t* out_cells = area_cells_all + (size_t)b * C;
  int32_t* out_start = area_start_all + (size_t)b * (MA + 1);
  int32_t* out_key = area_key_all + (size_t)b * MA;
  int32_t* out_label = label_all + (size_t)b * C;

  AreaScratch S = carve(scratch_all + (size_t)b * scratch_stride, C, MA);
  int32_t* ib = use_smem ? smem_i : S.ibuf;
  int32_t* lab = ib;                      // [C] area key of each cell, -1 = unassigned
  uint32_t* fkey = (uint32_t*)(ib + C);   // [C] frontier key (step 1)
  int32_t* flist = ib + 2 * C;            // [C] frontier cells (step 1)
  int32_t* hn = ib + 3 * C;               // [C] node list of the current area (step 1) / best area (step 2)
  int32_t* hc = ib + 4 * C;               // [C] cell list of the best area (step 2)

  for (int c = tid; c < C; c += NT) { lab[c] = -1; fkey[c] = NOKEY; out_label[c] = -1; }
  if (tid == 0) { n_areas_all[b] = 0; out_start[0] = 0; sh_work = 0ull; if (work_all) work_all[4 * b] = 0ull; }
  __syncthreads();
  if (first_nan_cell[b] < 0) {            // :50-51 IndexError in the reference
    if (tid == 0) status_all[b] = SIE_JOB_NO_NAN_CELL;
    return;
  }
  if (status_all[b] == SIE_JOB_CAPACITY) return;   // K1 already flagged this job

  // =============================================================== step 1 (:154-196)
  const long long clk0 = clock64();
  unsigned long long n_steps = 0, n_rounds = 0;
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
    Best cand;
    cand.idx = (dir >= 0) ? c : -1;
    cand.mean = 0.0;
    cand.key = (unsigned long long)(unsigned)c;        // earliest cell wins
    // lanes of an 8-group hold different cells here: reduce inside the group first
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      Best w = shfl_best(cand, o);
      if (better(w, cand)) cand = w;
    }
    // carry the chosen direction alongside (recomputed by the owner thread below)
    Best win = block_best(cand, slots);
    if (win.idx < 0) { c0 += NT; continue; }
    const int seed = win.idx;
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
    if (tid == 0) {
      lab[seed] = k; lab[nbr] = k;
      hn[0] = cnode[seed]; hn[1] = cnode[nbr];
      S.s1_cells[base] = seed; S.s1_cells[base + 1] = nbr;
      int cnt = 0;
      for (int p = 0; p < 2; ++p) {
        const int cc = (p == 0) ? seed : nbr;
        const int ci = cc / Y, cj = cc - ci * Y;
        for (int d = 0; d < 4; ++d) {     // gen_area_neighbours :80-94 (no wrap here)
          const int a = ci + (d == 0 ? -1 : (d == 1 ? 1 : 0));
          const int q = cj + (d == 2 ? -1 : (d == 3 ? 1 : 0));
          if (a < 0 || a >= X || q < 0 || q >= Y) continue;
          const int f = a * Y + q;
          const int fn = cnode[f];
          if (lab[f] >= 0 || fn < 0 || fn >= N) continue;
          const uint32_t key = ((uint32_t)d << 28) | (uint32_t)p;
          if (fkey[f] == NOKEY) flist[cnt++] = f;
          if (key < fkey[f]) fkey[f] = key;
        }
      }
      sh_i[1] = cnt;
    }
    __syncthreads();
    nf = sh_i[1];
    while (nf > 0) {
      Best loc; loc.idx = -1; loc.mean = 0.0; loc.key = 0;
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
          Best cur; cur.mean = mean; cur.key = fkey[f]; cur.idx = q;
          if (better(cur, loc)) loc = cur;
        }
      }
      const Best win2 = block_best(loc, slots);
      if (win2.idx < 0 || !(win2.mean > tau)) break;     // :134 (nanmax of all-NaN is NaN -> stop)
      if (tid == 0) {
        const int m = flist[win2.idx];
        lab[m] = k;
        hn[n] = cnode[m];
        S.s1_cells[base + n] = m;
        fkey[m] = NOKEY;
        int cnt = nf - 1;
        flist[win2.idx] = flist[cnt];
        const int ci = m / Y, cj = m - ci * Y;
        for (int d = 0; d < 4; ++d) {
          const int a = ci + (d == 0 ? -1 : (d == 1 ? 1 : 0));
          const int q = cj + (d == 2 ? -1 : (d == 3 ? 1 : 0));
          if (a < 0 || a >= X || q < 0 || q >= Y) continue;
          const int f = a * Y + q;
          const int fn = cnode[f];
          if (lab[f] >= 0 || fn < 0 || fn >= N) continue;
          const uint32_t key = ((uint32_t)d << 28) | (uint32_t)n;
          if (fkey[f] == NOKEY) flist[cnt++] = f;
          if (key < fkey[f]) fkey[f] = key;
        }
        sh_i[1] = cnt;
      }
      __syncthreads();
      nf = sh_i[1];
      ++n;
      ++n_steps;
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
  // `taken` is now "belongs to a finalised area"; lab[] keeps tracking the current owner key.
  int cur_best = -1, nb = 0;     // best area whose lists are materialised in hn/hc
  const size_t RM = rm_cap(C);
  while (true) {
    // --- largest not-yet-final area, first key on ties (:207-212)
    Best loc; loc.idx = -1; loc.mean = 0.0; loc.key = 0;
    for (int k = tid; k < nA; k += NT) {
      const int sz = S.size[k];
      if (sz > 0) {             // still a key of V
        Best cur; cur.mean = S.fin[k] ? 0.0 : (double)sz; cur.key = (unsigned long long)k; cur.idx = k;
        if (better(cur, loc)) loc = cur;
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      Best w = shfl_best(loc, o);
      if (better(w, loc)) loc = w;
    }
    const Best bw = block_best(loc, slots);
    if (bw.idx < 0 || bw.mean == 0.0) break;   // no areas at all (ValueError at :212) or all finalised
    const int best = bw.idx;
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
    }
    if (tid == 0) sh_i[2] = 0;
    __syncthreads();
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
    // --- hypothetical merges (:224-253), neighbours processed in chunks that fit the row-mean buffer
    int q0 = 0;
    while (q0 < nn) {
      // chunk = neighbours q0..q1-1 ; every thread computes the same partition
      int q1 = q0;
      size_t rows = 0, cells = 0;
      while (q1 < nn) {
        const int sz = S.size[S.nlist[q1]];
        if (q1 > q0 && (rows + nb + sz > RM || cells + sz > (size_t)C)) break;
        rows += nb + sz; cells += sz; ++q1;
      }
      // materialise the neighbours' node lists
      if (tid == 0) {
        int off = 0;
        for (int q = q0; q < q1; ++q) { S.noff[q] = off; off += S.size[S.nlist[q]]; }
        S.noff[q1] = off;
      }
      __syncthreads();
      for (int q = q0; q < q1; ++q) {
        int off = S.noff[q];
        for (int s = S.nlist[q]; s >= 0; s = S.seg_next[s]) {
          const int st = S.a_start[s], ln = S.a_len[s];
          for (int i = tid; i < ln; i += NT) S.nbuf[off + i] = cnode[S.s1_cells[st + i]];
          off += ln;
        }
      }
      __syncthreads();
      // row means r_p = nanmean(R[hyp_p, hyp_q], q > p); unit u = (neighbour q, row p)
      {
        size_t ubase = 0;
        for (int q = q0; q < q1; ++q) {
          const int nk = S.noff[q + 1] - S.noff[q];
          const int n = nb + nk;
          const int32_t* kn = S.nbuf + S.noff[q];
          double* rm = S.rowmean + ubase;
          for (int p = g; p < n; p += NG) {
            const int len = n - 1 - p;
            const int rownode = (p < nb) ? hn[p] : kn[p - nb];
            const double* row = R + (size_t)rownode * ldn;
            int nanc = 0;
            double sum = 0.0;
            if (len > 0) {
              if (j == 0) wk += (unsigned long long)len;
              sum = sie_pw_sum8(
                  [&](int i) { const int t = p + 1 + i; return __ldg(row + ((t < nb) ? hn[t] : kn[t - nb])); },
                  len, j, gmask, nanc);
              nanc += __shfl_xor_sync(gmask, nanc, 1);
              nanc += __shfl_xor_sync(gmask, nanc, 2);
              nanc += __shfl_xor_sync(gmask, nanc, 4);
            }
            if (j == 0) rm[p] = (len - nanc > 0) ? sum / (double)(len - nanc) : sie_nan();   // nanmean([]) = nan
          }
          ubase += n;
        }
      }
      __syncthreads();
      // stat_k = nanmean(r_0..r_{n-1})  (:253) -- one 8-lane group per neighbour
      {
        size_t ubase = 0;
        for (int q = q0; q < q1; ++q) {
          const int n = nb + S.noff[q + 1] - S.noff[q];
          if (g == (q - q0) % NG) {
            const double* rm = S.rowmean + ubase;
            int nanc = 0;
            const double sum = sie_pw_sum8([&](int i) { return rm[i]; }, n, j, gmask, nanc);
            nanc += __shfl_xor_sync(gmask, nanc, 1);
            nanc += __shfl_xor_sync(gmask, nanc, 2);
            nanc += __shfl_xor_sync(gmask, nanc, 4);
            if (j == 0) S.stat[S.nlist[q]] = (n - nanc > 0) ? sum / (double)(n - nanc) : sie_nan();
          }
          ubase += n;
        }
      }
      __syncthreads();
      q0 = q1;
    }
    // --- max(Anei_Rs.items(), key=itemgetter(1)) (:255): first in discovery order wins ties; a NaN in
    //     first position is never displaced (list comparison semantics)
    Best loc2; loc2.idx = -1; loc2.mean = 0.0; loc2.key = 0;
    unsigned long long first_key = NOKEY64; int first_idx = -1;
    for (int q = tid; q < nn; q += NT) {
      const int kk = S.nlist[q];
      const double st = S.stat[kk];
      const unsigned long long ok = S.okey[kk];
      if (ok < first_key) { first_key = ok; first_idx = kk; }
      if (st == st) {
        Best cur; cur.mean = st; cur.key = ok; cur.idx = kk;
        if (better(cur, loc2)) loc2 = cur;
      }
    }
    // first discovered neighbour (min okey) -- reuse the argmax machinery with mean fixed
    Best f1; f1.idx = first_idx; f1.mean = 0.0; f1.key = first_key;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      Best w = shfl_best(loc2, o);
      if (better(w, loc2)) loc2 = w;
      Best w1 = shfl_best(f1, o);
      if (better(w1, f1)) f1 = w1;
    }
    const Best firstn = block_best(f1, slots);
    const Best win = block_best(loc2, slots);
    bool merge = false;
    int kk = -1;
    if (nn > 0 && win.idx >= 0) {
      const double first_stat = S.stat[firstn.idx];
      if (first_stat == first_stat && win.mean > tau) { merge = true; kk = win.idx; }
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
      nb = off;
    } else {
      if (tid == 0) S.fin[best] = 1;           // :262-265 all cells of V[best] become unavailable
    }
    __syncthreads();
  }

  // =============================================================== output in dict order (ascending key)
  if (wk) atomicAdd(&sh_work, wk);
  __syncthreads();
  if (tid == 0 && work_all) {
    const long long clk2 = clock64();
    work_all[4 * b] = sh_work;                                  // correlations consumed
    work_all[4 * b + 1] = (unsigned long long)(clk1 - clk0);    // SM cycles in step 1
    work_all[4 * b + 2] = (unsigned long long)(clk2 - clk1);    // SM cycles in step 2
    work_all[4 * b + 3] = (n_steps << 32) | n_rounds;           // growth steps, merge rounds
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
  size_t smem = sizeof(int32_t) * (size_t)5 * C;
  int use_smem = (smem + 4096 <= (size_t)max_optin) ? 1 : 0;
  if (!use_smem) smem = 0;
  cudaFuncSetAttribute(k_area_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_area_level<<<B, NT, smem, (cudaStream_t)stream>>>(
      R, stencil, node_cell, cell_node, n_nodes, tau, first_nan_cell, X, Y, ldn, latlon, max_areas, area_cells,
      area_start, area_key, n_areas, label, status, (unsigned char*)scratch, per_job, use_smem,
      (unsigned long long*)work);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}
