// K2: all-pairs Pearson correlation R = Z Z^T (+ significance threshold tau) and K3: 4-neighbour stencil.
// Reference: np.corrcoef + fill_diagonal + t-test mean, ComplexNetworks.py:34-35,:41-47; seed-search
// gathers ComplexNetworks.py:166-172.
//
// Z rows are centred and scaled to unit norm by K1, so R_ij is a plain dot product over K = T <= Tp.
// The contraction runs on the FP64 tensor pipe (mma.sync m8n8k4 -> SASS DMMA.8x8x4; tcgen05 has no f64
// kind, and ptxas splits the larger sm_90 f64 shapes into 8x8x4 on sm_100a).  A 128-row panel of Z is one
// contiguous 128*Tp*8-byte run, so operands are staged with 1-D bulk async copies (cp.async.bulk -> SASS
// UBLKCP, the TMA engine) completing on an mbarrier; Tp = 4 (mod 8) makes the fragment loads
// bank-conflict free without swizzling.  Every kernel walks a precomputed table of 128x64 tiles on or right of the
// diagonal blocks; only those tiles are computed, and only the UPPER TRIANGLE of R is stored: every reader (K3, K4/K5,
// the host accessors) addresses R[min(i,j)][max(i,j)], so the matrix is symmetric by construction like numpy's
// syrk-based corrcoef and K2 writes 4 N^2 instead of 8 N^2 bytes.  Three kernels share the table:
//   k_corr_tma   (default when R is stored) finished sub-tiles leave as tensor-map bulk stores (TMA), see below;
//   k_corr_rows  (default for the tau-only pass) row-resident A panel, producer warp + 16 consumer warps;
//   k_corr_tiles (fallback for windows too long for the others' shared memory) persistent 256-thread CTAs, two per SM;
//                off-diagonal tiles are staged in shared memory (over the operand panels) and written as whole rows.
// Algorithmic work: N(N+1)T flop per network (upper triangle), 4 N(N+1) bytes if R is stored.
#include <cuda.h>
#include "common.cuh"

namespace {

constexpr int TILE = 128;        // rows of a tile (and the granularity of ldn / row shards)
constexpr int TILE_N = 64;       // columns of a tile
constexpr int NWARP = 8;         // 4 x 2 warps, each a 32 x 32 sub-tile = 4 x 4 DMMA blocks
constexpr int NTHREADS = NWARP * 32;
constexpr int CS_LD = TILE_N + 1;  // row stride of the staged output tile (odd: conflict-free column reads)
constexpr int CTAS_PER_SM = 2;   // co-resident CTAs in different phases: one's stores overlap another's MMAs

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// per-job tile bookkeeping for this shard: tile rows bi = rank, rank+count, ... ; row bi owns the column blocks
// (64 wide) 2*bi .. nc-1, i.e. everything on or right of the diagonal block
__device__ __host__ __forceinline__ long long shard_tiles_before(int q, int nc, int rank, int count) {
  // tiles in my first q rows: sum_{q'<q} (nc - 2*(rank + q'*count))
  return (long long)q * (nc - 2 * rank) - (long long)count * q * (q - 1);
}
__device__ __host__ __forceinline__ int shard_rows(int nb, int rank, int count) {
  return nb > rank ? (nb - rank + count - 1) / count : 0;
}

__global__ void k_tile_prefix(const int32_t* __restrict__ n_nodes, int B, int ldn, int rank, int count,
                              long long* __restrict__ prefix) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long acc = 0;
    for (int b = 0; b < B; ++b) {
      prefix[b] = acc;
      int N = min(n_nodes[b], ldn);
      int nb = (N + TILE - 1) / TILE;
      int nc = 2 * nb;                       // column blocks, padded to whole row blocks (z rows beyond N are zero)
      int rows = shard_rows(nb, rank, count);
      acc += shard_tiles_before(rows, nc, rank, count);
    }
    prefix[B] = acc;
  }
}

// item -> two int4: (job, 128-row block, 64-column block, N) and (k-steps, 0, int64 image of the threshold), written
// once so the tile loops decode with one 32-byte load and no dependent per-job loads
__global__ void k_tile_table(const int32_t* __restrict__ n_nodes, const int32_t* __restrict__ job_T,
                             const double* __restrict__ r_crit, int ldn, int rank, int count,
                             const long long* __restrict__ prefix, int4* __restrict__ table) {
  const int b = blockIdx.x, q = blockIdx.y;
  const int N = min(n_nodes[b], ldn);
  const int nb = (N + TILE - 1) / TILE, nc = 2 * nb;
  if (q >= shard_rows(nb, rank, count)) return;
  const int bi = rank + q * count;
  const long long base = prefix[b] + shard_tiles_before(q, nc, rank, count);
  const double rc = r_crit[b];
  const long long rcb = rc < 0.0 ? -1LL : __double_as_longlong(rc);
  const int4 meta = make_int4((job_T[b] + 3) >> 2, 0, (int)(unsigned)(rcb & 0xffffffffLL), (int)(rcb >> 32));
  for (int t = threadIdx.x; t < nc - 2 * bi; t += blockDim.x) {
    table[2 * (base + t)] = make_int4(b, bi, 2 * bi + t, N);
    table[2 * (base + t) + 1] = meta;
  }
}

__global__ void __launch_bounds__(NTHREADS, CTAS_PER_SM)
k_corr_tiles(const double* __restrict__ z, const int32_t* __restrict__ n_nodes,
             const int32_t* __restrict__ job_T, const double* __restrict__ r_crit,
             const long long* __restrict__ prefix, const int4* __restrict__ table, int B, int ldn, int Tp,
             double* __restrict__ R, double* __restrict__ tile_part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);          // [128][Tp]
  double* sB = sA + (size_t)TILE * Tp;                       // [64][Tp]
  __shared__ uint64_t full_bar, empty_bar;
  __shared__ double red_sum[NWARP];
  __shared__ double red_cnt[NWARP];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp >> 1, wc = warp & 1;           // warp position in the 4x2 grid
  if (tid == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&empty_bar, NWARP);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const long long total = prefix[B];
  const uint32_t a_bytes = (uint32_t)((size_t)TILE * Tp * sizeof(double));
  const uint32_t b_bytes = (uint32_t)((size_t)TILE_N * Tp * sizeof(double));

  auto decode = [&](long long item, int& b, int& bi, int& bj) {
    const int4 e = table[2 * item];
    b = e.x; bi = e.y; bj = e.z;
  };

  auto issue = [&](long long item) {
    int b, bi, bj;
    decode(item, b, bi, bj);
    const double* gA = z + ((size_t)b * ldn + (size_t)bi * TILE) * Tp;
    const double* gB = z + ((size_t)b * ldn + (size_t)bj * TILE_N) * Tp;
    mbar_expect_tx(&full_bar, a_bytes + b_bytes);
    bulk_g2s(sA, gA, a_bytes, &full_bar);
    bulk_g2s(sB, gB, b_bytes, &full_bar);
  };

  long long item = blockIdx.x;
  if (item >= total) return;
  if (tid == 0) issue(item);

  int eph = 0;                                       // uses of empty_bar so far (its phase)
  for (int it = 0; item < total; ++it, item += gridDim.x) {
    int b, bi, bj;
    decode(item, b, bi, bj);
    const int N = min(n_nodes[b], ldn);
    const int ksteps = (job_T[b] + 3) >> 2;
    const double rc_raw = r_crit[b];
    // `R >= 0 && R > rc` (ComplexNetworks.py:44-45) as ONE compare: rc >= 0 -> R > rc; rc < 0 -> R > -denorm_min (R >= 0)
    const double rc = (rc_raw < 0.0) ? -4.9406564584124654e-324 : rc_raw;

    mbar_wait(&full_bar, it & 1);

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const double* pa = sA + (size_t)(wr * 32 + (lane >> 2)) * Tp + (lane & 3);
    const double* pb = sB + (size_t)(wc * 32 + (lane >> 2)) * Tp + (lane & 3);
    for (int ks = 0; ks < ksteps; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        af[i] = pa[(size_t)i * 8 * Tp + ks * 4];
        bf[i] = pb[(size_t)i * 8 * Tp + ks * 4];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    const int row0 = bi * TILE, col0 = bj * TILE_N;
    const bool diag_tile = (col0 < row0 + TILE);     // the tile touches the diagonal band of its row block
    const bool staged = (R != nullptr) && !diag_tile;   // off-diagonal tiles are written through shared memory
    __syncwarp();
    if (!staged) {
      if (lane == 0) mbar_arrive(&empty_bar);        // this warp no longer reads the panels
      if (tid == 0 && item + gridDim.x < total) {    // refill them as soon as every warp is done: the load of the
        mbar_wait(&empty_bar, eph & 1);              // next tile overlaps this tile's epilogue
        issue(item + gridDim.x);
      }
      ++eph;
    }
    // ---- epilogue: clip, NaN diagonal, mirrored store, tau partials
    double lsum = 0.0, lcnt = 0.0;
    double* Rb = R ? R + (size_t)b * ldn * ldn : nullptr;
    if (staged) {
      // Every element of an off-diagonal tile is above the diagonal.  Only the upper triangle of R is stored (readers
      // address R[min(i,j)][max(i,j)]): stage the clipped tile in shared memory (over the operand panels, once every
      // warp has finished its MMAs) and write whole rows, 512-byte runs, instead of 64-byte pieces per fragment.
      __syncthreads();                               // all warps done with the panels
      double* Cs = reinterpret_cast<double*>(smem_raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int li = wr * 32 + i * 8 + (lane >> 2);
        const int gi = row0 + li;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int lj = wc * 32 + j * 8 + 2 * (lane & 3);
          const int gj = col0 + lj;
          double v0 = acc[i][j][0], v1 = acc[i][j][1];
          if (!(fabs(v0) <= 1.0)) v0 = v0 > 1.0 ? 1.0 : (v0 < -1.0 ? -1.0 : v0);   // np.clip keeps NaN
          if (!(fabs(v1) <= 1.0)) v1 = v1 > 1.0 ? 1.0 : (v1 < -1.0 ? -1.0 : v1);
          Cs[li * CS_LD + lj] = v0;
          Cs[li * CS_LD + lj + 1] = v1;
          if (gi < N) {
            if (gj < N && v0 > rc) { lsum += v0; lcnt += 1.0; }
            if (gj + 1 < N && v1 > rc) { lsum += v1; lcnt += 1.0; }
          }
        }
      }
      __syncthreads();
      // the tile: row r -> 64 doubles, lane l writes columns 2l, 2l+1 (4 rows per trip: the shared-memory reads of
      // all four are in flight before the first store)
      {
        const int gj = col0 + 2 * lane;
        for (int r0 = warp; r0 < TILE; r0 += 4 * NWARP) {
          double v[4][2];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * NWARP;
            v[u][0] = Cs[r * CS_LD + 2 * lane];
            v[u][1] = Cs[r * CS_LD + 2 * lane + 1];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int gi = row0 + r0 + u * NWARP;
            if (gi < N) {
              double* dst = Rb + (size_t)gi * ldn + gj;
              if (gj + 1 < N) *reinterpret_cast<double2*>(dst) = make_double2(v[u][0], v[u][1]);
              else if (gj < N) dst[0] = v[u][0];
            }
          }
        }
      }
      __syncthreads();                               // staging buffer free: the next panels may land
      if (tid == 0 && item + gridDim.x < total) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of Cs before the bulk copy lands
        issue(item + gridDim.x);
      }
    } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = row0 + wr * 32 + i * 8 + (lane >> 2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gj = col0 + wc * 32 + j * 8 + 2 * (lane & 3);
        double v0 = acc[i][j][0], v1 = acc[i][j][1];
        v0 = v0 > 1.0 ? 1.0 : (v0 < -1.0 ? -1.0 : v0);   // np.clip keeps NaN
        v1 = v1 > 1.0 ? 1.0 : (v1 < -1.0 ? -1.0 : v1);
        if (gi >= N) continue;
        const bool in0 = gj < N, in1 = gj + 1 < N;
        // an element counts when it is strictly above the diagonal; it stands for (i,j) and (j,i)
        const bool up0 = in0 && (!diag_tile || gj > gi), up1 = in1 && (!diag_tile || gj + 1 > gi);
        if (up0 && v0 > rc) { lsum += v0; lcnt += 1.0; }
        if (up1 && v1 > rc) { lsum += v1; lcnt += 1.0; }
        if (Rb) {
          if (diag_tile) {
            if (in0 && gj == gi) Rb[(size_t)gi * ldn + gj] = sie_nan();
            if (in1 && gj + 1 == gi) Rb[(size_t)gi * ldn + gj + 1] = sie_nan();
            if (up0) Rb[(size_t)gi * ldn + gj] = v0;
            if (up1) Rb[(size_t)gi * ldn + gj + 1] = v1;
          } else {
            if (in1) *reinterpret_cast<double2*>(Rb + (size_t)gi * ldn + gj) = make_double2(v0, v1);
            else if (in0) Rb[(size_t)gi * ldn + gj] = v0;
          }
        }
      }
    }
    }
    // deterministic CTA reduction of the tile partial (fixed shuffle tree, fixed warp order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
      lcnt += __shfl_xor_sync(0xffffffffu, lcnt, o);
    }
    if (lane == 0) { red_sum[warp] = lsum; red_cnt[warp] = lcnt; }
    __syncthreads();
    if (tid == 0) {
      double s = 0.0, c = 0.0;
      for (int w = 0; w < NWARP; ++w) { s += red_sum[w]; c += red_cnt[w]; }
      tile_part[2 * item] = 2.0 * s;       // both triangles
      tile_part[2 * item + 1] = 2.0 * c;
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// k_corr_rows: the row-resident, warp-specialised form of the same contraction (used whenever its shared
// memory fits; k_corr_tiles above remains the fallback for very long windows).
//   * one persistent CTA per SM walks a CONTIGUOUS run of the tile table, so the 128-row A panel of a tile row
//     stays in shared memory for the whole row and only the 64-row B panels stream (22.5 KB per tile instead of
//     67.6 KB out of L2);
//   * warp 16 is the producer: lane 0 keeps an S-stage ring of B panels full with bulk async copies (full/empty
//     mbarriers); warps 0-15 (4 x 4, a 32 x 16 sub-tile = 4 x 2 DMMA blocks each) compute.  Everything a tile needs
//     (job, block indices, N, k-steps, threshold image) sits in one 32-byte table entry that is prefetched a tile
//     ahead, so no dependent global load precedes the MMAs;
//   * STORE: the finished tile is staged in shared memory twice - as is (rows padded to 576 B) and transposed
//     (rows padded to 1040 B; both paddings make the fragment-layout writes bank-conflict free) - and written
//     with whole-row 16-byte stores: 512-byte runs for the tile, 1-KB runs for its mirror (tools/micro/wpat.cu:
//     plain stores of such runs reach > 6 TB/s, one bulk async store per run only 2.3-4.6 TB/s).  The stores are
//     fire-and-forget and drain while the CTA is already in the MMAs of the next tile - the panels are NOT
//     overlaid by the staging buffers.  Whole tiles are written, so the zero padding of R (rows/columns N..ldn of
//     the last blocks) is written too; the diagonal blocks are symmetrised in the staging buffer (only the strict
//     upper triangle is taken from the accumulators) so R stays bitwise symmetric;
//   * !STORE (25 km row shards): no staging, deeper ring, no CTA-wide barrier at all;
//   * thresholds are integer compares on the bit patterns, which leaves only the predicated DADD on the FP64 pipe.
// Per-tile tau partials are written per warp ([tile][16][2]), folded per tile by k_tau_tiles and reduced per
// job in a fixed order by k_tau_finalize.
// -------------------------------------------------------------------------------------------------
constexpr int RW_CWARPS = 16;                    // consumer warps (4 x 4)
constexpr int RW_SWARPS = 15;                    // store warps (STORE only): one warp sustains only ~3.6 GB/s of stores
                                                 // (tools/micro/wpat.cu), HBM's share per SM needs >= 12 of them
constexpr int RW_MAXS = 6;                       // deepest B ring
constexpr int RW_D_LD = TILE_N + 8;              // 72 doubles = 576 B per staged row
constexpr int RW_T_LD = TILE + 2;                // 130 doubles = 1040 B per staged transposed row
template <bool STORE> constexpr int rw_threads() { return (RW_CWARPS + 1 + (STORE ? RW_SWARPS : 0)) * 32; }

constexpr long long kOneBits = 0x3ff0000000000000LL;   // int64 image of 1.0: after clip_unit only NaNs lie above it
// clip to [-1, 1] (np.clip keeps NaN): one integer compare on the high word in the common |v| < 1 case
__device__ __forceinline__ double clip_unit(double v) {
  const unsigned hi = (unsigned)__double2hiint(v) & 0x7fffffffu;
  if (hi >= 0x3ff00000u) {
    if (v > 1.0) v = 1.0;
    else if (v < -1.0) v = -1.0;
  }
  return v;
}

template <bool STORE, bool MIRROR>
__global__ void __launch_bounds__(rw_threads<STORE>(), 1)
k_corr_rows(const double* __restrict__ z, const long long* __restrict__ prefix, const int4* __restrict__ table,
            int B, int ldn, int Tp, int S, double* __restrict__ R, double* __restrict__ parts) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);               // [128][Tp]
  double* sB = sA + (size_t)TILE * Tp;                            // [S][64][Tp]
  double* sD = sB + (size_t)S * TILE_N * Tp;                      // [128][RW_D_LD]   (STORE)
  double* sT = sD + (size_t)TILE * RW_D_LD;                       // [64][RW_T_LD]    (STORE && MIRROR)
  __shared__ uint64_t full_bar[RW_MAXS], empty_bar[RW_MAXS];
  __shared__ uint64_t d_full, d_free, t_full, t_free;             // staging handshakes consumers <-> store warps

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long total = prefix[B];
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long lo = (long long)blockIdx.x * per;
  const long long hi = lo + per < total ? lo + per : total;
  const int n = (int)(hi - lo);
  if (n <= 0) return;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], RW_CWARPS); }
    mbar_init(&d_full, RW_CWARPS); mbar_init(&t_full, RW_CWARPS);
    mbar_init(&d_free, RW_SWARPS); mbar_init(&t_free, RW_SWARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t a_bytes = (uint32_t)((size_t)TILE * Tp * sizeof(double));
  const uint32_t b_bytes = (uint32_t)((size_t)TILE_N * Tp * sizeof(double));

  if (warp >= RW_CWARPS) {
  // 1024 threads start with 64 registers each: the 4 consumer warpgroups take 80, the other 4 give back down to 48
  if (STORE) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
  if (warp == RW_CWARPS) {
    // ---------------- producer warp
    if (lane != 0) return;
    int cur_b = -1, cur_bi = -1;
    int4 e = table[2 * lo];
    for (int nl = 0; nl < n; ++nl) {
      const int4 en = table[2 * (lo + (nl + 1 < n ? nl + 1 : nl))];   // next entry in flight during the waits
      const int s = nl % S, f = nl / S;
      if (f > 0) mbar_wait(&empty_bar[s], (f - 1) & 1);               // the ring slot is free
      const bool new_row = (e.x != cur_b || e.y != cur_bi);
      if (new_row) {
        // every consumer must be done with the old A panel: wait for the tiles still in the ring
        for (int j = (nl - S + 1 > 0 ? nl - S + 1 : 0); j < nl; ++j) mbar_wait(&empty_bar[j % S], (j / S) & 1);
        cur_b = e.x; cur_bi = e.y;
      }
      mbar_expect_tx(&full_bar[s], b_bytes + (new_row ? a_bytes : 0u));
      if (new_row) bulk_g2s(sA, z + ((size_t)e.x * ldn + (size_t)e.y * TILE) * Tp, a_bytes, &full_bar[s]);
      bulk_g2s(sB + (size_t)s * TILE_N * Tp, z + ((size_t)e.x * ldn + (size_t)e.z * TILE_N) * Tp, b_bytes, &full_bar[s]);
      e = en;
    }
    return;
  }

  if (STORE && warp > RW_CWARPS) {
    // ---------------- store warps: staged tile -> R with whole-row 16-byte stores, while the consumers compute
    const int sw = warp - RW_CWARPS - 1;
    int4 e = table[2 * lo];
    for (int k = 0; k < n; ++k) {
      const int4 en = table[2 * (lo + (k + 1 < n ? k + 1 : k))];
      const int row0 = e.y * TILE, col0 = e.z * TILE_N, dk = e.z - 2 * e.y;
      double* Rb = R + (size_t)e.x * ldn * ldn;
      // the tile: staged row r -> 512 B of R row row0+r (rows 64.. of the first diagonal tile are mirrors: skipped)
      const int nrow = dk == 0 ? 64 : TILE;
      mbar_wait(&d_full, k & 1);
      for (int r0 = sw; r0 < nrow; r0 += 4 * RW_SWARPS) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * RW_SWARPS;
          if (r < nrow) v[u] = *reinterpret_cast<const double2*>(sD + (size_t)r * RW_D_LD + 2 * lane);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * RW_SWARPS;
          if (r < nrow) *reinterpret_cast<double2*>(Rb + (size_t)(row0 + r) * ldn + col0 + 2 * lane) = v[u];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_free);
      // its mirror: staged transposed row c -> 1 KB (512 B for the second diagonal tile) of R row col0+c
      if (!MIRROR) { e = en; continue; }              // upper triangle only: readers address R[min][max]
      mbar_wait(&t_full, k & 1);
      if (dk != 0) {
        const int nh = dk == 1 ? 1 : 2;
        for (int c0 = sw; c0 < TILE_N; c0 += 2 * RW_SWARPS) {
          double2 v[2][2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = c0 + u * RW_SWARPS;
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (c < TILE_N) v[u][h] = *reinterpret_cast<const double2*>(sT + (size_t)c * RW_T_LD + 64 * h + 2 * lane);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = c0 + u * RW_SWARPS;
            double* dst = Rb + (size_t)(col0 + c) * ldn + row0 + 2 * lane;
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (c < TILE_N && h < nh) *reinterpret_cast<double2*>(dst + 64 * h) = v[u][h];
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_free);
      e = en;
    }
    return;
  }
  return;
  }

  // ---------------- consumers
  if (STORE) asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
  const int wr = warp >> 2, wc = warp & 3;
  const double* pa = sA + (size_t)(wr * 32 + (lane >> 2)) * Tp + (lane & 3);
  int4 e0 = table[2 * lo], e1 = table[2 * lo + 1];
  for (int k = 0; k < n; ++k) {
    const long long item = lo + k;
    const int bi = e0.y, bj = e0.z, N = e0.w, ksteps = e1.x;
    // `R >= 0 && R > rc` (ComplexNetworks.py:44-45) on the bit patterns: for rc >= 0 a double is > rc exactly when
    // its int64 image is (negative doubles have negative images); rc < 0 leaves `R >= 0` = non-negative image
    const long long rcb = (long long)(((unsigned long long)(unsigned)e1.w << 32) | (unsigned)e1.z);
    {
      const long long nx = item + 1 < hi ? item + 1 : item;       // next tile's entry: in flight during the MMAs
      e0 = table[2 * nx];
      e1 = table[2 * nx + 1];
    }
    const int s = k % S;
    const double* pb = sB + (size_t)s * TILE_N * Tp + (size_t)(wc * 16 + (lane >> 2)) * Tp + (lane & 3);

    mbar_wait(&full_bar[s], (k / S) & 1);
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ++ks) {
      double af[4], bf[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = pa[(size_t)i * 8 * Tp + ks * 4];
#pragma unroll
      for (int j = 0; j < 2; ++j) bf[j] = pb[(size_t)j * 8 * Tp + ks * 4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);                     // this warp no longer reads the panels

    const int row0 = bi * TILE, col0 = bj * TILE_N;
    const int dk = bj - 2 * bi;                                    // 0 / 1: the two tiles of the diagonal block
    const bool interior = dk >= 2 && row0 + TILE <= N && col0 + TILE_N <= N;
    // the 64 x 64 square on the diagonal sits in rows dsq.. of the tile (dk = 0: rows 0-63, dk = 1: rows 64-127)
    const int dsq = dk == 0 ? 0 : (dk == 1 ? 64 : -1);
    double lsum = 0.0;
    int lcnt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { acc[i][j][0] = clip_unit(acc[i][j][0]); acc[i][j][1] = clip_unit(acc[i][j][1]); }
    // ---- pass 1: tau partial + the tile as is
    if (STORE && k > 0) mbar_wait(&d_free, (k - 1) & 1);           // tile k-1 has left sD
    if (interior) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int li = wr * 32 + i * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int lj = wc * 16 + j * 8 + 2 * (lane & 3);
          const double v0 = acc[i][j][0], v1 = acc[i][j][1];
          const long long i0 = __double_as_longlong(v0), i1 = __double_as_longlong(v1);
          const bool t0 = i0 > rcb && i0 <= kOneBits, t1 = i1 > rcb && i1 <= kOneBits;   // NaN images lie above 1.0's
          lsum += t0 ? v0 : 0.0;                                   // adding +0.0 leaves a non-negative sum unchanged
          lsum += t1 ? v1 : 0.0;
          lcnt += (int)t0 + (int)t1;
          if (STORE) *reinterpret_cast<double2*>(sD + li * RW_D_LD + lj) = make_double2(v0, v1);
        }
      }
    } else {
      // tiles of the diagonal block and of the ragged edge
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int li = wr * 32 + i * 8 + (lane >> 2);
        const int gi = row0 + li;
        const bool in_sq = dsq >= 0 && li >= dsq && li < dsq + 64;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int lj = wc * 16 + j * 8 + 2 * (lane & 3);
          const int gj = col0 + lj;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double v = acc[i][j][h];
            const bool up = gi < N && gj + h < N && (dk >= 2 || gj + h > gi);   // strictly above the diagonal
            const long long iv = __double_as_longlong(v);
            const bool t = up && iv > rcb && iv <= kOneBits;
            lsum += t ? v : 0.0;
            lcnt += (int)t;
            if (STORE) {
              if (in_sq) {
                const int a = li - dsq, c = lj + h;               // position inside the diagonal square
                if (c > a) { sD[li * RW_D_LD + c] = v; sD[(dsq + c) * RW_D_LD + a] = v; }
                else if (c == a) sD[li * RW_D_LD + c] = sie_nan();
              } else if (dk != 0) {
                sD[li * RW_D_LD + lj + h] = v;
              }
            }
          }
        }
      }
    }
    if (STORE) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_full);
    }
    if (STORE && MIRROR) {
      // ---- pass 2: the transposed tile (nothing for the first diagonal tile and for the diagonal square)
      if (k > 0) mbar_wait(&t_free, (k - 1) & 1);                  // tile k-1 has left sT
      if (dk != 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int li = wr * 32 + i * 8 + (lane >> 2);
          if (dk == 1 && li >= 64) continue;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int lj = wc * 16 + j * 8 + 2 * (lane & 3);
            sT[lj * RW_T_LD + li] = acc[i][j][0];
            sT[(lj + 1) * RW_T_LD + li] = acc[i][j][1];
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full);
    }
    double lc = (double)lcnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
      lc += __shfl_xor_sync(0xffffffffu, lc, o);
    }
    if (lane == 0) *reinterpret_cast<double2*>(parts + (item * RW_CWARPS + warp) * 2) = make_double2(lsum, lc);
  }
}

// -------------------------------------------------------------------------------------------------
// k_corr_tma: the stored-R kernel.  Same contraction and warp geometry as k_corr_rows (16 consumer warps, 4 x 4, a
// 32 x 16 sub-tile each; one producer warp; 128 x 64 tiles), but the finished tile leaves through the TMA engine:
//   * every consumer warp stages its 32 x 16 sub-tile (32 rows of 128 B, 4 KB) in its own shared-memory box and lane 0
//     issues ONE tensor-map bulk store (cp.async.bulk.tensor.3d, SASS UTMASTG; tensor map of R [B][ldn][ldn], box
//     16 x 32 x 1).  The store is asynchronous: the warp goes straight to the MMAs of its next tile and only waits
//     (cp.async.bulk.wait_group.read) before it overwrites the box again, a whole tile later.  No store warps, no CTA
//     barrier, no second trip of the tile through the LSU pipe (k_corr_tiles stages, reads back and stores with plain
//     st.global: store-issue bound);
//   * the fragment -> box writes are bank-conflict free without swizzling: a quarter warp holds two rows x 64 B of each
//     of the two 8-column blocks; even rows write their left block first, odd rows their right block, so the eight
//     16-byte pieces of one st.shared.v2.f64 wavefront cover all 32 banks;
//   * sub-tiles strictly below the diagonal or outside the N x N matrix are not stored; the diagonal gets NaN;
//   * work split: chunks of TM_CHUNK consecutive tiles dealt round-robin to the CTAs.  The item table is ordered by
//     job, and a sweep's jobs by window length, so a contiguous split (k_corr_rows) would give one CTA only long-window
//     tiles and another only short ones; round-robin chunks give every CTA the same mix, and a chunk mostly stays
//     inside one tile row, so the resident A panel is still reused;
//   * the A panel is double-buffered (a new tile row's panel loads while the old one is still being consumed); the
//     producer hands every ring slot its table entry + A buffer index through shared memory, a negative job ends the CTA.
// tau partials: per warp and tile, same layout and summation order as k_corr_rows (k_tau_tiles folds them).
// -------------------------------------------------------------------------------------------------
constexpr int TM_CWARPS = 16;
constexpr int TM_THREADS = (TM_CWARPS + 1) * 32;
constexpr int TM_MAXS = 6;
constexpr int TM_CHUNK = 8;
constexpr int TM_BOX_R = 32, TM_BOX_C = 16;       // one consumer warp's sub-tile = one TMA box

__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(TM_THREADS, 1)
k_corr_tma(const double* __restrict__ z, const long long* __restrict__ prefix, const int4* __restrict__ table, int B,
           int ldn, int Tp, int S, const __grid_constant__ CUtensorMap tmR, double* __restrict__ parts) {
  constexpr int chunk = TM_CHUNK;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t a_elems = (size_t)TILE * Tp, b_elems = (size_t)TILE_N * Tp;
  double* sA = reinterpret_cast<double*>(smem_raw);               // [2][128][Tp]
  double* sB = sA + 2 * a_elems;                                  // [S][64][Tp]
  double* sC = sB + (size_t)S * b_elems;                          // [16 warps][32][16]: the TMA store boxes
  __shared__ uint64_t full_bar[TM_MAXS], empty_bar[TM_MAXS];
  __shared__ int4 meta0[TM_MAXS], meta1[TM_MAXS];                 // ring slot -> table entry (+ A buffer in meta1.y)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long total = prefix[B];
  const long long G = gridDim.x, c = blockIdx.x;
  auto item_of = [&](int k) -> long long { const int q = k / chunk; return ((long long)q * G + c) * chunk + (k - q * chunk); };
  if (item_of(0) >= total) return;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], TM_CWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t a_bytes = (uint32_t)(a_elems * sizeof(double));
  const uint32_t b_bytes = (uint32_t)(b_elems * sizeof(double));

  if (warp == TM_CWARPS) {
    // ---------------- producer warp (one thread)
    if (lane != 0) return;
    int consumed = -1;                                            // tiles 0..consumed have left the ring and their A panel
    auto ensure = [&](int t) {
      while (consumed < t) { ++consumed; mbar_wait(&empty_bar[consumed % S], (consumed / S) & 1); }
    };
    int cur_b = -1, cur_bi = -1, abuf = 1;
    int a_last[2] = {-1, -1};                                     // last tile that reads each A buffer
    long long item = item_of(0);
    int4 e = table[2 * item], m = table[2 * item + 1];
    int nl = 0;
    while (true) {
      const long long nitem = item_of(nl + 1);
      const bool more = nitem < total;
      int4 en = e, mn = m;
      if (more) { en = table[2 * nitem]; mn = table[2 * nitem + 1]; }   // next entry in flight during the waits
      const int s = nl % S;
      ensure(nl - S);                                             // the ring slot is free
      const bool new_row = (e.x != cur_b || e.y != cur_bi);
      if (new_row) {
        abuf ^= 1;
        ensure(a_last[abuf]);                                     // every tile that read this A buffer is done
        cur_b = e.x; cur_bi = e.y;
      }
      a_last[abuf] = nl;
      meta0[s] = e;
      meta1[s] = make_int4(m.x, abuf, m.z, m.w);
      mbar_expect_tx(&full_bar[s], b_bytes + (new_row ? a_bytes : 0u));
      if (new_row) bulk_g2s(sA + (size_t)abuf * a_elems, z + ((size_t)e.x * ldn + (size_t)e.y * TILE) * Tp, a_bytes, &full_bar[s]);
      bulk_g2s(sB + (size_t)s * b_elems, z + ((size_t)e.x * ldn + (size_t)e.z * TILE_N) * Tp, b_bytes, &full_bar[s]);
      ++nl;
      if (!more) break;
      e = en; m = mn; item = nitem;
    }
    const int s = nl % S;                                         // end marker
    ensure(nl - S);
    meta0[s] = make_int4(-1, 0, 0, 0);
    mbar_arrive(&full_bar[s]);
    return;
  }

  // ---------------- consumers
  const int wr = warp >> 2, wc = warp & 3;
  const int r8 = lane >> 2, c4 = lane & 3;
  double* box = sC + (size_t)warp * (TM_BOX_R * TM_BOX_C);
  bool pending = false;                                           // lane 0: a bulk store of `box` may still be reading it
  int s = -1;
  uint32_t ring_phase = 1u;                                       // parity of the current pass over the ring
  for (int k = 0;; ++k) {
    if (++s == S) s = 0;
    if (s == 0) ring_phase ^= 1u;
    mbar_wait(&full_bar[s], ring_phase);
    const int4 e0 = meta0[s], e1 = meta1[s];
    if (e0.x < 0) break;
    const long long item = item_of(k);
    const int b = e0.x, bi = e0.y, bj = e0.z, N = e0.w, ksteps = e1.x;
    const long long rcb = (long long)(((unsigned long long)(unsigned)e1.w << 32) | (unsigned)e1.z);
    // `R >= 0 && R > rc` as one compare: rc >= 0 -> R > rc; rc < 0 (image -1) -> R > -denorm_min (R >= 0)
    const double rcd = rcb < 0 ? -4.9406564584124654e-324 : __longlong_as_double(rcb);
    const double* pa = sA + (size_t)e1.y * a_elems + (size_t)(wr * 32 + r8) * Tp + c4;
    const double* pb = sB + (size_t)s * b_elems + (size_t)(wc * 16 + r8) * Tp + c4;

    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ++ks) {
      double af[4], bf[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = pa[(size_t)i * 8 * Tp + ks * 4];
#pragma unroll
      for (int j = 0; j < 2; ++j) bf[j] = pb[(size_t)j * 8 * Tp + ks * 4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);                     // this warp no longer reads the panels

    const int row0 = bi * TILE, col0 = bj * TILE_N;
    const int dk = bj - 2 * bi;                                    // 0 / 1: the two tiles of the diagonal block
    const bool interior = dk >= 2 && row0 + TILE <= N && col0 + TILE_N <= N;
    double lsum = 0.0;
    int lcnt = 0;
    // ---- the sub-tile leaves through the TMA engine (skipped when it lies strictly below the diagonal or outside N)
    const int sr0 = row0 + wr * 32, sc0 = col0 + wc * 16;
    const bool stored = sc0 + TM_BOX_C - 1 >= sr0 && sr0 < N && sc0 < N;
    auto stage_store = [&](const double (&v)[4][2][2]) {
      if (!stored) return;
      if (lane == 0 && pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      const bool odd = r8 & 1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double* rowp = box + (size_t)(i * 8 + r8) * TM_BOX_C + 2 * c4;
        const double2 left = make_double2(v[i][0][0], v[i][0][1]), right = make_double2(v[i][1][0], v[i][1][1]);
        // even rows: left block then right block; odd rows the other way round (bank-conflict free wavefronts)
        *reinterpret_cast<double2*>(rowp + (odd ? 8 : 0)) = odd ? right : left;
        *reinterpret_cast<double2*>(rowp + (odd ? 0 : 8)) = odd ? left : right;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmR, box, sc0, sr0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        pending = true;
      }
    };
    // clip to [-1, 1] (np.clip keeps NaN): ONE test per tile on the largest |high word| instead of one per value
    unsigned hmax = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) hmax = max(hmax, (unsigned)__double2hiint(acc[i][j][h]) & 0x7fffffffu);
    if (interior && hmax < 0x3ff00000u) {
      // the common case touches the accumulators read-only (a conditional in-place clip / NaN diagonal makes the compiler
      // copy all 32 accumulator registers on this path too; the epilogue is what bounds this kernel)
      double ps[4] = {0.0, 0.0, 0.0, 0.0};                          // four independent partial sums: the 16 adds are a
#pragma unroll                                                     // chain of 4 + 2 dependent FP64 adds instead of 16
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double v0 = acc[i][j][0], v1 = acc[i][j][1];
          // one FP64 compare per value (NaN compares false): a third of the instructions of the integer-image test of
          // k_corr_rows; this kernel's epilogue is latency / issue-bound, not FP64-pipe-bound
          const bool t0 = v0 > rcd, t1 = v1 > rcd;
          ps[i] += t0 ? v0 : 0.0;                                  // adding +0.0 leaves a non-negative sum unchanged
          ps[i] += t1 ? v1 : 0.0;
          lcnt += (int)t0 + (int)t1;
        }
      lsum = (ps[0] + ps[1]) + (ps[2] + ps[3]);
      stage_store(acc);
    } else {
      // tiles of the diagonal block and of the ragged edge, or a value that needs clipping: count strictly above the
      // diagonal, NaN on it
      double cv[4][2][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gi = row0 + wr * 32 + i * 8 + r8;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int gj = col0 + wc * 16 + j * 8 + 2 * c4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double v = clip_unit(acc[i][j][h]);
            const bool up = gi < N && gj + h < N && (dk >= 2 || gj + h > gi);
            const bool t = up && v > rcd;
            lsum += t ? v : 0.0;
            lcnt += (int)t;
            cv[i][j][h] = (dk < 2 && gj + h == gi) ? sie_nan() : v;
          }
        }
      }
      stage_store(cv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
      lcnt += __shfl_xor_sync(0xffffffffu, lcnt, o);                 // the count stays an integer until it is stored
    }
    if (lane == 0) *reinterpret_cast<double2*>(parts + (item * TM_CWARPS + warp) * 2) = make_double2(lsum, (double)lcnt);
  }
  if (lane == 0 && pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores done before the CTA exits
}

// one thread per tile: the 16 warp partials of k_corr_rows in warp order, doubled (both triangles) -> one pair per tile
__global__ void k_tau_tiles(const double* __restrict__ parts, const long long* __restrict__ prefix, int B,
                            double* __restrict__ tile_pair) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= prefix[B]) return;
  const double2* p = reinterpret_cast<const double2*>(parts) + i * RW_CWARPS;
  double2 v[RW_CWARPS];
#pragma unroll
  for (int w = 0; w < RW_CWARPS; ++w) v[w] = p[w];
  double ts = 0.0, tc = 0.0;
#pragma unroll
  for (int w = 0; w < RW_CWARPS; ++w) { ts += v[w].x; tc += v[w].y; }
  reinterpret_cast<double2*>(tile_pair)[i] = make_double2(2.0 * ts, 2.0 * tc);
}

// one 256-thread block per job: fixed-order reduction of the per-tile pairs -> tau_sum, tau_cnt, tau
// (thread t adds tiles t, t+256, ... in order; then a fixed shuffle tree and a fixed warp order)
__global__ void __launch_bounds__(256)
k_tau_finalize(const double* __restrict__ tile_pair, const long long* __restrict__ prefix, double* __restrict__ tau_sum,
               int64_t* __restrict__ tau_cnt, double* __restrict__ tau) {
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ double rs[8], rc[8];
  const double2* tp = reinterpret_cast<const double2*>(tile_pair);
  double s = 0.0, c = 0.0;
  for (long long i = prefix[b] + threadIdx.x; i < prefix[b + 1]; i += 256) { const double2 v = tp[i]; s += v.x; c += v.y; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0) { rs[warp] = s; rc[warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0.0; c = 0.0;
    for (int w = 0; w < 8; ++w) { s += rs[w]; c += rc[w]; }
    tau_sum[b] = s;
    tau_cnt[b] = (int64_t)c;
    tau[b] = s / c;     // 0/0 -> NaN like np.mean([])
  }
}

// K3: one thread per (node, direction).  R == nullptr: the correlation is recomputed from the two unit-norm rows
// (sequential in k like the tensor-core accumulation, clipped), so the seed search works without a stored matrix.
__global__ void k_stencil(const double* __restrict__ R, const double* __restrict__ z,
                          const int32_t* __restrict__ job_T, int Tp, const int32_t* __restrict__ node_cell,
                          const int32_t* __restrict__ cell_node, const int32_t* __restrict__ n_nodes, int B,
                          int X, int Y, int ldn, int latlon, double* __restrict__ stencil) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * ldn * 4) return;
  const int d = (int)(idx & 3);
  const long long bn = idx >> 2;
  const int b = (int)(bn / ldn), n = (int)(bn - (long long)b * ldn);
  double out = sie_nan();
  const int N = min(n_nodes[b], ldn);
  if (n < N) {
    const int c = node_cell[(size_t)b * ldn + n];
    int i = c / Y, j = c - i * Y;
    int a = i + (d == 0 ? -1 : (d == 1 ? 1 : 0));
    int q = j + (d == 2 ? -1 : (d == 3 ? 1 : 0));
    bool ok = (a >= 0 && a < X);
    if (q < 0) { if (latlon) q = Y - 1; else ok = false; }
    if (q >= Y) { if (latlon) q = 0; else ok = false; }
    if (ok) {
      const int m = cell_node[(size_t)b * X * Y + a * Y + q];
      if (m >= 0 && m < N) {
        if (R) {
          out = (m == n) ? sie_nan()
                         : R[(size_t)b * ldn * ldn + (size_t)min(n, m) * ldn + max(n, m)];   // upper triangle stored
        } else {
          const double* za = z + ((size_t)b * ldn + n) * Tp;
          const double* zc = z + ((size_t)b * ldn + m) * Tp;
          const int kT = (job_T[b] + 3) & ~3;
          double acc = 0.0;
          for (int k = 0; k < kT; ++k) acc = fma(za[k], zc[k], acc);
          out = clip_unit(acc);
          if (m == n) out = sie_nan();
        }
      }
    }
  }
  stencil[idx] = out;
}

// selected rows of R recomputed from z: one thread per (row, column); the same sequential dot product as K3 / K4
__global__ void k_corr_rows_from_z(const double* __restrict__ z, const int32_t* __restrict__ rows, int n_rows, int N,
                                   int T, int Tp, double* __restrict__ out, int ld_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= N || r >= n_rows) return;
  const int a = rows[r];
  double v = sie_nan();
  if (a >= 0 && a < N && a != c) {
    const double* za = z + (size_t)a * Tp;
    const double* zc = z + (size_t)c * Tp;
    const int kT = (T + 3) & ~3;
    double acc = 0.0;
    for (int k = 0; k < kT; ++k) acc = fma(za[k], zc[k], acc);
    v = clip_unit(acc);
  }
  out[(size_t)r * ld_out + c] = v;
}

}  // namespace

extern "C" int sie_corr_rows(const double* z, const int32_t* rows, int n_rows, int N, int T, int Tp, double* out,
                             int ld_out, void* stream) {
  SIE_CHECK_ARG(z && rows && out, "null pointer");
  SIE_CHECK_ARG(n_rows > 0 && n_rows <= 65535 && N > 0 && T > 0 && Tp >= T && ld_out >= N, "bad size");
  k_corr_rows_from_z<<<dim3((unsigned)((N + 255) / 256), (unsigned)n_rows), 256, 0, (cudaStream_t)stream>>>(
      z, rows, n_rows, N, T, Tp, out, ld_out);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}

extern "C" size_t sie_corr_tau_scratch_bytes(int B, int ldn) {
  long long nb = (ldn + TILE - 1) / TILE;
  long long tiles = (long long)B * nb * (nb + 1);      // 128 x 64 tiles on or right of the diagonal blocks
  return (size_t)(tiles * (2 * RW_CWARPS * sizeof(double) + 2 * sizeof(int4) + 2 * sizeof(double)) +
                  (size_t)(B + 1) * sizeof(long long) + 1024);
}

extern "C" int sie_corr_tau(const double* z, const int32_t* n_nodes, const int32_t* job_T,
                            const double* r_crit, int B, int ldn, int Tp, double* R, double* tile_part,
                            size_t tile_part_bytes, double* tau_sum, int64_t* tau_cnt, double* tau,
                            int shard_rank, int shard_count, int kernel, void* stream) {
  SIE_CHECK_ARG(kernel == SIE_CORR_AUTO || kernel == SIE_CORR_TILES || kernel == SIE_CORR_ROWS ||
                    kernel == SIE_CORR_ROWS_MIRROR || kernel == SIE_CORR_TMA, "unknown kernel choice");
  SIE_CHECK_ARG(z && n_nodes && job_T && r_crit && tile_part && tau_sum && tau_cnt && tau, "null pointer");
  SIE_CHECK_ARG(B > 0 && ldn > 0 && (ldn % TILE) == 0, "ldn must be a positive multiple of 128");
  SIE_CHECK_ARG(Tp >= 4 && (Tp % 4) == 0, "Tp must be a multiple of 4");
  SIE_CHECK_ARG(shard_count >= 1 && shard_rank >= 0 && shard_rank < shard_count, "bad shard");
  SIE_CHECK_ARG(tile_part_bytes >= sie_corr_tau_scratch_bytes(B, ldn), "tile_part scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem_panels = (size_t)(TILE + TILE_N) * Tp * sizeof(double);
  const size_t smem_stage = R ? (size_t)TILE * CS_LD * sizeof(double) : 0;   // output tile staged over the panels
  const size_t smem = smem_panels > smem_stage ? smem_panels : smem_stage;
  const SieDevice* dev = sie_device();
  if (!dev) return SIE_ERR_LAUNCH;
  const int sms = dev->sm_count, max_optin = dev->max_smem_optin;
  // scratch layout: [prefix (B+1) int64][pad to 256][warp partials 16 pairs/tile][pad][item table 32 B/tile][pad][tile pairs]
  long long* prefix = reinterpret_cast<long long*>(tile_part);
  unsigned char* base = reinterpret_cast<unsigned char*>(tile_part);
  long long nb = ldn / TILE;
  long long max_items = (long long)B * nb * (nb + 1);
  size_t off = (((size_t)(B + 1) * sizeof(long long)) + 255) / 256 * 256;
  double* parts = reinterpret_cast<double*>(base + off);
  size_t off2 = (off + (size_t)max_items * 2 * RW_CWARPS * sizeof(double) + 255) / 256 * 256;
  int4* table = reinterpret_cast<int4*>(base + off2);
  size_t off3 = (off2 + (size_t)max_items * 2 * sizeof(int4) + 255) / 256 * 256;
  double* tile_pair = reinterpret_cast<double*>(base + off3);
  SIE_CHECK_ARG(off3 + (size_t)max_items * 2 * sizeof(double) <= tile_part_bytes, "tile_part scratch too small");
  k_tile_prefix<<<1, 32, 0, st>>>(n_nodes, B, ldn, shard_rank, shard_count, prefix);
  SIE_CHECK_LAUNCH();
  k_tile_table<<<dim3((unsigned)B, (unsigned)nb), 128, 0, st>>>(n_nodes, job_T, r_crit, ldn, shard_rank, shard_count, prefix,
                                                              table);
  SIE_CHECK_LAUNCH();
  // row-resident kernel when its shared memory fits (always for the 1979-2020 windows, Tp <= 44), else the tile kernel
  const bool mirror = (kernel == SIE_CORR_ROWS_MIRROR);           // also write R[j][i] (readers never need it)
  const size_t rw_fixed = (size_t)TILE * Tp * sizeof(double) +
                          (R ? ((size_t)TILE * RW_D_LD + (mirror ? (size_t)TILE_N * RW_T_LD : 0)) * sizeof(double) : 0);
  const size_t rw_stage = (size_t)TILE_N * Tp * sizeof(double);
  int S = (size_t)max_optin < rw_fixed + 256 ? 0 : (int)(((size_t)max_optin - 256 - rw_fixed) / rw_stage);
  if (S > RW_MAXS) S = RW_MAXS;
  // defaults (SIE_CORR_AUTO): rows kernel for the tau-only pass (25.2 vs 22.6 TFLOP/s at 25 km); TMA-store kernel when R is
  // stored (4.0 ms on 576 57x57 networks against 5.6 ms for the tile kernel and 6.4 ms for the rows kernel's store-warp
  // mode, tools/corr_ab.py), the tile kernel only when the window is too long for the TMA kernel's shared memory.
  // `kernel` = SIE_CORR_TILES / _ROWS / _ROWS_MIRROR / _TMA overrides (A/B timing, parity tests of every path).
  // stored R: the TMA-store kernel (tensor-map bulk stores straight from the consumer warps) whenever its shared memory
  // fits: two A panels + the 64-KB store boxes + >= 2 ring stages
  const size_t tm_fixed = 2 * (size_t)TILE * Tp * sizeof(double) + (size_t)TM_CWARPS * TM_BOX_R * TM_BOX_C * sizeof(double);
  int tmS = (size_t)max_optin < tm_fixed + 512 ? 0 : (int)(((size_t)max_optin - 512 - tm_fixed) / rw_stage);
  if (tmS > TM_MAXS) tmS = TM_MAXS;
  const bool use_tma = R != nullptr && tmS >= 2 && (kernel == SIE_CORR_AUTO || kernel == SIE_CORR_TMA);
  if (kernel == SIE_CORR_TMA && !use_tma) {
    sie_set_error("sie_corr_tau: SIE_CORR_TMA needs a stored R and Tp <= ~52 (Tp = %d)", Tp);
    return SIE_ERR_UNSUPPORTED;
  }
  const bool want_rows = kernel == SIE_CORR_AUTO ? (R == nullptr) : (kernel != SIE_CORR_TILES);
  const bool use_rows = S >= 2 && want_rows && !use_tma;
  if (use_tma) {
    alignas(64) CUtensorMap tm;
    if (int rc = sie_tensor_map_f64_3d(&tm, R, B, ldn, TM_BOX_C, TM_BOX_R)) return rc;
    const size_t tsmem = tm_fixed + (size_t)tmS * rw_stage;
    if (int rc = sie_ensure_smem(dev, SIE_K_CORR_TMA, (const void*)k_corr_tma, tsmem)) return rc;
    const long long chunks = (max_items + TM_CHUNK - 1) / TM_CHUNK;     // chunk sizes 2..64 and ring depths 2..6 time the same
    const int grid = (int)(chunks < sms ? chunks : sms);
    k_corr_tma<<<grid, TM_THREADS, tsmem, st>>>(z, prefix, table, B, ldn, Tp, tmS, tm, parts);
    SIE_CHECK_LAUNCH();
    k_tau_tiles<<<(unsigned)((max_items + 255) / 256), 256, 0, st>>>(parts, prefix, B, tile_pair);
    SIE_CHECK_LAUNCH();
  } else
  if (use_rows) {
    const size_t rsmem = rw_fixed + (size_t)S * rw_stage;
    const int grid = (int)(max_items < sms ? max_items : sms);
    if (R && mirror) {
      if (int rc = sie_ensure_smem(dev, SIE_K_CORR_ROWS_MIR, (const void*)k_corr_rows<true, true>, rsmem)) return rc;
      k_corr_rows<true, true><<<grid, rw_threads<true>(), rsmem, st>>>(z, prefix, table, B, ldn, Tp, S, R, parts);
    } else if (R) {
      if (int rc = sie_ensure_smem(dev, SIE_K_CORR_ROWS_ST, (const void*)k_corr_rows<true, false>, rsmem)) return rc;
      k_corr_rows<true, false><<<grid, rw_threads<true>(), rsmem, st>>>(z, prefix, table, B, ldn, Tp, S, R, parts);
    } else {
      if (int rc = sie_ensure_smem(dev, SIE_K_CORR_ROWS_TAU, (const void*)k_corr_rows<false, false>, rsmem)) return rc;
      k_corr_rows<false, false><<<grid, rw_threads<false>(), rsmem, st>>>(z, prefix, table, B, ldn, Tp, S, R, parts);
    }
    SIE_CHECK_LAUNCH();
    k_tau_tiles<<<(unsigned)((max_items + 255) / 256), 256, 0, st>>>(parts, prefix, B, tile_pair);
    SIE_CHECK_LAUNCH();
  } else {
    if ((int)smem + 1024 > max_optin) {
      sie_set_error("sie_corr_tau: Tp=%d needs %zu B of shared memory (> %d)", Tp, smem, max_optin);
      return SIE_ERR_UNSUPPORTED;
    }
    if (int rc = sie_ensure_smem(dev, SIE_K_CORR_TILES, (const void*)k_corr_tiles, smem)) return rc;
    const long long slots = (long long)CTAS_PER_SM * sms;
    int grid = (int)(max_items < slots ? max_items : slots);
    k_corr_tiles<<<grid, NTHREADS, smem, st>>>(z, n_nodes, job_T, r_crit, prefix, table, B, ldn, Tp, R, tile_pair);
    SIE_CHECK_LAUNCH();
  }
  k_tau_finalize<<<B, 256, 0, st>>>(tile_pair, prefix, tau_sum, tau_cnt, tau);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}

extern "C" int sie_corr_stencil(const double* R, const double* z, const int32_t* job_T, int Tp,
                                const int32_t* node_cell, const int32_t* cell_node,
                                const int32_t* n_nodes, int B, int X, int Y, int ldn, int latlon,
                                double* stencil, void* stream) {
  SIE_CHECK_ARG(node_cell && cell_node && n_nodes && stencil, "null pointer");
  SIE_CHECK_ARG(R || (z && job_T && Tp > 0), "without R the unit-norm rows z, job_T and Tp are needed");
  SIE_CHECK_ARG(B > 0 && X > 0 && Y > 0 && ldn > 0, "non-positive size");
  const long long total = (long long)B * ldn * 4;
  k_stencil<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(R, z, job_T, Tp, node_cell, cell_node,
                                                                               n_nodes, B, X, Y, ldn, latlon, stencil);
  SIE_CHECK_LAUNCH();
  return SIE_OK;
}
