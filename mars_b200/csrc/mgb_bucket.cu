// libmarsgb: bucketed aggregation for inputs that do not pre-aggregate (high-cardinality map chunks, reducer
// side merges of many partial blocks) -- the second implementation behind mgb_agg_* (a4 / a9 / a10;
// reference: DataFrameGroupByAgg._execute_map/_execute_combine/_execute_agg,
// mars/dataframe/groupby/aggregation.py:1043-1267).
//
// The hash-table path (mgb_agg.cu) pays one atomic per (row, accumulator) into a table sized by the row
// count; at 40 accumulators per row (min/max/var over 8 columns) or ~1e7 distinct keys per chunk that is
// latency- and init-bound.  Here the ROWS of all sources (update() batches, read in place) are moved into bucket
// order -- B buckets of <= 512 rows on average, bucket = top bits of the key hash -- by a multi-way split in one or
// two levels of <= 256 ways each (mgb_split.cuh: tile histogram -> scan -> TMA-staged column tiles, permuted on chip,
// written in runs with 16-byte stores; every level reads and writes each row once, sequentially on both sides):
//
//   k_bsplit_hist<1> / scan / k_bsplit_plan / k_bsplit_scatter<1>
//                   level 1: rows of all sources by the top b1 hash bits into a column-major staging array in
//                   which every partition starts on a tile boundary (so that a tile of level 2 lies in ONE partition)
//   k_bsplit_hist<2> / scan / k_bsplit_scatter<2> / k_bsplit_off
//                   level 2: every partition by the next b2 bits; output compact, in bucket order
//   k_bucket_agg    one CTA per bucket, reading ITS rows as one contiguous range of every column: key table in
//                   shared memory (dense group ids), counting sort of the bucket's rows by group id, then one
//                   thread (one warp for groups of >= 32 rows) per (group, value column) folds that group's values
//                   into five canonical REGISTER accumulators (sum, sum of squares, count, encoded min / max) and
//                   stores every accumulator the spec asks for once.  No atomics on values.
//   scan + k_bucket_emit   compaction of the per-bucket results into the caller's columns
//
// (Round 1 moved row IDS into bucket order and gathered the rows at random from a row-major copy: 99 B of DRAM
// traffic per 16-byte row.)  Inside a bucket the rows arrive in no particular order (the split ranks them with one
// shared-memory atomic per row): the fold does not depend on it beyond the rounding of the compensated float64 sums.
//
// A bucket that holds more groups than its table, or one far longer than the average (one hot key: it would
// serialise on one CTA), makes the fold kernel raise its error word; the host reads it together with the group count
// and the caller takes the hash-table path.
#include <stdlib.h>

#include <vector>

#include "mgb_layout.cuh"
#include "mgb_split.cuh"

#ifndef BK_THREADS
#define BK_THREADS 256
#endif
#define BK_TAB 1024           // key table slots per bucket (power of two)
#define BK_MAXG 896           // groups per bucket before overflow is declared (load <= 7/8)
#define BK_RT 1024            // rows per tile
#define BK_TARGET 512         // average rows per bucket (upper bound)
#define BK_BIG 32             // groups with >= BK_BIG rows in a tile are folded by a whole warp
#define BK_MAXA 8             // accumulators per value column handled in registers
#define BK_SKEW_ROWS (64 * BK_TARGET)   // a longer bucket (one hot key) would serialise on one CTA: decline

struct BktParams {
  const uint64_t* __restrict__ rows;   // column-major staging in bucket order: column c (keys first) at rows + c * T
  long long T;
  int logB;
  const int64_t* off;
  uint64_t* scratch;
  int64_t* gcount;
  int* err;
  Layout L;
};

__device__ __forceinline__ int bk_locate(const long long* __restrict__ begin, int nsrc, long long i) {
  int lo = 0, hi = nsrc - 1;   // largest s with begin[s] <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&begin[mid]) <= i)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// ---- multi-way split of the rows into bucket order ---------------------------------------------------------------
struct BsSources {
  const BktSource* srcs;        // device array
  const long long* row_begin;   // device [nsrc + 1]: first global row of every source
  const long long* tile_begin;  // device [nsrc + 1]: first tile of every source (sources are tiled one by one)
  BktSource src0;               // copy of source 0 (single-source fast path)
  long long rows0;
  int nsrc;
};

struct BsLevel2 {
  const uint64_t* stage;        // level-1 output: column c at stage + c * stride, partitions padded to tiles
  long long stride;
  const long long* po;          // [P1 + 1] first (padded) row of every partition
  const long long* cnt;         // [P1] rows of every partition
  const int32_t* tile_part;     // [NT] partition of every tile, -1 beyond the last one
};

// where a tile's rows lie: column c starts at col(c); `rows` of them are valid
template <int LEVEL>
struct BsTileSrc {
  const BktSource* src;      // LEVEL 1: the update() batch the tile belongs to
  const uint64_t* base;      // LEVEL 2: level-1 staging, column c at base + c * stride
  long long stride, r0;
  int nk, rows;
  __device__ __forceinline__ const uint64_t* col(int c) const {
    if (LEVEL == 1) return (c < nk ? src->keys[c] : src->vals[c - nk]) + r0;
    return base + (long long)c * stride + r0;
  }
};

template <int LEVEL>
__device__ __forceinline__ BsTileSrc<LEVEL> bs_tile(const BsSources& S, const BsLevel2& L2, long long tile, int nk) {
  BsTileSrc<LEVEL> t;
  t.nk = nk;
  t.src = &S.src0;
  t.base = L2.stage;
  t.stride = L2.stride;
  long long left;
  if (LEVEL == 1) {
    long long lt = tile, rows = S.rows0;
    if (S.nsrc > 1) {
      const int s_ = bk_locate(S.tile_begin, S.nsrc, tile);
      t.src = &S.srcs[s_];
      lt = tile - __ldg(&S.tile_begin[s_]);
      rows = __ldg(&S.row_begin[s_ + 1]) - __ldg(&S.row_begin[s_]);
    }
    t.r0 = lt * SP_TILE;
    left = rows - t.r0;
  } else {
    const int p = L2.tile_part[tile];
    t.r0 = tile * SP_TILE;
    left = p < 0 ? 0 : L2.po[p] + L2.cnt[p] - t.r0;
  }
  t.rows = (int)(left < 0 ? 0 : (left < SP_TILE ? left : SP_TILE));
  return t;
}

template <int NK, int LEVEL>
__device__ __forceinline__ uint32_t bs_digit(const BsTileSrc<LEVEL>& t, int j, int shift, uint32_t mask) {
  int64_t k[NK];
#pragma unroll
  for (int q = 0; q < NK; ++q) k[q] = (int64_t)mgb_ld_stream(t.col(q) + j);
  return (uint32_t)(mgb_slot_hash(mgb_hash_tuple<NK>(k)) >> shift) & mask;
}

template <int NK, int LEVEL>
__global__ void __launch_bounds__(SP_THREADS) k_bsplit_hist(const BsSources S, const BsLevel2 L2, long long NT, int shift,
                                                            int R, int64_t* __restrict__ hist) {
  __shared__ uint32_t cnt[SP_MAX_R];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < SP_MAX_R) cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long tile = blockIdx.x;
  const BsTileSrc<LEVEL> ts = bs_tile<LEVEL>(S, L2, tile, NK);
  const int nt = ts.rows;
  const uint32_t mask = (uint32_t)R - 1u;
  uint32_t dg[SP_ROUNDS];
#pragma unroll
  for (int r = 0; r < SP_ROUNDS; ++r) {   // all key loads of the thread in flight together
    const int j = warp * (SP_TILE / SP_WARPS) + r * 32 + lane;
    dg[r] = j < nt ? bs_digit<NK, LEVEL>(ts, j, shift, mask) : 0xffffffffu;
  }
#pragma unroll
  for (int r = 0; r < SP_ROUNDS; ++r)
    if (dg[r] != 0xffffffffu) atomicAdd(&cnt[dg[r]], 1u);
  __syncthreads();
  for (int p = threadIdx.x; p < R; p += SP_THREADS) hist[(long long)p * NT + tile] = cnt[p];
}

// after the scan of the level-1 histogram: rows per partition, partitions padded to whole tiles, tile -> partition
__global__ void __launch_bounds__(256) k_bsplit_plan(const int64_t* __restrict__ hist, long long NT, int P1, long long T,
                                                     long long* __restrict__ po, long long* __restrict__ cnt,
                                                     int32_t* __restrict__ tile_part, long long NT2) {
  __shared__ long long tiles[257];
  const int p = threadIdx.x;
  long long c = 0, nt = 0;
  if (p < P1) {
    c = (p + 1 < P1 ? hist[(long long)(p + 1) * NT] : T) - hist[(long long)p * NT];
    nt = (c + SP_TILE - 1) / SP_TILE;
    cnt[p] = c;
  }
  tiles[p + 1] = nt;
  if (p == 0) tiles[0] = 0;
  __syncthreads();
  if (p == 0)
    for (int q = 1; q <= 256; ++q) tiles[q] += tiles[q - 1];
  __syncthreads();
  if (p < P1) {
    po[p] = tiles[p] * SP_TILE;
    for (long long t = tiles[p]; t < tiles[p + 1] && t < NT2; ++t) tile_part[t] = p;
  }
  if (p == 0) po[P1] = tiles[P1] * SP_TILE;
  const long long last = tiles[256];
  for (long long t = last + p; t < NT2; t += 256) tile_part[t] = -1;
}

template <int NK, int LEVEL>
struct BucketTile {
  BsTileSrc<LEVEL> ts;
  const int64_t* __restrict__ hist;
  const long long* __restrict__ po;   // level 1 with a second level to follow: padded partition starts; else nullptr
  uint64_t* out;
  long long out_stride, NT, tile;
  int shift;
  uint32_t mask;
  __device__ __forceinline__ uint32_t digit(int j) const { return bs_digit<NK, LEVEL>(ts, j, shift, mask); }
  __device__ __forceinline__ int64_t dest_row(int d) const {
    const int64_t abs0 = hist[(long long)d * NT + tile];
    return po ? po[d] + (abs0 - hist[(long long)d * NT]) : abs0;
  }
  __device__ __forceinline__ const uint64_t* src(int c) const { return ts.col(c); }
  __device__ __forceinline__ uint64_t dst(int c, int) const { return (uint64_t)(out + (long long)c * out_stride); }
};

template <int NK, int LEVEL>
__global__ void __launch_bounds__(SP_THREADS, 3) k_bsplit_scatter(const BsSources S, const BsLevel2 L2, long long NT,
                                                                  int shift, int R, int ncols,
                                                                  const int64_t* __restrict__ hist,
                                                                  const long long* __restrict__ po, uint64_t* out,
                                                                  long long out_stride) {
  extern __shared__ __align__(128) unsigned char sp_raw[];
  SplitSmem& sm = *reinterpret_cast<SplitSmem*>(sp_raw);
  BucketTile<NK, LEVEL> f;
  f.ts = bs_tile<LEVEL>(S, L2, blockIdx.x, NK);
  f.hist = hist;
  f.po = po;
  f.out = out;
  f.out_stride = out_stride;
  f.NT = NT;
  f.tile = blockIdx.x;
  f.shift = shift;
  f.mask = (uint32_t)R - 1u;
  if (f.ts.rows == 0) return;
  sp_tile_split<false>(sm, f.ts.rows, R, ncols, f);
}

// bucket boundaries.  One level: bucket b = digit b, off[b] = hist[b][0].  Two levels: the rows leave level 2 ordered
// by (digit 2, partition, tile), so bucket (d2, p) is contiguous: off[d2 * P1 + p] = hist2[d2][first tile of p].
__global__ void __launch_bounds__(256) k_bsplit_off(const int64_t* __restrict__ hist, long long NT, int P1, int B,
                                                    const long long* __restrict__ po, long long T,
                                                    int64_t* __restrict__ off) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b <= B; b += gridDim.x * blockDim.x) {
    if (b == B)
      off[b] = T;
    else if (!po)
      off[b] = hist[(long long)b * NT];
    else
      off[b] = hist[(long long)(b / P1) * NT + po[b % P1] / SP_TILE];
  }
}

// find-or-insert in the CTA's key table; tab_idx[slot]: -1 empty, -2 being written, -3 poisoned (overflow),
// >= 0 dense group id.  Called by ALL threads of a warp together (`valid` = this lane has a row): a lane that
// meets a slot in the "being written" state does not spin on it -- the writer may be a lane of the same warp
// that is parked at a reconvergence point -- it retries in the next round, after the whole warp has
// reconverged at the vote.  The inserting thread also stores the key tuple of the new group into the
// bucket's scratch rows.  Returns the group id, or -1 on overflow.
template <int NK>
__device__ __forceinline__ int bk_find_or_insert(uint64_t* tab_keys, int* tab_idx, int* n_unique, bool valid,
                                                 const uint64_t* k, uint32_t slot, uint64_t* scratch, long long T,
                                                 long long lo) {
  bool done = !valid;
  int g = -1;
  uint32_t probe = 0;
  while (__any_sync(0xffffffffu, !done)) {
    if (!done) {
      int cur = *(volatile int*)&tab_idx[slot];
      if (cur == -1) {
        cur = atomicCAS(&tab_idx[slot], -1, -2);
        if (cur == -1) {
          const int id = atomicAdd(n_unique, 1);
          if (id >= BK_MAXG) {
            *(volatile int*)&tab_idx[slot] = -3;   // the CTA aborts after the tile
          } else {
            tab_keys[slot] = k[0];
            if (NK == 2) tab_keys[BK_TAB + slot] = k[1];
            __threadfence_block();
            *(volatile int*)&tab_idx[slot] = id;
#pragma unroll
            for (int j = 0; j < NK; ++j) scratch[(long long)j * T + lo + id] = k[j];
            g = id;
          }
          done = true;
        }
      }
      if (!done && cur != -2) {   // cur == -2: busy, look again next round
        if (cur == -3) {
          done = true;
        } else {
          bool eq = ((volatile uint64_t*)tab_keys)[slot] == k[0];
          if (NK == 2) eq = eq && ((volatile uint64_t*)tab_keys)[BK_TAB + slot] == k[1];
          if (eq) {
            g = cur;
            done = true;
          } else {
            slot = (slot + 1) & (BK_TAB - 1);
            if (++probe == BK_TAB) done = true;
          }
        }
      }
    }
  }
  return g;
}

// s += v with the rounding error of the addition accumulated in c (Neumaier's variant of Kahan summation)
__device__ __forceinline__ void bk_comp_add(double& s, double& c, const double v) {
  const double t = s + v;
  c += fabs(s) >= fabs(v) ? (s - t) + v : (v - t) + s;
  s = t;
}

// NEED: which canonical accumulators some column of the spec asks for (1 = sum of squares, 2 = min / max); the
// others are not computed
template <int NK, int NEED, int TH>
__global__ void __launch_bounds__(TH, 1280 / TH) k_bucket_agg(const BktParams p) {
  __shared__ uint64_t tab_keys[NK * BK_TAB];
  __shared__ int tab_idx[BK_TAB];
  __shared__ int cnt[BK_TAB + 1];      // rows per group of the current tile -> exclusive starts
  __shared__ int cursor[BK_TAB];
  __shared__ uint16_t r_gid[BK_RT];
  __shared__ uint16_t sorted[BK_RT];
  __shared__ int warp_tot[TH / 32];
  __shared__ int big[BK_RT / BK_BIG];   // groups with >= BK_BIG rows in the current tile
  __shared__ int n_unique, failed, n_big;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = blockIdx.x;
  const long long lo = p.off[b], hi = p.off[b + 1];
  if (lo == hi) {
    if (tid == 0) p.gcount[b] = 0;
    return;
  }
  if (hi - lo > BK_SKEW_ROWS) {   // one hot key: this bucket would serialise on one CTA -- the whole input is declined
    if (tid == 0) {
      p.gcount[b] = 0;
      atomicMax(p.err, 2);
    }
    return;
  }
  const long long T = p.T;
  const Layout& L = p.L;
  const int nk = L.nk, nv = L.nv;
  for (int i = tid; i < BK_TAB; i += TH) tab_idx[i] = -1;
  if (tid == 0) {
    n_unique = 0;
    failed = 0;
  }
  __syncthreads();

  for (long long tlo = lo; tlo < hi; tlo += BK_RT) {
    const int m = (int)((hi - tlo) < BK_RT ? (hi - tlo) : BK_RT);
    const int ng_before = n_unique;
    for (int i = tid; i <= BK_TAB; i += TH) cnt[i] = 0;
    if (tid == 0) n_big = 0;
    __syncthreads();
    // 1. rows of the tile: key gather, group id, histogram by group
    for (int r0 = 0; r0 < m; r0 += TH) {   // uniform trip count: the table insert votes per warp
      const int r = r0 + tid;
      const bool valid = r < m;
      uint64_t k[NK];
#pragma unroll
      for (int j = 0; j < NK; ++j) k[j] = valid ? mgb_ld_stream(p.rows + (long long)j * T + tlo + r) : 0;
      const uint64_t sh = mgb_slot_hash(mgb_hash_tuple<NK>((const int64_t*)k));
      const uint32_t slot = (uint32_t)(sh >> (54 - p.logB)) & (BK_TAB - 1);
      const int g = bk_find_or_insert<NK>(tab_keys, tab_idx, &n_unique, valid, k, slot, p.scratch, T, lo);
      if (valid) {
        if (g < 0) {
          failed = 1;
        } else {
          r_gid[r] = (uint16_t)g;
          atomicAdd(&cnt[g], 1);
        }
      }
    }
    __syncthreads();
    if (failed) {
      if (tid == 0) atomicMax(p.err, 1);
      return;
    }
    const int ng = n_unique;
    // 2. exclusive scan of cnt[0..ng) ((BK_TAB / TH) consecutive entries per thread)
    {
      const int base = tid * (BK_TAB / TH);
      int v[(BK_TAB / TH)];
      int tot = 0;
#pragma unroll
      for (int q = 0; q < (BK_TAB / TH); ++q) {
        v[q] = base + q < ng ? cnt[base + q] : 0;
        if (v[q] >= BK_BIG) big[atomicAdd(&n_big, 1)] = base + q;
        tot += v[q];
      }
      int x = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) warp_tot[warp] = x;
      __syncthreads();
      int e = x - tot;
      for (int w = 0; w < warp; ++w) e += warp_tot[w];
#pragma unroll
      for (int q = 0; q < (BK_TAB / TH); ++q) {
        if (base + q < ng) {
          cnt[base + q] = e;
          cursor[base + q] = e;
        }
        e += v[q];
      }
      if (tid == 0) cnt[ng] = m;
    }
    __syncthreads();
    // 3. counting sort of the tile's rows by group
    for (int r = tid; r < m; r += TH) sorted[atomicAdd(&cursor[r_gid[r]], 1)] = (uint16_t)r;
    __syncthreads();
    // 4. fold: one (group, value column) per thread / per warp
    // pass 0: one thread per (group, column) with < BK_BIG rows; pass 1: one warp per (big group, column)
    const int nb = n_big;
    const uint32_t inv_ng = 0xffffffffu / (uint32_t)ng + 1u;   // idx / ng == umulhi(idx, inv_ng) for idx < 2^22
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1 && nb == 0) break;
      const int n_items = (pass == 0 ? ng : nb) * nv;
      const int per = pass == 0 ? ng : nb;
      const uint32_t inv = pass == 0 ? inv_ng : 0xffffffffu / (uint32_t)nb + 1u;
      const int first = pass == 0 ? tid : warp;
      const int step = pass == 0 ? TH : TH / 32;
      for (int idx = first; idx < n_items; idx += step) {
        const int c = per == 1 ? idx : (int)__umulhi((uint32_t)idx, inv);
        const int gi = idx - c * per;
        const int g = pass == 0 ? gi : big[gi];
        const int st = cnt[g], n = cnt[g + 1] - st;
        if (n == 0 || (pass == 0 && n >= BK_BIG)) continue;
        const int a0 = L.col_begin[c], na_c = L.col_begin[c + 1] - a0;
        if (na_c == 0) continue;
        const int dt = L.acc[a0].dt;
        // the five canonical accumulators of the column, in registers: sum, sum of squares, count, and the
        // sortable encodings of min / max (table representation, mgb_common.cuh); whichever the spec asks
        // for is picked at the end.  Loads are issued four at a time (they are independent).
        const uint64_t* col0 = p.rows + (long long)(nk + c) * T + tlo;      // value c of row r of the tile at col0[r]
        const int j0 = pass == 0 ? 0 : lane, jstep = pass == 0 ? 1 : 32;
        // float64 sums are COMPENSATED (Neumaier: running sum + running error term): pandas' group_sum is a
        // Kahan sum (pandas/_libs/groupby.pyx), so the reference's partial sums are the correctly rounded true
        // sums except on a set of measure ~1e-15; a plain DADD chain differs from that in the last bit for a few
        // groups per million, which the sum-of-squares variance formula of the reference (var.py:38-48) turns
        // into relative errors beyond 1e-9 when a group's variance is tiny.  Squares are rounded before they are
        // added (no FMA), like `x ** 2` followed by the sum.
        double fs = 0.0, fq = 0.0, fsc = 0.0, fqc = 0.0;
        unsigned long long is = 0, iq = 0, cn = 0, mn = ~0ULL, mx = 0ULL;
        const bool f64 = dt == MGB_F64;
        for (int j = j0; j < n; j += 4 * jstep) {
          uint64_t x[4];
          bool have[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int jj = j + u * jstep;
            have[u] = jj < n;
            x[u] = have[u] ? __ldg(col0 + sorted[st + jj]) : 0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (!have[u]) continue;
            if (f64) {
              if (mgb_is_nan_bits(x[u])) continue;
              const double v = __longlong_as_double((long long)x[u]);
              bk_comp_add(fs, fsc, v);
              if (NEED & 1) bk_comp_add(fq, fqc, __dmul_rn(v, v));
              if (NEED & 2) {
                const uint64_t e = mgb_enc_f64(x[u]);
                mn = e < mn ? e : mn;
                mx = e > mx ? e : mx;
              }
            } else {
              is += x[u];
              if (NEED & 1) iq += x[u] * x[u];
              if (NEED & 2) {
                const uint64_t e = mgb_enc_i64(x[u]);
                mn = e < mn ? e : mn;
                mx = e > mx ? e : mx;
              }
            }
            ++cn;
          }
        }
        if (pass == 1) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double os = __shfl_down_sync(0xffffffffu, fs, o), osc = __shfl_down_sync(0xffffffffu, fsc, o);
            const double oq = (NEED & 1) ? __shfl_down_sync(0xffffffffu, fq, o) : 0.0;
            const double oqc = (NEED & 1) ? __shfl_down_sync(0xffffffffu, fqc, o) : 0.0;
            bk_comp_add(fs, fsc, os);
            fsc += osc;
            is += __shfl_down_sync(0xffffffffu, is, o);
            cn += __shfl_down_sync(0xffffffffu, cn, o);
            if (NEED & 1) {
              bk_comp_add(fq, fqc, oq);
              fqc += oqc;
              iq += __shfl_down_sync(0xffffffffu, iq, o);
            }
            if (NEED & 2) {
              const unsigned long long a = __shfl_down_sync(0xffffffffu, mn, o);
              const unsigned long long z = __shfl_down_sync(0xffffffffu, mx, o);
              mn = a < mn ? a : mn;
              mx = z > mx ? z : mx;
            }
          }
          if (lane != 0) continue;
        }
        // fold the error terms in (a NaN error term -- inf - inf on the way -- is dropped, as pandas does)
        if (fsc == fsc) fs += fsc;
        if (fqc == fqc) fq += fqc;
#pragma unroll
        for (int q = 0; q < BK_MAXA; ++q) {
          if (q < na_c) {
            const int op = L.acc[a0 + q].op;
            uint64_t v;
            if (op == MGB_SUM)
              v = f64 ? (uint64_t)__double_as_longlong(fs) : (uint64_t)is;
            else if (op == MGB_SUMSQ)
              v = f64 ? (uint64_t)__double_as_longlong(fq) : (uint64_t)iq;
            else if (op == MGB_COUNT)
              v = (uint64_t)cn;
            else
              v = op == MGB_MIN ? (uint64_t)mn : (uint64_t)mx;
            uint64_t* dst = p.scratch + (long long)(nk + L.acc_index[a0 + q]) * T + lo + g;
            *dst = g >= ng_before ? v : mgb_combine(L.acc[a0 + q].kind, *dst, v);
          }
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) p.gcount[b] = n_unique;
}

struct BktEmitCols {
  uint64_t* key[MGB_MAX_KEYS];
  uint64_t* acc[MGB_MAX_ACCS];
};

// one warp per bucket: scratch rows [off[b], off[b] + n_b) -> output rows [goff[b], goff[b] + n_b), decoded
__global__ void __launch_bounds__(256) k_bucket_emit(Layout L, const uint64_t* __restrict__ scratch, long long T,
                                                     const int64_t* __restrict__ off,
                                                     const int64_t* __restrict__ goff, int B, BktEmitCols out) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += nwarps) {
    const long long g0 = goff[b];
    const int n = (int)(goff[b + 1] - g0);
    if (n == 0) continue;
    const long long lo = off[b];
    for (int j = 0; j < L.nk; ++j)
      for (int g = lane; g < n; g += 32) out.key[j][g0 + g] = scratch[(long long)j * T + lo + g];
    for (int a = 0; a < L.na; ++a) {
      const int op = L.spec_op[a], dt = L.spec_dt[a];
      for (int g = lane; g < n; g += 32)
        out.acc[a][g0 + g] = mgb_acc_decode(op, dt, scratch[(long long)(L.nk + a) * T + lo + g]);
    }
  }
}

bool mgb_bucket_supports(const Layout& L) {
  for (int c = 0; c < L.nv; ++c)
    if (L.col_begin[c + 1] - L.col_begin[c] > BK_MAXA) return false;
  return true;
}

void mgb_bucket_free(BktResult* r, cudaStream_t st) {
  if (r->scratch) cudaFreeAsync(r->scratch, st);
  if (r->off) cudaFreeAsync(r->off, st);
  if (r->goff) cudaFreeAsync(r->goff, st);
  r->scratch = nullptr;
  r->off = nullptr;
  r->goff = nullptr;
}

static void bk_debug(const char* what, cudaStream_t st) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MGB_BUCKET_DEBUG");
    on = (e && *e && *e != '0') ? 1 : 0;
  }
  if (!on) return;
  fprintf(stderr, "[mgb bucket] %s ...", what);
  fflush(stderr);
  cudaError_t e = cudaStreamSynchronize(st);
  fprintf(stderr, " %s\n", cudaGetErrorString(e));
  fflush(stderr);
}

template <int NK>
static int bs_launch_level(int level, const BsSources& S, const BsLevel2& L2, long long NT, int shift, int R, int ncols,
                           int64_t* hist, const long long* po, uint64_t* out, long long out_stride, cudaStream_t st) {
  static bool attr_set[8] = {false};
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  const size_t smem = sizeof(SplitSmem);
  if (dev >= 0 && dev < 8 && !attr_set[dev]) {
    MGB_CUDA(cudaFuncSetAttribute(k_bsplit_scatter<NK, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MGB_CUDA(cudaFuncSetAttribute(k_bsplit_scatter<NK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev] = true;
  }
  if (level == 1)
    k_bsplit_scatter<NK, 1><<<(unsigned)NT, SP_THREADS, smem, st>>>(S, L2, NT, shift, R, ncols, hist, po, out, out_stride);
  else
    k_bsplit_scatter<NK, 2><<<(unsigned)NT, SP_THREADS, smem, st>>>(S, L2, NT, shift, R, ncols, hist, po, out, out_stride);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

template <int NK>
static int bs_launch_hist(int level, const BsSources& S, const BsLevel2& L2, long long NT, int shift, int R, int64_t* hist,
                          cudaStream_t st) {
  if (level == 1)
    k_bsplit_hist<NK, 1><<<(unsigned)NT, SP_THREADS, 0, st>>>(S, L2, NT, shift, R, hist);
  else
    k_bsplit_hist<NK, 2><<<(unsigned)NT, SP_THREADS, 0, st>>>(S, L2, NT, shift, R, hist);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

int mgb_bucket_aggregate(const Layout& L, const BktSource* srcs_host, int nsrc, long long T, cudaStream_t st,
                         BktResult* out, int* overflow) {
  *overflow = 0;
  memset(out, 0, sizeof(*out));
  MGB_REQUIRE(nsrc >= 1 && nsrc < 65536 && T >= 1 && T < 0x7fffffffLL, MGB_ERR_UNSUPPORTED,
              "bucketed aggregation: %d sources, %lld rows", nsrc, T);
  int logB = 0;
  while (logB < 22 && (T >> logB) > BK_TARGET) ++logB;
  if (logB > 16) {   // more than two levels of 256 ways: left to the table path
    *overflow = 3;
    return MGB_OK;
  }
  const int B = 1 << logB;
  const int b1 = logB <= 8 ? logB : (logB + 1) / 2, b2 = logB - b1;
  const int P1 = 1 << b1, R2 = 1 << b2;
  const int ncol = L.nk + L.na, W = L.nk + L.nv;
  // sources are tiled one by one
  std::vector<long long> row_begin((size_t)nsrc + 1), tile_begin((size_t)nsrc + 1);
  tile_begin[0] = 0;
  for (int s = 0; s < nsrc; ++s) {
    row_begin[s] = srcs_host[s].begin;
    const long long rows = (s + 1 < nsrc ? srcs_host[s + 1].begin : T) - srcs_host[s].begin;
    tile_begin[s + 1] = tile_begin[s] + (rows + SP_TILE - 1) / SP_TILE;
  }
  row_begin[nsrc] = T;
  const long long NT1 = tile_begin[nsrc] > 0 ? tile_begin[nsrc] : 1;
  const long long NT2 = b2 ? (T + SP_TILE - 1) / SP_TILE + P1 + 1 : 0;   // padded tiles of level 2 (upper bound)
  const long long SA = NT2 * SP_TILE;                                     // rows per column of the padded staging
  const size_t src_bytes = sizeof(BktSource) * (size_t)nsrc;
  const size_t tab_bytes = sizeof(long long) * ((size_t)nsrc + 1);
  const size_t src_pad = (src_bytes + 255) & ~(size_t)255, tab_pad = (tab_bytes + 255) & ~(size_t)255;
  const size_t plan_bytes = (sizeof(long long) * (2 * (size_t)P1 + 2) + sizeof(int32_t) * (size_t)(NT2 + 1) + 255) & ~(size_t)255;
  char* small = nullptr;     // srcs | row_begin | tile_begin | err | po | cnt | tile_part
  uint64_t *stageA = nullptr, *stageB = nullptr;
  int64_t *hist1 = nullptr, *hist2 = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&small, src_pad + 2 * tab_pad + 256 + plan_bytes, st));
  MGB_CUDA(cudaMallocAsync((void**)&stageB, sizeof(uint64_t) * (size_t)W * (size_t)T, st));
  MGB_CUDA(cudaMallocAsync((void**)&hist1, sizeof(int64_t) * (size_t)P1 * (size_t)NT1, st));
  if (b2) {
    MGB_CUDA(cudaMallocAsync((void**)&stageA, sizeof(uint64_t) * (size_t)W * (size_t)SA, st));
    MGB_CUDA(cudaMallocAsync((void**)&hist2, sizeof(int64_t) * (size_t)R2 * (size_t)NT2, st));
  }
  MGB_CUDA(cudaMallocAsync((void**)&out->off, sizeof(int64_t) * ((size_t)B + 1), st));
  MGB_CUDA(cudaMallocAsync((void**)&out->goff, sizeof(int64_t) * ((size_t)B + 1), st));
  MGB_CUDA(cudaMallocAsync((void**)&out->scratch, sizeof(uint64_t) * (size_t)ncol * (size_t)T, st));
  out->T = T;
  out->B = B;
  BktSource* d_srcs = (BktSource*)small;
  long long* d_row_begin = (long long*)(small + src_pad);
  long long* d_tile_begin = (long long*)(small + src_pad + tab_pad);
  int* d_err = (int*)(small + src_pad + 2 * tab_pad);
  long long* d_po = (long long*)(small + src_pad + 2 * tab_pad + 256);
  long long* d_cnt = d_po + P1 + 1;
  int32_t* d_tile_part = (int32_t*)(d_cnt + P1 + 1);
  // pageable sources: the runtime stages these small copies before cudaMemcpyAsync returns
  MGB_CUDA(cudaMemcpyAsync(d_srcs, srcs_host, src_bytes, cudaMemcpyHostToDevice, st));
  MGB_CUDA(cudaMemcpyAsync(d_row_begin, row_begin.data(), tab_bytes, cudaMemcpyHostToDevice, st));
  MGB_CUDA(cudaMemcpyAsync(d_tile_begin, tile_begin.data(), tab_bytes, cudaMemcpyHostToDevice, st));
  MGB_CUDA(cudaMemsetAsync(d_err, 0, 256, st));
  MGB_CUDA(cudaMemsetAsync(out->goff + B, 0, sizeof(int64_t), st));
  auto release = [&]() {
    if (stageA) cudaFreeAsync(stageA, st);
    cudaFreeAsync(stageB, st);
    cudaFreeAsync(hist1, st);
    if (hist2) cudaFreeAsync(hist2, st);
    cudaFreeAsync(small, st);
  };
  int status = MGB_OK;
  {
    MgbProfScope prof(MGB_K_BUCKET, st);
    BsSources S;
    S.srcs = d_srcs;
    S.row_begin = d_row_begin;
    S.tile_begin = d_tile_begin;
    S.src0 = srcs_host[0];
    S.rows0 = row_begin[1] - row_begin[0];
    S.nsrc = nsrc;
    BsLevel2 L2;
    memset(&L2, 0, sizeof(L2));
    const bool one = L.nk == 1;
    // level 1: by the top b1 bits.  One level: compact into stageB; two levels: into stageA, partitions padded to tiles
    const int shift1 = b1 ? 64 - b1 : 63;   // b1 == 0: one bucket (mask 0 makes the digit 0)
    status = one ? bs_launch_hist<1>(1, S, L2, NT1, shift1, P1, hist1, st) : bs_launch_hist<2>(1, S, L2, NT1, shift1, P1, hist1, st);
    if (status == MGB_OK) status = mgb_scan_exclusive_i64(hist1, (int64_t)P1 * NT1, nullptr, st);
    if (status == MGB_OK && b2) {
      k_bsplit_plan<<<1, 256, 0, st>>>(hist1, NT1, P1, T, d_po, d_cnt, d_tile_part, NT2);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) {
        mgb_set_error("bucket split plan launch: %s", cudaGetErrorString(e));
        status = MGB_ERR_CUDA;
      }
    }
    if (status == MGB_OK)
      status = one ? bs_launch_level<1>(1, S, L2, NT1, shift1, P1, W, hist1, b2 ? d_po : nullptr, b2 ? stageA : stageB, b2 ? SA : T, st)
                   : bs_launch_level<2>(1, S, L2, NT1, shift1, P1, W, hist1, b2 ? d_po : nullptr, b2 ? stageA : stageB, b2 ? SA : T, st);
    bk_debug("bucket split level 1", st);
    if (status == MGB_OK && b2) {   // level 2: every partition by the next b2 bits, compact, bucket order (digit 2, partition)
      L2.stage = stageA;
      L2.stride = SA;
      L2.po = d_po;
      L2.cnt = d_cnt;
      L2.tile_part = d_tile_part;
      const int shift2 = 64 - logB;
      status = one ? bs_launch_hist<1>(2, S, L2, NT2, shift2, R2, hist2, st) : bs_launch_hist<2>(2, S, L2, NT2, shift2, R2, hist2, st);
      if (status == MGB_OK) status = mgb_scan_exclusive_i64(hist2, (int64_t)R2 * NT2, nullptr, st);
      if (status == MGB_OK)
        status = one ? bs_launch_level<1>(2, S, L2, NT2, shift2, R2, W, hist2, nullptr, stageB, T, st)
                     : bs_launch_level<2>(2, S, L2, NT2, shift2, R2, W, hist2, nullptr, stageB, T, st);
      bk_debug("bucket split level 2", st);
    }
    if (status == MGB_OK) {
      const int gb = (B + 256) / 256;
      k_bsplit_off<<<gb > 1024 ? 1024 : gb, 256, 0, st>>>(b2 ? hist2 : hist1, b2 ? NT2 : NT1, P1, B, b2 ? d_po : nullptr, T,
                                                          out->off);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) {
        mgb_set_error("bucket offsets launch: %s", cudaGetErrorString(e));
        status = MGB_ERR_CUDA;
      }
    }
    // (a hot key makes one bucket as long as its row count: the CTA that meets such a bucket raises err = 2 and the
    // input is left to the table -- checked with the group count below, no extra host synchronisation)
    // one CTA per bucket
    if (status == MGB_OK) {
      BktParams p;
      p.rows = stageB;
      p.T = T;
      p.logB = logB;
      p.off = out->off;
      p.scratch = out->scratch;
      p.gcount = out->goff;
      p.err = d_err;
      p.L = L;
      int need = 0;
      for (int a = 0; a < L.na; ++a) need |= L.spec_op[a] == MGB_SUMSQ ? 1 : ((L.spec_op[a] == MGB_MIN || L.spec_op[a] == MGB_MAX) ? 2 : 0);
      // narrow rows (few value columns): 128-thread CTAs, twice as many resident -- the phases of a bucket are
      // separated by CTA barriers; wide rows keep 256 threads for the (group, column) fold
#define BK_LAUNCH(NK_, NEED_)                                       \
  if (L.nv <= 3)                                                    \
    k_bucket_agg<NK_, NEED_, 128><<<(unsigned)B, 128, 0, st>>>(p);  \
  else                                                              \
    k_bucket_agg<NK_, NEED_, 256><<<(unsigned)B, 256, 0, st>>>(p)
#define BK_PICK(NK_)                 \
  switch (need) {                    \
    case 0: BK_LAUNCH(NK_, 0); break;  \
    case 1: BK_LAUNCH(NK_, 1); break;  \
    case 2: BK_LAUNCH(NK_, 2); break;  \
    default: BK_LAUNCH(NK_, 3); break; \
  }
      if (L.nk == 1) {
        BK_PICK(1)
      } else {
        BK_PICK(2)
      }
#undef BK_PICK
#undef BK_LAUNCH
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) {
        mgb_set_error("bucket aggregation launch: %s", cudaGetErrorString(e));
        status = MGB_ERR_CUDA;
      }
    }
    bk_debug("k_bucket_agg", st);
    if (status == MGB_OK) status = mgb_scan_exclusive_i64(out->goff, (int64_t)B + 1, nullptr, st);
    bk_debug("scan of group counts", st);
  }
  int err_h = 0;
  int64_t G = 0;
  if (status == MGB_OK) {
    cudaError_t e = cudaMemcpyAsync(&err_h, d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&G, out->goff + B, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      mgb_set_error("bucket aggregation: %s", cudaGetErrorString(e));
      status = MGB_ERR_CUDA;
    }
  }
  release();
  if (status != MGB_OK || err_h) {
    mgb_bucket_free(out, st);
    *overflow = err_h;      // 1 = a bucket held more groups than its table, 2 = a bucket far longer than the average
    return status;
  }
  out->G = (long long)G;
  return MGB_OK;
}

int mgb_bucket_emit(const Layout& L, const BktResult& r, uint64_t* const* out_key, uint64_t* const* out_acc,
                    cudaStream_t st) {
  BktEmitCols out;
  memset(&out, 0, sizeof(out));
  for (int j = 0; j < L.nk; ++j) out.key[j] = out_key[j];
  for (int a = 0; a < L.na; ++a) out.acc[a] = out_acc[a];
  long long blocks = ((long long)r.B + 7) / 8, capb = (long long)mgb_sm_count() * 16;
  MgbProfScope prof(MGB_K_EMIT, st);
  k_bucket_emit<<<(unsigned)(blocks > capb ? capb : blocks), 256, 0, st>>>(L, r.scratch, r.T, r.off, r.goff, r.B,
                                                                          out);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}
