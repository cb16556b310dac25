// libmarsgb: bucketed aggregation for inputs that do not pre-aggregate (high-cardinality map chunks, reducer
// side merges of many partial blocks) -- the second implementation behind mgb_agg_* (a4 / a9 / a10;
// reference: DataFrameGroupByAgg._execute_map/_execute_combine/_execute_agg,
// mars/dataframe/groupby/aggregation.py:1043-1267).
//
// The hash-table path (mgb_agg.cu) pays one atomic per (row, accumulator) into a table sized by the row
// count; at 40 accumulators per row (min/max/var over 8 columns) or ~1e7 distinct keys per chunk that is
// latency- and init-bound.  Here the rows of ALL sources (update() batches, read in place) are split by the
// top bits of the key hash into B buckets of <= 512 rows on average:
//
//   k_bucket_pack   one sequential pass over the column-major batches: row -> bucket id, and the row itself
//                   (keys + values) into a ROW-MAJOR staging array through a shared-memory transpose.  The
//                   gathers of the next kernel then fetch one row = one or two 64-byte DRAM bursts instead
//                   of one burst per 8-byte column value (measured: 1118 -> ~190 B/row on 80-byte rows).
//   k_bucket_hist / scan / k_bucket_place
//                   row ids into bucket order through one atomic cursor per bucket (the order of the rows
//                   inside a bucket does not matter to the fold, so no stable radix passes are needed)
//   k_bucket_agg    one CTA per bucket: key table in shared memory (dense group ids), counting sort of the
//                   bucket's rows by group id, then one thread (one warp for groups of >= 32 rows) per
//                   (group, value column) folds that group's values into five canonical REGISTER accumulators
//                   (sum, sum of squares, count, encoded min / max) and stores every accumulator the spec
//                   asks for once.  No atomics on values; buckets of any length are processed in tiles.
//   scan + k_bucket_emit   compaction of the per-bucket results into the caller's columns
//
// A bucket that holds more groups than its table reports overflow, a bucket far longer than the average (one
// hot key) is declined before the aggregation kernel runs; the caller then takes the hash-table path.
#include <stdlib.h>

#include <vector>

#include "mgb_layout.cuh"

#ifndef BK_THREADS
#define BK_THREADS 256
#endif
#define BK_TAB 1024           // key table slots per bucket (power of two)
#define BK_MAXG 896           // groups per bucket before overflow is declared (load <= 7/8)
#define BK_RT 1024            // rows per tile
#define BK_TARGET 512         // average rows per bucket (upper bound)
#define BK_BIG 32             // groups with >= BK_BIG rows in a tile are folded by a whole warp
#define BK_EPT (BK_TAB / BK_THREADS)   // table entries per thread in the block scan
#define BK_MAXA 8             // accumulators per value column handled in registers
#define BK_SKEW_ROWS (64 * BK_TARGET)   // a longer bucket (one hot key) would serialise on one CTA: decline

struct BktParams {
  const uint64_t* __restrict__ rows;   // row-major staging: row i at rows + i * W, keys first, then the value columns
  int W;
  long long T;
  int logB;
  const int32_t* perm;
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

#define BKP_ROWS 256   // rows per tile of the pack kernel (== its block size)

struct BktPackParams {
  const BktSource* srcs;      // device array
  const long long* begin;     // device [nsrc + 1]
  BktSource src0;             // copy of source 0 (single-source fast path)
  int nsrc, nk, nv;
  long long T;
  int logB;
  int32_t* pid;
  uint64_t* rows;
};

// Tile of 256 rows: every thread reads its row column by column (coalesced across the warp), the tile is
// transposed through shared memory (row stride padded to an odd number of words: no bank conflicts) and
// written out as 256 * W contiguous words.
template <int NK>
__global__ void __launch_bounds__(BKP_ROWS) k_bucket_pack(const BktPackParams p) {
  extern __shared__ uint64_t tile[];
  const int W = p.nk + p.nv, Wp = W | 1;
  const long long ntiles = (p.T + BKP_ROWS - 1) / BKP_ROWS;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long i = t * BKP_ROWS + threadIdx.x;
    if (i < p.T) {
      const BktSource* src = &p.src0;
      long long loc = i;
      if (p.nsrc > 1) {
        const int s = bk_locate(p.begin, p.nsrc, i);
        src = &p.srcs[s];
        loc = i - __ldg(&p.begin[s]);
      }
      uint64_t* row = tile + threadIdx.x * Wp;
      int64_t k[NK];
#pragma unroll
      for (int j = 0; j < NK; ++j) {
        const uint64_t v = mgb_ld_stream(src->keys[j] + loc);
        k[j] = (int64_t)v;
        row[j] = v;
      }
      const uint64_t sh = mgb_slot_hash(mgb_hash_tuple<NK>(k));
      p.pid[i] = p.logB ? (int32_t)(sh >> (64 - p.logB)) : 0;
#pragma unroll 4
      for (int c = 0; c < p.nv; ++c) row[NK + c] = mgb_ld_stream(src->vals[c] + loc);
    }
    __syncthreads();
    const long long base = t * BKP_ROWS;
    const int nrow = (int)((p.T - base) < BKP_ROWS ? (p.T - base) : BKP_ROWS);
    const int nword = nrow * W;
    const uint32_t inv = 0xffffffffu / (uint32_t)W + 1u;
    uint64_t* out = p.rows + base * W;
    for (int idx = threadIdx.x; idx < nword; idx += BKP_ROWS) {
      const int r = W == 1 ? idx : (int)__umulhi((uint32_t)idx, inv);
      out[idx] = tile[r * Wp + (idx - r * W)];
    }
    __syncthreads();
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

template <int NK>
__global__ void __launch_bounds__(BK_THREADS, 5) k_bucket_agg(const BktParams p) {
  __shared__ uint64_t tab_keys[NK * BK_TAB];
  __shared__ int tab_idx[BK_TAB];
  __shared__ int cnt[BK_TAB + 1];      // rows per group of the current tile -> exclusive starts
  __shared__ int cursor[BK_TAB];
  __shared__ uint32_t r_loc[BK_RT];
  __shared__ uint16_t r_gid[BK_RT];
  __shared__ uint16_t sorted[BK_RT];
  __shared__ int warp_tot[BK_THREADS / 32];
  __shared__ int big[BK_RT / BK_BIG];   // groups with >= BK_BIG rows in the current tile
  __shared__ int n_unique, failed, n_big;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = blockIdx.x;
  const long long lo = p.off[b], hi = p.off[b + 1];
  if (lo == hi) {
    if (tid == 0) p.gcount[b] = 0;
    return;
  }
  const long long T = p.T;
  const Layout& L = p.L;
  const int nk = L.nk, nv = L.nv;
  for (int i = tid; i < BK_TAB; i += BK_THREADS) tab_idx[i] = -1;
  if (tid == 0) {
    n_unique = 0;
    failed = 0;
  }
  __syncthreads();

  for (long long tlo = lo; tlo < hi; tlo += BK_RT) {
    const int m = (int)((hi - tlo) < BK_RT ? (hi - tlo) : BK_RT);
    const int ng_before = n_unique;
    for (int i = tid; i <= BK_TAB; i += BK_THREADS) cnt[i] = 0;
    if (tid == 0) n_big = 0;
    __syncthreads();
    // 1. rows of the tile: key gather, group id, histogram by group
    for (int r0 = 0; r0 < m; r0 += BK_THREADS) {   // uniform trip count: the table insert votes per warp
      const int r = r0 + tid;
      const bool valid = r < m;
      const long long i = valid ? (long long)p.perm[tlo + r] : 0;
      uint64_t k[NK];
#pragma unroll
      for (int j = 0; j < NK; ++j) k[j] = valid ? __ldg(p.rows + i * p.W + j) : 0;
      const uint64_t sh = mgb_slot_hash(mgb_hash_tuple<NK>((const int64_t*)k));
      const uint32_t slot = (uint32_t)(sh >> (54 - p.logB)) & (BK_TAB - 1);
      const int g = bk_find_or_insert<NK>(tab_keys, tab_idx, &n_unique, valid, k, slot, p.scratch, T, lo);
      if (valid) {
        if (g < 0) {
          failed = 1;
        } else {
          r_gid[r] = (uint16_t)g;
          r_loc[r] = (uint32_t)i;
          atomicAdd(&cnt[g], 1);
        }
      }
    }
    __syncthreads();
    if (failed) {
      if (tid == 0) atomicExch(p.err, 1);
      return;
    }
    const int ng = n_unique;
    // 2. exclusive scan of cnt[0..ng) (BK_EPT consecutive entries per thread)
    {
      const int base = tid * BK_EPT;
      int v[BK_EPT];
      int tot = 0;
#pragma unroll
      for (int q = 0; q < BK_EPT; ++q) {
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
      for (int q = 0; q < BK_EPT; ++q) {
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
    for (int r = tid; r < m; r += BK_THREADS) sorted[atomicAdd(&cursor[r_gid[r]], 1)] = (uint16_t)r;
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
      const int step = pass == 0 ? BK_THREADS : BK_THREADS / 32;
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
        const uint64_t* col0 = p.rows + nk + c;      // value c of row i at col0[i * W]
        const long long W = p.W;
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
            x[u] = have[u] ? col0[(long long)r_loc[sorted[st + jj]] * W] : 0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (!have[u]) continue;
            if (f64) {
              if (mgb_is_nan_bits(x[u])) continue;
              const double v = __longlong_as_double((long long)x[u]);
              bk_comp_add(fs, fsc, v);
              bk_comp_add(fq, fqc, __dmul_rn(v, v));
              const uint64_t e = mgb_enc_f64(x[u]);
              mn = e < mn ? e : mn;
              mx = e > mx ? e : mx;
            } else {
              is += x[u];
              iq += x[u] * x[u];
              const uint64_t e = mgb_enc_i64(x[u]);
              mn = e < mn ? e : mn;
              mx = e > mx ? e : mx;
            }
            ++cn;
          }
        }
        if (pass == 1) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double os = __shfl_down_sync(0xffffffffu, fs, o), osc = __shfl_down_sync(0xffffffffu, fsc, o);
            const double oq = __shfl_down_sync(0xffffffffu, fq, o), oqc = __shfl_down_sync(0xffffffffu, fqc, o);
            bk_comp_add(fs, fsc, os);
            fsc += osc;
            bk_comp_add(fq, fqc, oq);
            fqc += oqc;
            is += __shfl_down_sync(0xffffffffu, is, o);
            iq += __shfl_down_sync(0xffffffffu, iq, o);
            cn += __shfl_down_sync(0xffffffffu, cn, o);
            const unsigned long long a = __shfl_down_sync(0xffffffffu, mn, o);
            const unsigned long long z = __shfl_down_sync(0xffffffffu, mx, o);
            mn = a < mn ? a : mn;
            mx = z > mx ? z : mx;
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

// bucket histogram and (unordered) placement of the row ids: within a bucket the order of the rows does not
// matter to the fold, so one pass with an atomic cursor per bucket replaces the stable radix passes
__global__ void __launch_bounds__(256) k_bucket_hist(const int32_t* __restrict__ pid, long long T,
                                                     unsigned long long* __restrict__ counts) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < T; i += (long long)gridDim.x * blockDim.x)
    atomicAdd(&counts[pid[i]], 1ULL);
}

__global__ void __launch_bounds__(256) k_bucket_place(const int32_t* __restrict__ pid, long long T,
                                                      unsigned long long* __restrict__ cursor,
                                                      int32_t* __restrict__ perm) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < T; i += (long long)gridDim.x * blockDim.x)
    perm[atomicAdd(&cursor[pid[i]], 1ULL)] = (int32_t)i;
}

__global__ void __launch_bounds__(256) k_copy_i64(const int64_t* __restrict__ a, int64_t* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    b[i] = a[i];
}

// longest bucket (skew check)
__global__ void __launch_bounds__(256) k_bucket_max(const int64_t* __restrict__ off, int B, int* __restrict__ out) {
  int m = 0;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    const long long n = off[b + 1] - off[b];
    m = max(m, (int)(n > 0x7fffffffLL ? 0x7fffffffLL : n));
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
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

int mgb_bucket_aggregate(const Layout& L, const BktSource* srcs_host, int nsrc, long long T, cudaStream_t st,
                         BktResult* out, int* overflow) {
  *overflow = 0;
  memset(out, 0, sizeof(*out));
  MGB_REQUIRE(nsrc >= 1 && nsrc < 65536 && T >= 1 && T < 0x7fffffffLL, MGB_ERR_UNSUPPORTED,
              "bucketed aggregation: %d sources, %lld rows", nsrc, T);
  int logB = 0;
  while (logB < 22 && (T >> logB) > BK_TARGET) ++logB;
  const int B = 1 << logB;
  const int ncol = L.nk + L.na, W = L.nk + L.nv;
  char* small = nullptr;     // srcs | begin | err, longest bucket
  const size_t src_bytes = sizeof(BktSource) * (size_t)nsrc;
  const size_t begin_bytes = sizeof(long long) * ((size_t)nsrc + 1);
  const size_t src_pad = (src_bytes + 255) & ~(size_t)255, begin_pad = (begin_bytes + 255) & ~(size_t)255;
  int32_t *pid = nullptr, *perm = nullptr;
  uint64_t* rows = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&small, src_pad + begin_pad + 256, st));
  MGB_CUDA(cudaMallocAsync((void**)&pid, sizeof(int32_t) * (size_t)T, st));
  MGB_CUDA(cudaMallocAsync((void**)&perm, sizeof(int32_t) * (size_t)T, st));
  MGB_CUDA(cudaMallocAsync((void**)&rows, sizeof(uint64_t) * (size_t)W * (size_t)T, st));
  MGB_CUDA(cudaMallocAsync((void**)&out->off, sizeof(int64_t) * ((size_t)B + 1), st));
  MGB_CUDA(cudaMallocAsync((void**)&out->goff, sizeof(int64_t) * ((size_t)B + 1), st));
  MGB_CUDA(cudaMallocAsync((void**)&out->scratch, sizeof(uint64_t) * (size_t)ncol * (size_t)T, st));
  out->T = T;
  out->B = B;
  BktSource* d_srcs = (BktSource*)small;
  long long* d_begin = (long long*)(small + src_pad);
  int* d_err = (int*)(small + src_pad + begin_pad);
  {
    std::vector<long long> begin((size_t)nsrc + 1);
    for (int s = 0; s < nsrc; ++s) begin[s] = srcs_host[s].begin;
    begin[nsrc] = T;
    // pageable sources: the runtime stages these small copies before cudaMemcpyAsync returns
    MGB_CUDA(cudaMemcpyAsync(d_srcs, srcs_host, src_bytes, cudaMemcpyHostToDevice, st));
    MGB_CUDA(cudaMemcpyAsync(d_begin, begin.data(), begin_bytes, cudaMemcpyHostToDevice, st));
  }
  MGB_CUDA(cudaMemsetAsync(d_err, 0, 256, st));
  MGB_CUDA(cudaMemsetAsync(out->goff + B, 0, sizeof(int64_t), st));
  auto release = [&]() {
    cudaFreeAsync(pid, st);
    cudaFreeAsync(perm, st);
    cudaFreeAsync(rows, st);
    cudaFreeAsync(small, st);
  };
  int status = MGB_OK;
  {
    MgbProfScope prof(MGB_K_BUCKET, st);
    // 1. bucket ids + row-major copy of the rows
    BktPackParams pp;
    pp.srcs = d_srcs;
    pp.begin = d_begin;
    pp.src0 = srcs_host[0];
    pp.nsrc = nsrc;
    pp.nk = L.nk;
    pp.nv = L.nv;
    pp.T = T;
    pp.logB = logB;
    pp.pid = pid;
    pp.rows = rows;
    const size_t sm = (size_t)BKP_ROWS * (size_t)(W | 1) * 8;
    const long long ntiles = (T + BKP_ROWS - 1) / BKP_ROWS, capb = (long long)mgb_sm_count() * 16;
    const unsigned grid = (unsigned)(ntiles > capb ? capb : ntiles);
    if (L.nk == 1) {
      if (sm > 48 * 1024) MGB_CUDA(cudaFuncSetAttribute(k_bucket_pack<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_bucket_pack<1><<<grid, BKP_ROWS, sm, st>>>(pp);
    } else {
      if (sm > 48 * 1024) MGB_CUDA(cudaFuncSetAttribute(k_bucket_pack<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_bucket_pack<2><<<grid, BKP_ROWS, sm, st>>>(pp);
    }
    MGB_CHECK_LAUNCH();
    bk_debug("k_bucket_pack", st);
    // 2. row ids in bucket order: histogram -> offsets -> placement through one atomic cursor per bucket
    {
      const long long blocks = (T + 255) / 256, capg = (long long)mgb_sm_count() * 16;
      const unsigned g2 = (unsigned)(blocks > capg ? capg : blocks);
      const unsigned gB = (unsigned)(((long long)B + 256) / 256 > capg ? capg : ((long long)B + 256) / 256);
      MGB_CUDA(cudaMemsetAsync(out->off, 0, sizeof(int64_t) * ((size_t)B + 1), st));
      k_bucket_hist<<<g2, 256, 0, st>>>(pid, T, (unsigned long long*)out->off);
      MGB_CHECK_LAUNCH();
      status = mgb_scan_exclusive_i64(out->off, (int64_t)B + 1, nullptr, st);
      if (status == MGB_OK) {
        // goff doubles as the cursor array until the aggregation kernel overwrites it with the group counts
        k_copy_i64<<<gB, 256, 0, st>>>(out->off, out->goff, (long long)B + 1);
        k_bucket_place<<<g2, 256, 0, st>>>(pid, T, (unsigned long long*)out->goff, perm);
        MGB_CHECK_LAUNCH();
        MGB_CUDA(cudaMemsetAsync(out->goff + B, 0, sizeof(int64_t), st));
      }
    }
    bk_debug("row ids in bucket order", st);
    if (status == MGB_OK) {   // a hot key makes one bucket as long as its row count: leave that input to the table
      const int gb = (B + 255) / 256;
      k_bucket_max<<<gb > 1024 ? 1024 : gb, 256, 0, st>>>(out->off, B, d_err + 1);
      int longest = 0;
      cudaError_t e = cudaMemcpyAsync(&longest, d_err + 1, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        mgb_set_error("bucket aggregation: %s", cudaGetErrorString(e));
        status = MGB_ERR_CUDA;
      } else if (longest > BK_SKEW_ROWS) {
        release();
        mgb_bucket_free(out, st);
        *overflow = 2;
        return MGB_OK;
      }
    }
    // 3. one CTA per bucket
    if (status == MGB_OK) {
      BktParams p;
      p.rows = rows;
      p.W = W;
      p.T = T;
      p.logB = logB;
      p.perm = perm;
      p.off = out->off;
      p.scratch = out->scratch;
      p.gcount = out->goff;
      p.err = d_err;
      p.L = L;
      if (L.nk == 1)
        k_bucket_agg<1><<<(unsigned)B, BK_THREADS, 0, st>>>(p);
      else
        k_bucket_agg<2><<<(unsigned)B, BK_THREADS, 0, st>>>(p);
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
    *overflow = err_h ? 1 : 0;
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
