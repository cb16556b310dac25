// libmarsgb: low-cardinality fast paths (sm_100a).
//
//  * mgb_preagg_dense  -- a4, DataFrameGroupByAgg._execute_map (mars/dataframe/groupby/aggregation.py:1043-1128)
//    for chunks whose number of distinct key tuples fits on chip (TPC-H Q1 shape: 4 groups).  ONE launch:
//      - every CTA streams its row tiles with register double-buffering (R rows per thread in flight while
//        the previous R are accumulated), finds the group id of each row in a 64-entry direct-mapped
//        shared-memory key table and updates accumulators that are privatised per warp AND per lane
//        ([warp][group][slot][lane]) -- plain LDS / DADD / STS, no atomics, no bank conflicts;
//      - count(x) is kept as rows-per-group minus NaNs-per-column, so a count costs one 32-bit add per row;
//      - at the end each CTA reduces lanes and warps and writes <= GC partial rows to a global staging
//        area; the last CTA to finish (atomic ticket, no waiting) merges all staged rows into unique
//        output rows and publishes the group count.
//    A CTA that meets more than GC distinct keys raises the fail flag: the call reports -1 groups and the
//    caller takes the general hash-table path (mgb_agg_*).
//
//  * mgb_merge_small   -- a9/a10, _execute_combine/_execute_agg (aggregation.py:1130-1267) and the concat
//    feeding them (merge/concat.py:339-348) for small partial tuples: a single CTA merges up to
//    MGB_SMALL_MAX_ROWS partial rows taken from several pieces (no concat copy), optionally rank-sorts the
//    unique keys (groupby(sort=True)) and writes the result.
#include "mgb_common.cuh"

#define DN_THREADS 256
#define DN_WARPS (DN_THREADS / 32)
#define DN_R 4                 // rows per thread per tile
#define DN_MAX_GC 16
#define DN_KEYTAB 64           // direct-mapped key table entries (power of two)
#define DN_SMEM_BUDGET (208 * 1024)
#define MGB_SMALL_MAX_ROWS 2048
#define SM_TAB 4096            // hash slots of the small merge (>= 2 * MGB_SMALL_MAX_ROWS)
#define SM_THREADS 512
#define SM_MAX_PIECES 16

#define OPB(op) (1u << (op))

// --------------------------------------------------------------------------------------------------
// shared "merge partial rows into unique output rows" machinery (one CTA)
// --------------------------------------------------------------------------------------------------
struct OutCols {
  uint64_t* key[MGB_MAX_KEYS];
  uint64_t* acc[MGB_MAX_ACCS];
  uint8_t op[MGB_MAX_ACCS];   // merge op of accumulator a: MGB_SUM (adds; also counts), MGB_MIN, MGB_MAX
  uint8_t dt[MGB_MAX_ACCS];   // MGB_F64 / MGB_I64 of the emitted column
  int na;
};

template <int NK>
__device__ __forceinline__ uint32_t small_hash(const uint64_t* k) {
  uint64_t h = mgb_mix64(k[0]);
  if (NK == 2) h = mgb_mix64(h ^ (k[1] * 0x9e3779b97f4a7c15ULL));
  return (uint32_t)(h >> 40);
}

// Insert-or-find a key tuple in the CTA's table.  tab_idx[slot] = -1 empty, -2 being written, >= 0 dense row.
// Keys are published before the index (threadfence_block) so readers that see idx >= 0 see the keys.
template <int NK>
__device__ __forceinline__ int small_find_or_insert(uint64_t* tab_keys, int* tab_idx, int* n_unique,
                                                    const uint64_t* k) {
  uint32_t s = small_hash<NK>(k) & (SM_TAB - 1);
  for (int probe = 0; probe < SM_TAB; ++probe, s = (s + 1) & (SM_TAB - 1)) {
    int cur = *(volatile int*)&tab_idx[s];
    if (cur == -1) {
      cur = atomicCAS(&tab_idx[s], -1, -2);
      if (cur == -1) {
        tab_keys[s] = k[0];
        if (NK == 2) tab_keys[SM_TAB + s] = k[1];
        const int idx = atomicAdd(n_unique, 1);
        __threadfence_block();
        *(volatile int*)&tab_idx[s] = idx;
        return idx;
      }
    }
    while (cur == -2) cur = *(volatile int*)&tab_idx[s];   // writer is a few instructions from publishing
    bool eq = ((volatile uint64_t*)tab_keys)[s] == k[0];
    if (NK == 2) eq = eq && ((volatile uint64_t*)tab_keys)[SM_TAB + s] == k[1];
    if (eq) return cur;
  }
  return -1;
}

__device__ __forceinline__ void out_init(const OutCols& o, int rows) {
  for (int i = threadIdx.x; i < rows * o.na; i += blockDim.x) {
    const int a = i / rows, r = i - a * rows;
    o.acc[a][r] = mgb_acc_init(o.op[a]);
  }
}

// accumulate decoded partial value v into output row idx (global atomics: the rows are few)
__device__ __forceinline__ void out_accumulate(const OutCols& o, int a, int idx, uint64_t v) {
  const int op = o.op[a], dt = o.dt[a];
  uint64_t* p = o.acc[a] + idx;
  if (op == MGB_MIN || op == MGB_MAX) {
    uint64_t e;
    if (dt == MGB_F64) {
      if (mgb_is_nan_bits(v)) return;   // NaN partial == "no value"
      e = mgb_enc_f64(v);
    } else {
      e = mgb_enc_i64(v);
    }
    if (op == MGB_MIN)
      atomicMin((unsigned long long*)p, (unsigned long long)e);
    else
      atomicMax((unsigned long long*)p, (unsigned long long)e);
  } else if (dt == MGB_F64) {
    atomicAdd((double*)p, __longlong_as_double((long long)v));
  } else {
    atomicAdd((unsigned long long*)p, (unsigned long long)v);
  }
}

__device__ __forceinline__ uint64_t ld_cg(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// decode min/max columns in place (rows [0, n)); call after a __syncthreads + __threadfence
__device__ __forceinline__ void out_decode(const OutCols& o, int n) {
  for (int a = 0; a < o.na; ++a) {
    if (o.op[a] != MGB_MIN && o.op[a] != MGB_MAX) continue;
    for (int r = threadIdx.x; r < n; r += blockDim.x)
      o.acc[a][r] = mgb_acc_decode(o.op[a], o.dt[a], ld_cg(o.acc[a] + r));
  }
}

template <int NK>
__device__ __forceinline__ bool key_less(const uint64_t* a, const uint64_t* b) {
  const int64_t a0 = (int64_t)a[0], b0 = (int64_t)b[0];
  if (NK == 1) return a0 < b0;
  if (a0 != b0) return a0 < b0;
  return (int64_t)a[1] < (int64_t)b[1];
}

// --------------------------------------------------------------------------------------------------
// dense pre-aggregation
// --------------------------------------------------------------------------------------------------
struct DenseParams {
  const uint64_t* keys[MGB_MAX_KEYS];
  const uint64_t* cols[8];
  long long n;
  int ncol;                 // loaded value columns (<= NV)
  int GC;                   // groups per CTA
  int stride_g;             // 32-bit words per (warp, group) accumulator block
  int off_nan, off_sum, off_sumsq, off_min, off_max;   // word offsets of column 0's slot inside a block
  uint8_t col_dt[8];
  uint8_t col_mask[8];      // OPB(MGB_SUM) | ...
  // global staging: row = cta * GC + g; words: keys[NK] | rows | per column: nan, sum, sumsq, min, max
  uint64_t* stg;
  int stg_w;
  int* cta_ng;
  unsigned int* ticket;
  int* fail;
  // output (spec order)
  OutCols out;
  uint8_t out_col[MGB_MAX_ACCS];   // loaded column of spec accumulator a, 0xFF: "rows of the group"
  uint8_t out_src[MGB_MAX_ACCS];   // mgb_op the staged value comes from
  long long* out_n;                // device: number of groups, -1 on failure
};

__device__ __forceinline__ uint32_t dense_slot(const uint64_t* k, int nk) {
  uint32_t h = (uint32_t)k[0] * 0x9E3779B1u + (uint32_t)(k[0] >> 32) * 0x85EBCA77u;
  if (nk == 2) h += (uint32_t)k[1] * 0xC2B2AE3Du + (uint32_t)(k[1] >> 32) * 0x27D4EB2Fu;
  return (h ^ (h >> 15)) * 0x2C1B3C6Du >> 26;
}

template <int NK, int NV, bool ALLOPS>
__global__ void __launch_bounds__(DN_THREADS, 1) k_preagg_dense(const DenseParams p) {
  extern __shared__ __align__(16) uint32_t dsm[];
  // layout: acc[warp][GC][stride_g] | tkeys[NK][KEYTAB] (u64) | tgid[KEYTAB] | gkeys[NK][MAX_GC] (u64) | ints
  uint32_t* acc = dsm;
  uint64_t* tkeys = (uint64_t*)(dsm + (size_t)DN_WARPS * p.GC * p.stride_g);
  int* tgid = (int*)(tkeys + NK * DN_KEYTAB);
  uint64_t* gkeys = (uint64_t*)(tgid + DN_KEYTAB);
  int* s_int = (int*)(gkeys + MGB_MAX_KEYS * DN_MAX_GC);
  volatile int* s_ng = s_int;
  int* s_lock = s_int + 1;
  int* s_fail = s_int + 2;
  int* s_last = s_int + 3;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int GC = p.GC;
  uint32_t* wacc = acc + (size_t)warp * GC * p.stride_g;
  // ---- init accumulators: zeros, except min (+inf / INT64_MAX) and max (-inf / INT64_MIN)
  for (int i = lane; i < GC * p.stride_g; i += 32) wacc[i] = 0;
  if (ALLOPS) {
    __syncwarp();
    for (int g = 0; g < GC; ++g)
      for (int c = 0; c < p.ncol; ++c) {
        uint64_t* b = (uint64_t*)(wacc + (size_t)g * p.stride_g);
        const bool f = p.col_dt[c] == MGB_F64;
        b[(p.off_min >> 1) + c * 32 + lane] = f ? 0x7ff0000000000000ULL : 0x7fffffffffffffffULL;
        b[(p.off_max >> 1) + c * 32 + lane] = f ? 0xfff0000000000000ULL : 0x8000000000000000ULL;
      }
  }
  for (int i = threadIdx.x; i < DN_KEYTAB; i += DN_THREADS) tgid[i] = -1;
  if (threadIdx.x == 0) {
    *s_ng = 0;
    *s_lock = 0;
    *s_fail = 0;
    *s_last = 0;
  }
  __syncthreads();

  const long long n = p.n;
  const long long tile_rows = (long long)DN_THREADS * DN_R;
  const long long ntiles = (n + tile_rows - 1) / tile_rows;

  uint64_t kb[2][DN_R][NK];
  uint64_t xb[2][DN_R][NV];

  auto load_tile = [&](int buf, long long t) {
    const long long base = t * tile_rows + threadIdx.x;
#pragma unroll
    for (int r = 0; r < DN_R; ++r) {
      const long long i = base + (long long)r * DN_THREADS;
      const bool v = i < n;
#pragma unroll
      for (int j = 0; j < NK; ++j) kb[buf][r][j] = v ? mgb_ld_stream(p.keys[j] + i) : 0;
#pragma unroll
      for (int c = 0; c < NV; ++c) xb[buf][r][c] = (v && c < p.ncol) ? mgb_ld_stream(p.cols[c] + i) : 0;
    }
  };

  bool failed = false;
  auto process_tile = [&](int buf, long long t) {
    const long long base = t * tile_rows + threadIdx.x;
#pragma unroll
    for (int r = 0; r < DN_R; ++r) {
      const bool valid = base + (long long)r * DN_THREADS < n;
      const uint64_t* k = kb[buf][r];
      int g = -1;
      if (valid) {
        const uint32_t s = dense_slot(k, NK);
        const int tg = *(volatile int*)&tgid[s];
        bool eq = tg >= 0 && ((volatile uint64_t*)tkeys)[s] == k[0];
        if (NK == 2) eq = eq && ((volatile uint64_t*)tkeys)[DN_KEYTAB + s] == k[1];
        if (eq) g = tg;
      }
      unsigned need = __ballot_sync(0xffffffffu, valid && g < 0);
      while (need) {   // slow path: new key, or direct-mapped conflict (then found in the group list)
        const int src = __ffs(need) - 1;
        uint64_t ks[NK];
#pragma unroll
        for (int j = 0; j < NK; ++j) ks[j] = __shfl_sync(0xffffffffu, k[j], src);
        int gi = 0;
        if (lane == 0) {
          while (atomicCAS(s_lock, 0, 1) != 0) {
          }
          __threadfence_block();
          const int ng = *s_ng;
          gi = -1;
          for (int q = 0; q < ng; ++q) {
            bool eq = ((volatile uint64_t*)gkeys)[q] == ks[0];
            if (NK == 2) eq = eq && ((volatile uint64_t*)gkeys)[DN_MAX_GC + q] == ks[1];
            if (eq) {
              gi = q;
              break;
            }
          }
          if (gi < 0) {
            if (ng < GC) {
              ((volatile uint64_t*)gkeys)[ng] = ks[0];
              if (NK == 2) ((volatile uint64_t*)gkeys)[DN_MAX_GC + ng] = ks[1];
              gi = ng;
              const uint32_t s = dense_slot(ks, NK);
              if (*(volatile int*)&tgid[s] < 0) {   // claim the direct-mapped entry if it is free
                ((volatile uint64_t*)tkeys)[s] = ks[0];
                if (NK == 2) ((volatile uint64_t*)tkeys)[DN_KEYTAB + s] = ks[1];
                __threadfence_block();
                *(volatile int*)&tgid[s] = gi;
              }
              __threadfence_block();
              *s_ng = ng + 1;
            } else {
              gi = -2;
            }
          }
          __threadfence_block();
          atomicExch(s_lock, 0);
        }
        gi = __shfl_sync(0xffffffffu, gi, 0);
        if (gi < 0) {
          failed = true;
          break;
        }
        bool mine = valid && g < 0;
#pragma unroll
        for (int j = 0; j < NK; ++j) mine = mine && (k[j] == ks[j]);
        if (mine) g = gi;
        need = __ballot_sync(0xffffffffu, valid && g < 0);
      }
      if (failed) return;
      if (valid) {
        uint32_t* blk = wacc + (size_t)g * p.stride_g + lane;
        blk[0] += 1;   // rows of the group
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          if (c < p.ncol) {
            const uint64_t x = xb[buf][r][c];
            const uint32_t m = p.col_mask[c];
            if (p.col_dt[c] == MGB_F64) {
              if (mgb_is_nan_bits(x)) {
                blk[p.off_nan + c * 32] += 1;
              } else {
                const double xv = __longlong_as_double((long long)x);
                double* d = (double*)(blk - lane) + lane;   // 64-bit view of the block, lane-private column
                if (m & OPB(MGB_SUM)) d[(p.off_sum >> 1) + c * 32] += xv;
                if (ALLOPS) {
                  if (m & OPB(MGB_SUMSQ)) d[(p.off_sumsq >> 1) + c * 32] += xv * xv;
                  if (m & OPB(MGB_MIN)) d[(p.off_min >> 1) + c * 32] = fmin(d[(p.off_min >> 1) + c * 32], xv);
                  if (m & OPB(MGB_MAX)) d[(p.off_max >> 1) + c * 32] = fmax(d[(p.off_max >> 1) + c * 32], xv);
                }
              }
            } else {
              long long* d = (long long*)(blk - lane) + lane;
              const long long xv = (long long)x;
              if (m & OPB(MGB_SUM)) d[(p.off_sum >> 1) + c * 32] += xv;
              if (ALLOPS) {
                if (m & OPB(MGB_SUMSQ)) d[(p.off_sumsq >> 1) + c * 32] += xv * xv;
                if (m & OPB(MGB_MIN)) d[(p.off_min >> 1) + c * 32] = min(d[(p.off_min >> 1) + c * 32], xv);
                if (m & OPB(MGB_MAX)) d[(p.off_max >> 1) + c * 32] = max(d[(p.off_max >> 1) + c * 32], xv);
              }
            }
          }
        }
      }
    }
  };

  // ---- main loop: register double-buffering, tiles blockIdx.x, +gridDim.x, ...
  long long t = blockIdx.x;
  if (t < ntiles) load_tile(0, t);
  int iter = 0;
  while (t < ntiles && !failed) {
    const long long t1 = t + gridDim.x;
    if (t1 < ntiles) load_tile(1, t1);
    process_tile(0, t);
    if (failed || t1 >= ntiles) break;
    const long long t2 = t1 + gridDim.x;
    if (t2 < ntiles) load_tile(0, t2);
    process_tile(1, t1);
    t = t2;
    if ((++iter & 7) == 0 && *(volatile int*)p.fail) failed = true;   // somebody else gave up: stop early
  }
  if (failed) {
    *s_fail = 1;
    *p.fail = 1;
  }
  __syncthreads();
  const bool cta_failed = *s_fail != 0;

  // ---- CTA reduction -> staging rows
  const int ng = cta_failed ? 0 : *s_ng;
  const int W = p.stg_w;
  for (int q = warp; q < ng * (1 + p.ncol * 5); q += DN_WARPS) {
    const int g = q / (1 + p.ncol * 5), s = q - g * (1 + p.ncol * 5);
    uint64_t* row = p.stg + ((size_t)blockIdx.x * GC + g) * W;
    if (s == 0) {   // rows
      uint32_t v = 0;
      for (int w = 0; w < DN_WARPS; ++w) v += acc[((size_t)w * GC + g) * p.stride_g + lane];
      v = __reduce_add_sync(0xffffffffu, v);
      if (lane == 0) row[NK] = v;
      continue;
    }
    const int c = (s - 1) / 5, kind = (s - 1) % 5;   // kind: 0 nan, 1 sum, 2 sumsq, 3 min, 4 max
    uint64_t* dst = row + NK + 1 + c * 5 + kind;
    if (kind == 0) {
      uint32_t v = 0;
      for (int w = 0; w < DN_WARPS; ++w) v += acc[((size_t)w * GC + g) * p.stride_g + p.off_nan + c * 32 + lane];
      v = __reduce_add_sync(0xffffffffu, v);
      if (lane == 0) *dst = v;
      continue;
    }
    if (!ALLOPS && kind > 1) continue;
    const uint32_t m = p.col_mask[c];
    const int op = kind == 1 ? MGB_SUM : kind == 2 ? MGB_SUMSQ : kind == 3 ? MGB_MIN : MGB_MAX;
    if (!(m & OPB(op))) continue;
    const int off = kind == 1 ? p.off_sum : kind == 2 ? p.off_sumsq : kind == 3 ? p.off_min : p.off_max;
    const bool f = p.col_dt[c] == MGB_F64;
    uint64_t v = 0;
    for (int w = 0; w < DN_WARPS; ++w) {
      const uint64_t y = ((const uint64_t*)(acc + ((size_t)w * GC + g) * p.stride_g))[(off >> 1) + c * 32 + lane];
      if (w == 0) {
        v = y;
      } else if (f) {
        const double a = __longlong_as_double((long long)v), b = __longlong_as_double((long long)y);
        const double r = kind <= 2 ? a + b : kind == 3 ? fmin(a, b) : fmax(a, b);
        v = (uint64_t)__double_as_longlong(r);
      } else {
        const long long a = (long long)v, b = (long long)y;
        v = (uint64_t)(kind <= 2 ? a + b : kind == 3 ? min(a, b) : max(a, b));
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t y = __shfl_xor_sync(0xffffffffu, v, o);
      if (f) {
        const double a = __longlong_as_double((long long)v), b = __longlong_as_double((long long)y);
        const double r = kind <= 2 ? a + b : kind == 3 ? fmin(a, b) : fmax(a, b);
        v = (uint64_t)__double_as_longlong(r);
      } else {
        const long long a = (long long)v, b = (long long)y;
        v = (uint64_t)(kind <= 2 ? a + b : kind == 3 ? min(a, b) : max(a, b));
      }
    }
    if (lane == 0) *dst = v;
  }
  for (int q = threadIdx.x; q < ng * NK; q += DN_THREADS) {
    const int g = q / NK, j = q - g * NK;
    p.stg[((size_t)blockIdx.x * GC + g) * W + j] = gkeys[j * DN_MAX_GC + g];
  }
  if (threadIdx.x == 0) p.cta_ng[blockIdx.x] = ng;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int tk = atomicAdd(p.ticket, 1u);
    *s_last = (tk == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!*s_last) return;

  // ---- last CTA: merge all staged rows into unique output rows (reuses the accumulator shared memory)
  __threadfence();
  if (*(volatile int*)p.fail) {
    if (threadIdx.x == 0) *p.out_n = -1;
    return;
  }
  uint64_t* tab_keys = (uint64_t*)dsm;                       // [NK][SM_TAB]
  int* tab_idx = (int*)(tab_keys + MGB_MAX_KEYS * SM_TAB);   // [SM_TAB]
  int* n_unique = tab_idx + SM_TAB;
  for (int i = threadIdx.x; i < SM_TAB; i += DN_THREADS) tab_idx[i] = -1;
  if (threadIdx.x == 0) *n_unique = 0;
  const int total = gridDim.x * GC;
  out_init(p.out, total);
  __threadfence();
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += DN_THREADS) {
    const int cta = i / GC, g = i - cta * GC;
    if (g >= ((volatile int*)p.cta_ng)[cta]) continue;
    const uint64_t* row = p.stg + (size_t)i * W;
    uint64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = ld_cg(row + j);
    const int idx = small_find_or_insert<NK>(tab_keys, tab_idx, n_unique, k);
    if (idx < 0) continue;   // cannot happen: table has 4096 slots for <= 2368 rows
#pragma unroll
    for (int j = 0; j < NK; ++j) p.out.key[j][idx] = k[j];   // same value from every writer
    const uint64_t rows = ld_cg(row + NK);
    for (int a = 0; a < p.out.na; ++a) {
      const int c = p.out_col[a], src = p.out_src[a];
      uint64_t v;
      if (src == MGB_COUNT) {
        v = c == 0xFF ? rows : rows - ld_cg(row + NK + 1 + c * 5);
      } else {
        const uint64_t cnt = p.col_dt[c] == MGB_F64 ? rows - ld_cg(row + NK + 1 + c * 5) : rows;
        const int kind = src == MGB_SUM ? 1 : src == MGB_SUMSQ ? 2 : src == MGB_MIN ? 3 : 4;
        v = ld_cg(row + NK + 1 + c * 5 + kind);
        if ((src == MGB_MIN || src == MGB_MAX) && cnt == 0) continue;   // no value in this partial
      }
      out_accumulate(p.out, a, idx, v);
    }
  }
  __threadfence();
  __syncthreads();
  const int nu = *n_unique;
  out_decode(p.out, nu);
  if (threadIdx.x == 0) *p.out_n = nu;
}

// --------------------------------------------------------------------------------------------------
// host side of the dense path
// --------------------------------------------------------------------------------------------------
struct DenseScratch {   // per (device, stream) scratch, grown on demand, reused across calls
  uint64_t* stg = nullptr;
  size_t stg_words = 0;
  int* ints = nullptr;          // cta_ng[grid] | ticket | fail
  long long* out_n = nullptr;   // device result
  long long* out_n_host = nullptr;   // pinned
};
static thread_local DenseScratch g_ds[8];

static int dense_scratch(int dev, size_t stg_words, int grid, DenseScratch** out) {
  MGB_REQUIRE(dev >= 0 && dev < 8, MGB_ERR_UNSUPPORTED, "device ordinal %d", dev);
  DenseScratch& s = g_ds[dev];
  if (s.stg_words < stg_words) {
    if (s.stg) cudaFree(s.stg);
    MGB_CUDA(cudaMalloc((void**)&s.stg, stg_words * 8));
    s.stg_words = stg_words;
  }
  if (!s.ints) {
    MGB_CUDA(cudaMalloc((void**)&s.ints, sizeof(int) * (1024 + 8)));
    MGB_CUDA(cudaMalloc((void**)&s.out_n, sizeof(long long) * 2));
    MGB_CUDA(cudaMallocHost((void**)&s.out_n_host, sizeof(long long) * 2));
  }
  MGB_REQUIRE(grid <= 1024, MGB_ERR_UNSUPPORTED, "grid %d", grid);
  *out = &s;
  return MGB_OK;
}

extern "C" int mgb_dense_capacity(int64_t* out_rows) {
  MGB_REQUIRE(out_rows, MGB_ERR_INVALID, "null pointer");
  *out_rows = (int64_t)mgb_sm_count() * DN_MAX_GC;
  return MGB_OK;
}

extern "C" int mgb_preagg_dense(const mgb_agg_spec* spec, const int64_t* const* key_cols,
                                const void* const* val_cols, int64_t n_rows, int64_t* const* out_key_cols,
                                void* const* out_acc_cols, int64_t out_capacity, int64_t* out_n_groups,
                                void* stream) {
  MGB_REQUIRE(spec && key_cols && val_cols && out_key_cols && out_acc_cols && out_n_groups, MGB_ERR_INVALID,
              "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  *out_n_groups = -1;
  const int nk = spec->n_keys, na = spec->n_accs;
  MGB_REQUIRE(nk >= 1 && nk <= MGB_MAX_KEYS, MGB_ERR_UNSUPPORTED, "n_keys=%d", nk);
  MGB_REQUIRE(na >= 1 && na <= MGB_MAX_ACCS && spec->n_vals >= 1 && spec->n_vals <= MGB_MAX_VALS,
              MGB_ERR_UNSUPPORTED, "spec outside the supported envelope");
  MGB_REQUIRE(n_rows >= 0, MGB_ERR_INVALID, "negative n_rows");
  if (n_rows == 0) {
    *out_n_groups = 0;
    return MGB_OK;
  }
  DenseParams p;
  memset(&p, 0, sizeof(p));
  // columns that must be read: everything except int64 columns that are only counted
  int loaded_of[MGB_MAX_VALS];
  int ncol = 0;
  bool allops = false;
  for (int c = 0; c < spec->n_vals; ++c) {
    uint32_t m = 0;
    for (int a = 0; a < na; ++a)
      if (spec->acc_col[a] == c) m |= OPB(spec->acc_op[a]);
    loaded_of[c] = -1;
    if (m == 0) continue;
    if (spec->val_dtype[c] == MGB_I64 && m == OPB(MGB_COUNT)) continue;
    if (ncol == 8) return MGB_OK;   // more than 8 live columns: not a dense-path shape (caller falls back)
    MGB_REQUIRE(val_cols[c], MGB_ERR_INVALID, "null value column %d", c);
    loaded_of[c] = ncol;
    p.cols[ncol] = (const uint64_t*)val_cols[c];
    p.col_dt[ncol] = (uint8_t)spec->val_dtype[c];
    p.col_mask[ncol] = (uint8_t)m;
    if (m & (OPB(MGB_SUMSQ) | OPB(MGB_MIN) | OPB(MGB_MAX))) allops = true;
    ++ncol;
  }
  for (int j = 0; j < nk; ++j) {
    MGB_REQUIRE(key_cols[j] && out_key_cols[j], MGB_ERR_INVALID, "null key column %d", j);
    p.keys[j] = (const uint64_t*)key_cols[j];
    p.out.key[j] = (uint64_t*)out_key_cols[j];
  }
  p.n = n_rows;
  p.ncol = ncol;
  // accumulator block layout (32-bit words): rows[32] | nan[ncol][32] | sum[ncol][64] | (sumsq|min|max)[ncol][64]
  p.off_nan = 32;
  p.off_sum = 32 + 32 * ncol;
  p.off_sumsq = p.off_sum + 64 * ncol;
  p.off_min = p.off_sumsq + 64 * ncol;
  p.off_max = p.off_min + 64 * ncol;
  p.stride_g = allops ? p.off_max + 64 * ncol : p.off_sumsq;
  const size_t fixed = (size_t)(MGB_MAX_KEYS * DN_KEYTAB + MGB_MAX_KEYS * DN_MAX_GC) * 8 + DN_KEYTAB * 4 + 64;
  int GC = (int)((DN_SMEM_BUDGET - fixed) / ((size_t)DN_WARPS * p.stride_g * 4));
  if (GC > DN_MAX_GC) GC = DN_MAX_GC;
  if (GC < 2) return MGB_OK;   // accumulators of a single group do not fit: general path
  p.GC = GC;
  const long long tile_rows = (long long)DN_THREADS * DN_R;
  const long long ntiles = (n_rows + tile_rows - 1) / tile_rows;
  const int sms = mgb_sm_count();
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  MGB_REQUIRE(out_capacity >= (int64_t)grid * GC, MGB_ERR_INVALID,
              "output capacity %lld < %d (mgb_dense_capacity)", (long long)out_capacity, grid * GC);
  p.stg_w = nk + 1 + 5 * ncol;
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  DenseScratch* s = nullptr;
  MGB_TRY(dense_scratch(dev, (size_t)grid * GC * p.stg_w, grid, &s));
  p.stg = s->stg;
  p.cta_ng = s->ints;
  p.ticket = (unsigned int*)(s->ints + 1024);
  p.fail = s->ints + 1025;
  p.out_n = s->out_n;
  MGB_CUDA(cudaMemsetAsync(s->ints + 1024, 0, 2 * sizeof(int), st));
  p.out.na = na;
  for (int a = 0; a < na; ++a) {
    MGB_REQUIRE(out_acc_cols[a], MGB_ERR_INVALID, "null output accumulator column %d", a);
    const int c = spec->acc_col[a], op = spec->acc_op[a];
    MGB_REQUIRE(c >= 0 && c < spec->n_vals && op >= MGB_SUM && op <= MGB_MAX, MGB_ERR_INVALID,
                "bad accumulator %d", a);
    p.out.acc[a] = (uint64_t*)out_acc_cols[a];
    p.out.op[a] = (uint8_t)(op == MGB_MIN ? MGB_MIN : op == MGB_MAX ? MGB_MAX : MGB_SUM);
    p.out.dt[a] = (uint8_t)(op == MGB_COUNT ? MGB_I64 : spec->val_dtype[c]);
    p.out_col[a] = (uint8_t)(loaded_of[c] < 0 ? 0xFF : loaded_of[c]);
    p.out_src[a] = (uint8_t)op;
  }
  size_t smem = (size_t)DN_WARPS * GC * p.stride_g * 4 + fixed;
  const size_t merge_smem = (size_t)MGB_MAX_KEYS * SM_TAB * 8 + SM_TAB * 4 + 64;
  if (smem < merge_smem) smem = merge_smem;
  {
    MgbProfScope prof(MGB_K_PREAGG_TINY, st);
#define LAUNCH(NK, NV, AO)                                                                              \
  do {                                                                                                  \
    MGB_CUDA(cudaFuncSetAttribute(k_preagg_dense<NK, NV, AO>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem));                                                          \
    k_preagg_dense<NK, NV, AO><<<grid, DN_THREADS, smem, st>>>(p);                                      \
  } while (0)
#define PICK_NV(NK, AO)                          \
  do {                                           \
    if (ncol <= 1) LAUNCH(NK, 1, AO);            \
    else if (ncol <= 2) LAUNCH(NK, 2, AO);       \
    else if (ncol <= 4) LAUNCH(NK, 4, AO);       \
    else if (ncol <= 6) LAUNCH(NK, 6, AO);       \
    else LAUNCH(NK, 8, AO);                      \
  } while (0)
    if (nk == 1) {
      if (allops) PICK_NV(1, true); else PICK_NV(1, false);
    } else {
      if (allops) PICK_NV(2, true); else PICK_NV(2, false);
    }
#undef PICK_NV
#undef LAUNCH
    MGB_CHECK_LAUNCH();
  }
  MGB_CUDA(cudaMemcpyAsync(s->out_n_host, s->out_n, sizeof(long long), cudaMemcpyDeviceToHost, st));
  MGB_CUDA(cudaStreamSynchronize(st));
  *out_n_groups = s->out_n_host[0];
  return MGB_OK;
}

// --------------------------------------------------------------------------------------------------
// small merge
// --------------------------------------------------------------------------------------------------
struct SmallPiece {
  const uint64_t* key[MGB_MAX_KEYS];
  int n;
};

struct SmallParams {
  int n_pieces;
  int total;
  SmallPiece piece[SM_MAX_PIECES];
  const uint64_t* const* acc_ptrs;   // device table [piece][acc] (written by the host through pinned staging)
  OutCols out;                       // final destination columns
  uint64_t* tmp;                     // scratch rows for the sorted variant: [(nk + na)][total]
  int sort;
  long long* out_n;
};

template <int NK>
__global__ void __launch_bounds__(SM_THREADS, 1) k_merge_small(const SmallParams p) {
  extern __shared__ __align__(16) uint32_t ssm[];
  uint64_t* tab_keys = (uint64_t*)ssm;                        // [NK][SM_TAB]
  int* tab_idx = (int*)(tab_keys + MGB_MAX_KEYS * SM_TAB);    // [SM_TAB]
  int& n_unique = tab_idx[SM_TAB];
  for (int i = threadIdx.x; i < SM_TAB; i += SM_THREADS) tab_idx[i] = -1;
  if (threadIdx.x == 0) n_unique = 0;
  // when sorting, accumulate into scratch columns first, then scatter by rank
  OutCols acc = p.out;
  if (p.sort) {
    for (int j = 0; j < NK; ++j) acc.key[j] = p.tmp + (size_t)j * p.total;
    for (int a = 0; a < p.out.na; ++a) acc.acc[a] = p.tmp + (size_t)(NK + a) * p.total;
  }
  out_init(acc, p.total);
  __threadfence();
  __syncthreads();
  for (int i = threadIdx.x; i < p.total; i += SM_THREADS) {
    int pc = 0, r = i;
    while (r >= p.piece[pc].n) {
      r -= p.piece[pc].n;
      ++pc;
    }
    uint64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = p.piece[pc].key[j][r];
    const int idx = small_find_or_insert<NK>(tab_keys, tab_idx, &n_unique, k);
    if (idx < 0) continue;
#pragma unroll
    for (int j = 0; j < NK; ++j) acc.key[j][idx] = k[j];
    for (int a = 0; a < p.out.na; ++a) {
      const uint64_t* col = p.acc_ptrs[pc * p.out.na + a];
      out_accumulate(acc, a, idx, col[r]);
    }
  }
  __threadfence();
  __syncthreads();
  const int nu = n_unique;
  out_decode(acc, nu);
  if (threadIdx.x == 0) *p.out_n = nu;
  if (!p.sort) return;
  __threadfence();
  __syncthreads();
  // rank sort (nu <= 2048): rank = number of smaller key tuples; keys are unique
  for (int i = threadIdx.x; i < nu; i += SM_THREADS) {
    uint64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = ld_cg(acc.key[j] + i);
    int rank = 0;
    for (int q = 0; q < nu; ++q) {
      uint64_t o[NK];
#pragma unroll
      for (int j = 0; j < NK; ++j) o[j] = ld_cg(acc.key[j] + q);
      rank += key_less<NK>(o, k) ? 1 : 0;
    }
#pragma unroll
    for (int j = 0; j < NK; ++j) p.out.key[j][rank] = k[j];
    for (int a = 0; a < p.out.na; ++a) p.out.acc[a][rank] = ld_cg(acc.acc[a] + i);
  }
}

struct SmallScratch {
  uint64_t* tmp = nullptr;
  size_t tmp_words = 0;
  const uint64_t** ptrs_dev = nullptr;
  const uint64_t** ptrs_host = nullptr;   // pinned ring: SLOTS tables of SM_MAX_PIECES * MGB_MAX_ACCS pointers
  int slot = 0;
  long long* out_n = nullptr;
  long long* out_n_host = nullptr;
};
#define SMALL_SLOTS 8
static thread_local SmallScratch g_ss[8];

extern "C" int mgb_merge_small(const mgb_agg_spec* spec, int32_t n_pieces, const int64_t* piece_rows,
                               const int64_t* const* const* piece_key_cols,
                               const void* const* const* piece_acc_cols, int64_t* const* out_key_cols,
                               void* const* out_acc_cols, int64_t out_capacity, int32_t sort_keys,
                               int64_t* out_n_groups, void* stream) {
  MGB_REQUIRE(spec && piece_rows && piece_key_cols && piece_acc_cols && out_key_cols && out_acc_cols &&
                  out_n_groups, MGB_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  *out_n_groups = -1;
  const int nk = spec->n_keys, na = spec->n_accs;
  MGB_REQUIRE(nk >= 1 && nk <= MGB_MAX_KEYS && na >= 1 && na <= MGB_MAX_ACCS, MGB_ERR_UNSUPPORTED,
              "spec outside the supported envelope");
  long long total = 0;
  int used = 0;
  for (int i = 0; i < n_pieces; ++i) {
    MGB_REQUIRE(piece_rows[i] >= 0, MGB_ERR_INVALID, "negative piece size");
    total += piece_rows[i];
    if (piece_rows[i] > 0) ++used;
  }
  if (total == 0) {
    *out_n_groups = 0;
    return MGB_OK;
  }
  if (total > MGB_SMALL_MAX_ROWS || used > SM_MAX_PIECES) return MGB_OK;   // not small: general path
  MGB_REQUIRE(out_capacity >= total, MGB_ERR_INVALID, "output capacity %lld < %lld rows",
              (long long)out_capacity, total);
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  MGB_REQUIRE(dev >= 0 && dev < 8, MGB_ERR_UNSUPPORTED, "device ordinal %d", dev);
  SmallScratch& s = g_ss[dev];
  const size_t table_words = (size_t)SM_MAX_PIECES * MGB_MAX_ACCS;
  if (!s.ptrs_dev) {
    MGB_CUDA(cudaMalloc((void**)&s.ptrs_dev, table_words * 8 * SMALL_SLOTS));
    MGB_CUDA(cudaMallocHost((void**)&s.ptrs_host, table_words * 8 * SMALL_SLOTS));
    MGB_CUDA(cudaMalloc((void**)&s.out_n, 16));
    MGB_CUDA(cudaMallocHost((void**)&s.out_n_host, 16));
  }
  const size_t tmp_words = (size_t)(nk + na) * MGB_SMALL_MAX_ROWS;
  if (sort_keys && s.tmp_words < tmp_words) {
    if (s.tmp) cudaFree(s.tmp);
    MGB_CUDA(cudaMalloc((void**)&s.tmp, tmp_words * 8));
    s.tmp_words = tmp_words;
  }
  SmallParams p;
  memset(&p, 0, sizeof(p));
  const int slot = s.slot;
  s.slot = (s.slot + 1) % SMALL_SLOTS;   // every call synchronises below, so a slot is free when reused
  const uint64_t** host_tab = s.ptrs_host + (size_t)slot * table_words;
  int q = 0;
  for (int i = 0; i < n_pieces; ++i) {
    if (piece_rows[i] == 0) continue;
    for (int j = 0; j < nk; ++j) {
      MGB_REQUIRE(piece_key_cols[i] && piece_key_cols[i][j], MGB_ERR_INVALID, "null key column");
      p.piece[q].key[j] = (const uint64_t*)piece_key_cols[i][j];
    }
    p.piece[q].n = (int)piece_rows[i];
    for (int a = 0; a < na; ++a) {
      MGB_REQUIRE(piece_acc_cols[i] && piece_acc_cols[i][a], MGB_ERR_INVALID, "null accumulator column");
      host_tab[q * na + a] = (const uint64_t*)piece_acc_cols[i][a];
    }
    ++q;
  }
  p.n_pieces = q;
  p.total = (int)total;
  const uint64_t** dev_tab = s.ptrs_dev + (size_t)slot * table_words;
  MGB_CUDA(cudaMemcpyAsync(dev_tab, host_tab, (size_t)q * na * 8, cudaMemcpyHostToDevice, st));
  p.acc_ptrs = (const uint64_t* const*)dev_tab;
  p.out.na = na;
  for (int j = 0; j < nk; ++j) {
    MGB_REQUIRE(out_key_cols[j], MGB_ERR_INVALID, "null output key column");
    p.out.key[j] = (uint64_t*)out_key_cols[j];
  }
  for (int a = 0; a < na; ++a) {
    MGB_REQUIRE(out_acc_cols[a], MGB_ERR_INVALID, "null output accumulator column");
    const int op = spec->acc_op[a], c = spec->acc_col[a];
    MGB_REQUIRE(op == MGB_SUM || op == MGB_MIN || op == MGB_MAX || op == MGB_COUNT, MGB_ERR_INVALID,
                "merge op of accumulator %d must be sum/min/max", a);
    MGB_REQUIRE(c >= 0 && c < spec->n_vals, MGB_ERR_INVALID, "bad accumulator column");
    p.out.acc[a] = (uint64_t*)out_acc_cols[a];
    p.out.op[a] = (uint8_t)(op == MGB_MIN ? MGB_MIN : op == MGB_MAX ? MGB_MAX : MGB_SUM);
    p.out.dt[a] = (uint8_t)spec->val_dtype[c];
  }
  p.tmp = s.tmp;
  p.sort = sort_keys ? 1 : 0;
  p.out_n = s.out_n;
  {
    MgbProfScope prof(MGB_K_MERGE, st);
    const size_t smem = (size_t)MGB_MAX_KEYS * SM_TAB * 8 + SM_TAB * 4 + 64;
    if (nk == 1) {
      MGB_CUDA(cudaFuncSetAttribute(k_merge_small<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_merge_small<1><<<1, SM_THREADS, smem, st>>>(p);
    } else {
      MGB_CUDA(cudaFuncSetAttribute(k_merge_small<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_merge_small<2><<<1, SM_THREADS, smem, st>>>(p);
    }
    MGB_CHECK_LAUNCH();
  }
  MGB_CUDA(cudaMemcpyAsync(s.out_n_host, s.out_n, sizeof(long long), cudaMemcpyDeviceToHost, st));
  MGB_CUDA(cudaStreamSynchronize(st));
  *out_n_groups = s.out_n_host[0];
  return MGB_OK;
}
