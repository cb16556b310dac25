// libmarsgb: device code shared by the low-cardinality fast paths (see mgb_dense.cu for the overview).
#pragma once
#include "mgb_common.cuh"

#define DN_MAX_GC 16
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
// Called by ALL lanes of a warp together (`valid` = this lane has a key): a lane that meets a slot in the
// "being written" state does not spin on it -- the writer may be a lane of its own warp, parked at a
// reconvergence point -- it looks again in the next round, after the warp has reconverged at the vote.
template <int NK>
__device__ __forceinline__ int small_find_or_insert(uint64_t* tab_keys, int* tab_idx, int* n_unique, bool valid,
                                                    const uint64_t* k, uint32_t mask = SM_TAB - 1) {
  uint32_t s = small_hash<NK>(k) & mask;
  bool done = !valid;
  int idx = -1;
  uint32_t probe = 0;
  while (__any_sync(0xffffffffu, !done)) {
    if (!done) {
      int cur = *(volatile int*)&tab_idx[s];
      if (cur == -1) {
        cur = atomicCAS(&tab_idx[s], -1, -2);
        if (cur == -1) {
          tab_keys[s] = k[0];
          if (NK == 2) tab_keys[SM_TAB + s] = k[1];
          idx = atomicAdd(n_unique, 1);
          __threadfence_block();
          *(volatile int*)&tab_idx[s] = idx;
          done = true;
        }
      }
      if (!done && cur != -2) {   // cur == -2: the writer publishes within its round, look again
        bool eq = ((volatile uint64_t*)tab_keys)[s] == k[0];
        if (NK == 2) eq = eq && ((volatile uint64_t*)tab_keys)[SM_TAB + s] == k[1];
        if (eq) {
          idx = cur;
          done = true;
        } else {
          s = (s + 1) & mask;
          if (++probe > mask) done = true;   // table full (callers size it at >= 2x the rows)
        }
      }
    }
  }
  return idx;
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

// L2 load (bypasses the non-coherent L1): data written by other CTAs before a __threadfence + ticket.
// A plain intrinsic, not volatile asm: the compiler may batch several of them (memory-level parallelism).
__device__ __forceinline__ uint64_t ld_cg(const uint64_t* p) { return __ldcg((const unsigned long long*)p); }

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
#define DN_TAB 64             // key-table slots per CTA (open addressing; entries never move)

#define DN_MAX_CHUNKS 64      // row batches (chunks) one launch streams through
#define DN_PTRS (MGB_MAX_KEYS + 8)

// The row batches of one launch: a band's chunks are streamed as ONE virtual sequence of 1024-row tiles
// (full tiles of chunk 0, then of chunk 1, ...); the < 1024-row remainders are handled afterwards, chunk c by
// CTA c % grid.  Passed by value (kernel parameter space) and copied to shared memory by every CTA.
struct DenseBatch {
  int n_chunks;
  int pad;
  long long tile_end[DN_MAX_CHUNKS];              // full tiles of chunks 0..c (inclusive prefix sum)
  long long n[DN_MAX_CHUNKS];                     // rows of chunk c
  const uint64_t* ptr[DN_MAX_CHUNKS][DN_PTRS];    // key columns [0, NK), then the loaded value columns
};

// Row program of the fused pre-path (TPC-H Q1: filter + derived columns + groupby in one pass over lineitem,
// benchmarks/tpch/run_queries.py:110-163): the kernel loads n_phys PHYSICAL columns per row, drops the rows a
// comparison rejects, and computes the LOGICAL value columns it aggregates -- logical column c is physical column
// c itself (nf[c] == 0) or a left-to-right product of up to three affine factors alpha + beta * physical column,
// each operation rounded on its own (no FMA), which is what the reference's chain of pandas operators computes
// (`price * (1 - discount) * (1 + tax)`).
struct DenseProgram {
  int n_phys;               // physical columns loaded per row (<= NV); 0: no program
  int filter_col;           // physical column the row filter reads, -1: none
  int filter_op;            // 0 <=, 1 <, 2 >=, 3 >, 4 ==, 5 !=
  int filter_f64;           // compare as float64 (else int64)
  long long filter_i;
  double filter_d;
  signed char nf[8];
  signed char fcol[8][3];
  double falpha[8][3], fbeta[8][3];
};

struct DenseParams {
  DenseProgram prog;
  int GC;                   // groups per CTA (accumulator capacity); group ids are dense, in insertion order
  int stride_g;             // 32-bit words per (warp, group) accumulator block
  int off_nan, off_sum, off_sumsq, off_min, off_max;   // word offsets of column 0's slot inside a block
  uint32_t f64_mask;        // bit c: loaded column c is float64 (else int64)
  uint32_t op_mask[5];      // per mgb_op: bit c set when column c carries that accumulator
  // global staging: row = cta * GC + g; words: keys[NK] | rows | per column: nan, then the kinds in use
  uint64_t* stg;
  int stg_w;                // words per staged row
  int wpc;                  // words per column in a staged row
  int kind_off[5];          // word of kind k (1 sum, 2 sumsq, 3 min, 4 max) inside a column's group; 0 is the NaN count
  int word_kind[5];         // inverse: kind stored at word k of a column's group (word 0: NaN count)
  int* cta_ng;              // per CTA: number of groups staged
  unsigned int* ticket;
  int* fail;
  // output (spec order)
  OutCols out;
  uint8_t out_col[MGB_MAX_ACCS];   // loaded column of spec accumulator a, 0xFF: "rows of the group"
  uint8_t out_src[MGB_MAX_ACCS];   // mgb_op the staged value comes from
  long long* out_n;                // device: number of groups, -1 on failure
  const int* run_if;               // optional: the kernel exits at once unless *run_if != 0 (set by k_preagg_tiny when it
                                   // gave up on the band, mgb_tiny.cu)
  long long* dbg;                  // optional [grid][4] phase timestamps (ns), profiling experiments only
};

__device__ __forceinline__ long long dense_now() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int NK>
__device__ __forceinline__ uint32_t dense_hash(const uint64_t* k) {
  uint32_t h = (uint32_t)k[0] * 0x9E3779B1u + (uint32_t)(k[0] >> 32) * 0x85EBCA77u;
  if (NK == 2) h += (uint32_t)k[1] * 0xC2B2AE3Du + (uint32_t)(k[1] >> 32) * 0x27D4EB2Fu;
  h ^= h >> 15;
  return h * 0x2C1B3C6Du;
}

// lock-free probe of the CTA's key table starting at slot s; returns the group id or -1 when the probe
// sequence reaches an empty slot.  tgid[s] < 0: empty; an entry is published (tgid written) after its keys.
template <int NK>
__device__ __forceinline__ int dense_probe(const uint64_t* tkeys, const int* tgid, int s, uint64_t k0, uint64_t k1) {
  for (int probe = 0; probe < DN_TAB; ++probe) {
    const int tg = ((const volatile int*)tgid)[s];
    if (tg < 0) return -1;
    bool eq = ((const volatile uint64_t*)tkeys)[s] == k0;
    if (NK == 2) eq = eq && ((const volatile uint64_t*)tkeys)[DN_TAB + s] == k1;
    if (eq) return tg;
    s = (s + 1) & (DN_TAB - 1);
  }
  return -1;
}

// Warp-collective slow path of the group lookup (kept out of line: it runs a handful of times per CTA).
// Lanes with live && g < 0 hold keys that missed in the table; one distinct key at a time is looked up
// again (another warp may have inserted it meanwhile) and otherwise inserted by lane 0 under the CTA lock
// (a single try-lock per round, so waiting warps spin on lock-free probes, not on the lock word).
// Returns the lane's group id, or -2 when the CTA already holds GC groups.
template <int NK>
__device__ __noinline__ int dense_slow_lookup(uint64_t* tkeys, int* tgid, uint64_t* gkeys, int* s_ng, int* s_lock,
                                              int GC, uint64_t k0, uint64_t k1, bool live, int g) {
  const int lane = threadIdx.x & 31;
  unsigned need = __ballot_sync(0xffffffffu, live && g < 0);
  while (need) {
    const int src = __ffs(need) - 1;
    const uint64_t s0 = __shfl_sync(0xffffffffu, k0, src);
    const uint64_t s1 = __shfl_sync(0xffffffffu, k1, src);
    const uint64_t ks[2] = {s0, s1};
    const int home = (int)(dense_hash<NK>(ks) >> 26);
    int gi = -1;
    if (lane == 0) {
      gi = dense_probe<NK>(tkeys, tgid, home, s0, s1);
      if (gi < 0 && atomicCAS(s_lock, 0, 1) == 0) {
        __threadfence_block();
        int q = home;
        gi = -2;
        for (int probe = 0; probe < DN_TAB; ++probe) {
          const int tg = ((volatile int*)tgid)[q];
          if (tg < 0) {   // free slot: the key is new
            const int ng = *(volatile int*)s_ng;
            if (ng < GC) {
              ((volatile uint64_t*)tkeys)[q] = s0;
              ((volatile uint64_t*)gkeys)[ng] = s0;
              if (NK == 2) {
                ((volatile uint64_t*)tkeys)[DN_TAB + q] = s1;
                ((volatile uint64_t*)gkeys)[DN_MAX_GC + ng] = s1;
              }
              __threadfence_block();
              ((volatile int*)tgid)[q] = ng;
              *(volatile int*)s_ng = ng + 1;
              gi = ng;
            }
            break;
          }
          bool eq = ((volatile uint64_t*)tkeys)[q] == s0;
          if (NK == 2) eq = eq && ((volatile uint64_t*)tkeys)[DN_TAB + q] == s1;
          if (eq) {
            gi = tg;
            break;
          }
          q = (q + 1) & (DN_TAB - 1);
        }
        __threadfence_block();
        atomicExch(s_lock, 0);
      }
    }
    gi = __shfl_sync(0xffffffffu, gi, 0);
    if (gi == -2) return -2;
    if (gi >= 0 && live && g < 0 && k0 == s0 && (NK == 1 || k1 == s1)) g = gi;
    need = __ballot_sync(0xffffffffu, live && g < 0);   // gi == -1: lock was busy, go round again
  }
  return g;
}

// combine two decoded partial values of one accumulator: kind 1,2 add; 3 min; 4 max (f64 min/max ignore NaN,
// which is the "no value" marker of a partial)
__device__ __forceinline__ uint64_t dense_combine(int kind, bool f, uint64_t v, uint64_t y) {
  if (f) {
    const double a = __longlong_as_double((long long)v), b = __longlong_as_double((long long)y);
    const double r = kind <= 2 ? a + b : kind == 3 ? fmin(a, b) : fmax(a, b);
    return (uint64_t)__double_as_longlong(r);
  }
  const long long a = (long long)v, b = (long long)y;
  return (uint64_t)(kind <= 2 ? a + b : kind == 3 ? min(a, b) : max(a, b));
}

// PLAIN: every loaded column is float64 and carries SUM, no SUMSQ / MIN / MAX (sum / mean / count shapes):
// straight-line DADD updates.  Otherwise the generic body tests the per-column masks (warp-uniform branches).
// PROG (PLAIN shapes only): the row program of DenseProgram runs between the loads and the group update.
template <int NK, int NV, bool PLAIN, int THREADS, bool PROG = false>
__global__ void __launch_bounds__(THREADS, 1) k_preagg_dense(const DenseParams p, const DenseBatch bt) {
  constexpr int DN_WARPS = THREADS / 32;
  constexpr int DN_R = 1024 / THREADS;   // rows per thread per tile: a tile is always 1024 rows
  constexpr int DN_THREADS = THREADS;
  constexpr bool ALLOPS = !PLAIN;
  extern __shared__ __align__(16) uint32_t dsm[];
  if (p.run_if && *(volatile const int*)p.run_if == 0) return;   // the register-accumulator kernel took the band
  // layout: acc[warp][GC][stride_g] | tkeys[NK][TAB] (u64) | gkeys[NK][MAX_GC] (u64) | tgid[TAB] | ints
  const int GC = p.GC;
  uint32_t* acc = dsm;
  uint64_t* tkeys = (uint64_t*)(dsm + (size_t)DN_WARPS * GC * p.stride_g);
  uint64_t* gkeys = tkeys + MGB_MAX_KEYS * DN_TAB;
  int* tgid = (int*)(gkeys + MGB_MAX_KEYS * DN_MAX_GC);
  int* s_int = tgid + DN_TAB;
  // chunk table (behind the ints; the last-CTA merge re-uses the memory before it, never this part... it does
  // re-use everything, but only after the streaming loop is over)
  long long* s_tile_end = (long long*)(s_int + 32);                 // [DN_MAX_CHUNKS]
  long long* s_rows = s_tile_end + DN_MAX_CHUNKS;                   // [DN_MAX_CHUNKS]
  const uint64_t** s_ptr = (const uint64_t**)(s_rows + DN_MAX_CHUNKS);   // [DN_MAX_CHUNKS][DN_PTRS]
  int* s_ng = s_int;
  int* s_lock = s_int + 1;
  int* s_fail = s_int + 2;
  int* s_last = s_int + 3;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 4 + 0] = dense_now();
  uint32_t* wacc = acc + (size_t)warp * GC * p.stride_g;
  const uint32_t f64m = p.f64_mask;
  const uint32_t summ = p.op_mask[MGB_SUM], sqm = p.op_mask[MGB_SUMSQ], minm = p.op_mask[MGB_MIN],
                 maxm = p.op_mask[MGB_MAX];
  const int o_nan = p.off_nan, o_sum = p.off_sum >> 1, o_sq = p.off_sumsq >> 1, o_min = p.off_min >> 1,
            o_max = p.off_max >> 1;
  // ---- init accumulators: zeros, except min (+inf / INT64_MAX) and max (-inf / INT64_MIN)
  for (int i = lane; i < GC * p.stride_g; i += 32) wacc[i] = 0;
  if (ALLOPS) {
    __syncwarp();
    for (int g = 0; g < GC; ++g)
      for (int c = 0; c < NV; ++c) {
        uint64_t* b = (uint64_t*)(wacc + (size_t)g * p.stride_g);
        const bool f = (f64m >> c) & 1;
        if ((minm >> c) & 1u) b[o_min + c * 32 + lane] = f ? 0x7ff0000000000000ULL : 0x7fffffffffffffffULL;
        if ((maxm >> c) & 1u) b[o_max + c * 32 + lane] = f ? 0xfff0000000000000ULL : 0x8000000000000000ULL;
      }
  }
  for (int i = threadIdx.x; i < DN_TAB; i += DN_THREADS) tgid[i] = -1;
  const int n_chunks = bt.n_chunks;
  for (int i = threadIdx.x; i < n_chunks; i += DN_THREADS) {
    s_tile_end[i] = bt.tile_end[i];
    s_rows[i] = bt.n[i];
  }
  for (int i = threadIdx.x; i < n_chunks * DN_PTRS; i += DN_THREADS)
    s_ptr[i] = bt.ptr[i / DN_PTRS][i % DN_PTRS];
  if (threadIdx.x == 0) {
    *s_ng = 0;
    *s_lock = 0;
    *s_fail = 0;
    *s_last = 0;
  }
  __syncthreads();

  bool failed = false;

  // one row: group lookup, then lane-private accumulator updates
  // row program: physical values in x[] -> logical values in x[]; returns whether the row passes the filter.
  // Operand selection is a chain of warp-uniform predicated moves (registers cannot be indexed dynamically).
  auto apply_program = [&](uint64_t* x) -> bool {
    double ph[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) ph[c] = __longlong_as_double((long long)x[c]);
    bool keep = true;
    if (p.prog.filter_col >= 0) {
      uint64_t raw = 0;
#pragma unroll
      for (int c = 0; c < NV; ++c)
        if (p.prog.filter_col == c) raw = x[c];
      int cmp;   // -1, 0, 1 (NaN compares as "unordered": every comparison but != fails, like pandas)
      bool unordered = false;
      if (p.prog.filter_f64) {
        const double v = __longlong_as_double((long long)raw);
        unordered = v != v;
        cmp = v < p.prog.filter_d ? -1 : (v > p.prog.filter_d ? 1 : 0);
      } else {
        const long long v = (long long)raw;
        cmp = v < p.prog.filter_i ? -1 : (v > p.prog.filter_i ? 1 : 0);
      }
      const int op = p.prog.filter_op;
      keep = op == 0 ? cmp <= 0 : op == 1 ? cmp < 0 : op == 2 ? cmp >= 0 : op == 3 ? cmp > 0 : op == 4 ? cmp == 0 : cmp != 0;
      if (unordered) keep = op == 5;
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int nf = p.prog.nf[c];
      if (nf == 0) continue;            // logical column c == physical column c
      double v = 1.0;
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        if (f < nf) {
          double o = 0.0;
#pragma unroll
          for (int j = 0; j < NV; ++j)
            if (p.prog.fcol[c][f] == j) o = ph[j];
          const double term = __dadd_rn(p.prog.falpha[c][f], __dmul_rn(p.prog.fbeta[c][f], o));
          v = f == 0 ? term : __dmul_rn(v, term);
        }
      }
      x[c] = (uint64_t)__double_as_longlong(v);
    }
    return keep;
  };

  // group id of one row (all 32 lanes together); sets `failed` when the CTA would need more than GC groups
  auto find_group = [&](const uint64_t* k, const bool live) -> int {
    const int s = (int)(dense_hash<NK>(k) >> 26);
    // first probe without branches (the common case); further probes only for the lanes that need them
    const int tg = ((const volatile int*)tgid)[s];
    bool hit = tg >= 0 && ((const volatile uint64_t*)tkeys)[s] == k[0];
    if (NK == 2) hit = hit && ((const volatile uint64_t*)tkeys)[DN_TAB + s] == k[1];
    int g = hit ? tg : -1;
    if (__any_sync(0xffffffffu, live && !hit)) {
      if (live && !hit && tg >= 0)
        g = dense_probe<NK>(tkeys, tgid, (s + 1) & (DN_TAB - 1), k[0], NK == 2 ? k[1] : 0);
      if (__any_sync(0xffffffffu, live && g < 0)) {   // some lane met a key this CTA has not seen yet
        g = dense_slow_lookup<NK>(tkeys, tgid, gkeys, s_ng, s_lock, GC, k[0], NK == 2 ? k[1] : 0, live, g);
        if (__any_sync(0xffffffffu, g == -2)) failed = true;   // more distinct keys than group slots: not a dense chunk
      }
    }
    return g;
  };

  // PLAIN shapes, two rows of a lane at once: the accumulator read-modify-write chains of the two rows are
  // independent unless both rows belong to the same group, so all loads are issued first, the adds of the second
  // row take the first row's sums when the groups coincide (a select, no branch), and the stores follow in
  // order -- half the shared-memory latency per row of the one-row-at-a-time update.  NaN values (rare) take the
  // one-row path.  Same summation order per (lane, group) as before: (a + x0) + x1.
  auto update_pair_plain = [&](const int g0, const uint64_t* x0, const int g1, const uint64_t* x1) -> bool {
    bool anynan = false;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const double u = __longlong_as_double((long long)x0[c]), v = __longlong_as_double((long long)x1[c]);
      anynan = anynan || (u != u) || (v != v);
    }
    if (anynan) return false;
    uint32_t* b0 = wacc + g0 * p.stride_g;
    uint32_t* b1 = wacc + g1 * p.stride_g;
    const bool same = g0 == g1;
    const uint32_t r0 = b0[lane], r1 = b1[lane];
    double a0[NV], a1[NV];
    double* d0 = (double*)b0 + lane;
    double* d1 = (double*)b1 + lane;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      a0[c] = d0[o_sum + c * 32];
      a1[c] = d1[o_sum + c * 32];
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      a0[c] += __longlong_as_double((long long)x0[c]);
      a1[c] = (same ? a0[c] : a1[c]) + __longlong_as_double((long long)x1[c]);
    }
    b0[lane] = r0 + 1;
    b1[lane] = (same ? r0 + 1 : r1) + 1;
#pragma unroll
    for (int c = 0; c < NV; ++c) d0[o_sum + c * 32] = a0[c];
#pragma unroll
    for (int c = 0; c < NV; ++c) d1[o_sum + c * 32] = a1[c];
    return true;
  };

  auto process_row = [&](const uint64_t* k, const uint64_t* x, const bool live, int g_known = -1) {
    int g = g_known;
    if (g_known < 0) {
      g = find_group(k, live);
      if (failed) return;
    }
    if (!live) return;
    uint32_t* blk = wacc + g * p.stride_g;   // block base; lane-private columns inside
    blk[lane] += 1;                          // rows of the group
    double* d = (double*)blk + lane;
    long long* q = (long long*)blk + lane;
    if (PLAIN) {
      bool anynan = false;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const double xv = __longlong_as_double((long long)x[c]);
        anynan = anynan || (xv != xv);
      }
      if (!anynan) {   // distinct slots: independent LDS / DADD / STS chains
        double a[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) a[c] = d[o_sum + c * 32];
#pragma unroll
        for (int c = 0; c < NV; ++c) a[c] += __longlong_as_double((long long)x[c]);
#pragma unroll
        for (int c = 0; c < NV; ++c) d[o_sum + c * 32] = a[c];
      } else {         // rare: count(x) = rows - nan, sums skip NaN
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          const double xv = __longlong_as_double((long long)x[c]);
          if (xv != xv)
            atomicAdd(&blk[o_nan + c], 1u);
          else
            d[o_sum + c * 32] += xv;
        }
      }
      return;
    }
    // generic: NaN bookkeeping first (rare): count(x) = rows - nan, every accumulator skips NaN
    uint32_t nanm = 0;
#pragma unroll
    for (int c = 0; c < NV; ++c)
      if (((f64m >> c) & 1u) && mgb_is_nan_bits(x[c])) nanm |= 1u << c;
    if (nanm) {
#pragma unroll
      for (int c = 0; c < NV; ++c)
        if ((nanm >> c) & 1u) atomicAdd(&blk[o_nan + c], 1u);
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      if ((nanm >> c) & 1u) continue;
      if ((f64m >> c) & 1u) {
        const double xv = __longlong_as_double((long long)x[c]);
        if ((summ >> c) & 1u) d[o_sum + c * 32] += xv;
        if ((sqm >> c) & 1u) d[o_sq + c * 32] += xv * xv;
        if ((minm >> c) & 1u) d[o_min + c * 32] = fmin(d[o_min + c * 32], xv);
        if ((maxm >> c) & 1u) d[o_max + c * 32] = fmax(d[o_max + c * 32], xv);
      } else {
        const long long xv = (long long)x[c];
        if ((summ >> c) & 1u) q[o_sum + c * 32] += xv;
        if ((sqm >> c) & 1u) q[o_sq + c * 32] += xv * xv;
        if ((minm >> c) & 1u) q[o_min + c * 32] = min(q[o_min + c * 32], xv);
        if ((maxm >> c) & 1u) q[o_max + c * 32] = max(q[o_max + c * 32], xv);
      }
    }
  };

  constexpr long long TILE = (long long)DN_THREADS * DN_R;
  const long long nfull = s_tile_end[n_chunks - 1];   // full tiles of all chunks: the unpredicated, double-buffered loop

  uint64_t kb[2][DN_R][NK];
  uint64_t xb[2][DN_R][NV];
  const uint64_t* pk[NK];
  const uint64_t* pc[NV];
  int cur = -1, cc = 0;          // chunk whose pointers are in pk / pc; chunk cursor (tiles are visited in order)
  long long cbeg = 0;            // first tile of chunk cc

  auto load_tile = [&](int buf, long long t) {
    while (t >= s_tile_end[cc]) {     // uniform across the CTA
      cbeg = s_tile_end[cc];
      ++cc;
    }
    if (cc != cur) {
      cur = cc;
#pragma unroll
      for (int j = 0; j < NK; ++j) pk[j] = s_ptr[cc * DN_PTRS + j] + threadIdx.x;
#pragma unroll
      for (int c = 0; c < NV; ++c) pc[c] = s_ptr[cc * DN_PTRS + MGB_MAX_KEYS + c] + threadIdx.x;
    }
    const long long off = (t - cbeg) * TILE;
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      const uint64_t* b = pk[j] + off;
#pragma unroll
      for (int r = 0; r < DN_R; ++r) kb[buf][r][j] = __ldcs((const unsigned long long*)b + r * DN_THREADS);
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      if (PROG && c >= p.prog.n_phys) {
#pragma unroll
        for (int r = 0; r < DN_R; ++r) xb[buf][r][c] = 0;
        continue;
      }
      const uint64_t* b = pc[c] + off;
#pragma unroll
      for (int r = 0; r < DN_R; ++r) xb[buf][r][c] = __ldcs((const unsigned long long*)b + r * DN_THREADS);
    }
  };
  auto process_tile = [&](int buf) {
    if (PROG) {     // filter + derived columns first; rows the filter drops take no part in the update
      auto rows = [&](uint64_t (&kk)[DN_R][NK], uint64_t (&xx)[DN_R][NV]) {
#pragma unroll
        for (int r = 0; r < DN_R; ++r) {
          if (failed) return;
          const bool keep = apply_program(xx[r]);
          process_row(kk[r], xx[r], keep);
        }
      };
      if (buf == 0)       // constant buffer index on either side: the tile buffers stay in registers
        rows(kb[0], xb[0]);
      else
        rows(kb[1], xb[1]);
      return;
    }
    if (PLAIN) {
#pragma unroll
      for (int r = 0; r < DN_R; r += 2) {
        if (failed) return;
        const int g0 = find_group(kb[buf][r], true);
        if (failed) return;
        const int g1 = find_group(kb[buf][r + 1], true);
        if (failed) return;
        if (!update_pair_plain(g0, xb[buf][r], g1, xb[buf][r + 1])) {
          process_row(kb[buf][r], xb[buf][r], true, g0);
          process_row(kb[buf][r + 1], xb[buf][r + 1], true, g1);
        }
      }
      return;
    }
#pragma unroll
    for (int r = 0; r < DN_R; ++r) {
      if (failed) return;
      process_row(kb[buf][r], xb[buf][r], true);
    }
  };

  long long t = blockIdx.x;
  if (t < nfull) load_tile(0, t);
  int iter = 0;
  while (t < nfull && !failed) {
    const long long t1 = t + gridDim.x;
    if (t1 < nfull) load_tile(1, t1);
    process_tile(0);
    if (failed || t1 >= nfull) break;
    const long long t2 = t1 + gridDim.x;
    if (t2 < nfull) load_tile(0, t2);
    process_tile(1);
    t = t2;
    if ((++iter & 7) == 0 && *(volatile int*)p.fail) failed = true;   // somebody else gave up: stop early
  }
  // ragged remainders (< TILE rows per chunk): chunk c by CTA c % grid, a warp-sized slice at a time; lanes
  // past the end stay in the warp-collective lookup but neither look up nor accumulate
  for (int c = (int)blockIdx.x; c < n_chunks && !failed; c += (int)gridDim.x) {
    const long long n = s_rows[c];
    const long long done = (s_tile_end[c] - (c ? s_tile_end[c - 1] : 0)) * TILE;
    const uint64_t* const* cp = s_ptr + c * DN_PTRS;
    for (long long i0 = done + warp * 32; i0 < n && !failed; i0 += DN_THREADS) {
      const long long i = i0 + lane;
      const bool live = i < n;
      uint64_t k[NK], x[NV];
#pragma unroll
      for (int j = 0; j < NK; ++j) k[j] = live ? cp[j][i] : 0;
#pragma unroll
      for (int cl = 0; cl < NV; ++cl) x[cl] = (live && !(PROG && cl >= p.prog.n_phys)) ? cp[MGB_MAX_KEYS + cl][i] : 0;
      bool keep = live;
      if (PROG) keep = apply_program(x) && live;
      process_row(k, x, keep);
    }
  }
  if (failed) {
    *s_fail = 1;
    *p.fail = 1;
  }
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 4 + 1] = dense_now();
  const bool cta_failed = *s_fail != 0;

  // ---- CTA reduction -> staging rows (one per group of this CTA), decoded values.  Rows are staged in key
  //      order, so CTAs that saw the same key set (the common case) stage the same key at the same position.
  const int ng = cta_failed ? 0 : *(volatile int*)s_ng;
  const int W = p.stg_w, WPC = p.wpc;
  // staging is transposed: the values of one (row, word) over all CTAs are contiguous, so that the last CTA
  // reads them as full sectors
#define STG(cta, r, w) (p.stg + ((size_t)((r) * W + (w)) * gridDim.x + (cta)))
  int* s_rank = s_int + 8;   // [MAX_GC]
  if (threadIdx.x < ng) {
    uint64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = gkeys[j * DN_MAX_GC + threadIdx.x];
    int r = 0;
    for (int q = 0; q < ng; ++q) {
      uint64_t o[NK];
#pragma unroll
      for (int j = 0; j < NK; ++j) o[j] = gkeys[j * DN_MAX_GC + q];
      r += key_less<NK>(o, k) ? 1 : 0;
    }
    s_rank[threadIdx.x] = r;
  }
  __syncthreads();
  constexpr int SLOTS = 1 + NV * 5;
  for (int qq = warp; qq < ng * SLOTS; qq += DN_WARPS) {
    const int g = qq / SLOTS, s = qq - g * SLOTS;
    const int sr = s_rank[g];
    if (s == 0) {   // rows
      uint32_t v = 0;
      for (int w = 0; w < DN_WARPS; ++w) v += acc[((size_t)w * GC + g) * p.stride_g + lane];
      v = __reduce_add_sync(0xffffffffu, v);
      if (lane == 0) *STG(blockIdx.x, sr, NK) = v;
      continue;
    }
    const int c = (s - 1) / 5, kind = (s - 1) % 5;   // kind: 0 nan, 1 sum, 2 sumsq, 3 min, 4 max
    if (kind == 0) {
      uint32_t v = lane < DN_WARPS ? acc[((size_t)lane * GC + g) * p.stride_g + o_nan + c] : 0u;
      v = __reduce_add_sync(0xffffffffu, v);
      if (lane == 0) *STG(blockIdx.x, sr, NK + 1 + c * WPC) = v;
      continue;
    }
    if (!ALLOPS && kind > 1) continue;
    const uint32_t m = kind == 1 ? summ : kind == 2 ? sqm : kind == 3 ? minm : maxm;
    if (!((m >> c) & 1u)) continue;
    const int off = kind == 1 ? o_sum : kind == 2 ? o_sq : kind == 3 ? o_min : o_max;
    const bool f = (f64m >> c) & 1u;
    // two independent chains over the warps' private copies, then the lanes
    uint64_t v0 = ((const uint64_t*)(acc + ((size_t)0 * GC + g) * p.stride_g))[off + c * 32 + lane];
    uint64_t v1 = ((const uint64_t*)(acc + ((size_t)1 * GC + g) * p.stride_g))[off + c * 32 + lane];
    for (int w = 2; w < DN_WARPS; w += 2) {
      v0 = dense_combine(kind, f, v0, ((const uint64_t*)(acc + ((size_t)w * GC + g) * p.stride_g))[off + c * 32 + lane]);
      v1 = dense_combine(kind, f, v1, ((const uint64_t*)(acc + ((size_t)(w + 1) * GC + g) * p.stride_g))[off + c * 32 + lane]);
    }
    uint64_t v = dense_combine(kind, f, v0, v1);
    for (int o = 16; o > 0; o >>= 1) v = dense_combine(kind, f, v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) *STG(blockIdx.x, sr, NK + 1 + c * WPC + p.kind_off[kind]) = v;
  }
  for (int qq = threadIdx.x; qq < ng * NK; qq += DN_THREADS) {
    const int g = qq / NK, j = qq - g * NK;
    *STG(blockIdx.x, s_rank[g], j) = gkeys[j * DN_MAX_GC + g];
  }
  if (threadIdx.x == 0) p.cta_ng[blockIdx.x] = ng;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int tk = atomicAdd(p.ticket, 1u);
    *s_last = (tk == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 4 + 2] = dense_now();
  if (!*s_last) return;

  // ---- last CTA: merge all staged rows into unique output rows.  Shared memory is re-used:
  //      tab_keys[NK][SM_TAB] | tab_idx[SM_TAB] | ints[8] | items[grid*GC] | item_idx[grid*GC] | sbuf[...]
  __threadfence();
  if (*(volatile int*)p.fail) {
    if (threadIdx.x == 0) *p.out_n = -1;
    return;
  }
  uint64_t* tab_keys = (uint64_t*)dsm;
  int* tab_idx = (int*)(tab_keys + MGB_MAX_KEYS * SM_TAB);
  int* n_unique = tab_idx + SM_TAB;
  int* n_items = n_unique + 1;
  int* items = n_unique + 8;
  const int n_slots = (gridDim.x * GC + 1) & ~1;   // even: keeps sbuf 8-byte aligned
  int* item_idx = items + n_slots;
  uint64_t* sbuf = (uint64_t*)(item_idx + n_slots);
  if (p.dbg && threadIdx.x == 0) p.dbg[gridDim.x * 4 + 0] = dense_now();
  // ---- fast path: every CTA staged the same keys (in key order).  Then staged row r of every CTA belongs
  //      to output row r: 8 lanes per (row, word) sum that word over the CTAs, no hashing, no sorting.  All
  //      loads of a phase are independent and issued back to back (bounded, fully unrolled loops): the
  //      phase costs one L2 round trip.
  if (gridDim.x <= 8 * 20) {
    const int G = (int)gridDim.x;
    const int ng0 = ((volatile int*)p.cta_ng)[0];
    uint64_t* red = sbuf;                 // [ng0][W]
    uint64_t* k0 = sbuf + DN_MAX_GC * W;  // [ng0][NK]: keys of CTA 0
    bool same = true;
    for (int cta = threadIdx.x; cta < G; cta += DN_THREADS) same = same && ((volatile int*)p.cta_ng)[cta] == ng0;
    for (int i = threadIdx.x; i < ng0 * NK; i += DN_THREADS) k0[i] = ld_cg(STG(0, i / NK, i % NK));
    __syncthreads();
    if (p.dbg && threadIdx.x == 0) p.dbg[gridDim.x * 4 + 1] = dense_now();
    {
      const int items = (G - 1) * ng0 * NK;
      uint64_t got[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = threadIdx.x + u * DN_THREADS;
        got[u] = 0;
        if (i < items) {   // consecutive threads read consecutive CTAs of one (row, key word): full sectors
          const int rj = i / (G - 1), cta = 1 + (i - rj * (G - 1));
          got[u] = ld_cg(STG(cta, rj / NK, rj % NK));
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = threadIdx.x + u * DN_THREADS;
        if (i < items) same = same && got[u] == k0[i / (G - 1)];
      }
      if (items > 8 * DN_THREADS) same = false;   // more keys than this path is unrolled for: generic path
    }
    if (__syncthreads_and(same ? 1 : 0)) {
      if (p.dbg && threadIdx.x == 0) p.dbg[gridDim.x * 4 + 2] = dense_now();
      const int nw = W - NK;
      const int sl8 = lane & 7;
      for (int pair0 = 0; pair0 < ng0 * nw; pair0 += DN_WARPS * 4) {
        const int pair = pair0 + warp * 4 + (lane >> 3);
        const bool on = pair < ng0 * nw;
        const int r = on ? pair / nw : 0, w = NK + (on ? pair - r * nw : 0);
        int kind = 1;      // rows / NaN counts: integer adds
        bool f = false;
        if (w > NK) {
          const int c = (w - NK - 1) / WPC, k = (w - NK - 1) - c * WPC;
          if (k > 0) {
            kind = p.word_kind[k];
            f = (f64m >> c) & 1u;
          }
        }
        uint64_t v = 0;
        if (kind == 3) v = f ? 0x7ff8000000000000ULL : 0x7fffffffffffffffULL;
        if (kind == 4) v = f ? 0x7ff8000000000000ULL : 0x8000000000000000ULL;
        uint64_t x[20];
#pragma unroll
        for (int u = 0; u < 20; ++u) {
          const int cta = sl8 + u * 8;
          x[u] = (on && cta < G) ? ld_cg(STG(cta, r, w)) : v;
        }
#pragma unroll
        for (int u = 0; u < 20; ++u)
          if (sl8 + u * 8 < G) v = dense_combine(kind, f, v, x[u]);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) v = dense_combine(kind, f, v, __shfl_xor_sync(0xffffffffu, v, o));
        if (on && sl8 == 0) red[r * W + w] = v;
      }
      __syncthreads();
      if (p.dbg && threadIdx.x == 0) p.dbg[gridDim.x * 4 + 3] = dense_now();
      const int na = p.out.na;
      for (int i = threadIdx.x; i < ng0 * na; i += DN_THREADS) {
        const int r = i / na, a = i - r * na;
        const int c = p.out_col[a], src = p.out_src[a];
        const uint64_t rows = red[r * W + NK];
        const bool fcol = c != 0xFF && ((f64m >> c) & 1u);
        const uint64_t nanc = fcol ? red[r * W + NK + 1 + c * WPC] : 0;
        uint64_t v;
        if (src == MGB_COUNT) {
          v = rows - nanc;
        } else {
          const int kind = src == MGB_SUM ? 1 : src == MGB_SUMSQ ? 2 : src == MGB_MIN ? 3 : 4;
          v = red[r * W + NK + 1 + c * WPC + p.kind_off[kind]];
          if (kind >= 3 && fcol && rows == nanc) v = 0x7ff8000000000000ULL;   // no value at all: NaN
        }
        p.out.acc[a][r] = v;
      }
      for (int i = threadIdx.x; i < ng0 * NK; i += DN_THREADS) p.out.key[i % NK][i / NK] = k0[i];
      if (threadIdx.x == 0) *p.out_n = ng0;
      if (p.dbg && threadIdx.x == 0) p.dbg[gridDim.x * 4 + 4] = dense_now();
      if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 4 + 3] = dense_now();
      return;
    }
  }
  for (int i = threadIdx.x; i < SM_TAB; i += DN_THREADS) tab_idx[i] = -1;
  if (threadIdx.x == 0) {
    *n_unique = 0;
    *n_items = 0;
  }
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) p.dbg[gridDim.x * 4 + 1] = dense_now();
  for (int cta = threadIdx.x; cta < (int)gridDim.x; cta += DN_THREADS) {
    const int cnt = ((volatile int*)p.cta_ng)[cta];
    int pos = cnt ? atomicAdd(n_items, cnt) : 0;
    for (int g = 0; g < cnt; ++g) items[pos++] = cta * GC + g;
  }
  __syncthreads();
  const int total = *n_items;
  if (p.dbg && threadIdx.x == 0) p.dbg[gridDim.x * 4 + 2] = dense_now();
  // pass 1: dense output row of every staged row (shared-memory hash table on the key tuple)
  for (int it0 = 0; it0 < total; it0 += DN_THREADS) {   // uniform trip count: the insert votes per warp
    const int it = it0 + (int)threadIdx.x;
    const bool valid = it < total;
    uint64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = 0;
    if (valid) {
      const int icta = items[it] / GC, ir = items[it] - icta * GC;
#pragma unroll
      for (int j = 0; j < NK; ++j) k[j] = ld_cg(STG(icta, ir, j));
    }
    const int idx = small_find_or_insert<NK>(tab_keys, tab_idx, n_unique, valid, k);
    if (valid) {
      item_idx[it] = idx;
#pragma unroll
      for (int j = 0; j < NK; ++j) p.out.key[j][idx] = k[j];   // same value from every writer
    }
  }
  __syncthreads();
  const int nu = *n_unique;
  const int na = p.out.na;
  if (p.dbg && threadIdx.x == 0) p.dbg[gridDim.x * 4 + 3] = dense_now();
  // pass 2: counting sort of the staged rows by output row (shared memory), then 8 lanes per
  // (output row, accumulator) pair reduce that row's contributions straight from L2 and combine by shuffles
  int* ucnt = (int*)sbuf;              // [nu + 1] counts -> exclusive starts
  int* ufill = ucnt + (nu + 1);        // [nu]
  int* sorted = ufill + nu;            // [total] item numbers grouped by output row
  for (int i = threadIdx.x; i <= nu; i += DN_THREADS) ucnt[i] = 0;
  for (int i = threadIdx.x; i < nu; i += DN_THREADS) ufill[i] = 0;
  __syncthreads();
  for (int it = threadIdx.x; it < total; it += DN_THREADS) atomicAdd(&ucnt[item_idx[it]], 1);
  __syncthreads();
  if (warp == 0) {   // exclusive scan of ucnt[0..nu] by one warp, 32 entries per step
    int carry = 0;
    for (int base = 0; base <= nu; base += 32) {
      const int i = base + lane;
      const int v = i <= nu ? ucnt[i] : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (i <= nu) ucnt[i] = carry + x - v;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
  }
  __syncthreads();
  for (int it = threadIdx.x; it < total; it += DN_THREADS) {
    const int u = item_idx[it];
    sorted[ucnt[u] + atomicAdd(&ufill[u], 1)] = it;
  }
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) p.dbg[gridDim.x * 4 + 4] = dense_now();
  const int sl = lane & 7, pl = lane >> 3;
  for (int pair0 = 0; pair0 < nu * na; pair0 += DN_WARPS * 4) {
    const int pair = pair0 + warp * 4 + pl;
    const bool on = pair < nu * na;
    const int u = on ? pair / na : 0, a = on ? pair - u * na : 0;
    const int c = p.out_col[a], src = p.out_src[a];
    const bool fcol = c != 0xFF && ((f64m >> c) & 1u);
    const bool f = src != MGB_COUNT && fcol;
    const int kind = src == MGB_SUM ? 1 : src == MGB_SUMSQ ? 2 : src == MGB_MIN ? 3 : src == MGB_MAX ? 4 : 1;
    const int voff = NK + 1 + (c == 0xFF ? 0 : c * WPC + p.kind_off[kind]);
    const int noff = NK + 1 + (c == 0xFF ? 0 : c * WPC);
    // identity: 0 for sums / counts, NaN (ignored by fmin / fmax) for float min / max, extremes for int
    uint64_t v = 0;
    if (kind == 3) v = f ? 0x7ff8000000000000ULL : 0x7fffffffffffffffULL;
    if (kind == 4) v = f ? 0x7ff8000000000000ULL : 0x8000000000000000ULL;
    const int q1 = on ? ucnt[u + 1] : 0;
#pragma unroll 4
    for (int q = (on ? ucnt[u] : 0) + sl; q < q1; q += 8) {
      const int item = items[sorted[q]];
      const int icta = item / GC, ir = item - icta * GC;
      const uint64_t rows = ld_cg(STG(icta, ir, NK));
      const uint64_t nanc = fcol ? ld_cg(STG(icta, ir, noff)) : 0;
      const uint64_t x = src == MGB_COUNT ? 0 : ld_cg(STG(icta, ir, voff));
      if (src == MGB_COUNT)
        v += rows - nanc;
      else if (kind <= 2 || rows != nanc)   // min / max of an all-NaN partial: no value
        v = dense_combine(kind, f, v, x);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) v = dense_combine(kind, f, v, __shfl_xor_sync(0xffffffffu, v, o));
    if (on && sl == 0) p.out.acc[a][u] = v;
  }
  if (threadIdx.x == 0) *p.out_n = nu;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 4 + 3] = dense_now();
}

// one translation unit per (NK, THREADS) pair instantiates the kernels (mgb_dense_inst.cu)
template <int NK, int TH>
int mgb_dense_launch(const DenseParams& p, const DenseBatch& bt, int ncol, bool plain, int grid, size_t smem,
                     cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
#define LAUNCH(NV, PL)                                                                                       \
  do {                                                                                                      \
    static bool attr_set[8] = {false};   /* per device: the opt-in is sticky, set it once */                \
    if (!attr_set[dev & 7]) {                                                                               \
      MGB_CUDA(cudaFuncSetAttribute(k_preagg_dense<NK, NV, PL, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    227 * 1024));                                                           \
      attr_set[dev & 7] = true;                                                                             \
    }                                                                                                       \
    k_preagg_dense<NK, NV, PL, TH><<<grid, TH, smem, st>>>(p, bt);                                             \
  } while (0)
#define PICK_NV(PL)                     \
  do {                                  \
    switch (ncol) {                     \
      case 1: LAUNCH(1, PL); break;     \
      case 2: LAUNCH(2, PL); break;     \
      case 3: LAUNCH(3, PL); break;     \
      case 4: LAUNCH(4, PL); break;     \
      case 5: LAUNCH(5, PL); break;     \
      case 6: LAUNCH(6, PL); break;     \
      case 7: LAUNCH(7, PL); break;     \
      default: LAUNCH(8, PL); break;    \
    }                                   \
  } while (0)
  if (p.prog.n_phys > 0) {
#define LAUNCH_PROG(NV)                                                                                        \
  do {                                                                                                        \
    static bool attr_set[8] = {false};                                                                        \
    if (!attr_set[dev & 7]) {                                                                                 \
      MGB_CUDA(cudaFuncSetAttribute(k_preagg_dense<NK, NV, true, TH, true>,                                    \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));                \
      attr_set[dev & 7] = true;                                                                               \
    }                                                                                                         \
    k_preagg_dense<NK, NV, true, TH, true><<<grid, TH, smem, st>>>(p, bt);                                     \
  } while (0)
    switch (ncol) {
      case 1: LAUNCH_PROG(1); break;
      case 2: LAUNCH_PROG(2); break;
      case 3: LAUNCH_PROG(3); break;
      case 4: LAUNCH_PROG(4); break;
      case 5: LAUNCH_PROG(5); break;
      case 6: LAUNCH_PROG(6); break;
      case 7: LAUNCH_PROG(7); break;
      default: LAUNCH_PROG(8); break;
    }
#undef LAUNCH_PROG
  } else if (plain) PICK_NV(true); else PICK_NV(false);
#undef PICK_NV
#undef LAUNCH
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}
