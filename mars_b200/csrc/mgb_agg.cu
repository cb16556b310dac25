// libmarsgb: hash aggregation (a4 map-side pre-aggregation, a9 combine, a10 agg) for sm_100a.
//
// Replaces `df.groupby(keys).agg(...)` in DataFrameGroupByAgg._execute_map/_execute_combine/_execute_agg
// (mars/dataframe/groupby/aggregation.py:1043-1267).  Data flow of one aggregator:
//
//   update(batch)  -> phase 1, asynchronous, one pass over the batch's key + value columns
//       k_preagg_hash  : sum / mean / count shapes.  Per-CTA shared-memory hash table (linear probing, 64-bit
//                        CAS claims), register double-buffered tiles; rows with a NaN / marker key and the
//                        misses of a frozen (3/4 full) table go to a row-index spill list.
//       k_preagg_smem  : every other shape (min / max / var, int columns, > 8 columns), descriptor driven;
//                        a full table is flushed to the staging list while that pays off, else frozen.
//       (skipped entirely when the caller set the distinct-keys hint)
//   finalize()     -> phase 2, sizes the global table from the exact number of staged + spilled rows
//       k_merge_staged        : staged partial rows -> global open-addressing table (64/128-bit CAS claims,
//                               RED.ADD.F64 / atomicMin / atomicMax on the accumulators)
//       k_update_global[_plain]: spilled raw rows   -> same table (gather through the spill list)
//   emit()         -> k_emit_count / scan / k_emit_write compaction (+ radix argsort for sort=True)
#include <new>
#include <vector>

#include "mgb_common.cuh"
#include "mgb_layout.cuh"

#define GROUP_COLS 8         // value columns handled per kernel instantiation
#define SMEM_THREADS 512
#define SMEM_ROWS 2          // rows per thread per tile (smem kernel)
#define SMEM_CHECK_EVERY 4   // tiles between fill checks
#define FLUSH_GAIN 8         // flush+reset only if >= FLUSH_GAIN input rows were absorbed per table entry
#define GLOBAL_THREADS 256
#define GLOBAL_ROWS 2
#define SMEM_BUDGET (200 * 1024)
#define MAX_INTER_ROWS (4LL << 20)
#define BUCKET_MIN_ROWS 4096 // smaller direct inputs take the hash-table path

struct Ctl {                 // per update, device memory, zero-initialised
  unsigned long long stg_count;   // staged partial rows
  unsigned long long miss_count[MGB_MAX_VALS / GROUP_COLS];  // spilled rows per column group
  int err;
};

struct GlobalCtl {           // per aggregator
  unsigned long long n_groups;
  int special_used;
  int err;
};

struct Batch {               // one update() call, passed by value
  const uint64_t* keys[MGB_MAX_KEYS];
  const uint64_t* cols[GROUP_COLS];   // value columns of this column group
  int c0, ncol;                       // first column / number of columns in the group
  int group;                          // column group index
  long long n;
};

struct Staging {             // partial rows, SoA: column j at base + j*cap  (j < nk: key j, else acc j-nk)
  uint64_t* base;
  long long cap;
};

// ----------------------------------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_stage(const Staging& s, int col, long long row, uint64_t v) {
  s.base[(long long)col * s.cap + row] = v;
}

__device__ __forceinline__ ulonglong2 ld_v2_volatile(const uint64_t* p) {
  ulonglong2 r;
  asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
  return r;
}

__device__ __forceinline__ ulonglong2 cas128(uint64_t* p, ulonglong2 cmp, ulonglong2 val) {
  ulonglong2 r;
  asm volatile(
      "{\n .reg .b128 c, n, o;\n mov.b128 c, {%2,%3};\n mov.b128 n, {%4,%5};\n"
      " atom.global.cas.b128 o, [%6], c, n;\n mov.b128 {%0,%1}, o;\n}"
      : "=l"(r.x), "=l"(r.y)
      : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(p)
      : "memory");
  return r;
}

// Find or claim the global slot of a key tuple.  Table: (C+1) slots of W words; slot C is reserved for the
// tuple whose every component equals MGB_EMPTY_KEY (the empty marker).  Returns -1 if the table is full.
template <int NK>
__device__ __forceinline__ long long global_find_or_insert(uint64_t* table, int W, unsigned long long Cmask,
                                                           int shift, const uint64_t* k, uint64_t h,
                                                           GlobalCtl* g, int* inserted) {
  const uint64_t E = (uint64_t)MGB_EMPTY_KEY;
  bool has_empty = (k[0] == E);
  bool all_empty = (k[0] == E);
  if (NK == 2) {
    has_empty = has_empty || (k[1] == E);
    all_empty = all_empty && (k[1] == E);
  }
  if (all_empty) {
    if (atomicExch(&g->special_used, 1) == 0) *inserted += 1;
    return (long long)(Cmask + 1);
  }
  unsigned long long idx = mgb_slot_hash(h) >> shift;
  for (unsigned long long probe = 0; probe <= Cmask; ++probe, idx = (idx + 1) & Cmask) {
    uint64_t* slot = table + idx * (unsigned long long)W;
    if (NK == 1) {
      uint64_t cur = *(volatile uint64_t*)slot;
      if (cur == k[0]) return (long long)idx;
      if (cur == E) {
        uint64_t old = atomicCAS((unsigned long long*)slot, (unsigned long long)E, (unsigned long long)k[0]);
        if (old == E) {
          *inserted += 1;
          return (long long)idx;
        }
        if (old == k[0]) return (long long)idx;
      }
    } else {
      ulonglong2 cur;
      // a plain 16-byte load may tear while another thread claims the slot (EMPTY,EMPTY)->(a,b); a torn value
      // can only alias our tuple if our tuple contains the EMPTY marker, in which case we ask the CAS.
      if (has_empty)
        cur = cas128(slot, make_ulonglong2(E, E), make_ulonglong2(k[0], k[1]));
      else
        cur = ld_v2_volatile(slot);
      if (has_empty && cur.x == E && cur.y == E) {
        *inserted += 1;
        return (long long)idx;
      }
      if (cur.x == k[0] && cur.y == k[1]) return (long long)idx;
      if (!has_empty && (cur.x == E || cur.y == E)) {
        ulonglong2 old = cas128(slot, make_ulonglong2(E, E), make_ulonglong2(k[0], k[1]));
        if (old.x == E && old.y == E) {
          *inserted += 1;
          return (long long)idx;
        }
        if (old.x == k[0] && old.y == k[1]) return (long long)idx;
      }
    }
  }
  return -1;
}

__device__ __forceinline__ void block_count_groups(int inserted, GlobalCtl* g) {
  // one atomic per warp at kernel end
  for (int o = 16; o > 0; o >>= 1) inserted += __shfl_xor_sync(0xffffffffu, inserted, o);
  if ((threadIdx.x & 31) == 0 && inserted) atomicAdd(&g->n_groups, (unsigned long long)inserted);
}

// ----------------------------------------------------------------------------------------------------
// table init
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_table_init(uint64_t* table, Layout L, long long n_slots) {
  const long long total = n_slots * L.W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % L.W);
    uint64_t v = 0;
    if (w < L.nk)
      v = (uint64_t)MGB_EMPTY_KEY;
    else if (w - L.nk < L.na)
      v = mgb_acc_init(L.spec_op[w - L.nk]);  // table word order == spec accumulator order
    table[i] = v;
  }
}

// ----------------------------------------------------------------------------------------------------
// phase 1a: lean shared-memory hash pre-aggregation for the sum / mean / count shapes (PLAIN: every value
// column float64 with SUM, no SUMSQ / MIN / MAX).  One pass, register double-buffered tiles of 1024 rows,
// 512 threads per CTA, one CTA per SM.  Per-CTA table: keys[NK][cap] | rows[cap] (u32) | sum[NV][cap] (f64),
// linear probing, 64-bit CAS claims.  count(x) == rows of the group because rows that carry a NaN (or a key
// equal to the EMPTY marker) are not aggregated here but appended to the spill list, which the generic
// global-table kernel handles.  When the table is 3/4 full it is frozen: hits keep aggregating on chip,
// misses are spilled.  At the end every occupied slot becomes one partial row of the staging list.
// ----------------------------------------------------------------------------------------------------
#define HASH_THREADS 512
#define HASH_R 2

struct HashParams {
  const uint64_t* keys[MGB_MAX_KEYS];
  const uint64_t* cols[8];
  long long n;
  int cap;                       // table slots (power of two)
  int na;                        // spec accumulators
  uint8_t acc_col[MGB_MAX_ACCS]; // spec accumulator a: value column it reads
  uint8_t acc_op[MGB_MAX_ACCS];  // MGB_SUM or MGB_COUNT
  Ctl* ctl;
  Staging stg;                   // column j < nk: key j, else spec accumulator j - nk
  int32_t* miss;                 // spill list (row indices)
};

template <int NK, int NV>
__global__ void __launch_bounds__(HASH_THREADS, 1) k_preagg_hash(const HashParams p) {
  extern __shared__ __align__(16) uint64_t hsm[];
  const int cap = p.cap;
  const uint32_t mask = (uint32_t)cap - 1u;
  uint64_t* sk = hsm;                                   // [NK][cap]
  double* ssum = (double*)(sk + (size_t)NK * cap);      // [NV][cap]
  uint32_t* srows = (uint32_t*)(ssum + (size_t)NV * cap);   // [cap]
  int* s_int = (int*)(srows + cap);
  int* s_count = s_int;   // occupied slots
  int* s_cursor = s_int + 1;
  long long* s_pos = (long long*)(s_int + 2);
  const uint64_t E = (uint64_t)MGB_EMPTY_KEY;

  for (int i = threadIdx.x; i < cap; i += HASH_THREADS) {
    sk[i] = E;
    if (NK == 2) sk[cap + i] = E;
    srows[i] = 0;
#pragma unroll
    for (int c = 0; c < NV; ++c) ssum[(size_t)c * cap + i] = 0.0;
  }
  if (threadIdx.x == 0) {
    *s_count = 0;
    *s_cursor = 0;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned lt = mgb_lanemask_lt();
  const int max_fill = cap - cap / 4;
  const long long n = p.n;
  constexpr long long TILE = (long long)HASH_THREADS * HASH_R;
  const long long nfull = n / TILE;

  int hot_slot = -1;   // warp-uniform: the slot several lanes of this warp shared in an earlier row
  auto process_row = [&](const uint64_t* k, const uint64_t* x, long long row, bool live) {
    int slot = -1;
    if (live) {
      bool ok = k[0] != E && (NK == 1 || k[1] != E);
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const double xv = __longlong_as_double((long long)x[c]);
        ok = ok && (xv == xv);
      }
      if (ok) {
        uint32_t idx = (uint32_t)(mgb_slot_hash(mgb_hash_tuple<NK>((const int64_t*)k)) >> 40) & mask;
        // a frozen (3/4 full) table is probed 4 slots deep only: a key that sits further away is treated as
        // a miss, which is still correct (the row is aggregated by the global path and merged with the
        // staged partial of that key) and bounds the cost of a chunk whose keys are mostly distinct
        const int max_probe = *(volatile int*)s_count >= max_fill ? 4 : cap;
        for (int probe = 0; probe < max_probe; ++probe, idx = (idx + 1) & mask) {
          uint64_t cur = *(volatile uint64_t*)&sk[idx];
          if (cur == E) {
            if (*(volatile int*)s_count >= max_fill) break;   // frozen: no more inserts
            cur = atomicCAS((unsigned long long*)&sk[idx], (unsigned long long)E, (unsigned long long)k[0]);
            if (cur == E) {
              if (NK == 2) *(volatile uint64_t*)&sk[cap + idx] = k[1];
              atomicAdd(s_count, 1);
              slot = (int)idx;
              break;
            }
          }
          if (cur == k[0]) {
            if (NK == 1) {
              slot = (int)idx;
              break;
            }
            // second key of a freshly claimed slot may not be published yet: reads as EMPTY -> mismatch ->
            // the row claims another slot; duplicate entries of one key are benign (merged later)
            if (*(volatile uint64_t*)&sk[cap + idx] == k[1]) {
              slot = (int)idx;
              break;
            }
          }
        }
      }
    }
    // Warp pre-combine of a hot slot.  Shared-memory float64 adds are compare-and-swap loops; when several lanes of a
    // warp hit the same slot (skewed keys: Zipf's first key is 18 % of BASELINE config 5) they serialise on it.  The
    // rows of the slot most lanes share are therefore summed inside the warp -- __reduce_add_sync for the row count,
    // shuffles for the float64 columns -- and cost ONE atomic per accumulator.  The candidate is the slot that was
    // hot in the previous row of this warp, else the slot of the first lane that has one (re-elected until a hot one
    // sticks); uniform keys pay three votes per row for it.
    bool combined = false;
    {
      const unsigned have = __ballot_sync(0xffffffffu, slot >= 0);
      if (have) {
        int cand = hot_slot;
        unsigned same = __ballot_sync(0xffffffffu, slot >= 0 && slot == cand);
        if (__popc(same) < 3) {
          cand = __shfl_sync(0xffffffffu, slot, __ffs(have) - 1);
          same = __ballot_sync(0xffffffffu, slot >= 0 && slot == cand);
        }
        if (__popc(same) >= 3) {
          hot_slot = cand;
          const bool mem = (same >> lane) & 1u;
          unsigned cnt = 0;
          if (mem) cnt = __reduce_add_sync(same, 1u);
          double v[NV];
#pragma unroll
          for (int c = 0; c < NV; ++c) {
            v[c] = mem ? __longlong_as_double((long long)x[c]) : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
          }
          if (lane == __ffs(same) - 1) {
            atomicAdd(&srows[cand], cnt);
#pragma unroll
            for (int c = 0; c < NV; ++c) atomicAdd(&ssum[(size_t)c * cap + cand], v[c]);
          }
          combined = mem;
        }
      }
    }
    if (slot >= 0 && !combined) {
      atomicAdd(&srows[slot], 1u);
#pragma unroll
      for (int c = 0; c < NV; ++c) atomicAdd(&ssum[(size_t)c * cap + slot], __longlong_as_double((long long)x[c]));
    }
    const bool miss = live && slot < 0;
    const unsigned mm = __ballot_sync(0xffffffffu, miss);
    if (mm) {   // warp-aggregated append to the spill list
      unsigned long long pos = 0;
      if (lane == 0) pos = atomicAdd(&p.ctl->miss_count[0], (unsigned long long)__popc(mm));
      pos = __shfl_sync(0xffffffffu, pos, 0);
      if (miss) p.miss[pos + __popc(mm & lt)] = (int32_t)row;
    }
  };

  uint64_t kb[2][HASH_R][NK];
  uint64_t xb[2][HASH_R][NV];
  const uint64_t* pk[NK];
  const uint64_t* pc[NV];
#pragma unroll
  for (int j = 0; j < NK; ++j) pk[j] = p.keys[j] + threadIdx.x;
#pragma unroll
  for (int c = 0; c < NV; ++c) pc[c] = p.cols[c] + threadIdx.x;
  auto load_tile = [&](int buf, long long t) {
    const long long off = t * TILE;
#pragma unroll
    for (int j = 0; j < NK; ++j)
#pragma unroll
      for (int r = 0; r < HASH_R; ++r) kb[buf][r][j] = __ldcs((const unsigned long long*)(pk[j] + off) + r * HASH_THREADS);
#pragma unroll
    for (int c = 0; c < NV; ++c)
#pragma unroll
      for (int r = 0; r < HASH_R; ++r) xb[buf][r][c] = __ldcs((const unsigned long long*)(pc[c] + off) + r * HASH_THREADS);
  };
  auto process_tile = [&](int buf, long long t) {
    const long long base = t * TILE + threadIdx.x;
#pragma unroll
    for (int r = 0; r < HASH_R; ++r) process_row(kb[buf][r], xb[buf][r], base + (long long)r * HASH_THREADS, true);
  };
  long long t = blockIdx.x;
  if (t < nfull) load_tile(0, t);
  while (t < nfull) {
    const long long t1 = t + gridDim.x;
    if (t1 < nfull) load_tile(1, t1);
    process_tile(0, t);
    if (t1 >= nfull) break;
    const long long t2 = t1 + gridDim.x;
    if (t2 < nfull) load_tile(0, t2);
    process_tile(1, t1);
    t = t2;
  }
  if ((int)(nfull % gridDim.x) == (int)blockIdx.x) {   // ragged tail, whole warps so the ballots stay converged
    for (long long i0 = nfull * TILE + (threadIdx.x & ~31); i0 < n; i0 += HASH_THREADS) {
      const long long i = i0 + lane;
      const bool live = i < n;
      uint64_t k[NK], x[NV];
#pragma unroll
      for (int j = 0; j < NK; ++j) k[j] = live ? p.keys[j][i] : 0;
#pragma unroll
      for (int c = 0; c < NV; ++c) x[c] = live ? p.cols[c][i] : 0;
      process_row(k, x, i, live);
    }
  }
  // flush: one staged partial row per occupied slot
  __syncthreads();
  const int cnt = *s_count;
  if (cnt == 0) return;
  if (threadIdx.x == 0) {
    const unsigned long long pos = atomicAdd(&p.ctl->stg_count, (unsigned long long)cnt);
    if (pos + cnt > (unsigned long long)p.stg.cap) {
      p.ctl->err = 1;
      *s_pos = -1;
    } else {
      *s_pos = (long long)pos;
    }
  }
  __syncthreads();
  const long long pos0 = *s_pos;
  if (pos0 < 0) return;
  for (int i = threadIdx.x; i < cap; i += HASH_THREADS) {
    const uint64_t k0 = sk[i];
    if (k0 == E) continue;
    const long long row = pos0 + atomicAdd(s_cursor, 1);
    st_stage(p.stg, 0, row, k0);
    if (NK == 2) st_stage(p.stg, 1, row, sk[cap + i]);
    for (int a = 0; a < p.na; ++a) {
      const uint64_t v = p.acc_op[a] == MGB_COUNT ? (uint64_t)srows[i]
                                                   : (uint64_t)__double_as_longlong(ssum[(size_t)p.acc_col[a] * cap + i]);
      st_stage(p.stg, NK + a, row, v);
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// phase 1b: shared-memory hash table pre-aggregation with staging flush and spill list
// ----------------------------------------------------------------------------------------------------
struct SmemParams {
  Batch b;
  Layout L;
  int cap;            // table slots (power of two)
  int a0, a1;         // sorted accumulator range of this column group
  Ctl* ctl;
  Staging stg;
  long long inter_cap;    // staging rows available to intermediate flushes
  int32_t* miss;          // spill list of this column group
};

__device__ __forceinline__ void smem_atomic_apply(uint64_t* p, int kind, uint64_t c) {
  switch (kind) {
    case 0: atomicAdd((double*)p, __longlong_as_double((long long)c)); break;
    case 1: atomicAdd((unsigned long long*)p, (unsigned long long)c); break;
    case 2: atomicMin((unsigned long long*)p, (unsigned long long)c); break;
    default: atomicMax((unsigned long long*)p, (unsigned long long)c); break;
  }
}

// returns slot, or -1 (miss: table frozen/full or key carries the EMPTY marker)
template <int NK>
__device__ __forceinline__ int smem_find_or_insert(uint64_t* sk, int cap, const uint64_t* k, uint64_t h,
                                                   bool allow_insert, int* s_count, int max_fill) {
  const uint64_t E = (uint64_t)MGB_EMPTY_KEY;
  if (k[0] == E) return -1;
  if (NK == 2 && k[1] == E) return -1;
  const int mask = cap - 1;
  int idx = (int)((mgb_slot_hash(h) >> 24) & (uint64_t)mask);
  for (int probe = 0; probe < cap; ++probe, idx = (idx + 1) & mask) {
    uint64_t cur = *(volatile uint64_t*)&sk[idx];
    if (cur == E) {
      if (!allow_insert || *(volatile int*)s_count >= max_fill) return -1;
      cur = atomicCAS((unsigned long long*)&sk[idx], (unsigned long long)E, (unsigned long long)k[0]);
      if (cur == E) {
        if (NK == 2) *(volatile uint64_t*)&sk[cap + idx] = k[1];
        atomicAdd(s_count, 1);
        return idx;
      }
    }
    if (cur == k[0]) {
      if (NK == 1) return idx;
      // second key: a claimed slot whose second component is not published yet reads as EMPTY -> treated as
      // a mismatch; the row then claims another slot.  Duplicate entries of one key inside a CTA table are
      // benign (they become two partial rows that the global merge unifies).
      if (*(volatile uint64_t*)&sk[cap + idx] == k[1]) return idx;
    }
  }
  return -1;
}

template <int NK, int NV>
__global__ void __launch_bounds__(SMEM_THREADS, 1) k_preagg_smem(const SmemParams p) {
  extern __shared__ __align__(16) uint64_t smem[];
  const int cap = p.cap;
  const int nag = p.a1 - p.a0;
  uint64_t* sk = smem;                       // [NK][cap]
  uint64_t* sa = smem + (size_t)NK * cap;    // [nag][cap]
  int* s_int = (int*)(sa + (size_t)nag * cap);
  int* s_count = s_int;       // occupied entries
  int* s_frozen = s_int + 1;
  int* s_flag = s_int + 2;    // flush decision
  int* s_cursor = s_int + 3;  // staging sub-cursor during a flush
  long long* s_pos = (long long*)(s_int + 4);

  auto reset_table = [&]() {
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
      sk[i] = (uint64_t)MGB_EMPTY_KEY;
      if (NK == 2) sk[cap + i] = (uint64_t)MGB_EMPTY_KEY;
    }
    for (int a = 0; a < nag; ++a) {
      const uint64_t init = mgb_acc_init(p.L.acc[p.a0 + a].op);
      for (int i = threadIdx.x; i < cap; i += blockDim.x) sa[(size_t)a * cap + i] = init;
    }
    if (threadIdx.x == 0) *s_count = 0;
  };
  // writes the occupied entries as partial rows at stg[pos0 ...)
  auto flush_rows = [&](long long pos0) {
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
      const uint64_t k0 = sk[i];
      if (k0 == (uint64_t)MGB_EMPTY_KEY) continue;
      const long long row = pos0 + atomicAdd(s_cursor, 1);
      st_stage(p.stg, 0, row, k0);
      if (NK == 2) st_stage(p.stg, 1, row, sk[cap + i]);
      for (int j = 0; j < p.L.na; ++j) {  // sorted accumulator j -> spec column acc_index[j]
        const AccDesc d = p.L.acc[j];
        const uint64_t t = (j >= p.a0 && j < p.a1) ? sa[(size_t)(j - p.a0) * cap + i] : mgb_acc_init(d.op);
        st_stage(p.stg, p.L.nk + p.L.acc_index[j], row, mgb_acc_decode(d.op, d.dt, t));
      }
    }
  };

  reset_table();
  if (threadIdx.x == 0) {
    *s_frozen = 0;
    *s_flag = 0;
    *s_cursor = 0;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned lt = mgb_lanemask_lt();
  const long long n = p.b.n;
  const long long tile_rows = (long long)blockDim.x * SMEM_ROWS;
  const long long ntiles = (n + tile_rows - 1) / tile_rows;
  const int fill_limit = cap / 2;
  const int max_fill = cap - cap / 4;
  long long rows_since_reset = 0;
  int tiles_since_check = 0;
  bool frozen = false;

  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long base = t * tile_rows + threadIdx.x;
    uint64_t k[SMEM_ROWS][NK];
    uint64_t x[SMEM_ROWS][NV];
    bool valid[SMEM_ROWS];
#pragma unroll
    for (int r = 0; r < SMEM_ROWS; ++r) {
      const long long i = base + (long long)r * blockDim.x;
      valid[r] = i < n;
#pragma unroll
      for (int j = 0; j < NK; ++j) k[r][j] = valid[r] ? mgb_ld_stream(p.b.keys[j] + i) : 0;
#pragma unroll
      for (int c = 0; c < NV; ++c) x[r][c] = (valid[r] && c < p.b.ncol) ? mgb_ld_stream(p.b.cols[c] + i) : 0;
    }
#pragma unroll
    for (int r = 0; r < SMEM_ROWS; ++r) {
      int slot = -1;
      if (valid[r]) {
        const uint64_t h = mgb_hash_tuple<NK>((const int64_t*)k[r]);
        slot = smem_find_or_insert<NK>(sk, cap, k[r], h, !frozen, s_count, max_fill);
      }
      if (slot >= 0) {
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          if (c < p.b.ncol) {
            const int jb = p.L.col_begin[p.b.c0 + c], je = p.L.col_begin[p.b.c0 + c + 1];
            for (int j = jb; j < je; ++j) {
              const AccDesc d = p.L.acc[j];
              uint64_t contrib;
              if (mgb_contrib(d.op, d.dt, x[r][c], &contrib))
                smem_atomic_apply(&sa[(size_t)(j - p.a0) * cap + slot], d.kind, contrib);
            }
          }
        }
      }
      // spill list (warp-aggregated append)
      const bool miss = valid[r] && slot < 0;
      const unsigned mm = __ballot_sync(0xffffffffu, miss);
      if (mm) {
        unsigned long long pos = 0;
        if (lane == 0) pos = atomicAdd(&p.ctl->miss_count[p.b.group], (unsigned long long)__popc(mm));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (miss) p.miss[pos + __popc(mm & lt)] = (int32_t)(base + (long long)r * blockDim.x);
      }
    }
    rows_since_reset += tile_rows;
    if (++tiles_since_check >= SMEM_CHECK_EVERY) {
      tiles_since_check = 0;
      __syncthreads();
      const int cnt = *s_count;
      __syncthreads();  // everybody has read cnt before anybody inserts again (uniform branch below)
      if (!frozen && cnt >= fill_limit) {
        // flush + reset only while it pays off and staging room is left; otherwise freeze the table.
        if (threadIdx.x == 0) {
          int ok = 0;
          if (rows_since_reset >= (long long)FLUSH_GAIN * cnt) {
            unsigned long long old = *(volatile unsigned long long*)&p.ctl->stg_count;
            while (old + (unsigned long long)cnt <= (unsigned long long)p.inter_cap) {
              unsigned long long prev = atomicCAS(&p.ctl->stg_count, old, old + cnt);
              if (prev == old) {
                ok = 1;
                break;
              }
              old = prev;
            }
            *s_pos = (long long)old;
          }
          *s_flag = ok;
          *s_cursor = 0;
        }
        __syncthreads();
        if (*s_flag) {
          flush_rows(*s_pos);
          __syncthreads();
          reset_table();
          rows_since_reset = 0;
        } else {
          frozen = true;
        }
        __syncthreads();
      }
    }
  }
  // final flush: room is reserved by construction (stg.cap - inter_cap >= gridDim.x * cap)
  __syncthreads();
  const int cnt = *s_count;
  if (cnt > 0) {
    if (threadIdx.x == 0) {
      const unsigned long long pos = atomicAdd(&p.ctl->stg_count, (unsigned long long)cnt);
      if (pos + cnt > (unsigned long long)p.stg.cap) {
        p.ctl->err = 1;
        *s_flag = 0;
      } else {
        *s_flag = 1;
      }
      *s_pos = (long long)pos;
      *s_cursor = 0;
    }
    __syncthreads();
    if (*s_flag) flush_rows(*s_pos);
  }
}

// ----------------------------------------------------------------------------------------------------
// phase 2: global table updates
// ----------------------------------------------------------------------------------------------------
struct GlobalParams {
  Batch b;
  Layout L;
  int a0, a1;
  uint64_t* table;
  unsigned long long Cmask;
  int shift;               // 64 - log2(C)
  GlobalCtl* g;
  const int32_t* gather;   // optional row indices
  long long n_rows;        // rows to process (length of gather list, or b.n)
};

template <int NK, int NV>
__global__ void __launch_bounds__(GLOBAL_THREADS) k_update_global(const GlobalParams p) {
  int inserted = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < p.n_rows;
       q0 += stride * GLOBAL_ROWS) {
    uint64_t k[GLOBAL_ROWS][NK];
    uint64_t x[GLOBAL_ROWS][NV];
    bool valid[GLOBAL_ROWS];
#pragma unroll
    for (int r = 0; r < GLOBAL_ROWS; ++r) {
      const long long q = q0 + (long long)r * stride;
      valid[r] = q < p.n_rows;
      long long i = 0;
      if (valid[r]) i = p.gather ? (long long)p.gather[q] : q;
#pragma unroll
      for (int j = 0; j < NK; ++j) k[r][j] = valid[r] ? mgb_ld_stream(p.b.keys[j] + i) : 0;
#pragma unroll
      for (int c = 0; c < NV; ++c) x[r][c] = (valid[r] && c < p.b.ncol) ? mgb_ld_stream(p.b.cols[c] + i) : 0;
    }
#pragma unroll
    for (int r = 0; r < GLOBAL_ROWS; ++r) {
      if (!valid[r]) continue;
      const uint64_t h = mgb_hash_tuple<NK>((const int64_t*)k[r]);
      const long long slot = global_find_or_insert<NK>(p.table, p.L.W, p.Cmask, p.shift, k[r], h, p.g, &inserted);
      if (slot < 0) {
        p.g->err = 1;
        continue;
      }
      uint64_t* acc = p.table + (unsigned long long)slot * p.L.W + p.L.nk;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        if (c < p.b.ncol) {
          const int jb = p.L.col_begin[p.b.c0 + c], je = p.L.col_begin[p.b.c0 + c + 1];
          for (int j = jb; j < je; ++j) {
            const AccDesc d = p.L.acc[j];
            uint64_t contrib;
            if (mgb_contrib(d.op, d.dt, x[r][c], &contrib))
              mgb_atomic_apply_global(acc + p.L.acc_index[j], d.kind, contrib);
          }
        }
      }
    }
  }
  block_count_groups(inserted, p.g);
}

// Lean variant for the sum / mean / count shapes (every value column float64, accumulators SUM / COUNT):
// straight-line per column, RED.ADD.F64 for the sums, one RED.ADD.U64 per counted column; NaN values are
// skipped per column exactly like pandas' skipna sums / counts.
struct GlobalPlainParams {
  const uint64_t* keys[MGB_MAX_KEYS];
  const uint64_t* cols[8];
  int8_t sum_word[8];      // table word (after the keys) of column c's SUM accumulator, -1: none
  int8_t cnt_word[8][4];   // up to 4 COUNT accumulators may share a column (dict forms); -1 terminated
  int W;
  uint64_t* table;
  unsigned long long Cmask;
  int shift;
  GlobalCtl* g;
  const int32_t* gather;
  long long n_rows;
};

template <int NK, int NV>
__global__ void __launch_bounds__(GLOBAL_THREADS) k_update_global_plain(const GlobalPlainParams p) {
  int inserted = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < p.n_rows; q0 += stride * 4) {
    uint64_t k[4][NK];
    uint64_t x[4][NV];
    bool valid[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long q = q0 + (long long)r * stride;
      valid[r] = q < p.n_rows;
      long long i = 0;
      if (valid[r]) i = p.gather ? (long long)p.gather[q] : q;
#pragma unroll
      for (int j = 0; j < NK; ++j) k[r][j] = valid[r] ? mgb_ld_stream(p.keys[j] + i) : 0;
#pragma unroll
      for (int c = 0; c < NV; ++c) x[r][c] = valid[r] ? mgb_ld_stream(p.cols[c] + i) : 0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (!valid[r]) continue;
      const uint64_t h = mgb_hash_tuple<NK>((const int64_t*)k[r]);
      const long long slot = global_find_or_insert<NK>(p.table, p.W, p.Cmask, p.shift, k[r], h, p.g, &inserted);
      if (slot < 0) {
        p.g->err = 1;
        continue;
      }
      uint64_t* acc = p.table + (unsigned long long)slot * p.W + NK;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const double xv = __longlong_as_double((long long)x[r][c]);
        if (xv != xv) continue;
        if (p.sum_word[c] >= 0) atomicAdd((double*)(acc + p.sum_word[c]), xv);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (p.cnt_word[c][q] >= 0) atomicAdd((unsigned long long*)(acc + p.cnt_word[c][q]), 1ULL);
      }
    }
  }
  block_count_groups(inserted, p.g);
}

// staged partial rows (decoded values, one column per spec accumulator) -> global table
struct MergeParams {
  Staging stg;
  long long n_rows;
  Layout L;
  uint64_t* table;
  unsigned long long Cmask;
  int shift;
  GlobalCtl* g;
};

template <int NK>
__global__ void __launch_bounds__(GLOBAL_THREADS) k_merge_staged(const MergeParams p) {
  int inserted = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_rows;
       i += (long long)gridDim.x * blockDim.x) {
    uint64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = p.stg.base[(long long)j * p.stg.cap + i];
    const uint64_t h = mgb_hash_tuple<NK>((const int64_t*)k);
    const long long slot = global_find_or_insert<NK>(p.table, p.L.W, p.Cmask, p.shift, k, h, p.g, &inserted);
    if (slot < 0) {
      p.g->err = 1;
      continue;
    }
    uint64_t* acc = p.table + (unsigned long long)slot * p.L.W + p.L.nk;
    for (int j = 0; j < p.L.na; ++j) {
      const AccDesc d = p.L.acc[j];
      const int a = p.L.acc_index[j];
      const uint64_t v = p.stg.base[(long long)(p.L.nk + a) * p.stg.cap + i];
      // merge step of a staged row of THIS aggregator (phase 1 -> phase 2): SUM/SUMSQ/COUNT add -- a NaN sum
      // (inf - inf) stays NaN, as in pandas' group_sum over the chunk; MIN/MAX re-encode (NaN == "no value")
      uint64_t contrib;
      bool ok = true;
      if (d.kind == 0) {
        contrib = v;
      } else if (d.kind == 1) {
        contrib = v;
        ok = v != 0;
      } else if (d.dt == MGB_F64) {
        ok = !mgb_is_nan_bits(v);
        contrib = mgb_enc_f64(v);
      } else {
        contrib = mgb_enc_i64(v);
      }
      if (ok) mgb_atomic_apply_global(acc + a, d.kind, contrib);
    }
  }
  block_count_groups(inserted, p.g);
}

// ----------------------------------------------------------------------------------------------------
// emit: compaction of the global table
// ----------------------------------------------------------------------------------------------------
#define EMIT_THREADS 256

template <int NK>
__device__ __forceinline__ bool slot_occupied(const uint64_t* table, int W, long long slot, long long C,
                                              const GlobalCtl* g) {
  if (slot == C) return g->special_used != 0;
  const uint64_t* s = table + (unsigned long long)slot * W;
  if (NK == 1) return s[0] != (uint64_t)MGB_EMPTY_KEY;
  return !(s[0] == (uint64_t)MGB_EMPTY_KEY && s[1] == (uint64_t)MGB_EMPTY_KEY);
}

template <int NK>
__global__ void __launch_bounds__(EMIT_THREADS) k_emit_count(const uint64_t* table, int W, long long C,
                                                             const GlobalCtl* g, int64_t* tile_counts) {
  const long long slot = (long long)blockIdx.x * EMIT_THREADS + threadIdx.x;
  const bool occ = slot <= C && slot_occupied<NK>(table, W, slot, C, g);
  const int c = __syncthreads_count(occ);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = c;
}

struct EmitCols {
  uint64_t* key[MGB_MAX_KEYS];
  uint64_t* acc[MGB_MAX_ACCS];
};

template <int NK>
__global__ void __launch_bounds__(EMIT_THREADS) k_emit_write(const uint64_t* table, Layout L, long long C,
                                                             const GlobalCtl* g, const int64_t* tile_offsets,
                                                             EmitCols out) {
  __shared__ int warp_cnt[EMIT_THREADS / 32];
  const long long slot = (long long)blockIdx.x * EMIT_THREADS + threadIdx.x;
  const bool occ = slot <= C && slot_occupied<NK>(table, L.W, slot, C, g);
  const unsigned m = __ballot_sync(0xffffffffu, occ);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[warp] = __popc(m);
  __syncthreads();
  if (!occ) return;
  int off = __popc(m & mgb_lanemask_lt());
  for (int w = 0; w < warp; ++w) off += warp_cnt[w];
  const long long row = tile_offsets[blockIdx.x] + off;
  const uint64_t* s = table + (unsigned long long)slot * L.W;
#pragma unroll
  for (int j = 0; j < NK; ++j) out.key[j][row] = s[j];
  for (int j = 0; j < L.na; ++j) {
    const AccDesc d = L.acc[j];
    const int a = L.acc_index[j];
    out.acc[a][row] = mgb_acc_decode(d.op, d.dt, s[L.nk + a]);
  }
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
struct UpdateRec {
  Batch batch;                    // keys / n (cols filled per group at launch time)
  const uint64_t* vals[MGB_MAX_VALS];
  Ctl* ctl;                       // device
  Staging stg;                    // staged partial rows of phase 1
  long long inter_cap;
  int32_t* miss;                  // [n_col_groups][n]
  bool direct;                    // no pre-aggregation pass: every row goes to the global table at finalize
};

struct mgb_agg {
  mgb_agg_spec spec;
  Layout L;
  cudaStream_t st;
  int n_col_groups;
  int group_a0[MGB_MAX_VALS / GROUP_COLS + 1];  // sorted accumulator range per column group
  int sms;
  int* sticky_fail;       // device
  GlobalCtl* gctl;        // device
  uint64_t* table;
  long long C;
  int shift;
  bool finalized;
  bool direct_hint;       // caller expects (nearly) distinct keys: skip the shared-memory pre-aggregation
  bool plain;             // sum / mean / count shape: lean kernels apply
  long long n_groups;
  std::vector<UpdateRec> updates;
  long long stats[6];
  bool bucket_mode;       // finalize took the bucketed path (mgb_bucket.cu): the result lives in `bkt`
  BktResult bkt;
};

static int build_layout(const mgb_agg_spec* s, Layout* L) {
  MGB_REQUIRE(s->n_keys >= 1 && s->n_keys <= MGB_MAX_KEYS, MGB_ERR_UNSUPPORTED, "n_keys=%d outside [1,%d]",
              s->n_keys, MGB_MAX_KEYS);
  MGB_REQUIRE(s->n_vals >= 1 && s->n_vals <= MGB_MAX_VALS, MGB_ERR_UNSUPPORTED, "n_vals=%d outside [1,%d]",
              s->n_vals, MGB_MAX_VALS);
  MGB_REQUIRE(s->n_accs >= 1 && s->n_accs <= MGB_MAX_ACCS, MGB_ERR_UNSUPPORTED, "n_accs=%d outside [1,%d]",
              s->n_accs, MGB_MAX_ACCS);
  memset(L, 0, sizeof(*L));
  L->nk = s->n_keys;
  L->nv = s->n_vals;
  L->na = s->n_accs;
  for (int c = 0; c < s->n_vals; ++c)
    MGB_REQUIRE(s->val_dtype[c] == MGB_F64 || s->val_dtype[c] == MGB_I64, MGB_ERR_INVALID, "bad dtype of column %d", c);
  int j = 0;
  for (int c = 0; c < s->n_vals; ++c) {
    L->col_begin[c] = (uint8_t)j;
    for (int a = 0; a < s->n_accs; ++a) {
      if (s->acc_col[a] != c) continue;
      MGB_REQUIRE(s->acc_op[a] >= MGB_SUM && s->acc_op[a] <= MGB_MAX, MGB_ERR_INVALID, "bad op of accumulator %d", a);
      L->acc[j].col = (uint8_t)c;
      L->acc[j].op = (uint8_t)s->acc_op[a];
      L->acc[j].dt = (uint8_t)s->val_dtype[c];
      L->acc[j].kind = (uint8_t)mgb_acc_kind(s->acc_op[a], s->val_dtype[c]);
      L->acc_index[j] = (uint8_t)a;
      L->spec_op[a] = (uint8_t)s->acc_op[a];
      L->spec_dt[a] = (uint8_t)s->val_dtype[c];
      L->spec_sorted[a] = (uint8_t)j;
      ++j;
    }
  }
  for (int c = s->n_vals; c <= MGB_MAX_VALS; ++c) L->col_begin[c] = (uint8_t)j;
  MGB_REQUIRE(j == s->n_accs, MGB_ERR_INVALID, "accumulator refers to a column outside [0,%d)", s->n_vals);
  int W = s->n_keys + s->n_accs;
  W = W <= 4 ? 4 : ((W + 1) & ~1);
  L->W = W;
  return MGB_OK;
}

extern "C" int mgb_agg_create(mgb_agg** out, const mgb_agg_spec* spec, void* stream) {
  MGB_REQUIRE(out && spec, MGB_ERR_INVALID, "null pointer");
  Layout L;
  MGB_TRY(build_layout(spec, &L));
  mgb_agg* a = new (std::nothrow) mgb_agg();
  MGB_REQUIRE(a, MGB_ERR_INVALID, "out of host memory");
  a->spec = *spec;
  a->L = L;
  a->st = (cudaStream_t)stream;
  a->sms = mgb_sm_count();
  a->n_col_groups = (spec->n_vals + GROUP_COLS - 1) / GROUP_COLS;
  for (int g = 0; g <= a->n_col_groups; ++g) {
    int c = g * GROUP_COLS;
    if (c > spec->n_vals) c = spec->n_vals;
    a->group_a0[g] = L.col_begin[c];
  }
  a->table = nullptr;
  a->C = 0;
  a->direct_hint = false;
  a->plain = a->n_col_groups == 1 && spec->n_vals <= 8;
  for (int c = 0; c < spec->n_vals && a->plain; ++c) a->plain = spec->val_dtype[c] == MGB_F64;
  for (int i = 0; i < spec->n_accs && a->plain; ++i) {
    a->plain = spec->acc_op[i] == MGB_SUM || spec->acc_op[i] == MGB_COUNT;
  }
  if (a->plain) {   // at most one SUM and four COUNT accumulators per column fit the lean kernels' tables
    for (int c = 0; c < spec->n_vals && a->plain; ++c) {
      int ns = 0, nc = 0;
      for (int i = 0; i < spec->n_accs; ++i)
        if (spec->acc_col[i] == c) (spec->acc_op[i] == MGB_SUM ? ns : nc)++;
      a->plain = ns <= 1 && nc <= 4;
    }
  }
  a->finalized = false;
  a->n_groups = 0;
  a->bucket_mode = false;
  memset(&a->bkt, 0, sizeof(a->bkt));
  memset(a->stats, 0, sizeof(a->stats));
  char* blk = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&blk, 256, a->st);
  if (e != cudaSuccess) {
    delete a;
    mgb_set_error("cudaMallocAsync: %s", cudaGetErrorString(e));
    return MGB_ERR_CUDA;
  }
  cudaMemsetAsync(blk, 0, 256, a->st);
  a->gctl = (GlobalCtl*)blk;
  a->sticky_fail = (int*)(blk + 128);
  *out = a;
  return MGB_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    MGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return MGB_OK;
}

#define DISPATCH_NK_NV(NKv, NVv, CALL)                  \
  do {                                                  \
    if ((NKv) == 1) {                                   \
      if ((NVv) <= 1) { CALL(1, 1); }                   \
      else if ((NVv) <= 2) { CALL(1, 2); }              \
      else if ((NVv) <= 4) { CALL(1, 4); }              \
      else { CALL(1, 8); }                              \
    } else {                                            \
      if ((NVv) <= 1) { CALL(2, 1); }                   \
      else if ((NVv) <= 2) { CALL(2, 2); }              \
      else if ((NVv) <= 4) { CALL(2, 4); }              \
      else { CALL(2, 8); }                              \
    }                                                   \
  } while (0)

static void fill_group(const mgb_agg* a, const UpdateRec& u, int g, Batch* b) {
  *b = u.batch;
  b->group = g;
  b->c0 = g * GROUP_COLS;
  b->ncol = a->spec.n_vals - b->c0 < GROUP_COLS ? a->spec.n_vals - b->c0 : GROUP_COLS;
  for (int c = 0; c < GROUP_COLS; ++c) b->cols[c] = c < b->ncol ? u.vals[b->c0 + c] : nullptr;
}

extern "C" int mgb_agg_update(mgb_agg* a, const int64_t* const* key_cols, const void* const* val_cols,
                              int64_t n_rows) {
  MGB_REQUIRE(a, MGB_ERR_INVALID, "null aggregator");
  MGB_REQUIRE(!a->finalized, MGB_ERR_STATE, "update after finalize");
  MGB_REQUIRE(n_rows >= 0 && n_rows < 0x7fffffffLL, MGB_ERR_UNSUPPORTED,
              "n_rows=%lld per update must be < 2^31 (split the batch)", (long long)n_rows);
  if (n_rows == 0) return MGB_OK;
  MGB_REQUIRE(key_cols && val_cols, MGB_ERR_INVALID, "null pointer");
  const Layout& L = a->L;
  cudaStream_t st = a->st;

  UpdateRec u;
  memset(&u, 0, sizeof(u));
  for (int j = 0; j < L.nk; ++j) {
    MGB_REQUIRE(key_cols[j], MGB_ERR_INVALID, "null key column %d", j);
    u.batch.keys[j] = (const uint64_t*)key_cols[j];
  }
  for (int c = 0; c < L.nv; ++c) {
    MGB_REQUIRE(val_cols[c], MGB_ERR_INVALID, "null value column %d", c);
    u.vals[c] = (const uint64_t*)val_cols[c];
  }
  u.batch.n = n_rows;

  // shared-memory table capacity for the widest column group
  int max_nag = 0;
  for (int g = 0; g < a->n_col_groups; ++g) {
    int nag = a->group_a0[g + 1] - a->group_a0[g];
    if (nag > max_nag) max_nag = nag;
  }
  int cap = 4096;
  while (cap > 64 && (size_t)cap * (L.nk + max_nag) * 8 > SMEM_BUDGET) cap >>= 1;
  const long long tile_rows = (long long)SMEM_THREADS * SMEM_ROWS;
  long long ntiles = (n_rows + tile_rows - 1) / tile_rows;
  int smem_grid = (int)(ntiles < a->sms ? ntiles : a->sms);

  if (a->direct_hint) {   // nothing to launch now: the whole batch is inserted at finalize
    u.direct = true;
    a->updates.push_back(u);
    return MGB_OK;
  }
  // PLAIN shape (every value column float64, only SUM / COUNT, one column group): lean hash kernel
  const bool plain = a->plain;
  int hcap = 4096;
  while (hcap > 64 && (size_t)hcap * ((L.nk + L.nv) * 8 + 4) + 64 > SMEM_BUDGET) hcap >>= 1;
  const long long htiles = (n_rows + 1023) / 1024;
  const int hash_grid = (int)(htiles < a->sms ? htiles : a->sms);
  if (plain) {
    cap = hcap;
    smem_grid = hash_grid;
  }

  // allocations: ctl | staging | spill lists
  MGB_CUDA(cudaMallocAsync((void**)&u.ctl, sizeof(Ctl), st));
  MGB_CUDA(cudaMemsetAsync(u.ctl, 0, sizeof(Ctl), st));
  const int ncolstg = L.nk + L.na;
  u.inter_cap = plain ? 0 : n_rows / FLUSH_GAIN;
  if (u.inter_cap > MAX_INTER_ROWS) u.inter_cap = MAX_INTER_ROWS;
  u.stg.cap = u.inter_cap + (long long)smem_grid * cap * a->n_col_groups;
  MGB_CUDA(cudaMallocAsync((void**)&u.stg.base, (size_t)u.stg.cap * ncolstg * 8, st));
  MGB_CUDA(cudaMallocAsync((void**)&u.miss, sizeof(int32_t) * (size_t)n_rows * a->n_col_groups, st));

  if (plain) {
    HashParams p;
    memset(&p, 0, sizeof(p));
    for (int j = 0; j < L.nk; ++j) p.keys[j] = u.batch.keys[j];
    for (int c = 0; c < L.nv; ++c) p.cols[c] = u.vals[c];
    p.n = n_rows;
    p.cap = hcap;
    p.na = L.na;
    for (int i = 0; i < L.na; ++i) {
      p.acc_col[i] = (uint8_t)a->spec.acc_col[i];
      p.acc_op[i] = (uint8_t)a->spec.acc_op[i];
    }
    p.ctl = u.ctl;
    p.stg = u.stg;
    p.miss = u.miss;
    const size_t sm = (size_t)hcap * ((L.nk + L.nv) * 8 + 4) + 64;
    MgbProfScope prof(MGB_K_PREAGG_SMEM, st);
#define CALLH(NK, NV)                                          \
  MGB_TRY(set_smem(k_preagg_hash<NK, NV>, sm));                \
  k_preagg_hash<NK, NV><<<hash_grid, HASH_THREADS, sm, st>>>(p);
#define PICKH(NK)                                  \
  switch (L.nv) {                                  \
    case 1: { CALLH(NK, 1) } break;                \
    case 2: { CALLH(NK, 2) } break;                \
    case 3: { CALLH(NK, 3) } break;                \
    case 4: { CALLH(NK, 4) } break;                \
    case 5: { CALLH(NK, 5) } break;                \
    case 6: { CALLH(NK, 6) } break;                \
    case 7: { CALLH(NK, 7) } break;                \
    default: { CALLH(NK, 8) } break;               \
  }
    if (L.nk == 1) { PICKH(1) } else { PICKH(2) }
#undef PICKH
#undef CALLH
    MGB_CHECK_LAUNCH();
    a->stats[5]++;
    a->updates.push_back(u);
    return MGB_OK;
  }
  for (int g = 0; g < a->n_col_groups; ++g) {
    SmemParams p;
    fill_group(a, u, g, &p.b);
    p.L = L;
    p.cap = cap;
    p.a0 = a->group_a0[g];
    p.a1 = a->group_a0[g + 1];
    p.ctl = u.ctl;
    p.stg = u.stg;
    p.inter_cap = u.inter_cap;
    p.miss = u.miss + (size_t)g * n_rows;
    size_t sm = (size_t)cap * (L.nk + (p.a1 - p.a0)) * 8 + 64;
    MgbProfScope prof(MGB_K_PREAGG_SMEM, st);
#define CALL(NK, NV)                                                               \
  MGB_TRY(set_smem(k_preagg_smem<NK, NV>, sm));                                    \
  k_preagg_smem<NK, NV><<<smem_grid, SMEM_THREADS, sm, st>>>(p);
    DISPATCH_NK_NV(L.nk, p.b.ncol, CALL);
#undef CALL
    MGB_CHECK_LAUNCH();
    a->stats[5]++;
  }
  a->updates.push_back(u);
  return MGB_OK;
}

static unsigned grid_rows(const mgb_agg* a, long long rows, int threads, int rows_per_thread) {
  long long b = (rows + (long long)threads * rows_per_thread - 1) / ((long long)threads * rows_per_thread);
  long long cap = (long long)a->sms * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

extern "C" int mgb_agg_finalize(mgb_agg* a, int64_t* out_n_groups) {
  MGB_REQUIRE(a && out_n_groups, MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(!a->finalized, MGB_ERR_STATE, "finalize called twice");
  cudaStream_t st = a->st;
  const Layout& L = a->L;
  const size_t U = a->updates.size();
  // 0. inputs that were not pre-aggregated (distinct-keys hint): bucketed aggregation -- no table sized by the
  //    row count, no atomics on the values (mgb_bucket.cu).  Overflow of a bucket falls through to the table.
  if (U > 0 && U < 65536 && mgb_bucket_supports(L)) {
    bool all_direct = true;
    long long T = 0;
    for (size_t i = 0; i < U; ++i) {
      all_direct = all_direct && a->updates[i].direct;
      T += a->updates[i].batch.n;
    }
    if (all_direct && T >= BUCKET_MIN_ROWS && T < 0x7fffffffLL) {
      std::vector<BktSource> srcs(U);
      long long begin = 0;
      for (size_t i = 0; i < U; ++i) {
        const UpdateRec& u = a->updates[i];
        memset(&srcs[i], 0, sizeof(BktSource));
        for (int j = 0; j < L.nk; ++j) srcs[i].keys[j] = u.batch.keys[j];
        for (int c = 0; c < L.nv; ++c) srcs[i].vals[c] = u.vals[c];
        srcs[i].begin = begin;
        begin += u.batch.n;
      }
      int overflow = 0;
      MGB_TRY(mgb_bucket_aggregate(L, srcs.data(), (int)U, T, st, &a->bkt, &overflow));
      if (!overflow) {
        a->bucket_mode = true;
        a->n_groups = a->bkt.G;
        a->stats[2] += T;
        a->stats[5] += 8;
        a->finalized = true;
        *out_n_groups = a->n_groups;
        return MGB_OK;
      }
      a->stats[0] = overflow;   // why the bucketed path declined (1 = table overflow, 2 = skewed keys)
    }
  }
  // 1. read back the phase-1 counters (the only mid-flight synchronisation)
  std::vector<Ctl> ctl(U);
  bool any_pre = false;
  for (size_t i = 0; i < U; ++i) {
    memset(&ctl[i], 0, sizeof(Ctl));
    if (a->updates[i].direct) continue;
    any_pre = true;
    MGB_CUDA(cudaMemcpyAsync(&ctl[i], a->updates[i].ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
  }
  if (any_pre) MGB_CUDA(cudaStreamSynchronize(st));
  long long total = 0;
  for (size_t i = 0; i < U; ++i) {
    const UpdateRec& u = a->updates[i];
    if (u.direct) {
      total += u.batch.n;
      a->stats[2] += u.batch.n;
      continue;
    }
    MGB_REQUIRE(ctl[i].err == 0, MGB_ERR_OVERFLOW, "staging overflow in update %zu", i);
    total += (long long)ctl[i].stg_count;
    long long missed = 0;
    for (int g = 0; g < a->n_col_groups; ++g) {
      total += (long long)ctl[i].miss_count[g];
      if ((long long)ctl[i].miss_count[g] > missed) missed = (long long)ctl[i].miss_count[g];
    }
    a->stats[1] += u.batch.n - missed;
    a->stats[2] += missed;
    a->stats[3] += (long long)ctl[i].stg_count;
  }
  // 2. global table sized from the exact row bound
  long long C = 1024;
  while (C < 2 * total) C <<= 1;
  int logC = 0;
  while ((1LL << logC) < C) ++logC;
  a->C = C;
  a->shift = 64 - logC;
  a->stats[4] = C;
  MGB_CUDA(cudaMallocAsync((void**)&a->table, (size_t)(C + 1) * L.W * 8, st));
  {
    long long words = (C + 1) * L.W;
    long long b = (words + 255) / 256, capb = (long long)a->sms * 16;
    MgbProfScope prof(MGB_K_TABLE_INIT, st);
    k_table_init<<<(unsigned)(b > capb ? capb : b), 256, 0, st>>>(a->table, L, C + 1);
    MGB_CHECK_LAUNCH();
    a->stats[5]++;
  }
  // 3. merges
  GlobalPlainParams gp;
  memset(&gp, 0, sizeof(gp));
  if (a->plain) {
    for (int c = 0; c < 8; ++c) {
      gp.sum_word[c] = -1;
      for (int q = 0; q < 4; ++q) gp.cnt_word[c][q] = -1;
    }
    for (int i = 0; i < L.na; ++i) {   // table word order == spec accumulator order
      const int c = a->spec.acc_col[i];
      if (a->spec.acc_op[i] == MGB_SUM) {
        gp.sum_word[c] = (int8_t)i;
      } else {
        for (int q = 0; q < 4; ++q)
          if (gp.cnt_word[c][q] < 0) {
            gp.cnt_word[c][q] = (int8_t)i;
            break;
          }
      }
    }
    gp.W = L.W;
    gp.table = a->table;
    gp.Cmask = (unsigned long long)(C - 1);
    gp.shift = a->shift;
    gp.g = a->gctl;
  }
  auto launch_plain = [&](const UpdateRec& u, const int32_t* gather, long long nrows) -> int {
    GlobalPlainParams q = gp;
    for (int j = 0; j < L.nk; ++j) q.keys[j] = u.batch.keys[j];
    for (int c = 0; c < L.nv; ++c) q.cols[c] = u.vals[c];
    q.gather = gather;
    q.n_rows = nrows;
    const unsigned grid = grid_rows(a, nrows, GLOBAL_THREADS, 4);
    MgbProfScope prof(MGB_K_UPDATE_GLOBAL, st);
#define CALLG(NK, NV) k_update_global_plain<NK, NV><<<grid, GLOBAL_THREADS, 0, st>>>(q);
#define PICKG(NK)                                  \
  switch (L.nv) {                                  \
    case 1: { CALLG(NK, 1) } break;                \
    case 2: { CALLG(NK, 2) } break;                \
    case 3: { CALLG(NK, 3) } break;                \
    case 4: { CALLG(NK, 4) } break;                \
    case 5: { CALLG(NK, 5) } break;                \
    case 6: { CALLG(NK, 6) } break;                \
    case 7: { CALLG(NK, 7) } break;                \
    default: { CALLG(NK, 8) } break;               \
  }
    if (L.nk == 1) { PICKG(1) } else { PICKG(2) }
#undef PICKG
#undef CALLG
    MGB_CHECK_LAUNCH();
    a->stats[5]++;
    return MGB_OK;
  };
  for (size_t i = 0; i < U; ++i) {
    const UpdateRec& u = a->updates[i];
    if (u.direct && a->plain) {
      MGB_TRY(launch_plain(u, nullptr, u.batch.n));
      continue;
    }
    if (u.direct) {   // generic shapes: the descriptor-driven kernel over the whole batch, per column group
      for (int g = 0; g < a->n_col_groups; ++g) {
        GlobalParams p;
        fill_group(a, u, g, &p.b);
        p.L = L;
        p.a0 = a->group_a0[g];
        p.a1 = a->group_a0[g + 1];
        p.table = a->table;
        p.Cmask = (unsigned long long)(C - 1);
        p.shift = a->shift;
        p.g = a->gctl;
        p.gather = nullptr;
        p.n_rows = u.batch.n;
        unsigned grid = grid_rows(a, u.batch.n, GLOBAL_THREADS, GLOBAL_ROWS);
        MgbProfScope prof(MGB_K_UPDATE_GLOBAL, st);
#define CALL(NK, NV) k_update_global<NK, NV><<<grid, GLOBAL_THREADS, 0, st>>>(p);
        DISPATCH_NK_NV(L.nk, p.b.ncol, CALL);
#undef CALL
        MGB_CHECK_LAUNCH();
        a->stats[5]++;
      }
      continue;
    }
    MergeParams m;
    m.L = L;
    m.table = a->table;
    m.Cmask = (unsigned long long)(C - 1);
    m.shift = a->shift;
    m.g = a->gctl;
    m.stg = u.stg;
    m.n_rows = (long long)ctl[i].stg_count;
    if (m.n_rows > 0) {
      unsigned g = grid_rows(a, m.n_rows, GLOBAL_THREADS, 1);
      MgbProfScope prof(MGB_K_MERGE, st);
      if (L.nk == 1)
        k_merge_staged<1><<<g, GLOBAL_THREADS, 0, st>>>(m);
      else
        k_merge_staged<2><<<g, GLOBAL_THREADS, 0, st>>>(m);
      MGB_CHECK_LAUNCH();
      a->stats[5]++;
    }
    if (a->plain) {
      const long long nm = (long long)ctl[i].miss_count[0];
      if (nm > 0) MGB_TRY(launch_plain(u, u.miss, nm));
      continue;
    }
    for (int g = 0; g < a->n_col_groups; ++g) {
      const long long nm = (long long)ctl[i].miss_count[g];
      if (nm == 0) continue;
      GlobalParams p;
      fill_group(a, u, g, &p.b);
      p.L = L;
      p.a0 = a->group_a0[g];
      p.a1 = a->group_a0[g + 1];
      p.table = a->table;
      p.Cmask = (unsigned long long)(C - 1);
      p.shift = a->shift;
      p.g = a->gctl;
      p.gather = u.miss + (size_t)g * u.batch.n;
      p.n_rows = nm;
      unsigned grid = grid_rows(a, nm, GLOBAL_THREADS, GLOBAL_ROWS);
      MgbProfScope prof(MGB_K_UPDATE_GLOBAL, st);
#define CALL(NK, NV) k_update_global<NK, NV><<<grid, GLOBAL_THREADS, 0, st>>>(p);
      DISPATCH_NK_NV(L.nk, p.b.ncol, CALL);
#undef CALL
      MGB_CHECK_LAUNCH();
      a->stats[5]++;
    }
  }
  // 4. group count
  GlobalCtl gh;
  MGB_CUDA(cudaMemcpyAsync(&gh, a->gctl, sizeof(GlobalCtl), cudaMemcpyDeviceToHost, st));
  MGB_CUDA(cudaStreamSynchronize(st));
  MGB_REQUIRE(gh.err == 0, MGB_ERR_OVERFLOW, "global hash table overflow (capacity %lld)", C);
  // phase-1 scratch is no longer needed
  for (size_t i = 0; i < U; ++i) {
    UpdateRec& u = a->updates[i];
    if (u.ctl) cudaFreeAsync(u.ctl, st);
    if (u.stg.base) cudaFreeAsync(u.stg.base, st);
    if (u.miss) cudaFreeAsync(u.miss, st);
    u.ctl = nullptr;
    u.stg.base = nullptr;
    u.miss = nullptr;
  }
  a->n_groups = (long long)gh.n_groups;
  a->finalized = true;
  *out_n_groups = a->n_groups;
  return MGB_OK;
}

extern "C" int mgb_agg_set_hint(mgb_agg* a, int32_t expect_distinct_keys) {
  MGB_REQUIRE(a, MGB_ERR_INVALID, "null aggregator");
  MGB_REQUIRE(a->updates.empty(), MGB_ERR_STATE, "hint must be set before the first update");
  a->direct_hint = expect_distinct_keys != 0;
  return MGB_OK;
}

extern "C" int mgb_agg_emit(mgb_agg* a, int64_t* const* out_key_cols, void* const* out_acc_cols,
                            int32_t sort_keys) {
  MGB_REQUIRE(a, MGB_ERR_INVALID, "null aggregator");
  MGB_REQUIRE(a->finalized, MGB_ERR_STATE, "emit before finalize");
  if (a->n_groups == 0) return MGB_OK;
  MGB_REQUIRE(out_key_cols && out_acc_cols, MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(!sort_keys || a->n_groups < 0x7fffffffLL, MGB_ERR_UNSUPPORTED, "sorted emit needs < 2^31 groups");
  cudaStream_t st = a->st;
  const Layout& L = a->L;
  const long long G = a->n_groups;
  const int ncol = L.nk + L.na;
  EmitCols out;
  memset(&out, 0, sizeof(out));
  uint64_t* tmp = nullptr;
  if (sort_keys) {
    MGB_CUDA(cudaMallocAsync((void**)&tmp, (size_t)G * ncol * 8, st));
    for (int j = 0; j < L.nk; ++j) out.key[j] = tmp + (size_t)j * G;
    for (int j = 0; j < L.na; ++j) out.acc[j] = tmp + (size_t)(L.nk + j) * G;
  } else {
    for (int j = 0; j < L.nk; ++j) {
      MGB_REQUIRE(out_key_cols[j], MGB_ERR_INVALID, "null output key column %d", j);
      out.key[j] = (uint64_t*)out_key_cols[j];
    }
    for (int j = 0; j < L.na; ++j) {
      MGB_REQUIRE(out_acc_cols[j], MGB_ERR_INVALID, "null output accumulator column %d", j);
      out.acc[j] = (uint64_t*)out_acc_cols[j];
    }
  }
  const long long slots = a->C + 1;
  const long long tiles = (slots + EMIT_THREADS - 1) / EMIT_THREADS;
  MgbProfScope prof(MGB_K_EMIT, st);
  int64_t* tile_counts = nullptr;
  int status = MGB_OK;
  if (a->bucket_mode) {
    if (!a->bkt.scratch) {   // the bucketed result is released by its first emit
      if (tmp) cudaFreeAsync(tmp, st);
      mgb_set_error("emit called twice on a bucketed aggregation result");
      return MGB_ERR_STATE;
    }
    status = mgb_bucket_emit(L, a->bkt, out.key, out.acc, st);
    mgb_bucket_free(&a->bkt, st);
  } else {
    MGB_CUDA(cudaMallocAsync((void**)&tile_counts, sizeof(int64_t) * tiles, st));
    if (L.nk == 1)
      k_emit_count<1><<<(unsigned)tiles, EMIT_THREADS, 0, st>>>(a->table, L.W, a->C, a->gctl, tile_counts);
    else
      k_emit_count<2><<<(unsigned)tiles, EMIT_THREADS, 0, st>>>(a->table, L.W, a->C, a->gctl, tile_counts);
    MGB_CHECK_LAUNCH();
    status = mgb_scan_exclusive_i64(tile_counts, tiles, nullptr, st);
  }
  if (status == MGB_OK && !a->bucket_mode) {
    if (L.nk == 1)
      k_emit_write<1><<<(unsigned)tiles, EMIT_THREADS, 0, st>>>(a->table, L, a->C, a->gctl, tile_counts, out);
    else
      k_emit_write<2><<<(unsigned)tiles, EMIT_THREADS, 0, st>>>(a->table, L, a->C, a->gctl, tile_counts, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      mgb_set_error("emit launch: %s", cudaGetErrorString(e));
      status = MGB_ERR_CUDA;
    }
  }
  a->stats[5] += 5;
  if (tile_counts) cudaFreeAsync(tile_counts, st);
  if (status == MGB_OK && sort_keys) {
    int32_t* perm = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&perm, sizeof(int32_t) * G, st);
    if (e != cudaSuccess) {
      mgb_set_error("cudaMallocAsync: %s", cudaGetErrorString(e));
      status = MGB_ERR_CUDA;
    } else {
      const int64_t* kc[MGB_MAX_KEYS];
      for (int j = 0; j < L.nk; ++j) kc[j] = (const int64_t*)out.key[j];
      status = mgb_radix_sort_perm_keys(kc, L.nk, G, perm, st);
      if (status == MGB_OK) {
        const void* in_cols[MGB_MAX_KEYS + MGB_MAX_ACCS];
        void* out_cols[MGB_MAX_KEYS + MGB_MAX_ACCS];
        for (int j = 0; j < L.nk; ++j) {
          in_cols[j] = out.key[j];
          out_cols[j] = out_key_cols[j];
        }
        for (int j = 0; j < L.na; ++j) {
          in_cols[L.nk + j] = out.acc[j];
          out_cols[L.nk + j] = out_acc_cols[j];
        }
        status = mgb_gather_rows(in_cols, out_cols, ncol, perm, G, st);
      }
      cudaFreeAsync(perm, st);
    }
  }
  if (tmp) cudaFreeAsync(tmp, st);
  return status;
}

extern "C" int mgb_agg_stats(mgb_agg* a, int64_t* stats6) {
  MGB_REQUIRE(a && stats6, MGB_ERR_INVALID, "null pointer");
  for (int i = 0; i < 6; ++i) stats6[i] = a->stats[i];
  return MGB_OK;
}

extern "C" int mgb_agg_destroy(mgb_agg* a) {
  if (!a) return MGB_OK;
  cudaStream_t st = a->st;
  for (auto& u : a->updates) {
    if (u.ctl) cudaFreeAsync(u.ctl, st);
    if (u.stg.base) cudaFreeAsync(u.stg.base, st);
    if (u.miss) cudaFreeAsync(u.miss, st);
  }
  if (a->table) cudaFreeAsync(a->table, st);
  if (a->gctl) cudaFreeAsync(a->gctl, st);
  mgb_bucket_free(&a->bkt, st);
  delete a;
  return MGB_OK;
}
