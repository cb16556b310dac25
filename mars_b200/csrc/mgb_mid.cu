// libmarsgb: mid-cardinality map-side pre-aggregation (sm_100a) -- a4, DataFrameGroupByAgg._execute_map
// (mars/dataframe/groupby/aggregation.py:1043-1128) for the sum / mean / count shapes when a band holds more
// distinct key tuples than the dense kernel keeps per CTA (mgb_dense.cuh: 8..16) but few enough for one
// shared-memory table: up to MID_MAX_GROUPS (BASELINE.json config 1: 1000 keys, 4 float64 columns).
//
// One launch streams ALL row batches of a band (same chunk table as the dense batch: the chunks form one virtual
// sequence of 1024-row tiles, register double-buffered).  Per CTA (one per SM):
//   * key table: open addressing over MID_TAB slots that map a key tuple to a dense group id (-1 empty, -2 being
//     published); lookups and inserts are warp-converged retry loops (a lane never spins on a slot that another
//     lane of its own warp is publishing).  The table is kept sparse (load <= 0.375) because a warp iterates as
//     long as its unluckiest lane probes (ncu, first version at load 0.5: 6.5 rounds per row);
//   * accumulators: ONE record per group -- [sum 0 .. sum NV-1 | rows (u32)] -- updated with shared-memory
//     atomics (float64 adds are compare-and-swap loops in SASS; the loops of a row's columns overlap).  A
//     per-record try-lock with plain loads / stores was measured and dropped: 1.1 TB/s against 1.5 TB/s, its
//     acquire -> load -> add -> store -> release chain is strictly serial (profiles/README.md, round 2).
//   * count(x) == rows of the group: a row that carries a NaN raises the "not this path" flag (the call then
//     reports -1 and the caller takes the general path, which keeps per-column NaN bookkeeping).
// At the end every CTA folds its occupied slots into a global open-addressing table (state claim + RED.ADD.F64 /
// RED.ADD.U64 per accumulator of the spec), and a second tiny kernel compacts that table into the caller's
// columns and publishes the group count on the device (no host synchronisation anywhere: the result feeds the
// speculative operand chain like the dense path's).
#include "mgb_dense.cuh"

#define MID_THREADS 512
#define MID_R 2                    // rows per thread per tile: a tile is 1024 rows
#define MID_TAB 4096               // slots of the per-CTA key table (load <= 0.375: short probe sequences)
#define MID_CTA_GROUPS 1536        // accumulator records per CTA == groups a CTA holds before it gives up
#define MID_GTAB 4096              // slots of the global merge table
#define MID_MAX_GROUPS 2048        // groups of a band this path reports (== capacity of the small merge)
#define MID_EMIT_THREADS 1024

struct MidParams {
  int nv;                          // loaded float64 value columns (1..8)
  int rec_w;                       // 8-byte words per accumulator record: nv sums + 1 (rows | lock), even
  int na;                          // spec accumulators
  uint8_t acc_col[MGB_MAX_ACCS];   // loaded column of accumulator a, 0xFF: a count that reads no column
  uint8_t acc_is_count[MGB_MAX_ACCS];
  // global merge table: state[MID_GTAB] (u32) | keys[NK][MID_GTAB] | acc[na][MID_GTAB], zero-initialised
  uint32_t* gstate;
  uint64_t* gkeys;
  uint64_t* gacc;
  int* flags;                      // [0] fail (too many groups / NaN value / table full), [1] groups in the global table
};

struct MidEmit {
  const uint32_t* gstate;
  const uint64_t* gkeys;
  const uint64_t* gacc;
  const int* flags;
  int nk, na, cap;
  uint64_t* out_key[MGB_MAX_KEYS];
  uint64_t* out_acc[MGB_MAX_ACCS];
  long long* out_n;
};

template <int NK>
__device__ __forceinline__ uint32_t mid_hash(const uint64_t* k) {
  uint64_t h = mgb_mix64(k[0]);
  if (NK == 2) h = mgb_mix64(h ^ (k[1] * 0x9e3779b97f4a7c15ULL));
  return (uint32_t)(h >> 32);
}

template <int NK, int NV>
__global__ void __launch_bounds__(MID_THREADS, 1) k_preagg_mid(const MidParams p, const DenseBatch bt) {
  extern __shared__ __align__(16) uint64_t msm[];
  // layout: rec[MID_CTA_GROUPS][rec_w] | tkeys[NK][MID_TAB] | tgid[MID_TAB] (int) | ints | chunk table
  const int RW = p.rec_w;
  uint64_t* rec = msm;
  uint64_t* tkeys = rec + (size_t)MID_CTA_GROUPS * RW;
  int* tgid = (int*)(tkeys + (size_t)NK * MID_TAB);
  int* s_int = tgid + MID_TAB;
  int* s_count = s_int;          // groups in this CTA's table
  int* s_fail = s_int + 1;
  long long* s_tile_end = (long long*)(s_int + 8);
  long long* s_rows = s_tile_end + DN_MAX_CHUNKS;
  const uint64_t** s_ptr = (const uint64_t**)(s_rows + DN_MAX_CHUNKS);

  for (int i = threadIdx.x; i < MID_CTA_GROUPS * RW; i += MID_THREADS) rec[i] = 0;
  for (int i = threadIdx.x; i < MID_TAB; i += MID_THREADS) tgid[i] = -1;
  const int n_chunks = bt.n_chunks;
  for (int i = threadIdx.x; i < n_chunks; i += MID_THREADS) {
    s_tile_end[i] = bt.tile_end[i];
    s_rows[i] = bt.n[i];
  }
  for (int i = threadIdx.x; i < n_chunks * DN_PTRS; i += MID_THREADS) s_ptr[i] = bt.ptr[i / DN_PTRS][i % DN_PTRS];
  if (threadIdx.x == 0) {
    *s_count = 0;
    *s_fail = 0;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  bool failed = false;

  // one row per lane: slot lookup / insert, then the locked record update.  Called by all 32 lanes together.
  auto process_row = [&](const uint64_t* k, const uint64_t* x, const bool live) {
    // ---- phase A: slot of the key tuple
    uint32_t s = mid_hash<NK>(k) & (MID_TAB - 1);
    int slot = -1;
    bool done = !live;
    int probes = 0;
    while (__any_sync(FULL, !done)) {
      if (!done) {
        const int tg = ((volatile int*)tgid)[s];
        if (tg >= 0) {
          bool eq = ((volatile uint64_t*)tkeys)[s] == k[0];
          if (NK == 2) eq = eq && ((volatile uint64_t*)tkeys)[MID_TAB + s] == k[1];
          if (eq) {
            slot = tg;
            done = true;
          } else {
            s = (s + 1) & (MID_TAB - 1);
            if (++probes > MID_TAB) done = true;      // cannot happen below the group limit
          }
        } else if (tg == -1) {
          if (*(volatile int*)s_count >= MID_CTA_GROUPS) {
            done = true;                               // too many groups for this path
          } else if (atomicCAS(&tgid[s], -1, -2) == -1) {
            const int id = atomicAdd(s_count, 1);
            if (id < MID_CTA_GROUPS) {
              ((volatile uint64_t*)tkeys)[s] = k[0];
              if (NK == 2) ((volatile uint64_t*)tkeys)[MID_TAB + s] = k[1];
              __threadfence_block();
              ((volatile int*)tgid)[s] = id;
              slot = id;
            } else {
              ((volatile int*)tgid)[s] = -3;           // poisoned: nobody waits on it, the launch fails
            }
            done = true;
          }
          // lost the claim: the winner may hold this very key -- look at the slot again next round
        } else if (tg == -3) {
          done = true;
        }
        // tg == -2: being published by another thread, look again next round
      }
    }
    bool bad = live && slot < 0;
    // ---- values: a NaN cannot be skipped here (count == rows of the group): not this path
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const double xv = __longlong_as_double((long long)x[c]);
      bad = bad || (live && xv != xv);
    }
    if (__any_sync(FULL, bad)) {
      failed = true;
      return;
    }
    // ---- phase B: the record of the group.  Shared memory has no native float64 add; the compare-and-swap loops
    // of the NV columns are independent of each other and overlap (measured faster than one per-record lock with
    // plain loads / stores, whose acquire -> load -> add -> store -> release chain is strictly serial).
    if (live) {
      uint64_t* r = rec + (size_t)slot * RW;
#pragma unroll
      for (int c = 0; c < NV; ++c) atomicAdd((double*)(r + c), __longlong_as_double((long long)x[c]));
      atomicAdd((uint32_t*)(r + NV), 1u);
    }
  };

  constexpr long long TILE = (long long)MID_THREADS * MID_R;
  const long long nfull = s_tile_end[n_chunks - 1];
  uint64_t kb[2][MID_R][NK];
  uint64_t xb[2][MID_R][NV];
  const uint64_t* pk[NK];
  const uint64_t* pc[NV];
  int cur = -1, cc = 0;
  long long cbeg = 0;
  auto load_tile = [&](int buf, long long t) {
    while (t >= s_tile_end[cc]) {
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
    for (int j = 0; j < NK; ++j)
#pragma unroll
      for (int r = 0; r < MID_R; ++r) kb[buf][r][j] = __ldcs((const unsigned long long*)(pk[j] + off) + r * MID_THREADS);
#pragma unroll
    for (int c = 0; c < NV; ++c)
#pragma unroll
      for (int r = 0; r < MID_R; ++r) xb[buf][r][c] = __ldcs((const unsigned long long*)(pc[c] + off) + r * MID_THREADS);
  };
  auto process_tile = [&](int buf) {
#pragma unroll
    for (int r = 0; r < MID_R; ++r) {
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
    if ((++iter & 7) == 0 && *(volatile int*)p.flags) failed = true;   // somebody else gave up: stop early
  }
  for (int c = (int)blockIdx.x; c < n_chunks && !failed; c += (int)gridDim.x) {   // < TILE-row remainders
    const long long n = s_rows[c];
    const long long done_rows = (s_tile_end[c] - (c ? s_tile_end[c - 1] : 0)) * TILE;
    const uint64_t* const* cp = s_ptr + c * DN_PTRS;
    for (long long i0 = done_rows + warp * 32; i0 < n && !failed; i0 += MID_THREADS) {
      const long long i = i0 + lane;
      const bool live = i < n;
      uint64_t k[NK], x[NV];
#pragma unroll
      for (int j = 0; j < NK; ++j) k[j] = live ? cp[j][i] : 0;
#pragma unroll
      for (int cl = 0; cl < NV; ++cl) x[cl] = live ? cp[MGB_MAX_KEYS + cl][i] : 0;
      process_row(k, x, live);
    }
  }
  if (failed) {
    *s_fail = 1;
    *p.flags = 1;
  }
  __syncthreads();
  if (*s_fail || *(volatile int*)p.flags) return;

  // ---- fold this CTA's groups into the global table (uniform trip count: the claim loop votes per warp)
  for (int i0 = 0; i0 < MID_TAB; i0 += MID_THREADS) {
    const int i = i0 + (int)threadIdx.x;
    const int my = tgid[i];
    const bool live = my >= 0;
    uint64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = live ? tkeys[(size_t)j * MID_TAB + i] : 0;
    uint32_t g = (mid_hash<NK>(k) >> 11) & (MID_GTAB - 1);
    int gs = -1;
    bool done = !live;
    int probes = 0;
    while (__any_sync(FULL, !done)) {
      if (!done) {
        const uint32_t st = *(volatile uint32_t*)&p.gstate[g];
        if (st == 2u) {
          bool eq = *(volatile uint64_t*)&p.gkeys[g] == k[0];
          if (NK == 2) eq = eq && *(volatile uint64_t*)&p.gkeys[MID_GTAB + g] == k[1];
          if (eq) {
            gs = (int)g;
            done = true;
          } else {
            g = (g + 1) & (MID_GTAB - 1);
            if (++probes > MID_GTAB) done = true;
          }
        } else if (st == 0u) {
          if (*(volatile int*)&p.flags[1] >= MID_MAX_GROUPS) {
            done = true;                                   // more groups in the band than this path reports
          } else if (atomicCAS(&p.gstate[g], 0u, 1u) == 0u) {
            *(volatile uint64_t*)&p.gkeys[g] = k[0];
            if (NK == 2) *(volatile uint64_t*)&p.gkeys[MID_GTAB + g] = k[1];
            atomicAdd(&p.flags[1], 1);
            __threadfence();
            *(volatile uint32_t*)&p.gstate[g] = 2u;
            gs = (int)g;
            done = true;
          }
        }
      }
    }
    if (live && gs < 0) {
      *p.flags = 1;
    } else if (live) {
      const uint64_t* r = rec + (size_t)my * RW;
      const unsigned long long rows = (unsigned long long)*(const uint32_t*)(r + NV);
      for (int a = 0; a < p.na; ++a) {
        uint64_t* dst = p.gacc + (size_t)a * MID_GTAB + gs;
        if (p.acc_is_count[a])
          atomicAdd((unsigned long long*)dst, rows);
        else
          atomicAdd((double*)dst, __longlong_as_double((long long)r[p.acc_col[a]]));
      }
    }
  }
}

// compaction of the global table into the caller's columns (slot order), group count to the device
__global__ void __launch_bounds__(MID_EMIT_THREADS) k_mid_emit(const MidEmit e) {
  __shared__ int warp_tot[MID_EMIT_THREADS / 32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (e.flags[0]) {
    if (threadIdx.x == 0) *e.out_n = -1;
    return;
  }
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < MID_GTAB; base += MID_EMIT_THREADS) {
    const int g = base + (int)threadIdx.x;
    const bool occ = e.gstate[g] == 2u;
    const unsigned m = __ballot_sync(0xffffffffu, occ);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = carry + __popc(m & mgb_lanemask_lt());
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (occ && off < e.cap) {
      for (int j = 0; j < e.nk; ++j) e.out_key[j][off] = e.gkeys[(size_t)j * MID_GTAB + g];
      for (int a = 0; a < e.na; ++a) e.out_acc[a][off] = e.gacc[(size_t)a * MID_GTAB + g];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < MID_EMIT_THREADS / 32; ++w) tot += warp_tot[w];
      carry += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *e.out_n = carry <= e.cap ? carry : -1;
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
struct MidScratch {
  char* table = nullptr;     // gstate | gkeys | gacc | flags
  size_t bytes = 0;
};
static thread_local MidScratch g_mid[8];

static size_t mid_table_bytes() {
  return (size_t)MID_GTAB * 4 + (size_t)MGB_MAX_KEYS * MID_GTAB * 8 + (size_t)MGB_MAX_ACCS * MID_GTAB * 8 + 64;
}

extern "C" int mgb_mid_capacity(int64_t* out_rows) {
  MGB_REQUIRE(out_rows, MGB_ERR_INVALID, "null pointer");
  *out_rows = MID_MAX_GROUPS;
  return MGB_OK;
}

static int mid_set_count(int64_t* out_n_dev, int value, cudaStream_t st) {
  MGB_CUDA(cudaMemsetAsync(out_n_dev, value < 0 ? 0xFF : 0, sizeof(int64_t), st));
  return MGB_OK;
}

extern "C" int mgb_preagg_mid_batch_async(const mgb_agg_spec* spec, int32_t n_chunks, const int64_t* chunk_rows,
                                          const int64_t* const* const* chunk_key_cols,
                                          const void* const* const* chunk_val_cols, int64_t* const* out_key_cols,
                                          void* const* out_acc_cols, int64_t out_capacity, int64_t* out_n_dev,
                                          void* stream) {
  MGB_REQUIRE(spec && chunk_rows && chunk_key_cols && chunk_val_cols && out_key_cols && out_acc_cols && out_n_dev,
              MGB_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int nk = spec->n_keys, na = spec->n_accs;
  MGB_REQUIRE(nk >= 1 && nk <= MGB_MAX_KEYS, MGB_ERR_UNSUPPORTED, "n_keys=%d", nk);
  MGB_REQUIRE(na >= 1 && na <= MGB_MAX_ACCS && spec->n_vals >= 1 && spec->n_vals <= MGB_MAX_VALS,
              MGB_ERR_UNSUPPORTED, "spec outside the supported envelope");
  MGB_REQUIRE(n_chunks >= 0 && n_chunks <= DN_MAX_CHUNKS, MGB_ERR_UNSUPPORTED,
              "%d chunks in one launch (mgb_dense_batch_max: %d)", n_chunks, DN_MAX_CHUNKS);
  MGB_REQUIRE(out_capacity >= MID_MAX_GROUPS, MGB_ERR_INVALID, "output capacity %lld < %d (mgb_mid_capacity)",
              (long long)out_capacity, MID_MAX_GROUPS);
  // this path takes the sum / mean / count shapes: float64 columns with SUM and / or COUNT; int64 columns that
  // are only counted are not read (count == rows of the group)
  MidParams p;
  memset(&p, 0, sizeof(p));
  int loaded_of[MGB_MAX_VALS];
  int ncol = 0;
  for (int c = 0; c < spec->n_vals; ++c) {
    uint32_t m = 0;
    for (int a = 0; a < na; ++a)
      if (spec->acc_col[a] == c) m |= OPB(spec->acc_op[a]);
    loaded_of[c] = -1;
    if (m == 0) continue;
    if (m & ~(OPB(MGB_SUM) | OPB(MGB_COUNT))) return mid_set_count(out_n_dev, -1, st);
    if (spec->val_dtype[c] == MGB_I64) {
      if (m != OPB(MGB_COUNT)) return mid_set_count(out_n_dev, -1, st);
      continue;
    }
    if (ncol == 8) return mid_set_count(out_n_dev, -1, st);
    loaded_of[c] = ncol++;
  }
  if (ncol == 0) return mid_set_count(out_n_dev, -1, st);
  p.nv = ncol;
  p.rec_w = (ncol + 1 + 1) & ~1;     // even number of words: records stay 16-byte aligned
  p.na = na;
  for (int a = 0; a < na; ++a) {
    const int c = spec->acc_col[a], op = spec->acc_op[a];
    MGB_REQUIRE(c >= 0 && c < spec->n_vals, MGB_ERR_INVALID, "bad accumulator %d", a);
    p.acc_is_count[a] = op == MGB_COUNT ? 1 : 0;
    p.acc_col[a] = (uint8_t)(loaded_of[c] < 0 ? 0xFF : loaded_of[c]);
    MGB_REQUIRE(out_acc_cols[a], MGB_ERR_INVALID, "null output accumulator column %d", a);
  }
  const long long tile_rows = 1024;
  DenseBatch bt;
  memset(&bt, 0, sizeof(bt));
  long long tiles = 0, work_tiles = 0;
  for (int i = 0; i < n_chunks; ++i) {
    MGB_REQUIRE(chunk_rows[i] >= 0, MGB_ERR_INVALID, "negative row count of chunk %d", i);
    if (chunk_rows[i] == 0) continue;
    MGB_REQUIRE(chunk_key_cols[i] && chunk_val_cols[i], MGB_ERR_INVALID, "null column table of chunk %d", i);
    const int q = bt.n_chunks++;
    for (int j = 0; j < nk; ++j) {
      MGB_REQUIRE(chunk_key_cols[i][j], MGB_ERR_INVALID, "null key column %d of chunk %d", j, i);
      bt.ptr[q][j] = (const uint64_t*)chunk_key_cols[i][j];
    }
    for (int c = 0; c < spec->n_vals; ++c) {
      if (loaded_of[c] < 0) continue;
      MGB_REQUIRE(chunk_val_cols[i][c], MGB_ERR_INVALID, "null value column %d of chunk %d", c, i);
      bt.ptr[q][MGB_MAX_KEYS + loaded_of[c]] = (const uint64_t*)chunk_val_cols[i][c];
    }
    bt.n[q] = chunk_rows[i];
    tiles += chunk_rows[i] / tile_rows;
    bt.tile_end[q] = tiles;
    work_tiles += (chunk_rows[i] + tile_rows - 1) / tile_rows;
  }
  if (bt.n_chunks == 0) return mid_set_count(out_n_dev, 0, st);
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  MGB_REQUIRE(dev >= 0 && dev < 8, MGB_ERR_UNSUPPORTED, "device ordinal %d", dev);
  MidScratch& sc = g_mid[dev];
  if (!sc.table) {
    MGB_CUDA(cudaMalloc((void**)&sc.table, mid_table_bytes()));
    sc.bytes = mid_table_bytes();
  }
  // only the words this call touches are cleared: state, keys, na accumulator columns, flags
  p.gstate = (uint32_t*)sc.table;
  p.gkeys = (uint64_t*)(sc.table + (size_t)MID_GTAB * 4);
  p.gacc = p.gkeys + (size_t)MGB_MAX_KEYS * MID_GTAB;
  p.flags = (int*)(sc.table + mid_table_bytes() - 64);
  MGB_CUDA(cudaMemsetAsync(sc.table, 0, (size_t)MID_GTAB * 4 + (size_t)MGB_MAX_KEYS * MID_GTAB * 8 +
                                            (size_t)na * MID_GTAB * 8, st));
  MGB_CUDA(cudaMemsetAsync(p.flags, 0, 64, st));
  const int sms = mgb_sm_count();
  const int grid = (int)(work_tiles < sms ? work_tiles : sms);
  const size_t smem = (size_t)MID_CTA_GROUPS * p.rec_w * 8 + (size_t)nk * MID_TAB * 8 + (size_t)MID_TAB * 4 + 32 +
                      (size_t)DN_MAX_CHUNKS * (2 + DN_PTRS) * 8 + 64;
  MGB_REQUIRE(smem <= 227 * 1024, MGB_ERR_OVERFLOW, "mid path shared-memory plan does not fit (%zu)", smem);
  {
    MgbProfScope prof(MGB_K_PREAGG_SMEM, st);
#define MID_LAUNCH(NK, NV)                                                                                        \
  do {                                                                                                            \
    static bool attr_set[8] = {false};                                                                            \
    if (!attr_set[dev]) {                                                                                         \
      MGB_CUDA(cudaFuncSetAttribute(k_preagg_mid<NK, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr_set[dev] = true;                                                                                       \
    }                                                                                                             \
    k_preagg_mid<NK, NV><<<grid, MID_THREADS, smem, st>>>(p, bt);                                                 \
  } while (0)
#define MID_PICK(NK)                          \
  switch (ncol) {                             \
    case 1: MID_LAUNCH(NK, 1); break;         \
    case 2: MID_LAUNCH(NK, 2); break;         \
    case 3: MID_LAUNCH(NK, 3); break;         \
    case 4: MID_LAUNCH(NK, 4); break;         \
    case 5: MID_LAUNCH(NK, 5); break;         \
    case 6: MID_LAUNCH(NK, 6); break;         \
    case 7: MID_LAUNCH(NK, 7); break;         \
    default: MID_LAUNCH(NK, 8); break;        \
  }
    if (nk == 1) { MID_PICK(1) } else { MID_PICK(2) }
#undef MID_PICK
#undef MID_LAUNCH
    MGB_CHECK_LAUNCH();
  }
  MidEmit e;
  memset(&e, 0, sizeof(e));
  e.gstate = p.gstate;
  e.gkeys = p.gkeys;
  e.gacc = p.gacc;
  e.flags = p.flags;
  e.nk = nk;
  e.na = na;
  e.cap = (int)(out_capacity < MID_MAX_GROUPS ? out_capacity : MID_MAX_GROUPS);
  for (int j = 0; j < nk; ++j) {
    MGB_REQUIRE(out_key_cols[j], MGB_ERR_INVALID, "null output key column %d", j);
    e.out_key[j] = (uint64_t*)out_key_cols[j];
  }
  for (int a = 0; a < na; ++a) e.out_acc[a] = (uint64_t*)out_acc_cols[a];
  e.out_n = (long long*)out_n_dev;
  {
    MgbProfScope prof(MGB_K_EMIT, st);
    k_mid_emit<<<1, MID_EMIT_THREADS, 0, st>>>(e);
    MGB_CHECK_LAUNCH();
  }
  return MGB_OK;
}
