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
#include <stdlib.h>

#include "mgb_dense.cuh"

int mgb_dense_launch_1_256(const DenseParams&, const DenseBatch&, int, bool, int, size_t, cudaStream_t);
int mgb_dense_launch_1_512(const DenseParams&, const DenseBatch&, int, bool, int, size_t, cudaStream_t);
int mgb_dense_launch_2_256(const DenseParams&, const DenseBatch&, int, bool, int, size_t, cudaStream_t);
int mgb_dense_launch_2_512(const DenseParams&, const DenseBatch&, int, bool, int, size_t, cudaStream_t);
// mgb_tiny.cu: <= 4 groups per CTA, accumulators in registers, TMA-staged tiles
bool mgb_tiny_eligible(const DenseBatch& bt, int nk, int ncol, int n_load);
int mgb_tiny_launch(const DenseParams& dp, const DenseBatch& bt, int nk, int ncol, int grid, int* ticket_fail2, int* sticky,
                    cudaStream_t st);

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
static thread_local int g_last_launches = 0;   // kernels enqueued by the last dense / batch call of this thread

extern "C" int mgb_dense_last_launches(int32_t* out_launches) {
  MGB_REQUIRE(out_launches, MGB_ERR_INVALID, "null pointer");
  *out_launches = g_last_launches;
  return MGB_OK;
}

static int dense_scratch(int dev, size_t stg_words, int grid, DenseScratch** out) {
  MGB_REQUIRE(dev >= 0 && dev < 8, MGB_ERR_UNSUPPORTED, "device ordinal %d", dev);
  DenseScratch& s = g_ds[dev];
  if (s.stg_words < stg_words) {
    if (s.stg) cudaFree(s.stg);
    MGB_CUDA(cudaMalloc((void**)&s.stg, stg_words * 8));
    s.stg_words = stg_words;
  }
  if (!s.ints) {
    MGB_CUDA(cudaMalloc((void**)&s.ints, sizeof(int) * (1024 + 128)));   // ... | tiny: ticket, fail | per-spec sticky flags
    MGB_CUDA(cudaMemset(s.ints, 0, sizeof(int) * (1024 + 128)));
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

// sets *out_n_dev (device) to `value` in stream order: -1 = "not a dense shape", 0 = empty input
static int dense_set_count(int64_t* out_n_dev, int value, cudaStream_t st) {
  MGB_CUDA(cudaMemsetAsync(out_n_dev, value < 0 ? 0xFF : 0, sizeof(int64_t), st));
  return MGB_OK;
}

extern "C" int mgb_dense_batch_max(int32_t* out_chunks) {
  MGB_REQUIRE(out_chunks, MGB_ERR_INVALID, "null pointer");
  *out_chunks = DN_MAX_CHUNKS;
  return MGB_OK;
}

static int dense_batch_impl(const mgb_agg_spec* spec, const mgb_row_program* prog, int32_t n_chunks,
                            const int64_t* chunk_rows, const int64_t* const* const* chunk_key_cols,
                            const void* const* const* chunk_val_cols, int64_t* const* out_key_cols,
                            void* const* out_acc_cols, int64_t out_capacity, int64_t* out_n_dev, void* stream) {
  g_last_launches = 0;
  MGB_REQUIRE(spec && chunk_rows && chunk_key_cols && chunk_val_cols && out_key_cols && out_acc_cols && out_n_dev,
              MGB_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int nk = spec->n_keys, na = spec->n_accs;
  MGB_REQUIRE(nk >= 1 && nk <= MGB_MAX_KEYS, MGB_ERR_UNSUPPORTED, "n_keys=%d", nk);
  MGB_REQUIRE(na >= 1 && na <= MGB_MAX_ACCS && spec->n_vals >= 1 && spec->n_vals <= MGB_MAX_VALS,
              MGB_ERR_UNSUPPORTED, "spec outside the supported envelope");
  MGB_REQUIRE(n_chunks >= 0 && n_chunks <= DN_MAX_CHUNKS, MGB_ERR_UNSUPPORTED,
              "%d chunks in one launch (mgb_dense_batch_max: %d)", n_chunks, DN_MAX_CHUNKS);
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
    if (ncol == 8) return dense_set_count(out_n_dev, -1, st);   // more than 8 live columns: not a dense shape
    loaded_of[c] = ncol;
    if (spec->val_dtype[c] == MGB_F64) p.f64_mask |= 1u << ncol;
    for (int o = MGB_SUM; o <= MGB_MAX; ++o)
      if (m & OPB(o)) p.op_mask[o] |= 1u << ncol;
    if (m & (OPB(MGB_SUMSQ) | OPB(MGB_MIN) | OPB(MGB_MAX))) allops = true;
    ++ncol;
  }
  if (ncol == 0) return dense_set_count(out_n_dev, -1, st);   // only counts of int64 columns: general path
  if (prog) {
    // fused pre-path: sum / mean / count shapes over float64 logical columns; the program's column indices count
    // the LOADED columns (the spec's value columns minus int64 columns that are only counted, in spec order)
    const uint32_t all = (1u << ncol) - 1u;
    if (allops || p.f64_mask != all || p.op_mask[MGB_SUM] != all) return dense_set_count(out_n_dev, -1, st);
    MGB_REQUIRE(prog->n_phys >= 1 && prog->n_phys <= ncol, MGB_ERR_UNSUPPORTED,
                "row program: %d physical columns for %d logical ones", prog->n_phys, ncol);
    p.prog.n_phys = prog->n_phys;
    p.prog.filter_col = prog->filter_col;
    MGB_REQUIRE(prog->filter_col < prog->n_phys && prog->filter_op >= 0 && prog->filter_op <= 5, MGB_ERR_INVALID,
                "bad row filter");
    p.prog.filter_op = prog->filter_op;
    p.prog.filter_f64 = prog->filter_dtype == MGB_F64;
    p.prog.filter_i = prog->filter_i64;
    p.prog.filter_d = prog->filter_f64;
    for (int c = 0; c < ncol; ++c) {
      const int nf = prog->n_factors[c];
      MGB_REQUIRE(nf >= 0 && nf <= 3, MGB_ERR_INVALID, "logical column %d: %d factors", c, nf);
      MGB_REQUIRE(nf > 0 || c < prog->n_phys, MGB_ERR_INVALID, "logical column %d has no physical column", c);
      p.prog.nf[c] = (signed char)nf;
      for (int f = 0; f < nf; ++f) {
        MGB_REQUIRE(prog->factors[c][f].col >= 0 && prog->factors[c][f].col < prog->n_phys, MGB_ERR_INVALID,
                    "logical column %d reads physical column %d of %d", c, prog->factors[c][f].col, prog->n_phys);
        p.prog.fcol[c][f] = (signed char)prog->factors[c][f].col;
        p.prog.falpha[c][f] = prog->factors[c][f].alpha;
        p.prog.fbeta[c][f] = prog->factors[c][f].beta;
      }
    }
  }
  // the batch: non-empty chunks as one virtual sequence of 1024-row tiles
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
    if (prog) {      // physical columns of the row program, in program order
      for (int j = 0; j < prog->n_phys; ++j) {
        MGB_REQUIRE(chunk_val_cols[i][j], MGB_ERR_INVALID, "null physical column %d of chunk %d", j, i);
        bt.ptr[q][MGB_MAX_KEYS + j] = (const uint64_t*)chunk_val_cols[i][j];
      }
    } else {
      for (int c = 0; c < spec->n_vals; ++c) {
        if (loaded_of[c] < 0) continue;
        MGB_REQUIRE(chunk_val_cols[i][c], MGB_ERR_INVALID, "null value column %d of chunk %d", c, i);
        bt.ptr[q][MGB_MAX_KEYS + loaded_of[c]] = (const uint64_t*)chunk_val_cols[i][c];
      }
    }
    bt.n[q] = chunk_rows[i];
    tiles += chunk_rows[i] / tile_rows;
    bt.tile_end[q] = tiles;
    work_tiles += (chunk_rows[i] + tile_rows - 1) / tile_rows;
  }
  if (bt.n_chunks == 0) return dense_set_count(out_n_dev, 0, st);
  for (int j = 0; j < nk; ++j) {
    MGB_REQUIRE(out_key_cols[j], MGB_ERR_INVALID, "null output key column %d", j);
    p.out.key[j] = (uint64_t*)out_key_cols[j];
  }
  // accumulator block of one (warp, group), 32-bit words: rows[32] | sum[ncol][64] | (sumsq|min|max)[ncol][64]
  // | nan[32] (lane-private 8-byte slots except the per-column NaN counters, which are touched rarely)
  const uint32_t all_cols = (1u << ncol) - 1u;
  const bool plain = !allops && p.f64_mask == all_cols && p.op_mask[MGB_SUM] == all_cols;
  // only the op kinds that occur get slots (accumulator block) and words (staged rows)
  int off = 32, w = 1;
  const int kind_op[5] = {-1, MGB_SUM, MGB_SUMSQ, MGB_MIN, MGB_MAX};
  int* kind_acc_off[5] = {nullptr, &p.off_sum, &p.off_sumsq, &p.off_min, &p.off_max};
  for (int kind = 1; kind <= 4; ++kind) {
    *kind_acc_off[kind] = off;
    p.kind_off[kind] = w;
    if (p.op_mask[kind_op[kind]]) p.word_kind[w] = kind;
    if (p.op_mask[kind_op[kind]]) {
      off += 64 * ncol;
      ++w;
    }
  }
  p.off_nan = off;
  p.stride_g = p.off_nan + 32;   // a multiple of 32 words: the bank of a lane-private slot depends on the lane only
  p.wpc = w;
  // key table + group keys + table ids + ints + the chunk table (tile ends, rows, column pointers)
  const size_t fixed = (size_t)(MGB_MAX_KEYS * DN_TAB + MGB_MAX_KEYS * DN_MAX_GC) * 8 + DN_TAB * 4 + 128 +
                       (size_t)DN_MAX_CHUNKS * (2 + DN_PTRS) * 8;
  // 16 warps per CTA when that still leaves 8 group slots (and registers for the double-buffered rows),
  // otherwise 8 warps with twice the rows per thread
  int threads = 512;
  int GC = (int)((DN_SMEM_BUDGET - fixed) / ((size_t)(threads / 32) * p.stride_g * 4));
  static const char* env_threads = getenv("MGB_DENSE_THREADS");   // experiments only
  static const char* env_gc = getenv("MGB_DENSE_GC");
  if (GC < 8 || nk + ncol > 8 || (env_threads && atoi(env_threads) == 256)) {
    threads = 256;
    GC = (int)((DN_SMEM_BUDGET - fixed) / ((size_t)(threads / 32) * p.stride_g * 4));
  }
  if (GC > DN_MAX_GC) GC = DN_MAX_GC;
  if (env_gc && atoi(env_gc) >= 2 && atoi(env_gc) < GC) GC = atoi(env_gc);
  if (GC < 2) return dense_set_count(out_n_dev, -1, st);   // accumulators of one group do not fit: general path
  p.GC = GC;
  const int n_warps = threads / 32;
  const int sms = mgb_sm_count();
  const int grid = (int)(work_tiles < sms ? work_tiles : sms);
  MGB_REQUIRE(out_capacity >= (int64_t)grid * GC, MGB_ERR_INVALID,
              "output capacity %lld < %d (mgb_dense_capacity)", (long long)out_capacity, grid * GC);
  p.stg_w = nk + 1 + p.wpc * ncol;
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  DenseScratch* s = nullptr;
  MGB_TRY(dense_scratch(dev, (size_t)grid * GC * p.stg_w, grid, &s));
  p.stg = s->stg;
  p.cta_ng = s->ints;
  p.ticket = (unsigned int*)(s->ints + 1024);
  p.fail = s->ints + 1025;
  p.out_n = (long long*)out_n_dev;
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
  // the CTA owns its SM: take (almost) all shared memory; the last CTA re-uses it for the cross-CTA merge
  const size_t acc_smem = (size_t)n_warps * GC * p.stride_g * 4 + fixed;
  size_t smem = acc_smem < 112 * 1024 ? 112 * 1024 : acc_smem;   // >= 112 KB for the cross-CTA merge tables
  // last-CTA merge: key table + 5 int arrays of at most grid * GC (+2) entries
  const size_t merge_smem = (size_t)MGB_MAX_KEYS * SM_TAB * 8 + SM_TAB * 4 + 32 + 5 * ((size_t)grid * GC + 4) * 4 +
                            (size_t)DN_MAX_GC * (p.stg_w + MGB_MAX_KEYS) * 8;
  MGB_REQUIRE(acc_smem <= smem && merge_smem <= smem, MGB_ERR_OVERFLOW, "dense path shared-memory plan does not fit");
  static const char* env_dbg = getenv("MGB_DENSE_DEBUG");   // experiments only: per-CTA phase timestamps
  long long* dbg = nullptr;
  if (env_dbg) {
    MGB_CUDA(cudaMalloc((void**)&dbg, sizeof(long long) * (4 * grid + 8)));
    MGB_CUDA(cudaMemset(dbg, 0, sizeof(long long) * (4 * grid + 8)));
    p.dbg = dbg;
  }
  g_last_launches = 1;
  {
    MgbProfScope prof(MGB_K_PREAGG_TINY, st);
    int status;
    // sum / mean / count shapes: first the kernel that keeps <= 4 groups per CTA in registers and stages its tiles by
    // TMA (mgb_tiny.cu); the dense kernel is enqueued right behind it and exits at once unless that one gave up
    static const char* env_tiny = getenv("MGB_TINY");
    if (plain && GC >= 4 && (!env_tiny || atoi(env_tiny) != 0) &&
        mgb_tiny_eligible(bt, nk, ncol, prog ? prog->n_phys : ncol)) {
      uint32_t h = (uint32_t)(nk * 131 + ncol * 31 + na + (prog ? 7919 : 0));
      for (int a = 0; a < na; ++a) h = h * 16777619u ^ (uint32_t)(spec->acc_col[a] * 8 + spec->acc_op[a]);
      int* tf = s->ints + 1026;
      MGB_CUDA(cudaMemsetAsync(tf, 0, 2 * sizeof(int), st));
      MGB_TRY(mgb_tiny_launch(p, bt, nk, ncol, grid, tf, s->ints + 1032 + (int)(h % 64u), st));
      p.run_if = tf + 1;
      g_last_launches = 2;
    }
    if (nk == 1)
      status = threads == 512 ? mgb_dense_launch_1_512(p, bt, ncol, plain, grid, smem, st)
                              : mgb_dense_launch_1_256(p, bt, ncol, plain, grid, smem, st);
    else
      status = threads == 512 ? mgb_dense_launch_2_512(p, bt, ncol, plain, grid, smem, st)
                              : mgb_dense_launch_2_256(p, bt, ncol, plain, grid, smem, st);
    MGB_TRY(status);
  }
  if (dbg) {
    MGB_CUDA(cudaStreamSynchronize(st));
    long long* h = (long long*)malloc(sizeof(long long) * (4 * grid + 8));
    cudaMemcpy(h, dbg, sizeof(long long) * (4 * grid + 8), cudaMemcpyDeviceToHost);
    long long t0 = h[0], mx[4] = {0, 0, 0, 0}, mn[4] = {1LL << 62, 1LL << 62, 1LL << 62, 1LL << 62};
    for (int i = 0; i < grid; ++i) if (h[i * 4] < t0) t0 = h[i * 4];
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 4; ++j) {
        if (!h[i * 4 + j]) continue;
        long long d = h[i * 4 + j] - t0;
        if (d > mx[j]) mx[j] = d;
        if (d < mn[j]) mn[j] = d;
      }
    fprintf(stderr, "dense phases (ns since first CTA start): start %lld..%lld | main loop done %lld..%lld | staged+ticket %lld..%lld | last CTA end %lld\n",
            mn[0], mx[0], mn[1], mx[1], mn[2], mx[2], mx[3]);
    fprintf(stderr, "  last CTA: enter %lld | phase a %lld | phase b %lld | phase c %lld | phase d %lld\n", h[4 * grid] - t0,
            h[4 * grid + 1] - t0, h[4 * grid + 2] - t0, h[4 * grid + 3] - t0, h[4 * grid + 4] - t0);
    free(h);
    cudaFree(dbg);
  }
  return MGB_OK;
}

extern "C" int mgb_preagg_dense_batch_async(const mgb_agg_spec* spec, int32_t n_chunks, const int64_t* chunk_rows,
                                            const int64_t* const* const* chunk_key_cols,
                                            const void* const* const* chunk_val_cols, int64_t* const* out_key_cols,
                                            void* const* out_acc_cols, int64_t out_capacity, int64_t* out_n_dev,
                                            void* stream) {
  return dense_batch_impl(spec, nullptr, n_chunks, chunk_rows, chunk_key_cols, chunk_val_cols, out_key_cols,
                          out_acc_cols, out_capacity, out_n_dev, stream);
}

extern "C" int mgb_preagg_dense_batch_fused_async(const mgb_agg_spec* spec, const mgb_row_program* prog,
                                                  int32_t n_chunks, const int64_t* chunk_rows,
                                                  const int64_t* const* const* chunk_key_cols,
                                                  const void* const* const* chunk_phys_cols,
                                                  int64_t* const* out_key_cols, void* const* out_acc_cols,
                                                  int64_t out_capacity, int64_t* out_n_dev, void* stream) {
  MGB_REQUIRE(prog, MGB_ERR_INVALID, "null row program");
  return dense_batch_impl(spec, prog, n_chunks, chunk_rows, chunk_key_cols, chunk_phys_cols, out_key_cols,
                          out_acc_cols, out_capacity, out_n_dev, stream);
}

extern "C" int mgb_preagg_dense_async(const mgb_agg_spec* spec, const int64_t* const* key_cols,
                                      const void* const* val_cols, int64_t n_rows, int64_t* const* out_key_cols,
                                      void* const* out_acc_cols, int64_t out_capacity, int64_t* out_n_dev,
                                      void* stream) {
  MGB_REQUIRE(key_cols && val_cols, MGB_ERR_INVALID, "null pointer");
  return mgb_preagg_dense_batch_async(spec, 1, &n_rows, &key_cols, &val_cols, out_key_cols, out_acc_cols,
                                      out_capacity, out_n_dev, stream);
}

extern "C" int mgb_preagg_dense(const mgb_agg_spec* spec, const int64_t* const* key_cols,
                                const void* const* val_cols, int64_t n_rows, int64_t* const* out_key_cols,
                                void* const* out_acc_cols, int64_t out_capacity, int64_t* out_n_groups,
                                void* stream) {
  MGB_REQUIRE(out_n_groups, MGB_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  *out_n_groups = -1;
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  DenseScratch* s = nullptr;
  MGB_TRY(dense_scratch(dev, 0, 1, &s));
  MGB_TRY(mgb_preagg_dense_async(spec, key_cols, val_cols, n_rows, out_key_cols, out_acc_cols, out_capacity,
                                 (int64_t*)s->out_n, stream));
  MGB_CUDA(cudaMemcpyAsync(s->out_n_host, s->out_n, sizeof(long long), cudaMemcpyDeviceToHost, st));
  MGB_CUDA(cudaStreamSynchronize(st));
  *out_n_groups = s->out_n_host[0];
  return MGB_OK;
}

// --------------------------------------------------------------------------------------------------
// small merge: one CTA, no atomics on the values, no host synchronisation required
// --------------------------------------------------------------------------------------------------
struct SmallPiece {
  const uint64_t* key[MGB_MAX_KEYS];
  const long long* n_dev;   // device row count of the piece (output count of an earlier call), or nullptr
  int n;                    // host-known row count (used when n_dev == nullptr)
};

struct SmallPost {      // one result column of the stage=agg post functions (mgb_post_step, packed)
  uint8_t kind;         // mgb_post_kind
  uint8_t a0, a1, a2;   // merged accumulators it reads
  int8_t ddof;
  uint8_t pad[3];
};

struct SmallParams {
  int n_pieces;
  int out_cap;
  SmallPiece piece[SM_MAX_PIECES];
  const uint64_t* const* acc_ptrs;   // device table [piece][acc]
  OutCols out;                       // destination columns; op: MGB_SUM / MGB_MIN / MGB_MAX, dt: column dtype
  int sort;
  long long* out_n;                  // device: number of groups, -1: not a small input / an input had failed
  int n_post;                        // > 0: evaluate the post functions on the merged rows (stage=agg in one launch)
  SmallPost post[MGB_MAX_POST];
  uint64_t* post_out[MGB_MAX_POST];
};

__device__ __forceinline__ double small_as_f64(uint64_t bits, int dt) {
  return dt == MGB_F64 ? __longlong_as_double((long long)bits) : (double)(long long)bits;
}

template <int NK>
__global__ void __launch_bounds__(SM_THREADS, 1) k_merge_small(const SmallParams p) {
  extern __shared__ __align__(16) uint32_t ssm[];
  constexpr int MAXR = MGB_SMALL_MAX_ROWS;
  // layout (bytes): tab_keys 2*4096*8 | ukeys 2*2048*8 | tab_idx 4096*4 | item_idx, item_r, ucnt(+1), ufill, sorted,
  //                 rank: 2048*4 each | item_pc 2048 | ints
  uint64_t* tab_keys = (uint64_t*)ssm;
  uint64_t* ukeys = tab_keys + MGB_MAX_KEYS * SM_TAB;
  int* tab_idx = (int*)(ukeys + MGB_MAX_KEYS * MAXR);
  int* item_idx = tab_idx + SM_TAB;
  int* item_r = item_idx + MAXR;
  int* ucnt = item_r + MAXR;            // MAXR + 1 (+ pad)
  int* ufill = ucnt + MAXR + 2;
  int* sorted = ufill + MAXR;
  int* rank = sorted + MAXR;
  unsigned char* item_pc = (unsigned char*)(rank + MAXR);
  int* s_int = (int*)(item_pc + MAXR);
  int* n_unique = s_int;                // [0]
  int* s_total = s_int + 1;
  int* pstart = s_int + 2;              // [SM_MAX_PIECES + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int na = p.out.na;

  if (threadIdx.x == 0) {
    int total = 0;
    bool bad = false;
    for (int i = 0; i < p.n_pieces; ++i) {
      long long n = p.piece[i].n_dev ? *(volatile const long long*)p.piece[i].n_dev : (long long)p.piece[i].n;
      if (n < 0) bad = true;
      pstart[i] = total;
      total += n > 0 ? (int)(n > MAXR + 1 ? MAXR + 1 : n) : 0;
      if (total > MAXR) bad = true;
    }
    pstart[p.n_pieces] = total;
    if (bad || total > p.out_cap) total = -1;
    *s_total = total;
    *n_unique = 0;
  }
  __syncthreads();
  const int total = *s_total;
  if (total < 0) {
    if (threadIdx.x == 0) *p.out_n = -1;
    return;
  }
  // hash table sized to the input: 2 * total rounded up to a power of two (>= 64)
  int tsize = 64;
  while (tsize < 2 * total) tsize <<= 1;
  const uint32_t tmask = (uint32_t)tsize - 1u;
  for (int i = threadIdx.x; i < tsize; i += SM_THREADS) tab_idx[i] = -1;
  __syncthreads();
  // pass 1: output row of every input row
  for (int it0 = 0; it0 < total; it0 += SM_THREADS) {   // uniform trip count: the insert votes per warp
    const int it = it0 + (int)threadIdx.x;
    const bool valid = it < total;
    int pc = 0, r = 0;
    uint64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = 0;
    if (valid) {
      while (it >= pstart[pc + 1]) ++pc;
      r = it - pstart[pc];
#pragma unroll
      for (int j = 0; j < NK; ++j) k[j] = ld_cg(p.piece[pc].key[j] + r);
    }
    const int idx = small_find_or_insert<NK>(tab_keys, tab_idx, n_unique, valid, k, tmask);
    if (valid) {
      item_idx[it] = idx;
      item_r[it] = r;
      item_pc[it] = (unsigned char)pc;
#pragma unroll
      for (int j = 0; j < NK; ++j) ukeys[j * MAXR + idx] = k[j];   // same value from every writer
    }
  }
  __syncthreads();
  const int nu = *n_unique;
  // output position of every unique key: rank by key tuple when sorting (keys are unique), else table order.
  // More than a handful of keys are ranked by a bitonic sort of their indices in shared memory (`sorted` is
  // free until the counting sort below): O(n log^2 n) instead of the n^2 comparisons of the direct ranking.
  const bool big_sort = p.sort && nu > 96;
  if (big_sort) {
    int n2 = 128;
    while (n2 < nu) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += SM_THREADS) sorted[i] = i < nu ? i : -1;   // -1: +infinity padding
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < n2; i += SM_THREADS) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const int a = sorted[i], b = sorted[ixj];
            bool b_less_a;   // is entry b strictly smaller than entry a ?
            if (b < 0)
              b_less_a = false;
            else if (a < 0)
              b_less_a = true;
            else {
              uint64_t ka[NK], kb2[NK];
#pragma unroll
              for (int q = 0; q < NK; ++q) {
                ka[q] = ukeys[q * MAXR + a];
                kb2[q] = ukeys[q * MAXR + b];
              }
              b_less_a = key_less<NK>(kb2, ka);
            }
            const bool up = (i & k) == 0;
            if (b_less_a == up) {
              sorted[i] = b;
              sorted[ixj] = a;
            }
          }
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < nu; i += SM_THREADS) rank[sorted[i]] = i;
    __syncthreads();
  }
  for (int u = threadIdx.x; u < nu; u += SM_THREADS) {
    int pos = u;
    if (big_sort) {
      pos = rank[u];
    } else if (p.sort) {
      uint64_t k[NK];
#pragma unroll
      for (int j = 0; j < NK; ++j) k[j] = ukeys[j * MAXR + u];
      pos = 0;
      for (int q = 0; q < nu; ++q) {
        uint64_t o[NK];
#pragma unroll
        for (int j = 0; j < NK; ++j) o[j] = ukeys[j * MAXR + q];
        pos += key_less<NK>(o, k) ? 1 : 0;
      }
    }
    rank[u] = pos;
#pragma unroll
    for (int j = 0; j < NK; ++j) p.out.key[j][pos] = ukeys[j * MAXR + u];
  }
  // counting sort of the input rows by output row
  for (int i = threadIdx.x; i <= nu; i += SM_THREADS) ucnt[i] = 0;
  for (int i = threadIdx.x; i < nu; i += SM_THREADS) ufill[i] = 0;
  __syncthreads();
  for (int it = threadIdx.x; it < total; it += SM_THREADS) atomicAdd(&ucnt[item_idx[it]], 1);
  __syncthreads();
  if (warp == 0) {
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
  for (int it = threadIdx.x; it < total; it += SM_THREADS) {
    const int u = item_idx[it];
    sorted[ucnt[u] + atomicAdd(&ufill[u], 1)] = it;
  }
  __syncthreads();
  // L lanes per (output row, accumulator): contributions straight from L2, combined by shuffles.  L = 8 when the
  // output rows have several contributions each (a tree merge of many pieces), L = 1 when there are about as many
  // output rows as input rows (one band's partials with hundreds of groups: the loop is bound by the latency of
  // its dependent loads, so every lane takes a pair of its own)
  const int L = total >= 3 * nu ? 8 : 1;
  const int sl = lane & (L - 1), pl = lane / L, ppw = 32 / L;
  constexpr int WARPS = SM_THREADS / 32;
  for (int pair0 = 0; pair0 < nu * na; pair0 += WARPS * ppw) {
    const int pair = pair0 + warp * ppw + pl;
    const bool on = pair < nu * na;
    const int u = on ? pair / na : 0, a = on ? pair - u * na : 0;
    const int op = p.out.op[a];
    const bool f = p.out.dt[a] == MGB_F64;
    const int kind = op == MGB_MIN ? 3 : op == MGB_MAX ? 4 : 1;
    // identity: 0 for sums / counts; NaN (ignored by fmin / fmax) for float min / max; extremes for int
    uint64_t v = 0;
    if (kind == 3) v = f ? 0x7ff8000000000000ULL : 0x7fffffffffffffffULL;
    if (kind == 4) v = f ? 0x7ff8000000000000ULL : 0x8000000000000000ULL;
    const int q1 = on ? ucnt[u + 1] : 0;
#pragma unroll 4
    for (int q = (on ? ucnt[u] : 0) + sl; q < q1; q += L) {
      const int it = sorted[q];
      const uint64_t* col = p.acc_ptrs[item_pc[it] * na + a];
      uint64_t y = ld_cg(col + item_r[it]);
      // a NaN partial SUM (inf - inf inside one chunk) is skipped, as the reference's combine / agg stages do: they
      // re-aggregate the partial frames with pandas' skipna sum (aggregation.py:1130-1164, 1023-1032)
      if (f && kind == 1 && mgb_is_nan_bits(y)) y = 0;
      v = dense_combine(kind, f, v, y);
    }
    for (int o = 1; o < L; o <<= 1) v = dense_combine(kind, f, v, __shfl_xor_sync(0xffffffffu, v, o));
    if (on && sl == 0) p.out.acc[a][rank[u]] = v;
  }
  if (p.n_post > 0) {
    // a10: the generated post functions of the reference on the merged rows, in its operation order and
    // without FMA contraction (mars/dataframe/reduction/mean.py:26-33, var.py:38-48)
    __syncthreads();   // the merged accumulators (global, written by this CTA) are visible to all its threads
    for (int i = threadIdx.x; i < nu * p.n_post; i += SM_THREADS) {
      const int j = i / nu, r = i - j * nu;
      const SmallPost ps = p.post[j];
      uint64_t v = p.out.acc[ps.a0][r];
      if (ps.kind == MGB_POST_MEAN) {
        const double s_ = small_as_f64(v, p.out.dt[ps.a0]);
        const double c_ = (double)(long long)p.out.acc[ps.a1][r];
        v = (uint64_t)__double_as_longlong(__ddiv_rn(s_, c_));
      } else if (ps.kind == MGB_POST_VAR) {
        const double s_ = small_as_f64(v, p.out.dt[ps.a0]);
        const double q_ = small_as_f64(p.out.acc[ps.a1][r], p.out.dt[ps.a1]);
        const double c_ = (double)(long long)p.out.acc[ps.a2][r];
        double res;
        if (ps.ddof == 0) {
          const double m2 = __ddiv_rn(q_, c_), m = __ddiv_rn(s_, c_);
          res = __dsub_rn(m2, __dmul_rn(m, m));
        } else {
          const double t_ = __ddiv_rn(__dmul_rn(s_, s_), c_);
          res = __ddiv_rn(__dsub_rn(q_, t_), __dsub_rn(c_, (double)ps.ddof));
        }
        v = (uint64_t)__double_as_longlong(res);
      }
      p.post_out[j][r] = v;
    }
  }
  if (threadIdx.x == 0) *p.out_n = nu;
}

struct SmallScratch {
  const uint64_t** ptrs_dev = nullptr;
  const uint64_t** ptrs_host = nullptr;   // pinned ring: SLOTS tables of SM_MAX_PIECES * MGB_MAX_ACCS pointers
  cudaEvent_t ev[64] = {nullptr};         // completion of the launch that last used a slot
  bool used[64] = {false};
  int slot = 0;
  long long* out_n = nullptr;             // scratch count for the synchronous entry point
  long long* out_n_host = nullptr;
};
#define SMALL_SLOTS 64
static thread_local SmallScratch g_ss[8];

static size_t small_smem_bytes() {
  return (size_t)MGB_MAX_KEYS * SM_TAB * 8 + (size_t)MGB_MAX_KEYS * MGB_SMALL_MAX_ROWS * 8 + SM_TAB * 4 +
         6 * ((size_t)MGB_SMALL_MAX_ROWS + 2) * 4 + MGB_SMALL_MAX_ROWS + (SM_MAX_PIECES + 8) * 4 + 64;
}

static int small_scratch(int dev, SmallScratch** out) {
  MGB_REQUIRE(dev >= 0 && dev < 8, MGB_ERR_UNSUPPORTED, "device ordinal %d", dev);
  SmallScratch& s = g_ss[dev];
  const size_t table_words = (size_t)SM_MAX_PIECES * MGB_MAX_ACCS;
  if (!s.ptrs_dev) {
    MGB_CUDA(cudaMalloc((void**)&s.ptrs_dev, table_words * 8 * SMALL_SLOTS));
    MGB_CUDA(cudaMallocHost((void**)&s.ptrs_host, table_words * 8 * SMALL_SLOTS));
    MGB_CUDA(cudaMalloc((void**)&s.out_n, 16));
    MGB_CUDA(cudaMallocHost((void**)&s.out_n_host, 16));
    for (int i = 0; i < SMALL_SLOTS; ++i) MGB_CUDA(cudaEventCreateWithFlags(&s.ev[i], cudaEventDisableTiming));
    const size_t smem = small_smem_bytes();
    MGB_CUDA(cudaFuncSetAttribute(k_merge_small<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MGB_CUDA(cudaFuncSetAttribute(k_merge_small<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  *out = &s;
  return MGB_OK;
}

static int merge_small_impl(const mgb_agg_spec* spec, int32_t n_pieces, const int64_t* piece_rows,
                            const int64_t* const* piece_rows_dev, const int64_t* const* const* piece_key_cols,
                            const void* const* const* piece_acc_cols, int64_t* const* out_key_cols,
                            void* const* out_acc_cols, int64_t out_capacity, int32_t sort_keys,
                            const mgb_post_step* post, int32_t n_post, void* const* out_post_cols,
                            int64_t* out_n_dev, void* stream) {
  MGB_REQUIRE(spec && piece_rows && piece_key_cols && piece_acc_cols && out_key_cols && out_acc_cols && out_n_dev,
              MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(n_post >= 0 && n_post <= MGB_MAX_POST && (n_post == 0 || (post && out_post_cols)), MGB_ERR_INVALID,
              "bad post steps (%d, at most %d)", n_post, MGB_MAX_POST);
  cudaStream_t st = (cudaStream_t)stream;
  const int nk = spec->n_keys, na = spec->n_accs;
  MGB_REQUIRE(nk >= 1 && nk <= MGB_MAX_KEYS && na >= 1 && na <= MGB_MAX_ACCS, MGB_ERR_UNSUPPORTED,
              "spec outside the supported envelope");
  MGB_REQUIRE(out_capacity >= 0, MGB_ERR_INVALID, "negative capacity");
  // pieces that are known to be empty are dropped; deferred ones (piece_rows < 0) must carry a device count
  long long known = 0;
  int used = 0;
  for (int i = 0; i < n_pieces; ++i) {
    if (piece_rows[i] < 0) {
      MGB_REQUIRE(piece_rows_dev && piece_rows_dev[i], MGB_ERR_INVALID, "piece %d: deferred size without a device count", i);
      ++used;
    } else if (piece_rows[i] > 0) {
      known += piece_rows[i];
      ++used;
    }
  }
  if (known > MGB_SMALL_MAX_ROWS || known > out_capacity || used > SM_MAX_PIECES) {   // not small: tell the caller
    MGB_CUDA(cudaMemsetAsync(out_n_dev, 0xFF, sizeof(int64_t), st));
    return MGB_OK;
  }
  if (used == 0) {
    MGB_CUDA(cudaMemsetAsync(out_n_dev, 0, sizeof(int64_t), st));
    return MGB_OK;
  }
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  SmallScratch* sp = nullptr;
  MGB_TRY(small_scratch(dev, &sp));
  SmallScratch& s = *sp;
  const size_t table_words = (size_t)SM_MAX_PIECES * MGB_MAX_ACCS;
  SmallParams p;
  memset(&p, 0, sizeof(p));
  const int slot = s.slot;
  s.slot = (s.slot + 1) % SMALL_SLOTS;
  if (s.used[slot]) MGB_CUDA(cudaEventSynchronize(s.ev[slot]));   // the launch that read this slot is done
  const uint64_t** host_tab = s.ptrs_host + (size_t)slot * table_words;
  int q = 0;
  for (int i = 0; i < n_pieces; ++i) {
    if (piece_rows[i] == 0) continue;
    for (int j = 0; j < nk; ++j) {
      MGB_REQUIRE(piece_key_cols[i] && piece_key_cols[i][j], MGB_ERR_INVALID, "null key column");
      p.piece[q].key[j] = (const uint64_t*)piece_key_cols[i][j];
    }
    p.piece[q].n = piece_rows[i] > 0 ? (int)piece_rows[i] : 0;
    p.piece[q].n_dev = piece_rows[i] < 0 ? (const long long*)piece_rows_dev[i] : nullptr;
    for (int a = 0; a < na; ++a) {
      MGB_REQUIRE(piece_acc_cols[i] && piece_acc_cols[i][a], MGB_ERR_INVALID, "null accumulator column");
      host_tab[q * na + a] = (const uint64_t*)piece_acc_cols[i][a];
    }
    ++q;
  }
  p.n_pieces = q;
  p.out_cap = (int)(out_capacity > MGB_SMALL_MAX_ROWS ? MGB_SMALL_MAX_ROWS : out_capacity);
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
  p.sort = sort_keys ? 1 : 0;
  p.out_n = (long long*)out_n_dev;
  p.n_post = n_post;
  for (int j = 0; j < n_post; ++j) {
    const mgb_post_step& q = post[j];
    const int need = q.kind == MGB_POST_IDENTITY ? 1 : q.kind == MGB_POST_MEAN ? 2 : 3;
    MGB_REQUIRE(q.kind >= MGB_POST_IDENTITY && q.kind <= MGB_POST_VAR && out_post_cols[j], MGB_ERR_INVALID,
                "bad post step %d", j);
    const int idx[3] = {q.a0, q.a1, q.a2};
    for (int t = 0; t < need; ++t)
      MGB_REQUIRE(idx[t] >= 0 && idx[t] < na, MGB_ERR_INVALID, "post step %d reads accumulator %d of %d", j, idx[t], na);
    p.post[j].kind = (uint8_t)q.kind;
    p.post[j].a0 = (uint8_t)q.a0;
    p.post[j].a1 = (uint8_t)(need > 1 ? q.a1 : 0);
    p.post[j].a2 = (uint8_t)(need > 2 ? q.a2 : 0);
    p.post[j].ddof = (int8_t)q.ddof;
    p.post_out[j] = (uint64_t*)out_post_cols[j];
  }
  {
    MgbProfScope prof(MGB_K_MERGE, st);
    const size_t smem = small_smem_bytes();
    if (nk == 1)
      k_merge_small<1><<<1, SM_THREADS, smem, st>>>(p);
    else
      k_merge_small<2><<<1, SM_THREADS, smem, st>>>(p);
    MGB_CHECK_LAUNCH();
  }
  MGB_CUDA(cudaEventRecord(s.ev[slot], st));
  s.used[slot] = true;
  return MGB_OK;
}

extern "C" int mgb_merge_small_async(const mgb_agg_spec* spec, int32_t n_pieces, const int64_t* piece_rows,
                                     const int64_t* const* piece_rows_dev,
                                     const int64_t* const* const* piece_key_cols,
                                     const void* const* const* piece_acc_cols, int64_t* const* out_key_cols,
                                     void* const* out_acc_cols, int64_t out_capacity, int32_t sort_keys,
                                     int64_t* out_n_dev, void* stream) {
  return merge_small_impl(spec, n_pieces, piece_rows, piece_rows_dev, piece_key_cols, piece_acc_cols, out_key_cols,
                          out_acc_cols, out_capacity, sort_keys, nullptr, 0, nullptr, out_n_dev, stream);
}

extern "C" int mgb_merge_small_post_async(const mgb_agg_spec* spec, int32_t n_pieces, const int64_t* piece_rows,
                                          const int64_t* const* piece_rows_dev,
                                          const int64_t* const* const* piece_key_cols,
                                          const void* const* const* piece_acc_cols, int64_t* const* out_key_cols,
                                          void* const* out_acc_cols, int64_t out_capacity, int32_t sort_keys,
                                          const mgb_post_step* post, int32_t n_post, void* const* out_post_cols,
                                          int64_t* out_n_dev, void* stream) {
  return merge_small_impl(spec, n_pieces, piece_rows, piece_rows_dev, piece_key_cols, piece_acc_cols, out_key_cols,
                          out_acc_cols, out_capacity, sort_keys, post, n_post, out_post_cols, out_n_dev, stream);
}

extern "C" int mgb_merge_small(const mgb_agg_spec* spec, int32_t n_pieces, const int64_t* piece_rows,
                               const int64_t* const* const* piece_key_cols,
                               const void* const* const* piece_acc_cols, int64_t* const* out_key_cols,
                               void* const* out_acc_cols, int64_t out_capacity, int32_t sort_keys,
                               int64_t* out_n_groups, void* stream) {
  MGB_REQUIRE(out_n_groups && piece_rows, MGB_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  *out_n_groups = -1;
  long long total = 0;
  for (int i = 0; i < n_pieces; ++i) {
    MGB_REQUIRE(piece_rows[i] >= 0, MGB_ERR_INVALID, "negative piece size");
    total += piece_rows[i];
  }
  MGB_REQUIRE(out_capacity >= total || total > MGB_SMALL_MAX_ROWS, MGB_ERR_INVALID,
              "output capacity %lld < %lld rows", (long long)out_capacity, total);
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  SmallScratch* sp = nullptr;
  MGB_TRY(small_scratch(dev, &sp));
  MGB_TRY(mgb_merge_small_async(spec, n_pieces, piece_rows, nullptr, piece_key_cols, piece_acc_cols, out_key_cols,
                                out_acc_cols, out_capacity, sort_keys, (int64_t*)sp->out_n, stream));
  MGB_CUDA(cudaMemcpyAsync(sp->out_n_host, sp->out_n, sizeof(long long), cudaMemcpyDeviceToHost, st));
  MGB_CUDA(cudaStreamSynchronize(st));
  *out_n_groups = sp->out_n_host[0];
  return MGB_OK;
}
