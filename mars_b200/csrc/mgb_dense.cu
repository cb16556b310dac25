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

int mgb_dense_launch_1_256(const DenseParams&, int, bool, int, size_t, cudaStream_t);
int mgb_dense_launch_1_512(const DenseParams&, int, bool, int, size_t, cudaStream_t);
int mgb_dense_launch_2_256(const DenseParams&, int, bool, int, size_t, cudaStream_t);
int mgb_dense_launch_2_512(const DenseParams&, int, bool, int, size_t, cudaStream_t);

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
    if (spec->val_dtype[c] == MGB_F64) p.f64_mask |= 1u << ncol;
    for (int o = MGB_SUM; o <= MGB_MAX; ++o)
      if (m & OPB(o)) p.op_mask[o] |= 1u << ncol;
    if (m & (OPB(MGB_SUMSQ) | OPB(MGB_MIN) | OPB(MGB_MAX))) allops = true;
    ++ncol;
  }
  for (int j = 0; j < nk; ++j) {
    MGB_REQUIRE(key_cols[j] && out_key_cols[j], MGB_ERR_INVALID, "null key column %d", j);
    p.keys[j] = (const uint64_t*)key_cols[j];
    p.out.key[j] = (uint64_t*)out_key_cols[j];
  }
  p.n = n_rows;
  if (ncol == 0) return MGB_OK;   // nothing but counts of int64 columns: leave it to the general path
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
    if (p.op_mask[kind_op[kind]]) {
      off += 64 * ncol;
      ++w;
    }
  }
  p.off_nan = off;
  p.stride_g = p.off_nan + 32;   // a multiple of 32 words: the bank of a lane-private slot depends on the lane only
  p.wpc = w;
  const size_t fixed = (size_t)(MGB_MAX_KEYS * DN_TAB + MGB_MAX_KEYS * DN_MAX_GC) * 8 + DN_TAB * 4 + 64;
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
  if (GC < 2) return MGB_OK;   // accumulators of a single group do not fit: general path
  p.GC = GC;
  const int n_warps = threads / 32;
  const long long tile_rows = 1024;
  const long long ntiles = (n_rows + tile_rows - 1) / tile_rows;
  const int sms = mgb_sm_count();
  const int grid = (int)(ntiles < sms ? ntiles : sms);
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
  // the CTA owns its SM: take (almost) all shared memory; the last CTA re-uses it for the cross-CTA merge
  const size_t acc_smem = (size_t)n_warps * GC * p.stride_g * 4 + fixed;
  size_t smem = acc_smem < 112 * 1024 ? 112 * 1024 : acc_smem;   // >= 112 KB for the cross-CTA merge tables
  // last-CTA merge: key table + 5 int arrays of at most grid * GC (+2) entries
  const size_t merge_smem = (size_t)MGB_MAX_KEYS * SM_TAB * 8 + SM_TAB * 4 + 32 + 5 * ((size_t)grid * GC + 4) * 4;
  MGB_REQUIRE(acc_smem <= smem && merge_smem <= smem, MGB_ERR_OVERFLOW, "dense path shared-memory plan does not fit");
  static const char* env_dbg = getenv("MGB_DENSE_DEBUG");   // experiments only: per-CTA phase timestamps
  long long* dbg = nullptr;
  if (env_dbg) {
    MGB_CUDA(cudaMalloc((void**)&dbg, sizeof(long long) * (4 * grid + 8)));
    MGB_CUDA(cudaMemset(dbg, 0, sizeof(long long) * (4 * grid + 8)));
    p.dbg = dbg;
  }
  {
    MgbProfScope prof(MGB_K_PREAGG_TINY, st);
    int status;
    if (nk == 1)
      status = threads == 512 ? mgb_dense_launch_1_512(p, ncol, plain, grid, smem, st)
                              : mgb_dense_launch_1_256(p, ncol, plain, grid, smem, st);
    else
      status = threads == 512 ? mgb_dense_launch_2_512(p, ncol, plain, grid, smem, st)
                              : mgb_dense_launch_2_256(p, ncol, plain, grid, smem, st);
    MGB_TRY(status);
  }
  MGB_CUDA(cudaMemcpyAsync(s->out_n_host, s->out_n, sizeof(long long), cudaMemcpyDeviceToHost, st));
  MGB_CUDA(cudaStreamSynchronize(st));
  *out_n_groups = s->out_n_host[0];
  if (dbg) {
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
    fprintf(stderr, "  last CTA: enter %lld | table init %lld | items %lld | pass1 %lld | staged copy %lld\n", h[4 * grid] - t0,
            h[4 * grid + 1] - t0, h[4 * grid + 2] - t0, h[4 * grid + 3] - t0, h[4 * grid + 4] - t0);
    free(h);
    cudaFree(dbg);
  }
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
