/*
 * mgb.h -- C ABI of libmarsgb.so: the B200-native (sm_100a) groupby-aggregation hot path of Mars.
 *
 * This is the drop-in boundary.  The reference has no native code on this path:
 * its operands' `execute(ctx, op)` classmethods call pandas.  Every entry point below names the reference
 * call site (paths relative to the reference checkout) whose arithmetic it replaces.  The Python executors
 * in mars_b200/executors.py bind these symbols with ctypes and are registered through the reference's own
 * plug-in seam `Operand.register_executor` (mars/core/operand/core.py:458-464).
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers unless a name ends in `_host`.
 *   - every column is a dense array of 8-byte elements (int64 keys; float64 or int64 values).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - every function returns an mgb_status; nothing calls exit()/abort(); mgb_last_error() returns a
 *     thread-local message for the last non-OK status (reference error contract: raise -> ExecutionError,
 *     mars/services/subtask/worker/processor.py:210-214).
 *   - the caller owns all input/output buffers (torch allocator); scratch memory is taken from the CUDA
 *     stream-ordered pool (cudaMallocAsync) on the given stream.
 */
#ifndef MGB_H_
#define MGB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_VERSION 100

typedef enum {
  MGB_OK = 0,
  MGB_ERR_INVALID = 1,     /* bad argument */
  MGB_ERR_CUDA = 2,        /* CUDA runtime error (message has cudaGetErrorString) */
  MGB_ERR_UNSUPPORTED = 3, /* valid request outside the implemented envelope */
  MGB_ERR_NCCL = 4,
  MGB_ERR_OVERFLOW = 5,    /* an internal capacity bound was exceeded (bug or corrupted input) */
  MGB_ERR_STATE = 6        /* call order violated (e.g. emit before finalize) */
} mgb_status;

typedef enum { MGB_F64 = 0, MGB_I64 = 1 } mgb_dtype;

/* Atomic reductions of the groupby map/combine/agg stages.  The reference decomposes user functions into
 * these steps in ReductionCompiler (mars/dataframe/reduction/core.py:976-1031):
 *   sum -> map sum / agg sum;  count -> map count / agg SUM (of int64 counts);  min, max -> same;
 *   mean = sum/count;  var = sum, sum(x**2), count.
 * NaN handling follows pandas skipna=True: NaN values are skipped by SUM/SUMSQ/MIN/MAX and not counted by
 * COUNT; an all-NaN group yields sum 0.0, count 0, min/max NaN.                                          */
typedef enum { MGB_SUM = 0, MGB_SUMSQ = 1, MGB_COUNT = 2, MGB_MIN = 3, MGB_MAX = 4 } mgb_op;

#define MGB_MAX_KEYS 2
#define MGB_MAX_VALS 64
#define MGB_MAX_ACCS 64

typedef struct {
  int32_t n_keys;                  /* 1..MGB_MAX_KEYS int64 key columns                               */
  int32_t n_vals;                  /* 0..MGB_MAX_VALS value columns                                   */
  int32_t n_accs;                  /* 1..MGB_MAX_ACCS accumulators                                    */
  int32_t val_dtype[MGB_MAX_VALS]; /* mgb_dtype of each value column                                  */
  int32_t acc_col[MGB_MAX_ACCS];   /* value column each accumulator reads                             */
  int32_t acc_op[MGB_MAX_ACCS];    /* mgb_op; output dtype: COUNT -> I64, otherwise the column dtype  */
} mgb_agg_spec;

const char* mgb_last_error(void);
int mgb_version(void);
/* Device sanity probe: returns MGB_OK and fills sm_count / cc (major*10+minor) of the current device. */
int mgb_device_info(int32_t* sm_count, int32_t* cc);

/* ---- a5: partition assignment --------------------------------------------------------------------
 * Replaces `pd.util.hash_pandas_object(idx, categorize=False) % size` in hash_dataframe_on
 * (mars/dataframe/utils.py:77-101) as used by DataFrameGroupByOperand.execute_map
 * (mars/dataframe/groupby/core.py:353-435).  Bit-exact: one int64 key -> pandas `_hash_ndarray`
 * (splitmix64 finaliser on the uint64 view); several keys -> `hash_tuples`/`combine_hash_arrays`.      */
int mgb_hash_keys(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows, uint64_t* out_hash,
                  void* stream);
int mgb_partition_ids(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows, int64_t n_partitions,
                      int32_t* out_pid, void* stream);

/* The same for shuffles keyed on COLUMNS of a frame -- `hash_dataframe_on(df, on=[...], size)` as used by the
 * plain groupby (mars/dataframe/groupby/core.py:353-435, `on = by`), the merge shuffle
 * (mars/dataframe/merge/merge.py:93-128) and drop_duplicates (mars/dataframe/base/drop_duplicates.py:150-168):
 * pandas hashes the rows of `df[on]` with hash_pandas_object(index=False), which combines the per-column hashes
 * also when there is ONE column (an Index of one level is hashed directly: mgb_partition_ids).              */
int mgb_partition_ids_frame(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows, int64_t n_partitions,
                            int32_t* out_pid, void* stream);

/* sort=True shuffle (PSRS): range partition by pivots.  Replaces the label slices
 * `in_df.loc[pivots[i-1]:pivots[i]].drop(index=pivots[i-1])` of DataFrameGroupbySortShuffle._execute_map
 * (mars/dataframe/groupby/sort.py:113-135): out_pid = number of pivot tuples strictly smaller than the key
 * tuple (pivots ascending; signed int64, lexicographic).  pivot_cols are device arrays of n_pivots entries. */
int mgb_range_partition_ids(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows,
                            const int64_t* const* pivot_cols, int32_t n_pivots, int32_t* out_pid, void* stream);

/* ---- a5/a6: split partial rows into per-reducer blocks --------------------------------------------
 * Replaces `RangeIndex.groupby(hash % size)` + the K `iloc[f_i]` takes (groupby/core.py:389-435).
 * Stable: inside every partition rows keep their input order (== ascending positions f_i).
 * out_cols[c][out_offsets[p] .. out_offsets[p+1]) holds partition p of column c; out_offsets has
 * n_partitions+1 int64 entries on the device.  1 <= n_partitions <= 65536.                              */
int mgb_partition_scatter(const int32_t* pid, int64_t n_rows, int32_t n_partitions,
                          const void* const* in_cols, void* const* out_cols, int32_t n_cols,
                          int64_t* out_offsets, void* stream);

/* a5 in one call, without exposing the partition ids and without any host synchronisation: partition id =
 * pandas hash of the key tuple % n_partitions (as mgb_partition_ids), stable split of the rows of in_cols into
 * per-reducer blocks of out_cols (as mgb_partition_scatter), out_offsets[n_partitions + 1] on the device.  For
 * n_partitions <= 256 every row is scattered straight to its destination (one pass over the columns, scattered
 * 8-byte writes combined in L2) instead of being gathered through a permutation (random 8-byte reads).      */
int mgb_hash_partition(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows, int32_t n_partitions,
                       const void* const* in_cols, void* const* out_cols, int32_t n_cols, int64_t* out_offsets,
                       void* stream);

/* The partition kernel in its two passes (north_star: per-destination histogram -> prefix scan -> coalesced 128-bit
 * vectorised scatter), for 1 <= n_partitions <= 256.  Same split as mgb_hash_partition; the passes are separate so
 * that a caller can learn the block sizes first (out_offsets), decide where every block goes, and then let the
 * scatter write each row ONCE straight to that place -- e.g. into the receive arena of the GPU that runs the reducer,
 * mapped through CUDA IPC (mgb_ipc_arena_slot / mgb_ipc_open_slot): partition + pack + transfer of the shuffle in one
 * kernel (replaces execute_map's block takes, groupby/core.py:389-435, AND the storage-service transfer of the
 * blocks, mars/services/storage/transfer.py:63-206).
 *   mode 0: partition id = pandas hash of the key tuple, Index semantics (mgb_partition_ids) % n_partitions
 *   mode 1: DataFrame-row semantics (mgb_partition_ids_frame);  mode 2: ids given in `pid` (clamped to range)
 *   mode | MGB_SPLIT_LONG_RUNS (all three calls alike): the destinations are peer GPUs -- 8192-row tiles (one CTA of
 *                     1024 threads per SM) instead of 2048-row tiles, i.e. four times longer store runs, which is
 *                     what NVLink wants (64 partitions: 453 -> 547 GB/s per GPU); also taken for >= 128 partitions
 *   mgb_split_tiles   number of row tiles T of a chunk; tile_hist is device scratch of n_partitions * T int64 that
 *                     must stay untouched between the two passes
 *   mgb_split_plan    pass 1: tile histograms, scanned; out_offsets[n_partitions + 1] on the device
 *   mgb_split_scatter pass 2: column tiles staged in shared memory by TMA (cp.async.bulk + mbarrier), permuted into
 *                     partition order on chip, written out in runs with 16-byte stores.  Destination of row j of
 *                     partition p in column c: out_cols[c] + out_offsets[p] + j when dst_table is NULL, else
 *                     (uint64_t*)dst_table[c * n_partitions + p] + j (device array of addresses; peer memory allowed;
 *                     16-byte aligned addresses get 16-byte stores).  Stable.  No host synchronisation.            */
#define MGB_SPLIT_LONG_RUNS 4
int mgb_split_tiles(int64_t n_rows, int32_t n_partitions, int32_t mode, int64_t* out_tiles);
int mgb_split_plan(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid, int64_t n_rows,
                   int32_t n_partitions, int64_t* tile_hist, int64_t* out_offsets, void* stream);
int mgb_split_scatter(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid, int64_t n_rows,
                      int32_t n_partitions, const int64_t* tile_hist, const void* const* in_cols, int32_t n_cols,
                      void* const* out_cols, const uint64_t* dst_table, void* stream);

/* ---- a4/a9/a10: hash aggregation -------------------------------------------------------------------
 * Replaces `df.groupby(keys).agg(...)` of DataFrameGroupByAgg._execute_map / _execute_combine /
 * _execute_agg (mars/dataframe/groupby/aggregation.py:1043-1267, pandas group_sum/min/max/count).
 * One aggregator consumes any number of row batches (chunks) and yields one row per distinct key tuple.
 *   create -> update* -> finalize -> emit -> destroy
 * update() is asynchronous on `stream`; the batch's column buffers must stay alive until finalize()
 * returns.  finalize() synchronises the stream once and reports the group count so the caller can
 * allocate the output columns.  emit() writes key columns (int64) and accumulator columns (8-byte, dtype
 * per spec) in table order; with sort_keys != 0 rows are ordered by the key tuple ascending
 * (groupby(sort=True)).                                                                                 */
typedef struct mgb_agg mgb_agg;
int mgb_agg_create(mgb_agg** out, const mgb_agg_spec* spec, void* stream);
/* Optional, before the first update: the caller expects that pre-aggregating a batch on its own does not
 * pay (an earlier chunk of the same computation spilled most of its rows; or the batches are partial frames
 * with unique keys, as on the reducer side): the shared-memory pre-aggregation pass is skipped and finalize
 * aggregates all batches in place with the bucketed aggregation (rows split by key hash into buckets of a
 * few hundred rows, one CTA folds a bucket in registers, no atomics on the values); inputs it declines (see
 * mgb_agg_stats) go row by row to the global hash table.                                                 */
int mgb_agg_set_hint(mgb_agg* agg, int32_t expect_distinct_keys);
int mgb_agg_update(mgb_agg* agg, const int64_t* const* key_cols, const void* const* val_cols,
                   int64_t n_rows);
int mgb_agg_finalize(mgb_agg* agg, int64_t* out_n_groups);
int mgb_agg_emit(mgb_agg* agg, int64_t* const* out_key_cols, void* const* out_acc_cols, int32_t sort_keys);
int mgb_agg_destroy(mgb_agg* agg);
/* Diagnostics of the last finalize (which path consumed the rows): stats[0]=why the bucketed aggregation
 * declined an input given under the distinct-keys hint (0 = it did not, 1 = a bucket overflowed its on-chip
 * table, 2 = skewed keys: one bucket far longer than the average), [1]=rows pre-aggregated by the
 * shared-memory hash table kernels, [2]=rows aggregated as raw rows (spilled, or all of them under the
 * distinct-keys hint), [3]=partial rows merged, [4]=global table capacity (0: the bucketed aggregation
 * produced the result, no table was built), [5]=kernels launched by this aggregator so far.               */
int mgb_agg_stats(mgb_agg* agg, int64_t* stats6);

/* Cardinality probe for the choice between the two map-side paths above: hashes a strided sample (at most
 * 16384 rows) of the batch's key tuples into an on-chip bitmap and reports how many rows were sampled and how
 * many distinct hashes they had.  One tiny launch + one 16-byte device->host read (synchronises the stream). */
int mgb_estimate_distinct(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows, int64_t* out_sampled,
                          int64_t* out_distinct, void* stream);


/* Low-cardinality fast paths (one kernel launch + one 8-byte device->host read each).
 *
 * mgb_preagg_dense: stage=map pre-aggregation (aggregation.py:1043-1128) of ONE row batch whose distinct key
 * tuples fit on chip (TPC-H Q1 shape).  Accumulators live in shared memory, privatised per warp and per
 * lane; the last CTA merges the per-CTA partial rows.  Output columns need capacity mgb_dense_capacity()
 * rows.  *out_n_groups = -1 means "not a dense shape" (too many groups / live columns): nothing was written
 * and the caller uses mgb_agg_*.  int64 columns that are only counted are not read at all.
 *
 * mgb_merge_small: stage=combine / stage=agg re-aggregation (aggregation.py:1130-1267) of small partial
 * tuples WITHOUT the concat copy of merge/concat.py:339-348: n_pieces partial blocks (piece i: piece_rows[i]
 * rows, key columns piece_key_cols[i][0..n_keys), accumulator columns piece_acc_cols[i][0..n_accs)) are
 * merged by a single CTA; acc_op is the merge op (MGB_SUM for sums and counts, MGB_MIN, MGB_MAX), val_dtype
 * of acc_col gives the column dtype.  sort_keys != 0 orders the result by key tuple.  Output columns need
 * capacity sum(piece_rows).  *out_n_groups = -1: more than MGB_SMALL_MAX_ROWS (2048) rows in total or more
 * than 16 non-empty pieces -- use mgb_agg_*.                                                             */
int mgb_dense_capacity(int64_t* out_rows);
int mgb_preagg_dense(const mgb_agg_spec* spec, const int64_t* const* key_cols, const void* const* val_cols,
                     int64_t n_rows, int64_t* const* out_key_cols, void* const* out_acc_cols,
                     int64_t out_capacity, int64_t* out_n_groups, void* stream);
int mgb_merge_small(const mgb_agg_spec* spec, int32_t n_pieces, const int64_t* piece_rows,
                    const int64_t* const* const* piece_key_cols, const void* const* const* piece_acc_cols,
                    int64_t* const* out_key_cols, void* const* out_acc_cols, int64_t out_capacity,
                    int32_t sort_keys, int64_t* out_n_groups, void* stream);
/* Asynchronous variants: no host synchronisation, so a chain map -> combine -> ... -> agg of one subtask
 * (the processor runs a subtask's operands back to back, processor.py:216-282) is enqueued without the host
 * waiting for any intermediate size.  The group count goes to *out_n_dev, a caller-owned DEVICE int64
 * (>= 0, or -1 for "not a dense / small input": then nothing useful was written and the caller redoes the
 * step on the general path once it has read the count).  mgb_merge_small_async accepts pieces whose size is
 * not known on the host yet: piece_rows[i] < 0 means "read it from the device int64 *piece_rows_dev[i]"
 * (the out_n_dev of an earlier asynchronous call); a -1 found there propagates to *out_n_dev.  Output columns
 * need capacity mgb_dense_capacity() rows resp. min(out_capacity, 2048).                                  */
int mgb_preagg_dense_async(const mgb_agg_spec* spec, const int64_t* const* key_cols, const void* const* val_cols,
                           int64_t n_rows, int64_t* const* out_key_cols, void* const* out_acc_cols,
                           int64_t out_capacity, int64_t* out_n_dev, void* stream);
int mgb_merge_small_async(const mgb_agg_spec* spec, int32_t n_pieces, const int64_t* piece_rows,
                          const int64_t* const* piece_rows_dev, const int64_t* const* const* piece_key_cols,
                          const void* const* const* piece_acc_cols, int64_t* const* out_key_cols,
                          void* const* out_acc_cols, int64_t out_capacity, int32_t sort_keys, int64_t* out_n_dev,
                          void* stream);

/* Band-level batch (one launch for a whole subtask): the worker runs the chunk operands of a subtask back to
 * back (mars/services/subtask/worker/processor.py:216-282); for the tree branch that is n stage=map operands
 * followed by the concat / stage=combine tree (aggregation.py:715-780).  mgb_preagg_dense_batch_async streams
 * ALL row batches of the band through one launch of the dense kernel (the chunks form one virtual sequence of
 * row tiles; one cross-CTA merge at the end instead of one per chunk) and yields what the map operands plus
 * the combine tree yield: one partial row per distinct key tuple of the band.  chunk_key_cols[i] /
 * chunk_val_cols[i] are the column tables of chunk i (a count-only int64 column may be null, as above);
 * at most mgb_dense_batch_max() chunks per call; out_n_dev as for mgb_preagg_dense_async.               */
int mgb_dense_batch_max(int32_t* out_chunks);
/* kernels the last mgb_preagg_dense* call of this thread enqueued (0: nothing to do / not a dense shape; 2: the
 * register-accumulator kernel for <= 4 groups per CTA went first, the dense kernel behind it) -- for launch accounting */
int mgb_dense_last_launches(int32_t* out_launches);
int mgb_preagg_dense_batch_async(const mgb_agg_spec* spec, int32_t n_chunks, const int64_t* chunk_rows,
                                 const int64_t* const* const* chunk_key_cols,
                                 const void* const* const* chunk_val_cols, int64_t* const* out_key_cols,
                                 void* const* out_acc_cols, int64_t out_capacity, int64_t* out_n_dev, void* stream);

/* Fused pre-path: TPC-H Q1 is filter + two derived columns + groupby
 * (benchmarks/tpch/run_queries.py:110-163); run operator by operator that is three extra passes over lineitem
 * (mask + compaction, `price * (1 - discount)`, `... * (1 + tax)`) before the map stage reads it again.  Here the
 * band-level batch kernel loads prog->n_phys PHYSICAL columns per row (chunk_phys_cols[i][0 .. n_phys)), drops
 * the rows that fail `physical column filter_col <op> constant` (op: 0 <=, 1 <, 2 >=, 3 >, 4 ==, 5 !=; filter_col
 * -1: no filter) and aggregates LOGICAL columns: logical column c is physical column c (n_factors[c] == 0) or the
 * left-to-right product of n_factors[c] <= 3 terms alpha + beta * physical column `col`, every operation rounded
 * on its own like the reference's chain of pandas operators.  Column indices count the loaded columns of the
 * spec (its float64 value columns in spec order; int64 columns that are only counted take no index).  Only the
 * sum / mean / count shapes are fused (*out_n_dev = -1 otherwise: the caller materialises the columns).      */
typedef struct {
  int32_t col;
  double alpha, beta;
} mgb_affine_factor;
typedef struct {
  int32_t n_phys;
  int32_t n_factors[8];
  mgb_affine_factor factors[8][3];
  int32_t filter_col, filter_op, filter_dtype;   /* filter_dtype: mgb_dtype of the filter column */
  int64_t filter_i64;
  double filter_f64;
} mgb_row_program;
int mgb_preagg_dense_batch_fused_async(const mgb_agg_spec* spec, const mgb_row_program* prog, int32_t n_chunks,
                                       const int64_t* chunk_rows, const int64_t* const* const* chunk_key_cols,
                                       const void* const* const* chunk_phys_cols, int64_t* const* out_key_cols,
                                       void* const* out_acc_cols, int64_t out_capacity, int64_t* out_n_dev,
                                       void* stream);

/* Elementwise evaluation of one such derived column (the unfused form: materialises it), and the row filter as a
 * stable compaction of n_cols 8-byte columns (*out_n_dev = rows kept; output columns need capacity n_rows).   */
int mgb_eval_product(const void* const* phys_cols, const mgb_affine_factor* factors, int32_t n_factors, double* out,
                     int64_t n_rows, void* stream);
int mgb_filter_rows(const void* filter_col, int32_t filter_dtype, int32_t filter_op, int64_t filter_i64,
                    double filter_f64, const void* const* in_cols, void* const* out_cols, int32_t n_cols,
                    int64_t n_rows, int64_t* out_n_host, void* stream);

/* The same band-level batch for the MID-cardinality sum / mean / count shapes (BASELINE.json config 1: 1000
 * keys): more distinct key tuples than the dense kernel keeps per CTA, few enough for one shared-memory table
 * per CTA -- float64 columns with SUM and / or COUNT (int64 columns that are only counted are not read), at
 * most mgb_mid_capacity() groups in the band, no NaN values.  Anything else reports -1 through out_n_dev
 * (nothing useful written) and the caller takes mgb_agg_*.  Accumulator records are updated under a per-slot
 * lock with plain shared-memory loads / stores: one atomic per row instead of one compare-and-swap loop per
 * (row, column).  Output columns need capacity mgb_mid_capacity() rows; rows come in table order.           */
int mgb_mid_capacity(int64_t* out_rows);
int mgb_preagg_mid_batch_async(const mgb_agg_spec* spec, int32_t n_chunks, const int64_t* chunk_rows,
                               const int64_t* const* const* chunk_key_cols,
                               const void* const* const* chunk_val_cols, int64_t* const* out_key_cols,
                               void* const* out_acc_cols, int64_t out_capacity, int64_t* out_n_dev, void* stream);

/* stage=agg in ONE launch (aggregation.py:1166-1267): mgb_merge_small_async followed, inside the same kernel,
 * by the post functions of the reference on the merged rows (reduction/mean.py:26-33, var.py:38-48, same
 * operation order as mgb_post_mean / mgb_post_var).  Result column j = post[j] evaluated on the merged
 * accumulators a0 (, a1, a2):  IDENTITY: a0;  MEAN: a0 = sum, a1 = count;  VAR: a0 = sum, a1 = sum of squares,
 * a2 = count.  out_post_cols[j] are 8-byte columns of out_capacity rows (float64, or the accumulator's own
 * dtype for IDENTITY); out_acc_cols is scratch for the merged accumulators.                             */
typedef enum { MGB_POST_IDENTITY = 0, MGB_POST_MEAN = 1, MGB_POST_VAR = 2 } mgb_post_kind;
typedef struct {
  int32_t kind;        /* mgb_post_kind */
  int32_t a0, a1, a2;  /* accumulator indices of the merge spec */
  int32_t ddof;        /* VAR only */
} mgb_post_step;
#define MGB_MAX_POST 128
int mgb_merge_small_post_async(const mgb_agg_spec* spec, int32_t n_pieces, const int64_t* piece_rows,
                               const int64_t* const* piece_rows_dev, const int64_t* const* const* piece_key_cols,
                               const void* const* const* piece_acc_cols, int64_t* const* out_key_cols,
                               void* const* out_acc_cols, int64_t out_capacity, int32_t sort_keys,
                               const mgb_post_step* post, int32_t n_post, void* const* out_post_cols,
                               int64_t* out_n_dev, void* stream);

/* Host-buffer convenience used by the end-to-end path: key/value columns live in (pinned) host memory,
 * are streamed to the device in row blocks overlapped with the kernels, and the result is written to
 * host arrays with capacity max_groups.  Returns the number of groups through out_n_groups
 * (MGB_ERR_OVERFLOW if it exceeds max_groups).                                                           */
int mgb_groupby_agg_host(const mgb_agg_spec* spec, const int64_t* const* key_cols_host,
                         const void* const* val_cols_host, int64_t n_rows, int64_t block_rows,
                         int64_t* const* out_key_cols_host, void* const* out_acc_cols_host,
                         int64_t max_groups, int32_t sort_keys, int64_t* out_n_groups, void* stream);

/* ---- a4 for chunks whose group keys are (nearly) all distinct: a row is its own partial ------------------------
 * The map stage of the reference groups every chunk (aggregation.py:1043-1128); when a chunk has about as many
 * groups as rows that fold reduces nothing, and the combine / agg stages re-group the partial rows anyway
 * (aggregation.py:1130-1267).  The partial of such a chunk is then the chunk itself: SUM / MIN / MAX partials are
 * the value column (aliased, no copy), and this entry writes the two that are not: op = MGB_COUNT -> int64
 * "value is not NaN" (1 for an int64 column), op = MGB_SUMSQ -> value * value rounded (no FMA), dtype of the column. */
int mgb_row_partial(const void* val, int32_t val_dtype, int32_t op, void* out, int64_t n, void* stream);

/* ---- a10: post functions -----------------------------------------------------------------------------
 * The generated post_funcs of the reference, evaluated in float64 in the reference's operation order
 * (mars/dataframe/reduction/mean.py:26-33, var.py:38-48):
 *   mean = S / C;   var(ddof!=0) = (Q - S*S/C) / (C - ddof);   var(ddof==0) = Q/C - (S/C)*(S/C)
 * sum_dtype tells whether S (and Q) are float64 or int64 columns; C is int64.                           */
int mgb_post_mean(const void* sum, int32_t sum_dtype, const int64_t* count, double* out, int64_t n,
                  void* stream);
int mgb_post_var(const void* sum, const void* sumsq, int32_t sum_dtype, const int64_t* count, int32_t ddof,
                 double* out, int64_t n, void* stream);

/* ---- a7/a8: row-wise concatenation of partial columns (pd.concat(axis=0), groupby/core.py:457,
 * merge/concat.py:339-348): copies n_rows 8-byte elements of src into dst starting at dst_offset.      */
int mgb_copy_rows(void* dst, int64_t dst_offset, const void* src, int64_t n_rows, void* stream);

/* a6: assembles the send buffer of the all-to-all: n_segments copies of n_rows[i] 8-byte elements from
 * src[i] to dst[i] in one launch (per 96 segments) instead of one cudaMemcpyAsync per block and column.  */
int mgb_copy_segments(const void* const* src, void* const* dst, const int64_t* n_rows, int32_t n_segments,
                      void* stream);

/* ---- the step before the path: DataFrameToGPU (mars/dataframe/base/to_gpu.py:28-35) of int64 columns ----------
 * The reference hands the whole pandas chunk to cudf.DataFrame.from_pandas, 8 bytes per int64 element over PCIe.
 * Here an int64 column whose values fit a narrower integer crosses the link narrow and is restored in HBM:
 *   mgb_host_narrow_i64  HOST code: dst[i] = (intW_t)src[i] for width W = 1 / 2 / 4 bytes; *fits = 1 when every value
 *                        survived the round trip (the encoding is lossless), 0 as soon as one does not (dst is then
 *                        garbage and the caller uploads the 8-byte column).  No device work, no stream: callers run
 *                        it from several threads on sub-ranges of a column, writing into pinned staging memory.
 *   mgb_widen_i64        device: dst[i] = (int64_t)src[i] (sign-extending), src / dst device pointers aligned to
 *                        16 bytes.  Every kernel of the path keeps reading 8-byte columns.                          */
int mgb_host_narrow_i64(const int64_t* src, int64_t n, int32_t width, void* dst, int32_t* fits);
int mgb_widen_i64(const void* src, int32_t width, int64_t n, int64_t* dst, void* stream);
/* The same for float64 columns that hold integers (quantities, counts kept as floats): dst[i] = (intW_t)src[i], *fits = 1
 * when every value is in range and converts back to the same BITS (fractions, -0.0, NaN and +-inf refuse);
 * mgb_widen_f64: dst[i] = (double)src[i].                                                                          */
int mgb_host_narrow_f64(const double* src, int64_t n, int32_t width, void* dst, int32_t* fits);
/* Which bodies the two host passes run: returns 2 (AVX2, picked at run time when the CPU has it) or 1 (SSE2 / scalar);
 * force_baseline != 0 pins the fallback bodies (tests compare the two), 0 restores the choice.                       */
int mgb_host_narrow_isa(int32_t force_baseline);
int mgb_widen_f64(const void* src, int32_t width, int64_t n, double* dst, void* stream);

/* ---- sort=True support: argsort of int64 key tuples (lexicographic, stable) and row gather ---------- */
int mgb_argsort_keys(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows, int32_t* out_perm,
                     void* stream);
int mgb_gather_rows(const void* const* in_cols, void* const* out_cols, int32_t n_cols, const int32_t* perm,
                    int64_t n_rows, void* stream);

/* ---- a6/e: multi-GPU exchange (grouped ncclSend/ncclRecv all-to-all over NVLink) ---------------------
 * Replaces the storage-service shuffle transport (mars/services/subtask/worker/processor.py:396-432,
 * mars/services/storage/transfer.py:63-206).  libnccl is dlopen()ed from `path` (NULL: default search). */
typedef struct mgb_comm mgb_comm;
int mgb_nccl_load(const char* libnccl_path);
int mgb_comm_unique_id(uint8_t out_id[128]);
int mgb_comm_init(mgb_comm** out, const uint8_t id[128], int32_t rank, int32_t n_ranks);
/* For each of n_cols 8-byte columns: rank r sends send_cols[c][send_offsets[p] .. send_offsets[p+1]) to
 * peer p and receives recv_offsets[p+1]-recv_offsets[p] elements from peer p at recv_cols[c]+recv_offsets[p].
 * Offsets are host arrays of n_ranks+1 int64.  One ncclGroupStart/End for all columns and peers.        */
int mgb_comm_alltoallv(mgb_comm* comm, const void* const* send_cols, const int64_t* send_offsets_host,
                       void* const* recv_cols, const int64_t* recv_offsets_host, int32_t n_cols,
                       void* stream);
int mgb_comm_allgather_i64(mgb_comm* comm, const int64_t* send, int64_t* recv, int64_t count, void* stream);
int mgb_comm_destroy(mgb_comm* comm);

/* ---- per-kernel device timing (bench.py roofline: CUDA events on the launching stream around each
 * kernel of the timed region).  Off by default; enabling adds two cudaEventRecord per launch.          */
typedef enum {
  MGB_K_PREAGG_TINY = 0,  /* map-side pre-aggregation, privatised tiny-cardinality kernel */
  MGB_K_PREAGG_SMEM = 1,  /* map-side pre-aggregation, shared-memory hash table kernel */
  MGB_K_MERGE = 2,        /* partial rows -> global table (reduce-side merge) */
  MGB_K_UPDATE_GLOBAL = 3,/* spilled raw rows -> global table */
  MGB_K_TABLE_INIT = 4,
  MGB_K_EMIT = 5,         /* table compaction (count, scan, write) */
  MGB_K_HASH = 6,         /* key hash / partition id */
  MGB_K_PARTITION = 7,    /* histogram + scan + scatter */
  MGB_K_SORT = 8,         /* radix argsort + gather for sort=True */
  MGB_K_POST = 9,
  MGB_K_BUCKET = 10,      /* bucketed aggregation: bucket ids, per-bucket aggregation, compaction */
  MGB_K_COUNT_ = 11
} mgb_kernel_id;
int mgb_prof_enable(int32_t on);
int mgb_prof_reset(void);
/* total device milliseconds and number of bracketed launches of kernel_id since the last reset
 * (synchronises on the recorded events). */
int mgb_prof_read(int32_t kernel_id, double* out_total_ms, int64_t* out_launches);

/* Peer-to-peer pull transport on one NVSwitch box (alternative to mgb_comm_alltoallv): mgb_ipc_arena returns a
 * library-owned device arena of at least `bytes` (grown by reallocation; the caller packs its send buffer
 * into it) and the arena's 64-byte CUDA IPC handle, which the caller publishes to the peers along with its
 * block counts; mgb_ipc_open maps a peer's arena (cached per peer, re-opened when the handle changes);
 * mgb_ipc_pull copies a slice out of it with cudaMemcpyAsync (copy engines over NVLink).  The caller
 * synchronises its pack before publishing and runs a barrier after its pulls, before the arena is reused. */
int mgb_ipc_arena(int64_t bytes, void** out_ptr, uint8_t out_handle[64]);
int mgb_ipc_open(int32_t peer, const uint8_t handle[64], void** out_ptr);
int mgb_ipc_pull(void* dst, const void* peer_src, int64_t bytes, void* stream);
/* Numbered arenas (slot 0 is the arena of mgb_ipc_arena; slots 1..3 are receive arenas of the push shuffle, rotated so
 * that the blocks of one shuffle stay valid while the next one is written).  mgb_ipc_arena_slot (re)allocates when
 * `bytes` exceeds the slot's capacity: the caller must first make every peer drop its mapping of that slot
 * (mgb_ipc_close_slot on the peers, then a barrier) -- an exported allocation is never freed while it is mapped.
 * out_capacity / out_fresh (may be NULL): capacity in bytes, 1 when this call allocated.                          */
int mgb_ipc_arena_slot(int32_t slot, int64_t bytes, void** out_ptr, uint8_t out_handle[64], int64_t* out_capacity,
                       int32_t* out_fresh);
int mgb_ipc_open_slot(int32_t peer, int32_t slot, const uint8_t handle[64], void** out_ptr);
int mgb_ipc_close_slot(int32_t peer, int32_t slot);

#ifdef __cplusplus
}
#endif
#endif /* MGB_H_ */
