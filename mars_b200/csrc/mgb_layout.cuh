// libmarsgb: accumulator layout of one aggregator, shared by the hash-table path (mgb_agg.cu) and the
// bucketed path (mgb_bucket.cu).
#pragma once
#include "mgb_common.cuh"

struct AccDesc {
  uint8_t col;   // value column (spec index)
  uint8_t op;    // mgb_op
  uint8_t dt;    // dtype of the value column
  uint8_t kind;  // mgb_acc_kind
};

// Device-visible layout, passed by value.
struct Layout {
  int nk, nv, na;
  int W;                                 // 8-byte words per global table slot (keys first)
  AccDesc acc[MGB_MAX_ACCS];             // accumulators sorted by column
  uint8_t acc_index[MGB_MAX_ACCS];       // spec index of sorted accumulator j
  uint8_t col_begin[MGB_MAX_VALS + 1];   // sorted accumulators of column c: [col_begin[c], col_begin[c+1])
  uint8_t spec_op[MGB_MAX_ACCS];         // op / dtype / sorted position of spec accumulator a
  uint8_t spec_dt[MGB_MAX_ACCS];
  uint8_t spec_sorted[MGB_MAX_ACCS];
};

// ----------------------------------------------------------------------------------------------------
// bucketed aggregation (mgb_bucket.cu): rows of all sources are split by key hash into buckets of a few
// hundred rows, one CTA aggregates one bucket in registers -- no atomics on the values, no table sized by
// the row count.  Used by mgb_agg_finalize for inputs that were not pre-aggregated.
// ----------------------------------------------------------------------------------------------------
struct BktSource {               // one update() batch: column pointers (device) and its first global row
  const uint64_t* keys[MGB_MAX_KEYS];
  const uint64_t* vals[MGB_MAX_VALS];
  long long begin;
};

struct BktResult {
  uint64_t* scratch;             // SoA [nk + na][T], table representation; bucket b owns rows [off[b], ...)
  int64_t* off;                  // [B + 1] first row of every bucket
  int64_t* goff;                 // [B + 1] first output group of every bucket (exclusive scan of the counts)
  long long T;
  int B;
  long long G;                   // number of groups
};

// status MGB_OK with *overflow != 0 means nothing was produced and the caller takes the hash-table path:
// 1 = a bucket held more groups than its on-chip table, 2 = a bucket is far longer than the average (skewed keys)
int mgb_bucket_aggregate(const Layout& L, const BktSource* srcs_host, int nsrc, long long T, cudaStream_t st,
                         BktResult* out, int* overflow);
int mgb_bucket_emit(const Layout& L, const BktResult& r, uint64_t* const* out_key, uint64_t* const* out_acc,
                    cudaStream_t st);
void mgb_bucket_free(BktResult* r, cudaStream_t st);
bool mgb_bucket_supports(const Layout& L);

