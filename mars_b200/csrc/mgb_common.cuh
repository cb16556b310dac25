// Shared device/host helpers for libmarsgb (sm_100a).  See include/mgb.h for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mgb.h"

// ----------------------------------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------------------------------
void mgb_set_error(const char* fmt, ...);

#define MGB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      mgb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));    \
      return MGB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define MGB_CHECK_LAUNCH()                                                                     \
  do {                                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess) {                                                                  \
      mgb_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return MGB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define MGB_REQUIRE(cond, status, ...) \
  do {                                 \
    if (!(cond)) {                     \
      mgb_set_error(__VA_ARGS__);      \
      return (status);                 \
    }                                  \
  } while (0)

#define MGB_TRY(expr)              \
  do {                             \
    int s__ = (expr);              \
    if (s__ != MGB_OK) return s__; \
  } while (0)

int mgb_sm_count();

// ----------------------------------------------------------------------------------------------------
// per-kernel device timing (mgb_prof_*): when enabled, every launch site brackets its kernel with two CUDA
// events on the launching stream; mgb_prof_read sums the elapsed times per kernel id.
// ----------------------------------------------------------------------------------------------------
bool mgb_prof_on();
void mgb_prof_record(int kernel_id, cudaStream_t st, bool begin);
struct MgbProfScope {
  int id;
  cudaStream_t st;
  bool on;
  MgbProfScope(int id_, cudaStream_t st_) : id(id_), st(st_), on(mgb_prof_on()) {
    if (on) mgb_prof_record(id, st, true);
  }
  ~MgbProfScope() {
    if (on) mgb_prof_record(id, st, false);
  }
};

// ----------------------------------------------------------------------------------------------------
// pandas-exact hashing (pandas/core/util/hashing.py: _hash_ndarray tail, combine_hash_arrays), as used by
// the reference's hash_dataframe_on (mars/dataframe/utils.py:77-101).
// ----------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mgb_mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}

// hash of a key tuple: one key -> mix64; N keys -> CPython-tuple-hash style combine of the per-level hashes.
template <int NK>
__host__ __device__ __forceinline__ uint64_t mgb_hash_tuple(const int64_t* k) {
  if (NK == 1) return mgb_mix64((uint64_t)k[0]);
  uint64_t out = 0x345678ULL;
  uint64_t mult = 1000003ULL;
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    uint64_t inv = (uint64_t)(NK - i);
    out ^= mgb_mix64((uint64_t)k[i]);
    out *= mult;
    mult += 82520ULL + inv + inv;
  }
  out += 97531ULL;
  return out;
}

// hash of a ROW of key columns with pandas' DataFrame semantics -- hash_pandas_object(df[on], index=False), what
// hash_dataframe_on(df, on=[...]) uses for plain groupby / merge / drop_duplicates shuffles
// (mars/dataframe/utils.py:88-100): the per-column hashes always go through combine_hash_arrays, also when there
// is a single column (an Index, in contrast, is hashed directly: mgb_hash_tuple<1>).
template <int NK>
__host__ __device__ __forceinline__ uint64_t mgb_hash_frame_row(const int64_t* k) {
  if (NK == 1) return (0x345678ULL ^ mgb_mix64((uint64_t)k[0])) * 1000003ULL + 97531ULL;
  return mgb_hash_tuple<NK>(k);
}

// slot hash: decorrelated from `hash % R` (all keys of one reducer share hash % R).
__host__ __device__ __forceinline__ uint64_t mgb_slot_hash(uint64_t h) { return h * 0x9e3779b97f4a7c15ULL; }

// ----------------------------------------------------------------------------------------------------
// sortable encodings for atomic min / max
// ----------------------------------------------------------------------------------------------------
#define MGB_SIGN64 0x8000000000000000ULL
#define MGB_EMPTY_KEY ((int64_t)0x8000000000000000LL)

__host__ __device__ __forceinline__ uint64_t mgb_enc_f64(uint64_t bits) {
  return (bits & MGB_SIGN64) ? ~bits : (bits | MGB_SIGN64);
}
__host__ __device__ __forceinline__ uint64_t mgb_dec_f64(uint64_t e) {
  return (e & MGB_SIGN64) ? (e & ~MGB_SIGN64) : ~e;
}
__host__ __device__ __forceinline__ uint64_t mgb_enc_i64(uint64_t bits) { return bits ^ MGB_SIGN64; }
__host__ __device__ __forceinline__ uint64_t mgb_dec_i64(uint64_t e) { return e ^ MGB_SIGN64; }

__host__ __device__ __forceinline__ bool mgb_is_nan_bits(uint64_t bits) {
  return (bits & 0x7fffffffffffffffULL) > 0x7ff0000000000000ULL;
}

// Table representation of an accumulator:
//   SUM/SUMSQ f64: raw double bits; SUM/SUMSQ i64, COUNT: int64; MIN/MAX: sortable-encoded uint64 where
//   ~0 (MIN) / 0 (MAX) mean "no value yet".
__host__ __device__ __forceinline__ uint64_t mgb_acc_init(int op) {
  return op == MGB_MIN ? ~0ULL : 0ULL;
}
// identity in the *decoded* (emitted) representation
__host__ __device__ __forceinline__ uint64_t mgb_acc_decode(int op, int dt, uint64_t t) {
  if (op == MGB_MIN) {
    if (t == ~0ULL) return dt == MGB_F64 ? 0x7ff8000000000000ULL : 0x7fffffffffffffffULL;
    return dt == MGB_F64 ? mgb_dec_f64(t) : mgb_dec_i64(t);
  }
  if (op == MGB_MAX) {
    if (t == 0ULL) return dt == MGB_F64 ? 0x7ff8000000000000ULL : 0x8000000000000000ULL;
    return dt == MGB_F64 ? mgb_dec_f64(t) : mgb_dec_i64(t);
  }
  return t;
}

// Contribution of one raw value to an accumulator, in table representation.  Returns false when the
// value is skipped (NaN).
__device__ __forceinline__ bool mgb_contrib(int op, int dt, uint64_t x, uint64_t* out) {
  if (dt == MGB_F64) {
    if (mgb_is_nan_bits(x)) {
      *out = 0;
      return false;
    }
    switch (op) {
      case MGB_SUM: *out = x; return true;
      case MGB_SUMSQ: {
        double v = __longlong_as_double((long long)x);
        *out = (uint64_t)__double_as_longlong(v * v);
        return true;
      }
      case MGB_COUNT: *out = 1; return true;
      default: *out = mgb_enc_f64(x); return true;
    }
  } else {
    switch (op) {
      case MGB_SUM: *out = x; return true;
      case MGB_SUMSQ: *out = x * x; return true;
      case MGB_COUNT: *out = 1; return true;
      default:
        // INT64_MAX / INT64_MIN are the identities of emitted i64 MIN / MAX partials: skipping them
        // keeps "no value" stable through merges and does not change a real min/max.
        *out = mgb_enc_i64(x);
        return true;
    }
  }
}

// kind of accumulate: 0 = f64 add, 1 = i64 add, 2 = u64 min, 3 = u64 max
__host__ __device__ __forceinline__ int mgb_acc_kind(int op, int dt) {
  if (op == MGB_MIN) return 2;
  if (op == MGB_MAX) return 3;
  if (op == MGB_COUNT) return 1;
  return dt == MGB_F64 ? 0 : 1;
}

__device__ __forceinline__ void mgb_atomic_apply_global(uint64_t* p, int kind, uint64_t c) {
  switch (kind) {
    case 0: atomicAdd((double*)p, __longlong_as_double((long long)c)); break;
    case 1: atomicAdd((unsigned long long*)p, (unsigned long long)c); break;
    case 2: atomicMin((unsigned long long*)p, (unsigned long long)c); break;
    default: atomicMax((unsigned long long*)p, (unsigned long long)c); break;
  }
}

__device__ __forceinline__ uint64_t mgb_combine(int kind, uint64_t a, uint64_t c) {
  switch (kind) {
    case 0:
      return (uint64_t)__double_as_longlong(__longlong_as_double((long long)a) +
                                            __longlong_as_double((long long)c));
    case 1: return a + c;
    case 2: return a < c ? a : c;
    default: return a > c ? a : c;
  }
}

__device__ __forceinline__ unsigned mgb_lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// streaming loads: inputs are read exactly once, keep them out of L1.
__device__ __forceinline__ uint64_t mgb_ld_stream(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// ----------------------------------------------------------------------------------------------------
// internal cross-file API
// ----------------------------------------------------------------------------------------------------
// exclusive prefix sum over n int64 (in place); total (sum of all) optionally written to *total_out (device).
int mgb_scan_exclusive_i64(int64_t* data, int64_t n, int64_t* total_out, cudaStream_t st);
// stable LSD radix passes over a permutation.  digit source: int32 pid array or int64 key column.
int mgb_radix_sort_perm_pid(const int32_t* pid, int64_t n, int32_t n_partitions, int32_t* perm_out,
                            int64_t* offsets_out, cudaStream_t st);
// same, any R < 2^31 (one pass per 8 bits); check_range = false skips the out-of-range test and its host
// synchronisation (for partition ids computed by the library itself)
int mgb_radix_sort_perm_pid_ex(const int32_t* pid, int64_t n, int32_t n_partitions, int32_t* perm_out,
                               int64_t* offsets_out, bool check_range, cudaStream_t st);
int mgb_radix_sort_perm_keys(const int64_t* const* key_cols_host_array, int32_t n_keys, int64_t n,
                             int32_t* perm_out, cudaStream_t st);
// histogram -> scan -> staged vectorised scatter into local columns (mgb_split.cu), n_partitions <= 256.
// mode 0 / 1: pandas hash of the keys (Index / DataFrame-row semantics); 2: ids given in pid (out-of-range ids set *bad)
int mgb_split_local(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid, int64_t n_rows,
                    int32_t n_partitions, const void* const* in_cols, void* const* out_cols, int32_t n_cols,
                    int64_t* out_offsets, int32_t* bad, cudaStream_t st);
