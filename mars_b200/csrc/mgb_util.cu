// libmarsgb: error plumbing, prefix scan, pandas-exact key hashing, post functions, row copies.
#include <stdarg.h>

#include "mgb_common.cuh"

// ----------------------------------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void mgb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mgb_last_error(void) { return g_err; }
extern "C" int mgb_version(void) { return MGB_VERSION; }

int mgb_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
    // first use of this device: keep freed scratch in the stream-ordered pool instead of returning it to the
    // driver at every synchronisation (the default threshold 0 makes each cudaMallocAsync a driver call)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  return cached[dev];
}

extern "C" int mgb_device_info(int32_t* sm_count, int32_t* cc) {
  int dev = 0, maj = 0, min = 0, sms = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  MGB_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  MGB_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  MGB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count) *sm_count = sms;
  if (cc) *cc = maj * 10 + min;
  return MGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// per-kernel device timing
// ----------------------------------------------------------------------------------------------------
#include <mutex>
#include <vector>
struct ProfRec {
  int id;
  cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;

bool mgb_prof_on() { return g_prof_on; }

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void mgb_prof_record(int kernel_id, cudaStream_t st, bool begin) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (begin) {
    ProfRec r{kernel_id, prof_event(), nullptr};
    if (r.e0) cudaEventRecord(r.e0, st);
    g_prof_recs.push_back(r);
  } else {
    for (size_t i = g_prof_recs.size(); i-- > 0;) {
      ProfRec& r = g_prof_recs[i];
      if (r.id == kernel_id && r.e1 == nullptr) {
        r.e1 = prof_event();
        if (r.e1) cudaEventRecord(r.e1, st);
        break;
      }
    }
  }
}

extern "C" int mgb_prof_enable(int32_t on) {
  g_prof_on = on != 0;
  return MGB_OK;
}

extern "C" int mgb_prof_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) {
    if (r.e0) g_prof_pool.push_back(r.e0);
    if (r.e1) g_prof_pool.push_back(r.e1);
  }
  g_prof_recs.clear();
  return MGB_OK;
}

extern "C" int mgb_prof_read(int32_t kernel_id, double* out_total_ms, int64_t* out_launches) {
  MGB_REQUIRE(out_total_ms && out_launches, MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(kernel_id >= 0 && kernel_id < MGB_K_COUNT_, MGB_ERR_INVALID, "kernel id %d", kernel_id);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double total = 0;
  int64_t n = 0;
  for (auto& r : g_prof_recs) {
    if (r.id != kernel_id || !r.e0 || !r.e1) continue;
    MGB_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0;
    MGB_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    total += ms;
    ++n;
  }
  *out_total_ms = total;
  *out_launches = n;
  return MGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// exclusive scan (int64), three-phase: per-block scan of SCAN_TILE elements, scan of block totals by a
// single block, add-back.  Used by the partition histogram, radix sort and table compaction.
// ----------------------------------------------------------------------------------------------------
#define SCAN_THREADS 512
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* total, int64_t* smem_warp) {
  // v: per-thread value; returns exclusive prefix in thread order, *total = block sum
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) smem_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int64_t w = lane < (SCAN_THREADS / 32) ? smem_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    if (lane < (SCAN_THREADS / 32)) smem_warp[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  int64_t warp_base = warp == 0 ? 0 : smem_warp[warp - 1];
  *total = smem_warp[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return warp_base + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(int64_t* data, int64_t n, int64_t* tile_sums) {
  __shared__ int64_t sw[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t v[SCAN_ITEMS];
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? data[base + i] : 0;
    s += v[i];
  }
  int64_t total;
  int64_t ex = block_exclusive_scan(s, &total, sw);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) data[base + i] = ex;
    ex += v[i];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of m tile sums (any m), total -> *total_out
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(int64_t* sums, int64_t m, int64_t* total_out) {
  __shared__ int64_t sw[32];
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t start = 0; start < m; start += SCAN_THREADS) {
    int64_t i = start + threadIdx.x;
    int64_t v = i < m ? sums[i] : 0;
    int64_t total;
    int64_t ex = block_exclusive_scan(v, &total, sw);
    int64_t carry = carry_s;
    if (i < m) sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int64_t* data, int64_t n, const int64_t* tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  const int64_t add = tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) data[base + i] += add;
}

int mgb_scan_exclusive_i64(int64_t* data, int64_t n, int64_t* total_out, cudaStream_t st) {
  if (n <= 0) {
    if (total_out) MGB_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
    return MGB_OK;
  }
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  int64_t* sums = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&sums, sizeof(int64_t) * tiles, st));
  k_scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(data, n, sums);
  k_scan_sums<<<1, SCAN_THREADS, 0, st>>>(sums, tiles, total_out);
  if (tiles > 1) k_scan_add<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(data, n, sums);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(sums, st);
  if (e != cudaSuccess) {
    mgb_set_error("scan launch failed: %s", cudaGetErrorString(e));
    return MGB_ERR_CUDA;
  }
  return MGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// a5: key hashing and partition ids
// ----------------------------------------------------------------------------------------------------
struct KeyCols {
  const int64_t* k[MGB_MAX_KEYS];
};

template <int NK, bool WANT_PID, bool FRAME = false>
__global__ void __launch_bounds__(256) k_hash(KeyCols kc, int64_t n, uint64_t R, uint64_t* out_hash,
                                              int32_t* out_pid) {
  const bool pow2 = (R & (R - 1)) == 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = (int64_t)mgb_ld_stream((const uint64_t*)kc.k[j] + i);
    uint64_t h = FRAME ? mgb_hash_frame_row<NK>(k) : mgb_hash_tuple<NK>(k);
    if (WANT_PID)
      out_pid[i] = (int32_t)(pow2 ? (h & (R - 1)) : (h % R));
    else
      out_hash[i] = h;
  }
}

static int launch_hash(const int64_t* const* key_cols, int32_t n_keys, int64_t n, uint64_t R,
                       uint64_t* out_hash, int32_t* out_pid, cudaStream_t st) {
  MGB_REQUIRE(n_keys >= 1 && n_keys <= MGB_MAX_KEYS, MGB_ERR_UNSUPPORTED,
              "n_keys=%d outside [1,%d]", n_keys, MGB_MAX_KEYS);
  MGB_REQUIRE(n >= 0, MGB_ERR_INVALID, "negative row count");
  if (n == 0) return MGB_OK;
  KeyCols kc;
  for (int j = 0; j < MGB_MAX_KEYS; ++j) kc.k[j] = j < n_keys ? key_cols[j] : nullptr;
  int64_t blocks = (n + 255) / 256;
  int64_t cap = (int64_t)mgb_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  const bool pid = out_pid != nullptr;
  MgbProfScope prof(MGB_K_HASH, st);
#define L(NK)                                                                                   \
  if (pid)                                                                                      \
    k_hash<NK, true><<<(unsigned)blocks, 256, 0, st>>>(kc, n, R, out_hash, out_pid);            \
  else                                                                                          \
    k_hash<NK, false><<<(unsigned)blocks, 256, 0, st>>>(kc, n, R, out_hash, out_pid);
  if (n_keys == 1) {
    L(1)
  } else {
    L(2)
  }
#undef L
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

extern "C" int mgb_hash_keys(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows,
                             uint64_t* out_hash, void* stream) {
  MGB_REQUIRE(key_cols && (out_hash || n_rows == 0), MGB_ERR_INVALID, "null pointer");
  return launch_hash(key_cols, n_keys, n_rows, 1, out_hash, nullptr, (cudaStream_t)stream);
}

extern "C" int mgb_partition_ids_frame(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows,
                                       int64_t n_partitions, int32_t* out_pid, void* stream) {
  MGB_REQUIRE(key_cols && (out_pid || n_rows == 0), MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(n_partitions >= 1 && n_partitions <= 0x7fffffffLL, MGB_ERR_INVALID, "n_partitions=%lld out of range",
              (long long)n_partitions);
  MGB_REQUIRE(n_keys >= 1 && n_keys <= MGB_MAX_KEYS, MGB_ERR_UNSUPPORTED, "n_keys=%d outside [1,%d]", n_keys,
              MGB_MAX_KEYS);
  MGB_REQUIRE(n_rows >= 0, MGB_ERR_INVALID, "negative row count");
  if (n_rows == 0) return MGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  KeyCols kc;
  for (int j = 0; j < MGB_MAX_KEYS; ++j) kc.k[j] = j < n_keys ? key_cols[j] : nullptr;
  int64_t blocks = (n_rows + 255) / 256, cap = (int64_t)mgb_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  MgbProfScope prof(MGB_K_HASH, st);
  if (n_keys == 1)
    k_hash<1, true, true><<<(unsigned)blocks, 256, 0, st>>>(kc, n_rows, (uint64_t)n_partitions, nullptr, out_pid);
  else
    k_hash<2, true, true><<<(unsigned)blocks, 256, 0, st>>>(kc, n_rows, (uint64_t)n_partitions, nullptr, out_pid);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

extern "C" int mgb_partition_ids(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows,
                                 int64_t n_partitions, int32_t* out_pid, void* stream) {
  MGB_REQUIRE(key_cols && (out_pid || n_rows == 0), MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(n_partitions >= 1 && n_partitions <= 0x7fffffffLL, MGB_ERR_INVALID,
              "n_partitions=%lld out of range", (long long)n_partitions);
  return launch_hash(key_cols, n_keys, n_rows, (uint64_t)n_partitions, nullptr, out_pid,
                     (cudaStream_t)stream);
}

// ----------------------------------------------------------------------------------------------------
// a4: cardinality probe.  Which map-side path pays -- on-chip pre-aggregation (few groups per chunk) or the
// bucketed aggregation (nearly distinct keys) -- depends on the data; the reference has no such choice (pandas
// factorizes either way, aggregation.py:951-984).  A strided sample of the chunk's key tuples is hashed into a
// shared-memory bitmap; the number of distinct hashes tells the two regimes apart for the cost of one tiny
// launch and a 16-byte read.
// ----------------------------------------------------------------------------------------------------
#define EST_THREADS 1024
#define EST_SAMPLE 16384
#define EST_BITS (1 << 18)

template <int NK>
__global__ void __launch_bounds__(EST_THREADS) k_estimate_distinct(KeyCols kc, int64_t n, int64_t* out2) {
  __shared__ uint32_t bm[EST_BITS / 32];
  __shared__ int total;
  for (int i = threadIdx.x; i < EST_BITS / 32; i += EST_THREADS) bm[i] = 0;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  const int64_t S = n < EST_SAMPLE ? n : EST_SAMPLE;
  const int64_t stride = n / S;
  int fresh = 0;
  for (int64_t i = threadIdx.x; i < S; i += EST_THREADS) {
    int64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = kc.k[j][i * stride];
    const uint64_t h = mgb_slot_hash(mgb_hash_tuple<NK>(k));
    const uint32_t bit = (uint32_t)(h >> (64 - 18));
    const uint32_t m = 1u << (bit & 31);
    const uint32_t old = atomicOr(&bm[bit >> 5], m);
    fresh += (old & m) ? 0 : 1;
  }
  for (int o = 16; o > 0; o >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
  if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(&total, fresh);
  __syncthreads();
  if (threadIdx.x == 0) {
    out2[0] = S;
    out2[1] = total;
  }
}

extern "C" int mgb_estimate_distinct(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows,
                                     int64_t* out_sampled, int64_t* out_distinct, void* stream) {
  MGB_REQUIRE(key_cols && out_sampled && out_distinct, MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(n_keys >= 1 && n_keys <= MGB_MAX_KEYS, MGB_ERR_UNSUPPORTED, "n_keys=%d outside [1,%d]", n_keys,
              MGB_MAX_KEYS);
  MGB_REQUIRE(n_rows >= 0, MGB_ERR_INVALID, "negative row count");
  *out_sampled = 0;
  *out_distinct = 0;
  if (n_rows == 0) return MGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  KeyCols kc;
  for (int j = 0; j < MGB_MAX_KEYS; ++j) kc.k[j] = j < n_keys ? key_cols[j] : nullptr;
  int64_t* d = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&d, 16, st));
  if (n_keys == 1)
    k_estimate_distinct<1><<<1, EST_THREADS, 0, st>>>(kc, n_rows, d);
  else
    k_estimate_distinct<2><<<1, EST_THREADS, 0, st>>>(kc, n_rows, d);
  cudaError_t e = cudaGetLastError();
  int64_t h[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFreeAsync(d, st);
  if (e != cudaSuccess) {
    mgb_set_error("cardinality probe: %s", cudaGetErrorString(e));
    return MGB_ERR_CUDA;
  }
  *out_sampled = h[0];
  *out_distinct = h[1];
  return MGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// sort=True shuffle: range partition of partial rows by pivots (DataFrameGroupbySortShuffle._execute_map,
// mars/dataframe/groupby/sort.py:113-135).  Partition i holds the key tuples in (pivot[i-1], pivot[i]], i.e.
// partition id = number of pivots strictly smaller than the key tuple (pivots ascending, signed int64,
// lexicographic for several keys) -- a binary search per row over the pivot table.
// ----------------------------------------------------------------------------------------------------
template <int NK>
__device__ __forceinline__ bool tuple_less(const int64_t* a, const int64_t* b) {
  if (a[0] != b[0]) return a[0] < b[0];
  if (NK == 1) return false;
  return a[1] < b[1];
}

template <int NK>
__global__ void __launch_bounds__(256) k_range_pid(KeyCols kc, KeyCols pv, int n_pivots, int64_t n, int32_t* out_pid) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = (int64_t)mgb_ld_stream((const uint64_t*)kc.k[j] + i);
    int lo = 0, hi = n_pivots;   // first pivot that is not < key
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      int64_t p[NK];
#pragma unroll
      for (int j = 0; j < NK; ++j) p[j] = pv.k[j][mid];
      if (tuple_less<NK>(p, k))
        lo = mid + 1;
      else
        hi = mid;
    }
    out_pid[i] = lo;
  }
}

extern "C" int mgb_range_partition_ids(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows,
                                       const int64_t* const* pivot_cols, int32_t n_pivots, int32_t* out_pid,
                                       void* stream) {
  MGB_REQUIRE(n_keys >= 1 && n_keys <= MGB_MAX_KEYS, MGB_ERR_UNSUPPORTED, "n_keys=%d outside [1,%d]", n_keys,
              MGB_MAX_KEYS);
  MGB_REQUIRE(n_rows >= 0 && n_pivots >= 0, MGB_ERR_INVALID, "negative size");
  if (n_rows == 0) return MGB_OK;
  MGB_REQUIRE(key_cols && out_pid && (n_pivots == 0 || pivot_cols), MGB_ERR_INVALID, "null pointer");
  KeyCols kc, pv;
  for (int j = 0; j < MGB_MAX_KEYS; ++j) {
    kc.k[j] = j < n_keys ? key_cols[j] : nullptr;
    pv.k[j] = (j < n_keys && n_pivots > 0) ? pivot_cols[j] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (n_rows + 255) / 256, cap = (int64_t)mgb_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  MgbProfScope prof(MGB_K_HASH, st);
  if (n_keys == 1)
    k_range_pid<1><<<(unsigned)blocks, 256, 0, st>>>(kc, pv, n_pivots, n_rows, out_pid);
  else
    k_range_pid<2><<<(unsigned)blocks, 256, 0, st>>>(kc, pv, n_pivots, n_rows, out_pid);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// a10: post functions.  Operation order follows the reference's generated source exactly.
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ double as_f64(const void* p, int dt, int64_t i) {
  return dt == MGB_F64 ? ((const double*)p)[i] : (double)((const int64_t*)p)[i];
}

__global__ void k_post_mean(const void* sum, int dt, const int64_t* cnt, double* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = as_f64(sum, dt, i) / (double)cnt[i];
}

__global__ void k_post_var(const void* sum, const void* sumsq, int dt, const int64_t* cnt, int ddof,
                           double* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double s = as_f64(sum, dt, i), q = as_f64(sumsq, dt, i), c = (double)cnt[i];
    double r;
    if (ddof == 0) {
      // (x**2).mean() - (x.mean())**2
      const double m2 = __ddiv_rn(q, c), m = __ddiv_rn(s, c);
      r = __dsub_rn(m2, __dmul_rn(m, m));
    } else {
      // ((x**2).sum() - x.sum()**2 / cnt) / (cnt - ddof)      (no FMA contraction: *_rn intrinsics)
      const double t = __ddiv_rn(__dmul_rn(s, s), c);
      r = __ddiv_rn(__dsub_rn(q, t), __dsub_rn(c, (double)ddof));
    }
    out[i] = r;
  }
}

static unsigned grid_for(int64_t n) {
  int64_t b = (n + 255) / 256, cap = (int64_t)mgb_sm_count() * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// a4 without a fold: the per-row partial of one accumulator for chunks whose keys are (nearly) all distinct.  A row IS
// its group's partial there: SUM / MIN / MAX are the value itself (the caller aliases the column), COUNT is
// "value is not NaN" and SUMSQ the rounded square -- written here, two rows per 16-byte access.
__global__ void __launch_bounds__(256) k_row_partial(const uint64_t* __restrict__ v, int dt, int op,
                                                     uint64_t* __restrict__ out, int64_t n) {
  const int64_t pairs = n >> 1;
  auto one = [&](uint64_t x) -> uint64_t {
    if (op == MGB_COUNT) return (dt == MGB_F64 && mgb_is_nan_bits(x)) ? 0ULL : 1ULL;
    if (dt == MGB_F64) {
      const double d = __longlong_as_double((long long)x);
      return (uint64_t)__double_as_longlong(__dmul_rn(d, d));
    }
    return x * x;
  };
  const bool vec = (((uintptr_t)v | (uintptr_t)out) & 15) == 0;
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (int64_t)gridDim.x * blockDim.x) {
      const ulonglong2 x = __ldg(reinterpret_cast<const ulonglong2*>(v) + i);
      ulonglong2 y;
      y.x = one(x.x);
      y.y = one(x.y);
      reinterpret_cast<ulonglong2*>(out)[i] = y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = one(v[n - 1]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      out[i] = one(mgb_ld_stream(v + i));
  }
}

extern "C" int mgb_row_partial(const void* val, int32_t val_dtype, int32_t op, void* out, int64_t n, void* stream) {
  MGB_REQUIRE(n >= 0, MGB_ERR_INVALID, "negative n");
  MGB_REQUIRE(op == MGB_COUNT || op == MGB_SUMSQ, MGB_ERR_INVALID, "op %d needs no per-row partial", op);
  if (n == 0) return MGB_OK;
  MGB_REQUIRE(val && out, MGB_ERR_INVALID, "null pointer");
  MgbProfScope prof(MGB_K_POST, (cudaStream_t)stream);
  k_row_partial<<<grid_for((n + 1) / 2), 256, 0, (cudaStream_t)stream>>>((const uint64_t*)val, val_dtype, op,
                                                                        (uint64_t*)out, n);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

extern "C" int mgb_post_mean(const void* sum, int32_t sum_dtype, const int64_t* count, double* out,
                             int64_t n, void* stream) {
  MGB_REQUIRE(n >= 0, MGB_ERR_INVALID, "negative n");
  if (n == 0) return MGB_OK;
  MGB_REQUIRE(sum && count && out, MGB_ERR_INVALID, "null pointer");
  MgbProfScope prof(MGB_K_POST, (cudaStream_t)stream);
  k_post_mean<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(sum, sum_dtype, count, out, n);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

extern "C" int mgb_post_var(const void* sum, const void* sumsq, int32_t sum_dtype, const int64_t* count,
                            int32_t ddof, double* out, int64_t n, void* stream) {
  MGB_REQUIRE(n >= 0, MGB_ERR_INVALID, "negative n");
  if (n == 0) return MGB_OK;
  MGB_REQUIRE(sum && sumsq && count && out, MGB_ERR_INVALID, "null pointer");
  MgbProfScope prof(MGB_K_POST, (cudaStream_t)stream);
  k_post_var<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(sum, sumsq, sum_dtype, count, ddof, out, n);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// a6: packing of shuffle blocks -- many (source slice -> destination slice) copies of 8-byte elements in one
// launch (the send buffer of the all-to-all is assembled from per-mapper, per-column slices).
// ----------------------------------------------------------------------------------------------------
#define SEG_MAX 96
struct SegTable {
  const uint64_t* src[SEG_MAX];
  uint64_t* dst[SEG_MAX];
  long long n[SEG_MAX];
  int count;
};

__global__ void __launch_bounds__(256) k_copy_segments(const SegTable t) {
  // blockIdx.y: segment; blockIdx.x strides over the segment
  const int sgi = blockIdx.y;
  const uint64_t* src = t.src[sgi];
  uint64_t* dst = t.dst[sgi];
  const long long n = t.n[sgi];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = mgb_ld_stream(src + i);
}

extern "C" int mgb_copy_segments(const void* const* src, void* const* dst, const int64_t* n_rows, int32_t n_segments,
                                 void* stream) {
  MGB_REQUIRE(n_segments >= 0, MGB_ERR_INVALID, "negative segment count");
  if (n_segments == 0) return MGB_OK;
  MGB_REQUIRE(src && dst && n_rows, MGB_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = mgb_sm_count();
  for (int s0 = 0; s0 < n_segments; s0 += SEG_MAX) {
    SegTable t;
    t.count = 0;
    long long longest = 0;
    for (int i = s0; i < n_segments && t.count < SEG_MAX; ++i) {
      MGB_REQUIRE(n_rows[i] >= 0, MGB_ERR_INVALID, "negative segment length");
      if (n_rows[i] == 0) continue;
      MGB_REQUIRE(src[i] && dst[i], MGB_ERR_INVALID, "null segment pointer");
      t.src[t.count] = (const uint64_t*)src[i];
      t.dst[t.count] = (uint64_t*)dst[i];
      t.n[t.count] = n_rows[i];
      if (n_rows[i] > longest) longest = n_rows[i];
      ++t.count;
    }
    if (t.count == 0) continue;
    long long bx = (longest + 256 * 8 - 1) / (256 * 8);
    const long long cap = (long long)sms * 8 / t.count + 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    k_copy_segments<<<dim3((unsigned)bx, (unsigned)t.count), 256, 0, st>>>(t);
    MGB_CHECK_LAUNCH();
  }
  return MGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// a7/a8: concatenation building block
// ----------------------------------------------------------------------------------------------------
extern "C" int mgb_copy_rows(void* dst, int64_t dst_offset, const void* src, int64_t n_rows, void* stream) {
  MGB_REQUIRE(n_rows >= 0 && dst_offset >= 0, MGB_ERR_INVALID, "negative size");
  if (n_rows == 0) return MGB_OK;
  MGB_REQUIRE(dst && src, MGB_ERR_INVALID, "null pointer");
  MGB_CUDA(cudaMemcpyAsync((char*)dst + dst_offset * 8, src, (size_t)n_rows * 8, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return MGB_OK;
}


// ----------------------------------------------------------------------------------------------------
// f3, unfused forms: a derived column as an elementwise kernel, the row filter as a stable compaction.  The fused
// form lives in the dense batch kernel (mgb_dense.cuh: DenseProgram); these serve the shapes it does not take.
// ----------------------------------------------------------------------------------------------------
struct ProductArgs {
  const double* col[3];
  double alpha[3], beta[3];
  int nf;
};

__global__ void __launch_bounds__(256) k_eval_product(ProductArgs a, double* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = 1.0;
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      if (f < a.nf) {
        const double term = __dadd_rn(a.alpha[f], __dmul_rn(a.beta[f], a.col[f][i]));
        v = f == 0 ? term : __dmul_rn(v, term);
      }
    }
    out[i] = v;
  }
}

extern "C" int mgb_eval_product(const void* const* phys_cols, const mgb_affine_factor* factors, int32_t n_factors,
                                double* out, int64_t n_rows, void* stream) {
  MGB_REQUIRE(n_rows >= 0 && n_factors >= 1 && n_factors <= 3, MGB_ERR_INVALID, "bad argument");
  if (n_rows == 0) return MGB_OK;
  MGB_REQUIRE(phys_cols && factors && out, MGB_ERR_INVALID, "null pointer");
  ProductArgs a;
  memset(&a, 0, sizeof(a));
  a.nf = n_factors;
  for (int f = 0; f < n_factors; ++f) {
    MGB_REQUIRE(factors[f].col >= 0 && phys_cols[factors[f].col], MGB_ERR_INVALID, "factor %d: bad column", f);
    a.col[f] = (const double*)phys_cols[factors[f].col];
    a.alpha[f] = factors[f].alpha;
    a.beta[f] = factors[f].beta;
  }
  MgbProfScope prof(MGB_K_POST, (cudaStream_t)stream);
  k_eval_product<<<grid_for(n_rows), 256, 0, (cudaStream_t)stream>>>(a, out, n_rows);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

__global__ void __launch_bounds__(256) k_filter_pid(const uint64_t* col, int f64, int op, long long ci, double cd,
                                                    int32_t* pid, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t raw = mgb_ld_stream(col + i);
    int cmp;
    bool unordered = false;
    if (f64) {
      const double v = __longlong_as_double((long long)raw);
      unordered = v != v;
      cmp = v < cd ? -1 : (v > cd ? 1 : 0);
    } else {
      const long long v = (long long)raw;
      cmp = v < ci ? -1 : (v > ci ? 1 : 0);
    }
    bool keep = op == 0 ? cmp <= 0 : op == 1 ? cmp < 0 : op == 2 ? cmp >= 0 : op == 3 ? cmp > 0 : op == 4 ? cmp == 0 : cmp != 0;
    if (unordered) keep = op == 5;
    pid[i] = keep ? 0 : 1;
  }
}

extern "C" int mgb_filter_rows(const void* filter_col, int32_t filter_dtype, int32_t filter_op, int64_t filter_i64,
                               double filter_f64, const void* const* in_cols, void* const* out_cols, int32_t n_cols,
                               int64_t n_rows, int64_t* out_n_host, void* stream) {
  MGB_REQUIRE(out_n_host && n_rows >= 0 && n_cols >= 0 && filter_op >= 0 && filter_op <= 5, MGB_ERR_INVALID,
              "bad argument");
  *out_n_host = 0;
  if (n_rows == 0) return MGB_OK;
  MGB_REQUIRE(filter_col && (n_cols == 0 || (in_cols && out_cols)), MGB_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* pid = nullptr;
  int64_t* off = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&pid, sizeof(int32_t) * n_rows, st));
  MGB_CUDA(cudaMallocAsync((void**)&off, sizeof(int64_t) * 3, st));
  k_filter_pid<<<grid_for(n_rows), 256, 0, st>>>((const uint64_t*)filter_col, filter_dtype == MGB_F64, filter_op,
                                                  (long long)filter_i64, filter_f64, pid, n_rows);
  int status = mgb_partition_scatter(pid, n_rows, 2, in_cols, out_cols, n_cols, off, stream);
  int64_t h[3] = {0, 0, 0};
  if (status == MGB_OK) {
    cudaError_t e = cudaMemcpyAsync(h, off, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      mgb_set_error("row filter: %s", cudaGetErrorString(e));
      status = MGB_ERR_CUDA;
    }
  }
  cudaFreeAsync(pid, st);
  cudaFreeAsync(off, st);
  *out_n_host = h[1];
  return status;
}
