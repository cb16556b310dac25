// libmarsgb: the shuffle's partition kernel (a5 / a6) -- per-destination histogram, prefix scan, coalesced
// 128-bit vectorised scatter, optionally straight into the receive arenas of peer GPUs.
//
// Replaces, for one mapper chunk, `hash_dataframe_on(df, on, size)` + the K `iloc[f_i]` takes of
// DataFrameGroupByOperand.execute_map (mars/dataframe/groupby/core.py:389-435) -- and, when the destination
// table points into IPC-mapped arenas of the reducers' GPUs, also the storage-service transfer of the blocks
// (mars/services/storage/transfer.py:63-206): a row is read once from this GPU's HBM and written once into the
// HBM of the GPU that reduces it.
//
//   pass 1  k_split_hist     tiles of 2048 rows: partition id of every row (pandas-exact hash % R, or a given id
//                            array), per-tile histogram -> hist[R][ntiles]; exclusive scan -> every (partition,
//                            tile) pair knows where its rows go; offsets[R + 1] = block boundaries
//   pass 2  k_split_scatter  per tile: stable rank of every row inside its partition (warp match + per-warp
//                            counters, no atomics), then column by column: the column tile is staged in shared
//                            memory by the TMA unit (cp.async.bulk + mbarrier, double buffered: column c+1 is in
//                            flight while column c is handled), permuted in shared memory into partition order,
//                            and written out run by run -- consecutive lanes write consecutive addresses, two rows
//                            per 16-byte store (runs are laid out in shared memory with the parity of their
//                            destination so that aligned pairs in shared memory are aligned pairs in memory).
//
// Stable: inside a partition rows keep their input order (the ascending positions f_i of the reference).
// No host synchronisation; 1 <= R <= 256 (more partitions: the radix passes of mgb_partition.cu).
#include <stdlib.h>

#include "mgb_split.cuh"

struct SplitSrc {
  const int64_t* k[MGB_MAX_KEYS];
  const int32_t* pid;   // MODE 2: partition ids given by the caller
  uint64_t R;
};

struct SplitCols {
  const uint64_t* in[SP_MAX_COLS];
  uint64_t* out[SP_MAX_COLS];   // local mode: out[c] + (absolute output row)
  int n;
};

// MODE 0: hash of the key tuple with Index semantics (mgb_partition_ids); 1: DataFrame-row semantics
// (mgb_partition_ids_frame); 2: ids given (range partition of the PSRS shuffle, mgb_range_partition_ids)
template <int NK, int MODE>
__device__ __forceinline__ uint32_t sp_pid(const SplitSrc& s, int64_t i, bool pow2, int32_t* bad) {
  if (MODE == 2) {
    uint32_t p = (uint32_t)s.pid[i];
    if (p >= (uint32_t)s.R) {
      *bad = 1;
      p = (uint32_t)s.R - 1;
    }
    return p;
  }
  int64_t k[NK];
#pragma unroll
  for (int j = 0; j < NK; ++j) k[j] = (int64_t)mgb_ld_stream((const uint64_t*)s.k[j] + i);
  const uint64_t h = MODE == 1 ? mgb_hash_frame_row<NK>(k) : mgb_hash_tuple<NK>(k);
  return (uint32_t)(pow2 ? (h & (s.R - 1)) : (h % s.R));
}

// counting needs no order: one shared-memory atomic per row (the stable ranks are pass 2's business)
template <int NK, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) k_split_hist(SplitSrc s, int64_t n, int64_t ntiles, int R,
                                                        int64_t* __restrict__ hist, int32_t* bad) {
  __shared__ uint32_t cnt[SP_MAX_R];
  if (threadIdx.x < SP_MAX_R) cnt[threadIdx.x] = 0;
  __syncthreads();
  const bool pow2 = (s.R & (s.R - 1)) == 0;
  const int64_t tile = blockIdx.x;
  const int64_t base = tile * (THREADS * SP_ROUNDS) + threadIdx.x;
  uint32_t dg[SP_ROUNDS];
#pragma unroll
  for (int r = 0; r < SP_ROUNDS; ++r) {   // all key loads of the thread in flight together
    const int64_t i = base + r * THREADS;
    dg[r] = i < n ? sp_pid<NK, MODE>(s, i, pow2, bad) : 0xffffffffu;
  }
#pragma unroll
  for (int r = 0; r < SP_ROUNDS; ++r)
    if (dg[r] != 0xffffffffu) atomicAdd(&cnt[dg[r]], 1u);
  __syncthreads();
  for (int p = threadIdx.x; p < R; p += THREADS) hist[(int64_t)p * ntiles + tile] = cnt[p];
}

// block boundaries out of the scanned histogram
__global__ void k_split_offsets(const int64_t* hist, int64_t ntiles, int R, int64_t n, int64_t* offsets) {
  for (int p = threadIdx.x; p <= R; p += blockDim.x) offsets[p] = p < R ? hist[(int64_t)p * ntiles] : n;
}

template <int NK, int MODE>
struct ShuffleTile {
  const SplitSrc& s;
  const SplitCols& cols;
  const int64_t* __restrict__ hist;
  const uint64_t* __restrict__ dst_tab;
  int64_t ntiles, tile, tbase;
  int R;
  bool pow2;
  __device__ __forceinline__ uint32_t digit(int j) const {
    int32_t unused;
    return sp_pid<NK, MODE>(s, tbase + j, pow2, &unused);
  }
  __device__ __forceinline__ int64_t dest_row(int d) const {
    const int64_t abs0 = hist[(int64_t)d * ntiles + tile];
    return dst_tab ? abs0 - hist[(int64_t)d * ntiles] : abs0;
  }
  __device__ __forceinline__ const uint64_t* src(int c) const { return cols.in[c] + tbase; }
  __device__ __forceinline__ uint64_t dst(int c, int d) const {
    return dst_tab ? dst_tab[(int64_t)c * R + d] : (uint64_t)cols.out[c];
  }
};

// THREADS = 256, NBUF = 2: 2048-row tiles, three CTAs per SM (the default).  THREADS = 1024, NBUF = 1: 8192-row tiles, one
// CTA per SM -- the runs a tile leaves are four times longer, which is what stores that cross NVLink want at many
// partitions (tools/push_microbench.py: 256-byte runs 455 GB/s, 2 KB runs 614 GB/s per GPU)
template <int NK, int MODE, int THREADS, int NBUF>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 3 : 1)
    k_split_scatter(const SplitSrc s, const SplitCols cols, int64_t n, int64_t ntiles, int R,
                    const int64_t* __restrict__ hist, const uint64_t* __restrict__ dst_tab) {
  extern __shared__ __align__(128) unsigned char sp_raw[];
  SplitSmemT<THREADS, NBUF>& sm = *reinterpret_cast<SplitSmemT<THREADS, NBUF>*>(sp_raw);
  constexpr int TILE = THREADS * SP_ROUNDS;
  const int64_t tile = blockIdx.x, tbase = tile * TILE;
  const int nt = (int)((n - tbase) < TILE ? (n - tbase) : TILE);
  ShuffleTile<NK, MODE> f{s, cols, hist, dst_tab, ntiles, tile, tbase, R, (s.R & (s.R - 1)) == 0};
  sp_tile_split_t<true, THREADS, NBUF>(sm, nt, R, cols.n, f);
}

// tile geometry of a split: big tiles when the rows go to many partitions (both passes must agree, so the rule
// depends on nothing but the partition count and the chunk's length)
#define SP_BIG_THREADS 1024
// long_runs (mode bit 2, MGB_SPLIT_LONG_RUNS): the caller's destinations are peer GPUs.  Measured at N = 4
// (tools/push_microbench.py, per GPU out): R = 8: 614 -> 628 GB/s, R = 64: 453 -> 547, R = 256: 238 -> 427.  Local
// destinations: R = 64 is 6 % slower with big tiles (one CTA per SM), R = 256 9 % faster.
static bool split_big(int64_t n_rows, int32_t R, bool long_runs) {
  static const char* env = getenv("MGB_SPLIT_BIG");   // experiments: 0 = never, 1 = always
  if (env && *env == '0') return false;
  if (env && *env == '1') return true;
  if (n_rows < (1 << 18)) return false;                // short chunks: keep the grid wide
  return long_runs || R >= 128;
}
static int64_t split_tile_rows(int64_t n_rows, int32_t R, bool long_runs) {
  return (int64_t)(split_big(n_rows, R, long_runs) ? SP_BIG_THREADS : SP_THREADS) * SP_ROUNDS;
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
static int split_check(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid, int64_t n_rows,
                       int32_t R) {
  MGB_REQUIRE(mode >= 0 && mode <= 2, MGB_ERR_INVALID, "mode=%d", mode);
  MGB_REQUIRE(R >= 1 && R <= SP_MAX_R, MGB_ERR_UNSUPPORTED, "n_partitions=%d outside [1,%d]", R, SP_MAX_R);
  MGB_REQUIRE(n_rows >= 0 && n_rows < 0x7fffffffLL, MGB_ERR_UNSUPPORTED, "n_rows=%lld", (long long)n_rows);
  if (mode == 2) {
    MGB_REQUIRE(pid || n_rows == 0, MGB_ERR_INVALID, "null partition ids");
  } else {
    MGB_REQUIRE(n_keys >= 1 && n_keys <= MGB_MAX_KEYS, MGB_ERR_UNSUPPORTED, "n_keys=%d outside [1,%d]", n_keys,
                MGB_MAX_KEYS);
    MGB_REQUIRE(key_cols || n_rows == 0, MGB_ERR_INVALID, "null key columns");
  }
  return MGB_OK;
}

static SplitSrc split_src(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid, int32_t R) {
  SplitSrc s;
  for (int j = 0; j < MGB_MAX_KEYS; ++j) s.k[j] = (mode != 2 && j < n_keys) ? key_cols[j] : nullptr;
  s.pid = pid;
  s.R = (uint64_t)R;
  return s;
}

#define SP_DISPATCH(KERNEL, GEO, ...)                                \
  do {                                                               \
    if (mode == 2) {                                                 \
      KERNEL<1, 2, GEO> __VA_ARGS__;                                 \
    } else if (mode == 1) {                                          \
      if (n_keys == 1) {                                             \
        KERNEL<1, 1, GEO> __VA_ARGS__;                               \
      } else {                                                       \
        KERNEL<2, 1, GEO> __VA_ARGS__;                               \
      }                                                              \
    } else {                                                         \
      if (n_keys == 1) {                                             \
        KERNEL<1, 0, GEO> __VA_ARGS__;                               \
      } else {                                                       \
        KERNEL<2, 0, GEO> __VA_ARGS__;                               \
      }                                                              \
    }                                                                \
  } while (0)
#define SP_GEO_SMALL SP_THREADS, 2
#define SP_GEO_BIG SP_BIG_THREADS, 1

extern "C" int mgb_split_tiles(int64_t n_rows, int32_t n_partitions, int32_t mode, int64_t* out_tiles) {
  MGB_REQUIRE(out_tiles && n_rows >= 0 && n_partitions >= 1, MGB_ERR_INVALID, "bad argument");
  const int64_t t = split_tile_rows(n_rows, n_partitions, (mode & MGB_SPLIT_LONG_RUNS) != 0);
  *out_tiles = (n_rows + t - 1) / t;
  return MGB_OK;
}

int mgb_split_plan_ex(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid, int64_t n_rows,
                      int32_t R, int64_t* tile_hist, int64_t* out_offsets, int32_t* bad, cudaStream_t st) {
  const bool long_runs = (mode & MGB_SPLIT_LONG_RUNS) != 0;
  mode &= 3;
  MGB_TRY(split_check(key_cols, n_keys, mode, pid, n_rows, R));
  MGB_REQUIRE(out_offsets, MGB_ERR_INVALID, "null out_offsets");
  if (n_rows == 0) {
    MGB_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(int64_t) * ((size_t)R + 1), st));
    return MGB_OK;
  }
  MGB_REQUIRE(tile_hist, MGB_ERR_INVALID, "null tile histogram");
  const int64_t trows = split_tile_rows(n_rows, R, long_runs);
  const int64_t ntiles = (n_rows + trows - 1) / trows;
  SplitSrc s = split_src(key_cols, n_keys, mode, pid, R);
  if (split_big(n_rows, R, long_runs))
    SP_DISPATCH(k_split_hist, SP_BIG_THREADS, <<<(unsigned)ntiles, SP_BIG_THREADS, 0, st>>>(s, n_rows, ntiles, R, tile_hist, bad));
  else
    SP_DISPATCH(k_split_hist, SP_THREADS, <<<(unsigned)ntiles, SP_THREADS, 0, st>>>(s, n_rows, ntiles, R, tile_hist, bad));
  MGB_CHECK_LAUNCH();
  MGB_TRY(mgb_scan_exclusive_i64(tile_hist, (int64_t)R * ntiles, nullptr, st));
  k_split_offsets<<<1, 288, 0, st>>>(tile_hist, ntiles, R, n_rows, out_offsets);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

extern "C" int mgb_split_plan(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid,
                              int64_t n_rows, int32_t n_partitions, int64_t* tile_hist, int64_t* out_offsets,
                              void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MGB_REQUIRE((mode & 3) != 2, MGB_ERR_UNSUPPORTED, "given partition ids are range-checked by mgb_partition_scatter");
  MgbProfScope prof(MGB_K_HASH, st);
  return mgb_split_plan_ex(key_cols, n_keys, mode, pid, n_rows, n_partitions, tile_hist, out_offsets, nullptr, st);
}

int mgb_split_scatter_ex(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid,
                         int64_t n_rows, int32_t R, const int64_t* tile_hist, const void* const* in_cols, int32_t n_cols,
                         void* const* out_cols, const uint64_t* dst_table, cudaStream_t st) {
  const bool long_runs = (mode & MGB_SPLIT_LONG_RUNS) != 0;
  mode &= 3;
  MGB_TRY(split_check(key_cols, n_keys, mode, pid, n_rows, R));
  MGB_REQUIRE(n_cols >= 0, MGB_ERR_INVALID, "negative column count");
  if (n_rows == 0 || n_cols == 0) return MGB_OK;
  MGB_REQUIRE(tile_hist && in_cols && (out_cols || dst_table), MGB_ERR_INVALID, "null pointer");
  static bool attr_set[8] = {false};
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  const size_t smem_small = sizeof(SplitSmemT<SP_THREADS, 2>), smem_big = sizeof(SplitSmemT<SP_BIG_THREADS, 1>);
  if (dev >= 0 && dev < 8 && !attr_set[dev]) {
#define SP_ATTR(NK, MODE)                                                                                            \
  MGB_CUDA(cudaFuncSetAttribute(k_split_scatter<NK, MODE, SP_GEO_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)smem_small));                                                                   \
  MGB_CUDA(cudaFuncSetAttribute(k_split_scatter<NK, MODE, SP_GEO_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                (int)smem_big))
    SP_ATTR(1, 0);
    SP_ATTR(2, 0);
    SP_ATTR(1, 1);
    SP_ATTR(2, 1);
    SP_ATTR(1, 2);
#undef SP_ATTR
    attr_set[dev] = true;
  }
  const bool big = split_big(n_rows, R, long_runs);
  const int64_t trows = split_tile_rows(n_rows, R, long_runs);
  const int64_t ntiles = (n_rows + trows - 1) / trows;
  SplitSrc s = split_src(key_cols, n_keys, mode, pid, R);
  for (int c0 = 0; c0 < n_cols; c0 += SP_MAX_COLS) {
    SplitCols sc;
    sc.n = n_cols - c0 < SP_MAX_COLS ? n_cols - c0 : SP_MAX_COLS;
    for (int c = 0; c < sc.n; ++c) {
      sc.in[c] = (const uint64_t*)in_cols[c0 + c];
      sc.out[c] = out_cols ? (uint64_t*)out_cols[c0 + c] : nullptr;
    }
    const uint64_t* tab = dst_table ? dst_table + (size_t)c0 * R : nullptr;
    if (big)
      SP_DISPATCH(k_split_scatter, SP_GEO_BIG, <<<(unsigned)ntiles, SP_BIG_THREADS, smem_big, st>>>(s, sc, n_rows, ntiles, R, tile_hist, tab));
    else
      SP_DISPATCH(k_split_scatter, SP_GEO_SMALL, <<<(unsigned)ntiles, SP_THREADS, smem_small, st>>>(s, sc, n_rows, ntiles, R, tile_hist, tab));
    MGB_CHECK_LAUNCH();
  }
  return MGB_OK;
}

extern "C" int mgb_split_scatter(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid,
                                 int64_t n_rows, int32_t n_partitions, const int64_t* tile_hist,
                                 const void* const* in_cols, int32_t n_cols, void* const* out_cols,
                                 const uint64_t* dst_table, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MgbProfScope prof(MGB_K_PARTITION, st);
  return mgb_split_scatter_ex(key_cols, n_keys, mode, pid, n_rows, n_partitions, tile_hist, in_cols, n_cols, out_cols,
                              dst_table, st);
}

// plan + scatter into local columns with a scratch histogram of its own (mgb_hash_partition / mgb_partition_scatter
// for n_partitions <= 256).  bad (device int32, may be null): set when a given partition id is out of range.
int mgb_split_local(const int64_t* const* key_cols, int32_t n_keys, int32_t mode, const int32_t* pid, int64_t n_rows,
                    int32_t R, const void* const* in_cols, void* const* out_cols, int32_t n_cols, int64_t* out_offsets,
                    int32_t* bad, cudaStream_t st) {
  MGB_TRY(split_check(key_cols, n_keys, mode & 3, pid, n_rows, R));
  if (n_rows == 0) {
    MGB_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(int64_t) * ((size_t)R + 1), st));
    return MGB_OK;
  }
  const int64_t trows = split_tile_rows(n_rows, R, (mode & MGB_SPLIT_LONG_RUNS) != 0);
  const int64_t ntiles = (n_rows + trows - 1) / trows;
  int64_t* hist = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&hist, sizeof(int64_t) * (size_t)R * ntiles, st));
  int status;
  {
    MgbProfScope prof(MGB_K_HASH, st);
    status = mgb_split_plan_ex(key_cols, n_keys, mode, pid, n_rows, R, hist, out_offsets, bad, st);
  }
  if (status == MGB_OK) {
    MgbProfScope prof(MGB_K_PARTITION, st);
    status = mgb_split_scatter_ex(key_cols, n_keys, mode, pid, n_rows, R, hist, in_cols, n_cols, out_cols, nullptr, st);
  }
  cudaFreeAsync(hist, st);
  return status;
}
