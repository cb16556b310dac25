// libmarsgb: stable radix passes over a row permutation -> partition scatter (a5/a6) and key argsort.
//
// One primitive serves both users: a stable split of rows by an 8-bit digit
//   (tile histogram -> exclusive scan over [digit][tile] -> in-tile stable ranking with __match_any_sync).
// * mgb_partition_scatter: digit = reducer id (`hash % R`), one pass for R <= 256, two for R <= 65536;
//   replaces RangeIndex.groupby(hash % size) + iloc takes of DataFrameGroupByOperand.execute_map
//   (mars/dataframe/groupby/core.py:389-435).  Stability == the ascending positions `f_i` of the reference.
// * mgb_argsort_keys: LSD radix sort of int64 key tuples (groupby(sort=True) output order).
#include "mgb_common.cuh"

#define RP_THREADS 256
#define RP_ROUNDS 8
#define RP_TILE (RP_THREADS * RP_ROUNDS)
#define RP_WARPS (RP_THREADS / 32)
#define MGB_MAX_COLS 80

struct DigitSrc {
  const int32_t* pid;   // digit source A: int32 partition ids
  const int64_t* key;   // digit source B: int64 keys (order-preserving: sign bit flipped)
  const int32_t* perm;  // optional indirection (nullptr = identity)
  int shift;
};

__device__ __forceinline__ uint32_t rp_digit(const DigitSrc& s, int32_t elem) {
  if (s.pid) return ((uint32_t)s.pid[elem] >> s.shift) & 255u;
  uint64_t k = (uint64_t)s.key[elem] ^ MGB_SIGN64;
  return (uint32_t)(k >> s.shift) & 255u;
}

__global__ void __launch_bounds__(RP_THREADS) k_radix_hist(DigitSrc s, int64_t n, int64_t ntiles,
                                                           int64_t* hist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t tile = blockIdx.x;
  const int64_t base = tile * RP_TILE;
#pragma unroll
  for (int r = 0; r < RP_ROUNDS; ++r) {
    int64_t i = base + (int64_t)r * RP_THREADS + threadIdx.x;
    if (i < n) {
      int32_t elem = s.perm ? s.perm[i] : (int32_t)i;
      atomicAdd(&sh[rp_digit(s, elem)], 1u);
    }
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * ntiles + tile] = sh[threadIdx.x];
}

// after the exclusive scan: the pass is trivial (all rows share one digit) iff every bin start is 0 or n.
__global__ void k_radix_trivial(const int64_t* hist, int64_t ntiles, int64_t n, int32_t* flag) {
  int64_t start = hist[(int64_t)threadIdx.x * ntiles];
  int bad = (start != 0 && start != n);
  int any = __syncthreads_or(bad);
  if (threadIdx.x == 0) *flag = any ? 0 : 1;
}

// perm_out[output row] = input row: a permutation to gather by (more than 256 partitions, key argsort; the splits
// of <= 256 ways move the rows themselves, mgb_split.cu)
__global__ void __launch_bounds__(RP_THREADS) k_radix_scatter(DigitSrc s, int64_t n, int64_t ntiles,
                                                              const int64_t* hist, const int32_t* trivial,
                                                              int32_t* perm_out) {
  __shared__ int64_t base[256];
  __shared__ int64_t wpos[RP_WARPS][256];
  __shared__ uint16_t wc[RP_WARPS][256];
  const int64_t tile = blockIdx.x;
  const int64_t tbase = tile * RP_TILE;
  if (trivial && *trivial) {
#pragma unroll
    for (int r = 0; r < RP_ROUNDS; ++r) {
      int64_t i = tbase + (int64_t)r * RP_THREADS + threadIdx.x;
      if (i < n) perm_out[i] = s.perm ? s.perm[i] : (int32_t)i;
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = mgb_lanemask_lt();
  base[threadIdx.x] = hist[(int64_t)threadIdx.x * ntiles + tile];
  for (int r = 0; r < RP_ROUNDS; ++r) {
#pragma unroll
    for (int w = 0; w < RP_WARPS; ++w) wc[w][threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = tbase + (int64_t)r * RP_THREADS + threadIdx.x;
    const bool valid = i < n;
    int32_t elem = 0;
    uint32_t d = 256u + (uint32_t)lane;  // invalid lanes never match anybody
    if (valid) {
      elem = s.perm ? s.perm[i] : (int32_t)i;
      d = rp_digit(s, elem);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int lr = __popc(peers & lt);
    if (valid && lr == 0) wc[warp][d] = (uint16_t)__popc(peers);
    __syncthreads();
    {
      int64_t run = base[threadIdx.x];
#pragma unroll
      for (int w = 0; w < RP_WARPS; ++w) {
        wpos[w][threadIdx.x] = run;
        run += wc[w][threadIdx.x];
      }
      base[threadIdx.x] = run;
    }
    __syncthreads();
    if (valid) {
      perm_out[wpos[warp][d] + lr] = elem;
    }
    // wc is re-zeroed at the top of the next round, after which a barrier follows: wpos reads above are
    // separated from the next round's wpos writes by that barrier and the one after the match.
  }
}

// One stable pass.  scratch hist must hold 256*ntiles int64.
static int radix_pass(const DigitSrc& s, int64_t n, int32_t* perm_out, int64_t* hist, int32_t* trivial_flag,
                      cudaStream_t st) {
  const int64_t ntiles = (n + RP_TILE - 1) / RP_TILE;
  k_radix_hist<<<(unsigned)ntiles, RP_THREADS, 0, st>>>(s, n, ntiles, hist);
  MGB_CHECK_LAUNCH();
  MGB_TRY(mgb_scan_exclusive_i64(hist, 256 * ntiles, nullptr, st));
  if (trivial_flag) {
    k_radix_trivial<<<1, 256, 0, st>>>(hist, ntiles, n, trivial_flag);
    MGB_CHECK_LAUNCH();
  }
  k_radix_scatter<<<(unsigned)ntiles, RP_THREADS, 0, st>>>(s, n, ntiles, hist, trivial_flag, perm_out);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

__global__ void k_iota(int32_t* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = (int32_t)i;
}

static unsigned grid_for(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads, cap = (int64_t)mgb_sm_count() * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// --------------------------------------------------------------------------------------------------
// partition offsets: histogram of pid over R bins -> exclusive scan -> offsets[R+1]
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pid_hist(const int32_t* pid, int64_t n, int32_t R, int64_t* counts,
                                                  int32_t* bad) {
  extern __shared__ uint32_t sh[];
  const bool use_smem = R <= 8192;
  if (use_smem) {
    for (int i = threadIdx.x; i < R; i += blockDim.x) sh[i] = 0;
    __syncthreads();
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int32_t p = pid[i];
    if (p < 0 || p >= R) {
      *bad = 1;
      continue;
    }
    if (use_smem)
      atomicAdd(&sh[p], 1u);
    else
      atomicAdd((unsigned long long*)&counts[p], 1ULL);
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < R; i += blockDim.x)
      if (sh[i]) atomicAdd((unsigned long long*)&counts[i], (unsigned long long)sh[i]);
  }
}

int mgb_radix_sort_perm_pid_ex(const int32_t* pid, int64_t n, int32_t R, int32_t* perm_out, int64_t* offsets_out,
                               bool check_range, cudaStream_t st) {
  // offsets
  MGB_CUDA(cudaMemsetAsync(offsets_out, 0, sizeof(int64_t) * ((size_t)R + 1), st));
  if (n == 0) return MGB_OK;
  int32_t* bad = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&bad, sizeof(int32_t) * 2, st));
  MGB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t) * 2, st));
  {
    unsigned g = grid_for(n);
    if (g > (unsigned)mgb_sm_count() * 4) g = mgb_sm_count() * 4;
    size_t sm = R <= 8192 ? sizeof(uint32_t) * R : 0;
    k_pid_hist<<<g, 256, sm, st>>>(pid, n, R, offsets_out, bad);
    MGB_CHECK_LAUNCH();
  }
  MGB_TRY(mgb_scan_exclusive_i64(offsets_out, (int64_t)R + 1, nullptr, st));
  // permutation: stable LSD passes on the 8-bit digits of pid (1 pass for R <= 256, 2 for <= 65536, ...)
  const int64_t ntiles = (n + RP_TILE - 1) / RP_TILE;
  int64_t* hist = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&hist, sizeof(int64_t) * 256 * ntiles, st));
  int npass = 1;
  while (npass < 4 && ((int64_t)1 << (8 * npass)) < (int64_t)R) ++npass;
  int status = MGB_OK;
  int32_t* tmp = nullptr;
  if (npass > 1) MGB_CUDA(cudaMallocAsync((void**)&tmp, sizeof(int32_t) * n, st));
  int32_t* bufs[2] = {tmp, perm_out};
  const int32_t* cur = nullptr;
  for (int p = 0; p < npass && status == MGB_OK; ++p) {
    int32_t* dst = bufs[((npass - 1 - p) & 1) ? 0 : 1];   // the last pass lands in perm_out
    DigitSrc s{pid, nullptr, cur, 8 * p};
    status = radix_pass(s, n, dst, hist, nullptr, st);
    cur = dst;
  }
  if (tmp) cudaFreeAsync(tmp, st);
  cudaFreeAsync(hist, st);
  if (status != MGB_OK || !check_range) {
    cudaFreeAsync(bad, st);
    return status;
  }
  // surface out-of-range partition ids (needs one small sync; partition ids come from mgb_partition_ids)
  int32_t bad_h = 0;
  MGB_CUDA(cudaMemcpyAsync(&bad_h, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MGB_CUDA(cudaStreamSynchronize(st));
  cudaFreeAsync(bad, st);
  MGB_REQUIRE(bad_h == 0, MGB_ERR_INVALID, "partition id outside [0, %d)", R);
  return MGB_OK;
}

int mgb_radix_sort_perm_pid(const int32_t* pid, int64_t n, int32_t R, int32_t* perm_out, int64_t* offsets_out,
                            cudaStream_t st) {
  return mgb_radix_sort_perm_pid_ex(pid, n, R, perm_out, offsets_out, true, st);
}

int mgb_radix_sort_perm_keys(const int64_t* const* key_cols, int32_t n_keys, int64_t n, int32_t* perm_out,
                             cudaStream_t st) {
  if (n == 0) return MGB_OK;
  const int64_t ntiles = (n + RP_TILE - 1) / RP_TILE;
  int64_t* hist = nullptr;
  int32_t* tmp = nullptr;
  int32_t* flag = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&hist, sizeof(int64_t) * 256 * ntiles, st));
  MGB_CUDA(cudaMallocAsync((void**)&tmp, sizeof(int32_t) * n, st));
  MGB_CUDA(cudaMallocAsync((void**)&flag, sizeof(int32_t), st));
  // 8 passes per key, last key first; ping-pong so that the final pass lands in perm_out.
  const int total = 8 * n_keys;
  int32_t* bufs[2] = {tmp, perm_out};
  // pass p writes bufs[(p + total) & 1 ...]: choose so that the last pass (p = total-1) writes perm_out.
  int status = MGB_OK;
  const int32_t* cur = nullptr;
  for (int p = 0; p < total && status == MGB_OK; ++p) {
    const int key_idx = n_keys - 1 - p / 8;
    const int shift = (p % 8) * 8;
    int32_t* dst = bufs[((total - 1 - p) & 1) ? 0 : 1];
    DigitSrc s{nullptr, key_cols[key_idx], cur, shift};
    status = radix_pass(s, n, dst, hist, flag, st);
    cur = dst;
  }
  cudaFreeAsync(flag, st);
  cudaFreeAsync(tmp, st);
  cudaFreeAsync(hist, st);
  return status;
}

// --------------------------------------------------------------------------------------------------
// gather rows of many 8-byte columns by a permutation
// --------------------------------------------------------------------------------------------------
struct GatherCols {
  const uint64_t* in[MGB_MAX_COLS];
  uint64_t* out[MGB_MAX_COLS];
  int n;
};

__global__ void __launch_bounds__(256) k_gather(GatherCols gc, const int32_t* perm, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t src = perm[i];
    for (int c = 0; c < gc.n; ++c) gc.out[c][i] = gc.in[c][src];
  }
}

static int gather_impl(const void* const* in_cols, void* const* out_cols, int32_t n_cols, const int32_t* perm,
                       int64_t n, cudaStream_t st) {
  for (int c0 = 0; c0 < n_cols; c0 += MGB_MAX_COLS) {
    GatherCols gc;
    gc.n = n_cols - c0 < MGB_MAX_COLS ? n_cols - c0 : MGB_MAX_COLS;
    for (int c = 0; c < gc.n; ++c) {
      gc.in[c] = (const uint64_t*)in_cols[c0 + c];
      gc.out[c] = (uint64_t*)out_cols[c0 + c];
    }
    k_gather<<<grid_for(n), 256, 0, st>>>(gc, perm, n);
    MGB_CHECK_LAUNCH();
  }
  return MGB_OK;
}

extern "C" int mgb_gather_rows(const void* const* in_cols, void* const* out_cols, int32_t n_cols,
                               const int32_t* perm, int64_t n_rows, void* stream) {
  MGB_REQUIRE(n_rows >= 0 && n_cols >= 0, MGB_ERR_INVALID, "negative size");
  if (n_rows == 0 || n_cols == 0) return MGB_OK;
  MGB_REQUIRE(in_cols && out_cols && perm, MGB_ERR_INVALID, "null pointer");
  return gather_impl(in_cols, out_cols, n_cols, perm, n_rows, (cudaStream_t)stream);
}

extern "C" int mgb_argsort_keys(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows,
                                int32_t* out_perm, void* stream) {
  MGB_REQUIRE(n_keys >= 1 && n_keys <= 8, MGB_ERR_INVALID, "n_keys=%d", n_keys);
  MGB_REQUIRE(n_rows >= 0 && n_rows < 0x7fffffffLL, MGB_ERR_UNSUPPORTED, "n_rows=%lld needs int32 row ids",
              (long long)n_rows);
  if (n_rows == 0) return MGB_OK;
  MGB_REQUIRE(key_cols && out_perm, MGB_ERR_INVALID, "null pointer");
  MgbProfScope prof(MGB_K_SORT, (cudaStream_t)stream);
  return mgb_radix_sort_perm_keys(key_cols, n_keys, n_rows, out_perm, (cudaStream_t)stream);
}

extern "C" int mgb_partition_scatter(const int32_t* pid, int64_t n_rows, int32_t n_partitions,
                                     const void* const* in_cols, void* const* out_cols, int32_t n_cols,
                                     int64_t* out_offsets, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MGB_REQUIRE(n_partitions >= 1 && n_partitions <= 65536, MGB_ERR_UNSUPPORTED,
              "n_partitions=%d outside [1,65536]", n_partitions);
  MGB_REQUIRE(n_rows >= 0 && n_rows < 0x7fffffffLL, MGB_ERR_UNSUPPORTED, "n_rows=%lld needs int32 row ids",
              (long long)n_rows);
  MGB_REQUIRE(out_offsets, MGB_ERR_INVALID, "null out_offsets");
  MGB_REQUIRE(n_rows == 0 || (pid && (n_cols == 0 || (in_cols && out_cols))), MGB_ERR_INVALID, "null pointer");
  if (n_rows == 0) {
    MGB_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(int64_t) * ((size_t)n_partitions + 1), st));
    return MGB_OK;
  }
  if (n_partitions <= 256) {
    // histogram -> scan -> staged, vectorised scatter (mgb_split.cu); the range check of the caller's ids costs the
    // one small synchronisation this entry has always had
    int32_t* bad = nullptr;
    MGB_CUDA(cudaMallocAsync((void**)&bad, sizeof(int32_t), st));
    MGB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
    int status = mgb_split_local(nullptr, 0, 2, pid, n_rows, n_partitions, in_cols, out_cols, n_cols, out_offsets, bad, st);
    int32_t bad_h = 0;
    if (status == MGB_OK) {
      cudaError_t e = cudaMemcpyAsync(&bad_h, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        mgb_set_error("partition range check: %s", cudaGetErrorString(e));
        status = MGB_ERR_CUDA;
      }
    }
    cudaFreeAsync(bad, st);
    MGB_TRY(status);
    MGB_REQUIRE(bad_h == 0, MGB_ERR_INVALID, "partition id outside [0, %d)", n_partitions);
    return MGB_OK;
  }
  MgbProfScope prof(MGB_K_PARTITION, st);
  int32_t* perm = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&perm, sizeof(int32_t) * n_rows, st));
  int status = mgb_radix_sort_perm_pid(pid, n_rows, n_partitions, perm, out_offsets, st);
  if (status == MGB_OK && n_cols > 0) status = gather_impl(in_cols, out_cols, n_cols, perm, n_rows, st);
  cudaFreeAsync(perm, st);
  return status;
}


// a5 in one call: partition id = pandas hash of the key tuple % n_partitions (mgb_partition_ids), stable split
// of the rows into per-reducer blocks.  n_partitions <= 256: histogram -> scan -> staged, 128-bit vectorised scatter
// (mgb_split.cu: no partition-id array, no permutation, no gather, no host synchronisation); more partitions take
// the general radix passes over a permutation.
extern "C" int mgb_hash_partition(const int64_t* const* key_cols, int32_t n_keys, int64_t n_rows, int32_t n_partitions,
                                  const void* const* in_cols, void* const* out_cols, int32_t n_cols,
                                  int64_t* out_offsets, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MGB_REQUIRE(n_partitions >= 1 && n_partitions <= 65536, MGB_ERR_UNSUPPORTED, "n_partitions=%d outside [1,65536]",
              n_partitions);
  MGB_REQUIRE(n_rows >= 0 && n_rows < 0x7fffffffLL, MGB_ERR_UNSUPPORTED, "n_rows=%lld needs int32 row ids",
              (long long)n_rows);
  MGB_REQUIRE(out_offsets && n_cols >= 0, MGB_ERR_INVALID, "bad argument");
  if (n_rows == 0) {
    MGB_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(int64_t) * ((size_t)n_partitions + 1), st));
    return MGB_OK;
  }
  MGB_REQUIRE(key_cols && (n_cols == 0 || (in_cols && out_cols)), MGB_ERR_INVALID, "null pointer");
  if (n_partitions <= 256)
    return mgb_split_local(key_cols, n_keys, 0, nullptr, n_rows, n_partitions, in_cols, out_cols, n_cols, out_offsets,
                           nullptr, st);
  // general path (ids come from the hash kernel: no range check, no synchronisation)
  int32_t* pid = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&pid, sizeof(int32_t) * n_rows, st));
  int status = mgb_partition_ids(key_cols, n_keys, n_rows, n_partitions, pid, stream);
  if (status == MGB_OK) {
    MgbProfScope prof(MGB_K_PARTITION, st);
    int32_t* perm = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&perm, sizeof(int32_t) * n_rows, st);
    if (e != cudaSuccess) {
      mgb_set_error("cudaMallocAsync: %s", cudaGetErrorString(e));
      status = MGB_ERR_CUDA;
    } else {
      status = mgb_radix_sort_perm_pid_ex(pid, n_rows, n_partitions, perm, out_offsets, false, st);
      if (status == MGB_OK && n_cols > 0) status = gather_impl(in_cols, out_cols, n_cols, perm, n_rows, st);
      cudaFreeAsync(perm, st);
    }
  }
  cudaFreeAsync(pid, st);
  return status;
}
