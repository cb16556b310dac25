// libmarsgb: host-buffer entry point (end-to-end path) and atomic-throughput microbenchmarks.
#include <vector>

#include "mgb_common.cuh"

// ----------------------------------------------------------------------------------------------------
// mgb_groupby_agg_host: host columns -> (blocked H2D on a copy stream, overlapped with phase-1 kernels)
// -> aggregator -> D2H of the result.  This is what a CPU band hands to the library when the chunk still
// lives in host memory (the reference reads it from shared-memory storage, processor.py:147-175).
// ----------------------------------------------------------------------------------------------------
extern "C" int mgb_groupby_agg_host(const mgb_agg_spec* spec, const int64_t* const* key_cols_host,
                                    const void* const* val_cols_host, int64_t n_rows, int64_t block_rows,
                                    int64_t* const* out_key_cols_host, void* const* out_acc_cols_host,
                                    int64_t max_groups, int32_t sort_keys, int64_t* out_n_groups,
                                    void* stream) {
  MGB_REQUIRE(spec && key_cols_host && val_cols_host && out_n_groups, MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(n_rows >= 0 && max_groups >= 0, MGB_ERR_INVALID, "negative size");
  if (block_rows <= 0) block_rows = 4LL << 20;
  if (block_rows >= 0x7fffffffLL) block_rows = 0x7ffffff0LL;
  cudaStream_t st = (cudaStream_t)stream;
  const int nk = spec->n_keys, nv = spec->n_vals, na = spec->n_accs;
  MGB_REQUIRE(nk >= 1 && nk <= MGB_MAX_KEYS && nv >= 1 && nv <= MGB_MAX_VALS && na >= 1 && na <= MGB_MAX_ACCS,
              MGB_ERR_UNSUPPORTED, "spec outside the supported envelope");
  mgb_agg* agg = nullptr;
  MGB_TRY(mgb_agg_create(&agg, spec, st));
  int status = MGB_OK;
  uint64_t* dev = nullptr;
  uint64_t* dout = nullptr;
  cudaStream_t copy = nullptr;
  std::vector<cudaEvent_t> events;
  const int ncol = nk + nv;
  do {
    if (n_rows > 0) {
      if (cudaMallocAsync((void**)&dev, (size_t)n_rows * ncol * 8, st) != cudaSuccess ||
          cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking) != cudaSuccess) {
        mgb_set_error("device allocation of %lld bytes failed: %s", (long long)n_rows * ncol * 8,
                      cudaGetErrorString(cudaGetLastError()));
        status = MGB_ERR_CUDA;
        break;
      }
      cudaEvent_t ready;
      cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
      events.push_back(ready);
      cudaEventRecord(ready, st);
      cudaStreamWaitEvent(copy, ready, 0);
      for (int64_t r0 = 0; r0 < n_rows && status == MGB_OK; r0 += block_rows) {
        const int64_t nb = n_rows - r0 < block_rows ? n_rows - r0 : block_rows;
        const int64_t* kp[MGB_MAX_KEYS];
        const void* vp[MGB_MAX_VALS];
        for (int c = 0; c < ncol; ++c) {
          const char* src = c < nk ? (const char*)key_cols_host[c] : (const char*)val_cols_host[c - nk];
          uint64_t* dst = dev + (size_t)c * n_rows + r0;
          if (cudaMemcpyAsync(dst, src + r0 * 8, (size_t)nb * 8, cudaMemcpyHostToDevice, copy) != cudaSuccess) {
            mgb_set_error("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            status = MGB_ERR_CUDA;
            break;
          }
          if (c < nk)
            kp[c] = (const int64_t*)dst;
          else
            vp[c - nk] = dst;
        }
        if (status != MGB_OK) break;
        cudaEvent_t ev;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        events.push_back(ev);
        cudaEventRecord(ev, copy);
        cudaStreamWaitEvent(st, ev, 0);
        status = mgb_agg_update(agg, kp, vp, nb);
      }
      if (status != MGB_OK) break;
    }
    int64_t G = 0;
    status = mgb_agg_finalize(agg, &G);
    if (status != MGB_OK) break;
    *out_n_groups = G;
    if (G > max_groups) {
      mgb_set_error("%lld groups exceed the output capacity %lld", (long long)G, (long long)max_groups);
      status = MGB_ERR_OVERFLOW;
      break;
    }
    if (G == 0) break;
    if (!out_key_cols_host || !out_acc_cols_host) {
      mgb_set_error("null output pointer");
      status = MGB_ERR_INVALID;
      break;
    }
    if (cudaMallocAsync((void**)&dout, (size_t)G * (nk + na) * 8, st) != cudaSuccess) {
      mgb_set_error("output allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
      status = MGB_ERR_CUDA;
      break;
    }
    int64_t* ok[MGB_MAX_KEYS];
    void* oa[MGB_MAX_ACCS];
    for (int j = 0; j < nk; ++j) ok[j] = (int64_t*)(dout + (size_t)j * G);
    for (int j = 0; j < na; ++j) oa[j] = dout + (size_t)(nk + j) * G;
    status = mgb_agg_emit(agg, ok, oa, sort_keys);
    if (status != MGB_OK) break;
    for (int j = 0; j < nk && status == MGB_OK; ++j)
      if (cudaMemcpyAsync(out_key_cols_host[j], ok[j], (size_t)G * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        status = MGB_ERR_CUDA;
    for (int j = 0; j < na && status == MGB_OK; ++j)
      if (cudaMemcpyAsync(out_acc_cols_host[j], oa[j], (size_t)G * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        status = MGB_ERR_CUDA;
    if (status != MGB_OK) mgb_set_error("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
  } while (0);
  cudaError_t se = cudaStreamSynchronize(st);
  if (status == MGB_OK && se != cudaSuccess) {
    mgb_set_error("stream sync failed: %s", cudaGetErrorString(se));
    status = MGB_ERR_CUDA;
  }
  if (copy) cudaStreamSynchronize(copy);
  mgb_agg_destroy(agg);
  if (dev) cudaFreeAsync(dev, st);
  if (dout) cudaFreeAsync(dout, st);
  for (auto e : events) cudaEventDestroy(e);
  if (copy) cudaStreamDestroy(copy);
  return status;
}

// ----------------------------------------------------------------------------------------------------
// Microbenchmarks (profiles/): how fast can one B200 SM apply 8 accumulator updates per row?
//   mode 0: shared-memory atomicAdd(double) (ATOMS.CAST.SPIN loop), random slot in [0, n_slots)
//   mode 1: per-lane privatised LDS/DADD/STS (no atomics), n_slots <= 8
//   mode 2: global RED.ADD.F64, random slot in [0, n_slots)
//   mode 3: shared-memory atomicAdd(uint32) (native ATOMS.ADD), random slot
//   mode 4: warp-aggregated (match_any leader) then shared atomicAdd(double)
// Rows are synthesised in registers (slot = mix64(i) % n_slots) so only the update path is timed.
// ----------------------------------------------------------------------------------------------------
#define MB_ACCS 8

__global__ void __launch_bounds__(512, 1) k_mb(int mode, long long n, int n_slots, double* gacc, double* sink) {
  extern __shared__ __align__(16) uint64_t sm[];
  double* sd = (double*)sm;
  unsigned* su = (unsigned*)sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int words = mode == 1 ? (blockDim.x / 32) * n_slots * MB_ACCS * 32 : n_slots * MB_ACCS;
  for (int i = threadIdx.x; i < words; i += blockDim.x) sd[i] = 0.0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int slot = (int)(mgb_mix64((uint64_t)i) % (uint64_t)n_slots);
    const double v = 1.0 + (double)(i & 7);
    if (mode == 0) {
#pragma unroll
      for (int a = 0; a < MB_ACCS; ++a) atomicAdd(&sd[a * n_slots + slot], v);
    } else if (mode == 1) {
      double* p = sd + ((size_t)(warp * n_slots + slot) * MB_ACCS) * 32 + lane;
#pragma unroll
      for (int a = 0; a < MB_ACCS; ++a) p[a * 32] += v;
    } else if (mode == 2) {
#pragma unroll
      for (int a = 0; a < MB_ACCS; ++a) atomicAdd(&gacc[(size_t)slot * MB_ACCS + a], v);
    } else if (mode == 3) {
#pragma unroll
      for (int a = 0; a < MB_ACCS; ++a) atomicAdd(&su[a * n_slots + slot], 1u);
    } else {
      const unsigned peers = __match_any_sync(__activemask(), slot);
      const int leader = __ffs(peers) - 1;
      double s = v;
      // peers share the slot: sum their values with a butterfly over the whole warp restricted to peers
      for (int o = 16; o > 0; o >>= 1) {
        double y = __shfl_xor_sync(0xffffffffu, s, o);
        int ys = __shfl_xor_sync(0xffffffffu, slot, o);
        if (ys == slot) s += y;  // approximation of the peer reduction cost (exact when peers are contiguous)
      }
      if (lane == leader) {
#pragma unroll
        for (int a = 0; a < MB_ACCS; ++a) atomicAdd(&sd[a * n_slots + slot], s);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && sink) sink[blockIdx.x] = sd[0] + (double)su[1];
}

extern "C" int mgb_bench_atomics(int32_t mode, int64_t n_rows, int32_t n_slots, int32_t iters, float* out_ms,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MGB_REQUIRE(mode >= 0 && mode <= 4 && n_rows > 0 && n_slots > 0 && iters > 0 && out_ms, MGB_ERR_INVALID,
              "bad argument");
  const int threads = 512;
  size_t sm = mode == 1 ? (size_t)(threads / 32) * n_slots * MB_ACCS * 32 * 8 : (size_t)n_slots * MB_ACCS * 8;
  if (mode == 2) sm = 64;
  MGB_REQUIRE(sm <= 200 * 1024, MGB_ERR_INVALID, "n_slots too large for shared memory in mode %d", mode);
  if (sm < 64) sm = 64;
  MGB_CUDA(cudaFuncSetAttribute(k_mb, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  double *gacc = nullptr, *sink = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&gacc, (size_t)n_slots * MB_ACCS * 8, st));
  MGB_CUDA(cudaMemsetAsync(gacc, 0, (size_t)n_slots * MB_ACCS * 8, st));
  const int grid = mgb_sm_count();
  MGB_CUDA(cudaMallocAsync((void**)&sink, (size_t)grid * 8, st));
  cudaEvent_t e0, e1;
  MGB_CUDA(cudaEventCreate(&e0));
  MGB_CUDA(cudaEventCreate(&e1));
  k_mb<<<grid, threads, sm, st>>>(mode, n_rows, n_slots, gacc, sink);  // warm-up
  MGB_CUDA(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) k_mb<<<grid, threads, sm, st>>>(mode, n_rows, n_slots, gacc, sink);
  MGB_CUDA(cudaEventRecord(e1, st));
  MGB_CUDA(cudaEventSynchronize(e1));
  MGB_CHECK_LAUNCH();
  float ms = 0;
  MGB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *out_ms = ms / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFreeAsync(gacc, st);
  cudaFreeAsync(sink, st);
  return MGB_OK;
}
