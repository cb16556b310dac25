// libmarsgb: host-buffer entry point (end-to-end path).
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
