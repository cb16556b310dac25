// libmarsgb: device half of the narrow transfer encoding (DataFrameToGPU, mars/dataframe/base/to_gpu.py:28-35).
//
// A chunk that still lives in host memory crosses PCIe before the map stage can read it, and PCIe (~55 GB/s) is what
// bounds the end-to-end path.  Integer columns of real tables (flags, status codes, dictionary-encoded strings, dates)
// rarely need their 8 bytes, and float64 columns often hold small integers: the host side (mgb_hostpack.cpp) rewrites
// such a column as int8 / int16 / int32 -- every value checked, so the encoding is lossless or refused -- the narrow
// bytes cross the link, and the kernels here restore the 8-byte column in HBM, where every kernel of the path expects
// 8-byte elements.
#include <cstdint>

#include "mgb_common.cuh"

namespace {

template <typename T, int PER>
__global__ void k_widen_f64(const T* __restrict__ src, int64_t n, double* __restrict__ dst) {
  const int64_t groups = n / PER;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    T v[PER];
    if constexpr (sizeof(T) * PER == 8) {
      *reinterpret_cast<uint2*>(v) = __ldcs(reinterpret_cast<const uint2*>(src) + g);
    } else {
      *reinterpret_cast<uint4*>(v) = __ldcs(reinterpret_cast<const uint4*>(src) + g);
    }
    double2* out = reinterpret_cast<double2*>(dst + g * PER);
#pragma unroll
    for (int j = 0; j < PER / 2; ++j) out[j] = make_double2((double)v[2 * j], (double)v[2 * j + 1]);
  }
  const int64_t tail = groups * PER;
  for (int64_t i = tail + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (double)src[i];
}

template <typename T, int PER>
int launch_widen_f64(const void* src, int64_t n, double* dst, cudaStream_t st) {
  const int threads = 256;
  int64_t blocks = (n / PER + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_widen_f64<T, PER><<<(unsigned)blocks, threads, 0, st>>>(static_cast<const T*>(src), n, dst);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

template <typename T, int PER>
__global__ void k_widen(const T* __restrict__ src, int64_t n, int64_t* __restrict__ dst) {
  // PER consecutive elements per thread: one 8- or 16-byte load of narrow values, PER / 2 16-byte stores
  const int64_t groups = n / PER;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    T v[PER];
    if constexpr (sizeof(T) * PER == 8) {
      *reinterpret_cast<uint2*>(v) = __ldcs(reinterpret_cast<const uint2*>(src) + g);
    } else {
      *reinterpret_cast<uint4*>(v) = __ldcs(reinterpret_cast<const uint4*>(src) + g);
    }
    longlong2* out = reinterpret_cast<longlong2*>(dst + g * PER);
#pragma unroll
    for (int j = 0; j < PER / 2; ++j) out[j] = make_longlong2((long long)v[2 * j], (long long)v[2 * j + 1]);
  }
  const int64_t tail = groups * PER;
  for (int64_t i = tail + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (int64_t)src[i];
}

template <typename T, int PER>
int launch_widen(const void* src, int64_t n, int64_t* dst, cudaStream_t st) {
  const int threads = 256;
  int64_t blocks = (n / PER + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_widen<T, PER><<<(unsigned)blocks, threads, 0, st>>>(static_cast<const T*>(src), n, dst);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

}  // namespace

extern "C" int mgb_widen_i64(const void* src, int32_t width, int64_t n, int64_t* dst, void* stream) {
  MGB_REQUIRE(n >= 0, MGB_ERR_INVALID, "negative size");
  if (n == 0) return MGB_OK;
  MGB_REQUIRE(src && dst, MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, MGB_ERR_INVALID,
              "columns must start on a 16-byte boundary");
  cudaStream_t st = (cudaStream_t)stream;
  switch (width) {
    case 1: return launch_widen<int8_t, 8>(src, n, dst, st);
    case 2: return launch_widen<int16_t, 8>(src, n, dst, st);
    case 4: return launch_widen<int32_t, 4>(src, n, dst, st);
    default: MGB_REQUIRE(false, MGB_ERR_INVALID, "width %d: 1, 2 or 4 bytes", (int)width);
  }
  return MGB_OK;
}

extern "C" int mgb_widen_f64(const void* src, int32_t width, int64_t n, double* dst, void* stream) {
  MGB_REQUIRE(n >= 0, MGB_ERR_INVALID, "negative size");
  if (n == 0) return MGB_OK;
  MGB_REQUIRE(src && dst, MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, MGB_ERR_INVALID,
              "columns must start on a 16-byte boundary");
  cudaStream_t st = (cudaStream_t)stream;
  switch (width) {
    case 1: return launch_widen_f64<int8_t, 8>(src, n, dst, st);
    case 2: return launch_widen_f64<int16_t, 8>(src, n, dst, st);
    case 4: return launch_widen_f64<int32_t, 4>(src, n, dst, st);
    default: MGB_REQUIRE(false, MGB_ERR_INVALID, "width %d: 1, 2 or 4 bytes", (int)width);
  }
  return MGB_OK;
}
