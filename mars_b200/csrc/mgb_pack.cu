// libmarsgb: narrow transfer encoding of int64 columns (DataFrameToGPU, mars/dataframe/base/to_gpu.py:28-35).
//
// A chunk that still lives in host memory crosses PCIe before the map stage can read it, and PCIe (~55 GB/s) is what
// bounds the end-to-end path.  Integer columns of real tables (flags, status codes, dictionary-encoded strings, dates)
// rarely need their 8 bytes: the host side rewrites such a column as int8 / int16 / int32 -- range-checked, so the
// encoding is lossless or refused -- the narrow bytes cross the link, and mgb_widen_i64 restores the int64 column in
// HBM, where every kernel of the path expects 8-byte elements.  The host loop is plain C++ (the compiler vectorises it);
// callers split a column over a few threads (ctypes releases the GIL).
#include <cstdint>
#include <cstring>
#include <limits>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "mgb_common.cuh"

namespace {

template <typename T>
int narrow_block(const int64_t* __restrict__ src, int64_t n, T* __restrict__ dst) {
  // |bad| collects the bits that do not survive the round trip; one test per block keeps the loop branch-free
  constexpr int64_t BLOCK = 4096;
  for (int64_t a = 0; a < n; a += BLOCK) {
    const int64_t b = a + BLOCK < n ? a + BLOCK : n;
    uint64_t bad = 0;
    for (int64_t i = a; i < b; ++i) {
      const int64_t v = src[i];
      const T t = (T)v;
      dst[i] = t;
      bad |= (uint64_t)(v ^ (int64_t)t);
    }
    if (bad) return 0;
  }
  return 1;
}

template <typename T>
int narrow_tail_f64(const double* __restrict__ src, int64_t a, int64_t b, T* __restrict__ dst) {
  constexpr double lo = (double)std::numeric_limits<T>::min(), hi = (double)std::numeric_limits<T>::max();
  uint64_t bad = 0;
  for (int64_t i = a; i < b; ++i) {
    const double v = src[i];
    const bool in = (v >= lo) & (v <= hi);
    const double vc = in ? v : 0.0;
    const int32_t t = (int32_t)vc;
    const double back = (double)t;
    uint64_t bv, bb;
    memcpy(&bv, &v, 8);
    memcpy(&bb, &back, 8);
    dst[i] = (T)t;
    bad |= (bv ^ bb);                   // out of range: vc = 0 -> back = +0.0, and v != +0.0 there
  }
  return bad == 0;
}

template <typename T>
int narrow_block_f64(const double* __restrict__ src, int64_t n, T* __restrict__ dst) {
  // integer-valued float64 columns (quantities, counts stored as floats): v -> (T)v is lossless when the value is in
  // range and converting back gives the same BITS (-0.0, fractions and NaN -- which fails the range test -- refuse)
  constexpr int64_t BLOCK = 4096;
  for (int64_t a = 0; a < n; a += BLOCK) {
    const int64_t b = a + BLOCK < n ? a + BLOCK : n;
    int64_t i = a;
#if defined(__SSE2__)
    // the compiler keeps the scalar cvttsd2si; two doubles per cvttpd2dq, eight values packed per store
    const __m128d lo = _mm_set1_pd((double)std::numeric_limits<T>::min());
    const __m128d hi = _mm_set1_pd((double)std::numeric_limits<T>::max());
    __m128i bad = _mm_setzero_si128();
    for (; i + 8 <= b; i += 8) {
      __m128i q[4];
      for (int j = 0; j < 4; ++j) {
        const __m128d v = _mm_loadu_pd(src + i + 2 * j);
        const __m128d in = _mm_and_pd(_mm_cmpge_pd(v, lo), _mm_cmple_pd(v, hi));
        const __m128d vc = _mm_and_pd(v, in);                          // out of range / NaN -> +0.0
        q[j] = _mm_cvttpd_epi32(vc);                                    // two int32 in the low half
        bad = _mm_or_si128(bad, _mm_xor_si128(_mm_castpd_si128(_mm_cvtepi32_pd(q[j])), _mm_castpd_si128(v)));
      }
      const __m128i w0 = _mm_unpacklo_epi64(q[0], q[1]), w1 = _mm_unpacklo_epi64(q[2], q[3]);   // 4 + 4 int32
      if constexpr (sizeof(T) == 4) {
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), w0);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 4), w1);
      } else {
        const __m128i h = _mm_packs_epi32(w0, w1);                      // 8 int16 (values are in range, or refused)
        if constexpr (sizeof(T) == 2) {
          _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), h);
        } else {
          _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i), _mm_packs_epi16(h, h));
        }
      }
    }
    if (_mm_movemask_epi8(_mm_cmpeq_epi8(bad, _mm_setzero_si128())) != 0xffff) return 0;
#endif
    if (!narrow_tail_f64<T>(src, i, b, dst)) return 0;
  }
  return 1;
}

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

extern "C" int mgb_host_narrow_i64(const int64_t* src, int64_t n, int32_t width, void* dst, int32_t* fits) {
  MGB_REQUIRE(fits && (n == 0 || (src && dst)), MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(n >= 0, MGB_ERR_INVALID, "negative size");
  switch (width) {
    case 1: *fits = narrow_block<int8_t>(src, n, static_cast<int8_t*>(dst)); break;
    case 2: *fits = narrow_block<int16_t>(src, n, static_cast<int16_t*>(dst)); break;
    case 4: *fits = narrow_block<int32_t>(src, n, static_cast<int32_t*>(dst)); break;
    default: MGB_REQUIRE(false, MGB_ERR_INVALID, "width %d: 1, 2 or 4 bytes", (int)width);
  }
  return MGB_OK;
}

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

extern "C" int mgb_host_narrow_f64(const double* src, int64_t n, int32_t width, void* dst, int32_t* fits) {
  MGB_REQUIRE(fits && (n == 0 || (src && dst)), MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(n >= 0, MGB_ERR_INVALID, "negative size");
  switch (width) {
    case 1: *fits = narrow_block_f64<int8_t>(src, n, static_cast<int8_t*>(dst)); break;
    case 2: *fits = narrow_block_f64<int16_t>(src, n, static_cast<int16_t*>(dst)); break;
    case 4: *fits = narrow_block_f64<int32_t>(src, n, static_cast<int32_t*>(dst)); break;
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
