// libmarsgb: HOST pass of the narrow transfer encoding (DataFrameToGPU, mars/dataframe/base/to_gpu.py:28-35; the
// device half -- k_widen -- is in mgb_pack.cu).  Plain C++ for the host compiler: no device code, no CUDA calls.
//
// An int64 column (or a float64 column that holds integers) is rewritten as int8 / int16 / int32.  Every value is
// checked: the encoding is lossless or the column is refused (*fits = 0; what was written is then garbage).  The pass
// is compute-bound with the baseline x86-64 instruction set (~12 GB/s per thread), so AVX2 bodies are selected at run
// time when the CPU has them; the SSE2 / scalar bodies are the fallback and handle the tails.
#include <cstdint>
#include <cstring>
#include <limits>

#if defined(__x86_64__)
#include <immintrin.h>
#define MGB_X86 1
#endif

#include "../../include/mgb.h"

void mgb_set_error(const char* fmt, ...);      // mgb_util.cu (thread-local message behind mgb_last_error)

namespace {

constexpr int64_t BLOCK = 4096;                // one refusal test per block keeps the loops branch-free

// ---- int64 -> intW --------------------------------------------------------------------------------------------
template <typename T>
inline bool narrow_i64_scalar(const int64_t* __restrict__ src, int64_t a, int64_t b, T* __restrict__ dst) {
  uint64_t bad = 0;                            // the bits that do not survive the round trip
  for (int64_t i = a; i < b; ++i) {
    const int64_t v = src[i];
    const T t = (T)v;
    dst[i] = t;
    bad |= (uint64_t)(v ^ (int64_t)t);
  }
  return bad == 0;
}

#if MGB_X86
// A value fits W bytes iff x = v ^ (v < 0 ? ~0 : 0) is below 2^(8W-1).  16 values per iteration: the low dwords of four
// 256-bit loads are gathered by permutevar8x32, then packed with signed saturation (exact for values that fit).
template <typename T>
__attribute__((target("avx2"))) inline bool narrow_i64_avx2(const int64_t* __restrict__ src, int64_t a, int64_t& i,
                                                            int64_t b, T* __restrict__ dst) {
  const __m256i zero = _mm256_setzero_si256();
  const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
  const __m128i shift = _mm_cvtsi32_si128((int)sizeof(T) * 8 - 1);
  __m256i bad = zero;
  for (i = a; i + 16 <= b; i += 16) {
    __m128i q[4];
    for (int j = 0; j < 4; ++j) {
      const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 4 * j));
      const __m256i x = _mm256_xor_si256(v, _mm256_cmpgt_epi64(zero, v));
      bad = _mm256_or_si256(bad, _mm256_srl_epi64(x, shift));
      q[j] = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(v, pick));
    }
    if constexpr (sizeof(T) == 4) {
      for (int j = 0; j < 4; ++j) _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 4 * j), q[j]);
    } else {
      const __m128i h0 = _mm_packs_epi32(q[0], q[1]), h1 = _mm_packs_epi32(q[2], q[3]);
      if constexpr (sizeof(T) == 2) {
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), h0);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 8), h1);
      } else {
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_packs_epi16(h0, h1));
      }
    }
  }
  return _mm256_testz_si256(bad, bad) != 0;
}
#endif

template <typename T>
int narrow_i64(const int64_t* src, int64_t n, T* dst, bool avx2) {
  for (int64_t a = 0; a < n; a += BLOCK) {
    const int64_t b = a + BLOCK < n ? a + BLOCK : n;
    int64_t i = a;
#if MGB_X86
    if (avx2 && !narrow_i64_avx2<T>(src, a, i, b, dst)) return 0;
#endif
    if (!narrow_i64_scalar<T>(src, i, b, dst)) return 0;
  }
  return 1;
}

// ---- float64 holding integers -> intW ------------------------------------------------------------------------
// v -> (T)v is lossless when v is in range and converting back gives the same BITS: -0.0 and fractions come back
// different, NaN fails the range test, and an out-of-range value is replaced by +0.0 before the conversion (it then
// differs from what comes back, too).
template <typename T>
inline bool narrow_f64_scalar(const double* __restrict__ src, int64_t a, int64_t b, T* __restrict__ dst) {
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
    bad |= (bv ^ bb);
  }
  return bad == 0;
}

#if MGB_X86
template <typename T>
inline bool narrow_f64_sse2(const double* __restrict__ src, int64_t a, int64_t& i, int64_t b, T* __restrict__ dst) {
  const __m128d lo = _mm_set1_pd((double)std::numeric_limits<T>::min());
  const __m128d hi = _mm_set1_pd((double)std::numeric_limits<T>::max());
  __m128i bad = _mm_setzero_si128();
  for (i = a; i + 8 <= b; i += 8) {
    __m128i q[4];
    for (int j = 0; j < 4; ++j) {
      const __m128d v = _mm_loadu_pd(src + i + 2 * j);
      const __m128d vc = _mm_and_pd(v, _mm_and_pd(_mm_cmpge_pd(v, lo), _mm_cmple_pd(v, hi)));
      q[j] = _mm_cvttpd_epi32(vc);                                        // two int32 in the low half
      bad = _mm_or_si128(bad, _mm_xor_si128(_mm_castpd_si128(_mm_cvtepi32_pd(q[j])), _mm_castpd_si128(v)));
    }
    const __m128i w0 = _mm_unpacklo_epi64(q[0], q[1]), w1 = _mm_unpacklo_epi64(q[2], q[3]);
    if constexpr (sizeof(T) == 4) {
      _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), w0);
      _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 4), w1);
    } else {
      const __m128i h = _mm_packs_epi32(w0, w1);
      if constexpr (sizeof(T) == 2) {
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), h);
      } else {
        _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i), _mm_packs_epi16(h, h));
      }
    }
  }
  return _mm_movemask_epi8(_mm_cmpeq_epi8(bad, _mm_setzero_si128())) == 0xffff;
}

template <typename T>
__attribute__((target("avx2"))) inline bool narrow_f64_avx2(const double* __restrict__ src, int64_t a, int64_t& i,
                                                            int64_t b, T* __restrict__ dst) {
  const __m256d lo = _mm256_set1_pd((double)std::numeric_limits<T>::min());
  const __m256d hi = _mm256_set1_pd((double)std::numeric_limits<T>::max());
  __m256i bad = _mm256_setzero_si256();
  for (i = a; i + 16 <= b; i += 16) {
    __m128i q[4];
    for (int j = 0; j < 4; ++j) {
      const __m256d v = _mm256_loadu_pd(src + i + 4 * j);
      const __m256d in = _mm256_and_pd(_mm256_cmp_pd(v, lo, _CMP_GE_OQ), _mm256_cmp_pd(v, hi, _CMP_LE_OQ));
      q[j] = _mm256_cvttpd_epi32(_mm256_and_pd(v, in));                   // four int32
      bad = _mm256_or_si256(bad, _mm256_xor_si256(_mm256_castpd_si256(_mm256_cvtepi32_pd(q[j])),
                                                  _mm256_castpd_si256(v)));
    }
    if constexpr (sizeof(T) == 4) {
      for (int j = 0; j < 4; ++j) _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 4 * j), q[j]);
    } else {
      const __m128i h0 = _mm_packs_epi32(q[0], q[1]), h1 = _mm_packs_epi32(q[2], q[3]);
      if constexpr (sizeof(T) == 2) {
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), h0);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 8), h1);
      } else {
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_packs_epi16(h0, h1));
      }
    }
  }
  return _mm256_testz_si256(bad, bad) != 0;
}
#endif

template <typename T>
int narrow_f64(const double* src, int64_t n, T* dst, bool avx2) {
  for (int64_t a = 0; a < n; a += BLOCK) {
    const int64_t b = a + BLOCK < n ? a + BLOCK : n;
    int64_t i = a;
#if MGB_X86
    if (avx2) {
      if (!narrow_f64_avx2<T>(src, a, i, b, dst)) return 0;
    } else if (!narrow_f64_sse2<T>(src, a, i, b, dst)) {
      return 0;
    }
#endif
    if (!narrow_f64_scalar<T>(src, i, b, dst)) return 0;
  }
  return 1;
}

int g_force_baseline = 0;                      // tests: run the fallback bodies on a CPU that has AVX2

bool use_avx2() {
#if MGB_X86
  static const bool has = __builtin_cpu_supports("avx2");
  return has && !g_force_baseline;
#else
  return false;
#endif
}

}  // namespace

#define HP_REQUIRE(cond, ...)     \
  do {                            \
    if (!(cond)) {                \
      mgb_set_error(__VA_ARGS__); \
      return MGB_ERR_INVALID;     \
    }                             \
  } while (0)

extern "C" int mgb_host_narrow_isa(int32_t force_baseline) {
  g_force_baseline = force_baseline != 0;
  return use_avx2() ? 2 : 1;                   // 2: AVX2 bodies, 1: SSE2 / scalar
}

extern "C" int mgb_host_narrow_i64(const int64_t* src, int64_t n, int32_t width, void* dst, int32_t* fits) {
  HP_REQUIRE(fits && (n == 0 || (src && dst)), "null pointer");
  HP_REQUIRE(n >= 0, "negative size");
  const bool avx2 = use_avx2();
  switch (width) {
    case 1: *fits = narrow_i64<int8_t>(src, n, static_cast<int8_t*>(dst), avx2); break;
    case 2: *fits = narrow_i64<int16_t>(src, n, static_cast<int16_t*>(dst), avx2); break;
    case 4: *fits = narrow_i64<int32_t>(src, n, static_cast<int32_t*>(dst), avx2); break;
    default: HP_REQUIRE(false, "width %d: 1, 2 or 4 bytes", (int)width);
  }
  return MGB_OK;
}

extern "C" int mgb_host_narrow_f64(const double* src, int64_t n, int32_t width, void* dst, int32_t* fits) {
  HP_REQUIRE(fits && (n == 0 || (src && dst)), "null pointer");
  HP_REQUIRE(n >= 0, "negative size");
  const bool avx2 = use_avx2();
  switch (width) {
    case 1: *fits = narrow_f64<int8_t>(src, n, static_cast<int8_t*>(dst), avx2); break;
    case 2: *fits = narrow_f64<int16_t>(src, n, static_cast<int16_t*>(dst), avx2); break;
    case 4: *fits = narrow_f64<int32_t>(src, n, static_cast<int32_t*>(dst), avx2); break;
    default: HP_REQUIRE(false, "width %d: 1, 2 or 4 bytes", (int)width);
  }
  return MGB_OK;
}
