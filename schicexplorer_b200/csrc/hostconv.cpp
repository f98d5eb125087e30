// hostconv.cpp -- host-side staging helpers of schic_mh_fit_csr_host (capi.cu): narrowing of scipy's float64 / float32
// values to the wire formats of the upload (uint8 / uint16 counts, float32) and plain copies into page-locked staging
// memory. Pure host loops, called from a small pool of worker threads, one chunk per call.
//
// The narrowing loops are written with AVX2 intrinsics (runtime dispatch, scalar loop elsewhere): the compiler's own
// vectorisation of the clamp / cast / compare-back loop ran at 4.4 GB/s of float64 per thread, a third of what one
// thread copies -- 16 threads then need 16 ms for the 1.15 GB of values of the benchmark's 10 000 cells, most of the
// class call's fit. With the explicit version a thread converts about as fast as it can read.
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define SCHIC_X86 1
#else
#define SCHIC_X86 0
#endif

namespace schic {

#define SCHIC_CLONES __attribute__((target_clones("default", "avx2", "arch=x86-64-v4")))

// dst[i] = (T)src[i]; returns false (dst unspecified) unless every value is an integer in [0, max(T)].
template <typename S, typename T>
static inline bool narrow_exact(const S* __restrict__ src, int64_t n, T* __restrict__ dst) {
    unsigned bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        const S x = src[i];
        const S c = x < (S)0 ? (S)0 : (x > (S)(T)~(T)0 ? (S)(T)~(T)0 : x);     // clamp: the cast below stays defined
        const T v = (T)c;
        dst[i] = v;
        bad |= (S)v != x;                                                       // NaN, fractions, out of range
    }
    return bad == 0;
}

#if SCHIC_X86
static bool has_avx2() {
    static const bool v = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("sse4.1");
    return v;
}

// 16 values per step: truncating conversion to int32, converted back and compared with the source (NaN, fractions and
// anything beyond int32 fail: an out-of-range conversion yields INT32_MIN), all int32 results OR-ed for the range test
// (a negative value or one above the mask sets a high bit), saturating packs to the destination width.
__attribute__((target("avx2,sse4.1")))
static bool narrow_f64_avx2(const double* __restrict__ s, int64_t n, void* __restrict__ dv, bool to_u16) {
    __m256d eq = _mm256_castsi256_pd(_mm256_set1_epi32(-1));
    __m128i acc = _mm_setzero_si128();
    uint8_t* d8 = (uint8_t*)dv;
    uint16_t* d16 = (uint16_t*)dv;
    int64_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256d x0 = _mm256_loadu_pd(s + i), x1 = _mm256_loadu_pd(s + i + 4);
        const __m256d x2 = _mm256_loadu_pd(s + i + 8), x3 = _mm256_loadu_pd(s + i + 12);
        const __m128i v0 = _mm256_cvttpd_epi32(x0), v1 = _mm256_cvttpd_epi32(x1);
        const __m128i v2 = _mm256_cvttpd_epi32(x2), v3 = _mm256_cvttpd_epi32(x3);
        eq = _mm256_and_pd(eq, _mm256_and_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(v0), x0, _CMP_EQ_OQ),
                                             _mm256_cmp_pd(_mm256_cvtepi32_pd(v1), x1, _CMP_EQ_OQ)));
        eq = _mm256_and_pd(eq, _mm256_and_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(v2), x2, _CMP_EQ_OQ),
                                             _mm256_cmp_pd(_mm256_cvtepi32_pd(v3), x3, _CMP_EQ_OQ)));
        acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(v0, v1), _mm_or_si128(v2, v3)));
        const __m128i p01 = _mm_packus_epi32(v0, v1), p23 = _mm_packus_epi32(v2, v3);      // 8 x uint16 each
        if (to_u16) {
            _mm_storeu_si128((__m128i*)(d16 + i), p01);
            _mm_storeu_si128((__m128i*)(d16 + i + 8), p23);
        } else {
            _mm_storeu_si128((__m128i*)(d8 + i), _mm_packus_epi16(p01, p23));
        }
    }
    bool ok = _mm256_movemask_pd(eq) == 0xF && _mm_testz_si128(acc, _mm_set1_epi32(to_u16 ? (int)0xFFFF0000 : (int)0xFFFFFF00));
    if (i < n) ok = (to_u16 ? narrow_exact(s + i, n - i, d16 + i) : narrow_exact(s + i, n - i, d8 + i)) && ok;
    return ok;
}

__attribute__((target("avx2,sse4.1")))
static bool narrow_f32_avx2(const float* __restrict__ s, int64_t n, void* __restrict__ dv, bool to_u16) {
    __m256 eq = _mm256_castsi256_ps(_mm256_set1_epi32(-1));
    __m256i acc = _mm256_setzero_si256();
    uint8_t* d8 = (uint8_t*)dv;
    uint16_t* d16 = (uint16_t*)dv;
    int64_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256 x0 = _mm256_loadu_ps(s + i), x1 = _mm256_loadu_ps(s + i + 8);
        const __m256i v0 = _mm256_cvttps_epi32(x0), v1 = _mm256_cvttps_epi32(x1);
        eq = _mm256_and_ps(eq, _mm256_and_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(v0), x0, _CMP_EQ_OQ),
                                             _mm256_cmp_ps(_mm256_cvtepi32_ps(v1), x1, _CMP_EQ_OQ)));
        acc = _mm256_or_si256(acc, _mm256_or_si256(v0, v1));
        // (an int32 that passes the range test is < 2^16, hence exactly representable in float32: the compare-back is exact)
        const __m128i p01 = _mm_packus_epi32(_mm256_castsi256_si128(v0), _mm256_extracti128_si256(v0, 1));
        const __m128i p23 = _mm_packus_epi32(_mm256_castsi256_si128(v1), _mm256_extracti128_si256(v1, 1));
        if (to_u16) {
            _mm_storeu_si128((__m128i*)(d16 + i), p01);
            _mm_storeu_si128((__m128i*)(d16 + i + 8), p23);
        } else {
            _mm_storeu_si128((__m128i*)(d8 + i), _mm_packus_epi16(p01, p23));
        }
    }
    bool ok = _mm256_movemask_ps(eq) == 0xFF && _mm256_testz_si256(acc, _mm256_set1_epi32(to_u16 ? (int)0xFFFF0000 : (int)0xFFFFFF00));
    if (i < n) ok = (to_u16 ? narrow_exact(s + i, n - i, d16 + i) : narrow_exact(s + i, n - i, d8 + i)) && ok;
    return ok;
}
#endif

bool narrow_f64_u8(const double* s, int64_t n, uint8_t* d) {
#if SCHIC_X86
    if (has_avx2()) return narrow_f64_avx2(s, n, d, false);
#endif
    return narrow_exact(s, n, d);
}
bool narrow_f64_u16(const double* s, int64_t n, uint16_t* d) {
#if SCHIC_X86
    if (has_avx2()) return narrow_f64_avx2(s, n, d, true);
#endif
    return narrow_exact(s, n, d);
}
bool narrow_f32_u8(const float* s, int64_t n, uint8_t* d) {
#if SCHIC_X86
    if (has_avx2()) return narrow_f32_avx2(s, n, d, false);
#endif
    return narrow_exact(s, n, d);
}
bool narrow_f32_u16(const float* s, int64_t n, uint16_t* d) {
#if SCHIC_X86
    if (has_avx2()) return narrow_f32_avx2(s, n, d, true);
#endif
    return narrow_exact(s, n, d);
}
SCHIC_CLONES void convert_f64_f32(const double* __restrict__ s, int64_t n, float* __restrict__ d) {
    for (int64_t i = 0; i < n; ++i) d[i] = (float)s[i];
}
// 64-bit ids -> low words (+ "some id needs more than 32 bits")
SCHIC_CLONES bool narrow_ids64(const uint64_t* __restrict__ s, int64_t n, uint32_t* __restrict__ lo) {
    uint64_t any = 0;
    for (int64_t i = 0; i < n; ++i) { lo[i] = (uint32_t)s[i]; any |= s[i] >> 32; }
    return any != 0;
}

}  // namespace schic
