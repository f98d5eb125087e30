// hostconv.cpp -- host-side staging helpers of schic_mh_fit_csr_host (capi.cu): narrowing of scipy's float64 / float32
// values to the wire formats of the upload (uint8 / uint16 counts, float32) and plain copies into page-locked staging
// memory. Pure host loops; target_clones gives SSE2 / AVX2 / AVX-512 bodies with runtime dispatch, so one prebuilt
// library is safe on whatever CPU the GPU box has. Called from a small pool of worker threads, one chunk per call.
#include <cstdint>
#include <cstring>

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

SCHIC_CLONES bool narrow_f64_u8(const double* s, int64_t n, uint8_t* d) { return narrow_exact(s, n, d); }
SCHIC_CLONES bool narrow_f64_u16(const double* s, int64_t n, uint16_t* d) { return narrow_exact(s, n, d); }
SCHIC_CLONES bool narrow_f32_u8(const float* s, int64_t n, uint8_t* d) { return narrow_exact(s, n, d); }
SCHIC_CLONES bool narrow_f32_u16(const float* s, int64_t n, uint16_t* d) { return narrow_exact(s, n, d); }
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
