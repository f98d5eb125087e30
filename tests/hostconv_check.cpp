// Stand-alone checker of csrc/hostconv.cpp (built and run by tests/test_hostconv.py on the CPU): the AVX2 narrowing loops
// against a scalar statement of their contract -- returns true iff every value is an integer in [0, max(T)], and then
// dst[i] == (T)src[i] -- on random arrays of every length class with special values planted at random positions.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <random>
#include <vector>

namespace schic {
bool narrow_f64_u8(const double* s, int64_t n, uint8_t* d);
bool narrow_f64_u16(const double* s, int64_t n, uint16_t* d);
bool narrow_f32_u8(const float* s, int64_t n, uint8_t* d);
bool narrow_f32_u16(const float* s, int64_t n, uint16_t* d);
void convert_f64_f32(const double* s, int64_t n, float* d);
bool narrow_ids64(const uint64_t* s, int64_t n, uint32_t* lo);
}

template <typename S, typename T, typename F>
static int run(F fn, const char* name, std::mt19937_64& rng) {
    const double specials[] = {std::nan(""), INFINITY, -INFINITY, -1.0, 0.5, 254.5, 255.0, 256.0, 65535.0, 65536.0, -0.0, 1e30,
                               2147483648.0, 4294967301.0, -2147483648.0, 1e-30, 3.0000001, 70000.0};
    const double maxT = (double)std::numeric_limits<T>::max();
    int fails = 0;
    for (int rep = 0; rep < 4000; ++rep) {
        const int64_t n = rep < 64 ? rep : (int64_t)(rng() % 3000);
        std::vector<S> src(n);
        const int kind = rep % 4;                      // 0: small counts, 1: full range of T, 2: one special, 3: several specials
        for (auto& x : src) x = (S)(kind == 1 ? (double)(rng() % ((uint64_t)maxT + 1)) : (double)(rng() % 200));
        if (kind >= 2 && n > 0)
            for (int k = 0; k < (kind == 2 ? 1 : 5); ++k) src[rng() % n] = (S)specials[rng() % (sizeof(specials) / sizeof(double))];
        bool want = true;
        for (auto x : src) want = want && (x >= (S)0 && (double)x <= maxT && (double)x == std::floor((double)x));
        std::vector<T> dst(n + 1, (T)0x5A);
        const bool got = fn(src.data(), n, dst.data());
        bool same = got == want && dst[n] == (T)0x5A;  // the guard element behind the output is untouched
        if (got && want)
            for (int64_t i = 0; i < n; ++i) same = same && dst[i] == (T)src[i];
        if (!same && ++fails < 5) std::printf("%s: mismatch at rep %d (n=%lld, want %d got %d)\n", name, rep, (long long)n, (int)want, (int)got);
    }
    return fails;
}

int main() {
    std::mt19937_64 rng(20260319);
    int fails = 0;
    fails += run<double, uint8_t>(schic::narrow_f64_u8, "f64->u8", rng);
    fails += run<double, uint16_t>(schic::narrow_f64_u16, "f64->u16", rng);
    fails += run<float, uint8_t>(schic::narrow_f32_u8, "f32->u8", rng);
    fails += run<float, uint16_t>(schic::narrow_f32_u16, "f32->u16", rng);
    // the two plain converters
    std::vector<double> a(1001);
    for (size_t i = 0; i < a.size(); ++i) a[i] = (double)i * 1.25 - 7.0;
    std::vector<float> f(a.size());
    schic::convert_f64_f32(a.data(), (int64_t)a.size(), f.data());
    for (size_t i = 0; i < a.size(); ++i) fails += f[i] != (float)a[i];
    std::vector<uint64_t> ids(777);
    for (size_t i = 0; i < ids.size(); ++i) ids[i] = i * 7919u;
    std::vector<uint32_t> lo(ids.size());
    fails += schic::narrow_ids64(ids.data(), (int64_t)ids.size(), lo.data()) ? 1 : 0;
    ids[300] = (5ull << 32) | 17u;
    fails += schic::narrow_ids64(ids.data(), (int64_t)ids.size(), lo.data()) ? 0 : 1;
    fails += lo[300] != 17u;
    std::printf(fails ? "HOSTCONV_FAIL %d\n" : "HOSTCONV_OK\n", fails);
    return fails ? 1 : 0;
}
