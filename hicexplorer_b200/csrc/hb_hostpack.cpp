// Host side of the narrow ingest (hb_xfer.cu: hb_ingest_values): fp64 counts -> uint16 while they are still in host
// memory, so that 2 bytes per entry cross PCIe instead of 8.  Lossless or nothing: a block is converted only if every
// value is an integer in [0, 65535] with the bit pattern of the widened result equal to the input (-0.0, NaN, Inf and
// fractions all fail), and the caller then falls back to the plain fp64 copy.  Plain C++ (g++), AVX2 when the CPU has it.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define HB_X86 1
#endif

namespace {

bool pack_scalar(const double* src, uint16_t* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const double v = src[i];
        if (!(v >= 0.0 && v <= 65535.0)) return false;   // NaN fails both comparisons
        const uint32_t t = (uint32_t)v;
        const double back = (double)t;
        uint64_t a, b;
        memcpy(&a, &v, 8);
        memcpy(&b, &back, 8);
        if (a != b) return false;                        // fraction, or -0.0
        dst[i] = (uint16_t)t;
    }
    return true;
}

#ifdef HB_X86
__attribute__((target("avx2"))) bool pack_avx2(const double* src, uint16_t* dst, size_t n) {
    const __m128i hi_mask = _mm_set1_epi32((int)0xFFFF0000u);
    size_t i = 0;
    __m256i all_eq = _mm256_set1_epi64x(-1);
    __m128i any_hi = _mm_setzero_si128();
    for (; i + 16 <= n; i += 16) {
        const __m256d v0 = _mm256_loadu_pd(src + i), v1 = _mm256_loadu_pd(src + i + 4);
        const __m256d v2 = _mm256_loadu_pd(src + i + 8), v3 = _mm256_loadu_pd(src + i + 12);
        const __m128i i0 = _mm256_cvttpd_epi32(v0), i1 = _mm256_cvttpd_epi32(v1);
        const __m128i i2 = _mm256_cvttpd_epi32(v2), i3 = _mm256_cvttpd_epi32(v3);
        // bit-exact round trip: rejects fractions, -0.0 and everything cvttpd maps to the "indefinite" 0x80000000
        all_eq = _mm256_and_si256(all_eq, _mm256_cmpeq_epi64(_mm256_castpd_si256(_mm256_cvtepi32_pd(i0)), _mm256_castpd_si256(v0)));
        all_eq = _mm256_and_si256(all_eq, _mm256_cmpeq_epi64(_mm256_castpd_si256(_mm256_cvtepi32_pd(i1)), _mm256_castpd_si256(v1)));
        all_eq = _mm256_and_si256(all_eq, _mm256_cmpeq_epi64(_mm256_castpd_si256(_mm256_cvtepi32_pd(i2)), _mm256_castpd_si256(v2)));
        all_eq = _mm256_and_si256(all_eq, _mm256_cmpeq_epi64(_mm256_castpd_si256(_mm256_cvtepi32_pd(i3)), _mm256_castpd_si256(v3)));
        any_hi = _mm_or_si128(any_hi, _mm_and_si128(_mm_or_si128(_mm_or_si128(i0, i1), _mm_or_si128(i2, i3)), hi_mask));
        _mm_storeu_si128((__m128i*)(dst + i), _mm_packus_epi32(i0, i1));
        _mm_storeu_si128((__m128i*)(dst + i + 8), _mm_packus_epi32(i2, i3));
    }
    if (_mm256_movemask_epi8(all_eq) != -1 || !_mm_testz_si128(any_hi, any_hi)) return false;
    return pack_scalar(src + i, dst + i, n - i);
}
#endif

}  // namespace

extern "C" int hb_host_pack_u16(const double* src, uint16_t* dst, size_t n) {
#ifdef HB_X86
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return pack_avx2(src, dst, n) ? 1 : 0;
#endif
    return pack_scalar(src, dst, n) ? 1 : 0;
}
