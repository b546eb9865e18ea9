// Streaming host copies of the staging engine (host_stage.cuh): plain C++ without any CUDA dependence, so the CPU test
// suite can compile and check them on their own (tests/test_capi_and_host.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define PVD_HOST_SSE2 1
#endif

namespace pvd {

#ifdef PVD_HOST_SSE2
// Streaming copies for the staging threads (SSE2 is part of x86-64): unaligned loads, 16-byte non-temporal stores once the
// destination is aligned, a store fence at the end (the DMA engine or another thread reads the lines next).
inline void stream_copy(void* dst, const void* src, size_t bytes) {
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    if (head > bytes) head = bytes;
    memcpy(d, s, head);
    d += head, s += head, bytes -= head;
    const size_t n64 = bytes / 64;
    for (size_t i = 0; i < n64; ++i, s += 64, d += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
    }
    memcpy(d, s, bytes - n64 * 64);
    _mm_sfence();
}
inline void stream_narrow(float* d, const double* s, size_t n) {  // float64 -> float32, round to nearest even like a C cast
    size_t i = 0;
    for (; i < n && (reinterpret_cast<uintptr_t>(d + i) & 15); ++i) d[i] = (float)s[i];
    for (; i + 4 <= n; i += 4) {
        const __m128 lo = _mm_cvtpd_ps(_mm_loadu_pd(s + i)), hi = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 2));
        _mm_stream_ps(d + i, _mm_movelh_ps(lo, hi));
    }
    for (; i < n; ++i) d[i] = (float)s[i];
    _mm_sfence();
}
inline void stream_widen(double* d, const float* s, size_t n) {  // float32 -> float64 (exact)
    size_t i = 0;
    for (; i < n && (reinterpret_cast<uintptr_t>(d + i) & 15); ++i) d[i] = (double)s[i];
    for (; i + 4 <= n; i += 4) {
        const __m128 v = _mm_loadu_ps(s + i);
        _mm_stream_pd(d + i, _mm_cvtps_pd(v));
        _mm_stream_pd(d + i + 2, _mm_cvtps_pd(_mm_movehl_ps(v, v)));
    }
    for (; i < n; ++i) d[i] = (double)s[i];
    _mm_sfence();
}
#else
inline void stream_copy(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }
inline void stream_narrow(float* d, const double* s, size_t n) {
    for (size_t i = 0; i < n; ++i) d[i] = (float)s[i];
}
inline void stream_widen(double* d, const float* s, size_t n) {
    for (size_t i = 0; i < n; ++i) d[i] = (double)s[i];
}
#endif

}  // namespace pvd
