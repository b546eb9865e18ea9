// Register-level DFT butterflies and the shared-memory Stockham stage engine.
//
// Everything here is plain fp32 CUDA-core arithmetic: the dose path is HBM-bound and needs no
// tensor cores (BASELINE.json north_star).  Twiddles inside a butterfly are compile-time
// constants (constexpr trig below); inter-stage twiddles come from a per-axis table of N-th
// roots of unity computed in double precision.
#pragma once
#include "pvd_common.cuh"
#include <type_traits>
#include <utility>

namespace pvd {

// ------------------------------------------------------------------ constexpr trig
constexpr double cx_sin_small(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
constexpr double cx_cos_small(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
struct CxCS {
    double c, s;
};
// cos/sin of 2*pi*num/den with exact integer octant reduction
constexpr CxCS cx_cossin2pi(long long num, long long den) {
    constexpr double kTwoPi = 6.283185307179586476925286766559;
    long long D = 8 * den;
    long long A = ((8 * num) % D + D) % D;
    double cs = 1.0, ss = 1.0;
    if (A > D / 2) {
        A = D - A;
        ss = -1.0;
    }
    if (A > D / 4) {
        A = D / 2 - A;
        cs = -1.0;
    }
    double c = 0, s = 0;
    if (A > D / 8) {
        long long Ap = D / 4 - A;
        double x = kTwoPi * double(Ap) / double(D);
        c = cx_sin_small(x);
        s = cx_cos_small(x);
    } else {
        double x = kTwoPi * double(A) / double(D);
        c = cx_cos_small(x);
        s = cx_sin_small(x);
    }
    return CxCS{cs * c, ss * s};
}
template <int NUM, int DEN>
struct TwC {
    static constexpr float c = (float)cx_cossin2pi(NUM, DEN).c;
    static constexpr float s = (float)cx_cossin2pi(NUM, DEN).s;
};

template <int I, int N, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// a * exp(DIR * 2*pi*i * NUM/DEN)
template <int NUM, int DEN, int DIR>
__host__ __device__ __forceinline__ float2 mulw(float2 a) {
    constexpr int n = ((NUM % DEN) + DEN) % DEN;
    if constexpr (n == 0) {
        return a;
    } else if constexpr (4 * n == DEN) {  // * (DIR * i)
        return make_float2(-DIR * a.y, DIR * a.x);
    } else if constexpr (2 * n == DEN) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (4 * n == 3 * DEN) {  // * (-DIR * i)
        return make_float2(DIR * a.y, -DIR * a.x);
    } else {
        constexpr float c = TwC<n, DEN>::c;
        constexpr float s = DIR * TwC<n, DEN>::s;
#if defined(__CUDA_ARCH__) && !defined(PVD_EMULATE) && PVD_F32X2 > 1
        // a*c + (-a.y*s, a.x*s): two scalar products feed one packed fma (3 instructions instead of 4); the rounding
        // differs from the scalar form only in which product is fused
        return __ffma2_rn(a, make_float2(c, c), make_float2(-a.y * s, a.x * s));
#else
        return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
#endif
    }
}

constexpr int first_factor(int r) {
    if (r % 4 == 0 && r > 4) return 4;
    if (r % 2 == 0 && r > 2) return 2;
    for (int p = 3; p * p <= r; p += 2)
        if (r % p == 0) return p;
    return r;  // prime (or 2 / 4 handled by specialisations)
}

// ------------------------------------------------------------------ in-register DFT of size R
// a[] in natural order in, natural order out.  DIR = -1 forward, +1 inverse (unnormalised).
template <int R, int DIR>
struct Dft {
    __host__ __device__ static __forceinline__ void run(float2 (&a)[R]) {
        constexpr int R1 = first_factor(R);
        if constexpr (R1 == R) {
            // odd prime: symmetric-pair form
            constexpr int H = (R - 1) / 2;
            float2 sp[H + 1], dm[H + 1];
            static_for<1, H + 1>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                sp[j] = cadd(a[j], a[R - j]);
                dm[j] = csub(a[j], a[R - j]);
            });
            float2 a0 = a[0];
            float2 x0 = a0;
            static_for<1, H + 1>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                x0 = cadd(x0, sp[j]);
            });
            a[0] = x0;
            static_for<1, H + 1>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                float2 A = a0, B = make_float2(0.f, 0.f);
                static_for<1, H + 1>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    constexpr float c = TwC<(j * k) % R, R>::c;
                    constexpr float s = TwC<(j * k) % R, R>::s;
#if defined(__CUDA_ARCH__) && !defined(PVD_EMULATE) && PVD_F32X2
                    A = __ffma2_rn(sp[j], make_float2(c, c), A);
                    B = __ffma2_rn(dm[j], make_float2(s, s), B);
#else
                    A.x += sp[j].x * c;
                    A.y += sp[j].y * c;
                    B.x += dm[j].x * s;
                    B.y += dm[j].y * s;
#endif
                });
                // X_k = A + DIR*i*B ; X_{R-k} = A - DIR*i*B
                a[k] = make_float2(A.x - DIR * B.y, A.y + DIR * B.x);
                a[R - k] = make_float2(A.x + DIR * B.y, A.y - DIR * B.x);
            });
        } else {
            constexpr int R2 = R / R1;
            float2 b[R];
            static_for<0, R2>([&](auto j2c) {
                constexpr int j2 = decltype(j2c)::value;
                float2 t[R1];
                static_for<0, R1>([&](auto j1c) {
                    constexpr int j1 = decltype(j1c)::value;
                    t[j1] = a[j1 * R2 + j2];
                });
                Dft<R1, DIR>::run(t);
                static_for<0, R1>([&](auto k1c) {
                    constexpr int k1 = decltype(k1c)::value;
                    b[k1 * R2 + j2] = mulw<j2 * k1, R, DIR>(t[k1]);
                });
            });
            static_for<0, R1>([&](auto k1c) {
                constexpr int k1 = decltype(k1c)::value;
                float2 u[R2];
                static_for<0, R2>([&](auto j2c) {
                    constexpr int j2 = decltype(j2c)::value;
                    u[j2] = b[k1 * R2 + j2];
                });
                Dft<R2, DIR>::run(u);
                static_for<0, R2>([&](auto k2c) {
                    constexpr int k2 = decltype(k2c)::value;
                    a[k1 + R1 * k2] = u[k2];
                });
            });
        }
    }
};
template <int DIR>
struct Dft<1, DIR> {
    __host__ __device__ static __forceinline__ void run(float2 (&)[1]) {}
};
template <int DIR>
struct Dft<2, DIR> {
    __host__ __device__ static __forceinline__ void run(float2 (&a)[2]) {
        float2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    }
};
template <int DIR>
struct Dft<4, DIR> {
    __host__ __device__ static __forceinline__ void run(float2 (&a)[4]) {
        float2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
        float2 t2 = cadd(a[1], a[3]), t3 = csub(a[1], a[3]);
        float2 t3r = make_float2(-DIR * t3.y, DIR * t3.x);  // t3 * (DIR*i)
        a[0] = cadd(t0, t2);
        a[1] = cadd(t1, t3r);
        a[2] = csub(t0, t2);
        a[3] = csub(t1, t3r);
    }
};

// ------------------------------------------------------------------ one Stockham stage in smem
// Data layout: element (idx, line) at buf[idx * LS + line]; L = 1 << Llog lines per tile.
// Stage with current stride s (product of previous radices), radix R (DIF autosort):
//   Y[q + s*(R*p + k)] = ( sum_j X[q + s*(p + m*j)] w_R^{jk} ) * W_N^{p*s*k},  m = N/(s*R)
// tw[i] = exp(-2*pi*i * i/N) (forward sign); the inverse conjugates.
template <int R, int DIR>
__device__ __forceinline__ void stockham_stage(const float2* __restrict__ X, float2* __restrict__ Y,
                                               const float2* __restrict__ tw, int N, int s, int Llog, int LS) {
    const int nb = N / R;
    const int m = nb / s;
    const int total = nb << Llog;
    const int Lmask = (1 << Llog) - 1;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int line = i & Lmask;
        const int b = i >> Llog;
        const int p = b / s;
        const int q = b - p * s;
        float2 a[R];
        const float2* src = X + (q + s * p) * LS + line;
        const int jstride = s * m * LS;
        PVD_UNROLL
        for (int j = 0; j < R; ++j) a[j] = src[j * jstride];
        Dft<R, DIR>::run(a);
        if (p != 0) {
            const int tstep = p * s;
            PVD_UNROLL
            for (int k = 1; k < R; ++k) {
                float2 w = __ldg(&tw[tstep * k]);
                if (DIR > 0) w.y = -w.y;
                a[k] = cmul(a[k], w);
            }
        }
        float2* dst = Y + (q + s * R * p) * LS + line;
        const int kstride = s * LS;
        PVD_UNROLL
        for (int k = 0; k < R; ++k) dst[k * kstride] = a[k];
    }
}

// Fallback for any other prime radix r: O(r^2), one output per work item.
template <int DIR>
__device__ __forceinline__ void stockham_stage_generic(const float2* __restrict__ X, float2* __restrict__ Y,
                                                       const float2* __restrict__ tw, int N, int s, int r, int Llog,
                                                       int LS) {
    const int nb = N / r;
    const int m = nb / s;
    const int total = (nb * r) << Llog;
    const int Lmask = (1 << Llog) - 1;
    const int wr = N / r;  // w_r = W_N^(N/r)
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int line = i & Lmask;
        int t = i >> Llog;
        const int k = t % r;
        const int b = t / r;
        const int p = b / s;
        const int q = b - p * s;
        const float2* src = X + (q + s * p) * LS + line;
        const int jstride = s * m * LS;
        float2 acc = make_float2(0.f, 0.f);
        int e = 0;  // (j*k) mod r
        for (int j = 0; j < r; ++j) {
            float2 w = __ldg(&tw[e * wr]);
            if (DIR > 0) w.y = -w.y;
            float2 v = src[j * jstride];
            acc.x += v.x * w.x - v.y * w.y;
            acc.y += v.x * w.y + v.y * w.x;
            e += k;
            if (e >= r) e -= r;
        }
        if (p != 0 && k != 0) {
            float2 w = __ldg(&tw[p * s * k]);
            if (DIR > 0) w.y = -w.y;
            acc = cmul(acc, w);
        }
        Y[(q + s * (r * p + k)) * LS + line] = acc;
    }
}

// Full length-N transform of 2^Llog lines held in smem buffer A (scratch B).  All threads of
// the block must call it; data must be visible (a __syncthreads() before).  Returns the buffer
// that holds the (naturally ordered) result; the result is visible to all threads on return.
template <int DIR>
__device__ __forceinline__ float2* smem_fft(float2* A, float2* B, const float2* __restrict__ tw, int N, int Llog,
                                            int LS, const Stages& st) {
    int s = 1;
    for (int i = 0; i < st.n; ++i) {
        const int r = st.radix[i];
        switch (r) {
            case 2: stockham_stage<2, DIR>(A, B, tw, N, s, Llog, LS); break;
            case 3: stockham_stage<3, DIR>(A, B, tw, N, s, Llog, LS); break;
            case 4: stockham_stage<4, DIR>(A, B, tw, N, s, Llog, LS); break;
            case 5: stockham_stage<5, DIR>(A, B, tw, N, s, Llog, LS); break;
            case 7: stockham_stage<7, DIR>(A, B, tw, N, s, Llog, LS); break;
            case 8: stockham_stage<8, DIR>(A, B, tw, N, s, Llog, LS); break;
            case 9: stockham_stage<9, DIR>(A, B, tw, N, s, Llog, LS); break;
            case 16: stockham_stage<16, DIR>(A, B, tw, N, s, Llog, LS); break;
            case 25: stockham_stage<25, DIR>(A, B, tw, N, s, Llog, LS); break;
            default: stockham_stage_generic<DIR>(A, B, tw, N, s, r, Llog, LS); break;
        }
        __syncthreads();
        float2* t = A;
        A = B;
        B = t;
        s *= r;
    }
    return A;
}

}  // namespace pvd
