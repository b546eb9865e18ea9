// Persistent pipelined column passes, two adjacent columns per thread.
//
// ncu on the one-column-per-thread kernels (profiles/r01_ncu_pipe.md) shows them SM-side bound, not HBM
// bound: issue slots ~46 % active, top stall mio_throttle (the shared-memory / LSU instruction queue), and
// the x pass takes exactly twice the y pass with twice the instructions.  Here every thread owns the SAME
// butterfly of two neighbouring frequency columns, so every shared-memory, global and twiddle access is a
// 128-bit instruction serving both columns and all address / predicate arithmetic is shared: half the
// memory instructions, ~25 % fewer instructions overall, identical arithmetic.
#pragma once
#include "fft_pipe.cuh"

namespace pvd {

__device__ __forceinline__ float4 ldg128_ro(const float4* p) {
#ifdef PVD_EMULATE
    return *p;
#else
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ void stg128(float4* p, float4 v) {
#ifdef PVD_EMULATE
    *p = v;
#else
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#endif
}

// Stage on float4 slots: (x, y) = column 2w, (z, w) = column 2w+1.  in(u, j, idx, w2) -> float4.
template <int N, int NT, int R, int S, int DIR, bool SYNC_AFTER_READ, class In, class Out>
__device__ __forceinline__ void fast_stage2(In&& in, Out&& out, const float2* __restrict__ tws) {
    constexpr int WT = 8;              // thread columns (pairs of frequency columns)
    constexpr int NB = N / R;
    constexpr int M = NB / S;
    constexpr int TPC = NT / WT;
    constexpr int BPT = (NB + TPC - 1) / TPC;
    constexpr bool GUARD = (NB % TPC) != 0;
    constexpr int RP = tw_row<R>();
    static_assert(N % R == 0 && NB % S == 0 && NT % WT == 0, "bad radix schedule");
    const int w2 = threadIdx.x % WT;
    const int b0 = threadIdx.x / WT;
    float2 a0[BPT][R], a1[BPT][R];
    PVD_UNROLL
    for (int u = 0; u < BPT; ++u) {
        const int b = b0 + u * TPC;
        if (!GUARD || b < NB) {
            PVD_UNROLL
            for (int j = 0; j < R; ++j) {
                const float4 v = in(u, j, b + NB * j, w2);
                a0[u][j] = make_float2(v.x, v.y);
                a1[u][j] = make_float2(v.z, v.w);
            }
        }
    }
    if (SYNC_AFTER_READ) __syncthreads();
    PVD_UNROLL
    for (int u = 0; u < BPT; ++u) {
        const int b = b0 + u * TPC;
        if (!GUARD || b < NB) {
            Dft<R, DIR>::run(a0[u]);
            Dft<R, DIR>::run(a1[u]);
            int obase;
            if constexpr (S == 1) {
                obase = R * b;
            } else if constexpr (M == 1) {
                obase = b;
            } else {
                obase = (b % S) + (R * S) * (b / S);
            }
            if constexpr (M > 1) {
                const float4* __restrict__ tp = reinterpret_cast<const float4*>(tws + (b / S) * RP);
                float4 tv[RP / 2];
                PVD_UNROLL
                for (int i = 0; i < RP / 2; ++i) tv[i] = tp[i];
                PVD_UNROLL
                for (int k = 1; k < R; ++k) {
                    const float2 t = (k & 1) ? make_float2(tv[k / 2].z, tv[k / 2].w) : make_float2(tv[k / 2].x, tv[k / 2].y);
                    a0[u][k] = (DIR < 0) ? cmul(a0[u][k], t) : cmulc(a0[u][k], t);
                    a1[u][k] = (DIR < 0) ? cmul(a1[u][k], t) : cmulc(a1[u][k], t);
                }
            }
            PVD_UNROLL
            for (int k = 0; k < R; ++k)
                out(u, k, obase + S * k, w2, make_float4(a0[u][k].x, a0[u][k].y, a1[u][k].x, a1[u][k].y));
        }
    }
}

template <int N, int NT, int DIR, int R1, int R2, int R3, bool IN_SMEM, bool OUT_SMEM, class In, class Out>
__device__ __forceinline__ void fast_fft2(In&& in, Out&& out, float2* tile, const float2* __restrict__ tws) {
    static_assert(R1 * R2 * R3 == N, "radix schedule must multiply to N");
    float4* t4 = reinterpret_cast<float4*>(tile);  // row = 16 float2 = 8 float4
    auto sm_in = [&](int, int, int idx, int w2) -> float4 { return t4[idx * 8 + w2]; };
    auto sm_out = [&](int, int, int idx, int w2, float4 v) { t4[idx * 8 + w2] = v; };
    fast_stage2<N, NT, R1, 1, DIR, IN_SMEM>(in, sm_out, tws);
    __syncthreads();
    if constexpr (R3 > 1) {
        fast_stage2<N, NT, R2, R1, DIR, true>(sm_in, sm_out, tws + Sched<N, R1, R2, R3>::T1);
        __syncthreads();
        fast_stage2<N, NT, R3, R1 * R2, DIR, OUT_SMEM>(sm_in, out, tws);
    } else {
        fast_stage2<N, NT, R2, R1, DIR, OUT_SMEM>(sm_in, out, tws);
    }
}

template <int N, int NT, int MINB, int R1, int R2, int R3, int MODE>
__global__ void __launch_bounds__(NT, MINB) cols_pipe2_kernel(const ColPipeArgs pa) {
    constexpr int W = 16, WT = 8;
    constexpr int RL = (R3 > 1) ? R3 : R2;
    constexpr int STEP = N / RL;
    constexpr int TPC = NT / WT;
    constexpr int BPTL = (STEP + TPC - 1) / TPC;
    using Fwd = Sched<N, R1, R2, R3>;
    using Rev = Sched<N, (R3 > 1 ? R3 : R2), (R3 > 1 ? R2 : R1), (R3 > 1 ? R1 : 1)>;
    constexpr bool SYM = (R3 > 1) ? (R1 == R3) : (R1 == R2);
    constexpr int CHUNKS = N * 8;
    static_assert(CHUNKS % NT == 0 && NT % 8 == 0, "tile must split evenly over the threads");
    const ColArgs& g = pa.c;
    PVD_DYN_SMEM(float2, smem);
    float2* tws = smem + 2 * N * W;
    float2* twr = SYM ? tws : tws + Fwd::TOTAL;
    Fwd::build(tws, g.tw);
    if constexpr (MODE == COL_CONV && !SYM) Rev::build(twr, g.tw);
    const unsigned es = (unsigned)g.es;
    const unsigned esb = es * (unsigned)sizeof(float2);
    const int n_in = g.n_in;
    const unsigned cnt = (unsigned)g.out_n;
    const int ntz = pa.ntz, ntiles = pa.ntiles;
    const unsigned magic = pa.ntz_magic;
    const int crow = threadIdx.x >> 3, ccol = (threadIdx.x & 7) * 2;  // cp.async chunk and FFT column pair coincide
    const int b0 = crow;
    const int blo = b0 - g.out_lo;
    const size_t toff = (size_t)b0 * es + ccol;
    auto tile_base = [&](int t, int& zt) -> long long {
        const int outer = (ntz == 1) ? t : (int)__umulhi((unsigned)t, magic);
        zt = t - outer * ntz;
        return (long long)(g.outer0 + outer) * g.os + (long long)zt * W;
    };
    auto issue = [&](float2* buf, int t) {
        int zt;
        const float2* src = opaque(g.in + tile_base(t, zt) + toff);
        float2* dstp = buf + crow * W + ccol;
        PVD_UNROLL
        for (int i = 0; i < CHUNKS / NT; ++i)
            cp_async16(dstp + i * ((NT / 8) * W), eptr(src, esb, i * (NT / 8)), crow + i * (NT / 8) < n_in);
    };
    int t = blockIdx.x;
    if (t < ntiles) issue(smem, t);
    cp_async_commit();
    int cur = 0;
    for (; t < ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        if (tn < ntiles) issue(smem + (cur ^ 1) * (N * W), tn);
        cp_async_commit();
        int zt;
        const long long base = tile_base(t, zt);
        const bool wok = ccol < g.nzf - zt * W;  // the odd column of the last pair may be padding (harmless)
        float4* dst = reinterpret_cast<float4*>(opaque(g.out + base + toff));
        float4 sp[BPTL][RL];
        if constexpr (MODE == COL_CONV) {
            const float4* spp = reinterpret_cast<const float4*>(opaque(g.spec + base + toff));
            PVD_UNROLL
            for (int u = 0; u < BPTL; ++u) {
                PVD_UNROLL
                for (int k = 0; k < RL; ++k)
                    sp[u][k] = (wok && b0 + u * TPC < STEP) ? ldg128_ro(eptr(spp, esb, u * TPC + STEP * k))
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        cp_async_wait<1>();
        __syncthreads();
        float2* tile = smem + cur * (N * W);
        float4* t4 = reinterpret_cast<float4*>(tile);
        auto sm_in = [&](int, int, int idx, int w2) -> float4 { return t4[idx * 8 + w2]; };
        auto gout = [&](int u, int k, int, int, float4 v) {
            if (wok && (unsigned)(blo + u * TPC + STEP * k) < cnt) stg128(eptr(dst, esb, u * TPC + STEP * k), v);
        };
        if constexpr (MODE == COL_FWD) {
            fast_fft2<N, NT, -1, R1, R2, R3, true, false>(sm_in, gout, tile, tws);
        } else if constexpr (MODE == COL_SPEC) {
            const float sc = g.scale;
            auto sout = [&](int u, int k, int r, int w2, float4 v) {
                gout(u, k, r, w2, make_float4(v.x * sc, v.y * sc, v.z * sc, v.w * sc));
            };
            fast_fft2<N, NT, -1, R1, R2, R3, true, false>(sm_in, sout, tile, tws);
        } else if constexpr (MODE == COL_INV) {
            fast_fft2<N, NT, +1, R1, R2, R3, true, false>(sm_in, gout, tile, tws);
        } else {
            float4 hold[BPTL][RL];
            auto rout = [&](int u, int k, int, int, float4 v) { hold[u][k] = v; };
            fast_fft2<N, NT, -1, R1, R2, R3, true, false>(sm_in, rout, tile, tws);
            PVD_UNROLL
            for (int u = 0; u < BPTL; ++u) {
                PVD_UNROLL
                for (int k = 0; k < RL; ++k) {
                    const float4 h = hold[u][k], s = sp[u][k];
                    hold[u][k] = make_float4(h.x * s.x - h.y * s.y, h.x * s.y + h.y * s.x, h.z * s.z - h.w * s.w,
                                             h.z * s.w + h.w * s.z);
                }
            }
            __syncthreads();
            auto rin = [&](int u, int j, int, int) -> float4 { return hold[u][j]; };
            constexpr int STEPR = N / R1;
            auto gout_rev = [&](int u, int k, int, int, float4 v) {
                if (wok && (unsigned)(blo + u * TPC + STEPR * k) < cnt) stg128(eptr(dst, esb, u * TPC + STEPR * k), v);
            };
            if constexpr (R3 > 1)
                fast_fft2<N, NT, +1, R3, R2, R1, false, false>(rin, gout_rev, tile, twr);
            else
                fast_fft2<N, NT, +1, R2, R1, 1, false, false>(rin, gout_rev, tile, twr);
        }
        __syncthreads();
        cur ^= 1;
    }
    cp_async_wait<0>();
}

}  // namespace pvd
