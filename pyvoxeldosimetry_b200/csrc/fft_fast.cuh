// Size-specialised ("fast") versions of the five passes.
//
// Same mathematics and argument structs as fft_passes.cuh, but everything about the transform is a
// compile-time constant: length N, radix schedule (R1,R2[,R3]), tile width W = 16 lines (one 128-byte
// line of float2 per transform index) and the thread count NT.  Consequences:
//   * every radix stage runs in registers; the first stage of a column pass loads straight from HBM
//     (each half-warp reads one full 128-byte line), the last stage stores straight back - shared
//     memory is only the exchange buffer between stages (ONE buffer, N*16*8 bytes, in place);
//   * in the x pass the forward transform ends and the inverse transform starts in the same registers:
//     forward -> multiply by the cached kernel spectrum -> inverse without touching memory;
//   * index arithmetic folds to immediates: a butterfly b reads indices b + (N/R)*j; inter-stage
//     twiddles sit in shared memory as per-stage tables laid out [p][k] so one butterfly fetches its
//     R-1 factors with R/2 128-bit loads; the inverse multiplies by the conjugate instead of negating.
// Sizes outside the instantiated menu fall back to the generic engine in fft_passes.cuh.
#pragma once
#include "fft_passes.cuh"

namespace pvd {

__host__ __device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// Address helpers.  `opaque` hides a per-thread base pointer from the optimiser so that every access is
// formed as base + (32-bit stride in bytes) * (compile-time count) = ONE IMAD.WIDE.U32, instead of being
// re-derived from the kernel parameters with a 4-instruction 64-bit LEA sequence per access.
template <class T>
__device__ __forceinline__ T* opaque(T* p) {
#ifndef PVD_EMULATE
    asm volatile("" : "+l"(p));
#endif
    return p;
}
template <class T>
__device__ __forceinline__ T* eptr(T* base, unsigned stride_bytes, int count) {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(const_cast<typename std::remove_const<T>::type*>(base)) +
                                (unsigned long long)stride_bytes * (unsigned)count);
}

// The opaque base pointers are generic as far as the compiler knows; these keep the accesses on the
// global path (LDG/STG instead of generic LD/ST).
__device__ __forceinline__ float2 ldg64(const float2* p) {
#ifdef PVD_EMULATE
    return *p;
#else
    float2 v;
    asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ float2 ldg64_ro(const float2* p) {  // data that is read-only for the whole kernel (spectrum)
#ifdef PVD_EMULATE
    return *p;
#else
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef PVD_EMULATE
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
__device__ __forceinline__ void stg64(float2* p, float2 v) {
#ifdef PVD_EMULATE
    *p = v;
#else
    asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
#endif
}

// ---- per-stage twiddle tables ------------------------------------------------------------------
// Stage (radix R, stride S) of a length-N transform multiplies output k of butterfly p by W_N^{p*S*k},
// p in [0, M), M = N/(R*S).  Table row p holds k = 0..R-1, padded to an even count (16-byte rows).
template <int R>
constexpr int tw_row() { return R + (R & 1); }
template <int N, int R, int S>
constexpr int tw_size() { return (N / (R * S) > 1) ? (N / (R * S)) * tw_row<R>() : 0; }

template <int N, int R, int S>
__device__ __forceinline__ void build_stage_table(float2* dst, const float2* __restrict__ gtw) {
    constexpr int M = N / (R * S), RP = tw_row<R>();
    if constexpr (M > 1) {
        for (int i = threadIdx.x; i < M * RP; i += blockDim.x) {
            const int p = i / RP, k = i - p * RP;
            dst[i] = (k < R) ? gtw[p * S * k] : make_float2(0.f, 0.f);
        }
    }
}

// Tables of one radix schedule (RA, RB, RC); RC == 1 means two stages.  Only non-final stages have twiddles.
template <int N, int RA, int RB, int RC>
struct Sched {
    static constexpr int T1 = tw_size<N, RA, 1>();
    static constexpr int T2 = (RC > 1) ? tw_size<N, RB, RA>() : 0;
    static constexpr int TOTAL = T1 + T2;
    static __device__ __forceinline__ void build(float2* dst, const float2* __restrict__ gtw) {
        build_stage_table<N, RA, 1>(dst, gtw);
        if constexpr (RC > 1) build_stage_table<N, RB, RA>(dst + T1, gtw);
    }
};

// One Stockham DIF stage with radix R and stride S.  in(u, j, idx, w) -> float2, out(u, k, idx, w, v).
// (u, j)/(u, k) are the register slots (compile-time after unrolling), idx the transform index, w the line.
// Who runs a stage: the whole CTA (default) or a 256-thread group of it with its own named barrier.
struct BlockCtx {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
#ifndef PVD_EMULATE
template <int THREADS>
struct GroupCtx {
    int t, id;  // thread index inside the group, barrier id (1..15)
    __device__ __forceinline__ int tid() const { return t; }
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(THREADS) : "memory"); }
};
#endif
template <int N, int W, int NT, int R, int S, int DIR, bool SYNC_AFTER_READ, class In, class Out, class Ctx = BlockCtx>
__device__ __forceinline__ void fast_stage(In&& in, Out&& out, const float2* __restrict__ tws, Ctx ctx = Ctx()) {
    constexpr int NB = N / R;          // butterflies per line
    constexpr int M = NB / S;
    constexpr int TPC = NT / W;        // butterflies of one line processed concurrently
    constexpr int BPT = (NB + TPC - 1) / TPC;
    constexpr bool GUARD = (NB % TPC) != 0;
    constexpr int RP = tw_row<R>();
    static_assert(N % R == 0 && NB % S == 0 && NT % W == 0, "bad radix schedule");
    const int w = ctx.tid() % W;
    const int b0 = ctx.tid() / W;
    float2 a[BPT][R];
    PVD_UNROLL
    for (int u = 0; u < BPT; ++u) {
        const int b = b0 + u * TPC;
        if (!GUARD || b < NB) {
            PVD_UNROLL
            for (int j = 0; j < R; ++j) a[u][j] = in(u, j, b + NB * j, w);  // q + S*(p + M*j) == b + NB*j
        }
    }
    if (SYNC_AFTER_READ) ctx.sync();
    PVD_UNROLL
    for (int u = 0; u < BPT; ++u) {
        const int b = b0 + u * TPC;
        if (!GUARD || b < NB) {
            Dft<R, DIR>::run(a[u]);
            int obase;  // q + S*R*p
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
                    a[u][k] = (DIR < 0) ? cmul(a[u][k], t) : cmulc(a[u][k], t);
                }
            }
            PVD_UNROLL
            for (int k = 0; k < R; ++k) out(u, k, obase + S * k, w, a[u][k]);
        }
    }
}

// Whole transform.  Stage list (R1, R2, R3) with R3 == 1 meaning two stages.  IN_SMEM / OUT_SMEM say
// whether `in` / `out` address the exchange tile itself (then reads must complete before writes).
// `tws` = Sched<N,R1,R2,R3> tables.
template <int N, int W, int LS, int NT, int DIR, int R1, int R2, int R3, bool IN_SMEM, bool OUT_SMEM, class In, class Out, class Ctx = BlockCtx>
__device__ __forceinline__ void fast_fft(In&& in, Out&& out, float2* tile, const float2* __restrict__ tws, Ctx ctx = Ctx()) {
    static_assert(R1 * R2 * R3 == N, "radix schedule must multiply to N");
    auto sm_in = [&](int, int, int idx, int w) -> float2 { return tile[idx * LS + w]; };
    auto sm_out = [&](int, int, int idx, int w, float2 v) { tile[idx * LS + w] = v; };
    fast_stage<N, W, NT, R1, 1, DIR, IN_SMEM>(in, sm_out, tws, ctx);
    ctx.sync();
    if constexpr (R3 > 1) {
        fast_stage<N, W, NT, R2, R1, DIR, true>(sm_in, sm_out, tws + Sched<N, R1, R2, R3>::T1, ctx);
        ctx.sync();
        fast_stage<N, W, NT, R3, R1 * R2, DIR, OUT_SMEM>(sm_in, out, tws, ctx);
    } else {
        fast_stage<N, W, NT, R2, R1, DIR, OUT_SMEM>(sm_in, out, tws, ctx);
    }
}

// Radix of the last stage and the register-slot geometry of its outputs.
template <int N, int NT, int R1, int R2, int R3, int W = 16>
struct LastStage {
    static constexpr int RL = (R3 > 1) ? R3 : R2;
    static constexpr int STEP = N / RL;            // output index of slot (u, k) = b0 + u*TPC + STEP*k
    static constexpr int TPC = NT / W;
    static constexpr int BPT = (STEP + TPC - 1) / TPC;
};

// ------------------------------------------------------------------------------------------------
// Column passes (axis 1 forward / inverse, axis 0 forward * spectrum * inverse, kernel spectrum).
// One tile per CTA; used where the double-buffered persistent variant (fft_pipe.cuh) does not fit.
// W = frequencies per tile (16 = full 128-byte lines; 8 = half lines, half the shared memory per CTA: more CTAs per SM).
template <int N, int NT, int R1, int R2, int R3, int MODE, int MINB = 1, int W = 16>
__global__ void __launch_bounds__(NT, MINB) cols_fast_kernel(const ColArgs g) {
    using LS_ = LastStage<N, NT, R1, R2, R3, W>;
    constexpr int RL = LS_::RL, TPC = LS_::TPC, BPTL = LS_::BPT;
    using Fwd = Sched<N, R1, R2, R3>;
    using Rev = Sched<N, (R3 > 1 ? R3 : R2), (R3 > 1 ? R2 : R1), (R3 > 1 ? R1 : 1)>;
    constexpr bool SYM = (R3 > 1) ? (R1 == R3) : (R1 == R2);
    PVD_DYN_SMEM(float2, smem);
    float2* tile = smem;
    float2* tws = smem + N * W;
    float2* twr = SYM ? tws : tws + Fwd::TOTAL;
    // The one-tile-per-CTA form is launched normally (fully serialised behind its predecessor): as a programmatic
    // dependent every one of its thousands of CTAs paid the grid-dependency wait (C3: +70 us per volume, measured).
    // It still lets ITS successor start early.
    grid_dep_launch();
    Fwd::build(tws, g.tw);
    if constexpr (MODE == COL_CONV && !SYM) Rev::build(twr, g.tw);
    // Tile walk.  2-D grid: one tile per CTA (zt = blockIdx.x, outer = blockIdx.y).  loop_ntiles > 0: a 1-D grid of
    // resident CTAs strides over the linear tile list - the table build and the CTA start-up are paid once per CTA
    // instead of once per tile.  The first radix stage issues its global loads and THEN meets the block barrier
    // (IN_SMEM = true below): that barrier publishes the twiddle tables on the first tile and frees the exchange
    // tile on the following ones, and it does not wait for the loads themselves.
    const bool looped = g.loop_ntiles > 0;
    if (looped) grid_dep_wait();  // the persistent form is launched as a programmatic dependent: one wait per CTA
    const int tend = looped ? g.loop_ntiles : 1, tstep = looped ? (int)gridDim.x : 1;
    const unsigned es = (unsigned)g.es;
    const unsigned esb = es * (unsigned)sizeof(float2);  // stride between transform indices in bytes
    const int wl = threadIdx.x % W, b0 = threadIdx.x / W;
    const int n_in = g.n_in;
    const unsigned cnt = (unsigned)g.out_n;
    const int blo = b0 - g.out_lo;
    constexpr int NB1 = N / R1;
    for (int t = looped ? (int)blockIdx.x : 0; t < tend; t += tstep) {
    int zt = blockIdx.x, outer = blockIdx.y;
    if (looped) {
        outer = t / g.loop_ntz;
        zt = t - outer * g.loop_ntz;
    }
    const int z0 = zt * W;
    const long long base = (long long)(g.outer0 + outer) * g.os + z0;
    const int zlim = g.nzf - z0;
    const bool wok = wl < zlim;
    const float2* src = opaque(g.in + base + (size_t)b0 * es + wl);
    float2* dst = opaque(g.out + base + (size_t)b0 * es + wl);
    auto gin = [&](int u, int j, int r, int) -> float2 {
        return (wok && r < n_in) ? ldg64(eptr(src, esb, u * TPC + NB1 * j)) : make_float2(0.f, 0.f);
    };
    auto gout = [&](int u, int k, int, int, float2 v) {
        if (wok && (unsigned)(blo + u * TPC + LS_::STEP * k) < cnt) stg64(eptr(dst, esb, u * TPC + LS_::STEP * k), v);
    };
    if constexpr (MODE == COL_FWD) {
        fast_fft<N, W, W, NT, -1, R1, R2, R3, true, false>(gin, gout, tile, tws);
    } else if constexpr (MODE == COL_SPEC) {
        const float sc = g.scale;
        auto sout = [&](int u, int k, int r, int w, float2 v) { gout(u, k, r, w, make_float2(v.x * sc, v.y * sc)); };
        fast_fft<N, W, W, NT, -1, R1, R2, R3, true, false>(gin, sout, tile, tws);
    } else if constexpr (MODE == COL_INV) {
        fast_fft<N, W, W, NT, +1, R1, R2, R3, true, false>(gin, gout, tile, tws);
    } else {  // COL_CONV: forward -> * spectrum -> inverse, the middle never leaves registers
        float2 hold[BPTL][RL];
        auto rout = [&](int u, int k, int, int, float2 v) { hold[u][k] = v; };
        const float2* sp = opaque(g.spec + base + (size_t)b0 * es + wl);
        fast_fft<N, W, W, NT, -1, R1, R2, R3, true, false>(gin, rout, tile, tws);
        PVD_UNROLL
        for (int u = 0; u < BPTL; ++u) {
            if (b0 + u * TPC < LS_::STEP) {
                PVD_UNROLL
                for (int k = 0; k < RL; ++k)
                    if (wok) hold[u][k] = cmul(hold[u][k], ldg64_ro(eptr(sp, esb, u * TPC + LS_::STEP * k)));
            }
        }
        __syncthreads();  // every thread finished reading the tile in the last forward stage
        auto rin = [&](int u, int j, int, int) -> float2 { return hold[u][j]; };
        constexpr int STEPR = N / R1;  // the reversed schedule ends with radix R1
        auto gout_rev = [&](int u, int k, int, int, float2 v) {
            if (wok && (unsigned)(blo + u * TPC + STEPR * k) < cnt) stg64(eptr(dst, esb, u * TPC + STEPR * k), v);
        };
        if constexpr (R3 > 1)
            fast_fft<N, W, W, NT, +1, R3, R2, R1, false, false>(rin, gout_rev, tile, twr);
        else
            fast_fft<N, W, W, NT, +1, R2, R1, 1, false, false>(rin, gout_rev, tile, twr);
    }
    }  // tile walk
}

// ------------------------------------------------------------------------------------------------
// Row passes: 32 real rows = 16 packed complex lines per block; tile stride 17 keeps the transposing
// loads/stores at 2-way bank conflicts and the butterflies conflict-free.
template <int N, int NT, int R1, int R2, int R3>
__global__ void __launch_bounds__(NT) rows_fwd_fast_kernel(const RowFwdArgs g) {
    constexpr int W = 16, LS = 17;
    PVD_DYN_SMEM(float2, smem);
    float2* tile = smem;
    float2* tws = smem + N * LS;
    Sched<N, R1, R2, R3>::build(tws, g.tw);
    const long long nrows = (long long)g.n0 * g.n1;
    const long long row0 = (long long)blockIdx.x * (2 * W);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NWARPS = NT / 32;
    float* Af = reinterpret_cast<float*>(tile);
    const int T = g.T, n2 = g.n2;
    for (int rr = warp; rr < 2 * W; rr += NWARPS) {
        const long long R = row0 + rr;
        const bool valid = R < nrows;
        long long off = 0;
        if (valid) {
            const long long x = R / g.n1, y = R - x * g.n1;
            off = x * g.in_s0 + y * g.in_s1;
        }
        float* dstf = Af + (rr >> 1) * 2 + (rr & 1);
        if (T == 1) {
            const float* __restrict__ p0 = g.in[0] + off;
            const float w0 = g.w[0];
            PVD_UNROLL
            for (int z = lane; z < N; z += 32) dstf[z * (2 * LS)] = (valid && z < n2) ? w0 * __ldg(p0 + z) : 0.f;
        } else {
            // time-weighted sum: one batch of independent loads per time point (memory-level parallelism)
            constexpr int ZIT = (N + 31) / 32;
            float acc[ZIT];
            PVD_UNROLL
            for (int i = 0; i < ZIT; ++i) acc[i] = 0.f;
            // TB time points per batch (TB * ZIT independent loads in flight per lane, bounded by the register budget)
            constexpr int TB = ZIT <= 8 ? 4 : (ZIT <= 16 ? 2 : 1);
            int t = 0;
            for (; TB > 1 && t + TB <= T; t += TB) {
                float v[TB][ZIT];
                PVD_UNROLL
                for (int q = 0; q < TB; ++q) {
                    const float* __restrict__ pt = g.in[t + q] + off;
                    PVD_UNROLL
                    for (int i = 0; i < ZIT; ++i) {
                        const int z = lane + 32 * i;
                        v[q][i] = (valid && z < n2) ? __ldg(pt + z) : 0.f;
                    }
                }
                PVD_UNROLL
                for (int q = 0; q < TB; ++q) {
                    const float wt = g.w[t + q];
                    PVD_UNROLL
                    for (int i = 0; i < ZIT; ++i) acc[i] = fmaf(wt, v[q][i], acc[i]);
                }
            }
            for (; t < T; ++t) {
                const float* __restrict__ pt = g.in[t] + off;
                const float wt = g.w[t];
                PVD_UNROLL
                for (int i = 0; i < ZIT; ++i) {
                    const int z = lane + 32 * i;
                    if (valid && z < n2) acc[i] = fmaf(wt, __ldg(pt + z), acc[i]);
                }
            }
            PVD_UNROLL
            for (int i = 0; i < ZIT; ++i) {
                const int z = lane + 32 * i;
                if (z < N) dstf[z * (2 * LS)] = acc[i];
            }
        }
    }
    __syncthreads();
    auto sm_in = [&](int, int, int idx, int w) -> float2 { return tile[idx * LS + w]; };
    auto sm_out = [&](int, int, int idx, int w, float2 v) { tile[idx * LS + w] = v; };
    fast_fft<N, W, LS, NT, -1, R1, R2, R3, true, true>(sm_in, sm_out, tile, tws);
    __syncthreads();
    const int Nh = g.Nh;
    for (int rr = warp; rr < 2 * W; rr += NWARPS) {
        const long long R = row0 + rr;
        if (R >= nrows) continue;
        const long long x = R / g.n1, y = R - x * g.n1;
        float2* __restrict__ dst = g.out + x * g.out_s0 + y * g.out_s1;
        const int line = rr >> 1;
        const bool odd = rr & 1;
        for (int k = lane; k < Nh; k += 32) {
            const float2 zk = tile[k * LS + line];
            const float2 zm = tile[((k == 0) ? 0 : N - k) * LS + line];
            dst[k] = odd ? make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x))
                         : make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        }
    }
}

template <int N, int NT, int R1, int R2, int R3>
__global__ void __launch_bounds__(NT) rows_inv_fast_kernel(const RowInvArgs g) {
    constexpr int W = 16, LS = 17;
    PVD_DYN_SMEM(float2, smem);
    float2* tile = smem;
    float2* tws = smem + N * LS;
    Sched<N, R1, R2, R3>::build(tws, g.tw);
    const long long nrows = (long long)g.O0 * g.O1;
    const long long row0 = (long long)blockIdx.x * (2 * W);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NWARPS = NT / 32;
    const int Nh = g.Nh;
    for (int line = warp; line < W; line += NWARPS) {
        const long long Ra = row0 + 2 * line, Rb = Ra + 1;
        const bool va = Ra < nrows, vb = Rb < nrows;
        const float2 *pa = g.in, *pb = g.in;
        if (va) {
            const long long xa = Ra / g.O1, ya = Ra - xa * g.O1;
            pa = g.in + (xa + g.x_lo) * g.in_s0 + (ya + g.y_lo) * g.in_s1;
        }
        if (vb) {
            const long long xb = Rb / g.O1, yb = Rb - xb * g.O1;
            pb = g.in + (xb + g.x_lo) * g.in_s0 + (yb + g.y_lo) * g.in_s1;
        }
        // all global loads of the line first (memory-level parallelism), then the packing
        constexpr int KIT = (N / 2 + 1 + 31) / 32;
        float2 av[KIT], bv[KIT];
        PVD_UNROLL
        for (int i = 0; i < KIT; ++i) {
            const int k = lane + 32 * i;
            av[i] = (va && k < Nh) ? pa[k] : make_float2(0.f, 0.f);
            bv[i] = (vb && k < Nh) ? pb[k] : make_float2(0.f, 0.f);
        }
        PVD_UNROLL
        for (int i = 0; i < KIT; ++i) {
            const int k = lane + 32 * i;
            if (k < Nh) {
                float2 a = av[i], b = bv[i];
                const int mk = N - k;
                const bool self = (k == 0) || (mk == k);
                if (self) {
                    a.y = 0.f;
                    b.y = 0.f;
                }
                tile[k * LS + line] = make_float2(a.x - b.y, a.y + b.x);
                if (!self) tile[mk * LS + line] = make_float2(a.x + b.y, b.x - a.y);
            }
        }
    }
    __syncthreads();
    auto sm_in = [&](int, int, int idx, int w) -> float2 { return tile[idx * LS + w]; };
    auto sm_out = [&](int, int, int idx, int w, float2 v) { tile[idx * LS + w] = v; };
    fast_fft<N, W, LS, NT, +1, R1, R2, R3, true, true>(sm_in, sm_out, tile, tws);
    __syncthreads();
    const float* Zf = reinterpret_cast<const float*>(tile);
    const float scale = g.scale, rho_ref = g.rho_ref, rho_min = g.rho_min, rho_cut = g.rho_cut;
    const int O2 = g.O2, z_lo = g.z_lo;
    for (int rr = warp; rr < 2 * W; rr += NWARPS) {
        const long long R = row0 + rr;
        if (R >= nrows) continue;
        const long long x = R / g.O1, y = R - x * g.O1;
        float* __restrict__ dst = g.out + x * g.out_s0 + y * g.out_s1;
        const float* srcf = Zf + (rr >> 1) * 2 + (rr & 1) + (size_t)z_lo * (2 * LS);
        constexpr int ZIT = (N + 31) / 32;
        if (g.density) {
            const float* __restrict__ den = g.density + x * g.den_s0 + y * g.den_s1;
            float rho[ZIT];
            PVD_UNROLL
            for (int i = 0; i < ZIT; ++i) {
                const int z = lane + 32 * i;
                rho[i] = (z < O2) ? __ldg(den + z) : 1.f;
            }
            PVD_UNROLL
            for (int i = 0; i < ZIT; ++i) {
                const int z = lane + 32 * i;
                if (z < O2) {
                    const float v = srcf[z * (2 * LS)] * scale;
                    dst[z] = (rho[i] < rho_cut) ? 0.f : v * __fdividef(rho_ref, fmaxf(rho[i], rho_min));
                }
            }
        } else {
            PVD_UNROLL
            for (int i = 0; i < ZIT; ++i) {
                const int z = lane + 32 * i;
                if (z < O2) dst[z] = srcf[z * (2 * LS)] * scale;
            }
        }
    }
}

}  // namespace pvd
