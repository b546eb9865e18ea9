// Persistent, software-pipelined column passes.
//
// One CTA per SM walks a strided list of column tiles (N transform indices x 16 frequencies = N full
// 128-byte lines).  While the FFT of tile i runs out of shared-memory buffer i&1, the HBM loads of tile
// i+1 are already in flight into the other buffer (cp.async.cg 16-byte copies, no registers held, zero
// fill for the padded rows); results leave straight from registers (fire-and-forget stores).  In the
// x pass the matching tile of the cached kernel spectrum is prefetched into registers at the start of the
// tile and consumed after the forward transform.  The point is memory-level parallelism: ncu showed the
// non-pipelined kernels stalled on long_scoreboard with ~50 % issue activity (profiles/r01_ncu_fast.md).
#pragma once
#include "direct_conv.cuh"  // CUtensorMap, mbarrier / TMA helpers
#include "fft_fast.cuh"

namespace pvd {

#ifdef PVD_EMULATE
static inline void cp_async16(void* dst, const void* src, bool valid) {
    if (valid) std::memcpy(dst, src, 16);
    else std::memset(dst, 0, 16);
}
static inline void cp_async16_full(void* dst, const void* src) { std::memcpy(dst, src, 16); }
static inline void st_zero16(void* dst) { std::memset(dst, 0, 16); }
static inline void cp_async_commit() {}
template <int K> static inline void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros, nothing read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
// Unconditional 16-byte copy: ONE LDGSTS.  The src-size form above costs ~8 more instructions per copy (ptxas
// lowers the run-time size with address arithmetic and three dummy shared loads), so callers use it only where a
// chunk is really partial and write plain zeros (st_zero16) for padding.  (Measured: the row passes gain ~1 %; the
// HBM-bound column passes got 5 % SLOWER with it - 0.152 -> 0.160 ms - and keep the src-size form.)
__device__ __forceinline__ void cp_async16_full(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void st_zero16(void* smem_dst) { *reinterpret_cast<float4*>(smem_dst) = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int K> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(K) : "memory"); }
#endif

struct ColPipeArgs {
    // TMA variant (use_tma): the tile [N indices][W frequencies] is a 3-D box of this tensor map over the work buffer
    // - dims (2*Sz floats, rows, outer), box (2*W floats, BOXR rows, 1) - so ONE thread issues N/BOXR bulk copies per
    // tile instead of N*8/NT 16-byte cp.async per thread (ncu: mio_throttle was the top stall of the y passes).
    alignas(64) CUtensorMap tmap;
    ColArgs c;
    int ntz;     // tiles along the frequency axis
    int ntiles;  // ntz * number of outer indices
    unsigned ntz_magic;  // ceil(2^32 / ntz): t / ntz == __umulhi(t, magic) for t * ntz < 2^32
    int use_tma;
    int* error_flag;  // set if a bulk copy never completes (bad descriptor) instead of hanging the GPU
};

// rows per TMA box: the largest divisor of N that fits the 256-element box limit
constexpr int tma_box_rows(int n) {
    int best = 1;
    for (int d = 1; d <= 256 && d <= n; ++d)
        if (n % d == 0) best = d;
    return best;
}

// W = frequencies per tile: 16 (full 128-byte lines) wherever two N x 16 buffers fit one SM; 8 (64-byte half lines) for
// the 1024- and 1152-point y passes, whose full-line tile (131 / 147 KB) could only run unpipelined, one tile per CTA.
template <int N, int NT, int MINB, int R1, int R2, int R3, int MODE, int W = 16>
__global__ void __launch_bounds__(NT, MINB) cols_pipe_kernel(const __grid_constant__ ColPipeArgs pa) {
    static_assert(W == 16 || W == 8, "tile width: full or half 128-byte lines");
    constexpr int CPR = W / 2;  // 16-byte chunks per tile row
    using LS_ = LastStage<N, NT, R1, R2, R3, W>;
    constexpr int RL = LS_::RL, TPC = LS_::TPC, BPTL = LS_::BPT;
    using Fwd = Sched<N, R1, R2, R3>;
    using Rev = Sched<N, (R3 > 1 ? R3 : R2), (R3 > 1 ? R2 : R1), (R3 > 1 ? R1 : 1)>;
    constexpr bool SYM = (R3 > 1) ? (R1 == R3) : (R1 == R2);
    constexpr int CHUNKS = N * CPR;  // 16-byte chunks per tile
    static_assert(CHUNKS % NT == 0 && NT % CPR == 0, "tile must split evenly over the threads");
    const ColArgs& g = pa.c;
    PVD_DYN_SMEM(float2, smem);
    float2* tws = smem + 2 * N * W;
    float2* twr = SYM ? tws : tws + Fwd::TOTAL;
    grid_dep_launch();
    Fwd::build(tws, g.tw);
    if constexpr (MODE == COL_CONV && !SYM) Rev::build(twr, g.tw);
    grid_dep_wait();
    const unsigned es = (unsigned)g.es;
    const unsigned esb = es * (unsigned)sizeof(float2);
    const int n_in = g.n_in;
    const unsigned cnt = (unsigned)g.out_n;
    const int ntz = pa.ntz, ntiles = pa.ntiles;
    const int crow = threadIdx.x / CPR, ccol = (threadIdx.x % CPR) * 2;  // this thread's chunk inside a row group
    const int wl = threadIdx.x % W, b0 = threadIdx.x / W;
    const int blo = b0 - g.out_lo;
    const size_t toff = (size_t)b0 * es + wl;  // this thread's element inside a tile (stage-1 input / last-stage output)
    const unsigned magic = pa.ntz_magic;
    auto tile_base = [&](int t, int& zt) -> long long {
        const int outer = (ntz == 1) ? t : (int)__umulhi((unsigned)t, magic);
        zt = t - outer * ntz;
        return (long long)(g.outer0 + outer) * g.os + (long long)zt * W;
    };
    const size_t coff = (size_t)crow * es + ccol;
    auto issue = [&](float2* buf, int t) {
        int zt;
        const float2* src = opaque(g.in + tile_base(t, zt) + coff);
        float2* dstp = buf + crow * W + ccol;
        PVD_UNROLL
        for (int i = 0; i < CHUNKS / NT; ++i)
            cp_async16(dstp + i * ((NT / CPR) * W), eptr(src, esb, i * (NT / CPR)), crow + i * (NT / CPR) < n_in);
    };
#ifndef PVD_EMULATE
    constexpr int BOXR = tma_box_rows(N);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + 2 * N * W + 4 * N) - 2;  // tail of the table area
    const bool tma = pa.use_tma && ((unsigned)__cvta_generic_to_shared(smem) & 127u) == 0;
    unsigned ph0 = 0, ph1 = 0;  // phase parity of the two buffers' mbarriers
    auto issue_tma = [&](int b, int t) {  // one thread: N/BOXR bulk copies of BOXR rows x 128 bytes
        int zt;
        const int outer = (ntz == 1) ? t : (int)__umulhi((unsigned)t, magic);
        zt = t - outer * ntz;
        mbar_expect_tx(&bars[b], (unsigned)(N * W * sizeof(float2)));
        PVD_UNROLL
        for (int i = 0; i < N / BOXR; ++i)
            tma_load_3d(smem + b * (N * W) + i * (BOXR * W), &pa.tmap, &bars[b], zt * (2 * W), i * BOXR, g.outer0 + outer);
    };
    if (tma) {
        if (threadIdx.x == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
        }
        __syncthreads();
    }
#else
    const bool tma = false;
#endif
    int t = blockIdx.x;
    if (!tma) {
        if (t < ntiles) issue(smem, t);
        cp_async_commit();
    }
#ifndef PVD_EMULATE
    else if (t < ntiles && threadIdx.x == 0) issue_tma(0, t);
#endif
    int cur = 0;
    for (; t < ntiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        if (!tma) {
            if (tn < ntiles) issue(smem + (cur ^ 1) * (N * W), tn);
            cp_async_commit();
        }
#ifndef PVD_EMULATE
        else if (tn < ntiles && threadIdx.x == 0) issue_tma(cur ^ 1, tn);
#endif
        int zt;
        const long long base = tile_base(t, zt);
        const bool wok = wl < g.nzf - zt * W;
        float2* dst = opaque(g.out + base + toff);
        float2 sp[BPTL][RL];
        if constexpr (MODE == COL_CONV) {  // spectrum tile -> registers, consumed after the forward transform
            const float2* spp = opaque(g.spec + base + toff);
            PVD_UNROLL
            for (int u = 0; u < BPTL; ++u) {
                PVD_UNROLL
                for (int k = 0; k < RL; ++k)
                    sp[u][k] = (wok && b0 + u * TPC < LS_::STEP) ? ldg64_ro(eptr(spp, esb, u * TPC + LS_::STEP * k))
                                                                  : make_float2(0.f, 0.f);
            }
        }
        if (!tma) {
            cp_async_wait<1>();  // everything but the prefetch just issued has landed
            __syncthreads();
        }
#ifndef PVD_EMULATE
        else {
            const unsigned ph = cur ? ph1 : ph0;
            const long long t0 = clock64();
            while (!mbar_try_wait(&bars[cur], ph)) {
                if (clock64() - t0 > kWatchdogCycles) {  // ~10 s: a broken descriptor must not hang the GPU
                    if (threadIdx.x == 0) *pa.error_flag = 2;
                    break;
                }
            }
            if (cur) ph1 ^= 1; else ph0 ^= 1;
        }
#endif
        float2* tile = smem + cur * (N * W);
        auto sm_in = [&](int, int, int idx, int w) -> float2 { return tile[idx * W + w]; };
        auto gout = [&](int u, int k, int, int, float2 v) {
            if (wok && (unsigned)(blo + u * TPC + LS_::STEP * k) < cnt) stg64(eptr(dst, esb, u * TPC + LS_::STEP * k), v);
        };
        if constexpr (MODE == COL_FWD) {
            fast_fft<N, W, W, NT, -1, R1, R2, R3, true, false>(sm_in, gout, tile, tws);
        } else if constexpr (MODE == COL_SPEC) {
            const float sc = g.scale;
            auto sout = [&](int u, int k, int r, int w, float2 v) { gout(u, k, r, w, make_float2(v.x * sc, v.y * sc)); };
            fast_fft<N, W, W, NT, -1, R1, R2, R3, true, false>(sm_in, sout, tile, tws);
        } else if constexpr (MODE == COL_INV) {
            fast_fft<N, W, W, NT, +1, R1, R2, R3, true, false>(sm_in, gout, tile, tws);
        } else {
            float2 hold[BPTL][RL];
            auto rout = [&](int u, int k, int, int, float2 v) { hold[u][k] = v; };
            fast_fft<N, W, W, NT, -1, R1, R2, R3, true, false>(sm_in, rout, tile, tws);
            PVD_UNROLL
            for (int u = 0; u < BPTL; ++u) {
                PVD_UNROLL
                for (int k = 0; k < RL; ++k) hold[u][k] = cmul(hold[u][k], sp[u][k]);
            }
            __syncthreads();  // all reads of the tile by the last forward stage are done
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
#ifndef PVD_EMULATE
        // the exchange stages wrote this buffer through the generic proxy; the bulk copy that refills it writes through
        // the async proxy: order the two before the barrier that hands the buffer back
        if (tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        __syncthreads();  // tile buffer may be refilled by the next iteration's prefetch
        cur ^= 1;
    }
    cp_async_wait<0>();
}

}  // namespace pvd
