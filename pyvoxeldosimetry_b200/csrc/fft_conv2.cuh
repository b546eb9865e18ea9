// x pass (forward * cached spectrum * inverse) with the spectrum tile staged by TMA.
//
// The one-tile-per-CTA / tile-walk kernel in fft_fast.cuh reads its spectrum tile with 32 global loads per thread AFTER
// the forward transform: ncu's source page puts 41 % of that kernel's stall samples on the first uses of those values
// (profiles/r01g_ncu_p3_phases.txt) - a full DRAM round trip per tile that two CTAs per SM do not hide, and neither the
// register file (64 data registers per thread) nor shared memory (2 x 80 KB) has room for a second copy per CTA.
// Here ONE 512-thread CTA per SM runs two independent 256-thread groups (named barriers 1 and 2), each walking its own
// tiles exactly like a CTA of the old kernel, and the SM's spare 64 KB hold ONE spectrum tile that the groups use in
// turn: at the start of a tile the group's elected thread takes a ticket, waits until the previous user has released
// the buffer, and issues two bulk copies (TMA, mbarrier completion); the tile arrives during the group's input load and
// forward transform, is consumed from shared memory, and is released.  Waiting for the buffer staggers the two groups
// by half a tile, which is also what spreads their memory bursts.  All waits carry the 2 s watchdog.
#pragma once
#include "fft_pipe.cuh"

#ifndef PVD_EMULATE
namespace pvd {

struct Conv2Args {
    alignas(64) CUtensorMap tmap_spec;  // 3-D map of the spectrum buffer: dims (2*Sz floats, M1, M0), box (32, 1, BOXR)
    ColArgs c;
    int ntz, ntiles;
    int* error_flag;
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int N, int R1, int R2>
__global__ void __launch_bounds__(512, 1) cols_conv2_kernel(const __grid_constant__ Conv2Args pa) {
    constexpr int W = 16, GT = 256, TPC = GT / W;
    constexpr int RL = R2, STEP = N / RL, BPTL = (STEP + TPC - 1) / TPC;
    constexpr int NB1 = N / R1, STEPR = N / R1;
    constexpr int BOXR = tma_box_rows(N);
    using Fwd = Sched<N, R1, R2, 1>;
    using Rev = Sched<N, R2, R1, 1>;
    static_assert(R1 * R2 == N && STEP % TPC == 0 && NB1 % TPC == 0, "schedule must tile the 256-thread group exactly");
    const ColArgs& g = pa.c;
    PVD_DYN_SMEM(float2, smem);
    const int grp = threadIdx.x >> 8, tid = threadIdx.x & (GT - 1);
    float2* tile = smem + grp * (N * W);
    float2* S = smem + 2 * N * W;
    float2* tws = smem + 3 * N * W;
    float2* twr = tws + Fwd::TOTAL;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(twr + Rev::TOTAL);  // [0] spectrum landed, [1] buffer released
    int* sh = reinterpret_cast<int*>(bars + 2);                                          // [0] ticket counter, [1 + grp] group's ticket
    grid_dep_launch();
    Fwd::build(tws, g.tw);
    Rev::build(twr, g.tw);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        sh[0] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive(&bars[1]);  // the buffer starts released: phase 0 of `released` completes
    grid_dep_wait();
    GroupCtx<GT> ctx{tid, 1 + grp};
    const unsigned es = (unsigned)g.es, esb = es * (unsigned)sizeof(float2);
    const int wl = tid % W, b0 = tid / W;
    const int n_in = g.n_in;
    const unsigned cnt = (unsigned)g.out_n;
    const int blo = b0 - g.out_lo;
    const int ntz = pa.ntz;
    for (int t = 2 * (int)blockIdx.x + grp; t < pa.ntiles; t += 2 * (int)gridDim.x) {
        const int outer = t / ntz, zt = t - outer * ntz;
        const long long base = (long long)(g.outer0 + outer) * g.os + zt * W;
        const bool wok = wl < g.nzf - zt * W;
        if (tid == 0) {  // ticket, wait for the buffer, request this tile's spectrum
            const int u = atomicAdd(&sh[0], 1);
            sh[1 + grp] = u;
            mbar_wait_guarded(&bars[1], (unsigned)(u & 1), pa.error_flag, 5);
            mbar_expect_tx(&bars[0], (unsigned)(N * W * sizeof(float2)));
            PVD_UNROLL
            for (int i = 0; i < N / BOXR; ++i)
                tma_load_3d(S + i * (BOXR * W), &pa.tmap_spec, &bars[0], zt * (2 * W), g.outer0 + outer, i * BOXR);
        }
        const float2* src = opaque(g.in + base + (size_t)b0 * es + wl);
        float2* dst = opaque(g.out + base + (size_t)b0 * es + wl);
        auto gin = [&](int u, int j, int r, int) -> float2 {
            return (wok && r < n_in) ? ldg64(eptr(src, esb, u * TPC + NB1 * j)) : make_float2(0.f, 0.f);
        };
        float2 hold[BPTL][RL];
        auto rout = [&](int u, int k, int, int, float2 v) { hold[u][k] = v; };
        // the first stage issues its loads, then meets the group barrier (tables visible / exchange tile free)
        fast_fft<N, W, W, GT, -1, R1, R2, 1, true, false>(gin, rout, tile, tws, NoHook(), NoHook(), ctx);
        const int u = sh[1 + grp];  // written by tid 0 before the group barriers above
        mbar_wait_guarded(&bars[0], (unsigned)(u & 1), pa.error_flag, 6);
        PVD_UNROLL
        for (int uu = 0; uu < BPTL; ++uu) {
            PVD_UNROLL
            for (int k = 0; k < RL; ++k) hold[uu][k] = cmul(hold[uu][k], S[(b0 + uu * TPC + STEP * k) * W + wl]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // our generic reads of S before the next bulk copy into it
        ctx.sync();  // the whole group is done with S and with the exchange tile (last forward stage)
        if (tid == 0) mbar_arrive(&bars[1]);
        auto rin = [&](int uu, int j, int, int) -> float2 { return hold[uu][j]; };
        auto gout_rev = [&](int uu, int k, int, int, float2 v) {
            if (wok && (unsigned)(blo + uu * TPC + STEPR * k) < cnt) stg64(eptr(dst, esb, uu * TPC + STEPR * k), v);
        };
        fast_fft<N, W, W, GT, +1, R2, R1, 1, false, false>(rin, gout_rev, tile, twr, NoHook(), NoHook(), ctx);
    }
}

}  // namespace pvd
#endif
