// Direct tiled 3-D convolution for small dose kernels ('same' / zero boundary), TMA-staged halo tiles.
//
// One CTA computes an 8 x 8 x 64 (x, y, z) block of outputs.  Its input neighbourhood
// (8+K0-1) x (8+K1-1) x (64+K2-1) is ONE TMA box load (cp.async.bulk.tensor.3d into shared memory,
// completion on an mbarrier): the tensor map describes the activity volume, the box origin is the block
// origin minus the kernel reach, and everything outside the volume arrives as zeros - exactly the zero
// boundary of 'same' mode, so no halo logic exists in the kernel.  256 threads; each owns a 4 (x) x 4 (z)
// register tile, reads its input runs with 128-bit shared-memory loads (conflict-free: 16 lanes cover a
// contiguous 256-byte run) and the taps as warp-uniform broadcasts.  K0*K1*K2 FMAs per voxel, so this wins
// over the FFT path only for small kernels (K <= 5: 12 B/voxel of HBM traffic instead of 48).
// Several CTAs (39 KB each for K = 5) share an SM, so one CTA's TMA wait hides behind the others' FMAs.
#pragma once
#include "pvd_common.cuh"

#ifndef PVD_EMULATE
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#else
struct CUtensorMap { unsigned long long opaque[16]; };
#define __grid_constant__
#endif

namespace pvd {

// Device-side watchdog of every TMA wait (~10 s at 1.9 GHz): a transaction that never completes (bad descriptor) raises a
// flag in the workspace instead of hanging the GPU; the host API reads and clears the flag on every host-returning path.
constexpr long long kWatchdogCycles = 20000000000LL;
constexpr int kDirTX = 8, kDirTY = 8, kDirTZ = 64, kDirThreads = 256;

struct DirectArgs {
    const float* in;     // activity volume (used directly only by the CPU emulation of the TMA load)
    int n0, n1, n2;
    const float* taps;   // flipped kernel kf[i] = k[K-1-i], dense [K0][K1][K2]
    int K0, K1;
    int o0, o1, o2;      // box origin = block origin + o; o0,o1 = c - (K-1) <= 0, o2 = -4 (16-byte aligned)
    int bx, by, bz;      // box extents (bz = 72: 4 aligned lead-in + 64 + right reach, multiple of 4)
    float* out;
    const float* density;
    float rho_ref, rho_min, rho_cut, scale;
    int* error_flag;     // set when the TMA transaction never completes (bad descriptor) instead of hanging
};

#ifndef PVD_EMULATE
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned phase) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1),
        "r"(c2)
        : "memory");
}
#endif

#ifndef PVD_EMULATE
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// bulk L2 prefetch of one box (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
// spin on an mbarrier phase with a ~10 s watchdog: a broken descriptor raises *flag instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_guarded(unsigned long long* bar, unsigned phase, int* flag, int code) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, phase)) {
        if (clock64() - t0 > kWatchdogCycles) {
            if (threadIdx.x == 0) *flag = code;
            break;
        }
    }
}
#endif

// KX > 0: cubic fast path with K0 == KX known at compile time; KX == 0: any K0 <= 9 (runtime loop).
template <int KZ, int KX>
__global__ void __launch_bounds__(kDirThreads) direct_conv_kernel(const __grid_constant__ CUtensorMap tmap, const DirectArgs g) {
    // The TMA box must start on a 16-byte boundary along z (measured: a misaligned innermost coordinate
    // raises "illegal instruction", scripts/tma_probe), so the box starts at z0 - 4 and the taps index
    // with a constant shift ZS = 4 - (left reach of the kernel).
    constexpr int ZS = 4 - (KZ - 1 - KZ / 2);
    constexpr int RL = 3;  // float4 loads per input run: 12 floats cover ZS + 3 + KZ - 1 <= 11
    static_assert(ZS >= 0 && ZS + 3 + KZ - 1 <= 4 * RL - 1, "kernel too long for the staged run");
    PVD_DYN_SMEM(unsigned char, raw);
    // 128-byte aligned carve-up: [tile bx*by*bz floats][taps K0*K1*KZ floats][mbarrier]
    // 128-byte aligned start, derived as an OFFSET from the dynamic shared-memory symbol so that the compiler keeps the
    // accesses on the shared path (LDS) - rounding the pointer through an integer made them generic loads (LD)
    float* tile = reinterpret_cast<float*>(raw + ((128u - ((unsigned)reinterpret_cast<uintptr_t>(raw) & 127u)) & 127u));
    const int box = g.bx * g.by * g.bz;
    float* taps = tile + box;
    const int ntaps = g.K0 * g.K1 * KZ;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(taps + ((ntaps + 3) & ~3) + 2);
    bar = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(bar) + 7) & ~(uintptr_t)7);
    const int tid = threadIdx.x;
    const int z0 = blockIdx.x * kDirTZ, y0 = blockIdx.y * kDirTY, x0 = blockIdx.z * kDirTX;
#ifdef PVD_EMULATE
    // CPU emulation of the TMA box load: gather with zero fill outside the volume
    for (int i = tid; i < box; i += blockDim.x) {
        const int bzz = i % g.bz, t = i / g.bz, byy = t % g.by, bxx = t / g.by;
        const int gx = x0 + g.o0 + bxx, gy = y0 + g.o1 + byy, gz = z0 + g.o2 + bzz;
        const bool inside = gx >= 0 && gx < g.n0 && gy >= 0 && gy < g.n1 && gz >= 0 && gz < g.n2;
        tile[i] = inside ? g.in[((size_t)gx * g.n1 + gy) * g.n2 + gz] : 0.f;
    }
    for (int i = tid; i < ntaps; i += blockDim.x) taps[i] = g.taps[i];
    __syncthreads();
#else
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (unsigned)(box * sizeof(float)));
        tma_load_3d(tile, &tmap, bar, z0 + g.o2, y0 + g.o1, x0 + g.o0);  // innermost coordinate first
    }
    for (int i = tid; i < ntaps; i += kDirThreads) taps[i] = __ldg(g.taps + i);
    {
        const long long t0 = clock64();
        while (!mbar_try_wait(bar, 0)) {
            if (clock64() - t0 > kWatchdogCycles) {  // ~10 s: a broken descriptor must not hang the GPU
                if (tid == 0) *g.error_flag = 1;
                break;
            }
        }
    }
    __syncthreads();
#endif
    const int zq = tid & 15, y = (tid >> 4) & 7, xh = tid >> 7;
    float acc[4][4];
    PVD_UNROLL
    for (int x = 0; x < 4; ++x) {
        PVD_UNROLL
        for (int z = 0; z < 4; ++z) acc[x][z] = 0.f;
    }
    const int rowstride = g.bz, planestride = g.by * g.bz;
    if constexpr (KX > 0) {
        // Cubic-kernel fast path (K0 == KX compile time): per ky the KX*KZ taps sit in registers and each of
        // the 4+KX-1 input rows this thread needs is loaded ONCE (3 x LDS.128) and reused by all four x
        // outputs - 2.5x fewer shared-memory loads than the generic loop below (ncu: that one is l1tex-bound).
        for (int ky = 0; ky < g.K1; ++ky) {
            float t[KX][KZ];
            PVD_UNROLL
            for (int kx = 0; kx < KX; ++kx) {
                PVD_UNROLL
                for (int kz = 0; kz < KZ; ++kz) t[kx][kz] = taps[(kx * g.K1 + ky) * KZ + kz];
            }
            const float* rowbase = tile + (xh * 4) * planestride + (y + ky) * rowstride + 4 * zq;
            PVD_UNROLL
            for (int r = 0; r < 4 + KX - 1; ++r) {
                float rr[RL * 4];
                const float4* rp = reinterpret_cast<const float4*>(rowbase + r * planestride);
                PVD_UNROLL
                for (int i = 0; i < RL; ++i) {
                    const float4 v = rp[i];
                    rr[4 * i] = v.x;
                    rr[4 * i + 1] = v.y;
                    rr[4 * i + 2] = v.z;
                    rr[4 * i + 3] = v.w;
                }
                PVD_UNROLL
                for (int x = 0; x < 4; ++x) {
                    const int kx = r - x;  // compile time after unrolling
                    if (kx >= 0 && kx < KX) {
                        PVD_UNROLL
                        for (int kz = 0; kz < KZ; ++kz) {
                            PVD_UNROLL
                            for (int z = 0; z < 4; ++z) acc[x][z] = fmaf(t[kx][kz], rr[z + kz + ZS], acc[x][z]);
                        }
                    }
                }
            }
        }
    } else
    for (int kx = 0; kx < g.K0; ++kx) {
        for (int ky = 0; ky < g.K1; ++ky) {
            float t[KZ];
            const float* tp = taps + (kx * g.K1 + ky) * KZ;
            PVD_UNROLL
            for (int kz = 0; kz < KZ; ++kz) t[kz] = tp[kz];
            const float* rowbase = tile + (xh * 4 + kx) * planestride + (y + ky) * rowstride + 4 * zq;
            PVD_UNROLL
            for (int x = 0; x < 4; ++x) {
                float r[RL * 4];
                const float4* rp = reinterpret_cast<const float4*>(rowbase + x * planestride);
                PVD_UNROLL
                for (int i = 0; i < RL; ++i) {
                    const float4 v = rp[i];
                    r[4 * i] = v.x;
                    r[4 * i + 1] = v.y;
                    r[4 * i + 2] = v.z;
                    r[4 * i + 3] = v.w;
                }
                PVD_UNROLL
                for (int kz = 0; kz < KZ; ++kz) {
                    PVD_UNROLL
                    for (int z = 0; z < 4; ++z) acc[x][z] = fmaf(t[kz], r[z + kz + ZS], acc[x][z]);
                }
            }
        }
    }
    const int gz = z0 + 4 * zq, gy = y0 + y;
    if (gz < g.n2 && gy < g.n1) {  // n2 % 4 == 0 is a precondition of this path
        PVD_UNROLL
        for (int x = 0; x < 4; ++x) {
            const int gx = x0 + xh * 4 + x;
            if (gx < g.n0) {
                const size_t off = ((size_t)gx * g.n1 + gy) * g.n2 + gz;
                float4 v = make_float4(acc[x][0] * g.scale, acc[x][1] * g.scale, acc[x][2] * g.scale, acc[x][3] * g.scale);
                if (g.density) {
                    const float4 rho = *reinterpret_cast<const float4*>(g.density + off);
                    v.x = (rho.x < g.rho_cut) ? 0.f : v.x * __fdividef(g.rho_ref, fmaxf(rho.x, g.rho_min));
                    v.y = (rho.y < g.rho_cut) ? 0.f : v.y * __fdividef(g.rho_ref, fmaxf(rho.y, g.rho_min));
                    v.z = (rho.z < g.rho_cut) ? 0.f : v.z * __fdividef(g.rho_ref, fmaxf(rho.z, g.rho_min));
                    v.w = (rho.w < g.rho_cut) ? 0.f : v.w * __fdividef(g.rho_ref, fmaxf(rho.w, g.rho_min));
                }
                *reinterpret_cast<float4*>(g.out + off) = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cubic kernels (K = 3, 5, 7), both boundary modes.
//
//  * the K^3 taps travel inside the __grid_constant__ argument struct: after full unrolling every tap is a
//    constant-bank operand of its FFMA (c[0x0][imm]) - no tap registers, no tap loads;
//  * each thread owns a 4 (x) x 2 (y) x 4 (z) register tile: an input row (3 x LDS.128) is loaded once and feeds
//    4 x-outputs and 2 y-outputs - 0.6 of the shared-memory traffic of the 4 x 1 x 4 tile above, which ncu showed
//    bound by the shared-memory pipe (l1tex 99 %);
//  * CTA tile 8 x 16 x 64; its (8+K-1) x (16+K-1) x 72 halo box is ONE TMA load; several CTAs per SM
//    (69 KB each for K = 5) keep the FMA pipe fed while others wait for their box;
//  * WRAP (reference / circular boundary, kernel anchored at the origin: core/kernel_convolution.py:71-74): the halo
//    reaches only towards lower indices; tiles whose box crosses index 0 on any axis (one tile row per axis) gather
//    their box with modulo indexing instead of the TMA load, every other tile takes the TMA path unchanged.
constexpr int kCubTX = 8, kCubTY = 16, kCubTZ = 64, kCubThreads = 256;

template <int K>
struct CubicArgs {
    float taps[K * K * K];  // flipped kernel kf[i] = k[K-1-i], [x][y][z]
    const float* in;
    int n0, n1, n2;
    int o0, o1, o2;         // box origin = block origin + o
    float* out;
    const float* density;
    float rho_ref, rho_min, rho_cut, scale;
    int wrap;               // circular boundary
    int* error_flag;
};

// ZS: index of the first tap inside a staged run = (global z of tap 0 for output z) - (box origin z) - z
template <int K, int ZS>
__global__ void __launch_bounds__(kCubThreads) direct_conv_cubic_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                        const __grid_constant__ CubicArgs<K> g) {
    constexpr int BX = kCubTX + K - 1, BY = kCubTY + K - 1, BZ = kCubTZ + 8;
    constexpr int RL = 3;  // float4 loads per input run: 12 floats cover ZS + 3 + K - 1
    static_assert(ZS >= 0 && ZS + 3 + K - 1 <= 4 * RL - 1, "kernel too long for the staged run");
    PVD_DYN_SMEM(unsigned char, raw);
    // 128-byte aligned start, derived as an OFFSET from the dynamic shared-memory symbol so that the compiler keeps the
    // accesses on the shared path (LDS) - rounding the pointer through an integer made them generic loads (LD)
    float* tile = reinterpret_cast<float*>(raw + ((128u - ((unsigned)reinterpret_cast<uintptr_t>(raw) & 127u)) & 127u));
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tile + BX * BY * BZ);
    const int tid = threadIdx.x;
    const int z0 = blockIdx.x * kCubTZ, y0 = blockIdx.y * kCubTY, x0 = blockIdx.z * kCubTX;
    const int bx0 = x0 + g.o0, by0 = y0 + g.o1, bz0 = z0 + g.o2;
#ifdef PVD_EMULATE
    // CPU emulation of the box load: zero-filled gather
    for (int i = tid; i < BX * BY * BZ; i += kCubThreads) {
        const int bz = i % BZ, t = i / BZ, by = t % BY, bx = t / BY;
        const int gx = bx0 + bx, gy = by0 + by, gz = bz0 + bz;
        const bool inside = gx >= 0 && gx < g.n0 && gy >= 0 && gy < g.n1 && gz >= 0 && gz < g.n2;
        tile[i] = inside ? g.in[((size_t)gx * g.n1 + gy) * g.n2 + gz] : 0.f;
    }
    __syncthreads();
#else
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (unsigned)(BX * BY * BZ * sizeof(float)));
        tma_load_3d(tile, &tmap, bar, bz0, by0, bx0);  // innermost coordinate first; out of bounds = zeros
    }
    mbar_wait_guarded(bar, 0, g.error_flag, 1);
#endif
    if (g.wrap && (bx0 < 0 || by0 < 0 || bz0 < 0)) {
        // circular boundary: the part of the box below index 0 (zeros so far) is the wrapped data from the far end;
        // only the one tile row per axis that touches index 0 comes here, and only its negative part is fetched
        for (int i = tid; i < BX * BY * BZ; i += kCubThreads) {
            const int bz = i % BZ, t = i / BZ, by = t % BY, bx = t / BY;
            int gx = bx0 + bx, gy = by0 + by, gz = bz0 + bz;
            if (gx < 0 || gy < 0 || gz < 0) {
                if (gx < 0) gx += g.n0;
                if (gy < 0) gy += g.n1;
                if (gz < 0) gz += g.n2;
                // beyond the high end nothing is ever used (the reach is towards lower indices only)
                tile[i] = (gx < g.n0 && gy < g.n1 && gz < g.n2) ? g.in[((size_t)gx * g.n1 + gy) * g.n2 + gz] : 0.f;
            }
        }
        __syncthreads();
    }
    const int zq = tid & 15, yp = (tid >> 4) & 7, xh = tid >> 7;
    // Accumulators as float2 pairs (z, z+1): the arithmetic is fma.rn.f32x2 (FFMA2).  A plain FFMA issues every second
    // cycle per scheduler on this part (the 4 x 1 x 4 kernel above and a scalar version of this one both sat at
    // 17-18 TFMA/s = the scalar FP32 pipe limit); the packed form does two FMAs per issue slot.  Bit-identical results.
    float2 acc[4][2][2];
    PVD_UNROLL
    for (int x = 0; x < 4; ++x) {
        PVD_UNROLL
        for (int y = 0; y < 2; ++y) acc[x][y][0] = acc[x][y][1] = make_float2(0.f, 0.f);
    }
    const float* base = tile + (xh * 4) * (BY * BZ) + (2 * yp) * BZ + 4 * zq;
    PVD_UNROLL
    for (int ry = 0; ry < K + 1; ++ry) {      // input row ry feeds y-output 0 with ky = ry and y-output 1 with ky = ry - 1
        PVD_UNROLL
        for (int rx = 0; rx < 4 + K - 1; ++rx) {  // input plane rx feeds x-output x with kx = rx - x
            float rr[RL * 4];
            const float4* rp = reinterpret_cast<const float4*>(base + rx * (BY * BZ) + ry * BZ);
            PVD_UNROLL
            for (int i = 0; i < RL; ++i) {
                const float4 v = rp[i];
                rr[4 * i] = v.x;
                rr[4 * i + 1] = v.y;
                rr[4 * i + 2] = v.z;
                rr[4 * i + 3] = v.w;
            }
            // the run as pairs starting at every index: (rr[i], rr[i+1]); even starts are the loaded register pairs
            float2 pr[RL * 4 - 1];
            PVD_UNROLL
            for (int i = 0; i < RL * 4 - 1; ++i) pr[i] = make_float2(rr[i], rr[i + 1]);
            PVD_UNROLL
            for (int x = 0; x < 4; ++x) {
                const int kx = rx - x;
                if (kx >= 0 && kx < K) {
                    PVD_UNROLL
                    for (int y = 0; y < 2; ++y) {
                        const int ky = ry - y;
                        if (ky >= 0 && ky < K) {
                            PVD_UNROLL
                            for (int kz = 0; kz < K; ++kz) {
                                const float t = g.taps[(kx * K + ky) * K + kz];  // constant bank -> uniform register
                                const float2 tt = make_float2(t, t);
                                acc[x][y][0] = cfma2(pr[kz + ZS], tt, acc[x][y][0]);
                                acc[x][y][1] = cfma2(pr[kz + ZS + 2], tt, acc[x][y][1]);
                            }
                        }
                    }
                }
            }
        }
    }
    const int gz = z0 + 4 * zq;
    if (gz < g.n2) {  // n2 % 4 == 0 is a precondition of this path
        PVD_UNROLL
        for (int y = 0; y < 2; ++y) {
            const int gy = y0 + 2 * yp + y;
            if (gy >= g.n1) continue;
            PVD_UNROLL
            for (int x = 0; x < 4; ++x) {
                const int gx = x0 + xh * 4 + x;
                if (gx < g.n0) {
                    const size_t off = ((size_t)gx * g.n1 + gy) * g.n2 + gz;
                    float4 v = make_float4(acc[x][y][0].x * g.scale, acc[x][y][0].y * g.scale, acc[x][y][1].x * g.scale, acc[x][y][1].y * g.scale);
                    if (g.density) {
                        const float4 rho = *reinterpret_cast<const float4*>(g.density + off);
                        v.x = (rho.x < g.rho_cut) ? 0.f : v.x * __fdividef(g.rho_ref, fmaxf(rho.x, g.rho_min));
                        v.y = (rho.y < g.rho_cut) ? 0.f : v.y * __fdividef(g.rho_ref, fmaxf(rho.y, g.rho_min));
                        v.z = (rho.z < g.rho_cut) ? 0.f : v.z * __fdividef(g.rho_ref, fmaxf(rho.z, g.rho_min));
                        v.w = (rho.w < g.rho_cut) ? 0.f : v.w * __fdividef(g.rho_ref, fmaxf(rho.w, g.rho_min));
                    }
                    *reinterpret_cast<float4*>(g.out + off) = v;
                }
            }
        }
    }
}

// kf[i0][i1][i2] = k[K0-1-i0][K1-1-i1][K2-1-i2]
__global__ void flip_kernel_kernel(const float* __restrict__ k, float* __restrict__ kf, int K0, int K1, int K2) {
    const int n = K0 * K1 * K2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int i2 = i % K2, t = i / K2, i1 = t % K1, i0 = t / K1;
        kf[i] = k[((K0 - 1 - i0) * K1 + (K1 - 1 - i1)) * K2 + (K2 - 1 - i2)];
    }
}

}  // namespace pvd
