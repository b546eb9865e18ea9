// Persistent row passes with cp.async staging (axis 2, the contiguous axis).
//
// P1  real rows -> half spectrum.  32 real rows are copied row-major into a staging buffer with 16-byte
//     cp.async (zero fill for padding / invalid rows, no registers held); the first radix stage reads the
//     two rows of a line as (re, im) straight from the staging buffer, so the transposition into the
//     [index][line] exchange tile costs no extra pass.  As soon as stage 1 has consumed the staging buffer
//     the next tile's rows are already being fetched while stage 2, the Hermitian split and the stores run.
// P5  half spectrum -> real rows * density.  The complex rows are staged the same way, packed into the
//     Hermitian line Z = A + iB, and while the inverse FFT runs the density rows of the same tile are
//     prefetched into the (now free) staging buffer.
// Two CTAs per SM (109 KB each) interleave so one is always fetching.
#pragma once
#include "fft_pipe.cuh"

namespace pvd {

#ifdef PVD_EMULATE
static inline void cp_async16_partial(void* dst, const void* src, int valid_bytes) {
    std::memset(dst, 0, 16);
    if (valid_bytes > 0) std::memcpy(dst, src, valid_bytes > 16 ? 16 : valid_bytes);
}
#else
__device__ __forceinline__ void cp_async16_partial(void* smem_dst, const void* gsrc, int valid_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid_bytes < 0 ? 0 : (valid_bytes > 16 ? 16 : valid_bytes);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
#endif

// Staging geometry: 32 rows, row stride LSF floats (16-byte multiple).  404 = 400 + 4 keeps the
// (line-strided) stage-1 reads at 2-way bank conflicts.
template <int N>
struct RowStage {
    static constexpr int LSF = ((N + 3) / 4) * 4 + 4;     // floats per staged real row
    static constexpr int CHR = (N + 3) / 4;               // 16-byte chunks per real row
    static constexpr int LSC = LSF / 2;                   // float2 per staged complex row
    static constexpr int CHC = (N / 2 + 1 + 1) / 2;       // 16-byte chunks per complex row (Nh float2)
    static constexpr int BYTES = 32 * LSF * 4;
    static_assert(CHC * 2 <= LSC, "complex row must fit the staging row");
};

template <int N, int NT, int MINB, int R1, int R2, int R3>
__global__ void __launch_bounds__(NT, MINB) rows_fwd_pipe_kernel(const RowFwdArgs g) {
    constexpr int W = 16, LS = 17;
    using RS = RowStage<N>;
    constexpr int LSF = RS::LSF, CHR = RS::CHR;
    PVD_DYN_SMEM(float2, smem);
    float2* tile = smem;
    float* raw = reinterpret_cast<float*>(smem + N * LS);
    float2* tws = reinterpret_cast<float2*>(raw + 32 * LSF);
    Sched<N, R1, R2, R3>::build(tws, g.tw);
    const long long nrows = (long long)g.n0 * g.n1;
    const int ntiles = (int)((nrows + 31) / 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NWARPS = NT / 32;
    const float* __restrict__ in0 = g.in[0];
    const float w0 = g.w[0];
    const int n1 = g.n1, n2b = g.n2 * 4;
    auto issue = [&](int t) {  // one warp per row: the row address is warp-uniform, lanes walk the 16-byte chunks
        const int row0 = t * 32;
        for (int rr = warp; rr < 32; rr += NWARPS) {
            const int R = row0 + rr;
            const bool valid = R < (int)nrows;
            long long off = 0;
            if (valid) {
                const int x = R / n1, y = R - x * n1;
                off = x * g.in_s0 + y * g.in_s1;
            }
            const float* src = in0 + off;
            float* dstp = raw + rr * LSF;
            PVD_UNROLL
            for (int i = 0; i < (CHR + 31) / 32; ++i) {
                const int ch = lane + 32 * i;
                if (ch < CHR) cp_async16_partial(dstp + ch * 4, src + ch * 4, valid ? n2b - ch * 16 : 0);
            }
        }
    };
    int t = blockIdx.x;
    if (t < ntiles) issue(t);
    cp_async_commit();
    const int wl = threadIdx.x % W;
    const float* rawA = raw + (2 * wl) * LSF;
    const float* rawB = rawA + LSF;
    for (; t < ntiles; t += gridDim.x) {
        cp_async_wait<0>();
        __syncthreads();  // staged rows of tile t visible; previous tile's split phase finished with `tile`
        auto raw_in = [&](int, int, int idx, int) -> float2 { return make_float2(w0 * rawA[idx], w0 * rawB[idx]); };
        auto sm_in = [&](int, int, int idx, int w) -> float2 { return tile[idx * LS + w]; };
        auto sm_out = [&](int, int, int idx, int w, float2 v) { tile[idx * LS + w] = v; };
        fast_stage<N, W, NT, R1, 1, -1, false>(raw_in, sm_out, tws);
        __syncthreads();  // staging buffer consumed -> refill it with the next tile while the rest runs
        const int tn = t + gridDim.x;
        if (tn < ntiles) issue(tn);
        cp_async_commit();
        if constexpr (R3 > 1) {
            fast_stage<N, W, NT, R2, R1, -1, true>(sm_in, sm_out, tws + Sched<N, R1, R2, R3>::T1);
            __syncthreads();
            fast_stage<N, W, NT, R3, R1 * R2, -1, true>(sm_in, sm_out, tws);
        } else {
            fast_stage<N, W, NT, R2, R1, -1, true>(sm_in, sm_out, tws);
        }
        __syncthreads();
        // Hermitian split: A[k] = (Z[k] + conj(Z[N-k]))/2, B[k] = (Z[k] - conj(Z[N-k]))/(2i)
        const int row0 = t * 32;
        const int Nh = g.Nh;
        for (int rr = warp; rr < 32; rr += NWARPS) {
            const int R = row0 + rr;
            if (R >= (int)nrows) continue;
            const int x = R / n1, y = R - x * n1;
            float2* __restrict__ dst = g.out + x * g.out_s0 + y * g.out_s1;
            const int line = rr >> 1;
            const bool odd = rr & 1;
            constexpr int KIT = (N / 2 + 1 + 31) / 32;
            PVD_UNROLL
            for (int i = 0; i < KIT; ++i) {
                const int k = lane + 32 * i;
                if (k < Nh) {
                    const float2 zk = tile[k * LS + line];
                    const float2 zm = tile[((k == 0) ? 0 : N - k) * LS + line];
                    dst[k] = odd ? make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x))
                                 : make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                }
            }
        }
    }
    cp_async_wait<0>();
}

template <int N, int NT, int MINB, int R1, int R2, int R3>
__global__ void __launch_bounds__(NT, MINB) rows_inv_pipe_kernel(const RowInvArgs g) {
    constexpr int W = 16, LS = 17;
    using RS = RowStage<N>;
    constexpr int LSF = RS::LSF, LSC = RS::LSC, CHR = RS::CHR, CHC = RS::CHC;
    PVD_DYN_SMEM(float2, smem);
    float2* tile = smem;
    float2* rawc = smem + N * LS;                       // staged complex rows, later the density rows
    float* rawf = reinterpret_cast<float*>(rawc);
    float2* tws = rawc + 32 * LSC;
    Sched<N, R1, R2, R3>::build(tws, g.tw);
    const long long nrows = (long long)g.O0 * g.O1;
    const int ntiles = (int)((nrows + 31) / 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NWARPS = NT / 32;
    const int Nh = g.Nh, O1 = g.O1, O2 = g.O2, z_lo = g.z_lo;
    const bool has_den = g.density != nullptr;
    auto issue_spec = [&](int t) {
        const int row0 = t * 32;
        for (int rr = warp; rr < 32; rr += NWARPS) {
            const int R = row0 + rr;
            const bool valid = R < (int)nrows;
            long long off = 0;
            if (valid) {
                const int x = R / O1, y = R - x * O1;
                off = (x + g.x_lo) * g.in_s0 + (y + g.y_lo) * g.in_s1;
            }
            const float2* src = g.in + off;
            float2* dstp = rawc + rr * LSC;
            PVD_UNROLL
            for (int i = 0; i < (CHC + 31) / 32; ++i) {
                const int ch = lane + 32 * i;
                if (ch < CHC) cp_async16(dstp + ch * 2, src + ch * 2, valid);
            }
        }
    };
    auto issue_density = [&](int t) {
        const int row0 = t * 32;
        const int o2b = O2 * 4;
        for (int rr = warp; rr < 32; rr += NWARPS) {
            const int R = row0 + rr;
            const bool valid = R < (int)nrows;
            long long off = 0;
            if (valid) {
                const int x = R / O1, y = R - x * O1;
                off = x * g.den_s0 + y * g.den_s1;
            }
            const float* src = g.density + off;
            float* dstp = rawf + rr * LSF;
            PVD_UNROLL
            for (int i = 0; i < (CHR + 31) / 32; ++i) {
                const int ch = lane + 32 * i;
                if (ch < CHR) cp_async16_partial(dstp + ch * 4, src + ch * 4, valid ? o2b - ch * 16 : 0);
            }
        }
    };
    int t = blockIdx.x;
    if (t < ntiles) issue_spec(t);
    cp_async_commit();
    const float scale = g.scale, rho_ref = g.rho_ref, rho_min = g.rho_min, rho_cut = g.rho_cut;
    for (; t < ntiles; t += gridDim.x) {
        cp_async_wait<0>();
        __syncthreads();
        // rebuild the packed Hermitian line Z = A + i*B for each pair of rows (lanes along k)
        for (int line = warp; line < W; line += NWARPS) {
            const float2* pa = rawc + (2 * line) * LSC;
            const float2* pb = pa + LSC;
            constexpr int KIT = (N / 2 + 1 + 31) / 32;
            PVD_UNROLL
            for (int i = 0; i < KIT; ++i) {
                const int k = lane + 32 * i;
                if (k < Nh) {
                    float2 a = pa[k], b = pb[k];
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
        __syncthreads();  // staging buffer consumed
        const int tn = t + gridDim.x;
        // den_ldg: density comes through batched LDGs in the store phase, so the staging buffer is free NOW and
        // the next tile's spectrum rows stream in during the inverse FFT and the stores.  Otherwise the density
        // rows of this tile are staged here and the next tile is requested only after the store phase.
        if (g.den_ldg) {
            if (tn < ntiles) issue_spec(tn);
        } else if (has_den) {
            issue_density(t);
        }
        cp_async_commit();
        auto sm_in = [&](int, int, int idx, int w) -> float2 { return tile[idx * LS + w]; };
        auto sm_out = [&](int, int, int idx, int w, float2 v) { tile[idx * LS + w] = v; };
        fast_fft<N, W, LS, NT, +1, R1, R2, R3, true, true>(sm_in, sm_out, tile, tws);
        if (!g.den_ldg) cp_async_wait<0>();
        __syncthreads();  // transform done (and staged density rows landed)
        const int row0 = t * 32;
        const float* Zf = reinterpret_cast<const float*>(tile);
        for (int rr = warp; rr < 32; rr += NWARPS) {
            const int R = row0 + rr;
            if (R >= (int)nrows) continue;
            const int x = R / O1, y = R - x * O1;
            float* __restrict__ dst = g.out + x * g.out_s0 + y * g.out_s1;
            const float* srcf = Zf + (rr >> 1) * 2 + (rr & 1) + (size_t)z_lo * (2 * LS);
            constexpr int ZIT = (N + 31) / 32;
            if (has_den && g.den_ldg) {
                const float* __restrict__ dg = g.density + x * g.den_s0 + y * g.den_s1;
                float rho[ZIT];
                PVD_UNROLL
                for (int i = 0; i < ZIT; ++i) {
                    const int z = lane + 32 * i;
                    rho[i] = (z < O2) ? __ldg(dg + z) : 1.f;
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
                const float* den = rawf + rr * LSF;
                PVD_UNROLL
                for (int i = 0; i < ZIT; ++i) {
                    const int z = lane + 32 * i;
                    if (z < O2) {
                        float v = srcf[z * (2 * LS)] * scale;
                        if (has_den) {
                            const float rho = den[z];
                            v = (rho < rho_cut) ? 0.f : v * __fdividef(rho_ref, fmaxf(rho, rho_min));
                        }
                        dst[z] = v;
                    }
                }
            }
        }
        if (!g.den_ldg) {
            __syncthreads();  // staging buffer (density) and tile free again
            if (tn < ntiles) issue_spec(tn);
            cp_async_commit();
        }
    }
    cp_async_wait<0>();
}

}  // namespace pvd
