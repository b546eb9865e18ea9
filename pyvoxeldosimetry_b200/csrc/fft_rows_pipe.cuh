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
__global__ void __launch_bounds__(NT, MINB) rows_fwd_pipe_kernel(const __grid_constant__ RowFwdArgs g) {
    constexpr int W = 16, LS = 17;
    using RS = RowStage<N>;
    constexpr int LSF = RS::LSF, CHR = RS::CHR;
    // TMA staging: the 32 rows arrive as 4 boxes of N/4 floats x 32 rows, laid out [quarter][row][N/4] (a 100-float row
    // pitch keeps the line-strided first-stage reads at the same 2-way bank conflicts as the 404-float pitch above)
    constexpr int QW = N / 4;
    constexpr bool TMA_OK = (N % 16 == 0) && (QW <= 256) && (32 * N * 4 <= RS::BYTES);
    PVD_DYN_SMEM(float2, smem);
    float2* tile = smem;
    float* raw = reinterpret_cast<float*>(smem + N * LS);
    float2* tws = reinterpret_cast<float2*>(raw + 32 * LSF);
#ifndef PVD_EMULATE
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tws + Sched<N, R1, R2, R3>::TOTAL);
    const bool tma = TMA_OK && g.use_tma && ((unsigned)__cvta_generic_to_shared(raw) & 127u) == 0;
    if (tma && threadIdx.x == 0) mbar_init(bar, 1);
    unsigned ph = 0;
#else
    const bool tma = false;
#endif
    grid_dep_launch();
    Sched<N, R1, R2, R3>::build(tws, g.tw);  // tables come from the plan's twiddle buffer (complete since plan set-up)
    if (tma) __syncthreads();                // mbarrier initialised before anyone polls it
    grid_dep_wait();                         // predecessor grid done: its output / our output buffer may be touched now
    const long long nrows = (long long)g.n0 * g.n1;
    const int ntiles = (int)((nrows + 31) / 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NWARPS = NT / 32;
    const float* __restrict__ in0 = g.in[0];
    const float w0 = 0.5f * g.w[0];  // the 1/2 of the Hermitian split, applied once at the input
    const int n1 = g.n1, n2b = g.n2 * 4;
    const bool dense = g.dense != 0;  // row address = base + linear row index * row stride: no (x, y) bookkeeping
    // (x, y) of a tile row from the tile's first row: one division per tile instead of one per row
    auto row_xy = [&](int x0, int y0, int rr, int& x, int& y) {
        x = x0;
        y = y0 + rr;
        while (y >= n1) {
            y -= n1;
            ++x;
        }
    };
    auto issue = [&](int t, int x0, int y0) {  // one warp per row: the row address is warp-uniform, lanes walk the 16-byte chunks
        const int row0 = t * 32;
        for (int rr = warp; rr < 32; rr += NWARPS) {
            const bool valid = row0 + rr < (int)nrows;
            long long off = 0;
            if (valid) {
                if (g.dense_in) {
                    off = (long long)(row0 + rr) * g.in_s1;
                } else {
                    int x, y;
                    row_xy(x0, y0, rr, x, y);
                    off = x * g.in_s0 + y * g.in_s1;
                }
            }
            const float* src = in0 + off;
            float* dstp = raw + rr * LSF;
            const int nfull = valid ? (n2b >> 4) : 0;  // chunks entirely inside the row; at most one partial chunk follows
            const int rem = valid ? (n2b & 15) : 0;
            PVD_UNROLL
            for (int i = 0; i < (CHR + 31) / 32; ++i) {
                const int ch = lane + 32 * i;
                if (ch < nfull) cp_async16_full(dstp + ch * 4, src + ch * 4);
                else if (ch == nfull && rem) cp_async16_partial(dstp + ch * 4, src + ch * 4, rem);
                else if (ch < CHR) st_zero16(dstp + ch * 4);
            }
        }
    };
    int t = blockIdx.x;
    // (x, y) of the first row of the current tile and of the next one, advanced without divisions
    const int step = 32 * (int)gridDim.x;
    const int sx = step / n1, sy = step - sx * n1;
    int cx = (32 * t) / n1, cy = 32 * t - cx * n1;
    auto advance = [&](int& x, int& y) {
        x += sx;
        y += sy;
        if (y >= n1) {
            y -= n1;
            ++x;
        }
    };
#ifndef PVD_EMULATE
    auto issue_tma = [&](int tt) {  // one thread: 4 box copies, rows [32 tt, 32 tt + 32), out-of-range rows / columns arrive as zeros
        mbar_expect_tx(bar, 32u * N * 4u);
        PVD_UNROLL
        for (int q = 0; q < 4; ++q) tma_load_2d(raw + q * (32 * QW), &g.tmap, bar, q * QW, tt * 32);
    };
#endif
    if (!tma) {
        if (t < ntiles) issue(t, cx, cy);
        cp_async_commit();
    }
#ifndef PVD_EMULATE
    else if (t < ntiles && threadIdx.x == 0) issue_tma(t);
#endif
    const int wl = threadIdx.x % W;
    // line wl packs staged rows wl (re) and wl + 16 (im): row starts are then 20 (mod 32) banks apart for
    // LSF = 404, i.e. 2-way conflicts over the 16 lines of a warp (adjacent rows 2wl, 2wl+1 gave 4-way)
    const float* rawA = raw + wl * LSF;
    const float* rawB = rawA + W * LSF;
    const float* rawTA = raw + wl * QW;  // TMA layout
    const float* rawTB = rawTA + W * QW;
    for (; t < ntiles; t += gridDim.x) {
        if (!tma) cp_async_wait<0>();
#ifndef PVD_EMULATE
        else {
            mbar_wait_guarded(bar, ph, g.error_flag, 3);
            ph ^= 1;
        }
#endif
        __syncthreads();  // staged rows of tile t visible; previous tile's split phase finished with `tile`
        auto raw_in = [&](int, int, int idx, int) -> float2 { return make_float2(w0 * rawA[idx], w0 * rawB[idx]); };  // w0 = w/2
        auto raw_in_tma = [&](int, int j, int idx, int) -> float2 {
            // quarter of idx: a literal for every slot j whose index range stays inside one quarter
            constexpr int NBl = N / R1;
            const int lo = NBl * j, hi = lo + NBl - 1;
            const int q = (lo / QW == hi / QW) ? lo / QW : idx / QW;
            const int o = q * (32 * QW - QW) + idx;  // q * 32 * QW + (idx - q * QW)
            return make_float2(w0 * rawTA[o], w0 * rawTB[o]);
        };
        auto sm_in = [&](int, int, int idx, int w) -> float2 { return tile[idx * LS + w]; };
        auto sm_out = [&](int, int, int idx, int w, float2 v) { tile[idx * LS + w] = v; };
        if (tma) fast_stage<N, W, NT, R1, 1, -1, false>(raw_in_tma, sm_out, tws);
        else fast_stage<N, W, NT, R1, 1, -1, false>(raw_in, sm_out, tws);
        __syncthreads();  // staging buffer consumed -> refill it with the next tile while the rest runs
        const int tn = t + gridDim.x;
        int nx = cx, ny = cy;
        advance(nx, ny);
        if (!tma) {
            if (tn < ntiles) issue(tn, nx, ny);
            cp_async_commit();
        }
#ifndef PVD_EMULATE
        else if (tn < ntiles && threadIdx.x == 0) issue_tma(tn);
#endif
        if constexpr (R3 > 1) {
            fast_stage<N, W, NT, R2, R1, -1, true>(sm_in, sm_out, tws + Sched<N, R1, R2, R3>::T1);
            __syncthreads();
            fast_stage<N, W, NT, R3, R1 * R2, -1, true>(sm_in, sm_out, tws);
        } else {
            fast_stage<N, W, NT, R2, R1, -1, true>(sm_in, sm_out, tws);
        }
        __syncthreads();
        // Hermitian split: A[k] = (Z[k] + conj(Z[N-k]))/2, B[k] = (Z[k] - conj(Z[N-k]))/(2i).  One warp per LINE: the
        // pair (Z[k], Z[N-k]) is read once and yields both rows; the factor 1/2 is already in the input weight.
        const int row0 = t * 32;
        const int x0 = cx, y0 = cy;
        const int Nh = g.Nh;
        for (int line = warp; line < W; line += NWARPS) {
            const bool va = row0 + line < (int)nrows, vb = row0 + line + W < (int)nrows;
            if (!va) continue;  // rows are consecutive: b valid implies a valid
            float2* __restrict__ dsta;
            float2* __restrict__ dstb;
            if (dense) {
                dsta = g.out + (long long)(row0 + line) * g.out_s1;
                dstb = vb ? dsta + W * g.out_s1 : dsta;
            } else {
                int x, y;
                row_xy(x0, y0, line, x, y);
                dsta = g.out + x * g.out_s0 + y * g.out_s1;
                dstb = dsta;
                if (vb) {
                    row_xy(x0, y0, line + W, x, y);
                    dstb = g.out + x * g.out_s0 + y * g.out_s1;
                }
            }
            constexpr int KIT = (N / 2 + 1 + 31) / 32;
            PVD_UNROLL
            for (int i = 0; i < KIT; ++i) {
                const int k = lane + 32 * i;
                if (k < Nh) {
                    const float2 zk = tile[k * LS + line];
                    const float2 zm = tile[((k == 0) ? 0 : N - k) * LS + line];
                    dsta[k] = make_float2(zk.x + zm.x, zk.y - zm.y);
                    if (vb) dstb[k] = make_float2(zk.y + zm.y, zm.x - zk.x);
                }
            }
        }
        cx = nx;
        cy = ny;
    }
    cp_async_wait<0>();
}

// 128-bit global accesses of the store phase (density rows are read once, dose rows written once)
#ifdef PVD_EMULATE
static inline float4 ldg128_ro(const float* p) { return *reinterpret_cast<const float4*>(p); }
static inline void stg128(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
#else
__device__ __forceinline__ float4 ldg128_ro(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg128(float* p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
#endif

// 1/x as ONE MUFU.RCP (rcp.approx.ftz, 1 ulp): __fdividef wraps it in a denormal-rescaling sequence (5 more
// instructions per voxel) that max(rho, rho_min) never needs
__device__ __forceinline__ float fast_rcp(float x) {
#ifdef PVD_EMULATE
    return 1.0f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
// dose = v * sr / max(rho, rho_min), zero below rho_cut   (sr = scale * rho_ref)
__device__ __forceinline__ float den_apply(float v, float rho, float sr, float rho_min, float rho_cut) {
    const float f = v * (sr * fast_rcp(fmaxf(rho, rho_min)));
    return (rho < rho_cut) ? 0.f : f;
}

// P5.  The last radix stage does not return to the [index][line] exchange layout: it writes the two real rows
// of every line ROW-MAJOR (and already cropped: index - z_lo) over the tile, so the store phase is a straight
// copy with 128-bit shared loads, 128-bit density loads and 128-bit dose stores - a quarter of the memory
// instructions and guards of the element-wise version.  Row r of the tile pairs with row r + 16 in one line.
template <int N, int NT, int MINB, int R1, int R2, int R3>
__global__ void __launch_bounds__(NT, MINB) rows_inv_pipe_kernel(const __grid_constant__ RowInvArgs g) {
    constexpr int W = 16, LS = 17;
    using RS = RowStage<N>;
    constexpr int LSC = RS::LSC, CHC = RS::CHC;
    constexpr int LSA = RS::LSF;  // floats per row of the row-major result (16-byte multiple, 32 rows fit the tile)
    static_assert(32 * LSA * 4 <= N * LS * 8, "row-major result must fit the exchange tile");
    PVD_DYN_SMEM(float2, smem);
    float2* tile = smem;
    float* rowbuf = reinterpret_cast<float*>(smem);
    float2* rawc = smem + N * LS;  // staged complex rows
    float2* tws = rawc + 32 * LSC;
    // TMA staging: the 32 half-spectrum rows are ONE box (LSC 8-byte elements x 32 rows) of a 2-D map of the work
    // buffer; the box lands in exactly the [row][LSC] layout the cp.async path builds.  The density rows of the tile
    // are pulled into L2 by one bulk prefetch per tile instead of one prefetch instruction per 128-byte line.
    constexpr bool TMA_OK = LSC <= 256;
#ifndef PVD_EMULATE
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tws + Sched<N, R1, R2, R3>::TOTAL);
    const bool tma = TMA_OK && g.use_tma && ((unsigned)__cvta_generic_to_shared(rawc) & 127u) == 0;
    const bool tma_den = g.use_tma_den != 0;
    if (tma && threadIdx.x == 0) mbar_init(bar, 1);
    unsigned ph = 0;
#else
    const bool tma = false, tma_den = false;
#endif
    grid_dep_launch();
    Sched<N, R1, R2, R3>::build(tws, g.tw);
    if (tma) __syncthreads();  // mbarrier initialised before anyone polls it
    grid_dep_wait();
    const long long nrows = (long long)g.O0 * g.O1;
    const int ntiles = (int)((nrows + 31) / 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NWARPS = NT / 32;
    const int Nh = g.Nh, O1 = g.O1, O2 = g.O2, z_lo = g.z_lo;
    const bool has_den = g.density != nullptr;
    const bool vec4 = g.vec4 != 0;
    const bool dense = g.dense != 0;          // row address = base + linear row index * row stride
    const bool plain_den = g.plain_den != 0;  // dose = v / max(rho, rho_min)
    // (x, y) of a tile row from the tile's first row: one division per tile instead of one per row
    auto row_xy = [&](int x0, int y0, int rr, int& x, int& y) {
        x = x0;
        y = y0 + rr;
        while (y >= O1) {
            y -= O1;
            ++x;
        }
    };
    auto issue_spec = [&](int t, int x0, int y0) {
        const int row0 = t * 32;
        for (int rr = warp; rr < 32; rr += NWARPS) {
            const bool valid = row0 + rr < (int)nrows;
            long long off = 0;
            if (valid) {
                if (dense) {
                    off = (long long)(row0 + rr) * g.in_s1;
                } else {
                    int x, y;
                    row_xy(x0, y0, rr, x, y);
                    off = (x + g.x_lo) * g.in_s0 + (y + g.y_lo) * g.in_s1;
                }
            }
            const float2* src = g.in + off;
            float2* dstp = rawc + rr * LSC;
            PVD_UNROLL
            for (int i = 0; i < (CHC + 31) / 32; ++i) {
                const int ch = lane + 32 * i;
                if (ch < CHC) {
                    if (valid) cp_async16_full(dstp + ch * 2, src + ch * 2);
                    else st_zero16(dstp + ch * 2);
                }
            }
        }
    };
    int t = blockIdx.x;
    // (x, y) of the first row of the current tile and of the next one, advanced without divisions
    const int step = 32 * (int)gridDim.x;
    const int sx = step / O1, sy = step - sx * O1;
    int cx = (32 * t) / O1, cy = 32 * t - cx * O1;
    auto advance = [&](int& x, int& y) {
        x += sx;
        y += sy;
        if (y >= O1) {
            y -= O1;
            ++x;
        }
    };
#ifndef PVD_EMULATE
    auto issue_tma = [&](int tt, int x, int y) {  // one thread, one box: the tile's 32 rows of the work buffer; rows past the end arrive as zeros
        mbar_expect_tx(bar, 32u * LSC * 8u);
        if (g.tma3d) tma_load_3d(rawc, &g.tmap, bar, 0, y + g.y_lo, x + g.x_lo);
        else tma_load_2d(rawc, &g.tmap, bar, 0, tt * 32);
    };
#endif
    if (!tma) {
        if (t < ntiles) issue_spec(t, cx, cy);
        cp_async_commit();
    }
#ifndef PVD_EMULATE
    else if (t < ntiles && threadIdx.x == 0) issue_tma(t, cx, cy);
#endif
    const float sr = g.scale * (has_den ? g.rho_ref : 1.f), rho_min = g.rho_min, rho_cut = g.rho_cut;
    const int wl = threadIdx.x % W;
    for (; t < ntiles; t += gridDim.x) {
        if (!tma) cp_async_wait<0>();
#ifndef PVD_EMULATE
        else {
            mbar_wait_guarded(bar, ph, g.error_flag, 4);
            ph ^= 1;
        }
#endif
        __syncthreads();  // staged rows landed; the previous tile's store phase is done with the tile
        // rebuild the packed Hermitian line Z = A + i*B for each pair of rows (lanes along k)
        for (int line = warp; line < W; line += NWARPS) {
            const float2* pa = rawc + line * LSC;
            const float2* pb = pa + W * LSC;
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
        __syncthreads();  // staging buffer consumed: the next tile's spectrum rows stream in during the FFT and the stores
        const int tn = t + gridDim.x;
        int nx = cx, ny = cy;
        advance(nx, ny);
        if (!tma) {
            if (tn < ntiles) issue_spec(tn, nx, ny);
            cp_async_commit();
        }
#ifndef PVD_EMULATE
        else if (threadIdx.x == 0) {
            if (tn < ntiles) issue_tma(tn, nx, ny);
        }
        if (has_den && tma_den) {
            if (threadIdx.x == 32) tma_prefetch_2d(&g.tmap_den, 0, t * 32);  // this tile's 32 density rows -> L2
        } else
#endif
        if (has_den) {
            // the density rows of THIS tile are needed after the inverse transform (~10 us from now): pull their
            // 128-byte lines into L2 now so that the store phase does not wait a full DRAM round trip per row
            const int r0 = t * 32;
            const int px0 = cx, py0 = cy;
            const int lines_per_row = (O2 * 4 + 127) / 128;
            for (int i = threadIdx.x; i < 32 * lines_per_row; i += NT) {
                const int rr = i / lines_per_row, ln = i - rr * lines_per_row;
                if (r0 + rr < (int)nrows) {
                    if (dense) {
                        prefetch_l2(g.density + (long long)(r0 + rr) * g.den_s1 + ln * 32);
                    } else {
                        int x, y;
                        row_xy(px0, py0, rr, x, y);
                        prefetch_l2(g.density + x * g.den_s0 + y * g.den_s1 + ln * 32);
                    }
                }
            }
        }
        auto sm_in = [&](int, int, int idx, int w) -> float2 { return tile[idx * LS + w]; };
        float* rowA = rowbuf + wl * LSA - z_lo;
        auto row_out = [&](int, int, int idx, int, float2 v) {
            if ((unsigned)(idx - z_lo) < (unsigned)O2) {
                rowA[idx] = v.x;
                rowA[idx + W * LSA] = v.y;
            }
        };
        fast_fft<N, W, LS, NT, +1, R1, R2, R3, true, true>(sm_in, row_out, tile, tws);
        __syncthreads();  // all rows complete
        const int row0 = t * 32;
        const int x0 = cx, y0 = cy;
        for (int rr = warp; rr < 32; rr += NWARPS) {
            if (row0 + rr >= (int)nrows) continue;
            float* __restrict__ dst;
            const float* __restrict__ dg;
            if (dense) {
                dst = g.out + (long long)(row0 + rr) * g.out_s1;
                dg = has_den ? g.density + (long long)(row0 + rr) * g.den_s1 : nullptr;
            } else {
                int x, y;
                row_xy(x0, y0, rr, x, y);
                dst = g.out + x * g.out_s0 + y * g.out_s1;
                dg = has_den ? g.density + x * g.den_s0 + y * g.den_s1 : nullptr;
            }
            const float* srow = rowbuf + rr * LSA;
            if (vec4 && has_den && plain_den) {
                // the common call (scale * rho_ref folded into the input weights by the host, no density cut-off):
                // max, reciprocal, multiply - half the instructions of the general form below
                constexpr int QIT = ((N + 3) / 4 + 31) / 32;
                float4 rho[QIT];
                PVD_UNROLL
                for (int i = 0; i < QIT; ++i) {
                    const int z = 4 * (lane + 32 * i);
                    rho[i] = (z < O2) ? ldg128_ro(dg + z) : make_float4(1.f, 1.f, 1.f, 1.f);
                }
                __syncwarp();  // keep the row's loads in flight together (see below)
                PVD_UNROLL
                for (int i = 0; i < QIT; ++i) {
                    const int z = 4 * (lane + 32 * i);
                    float4 v = *reinterpret_cast<const float4*>(srow + (z < O2 ? z : 0));
                    v.x *= fast_rcp(fmaxf(rho[i].x, rho_min));
                    v.y *= fast_rcp(fmaxf(rho[i].y, rho_min));
                    v.z *= fast_rcp(fmaxf(rho[i].z, rho_min));
                    v.w *= fast_rcp(fmaxf(rho[i].w, rho_min));
                    if (z < O2) stg128(dst + z, v);
                }
            } else if (vec4 && has_den) {
                constexpr int QIT = ((N + 3) / 4 + 31) / 32;
                float4 rho[QIT];
                PVD_UNROLL
                for (int i = 0; i < QIT; ++i) {
                    const int z = 4 * (lane + 32 * i);
                    rho[i] = (z < O2) ? ldg128_ro(dg + z) : make_float4(1.f, 1.f, 1.f, 1.f);
                }
                // Keep the QIT loads of the row in flight TOGETHER: ptxas otherwise sinks each load to just before its own
                // use in the register-heavy instantiations (one load in flight instead of four: P5<432> 0.42 -> 0.53 ms).
                // A warp barrier is a scheduling fence for memory operations; the lanes of the warp are convergent here.
                __syncwarp();
                PVD_UNROLL
                for (int i = 0; i < QIT; ++i) {
                    const int z = 4 * (lane + 32 * i);
                    float4 v = *reinterpret_cast<const float4*>(srow + (z < O2 ? z : 0));
                    v.x = den_apply(v.x, rho[i].x, sr, rho_min, rho_cut);
                    v.y = den_apply(v.y, rho[i].y, sr, rho_min, rho_cut);
                    v.z = den_apply(v.z, rho[i].z, sr, rho_min, rho_cut);
                    v.w = den_apply(v.w, rho[i].w, sr, rho_min, rho_cut);
                    if (z < O2) stg128(dst + z, v);
                }
            } else if (vec4) {
                constexpr int QIT = ((N + 3) / 4 + 31) / 32;
                PVD_UNROLL
                for (int i = 0; i < QIT; ++i) {
                    const int z = 4 * (lane + 32 * i);
                    if (z < O2) {
                        float4 v = *reinterpret_cast<const float4*>(srow + z);
                        v.x *= sr;
                        v.y *= sr;
                        v.z *= sr;
                        v.w *= sr;
                        stg128(dst + z, v);
                    }
                }
            } else {
                constexpr int ZIT = (N + 31) / 32;
                PVD_UNROLL
                for (int i = 0; i < ZIT; ++i) {
                    const int z = lane + 32 * i;
                    if (z < O2) dst[z] = has_den ? den_apply(srow[z], __ldg(dg + z), sr, rho_min, rho_cut) : srow[z] * sr;
                }
            }
        }
        cx = nx;
        cy = ny;
    }
    cp_async_wait<0>();
}

}  // namespace pvd
