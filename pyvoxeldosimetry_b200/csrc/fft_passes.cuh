// The five HBM passes of the 3-D real FFT convolution (generic-size engine).
//
//   P1 rows_fwd : real rows along the contiguous axis (axis 2) -> half spectrum, two real rows
//                 packed into one complex line; fuses  sum_t w_t * activity_t  and zero padding.
//   P2 cols     : axis-1 forward transform on tiles of W contiguous frequencies.
//   P3 cols     : axis-0 forward * kernel spectrum * axis-0 inverse in one kernel.
//   P4 cols     : axis-1 inverse.
//   P5 rows_inv : half spectrum -> real rows; fuses crop / centre shift, scale and the
//                 voxel-wise density correction  D * rho_ref / max(rho, rho_min).
//
// Volume layout follows the reference's NumPy C order arr[x, y, z] (z contiguous), complex
// work buffer is [M0][M1][Sz] float2 with Sz = round_up(M2/2+1, 16).
#pragma once
#include "direct_conv.cuh"  // CUtensorMap, mbarrier / TMA helpers
#include "fft_core.cuh"

namespace pvd {

struct RowFwdArgs {
    // pipelined kernel, TMA staging (use_tma): 2-D map of the dense activity volume, dims (n2 floats, rows), box (N/4, 32)
    alignas(64) CUtensorMap tmap;
    int use_tma;
    int* error_flag;
    const float* in[kMaxT];
    float w[kMaxT];
    int T;
    long long in_s0, in_s1;  // element strides of axis 0 / axis 1 of the real input (axis 2 contiguous)
    int n0, n1, n2;          // extents that are actually read (rest of the M2 line is zero)
    float2* out;
    long long out_s0, out_s1;  // float2 strides of the work buffer
    int M2, Nh;
    int Llog;  // 2^Llog complex lines (= 2^(Llog+1) real rows) per block
    int dense; // pipelined kernel: rows of input and work buffer are equally spaced in the linear row index (s0 == n1 * s1)
    int dense_in;  // the same for the input alone (enough for the TMA staging)
    const float2* tw;
    Stages st;
};

__global__ void __launch_bounds__(PVD_BLOCK) rows_fwd_kernel(const RowFwdArgs g) {
    PVD_DYN_SMEM(float2, smem);
    const int L = 1 << g.Llog, LS = L + 1, M = g.M2;
    float2* A = smem;
    float2* B = smem + (size_t)M * LS;
    const long long nrows = (long long)g.n0 * g.n1;
    const long long row0 = (long long)blockIdx.x * (2 * L);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* Af = reinterpret_cast<float*>(A);
    for (int rr = warp; rr < 2 * L; rr += nwarps) {
        const long long R = row0 + rr;
        const bool valid = R < nrows;
        long long off = 0;
        if (valid) {
            const long long x = R / g.n1, y = R - x * g.n1;
            off = x * g.in_s0 + y * g.in_s1;
        }
        for (int z = lane; z < M; z += 32) {
            float v = 0.f;
            if (valid && z < g.n2) {
                for (int t = 0; t < g.T; ++t) v += g.w[t] * __ldg(g.in[t] + off + z);
            }
            Af[((size_t)z * LS + (rr >> 1)) * 2 + (rr & 1)] = v;
        }
    }
    __syncthreads();
    const float2* Z = smem_fft<-1>(A, B, g.tw, M, g.Llog, LS, g.st);
    // split the packed transform: A[k] = (Z[k] + conj(Z[M-k]))/2, B[k] = (Z[k] - conj(Z[M-k]))/(2i)
    for (int rr = warp; rr < 2 * L; rr += nwarps) {
        const long long R = row0 + rr;
        if (R >= nrows) continue;
        const long long x = R / g.n1, y = R - x * g.n1;
        float2* dst = g.out + x * g.out_s0 + y * g.out_s1;
        const int line = rr >> 1;
        for (int k = lane; k < g.Nh; k += 32) {
            const float2 zk = Z[(size_t)k * LS + line];
            const int mk = (k == 0) ? 0 : M - k;
            const float2 zm = Z[(size_t)mk * LS + line];
            float2 o;
            if ((rr & 1) == 0)
                o = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            else
                o = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
            dst[k] = o;
        }
    }
}

enum ColMode { COL_FWD = 0, COL_INV = 1, COL_CONV = 2, COL_SPEC = 3 };

struct ColArgs {
    const float2* in;
    float2* out;
    const float2* spec;
    long long es;  // stride between consecutive transform indices (float2)
    long long os;  // stride of the outer (blockIdx.y) index
    int outer0;    // first outer index
    int n_in;      // rows loaded (the rest of the M-line is zero)
    int M;
    int out_lo, out_n;  // rows stored
    int nzf;            // valid frequencies along the contiguous axis (Nh)
    int Wlog;           // tile width 2^Wlog frequencies
    int mode;
    float scale;
    int loop_ntz, loop_ntiles;  // cols_fast_kernel as a persistent 1-D grid: tiles along the frequency axis / in total (0: 2-D grid, one tile per CTA)
    const float2* tw;
    Stages st;
};

__global__ void __launch_bounds__(PVD_BLOCK) cols_kernel(const ColArgs g) {
    PVD_DYN_SMEM(float2, smem);
    const int W = 1 << g.Wlog, M = g.M;
    float2* A = smem;
    float2* B = smem + (size_t)M * W;
    const int z0 = blockIdx.x * W;
    const long long base = (long long)(g.outer0 + (int)blockIdx.y) * g.os + z0;
    const int total = M << g.Wlog;
    const int zlim = g.nzf - z0;  // columns w < zlim are valid
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int w = i & (W - 1), r = i >> g.Wlog;
        float2 v = make_float2(0.f, 0.f);
        if (r < g.n_in && w < zlim) v = g.in[base + (long long)r * g.es + w];
        A[i] = v;
    }
    __syncthreads();
    float2* Z;
    if (g.mode == COL_INV) {
        Z = smem_fft<+1>(A, B, g.tw, M, g.Wlog, W, g.st);
    } else {
        Z = smem_fft<-1>(A, B, g.tw, M, g.Wlog, W, g.st);
        if (g.mode == COL_CONV) {
            for (int i = threadIdx.x; i < total; i += blockDim.x) {
                const int w = i & (W - 1), r = i >> g.Wlog;
                if (w < zlim) Z[i] = cmul(Z[i], __ldg(&g.spec[base + (long long)r * g.es + w]));
            }
            __syncthreads();
            float2* other = (Z == A) ? B : A;
            Z = smem_fft<+1>(Z, other, g.tw, M, g.Wlog, W, g.st);
        }
    }
    const float sc = g.scale;
    const int lo = g.out_lo, hi = g.out_lo + g.out_n;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int w = i & (W - 1), r = i >> g.Wlog;
        if (r >= lo && r < hi && w < zlim) {
            float2 v = Z[i];
            if (g.mode == COL_SPEC) v = make_float2(v.x * sc, v.y * sc);
            g.out[base + (long long)r * g.es + w] = v;
        }
    }
}

struct RowInvArgs {
    // pipelined kernel, TMA staging (use_tma): 2-D map of the work buffer as 8-byte elements, dims (Sz, rows), box
    // (LSC, 32); tmap_den (use_tma_den): 2-D map of the density volume as 8-byte elements for bulk L2 prefetches
    alignas(64) CUtensorMap tmap;
    alignas(64) CUtensorMap tmap_den;
    int use_tma, use_tma_den;
    int tma3d;  // tmap is the 3-D view (Sz, M1, M0) of the work buffer: a tile is the box at (0, y + y_lo, x + x_lo) - cropped outputs, O1 % 32 == 0
    int* error_flag;
    const float2* in;
    long long in_s0, in_s1;
    int x_lo, y_lo, z_lo;
    int O0, O1, O2;
    float* out;
    long long out_s0, out_s1;
    const float* density;  // may be null
    long long den_s0, den_s1;
    float rho_ref, rho_min, rho_cut, scale;
    int M2, Nh;
    int Llog;
    int vec4;  // pipelined kernel: dose/density rows are 16-byte aligned and O2 % 4 == 0 -> 128-bit store phase
    int dense;   // pipelined kernel: no crop offset along axes 0/1 and every row stride pair satisfies s0 == O1 * s1
    int plain_den;  // pipelined kernel: scale * rho_ref == 1 and rho_cut <= 0 -> dose = v / max(rho, rho_min), 3 instructions per voxel
    const float2* tw;
    Stages st;
};

__global__ void __launch_bounds__(PVD_BLOCK) rows_inv_kernel(const RowInvArgs g) {
    PVD_DYN_SMEM(float2, smem);
    const int L = 1 << g.Llog, LS = L + 1, M = g.M2;
    float2* A = smem;
    float2* B = smem + (size_t)M * LS;
    const long long nrows = (long long)g.O0 * g.O1;
    const long long row0 = (long long)blockIdx.x * (2 * L);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // rebuild the packed Hermitian line Z = A + i*B for each pair of rows
    for (int line = warp; line < L; line += nwarps) {
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
        for (int k = lane; k < g.Nh; k += 32) {
            float2 a = va ? pa[k] : make_float2(0.f, 0.f);
            float2 b = vb ? pb[k] : make_float2(0.f, 0.f);
            const int mk = M - k;
            const bool self = (k == 0) || (mk == k);  // DC / Nyquist: purely real bins (C2R convention)
            if (self) {
                a.y = 0.f;
                b.y = 0.f;
            }
            A[(size_t)k * LS + line] = make_float2(a.x - b.y, a.y + b.x);
            if (!self) A[(size_t)mk * LS + line] = make_float2(a.x + b.y, b.x - a.y);
        }
    }
    __syncthreads();
    const float2* Z = smem_fft<+1>(A, B, g.tw, M, g.Llog, LS, g.st);
    const float* Zf = reinterpret_cast<const float*>(Z);
    for (int rr = warp; rr < 2 * L; rr += nwarps) {
        const long long R = row0 + rr;
        if (R >= nrows) continue;
        const long long x = R / g.O1, y = R - x * g.O1;
        float* dst = g.out + x * g.out_s0 + y * g.out_s1;
        const float* den = g.density ? g.density + x * g.den_s0 + y * g.den_s1 : nullptr;
        for (int z = lane; z < g.O2; z += 32) {
            float v = Zf[((size_t)(z + g.z_lo) * LS + (rr >> 1)) * 2 + (rr & 1)] * g.scale;
            if (den) {
                const float rho = __ldg(den + z);
                v = (rho < g.rho_cut) ? 0.f : v * (g.rho_ref / fmaxf(rho, g.rho_min));
            }
            dst[z] = v;
        }
    }
}

// tw[i] = exp(-2*pi*i * i / n), evaluated in double precision
__global__ void twiddle_kernel(float2* tw, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double s, c;
        sincospi(2.0 * (double)i / (double)n, &s, &c);
        tw[i] = make_float2((float)c, (float)-s);
    }
}

// flag[0] |= 1 if any element of x[0..n) is non-finite
__global__ void finite_check_kernel(const float* x, long long n, int* flag) {
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        if (!(fabsf(v) <= 3.402823466e38f)) bad = 1;
    }
    if (bad) *flag = 1;
}

}  // namespace pvd
