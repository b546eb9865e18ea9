// Bandwidth-bound elementwise kernels of the dose path: dose-voxel-kernel evaluation on the image
// grid (A5/A6/A10), HU -> density (A9), time-weighted accumulation (A3), closed-form mono-exponential
// integral (A11) and the standalone density scale.
#pragma once
#include "pvd_common.cuh"

namespace pvd {

// Radial dose-point-kernel model shared by the Y90 and Lu177 generators:
//   k(r) = scaling * [ sum_b amp_b (1 - r/R_b)^2 exp(-2 r / R_b) [r <= R_b]
//                      + sum_p amp_p exp(-mu_p r / 10) / (4 pi r^2) [r > 0] ]
struct RadialModel {
    int nb, np;
    double beta_range[4], beta_amp[4];
    double phot_mu[4], phot_amp[4];
    double scaling;
    double sp[3];
    int g[3];
};

__global__ void kernel_eval_kernel(const RadialModel m, float* out) {
    const long long n = (long long)m.g[0] * m.g[1] * m.g[2];
    const int c0 = m.g[0] / 2, c1 = m.g[1] / 2, c2 = m.g[2] / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % m.g[2]);
        const long long t = i / m.g[2];
        const int y = (int)(t % m.g[1]);
        const int x = (int)(t / m.g[1]);
        const double dx = (x - c0) * m.sp[0], dy = (y - c1) * m.sp[1], dz = (z - c2) * m.sp[2];
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        double v = 0.0;
        for (int b = 0; b < m.nb; ++b) {
            if (r <= m.beta_range[b]) {
                const double u = 1.0 - r / m.beta_range[b];
                v += m.beta_amp[b] * (u * u * exp(-2.0 * r / m.beta_range[b]));
            }
        }
        if (r > 0.0) {
            for (int p = 0; p < m.np; ++p)
                v += m.phot_amp[p] * exp(-m.phot_mu[p] * r / 10.0) / (4.0 * 3.141592653589793 * (r * r));
        }
        out[i] = (float)(v * m.scaling);
    }
}

struct Knots {
    int nk;
    float hu[32], rho[32];
};

template <class T>
__global__ void hu_to_density_kernel(const T* __restrict__ hu, const Knots k, float* __restrict__ rho, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float h = (float)hu[i];
        float r;
        if (h <= k.hu[0]) {
            r = k.rho[0];
        } else if (h >= k.hu[k.nk - 1]) {
            r = k.rho[k.nk - 1];
        } else {
            int j = 1;
            while (h > k.hu[j]) ++j;
            const float t = (h - k.hu[j - 1]) / (k.hu[j] - k.hu[j - 1]);
            r = k.rho[j - 1] + t * (k.rho[j] - k.rho[j - 1]);
        }
        rho[i] = r;
    }
}

// 16-bit stored activity (PET DICOM pixel data) -> float32: out = slope * stored + intercept (io/dicom.py:27-47)
template <class T>
__global__ void rescale_to_f32_kernel(const T* __restrict__ in, float slope, float intercept, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = fmaf(slope, (float)in[i], intercept);
}

// J weighted combinations of T volumes in one pass: out_j = sum_t w[j][t] * v_t.  Every input is read once and every output
// written once: 4 (T + J) bytes per voxel.  This is ActivitySampler._trapezoid_integration (J = 1, core/activity_sampler.py:
// 69-79), the trapezoid of precomputed dose rates, and interpolate_timepoints (core/utils.py:154-191) - scipy's interp1d of
// kind linear / cubic / previous is a LINEAR map of the sampled volumes whose T x J weights depend on the time points only,
// so the host computes the weights and the volumes make one trip through HBM.  A zero weight skips its product (an unused
// volume may hold anything, as with interp1d), a NaN weight makes the output NaN ('previous' before the first time point).
constexpr int kMaxJ = 16;
struct WcombArgs {
    const float* v[kMaxT];
    float* o[kMaxJ];
    float w[kMaxJ][kMaxT];
    unsigned used;  // bit t: some output has a non-zero weight on volume t (others are never loaded)
    int T, J;
};

template <int JJ>  // compile-time bound of the outputs (accumulators stay in registers); a.J <= JJ of them are stored
__global__ void __launch_bounds__(256) weighted_combine_kernel(const WcombArgs a, size_t n, int vec) {
    const size_t ngroups = (n + 3) / 4;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (size_t)gridDim.x * blockDim.x) {
        const size_t i0 = 4 * g;
        const bool full = vec && i0 + 4 <= n;
        float4 acc[JJ];
        PVD_UNROLL
        for (int j = 0; j < JJ; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        // four volumes per trip: their loads are issued together (memory-level parallelism), then the products
        for (int t0 = 0; t0 < a.T; t0 += 4) {
            float4 x[4];
            PVD_UNROLL
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + q;
                x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < a.T && ((a.used >> t) & 1u)) {
                    const float* __restrict__ src = a.v[t] + i0;
                    if (full) {
                        x[q] = *reinterpret_cast<const float4*>(src);
                    } else {
                        x[q].x = src[0];
                        if (i0 + 1 < n) x[q].y = src[1];
                        if (i0 + 2 < n) x[q].z = src[2];
                        if (i0 + 3 < n) x[q].w = src[3];
                    }
                }
            }
            PVD_UNROLL
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + q;
                if (t < a.T) {
                    PVD_UNROLL
                    for (int j = 0; j < JJ; ++j) {
                        const float w = a.w[j][t];
                        if (w != 0.f) {  // NaN != 0: a NaN weight propagates
                            acc[j].x = fmaf(w, x[q].x, acc[j].x);
                            acc[j].y = fmaf(w, x[q].y, acc[j].y);
                            acc[j].z = fmaf(w, x[q].z, acc[j].z);
                            acc[j].w = fmaf(w, x[q].w, acc[j].w);
                        }
                    }
                }
            }
        }
        PVD_UNROLL
        for (int j = 0; j < JJ; ++j) {
            if (j < a.J) {
                float* __restrict__ dst = a.o[j] + i0;
                if (full) {
                    *reinterpret_cast<float4*>(dst) = acc[j];
                } else {
                    dst[0] = acc[j].x;
                    if (i0 + 1 < n) dst[1] = acc[j].y;
                    if (i0 + 2 < n) dst[2] = acc[j].z;
                    if (i0 + 3 < n) dst[3] = acc[j].w;
                }
            }
        }
    }
}

__global__ void monoexp_integral_kernel(const float* __restrict__ A0, const float* __restrict__ lam, float tlim,
                                        float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float l = lam[i];
        // -expm1(-x) keeps full relative precision for small lambda*T
        out[i] = A0[i] / l * (-expm1f(-l * tlim));
    }
}

__global__ void density_scale_kernel(const float* __restrict__ dose, const float* __restrict__ den, float rho_ref,
                                     float rho_min, float rho_cut, float scale, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float rho = den[i];
        const float v = dose[i] * scale;
        out[i] = (rho < rho_cut) ? 0.f : v * (rho_ref / fmaxf(rho, rho_min));
    }
}

}  // namespace pvd
