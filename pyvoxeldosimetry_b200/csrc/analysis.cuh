// Kernels for the steps either side of the convolution:
//   * per-voxel mono-exponential time-activity fit + closed-form integral (feeds the convolution),
//   * CT preparation: metal-artifact fill, HU -> density and tissue-class labels (feeds the density correction),
//   * dose-volume histogram of a dose map over an ROI mask (consumes the dose map on the device).
// All are single-pass, bandwidth-bound elementwise / reduction kernels.
#pragma once
#include "elementwise.cuh"

#ifdef PVD_EMULATE
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline unsigned atomicMin(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
using std::isfinite;
#endif

namespace pvd {

// ------------------------------------------------------------------------------------------------
// Mono-exponential fit  y(t) = A0 exp(-lambda t), weighted least squares  sum_i (w_i (y_i - f(t_i)))^2
// (reference: time_integration/curve_fitting.py:46-59 - one scipy.optimize.curve_fit call per voxel with
// p0 = [y(t_0), ln2/half_life], sigma = 1/weight_factors).  Per voxel, in registers:
//   1. up to 16 damped Gauss-Newton (Levenberg-Marquardt) steps on (A0, lambda) from the reference's p0 - the
//      globalisation: a step is accepted only if the cost does not increase; stops when the cost stalls;
//   2. up to 4 Newton steps on the variable-projection stationarity condition h(lambda) = P Q1 - Q P1 = 0
//      (P = sum w^2 y e, Q = sum w^2 e^2, P1 = sum w^2 y t e, Q1 = sum w^2 t e^2, e = exp(-lambda t)), which
//      polishes lambda to float32 round-off; A0 = P/Q is the exact linear optimum for that lambda.
// A voxel whose result is not finite or whose Jacobian is rank-deficient at the end point (no finite minimiser)
// gets [0, lambda0], as the reference does when curve_fit raises (:58-59).
// The integral A0/lambda (1 - exp(-lambda T_lim)) (curve_fitting.py:74-84) is fused into the same pass.
struct FitArgs {
    const float* v[kMaxT];
    float t[kMaxT];
    float w[kMaxT];
    int T;
    float lam0, tlim;
};

template <int TT>  // TT = compile-time number of time points (0: runtime, up to kMaxT)
__global__ void monoexp_fit_kernel(const FitArgs a, float* __restrict__ A0out, float* __restrict__ lamout,
                                   float* __restrict__ accout, size_t n) {
    constexpr int TM = TT > 0 ? TT : kMaxT;
    const int T = TT > 0 ? TT : a.T;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float y[TM];
        PVD_UNROLL
        for (int k = 0; k < TM; ++k) y[k] = (k < T) ? __ldg(a.v[k] + i) : 0.f;
        // e[k] = exp(-lambda t_k) at the current iterate is carried along: one set of exponentials per step
        float e[TM];
        float A = y[0], lam = a.lam0, mu = 1e-3f, c = 0.f;
        PVD_UNROLL
        for (int k = 0; k < TM; ++k)
            if (k < T) {
                e[k] = expf(-lam * a.t[k]);
                const float r = a.w[k] * (y[k] - A * e[k]);
                c += r * r;
            }
        for (int it = 0; it < 16; ++it) {
            float a11 = 0.f, a12 = 0.f, a22 = 0.f, g1 = 0.f, g2 = 0.f;
            PVD_UNROLL
            for (int k = 0; k < TM; ++k)
                if (k < T) {
                    const float ja = a.w[k] * e[k], jl = -a.w[k] * A * a.t[k] * e[k];
                    const float r = a.w[k] * (y[k] - A * e[k]);
                    a11 += ja * ja;
                    a12 += ja * jl;
                    a22 += jl * jl;
                    g1 += ja * r;
                    g2 += jl * r;
                }
            const float d11 = a11 * (1.f + mu), d22 = a22 * (1.f + mu) + (a22 == 0.f ? 1.f : 0.f);
            float det = d11 * d22 - a12 * a12;
            if (det == 0.f) det = 1.f;
            const float An = A + (d22 * g1 - a12 * g2) / det, ln = lam + (d11 * g2 - a12 * g1) / det;
            float en[TM], cn = 0.f;
            PVD_UNROLL
            for (int k = 0; k < TM; ++k)
                if (k < T) {
                    en[k] = expf(-ln * a.t[k]);
                    const float r = a.w[k] * (y[k] - An * en[k]);
                    cn += r * r;
                }
            const bool ok = cn <= c;  // false for NaN
            const float gain = c - cn, cold = c;
            if (ok) {
                A = An;
                lam = ln;
                c = cn;
                PVD_UNROLL
                for (int k = 0; k < TM; ++k) e[k] = en[k];
            }
            mu = fminf(fmaxf(ok ? mu * 0.2f : mu * 10.f, 1e-9f), 1e9f);
            // stop when the step no longer changes the cost or the parameters beyond float32 noise (the Newton polish
            // below finishes the digits), or when the damping has run away without finding a descent step
            const bool small_step = fabsf(An - A) <= 1e-5f * fabsf(A) && fabsf(ln - lam) <= 1e-5f * fabsf(lam);
            if (fabsf(gain) <= 1e-5f * cold || (!ok && (small_step || mu >= 1e9f))) break;
        }
        float P = 0.f, Q = 0.f, h = 0.f, dh = 0.f;
        auto hfun = [&](float l, float& P_, float& Q_, float& h_, float& dh_) {
            float p = 0.f, q = 0.f, p1 = 0.f, q1 = 0.f, p2 = 0.f, q2 = 0.f;
            PVD_UNROLL
            for (int k = 0; k < TM; ++k)
                if (k < T) {
                    const float e = expf(-l * a.t[k]), w2 = a.w[k] * a.w[k], t = a.t[k];
                    const float wye = w2 * y[k] * e, wee = w2 * e * e;
                    p += wye;
                    q += wee;
                    p1 += wye * t;
                    q1 += wee * t;
                    p2 += wye * t * t;
                    q2 += wee * t * t;
                }
            P_ = p;
            Q_ = q;
            h_ = p * q1 - q * p1;
            dh_ = -p1 * q1 - 2.f * p * q2 + 2.f * q1 * p1 + q * p2;
        };
        hfun(lam, P, Q, h, dh);
        for (int it = 0; it < 4; ++it) {
            const float step = dh != 0.f ? h / dh : 0.f;
            if (fabsf(step) <= 1e-7f * fabsf(lam)) break;
            const float ln = lam - step;
            float Pn, Qn, hn, dhn;
            hfun(ln, Pn, Qn, hn, dhn);
            if (!(fabsf(hn) < fabsf(h))) break;
            lam = ln;
            P = Pn;
            Q = Qn;
            h = hn;
            dh = dhn;
        }
        A = Q > 0.f ? P / Q : 0.f;
        // Ill-posed curves (e.g. activity that drops to zero after the first point) have no finite minimiser:
        // lambda runs away, curve_fit raises and the reference stores [0, lambda0].  Here that shows as a
        // rank-deficient Jacobian at the end point: det(J^T J) <= 1e-5 a11 a22 (well-posed fits sit at 0.3-0.7).
        {
            float a11 = 0.f, a12 = 0.f, a22 = 0.f;
            PVD_UNROLL
            for (int k = 0; k < TM; ++k)
                if (k < T) {
                    const float e = expf(-lam * a.t[k]);
                    const float ja = a.w[k] * e, jl = a.w[k] * a.t[k] * e;  // d/dlambda without the common factor -A
                    a11 += ja * ja;
                    a12 += ja * jl;
                    a22 += jl * jl;
                }
            const bool rank_ok = (a11 * a22 - a12 * a12) > 1e-5f * (a11 * a22);
            if (!(isfinite(A) && isfinite(lam)) || (A != 0.f && !rank_ok)) {
                A = 0.f;
                lam = a.lam0;
            }
        }
        if (A0out) A0out[i] = A;
        if (lamout) lamout[i] = lam;
        if (accout) accout[i] = A / lam * (-expm1f(-lam * a.tlim));
    }
}

// ------------------------------------------------------------------------------------------------
// CT preparation.  One pass over the CT volume (float32 HU):
//   * metal-artifact fill (reference tissue/composition.py:73-93): voxels above the threshold are replaced
//     by the value at that voxel of gaussian_filter(ct with those voxels zeroed, sigma=1) - scipy defaults:
//     radius 4 (truncate 4.0), 'reflect' boundary.  Only metal voxels need the 9^3 gather, so it is
//     evaluated at those voxels only (weights g(dx) g(dy) g(dz), identical to the separable passes);
//   * HU -> mass density by the piecewise-linear knot table (A9 helper);
//   * tissue-class bit mask, bit c set when lo_c <= HU <= hi_c (composition.py:63-67, ranges :40-46).
struct CtArgs {
    const float* hu;
    int n0, n1, n2;
    float metal_thr;  // +inf disables the fill
    float g[5];       // normalised Gaussian weights for |d| = 0..4
    // density: rho(h) = rho0 + sum_j slope_j * clamp(h - hu_j, 0, len_j) - the piecewise-linear knot table without
    // a data-dependent search (every lane reads the same constants; random HU would otherwise serialise the
    // constant-bank lookups lane by lane)
    int nseg;
    float rho0, seg_hu[31], seg_len[31], seg_slope[31];
    int nr;
    float lo[8], hi[8];
    // Merged interval table (built on the host, pvd_ct_prepare): the density knots, every range start lo_c and the float
    // after every range end hi_c, sorted, cut the HU axis into intervals on which the density is ONE linear piece and the
    // class mask is constant.  thr[] = the sorted cuts padded with +inf to 2^lut_log - 1 entries; entry k (k = number of
    // cuts <= h, found by a branch-free binary search in shared memory: the tables are a few dozen words, so 32 different
    // HU values read different banks) = (a, rho(a), slope, mask bits): rho(h) = rho(a) + slope (h - a).
    int lut_log;  // 0: table too large, use the sums above
    float knot_lo, knot_hi;  // first / last density knot
    float thr[63];
    float ent[64][4];
    float* corrected;
    float* rho;
    unsigned char* labels;
};

__device__ __forceinline__ int reflect_idx(int i, int n) {  // scipy 'reflect': (d c b a | a b c d | d c b a)
    while (i < 0 || i >= n) i = (i < 0) ? (-i - 1) : (2 * n - 1 - i);
    return i;
}

constexpr int kCtThreads = 256;

__device__ __forceinline__ float4 ct_ld128(const float* p) {
#ifdef PVD_EMULATE
    return *reinterpret_cast<const float4*>(p);
#else
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}

// Streaming pass, four consecutive voxels per thread (128-bit loads / stores when the volume is 16-byte aligned), no
// block-wide synchronisation.  Metal voxels are rare, so their 9^3 gathers are not done by the owning lane alone
// (it would run 729 taps while 31 lanes idle): the warp votes, and for every metal voxel found all 32 lanes share
// the work - the 81 (dx, dy) columns of 9 contiguous z taps are dealt round-robin over the lanes and the partial
// sums are combined with a butterfly of shuffles (fixed order: deterministic result).
template <int LOG>  // LOG > 0: merged interval table of 2^LOG entries in shared memory; 0: direct sums over segments and ranges
__global__ void __launch_bounds__(kCtThreads) ct_prepare_kernel(const CtArgs a, const int vec) {
    __shared__ float s_thr[LOG > 0 ? (1 << LOG) : 1];
    __shared__ float4 s_ent[LOG > 0 ? (1 << LOG) : 1];
    if constexpr (LOG > 0) {
        for (int i = threadIdx.x; i < (1 << LOG); i += blockDim.x) {
            s_thr[i] = i < (1 << LOG) - 1 ? a.thr[i] : 0.f;
            s_ent[i] = make_float4(a.ent[i][0], a.ent[i][1], a.ent[i][2], a.ent[i][3]);
        }
        __syncthreads();
    }
    const size_t n = (size_t)a.n0 * a.n1 * a.n2;
    const size_t ngroups = (n + 3) / 4;
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count: every lane of a warp runs the same number of iterations (the votes need all lanes)
    const size_t gstride = (size_t)gridDim.x * blockDim.x;
    const size_t warp_first = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    // Software pipeline: the 128-bit load of the NEXT trip's group is issued before this trip's ~260 instructions of table
    // arithmetic, so every warp always has a load in flight (one load per thread and trip left the pass at 38 % of its
    // issue slots with 16 KB per SM in flight).  The loads are volatile asm so the compiler keeps them where they are.
    auto fetch = [&](size_t gb, float (&hv)[4]) {
        hv[0] = hv[1] = hv[2] = hv[3] = 0.f;
        if (gb >= ngroups) return;
        const size_t j0 = 4 * (gb + lane);
        if (vec && j0 + 4 <= n) {
            const float4 v = ct_ld128(a.hu + j0);
            hv[0] = v.x;
            hv[1] = v.y;
            hv[2] = v.z;
            hv[3] = v.w;
        } else {
            PVD_UNROLL
            for (int q = 0; q < 4; ++q)
                if (j0 + q < n) hv[q] = a.hu[j0 + q];
        }
    };
    float hn[4];
    fetch(warp_first, hn);
    for (size_t gbase = warp_first; gbase < ngroups; gbase += gstride) {
        const size_t i0 = 4 * (gbase + lane);
        const bool full = vec && i0 + 4 <= n;
        float h[4] = {hn[0], hn[1], hn[2], hn[3]};
        fetch(gbase + gstride, hn);
        // metal voxels are rare: ONE vote over the four voxels of every lane decides whether the warp looks closer
        const bool any_metal = __any_sync(0xFFFFFFFFu, fmaxf(fmaxf(h[0], h[1]), fmaxf(h[2], h[3])) > a.metal_thr);
        if (any_metal) {
        PVD_UNROLL
        for (int q = 0; q < 4; ++q) {
            unsigned metal = __ballot_sync(0xFFFFFFFFu, i0 + q < n && h[q] > a.metal_thr);
            while (metal) {
                const int owner = __ffs(metal) - 1;
                metal &= metal - 1;
                const size_t iv = 4 * (gbase + owner) + q;
                const int z = (int)(iv % a.n2);
                const size_t r = iv / a.n2;
                const int y = (int)(r % a.n1), x = (int)(r / a.n1);
                int zi[9];
                PVD_UNROLL
                for (int d = 0; d < 9; ++d) zi[d] = reflect_idx(z + d - 4, a.n2);
                float part = 0.f;
                for (int col = lane; col < 81; col += 32) {
                    const int dx = col / 9 - 4, dy = col % 9 - 4;
                    const float* __restrict__ src = a.hu + ((size_t)reflect_idx(x + dx, a.n0) * a.n1 + reflect_idx(y + dy, a.n1)) * a.n2;
                    float v[9];
                    PVD_UNROLL
                    for (int d = 0; d < 9; ++d) v[d] = __ldg(src + zi[d]);
                    float line = 0.f;
                    PVD_UNROLL
                    for (int d = 0; d < 9; ++d) line += a.g[d < 4 ? 4 - d : d - 4] * (v[d] > a.metal_thr ? 0.f : v[d]);
                    part += a.g[dx < 0 ? -dx : dx] * a.g[dy < 0 ? -dy : dy] * line;
                }
                PVD_UNROLL
                for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, s);
                if (lane == owner) h[q] = part;
            }
        }
        }
        float rr[4];
        unsigned lab[4];
        PVD_UNROLL
        for (int q = 0; q < 4; ++q) {
            if constexpr (LOG > 0) {
                int pos = 0;  // number of cuts <= h (NaN: 0 -> the entry below every cut: rho of the first knot, no class)
                PVD_UNROLL
                for (int st = 1 << (LOG - 1); st > 0; st >>= 1) pos += (s_thr[pos + st - 1] <= h[q]) ? st : 0;
                const float4 e = s_ent[pos];
                // h clamped to the knot range for the linear piece: beyond it the slope is 0 and 0 * inf would be NaN
                rr[q] = fmaf(e.z, fminf(fmaxf(h[q], a.knot_lo), a.knot_hi) - e.x, e.y);
                lab[q] = __float_as_uint(e.w);
            } else {
                float acc = a.rho0;
                for (int j = 0; j < a.nseg; ++j) acc = fmaf(a.seg_slope[j], fminf(fmaxf(h[q] - a.seg_hu[j], 0.f), a.seg_len[j]), acc);
                rr[q] = acc;
                unsigned m = 0;
                PVD_UNROLL
                for (int c = 0; c < 8; ++c) m |= (h[q] >= a.lo[c] && h[q] <= a.hi[c]) ? (1u << c) : 0u;  // unused ranges are empty (lo > hi)
                lab[q] = m;
            }
        }
        if (full) {
            if (a.corrected) *reinterpret_cast<float4*>(a.corrected + i0) = make_float4(h[0], h[1], h[2], h[3]);
            if (a.rho) *reinterpret_cast<float4*>(a.rho + i0) = make_float4(rr[0], rr[1], rr[2], rr[3]);
            if (a.labels) *reinterpret_cast<unsigned*>(a.labels + i0) = lab[0] | (lab[1] << 8) | (lab[2] << 16) | (lab[3] << 24);
        } else {
            PVD_UNROLL
            for (int q = 0; q < 4; ++q)
                if (i0 + q < n) {
                    if (a.corrected) a.corrected[i0 + q] = h[q];
                    if (a.rho) a.rho[i0 + q] = rr[q];
                    if (a.labels) a.labels[i0 + q] = (unsigned char)lab[q];
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dose-volume histogram (reference core/utils.py:233-262: np.histogram(dose[mask > 0], bins) + cumulative).
// Pass 1: min / max / count of the ROI doses.  Floats are mapped to order-preserving unsigned keys so one
// atomicMin/atomicMax per block suffices.  Pass 2: bin counts with numpy's uniform-bin rule, reproduced
// operation by operation in float32 (numpy/lib/_histograms_impl.py: index = int((x - first) / (last - first) * bins),
// clamp of the last edge, then the two corrections against the edge array, which the host computes with
// np.linspace exactly as numpy does) - so the counts are bit-identical to numpy's on the same float32 doses.
__device__ __forceinline__ unsigned f2key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct RoiStats {
    unsigned min_key, max_key;
    unsigned long long count;
};

// four mask values at once: a 32-bit word of uint8 flags or a float4
__device__ __forceinline__ void load_mask4(const unsigned char* m, size_t i4, bool in[4]) {
    const unsigned u = reinterpret_cast<const unsigned*>(m)[i4];
    in[0] = (u & 0xFFu) != 0;
    in[1] = (u & 0xFF00u) != 0;
    in[2] = (u & 0xFF0000u) != 0;
    in[3] = (u & 0xFF000000u) != 0;
}
__device__ __forceinline__ void load_mask4(const float* m, size_t i4, bool in[4]) {
    const float4 v = reinterpret_cast<const float4*>(m)[i4];
    in[0] = v.x > 0.f;
    in[1] = v.y > 0.f;
    in[2] = v.z > 0.f;
    in[3] = v.w > 0.f;
}

// n4 = number of 4-voxel groups read with 128-bit loads (0 when the arrays are not aligned); the rest is scalar.
template <class M>
__global__ void roi_minmax_kernel(const float* __restrict__ dose, const M* __restrict__ mask, size_t n, size_t n4, RoiStats* out) {
    __shared__ unsigned s_min, s_max;
    __shared__ unsigned long long s_cnt;
    if (threadIdx.x == 0) {
        s_min = 0xFFFFFFFFu;
        s_max = 0u;
        s_cnt = 0ull;
    }
    __syncthreads();
    unsigned lo = 0xFFFFFFFFu, hi = 0u;
    unsigned long long cnt = 0;
    auto take = [&](float d) {
        const unsigned k = f2key(d);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
        ++cnt;
    };
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t i = gtid;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // four independent 128-bit loads in flight per thread (memory-level parallelism)
        float4 d[4];
        bool in[4][4];
        PVD_UNROLL
        for (int u = 0; u < 4; ++u) {
            d[u] = reinterpret_cast<const float4*>(dose)[i + u * stride];
            load_mask4(mask, i + u * stride, in[u]);
        }
        PVD_UNROLL
        for (int u = 0; u < 4; ++u) {
            if (in[u][0]) take(d[u].x);
            if (in[u][1]) take(d[u].y);
            if (in[u][2]) take(d[u].z);
            if (in[u][3]) take(d[u].w);
        }
    }
    for (; i < n4; i += stride) {
        const float4 d = reinterpret_cast<const float4*>(dose)[i];
        bool in[4];
        load_mask4(mask, i, in);
        if (in[0]) take(d.x);
        if (in[1]) take(d.y);
        if (in[2]) take(d.z);
        if (in[3]) take(d.w);
    }
    for (size_t j = 4 * n4 + gtid; j < n; j += stride)
        if (mask[j] > (M)0) take(dose[j]);
    if (cnt) {
        atomicMin(&s_min, lo);
        atomicMax(&s_max, hi);
        atomicAdd(&s_cnt, cnt);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) {
        atomicMin(&out->min_key, s_min);
        atomicMax(&out->max_key, s_max);
        atomicAdd(&out->count, s_cnt);
    }
}

constexpr int kDvhSmemBins = 4096;

// The bin index numpy derives with a true division is only a first guess that the two comparisons against the edge
// array then correct by at most one bin; a multiplication by the host-computed reciprocal gives a guess within the same
// +-1, so the corrected index - the edge-defined bin - is identical (counts stay bit-identical to np.histogram) while
// the IEEE division (~10 instructions) and the two global edge loads per voxel go away: the edges sit in shared memory.
template <class M>
__global__ void __launch_bounds__(256) dvh_hist_kernel(const float* __restrict__ dose, const M* __restrict__ mask, size_t n, size_t n4,
                                                       const float* __restrict__ edges, int bins, float first, float last,
                                                       unsigned long long* __restrict__ hist) {
    __shared__ unsigned s_hist[kDvhSmemBins];
    __shared__ float s_edges[kDvhSmemBins + 1];
    const bool use_smem = bins <= kDvhSmemBins;
    if (use_smem) {
        for (int b = threadIdx.x; b < bins; b += blockDim.x) s_hist[b] = 0u;
        for (int b = threadIdx.x; b <= bins; b += blockDim.x) s_edges[b] = edges[b];
        __syncthreads();
    }
    const float* __restrict__ ed = use_smem ? s_edges : edges;
    const float inv = (float)bins / __fsub_rn(last, first);
    // How far (in bins) the float32 guess f can sit from the position the float32 EDGES imply: the guess carries a relative
    // error below 2^-22 (subtraction, reciprocal, product), an edge is off its ideal place by at most half an ulp of the
    // larger outer edge.  A voxel whose fractional position is further than that (with a 4x margin) from both ends of its
    // bin cannot be moved by the edge comparisons - it skips the two shared-memory edge loads (99.9 % of the voxels for
    // 1000 bins); the rest take the exact path, so the counts stay bit-identical to np.histogram.
    const float tol_edges = 2.5e-7f * fmaxf(fabsf(first), fabsf(last)) * inv;
    auto take = [&](float x) {
        if (!(x >= first && x <= last)) return;  // NaN doses are dropped, as numpy's `keep` mask does
        const float f = (x - first) * inv;
        int idx = (int)f;
        const float frac = f - (float)idx, tol = fmaf(1e-6f, f, tol_edges);
        if (!(frac > tol && frac < 1.f - tol) || idx >= bins) {
            idx = idx < 0 ? 0 : (idx > bins - 1 ? bins - 1 : idx);
            // one step to the edge-defined bin: the guess is within +-1 (its relative error is a few 2^-24, bins <= 2^20)
            if (x < ed[idx]) --idx;
            else if (idx != bins - 1 && x >= ed[idx + 1]) ++idx;
        }
        if (use_smem) atomicAdd(&s_hist[idx], 1u);
        else atomicAdd(&hist[idx], 1ull);
    };
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t i = gtid;
    for (; i + stride < n4; i += 2 * stride) {  // two independent 128-bit loads in flight per thread
        const float4 d0 = reinterpret_cast<const float4*>(dose)[i], d1 = reinterpret_cast<const float4*>(dose)[i + stride];
        bool in0[4], in1[4];
        load_mask4(mask, i, in0);
        load_mask4(mask, i + stride, in1);
        if (in0[0]) take(d0.x);
        if (in0[1]) take(d0.y);
        if (in0[2]) take(d0.z);
        if (in0[3]) take(d0.w);
        if (in1[0]) take(d1.x);
        if (in1[1]) take(d1.y);
        if (in1[2]) take(d1.z);
        if (in1[3]) take(d1.w);
    }
    for (; i < n4; i += stride) {
        const float4 d = reinterpret_cast<const float4*>(dose)[i];
        bool in[4];
        load_mask4(mask, i, in);
        if (in[0]) take(d.x);
        if (in[1]) take(d.y);
        if (in[2]) take(d.z);
        if (in[3]) take(d.w);
    }
    for (size_t j = 4 * n4 + gtid; j < n; j += stride)
        if (mask[j] > (M)0) take(dose[j]);
    if (use_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < bins; b += blockDim.x)
            if (s_hist[b]) atomicAdd(&hist[b], (unsigned long long)s_hist[b]);
    }
}

// Lane-private variant for <= kDvhLaneBins bins: 32 copies of the histogram, copy l entirely in shared-memory bank l
// (counter of bin b for lane l at word 32 b + l), so the 32 atomics of a warp never share a bank - with one copy, 32 random
// bins fall 3-4 deep into some bank and the shared-memory atomics were half of the kernel's time.  One 1024-thread CTA per
// SM walks the volume; the copies are summed at the end with rotated (conflict-free) reads.  Same bin rule as above.
constexpr int kDvhLaneBins = 1536;
constexpr int kDvhLaneThreads = 1024;

template <class M>
__global__ void __launch_bounds__(kDvhLaneThreads, 1) dvh_hist_lanes_kernel(const float* __restrict__ dose, const M* __restrict__ mask, size_t n,
                                                                            size_t n4, const float* __restrict__ edges, int bins, float first,
                                                                            float last, unsigned long long* __restrict__ hist) {
    PVD_DYN_SMEM(unsigned, sm);
    unsigned* s_hist = sm;                                              // [bins][32]
    float* ed = reinterpret_cast<float*>(sm + (size_t)bins * 32);       // [bins + 1]
    for (int b = threadIdx.x; b < bins * 32; b += blockDim.x) s_hist[b] = 0u;
    for (int b = threadIdx.x; b <= bins; b += blockDim.x) ed[b] = edges[b];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float inv = (float)bins / __fsub_rn(last, first);
    const float tol_edges = 2.5e-7f * fmaxf(fabsf(first), fabsf(last)) * inv;
    auto take = [&](float x) {
        if (!(x >= first && x <= last)) return;
        const float f = (x - first) * inv;
        int idx = (int)f;
        const float frac = f - (float)idx, tol = fmaf(1e-6f, f, tol_edges);
        if (!(frac > tol && frac < 1.f - tol) || idx >= bins) {
            idx = idx < 0 ? 0 : (idx > bins - 1 ? bins - 1 : idx);
            if (x < ed[idx]) --idx;
            else if (idx != bins - 1 && x >= ed[idx + 1]) ++idx;
        }
        atomicAdd(&s_hist[idx * 32 + lane], 1u);
    };
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t i = gtid;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // four independent 128-bit loads in flight per thread
        float4 d[4];
        bool in[4][4];
        PVD_UNROLL
        for (int u = 0; u < 4; ++u) {
            d[u] = reinterpret_cast<const float4*>(dose)[i + u * stride];
            load_mask4(mask, i + u * stride, in[u]);
        }
        PVD_UNROLL
        for (int u = 0; u < 4; ++u) {
            if (in[u][0]) take(d[u].x);
            if (in[u][1]) take(d[u].y);
            if (in[u][2]) take(d[u].z);
            if (in[u][3]) take(d[u].w);
        }
    }
    for (; i < n4; i += stride) {
        const float4 d = reinterpret_cast<const float4*>(dose)[i];
        bool in[4];
        load_mask4(mask, i, in);
        if (in[0]) take(d.x);
        if (in[1]) take(d.y);
        if (in[2]) take(d.z);
        if (in[3]) take(d.w);
    }
    for (size_t j = 4 * n4 + gtid; j < n; j += stride)
        if (mask[j] > (M)0) take(dose[j]);
    __syncthreads();
    for (int b = threadIdx.x; b < bins; b += blockDim.x) {
        unsigned long long c = 0;
        PVD_UNROLL
        for (int l = 0; l < 32; ++l) c += s_hist[b * 32 + ((l + lane) & 31)];
        if (c) atomicAdd(&hist[b], c);
    }
}

}  // namespace pvd
