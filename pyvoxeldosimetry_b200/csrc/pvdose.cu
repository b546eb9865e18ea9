// libpvdose: plan management and the C ABI (include/pvdose.h).
#include "../../include/pvdose.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "direct_conv.cuh"
#include "analysis.cuh"
#include "fft_fast.cuh"
#include "fft_pipe.cuh"
#include "fft_rows_pipe.cuh"
#include "fft_passes.cuh"
#include "host_stage.cuh"

using namespace pvd;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define PVD_CUDA_CHECK(what)                                                                  \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) return fail(PVD_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__)); \
    } while (0)

constexpr size_t kMaxSmem = 227 * 1024;

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- radix schedule -------------------------------------------------------------------------
bool make_stages(int M, Stages& st) {
    st.n = 0;
    if (M < 1) return false;
    std::vector<int> r;
    int m = M, e2 = 0;
    while (m % 2 == 0) {
        m /= 2;
        ++e2;
    }
    int e3 = 0, e5 = 0, e7 = 0;
    while (m % 3 == 0) { m /= 3; ++e3; }
    while (m % 5 == 0) { m /= 5; ++e5; }
    while (m % 7 == 0) { m /= 7; ++e7; }
    for (; e5 >= 2; e5 -= 2) r.push_back(25);
    if (e5) r.push_back(5);
    for (; e3 >= 2; e3 -= 2) r.push_back(9);
    if (e3) r.push_back(3);
    for (; e7 > 0; --e7) r.push_back(7);
    for (int p = 11; m > 1; p += 2) {  // remaining primes -> generic O(p^2) stage
        while (m % p == 0) {
            r.push_back(p);
            m /= p;
        }
        if ((long long)p * p > m && m > 1) {
            r.push_back(m);
            m = 1;
        }
    }
    if (e2 > 0) {
        const int nst = (e2 + 3) / 4;
        const int base = e2 / nst, rem = e2 % nst;
        for (int i = 0; i < nst; ++i) r.push_back(1 << (base + (i < rem ? 1 : 0)));
    }
    if ((int)r.size() > kMaxStages) return false;
    st.n = (int)r.size();
    for (int i = 0; i < st.n; ++i) st.radix[i] = r[i];
    return true;
}

bool is_smooth7(int m) {
    for (int p : {2, 3, 5, 7})
        while (m % p == 0) m /= p;
    return m == 1;
}

// Relative cost of an axis of length m: every pass streams the axis through HBM (weight 8) and each
// radix-r stage costs one shared-memory round trip plus ~log2(r) butterfly levels.
double size_cost(int m) {
    Stages st;
    if (!make_stages(m, st)) return 1e30;
    double c = 8.0;
    for (int i = 0; i < st.n; ++i) c += 0.3 + 0.35 * std::log2((double)st.radix[i]);
    return (double)m * c;
}

int good_size(int n) {
    if (n <= 1) return 1;
    int best = -1;
    double bc = 1e30;
    for (int m = n; m <= 2 * n; ++m) {
        if (!is_smooth7(m)) continue;
        const double c = size_cost(m);
        if (c < bc) {
            bc = c;
            best = m;
        }
        if ((double)m * 8.0 > bc) break;  // no later candidate can win
    }
    return best;
}

// ---- size-specialised kernel menu (fft_fast.cuh) ------------------------------------------------
typedef void (*ColKernelFn)(const ColArgs);
typedef void (*RowFwdKernelFn)(const RowFwdArgs);
typedef void (*RowInvKernelFn)(const RowInvArgs);
typedef void (*ColPipeKernelFn)(const ColPipeArgs);
struct FastCols {
    int N, NT;
    ColKernelFn fn[4];        // one tile per CTA, indexed by ColMode
    ColPipeKernelFn pipe[4];  // persistent cp.async-pipelined variant (null: use fn for that mode)
    int pipeNT[4];            // threads per CTA of each pipelined variant
    int fnNT[4];              // threads per CTA of each one-tile-per-CTA variant
    int pipeW;                // frequencies per tile of the pipelined variants (16 = full 128-byte lines, 8 = half lines)
};
struct FastRows {
    int N, NT;
    RowFwdKernelFn fwd;
    RowInvKernelFn inv;
    RowFwdKernelFn fwdPipe;  // persistent cp.async-staged variants (fft_rows_pipe.cuh)
    RowInvKernelFn invPipe;
    size_t smemPipe;
    int NTinv;               // threads per CTA of the inverse kernels (the two directions may use different radix schedules)
};
#define PVD_COLS_FN(N, NT, R1, R2, R3)                                                             \
    {                                                                                                  \
        cols_fast_kernel<N, NT, R1, R2, R3, COL_FWD>, cols_fast_kernel<N, NT, R1, R2, R3, COL_INV>,    \
            cols_fast_kernel<N, NT, R1, R2, R3, COL_CONV>, cols_fast_kernel<N, NT, R1, R2, R3, COL_SPEC> \
    }
#define PVD_COLS_PIPE(N, NT, MINB, R1, R2, R3)                                                                     \
    {                                                                                                              \
        cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_FWD>, cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_INV>,    \
            cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_CONV>, cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_SPEC> \
    }
#define PVD_COLS(N, NT, MINB, R1, R2, R3) \
    { N, NT, PVD_COLS_FN(N, NT, R1, R2, R3), PVD_COLS_PIPE(N, NT, MINB, R1, R2, R3), {NT, NT, NT, NT}, {NT, NT, NT, NT}, 16 }
// As PVD_COLS, but the forward*spectrum*inverse pass (issue-bound, 2 transforms per tile) runs the one-tile-per-CTA
// kernel with its own schedule (X1, X2, X3), NTX threads and MINBX CTAs per SM: fewer radix stages = fewer
// block-wide barriers, and independent CTAs overlap each other's memory and arithmetic phases
// (512: 16*32 at 2 x 256 threads 0.280 ms vs 0.304 ms for the pipelined 8*8*8).
#define PVD_COLS_CX(N, NT, MINB, R1, R2, R3, NTX, MINBX, X1, X2, X3)                                               \
    {                                                                                                              \
        N, NT,                                                                                                     \
            {cols_fast_kernel<N, NT, R1, R2, R3, COL_FWD>, cols_fast_kernel<N, NT, R1, R2, R3, COL_INV>,           \
             cols_fast_kernel<N, NTX, X1, X2, X3, COL_CONV, MINBX>, cols_fast_kernel<N, NT, R1, R2, R3, COL_SPEC>}, \
            {cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_FWD>, cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_INV>, \
             nullptr, cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_SPEC>},                                         \
            {NT, NT, 0, NT}, {NT, NT, NTX, NT}, 16                                                                 \
    }
// NTC threads for the (issue-bound) forward*spectrum*inverse variant, NT for the others
#define PVD_COLS_C(N, NT, MINB, NTC, R1, R2, R3)                                                                \
    {                                                                                                           \
        N, NT, PVD_COLS_FN(N, NT, R1, R2, R3),                                                                  \
            {cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_FWD>, cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_INV>, \
             cols_pipe_kernel<N, NTC, 1, R1, R2, R3, COL_CONV>, cols_pipe_kernel<N, NT, MINB, R1, R2, R3, COL_SPEC>}, \
            {NT, NT, NTC, NT}, {NT, NT, NT, NT}, 16                                                             \
    }
#define PVD_COLS_NOPIPE(N, NT, R1, R2, R3) \
    { N, NT, PVD_COLS_FN(N, NT, R1, R2, R3), {nullptr, nullptr, nullptr, nullptr}, {0, 0, 0, 0}, {NT, NT, NT, NT}, 16 }
// Long transforms (1024, 1152): two full-line tiles do not fit one SM, so the forward / inverse / spectrum passes run the
// persistent double-buffered kernel on HALF-line tiles (N x 8 frequencies, NTP threads, TMA-staged), while the
// forward*spectrum*inverse pass keeps the full-line kernel (64-byte accesses at a plane stride waste DRAM pages: measured
// 0.457 vs 0.262 ms at 512) as a tile walk; prefetching the next tile's lines into L2 from inside the walk made it SLOWER
// (1024: 3.35 -> 4.18 ms, 1152: 3.70 -> 5.34 ms, profiles/r02_ab_long_columns.jsonl) and is not in the code.
#define PVD_COLS_HALF(N, NT, NTP, R1, R2, R3, NTX, X1, X2, X3)                                                         \
    {                                                                                                                  \
        N, NT,                                                                                                         \
            {cols_fast_kernel<N, NT, R1, R2, R3, COL_FWD>, cols_fast_kernel<N, NT, R1, R2, R3, COL_INV>,               \
             cols_fast_kernel<N, NTX, X1, X2, X3, COL_CONV, 1>, cols_fast_kernel<N, NT, R1, R2, R3, COL_SPEC>},        \
            {cols_pipe_kernel<N, NTP, 1, R1, R2, R3, COL_FWD, 8>, cols_pipe_kernel<N, NTP, 1, R1, R2, R3, COL_INV, 8>, \
             nullptr, cols_pipe_kernel<N, NTP, 1, R1, R2, R3, COL_SPEC, 8>},                                           \
            {NTP, NTP, 0, NTP}, {NT, NT, NTX, NT}, 8                                                                   \
    }
#define PVD_ROWS(N, NT, MINB, R1, R2, R3)                                                                  \
    {                                                                                                      \
        N, NT, rows_fwd_fast_kernel<N, NT, R1, R2, R3>, rows_inv_fast_kernel<N, NT, R1, R2, R3>,           \
            rows_fwd_pipe_kernel<N, NT, MINB, R1, R2, R3>, rows_inv_pipe_kernel<N, NT, MINB, R1, R2, R3>,  \
            (size_t)N * 17 * sizeof(float2) + RowStage<N>::BYTES + Sched<N, R1, R2, R3>::TOTAL * sizeof(float2) + 16, NT \
    }
// separate radix schedules for the forward (NTF threads, F1 F2 F3) and the inverse (NTI threads, I1 I2 I3) direction
#define PVD_ROWS2(N, NTF, F1, F2, F3, NTI, I1, I2, I3) PVD_ROWS2M(N, 1, NTF, F1, F2, F3, NTI, I1, I2, I3)
#define PVD_ROWS2M(N, MINB, NTF, F1, F2, F3, NTI, I1, I2, I3)                                                            \
    {                                                                                                                    \
        N, NTF, rows_fwd_fast_kernel<N, NTF, F1, F2, F3>, rows_inv_fast_kernel<N, NTI, I1, I2, I3>,                      \
            rows_fwd_pipe_kernel<N, NTF, MINB, F1, F2, F3>, rows_inv_pipe_kernel<N, NTI, MINB, I1, I2, I3>,              \
            (size_t)N * 17 * sizeof(float2) + RowStage<N>::BYTES +                                                       \
                (Sched<N, F1, F2, F3>::TOTAL > Sched<N, I1, I2, I3>::TOTAL ? Sched<N, F1, F2, F3>::TOTAL                 \
                                                                           : Sched<N, I1, I2, I3>::TOTAL) * sizeof(float2) + 16, \
            NTI                                                                                                          \
    }
#define PVD_ROWS_NOPIPE(N, NT, R1, R2, R3) \
    { N, NT, rows_fwd_fast_kernel<N, NT, R1, R2, R3>, rows_inv_fast_kernel<N, NT, R1, R2, R3>, nullptr, nullptr, 0, NT }
const FastCols kFastCols[] = {
    PVD_COLS_CX(512, 512, 1, 8, 8, 8, 256, 2, 16, 32, 1),   // x pass 16*32; 32*16 measured 0.266 vs 0.263 ms (profiles/r02_ab_p3_without_bounds_predicates.jsonl)
    PVD_COLS(256, 256, 2, 16, 16, 1),
    PVD_COLS(400, 320, 1, 20, 20, 1),
    PVD_COLS(576, 384, 1, 24, 24, 1),   // 512 + kernel reach ('same' mode of 512-wide volumes)
    PVD_COLS(432, 384, 1, 18, 24, 1),
    PVD_COLS_CX(288, 288, 2, 16, 18, 1, 288, 2, 16, 18, 1),   // 256 + kernel reach; x pass as at 320 / 180 below (0.081 -> 0.074 ms)
    PVD_COLS_HALF(1024, 1024, 512, 16, 8, 8, 512, 32, 32, 1),   // x pass: two radix-32 stages, 512 threads (one exchange per transform)
    // slab decomposition of the 1024 x 1024 x 800 volume ('same' mode): 1024 + reach -> 1152, slabs of
    // 256 / 128 planes + 50 halo planes -> 320 / 180 (192: other kernel sizes)
    PVD_COLS_HALF(1152, 768, 384, 8, 12, 12, 768, 8, 12, 12),
    // x pass of the slab plans as the one-tile-per-CTA walk at 2 CTAs per SM (as at 512): the persistent pipelined form of the
    // forward*spectrum*inverse kernel spilled at these lengths (320: 324 B, 180: 276 B) - per 4-rank slab 0.898 -> 0.659 ms,
    // per 8-rank slab 0.485 -> 0.439 ms (profiles/r02_ab_slab_x_pass.jsonl)
    PVD_COLS_CX(320, 320, 2, 16, 20, 1, 320, 2, 16, 20, 1),
    PVD_COLS(192, 256, 3, 12, 16, 1),
    // 128-plane slab + 50 halo planes = 178 -> 180 (8 ranks; 192 costs 6.7 % more points)
    PVD_COLS_CX(180, 288, 3, 10, 18, 1, 288, 2, 10, 18, 1),
};
const FastRows kFastRows[] = {
    PVD_ROWS(400, 320, 2, 20, 20, 1),
    PVD_ROWS(256, 256, 3, 16, 16, 1),
    PVD_ROWS(512, 512, 1, 8, 8, 8),
    // 400 / 256 + kernel reach ('same' mode).  Schedules per direction (profiles/r02_ab_rows_small.jsonl): the 432-point inverse
    // with three stages and 576 threads 0.308 -> 0.285 ms (its forward pass is faster with two: 0.202 vs 0.211-0.221), the
    // 288-point inverse as 18*16 0.041 -> 0.039 ms
    PVD_ROWS2M(432, 1, 384, 18, 24, 1, 576, 6, 6, 12),
    PVD_ROWS2M(288, 2, 288, 16, 18, 1, 288, 18, 16, 1),
    // The long rows leave room for ONE CTA per SM, so every block-wide barrier idles the SM: two big-radix stages (one
    // exchange per transform) beat three small ones wherever the registers hold (profiles/r02_ab_long_rows.jsonl: 800 forward
    // 1.74 -> 1.42 ms, 840 forward 2.54 -> 1.69 ms, 840 inverse 3.13 -> 2.39 ms per 1024^2 / 1152^2 rows; the 800-point inverse
    // is faster with three stages, 2.02 vs 2.12 ms).
    PVD_ROWS2(800, 512, 32, 25, 1, 640, 8, 10, 10),   // 1024 x 1024 x 800, reference mode
    PVD_ROWS2(840, 480, 30, 28, 1, 480, 28, 30, 1),   // 800 + kernel reach ('same' mode: 825 points needed): the longest length whose
                                                      // exchange tile + 32-row staging buffer still fit one SM, so it runs pipelined
    PVD_ROWS_NOPIPE(864, 576, 8, 9, 12),       // tile + staging buffer exceed one SM's shared memory
};
const FastCols* find_fast_cols(int n) {
    for (const auto& e : kFastCols)
        if (e.N == n) return &e;
    return nullptr;
}
const FastRows* find_fast_rows(int n) {
    for (const auto& e : kFastRows)
        if (e.N == n) return &e;
    return nullptr;
}

// The ONE environment hook the library reads (tests only): PVD_FORCE_GENERIC=1 routes every length through the
// any-length engine of fft_passes.cuh, so the parity tests can cover it at sizes the specialised menu would take.
bool force_generic() {
    const char* e = getenv("PVD_FORCE_GENERIC");
    return e && e[0] == '1';
}

// Transform length for an axis that needs at least n points: a length on the specialised menu wins when
// it costs at most 15 % more points than the best generic {2,3,5,7}-smooth length.
int good_size_axis(int n, int axis) {
    const int g = good_size(n);
    int best = -1;
    if (axis == 2) {
        for (const auto& e : kFastRows)
            if (e.N >= n && (best < 0 || e.N < best)) best = e.N;
    } else {
        for (const auto& e : kFastCols)
            if (e.N >= n && (best < 0 || e.N < best)) best = e.N;
    }
    if (force_generic()) return g;
    return (best > 0 && (double)best <= 1.15 * (double)n) ? best : g;
}

}  // namespace

struct pvd_plan {
    int n[3], m[3], olo[3], on[3], k[3], ke[3];
    int Nh, Sz;
    int algo;
    Stages st[3];
    int rowLlog, colWlog[2];          // tile shapes (generic engine)
    size_t rowSmem, colSmem[2];
    const FastCols* fastCols[2] = {nullptr, nullptr};  // size-specialised kernels, when the length is on the menu
    const FastRows* fastRows = nullptr;
    bool usePipe = true;
    bool pdl = true;  // programmatic dependent launch of the specialised kernels
    int sms = 0;          // SMs of the device (set with the workspace)
    int reserve_sms = 0;  // SMs the persistent grids leave free (pvd_plan_reserve_sms): room for a concurrent NCCL kernel
    // persistent grid of a kernel that keeps `per_sm` CTAs resident per SM: `full` CTAs fill the GPU
    int pgrid(int full) const {
        if (reserve_sms <= 0 || sms <= 0 || full < sms) return full;
        const int per_sm = full / sms;
        return std::max(per_sm, full - reserve_sms * per_sm);
    }
    // TMA variant of the persistent y passes: tensor maps over the work buffer (forward: n[1] rows, inverse: m[1] rows)
    bool tmaRows = false;  // TMA staging in the persistent row passes
    bool tmaCols = false;
    CUtensorMap tmapCols[2];
    int tmapRows[2] = {0, 0};
    int pipeGrid[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};  // persistent grid size per axis / mode
    int rowPipeGrid[2] = {0, 0};                        // persistent grid size of the row passes (fwd, inv)
    int fnGrid[2] = {0, 0};                             // CTAs of the one-tile-per-CTA CONV kernel resident on the GPU
    // direct (TMA) path
    int want_algo = PVD_ALGO_AUTO;
    int dbox[3] = {0, 0, 0};
    size_t off_taps = 0, dsmem = 0;
    bool dcubic = false;       // cubic K in {3, 5, 7}: register-tiled kernel with the taps in the argument struct
    bool dwrap = false;        // circular (reference) boundary handled by the direct kernel
    std::vector<float> h_taps; // flipped kernel on the host (cubic path)
    float* taps() const { return reinterpret_cast<float*>(ws + off_taps); }
    size_t off_tw[3], off_buf, off_spec, off_flag, ws_bytes;
    char* ws = nullptr;
    bool kernel_set = false;
    // measurement hook
    bool prof = false;
    bool ev_made = false;
    cudaEvent_t ev[PVD_MAX_PASSES + 1];
    int npass = 0;
    double pass_bytes[PVD_MAX_PASSES];
    const char* pass_names[PVD_MAX_PASSES];
    void mark(cudaStream_t s, const char* name, double bytes) {  // call before each launch
        if (!prof || npass >= PVD_MAX_PASSES) return;
        cudaEventRecord(ev[npass], s);
        pass_names[npass] = name;
        pass_bytes[npass] = bytes;
        ++npass;
    }
    void mark_end(cudaStream_t s) {
        if (prof) cudaEventRecord(ev[npass], s);
    }
    float2* tw(int a) const { return reinterpret_cast<float2*>(ws + off_tw[a]); }
    float2* buf() const { return reinterpret_cast<float2*>(ws + off_buf); }
    float2* spec() const { return reinterpret_cast<float2*>(ws + off_spec); }
    int* flag() const { return reinterpret_cast<int*>(ws + off_flag); }
};

namespace {

#ifndef PVD_EMULATE
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}
#endif

int launch_rows_fwd(const pvd_plan* p, const float* const* in, const float* w, int T, long long s0, long long s1,
                    const int ext[3], cudaStream_t stream, int out_plane0 = 0) {
    RowFwdArgs a;
    memset(&a, 0, sizeof a);
    for (int t = 0; t < T; ++t) {
        a.in[t] = in[t];
        a.w[t] = w ? w[t] : 1.0f;
    }
    a.T = T;
    a.in_s0 = s0;
    a.in_s1 = s1;
    a.n0 = ext[0];
    a.n1 = ext[1];
    a.n2 = ext[2];
    a.out_s0 = (long long)p->m[1] * p->Sz;
    a.out_s1 = p->Sz;
    a.out = p->buf() + (long long)out_plane0 * a.out_s0;
    a.M2 = p->m[2];
    a.Nh = p->Nh;
    a.Llog = p->rowLlog;
    a.tw = p->tw(2);
    a.st = p->st[2];
    a.dense_in = (s0 == (long long)ext[1] * s1) ? 1 : 0;
    a.dense = (a.dense_in && a.out_s0 == (long long)ext[1] * a.out_s1) ? 1 : 0;

    const long long nrows = (long long)ext[0] * ext[1];
    const long long per = 2LL << p->rowLlog;
    const long long nblk = (nrows + per - 1) / per;
    if (nblk <= 0) return PVD_OK;
    if (p->fastRows && p->usePipe && p->rowPipeGrid[0] > 0 && T == 1 && nrows < 2000000000LL && s0 % 4 == 0 && s1 % 4 == 0 &&
        ((uintptr_t)in[0] & 15) == 0) {
        const FastRows* f = p->fastRows;
        const int grid = (int)std::min<long long>((nrows + 31) / 32, p->pgrid(p->rowPipeGrid[0]));
        a.use_tma = 0;
        a.error_flag = p->flag() + 1;
#ifndef PVD_EMULATE
        // TMA staging of the 32-row tiles: a 2-D map of the (dense) activity volume, encoded per call (the pointer is the caller's)
        if (p->tmaRows && a.dense_in && f->N % 16 == 0 && f->N / 4 <= 256 && ext[2] % 4 == 0) {
            const cuuint64_t gdim[2] = {(cuuint64_t)ext[2], (cuuint64_t)nrows};
            const cuuint64_t gstr[1] = {(cuuint64_t)s1 * 4};
            const cuuint32_t box[2] = {(cuuint32_t)(f->N / 4), 32};
            const cuuint32_t estr[2] = {1, 1};
            if (get_encode_tiled()(&a.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(in[0]), gdim, gstr, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                a.use_tma = 1;
        }
#endif
        // with SMs reserved for a concurrent exchange kernel the first pass is a plain launch: as a programmatic dependent it
        // would be resident (and hold every CTA slot) before the exchange kernel of the other stream becomes eligible
        PVD_LAUNCH_PDL(p->pdl && p->reserve_sms == 0, f->fwdPipe, dim3((unsigned)grid), dim3(f->NT), f->smemPipe, stream, a);
        PVD_CUDA_CHECK("rows_fwd_pipe_kernel");
        return PVD_OK;
    }
    if (p->fastRows) {
        const FastRows* f = p->fastRows;
        const size_t smem = ((size_t)f->N * 17 + 4 * f->N) * sizeof(float2);
        PVD_LAUNCH(f->fwd, dim3((unsigned)((nrows + 31) / 32)), dim3(f->NT), smem, stream, a);
        PVD_CUDA_CHECK("rows_fwd_fast_kernel");
        return PVD_OK;
    }
    PVD_LAUNCH(rows_fwd_kernel, dim3((unsigned)nblk), dim3(PVD_BLOCK), p->rowSmem, stream, a);
    PVD_CUDA_CHECK("rows_fwd_kernel");
    return PVD_OK;
}

int launch_cols(const pvd_plan* p, int axis, int mode, const float2* in, float2* out, int outer0, int nouter,
                int n_in, int out_lo, int out_n, float scale, cudaStream_t stream) {
    ColArgs a;
    memset(&a, 0, sizeof a);
    a.in = in;
    a.out = out;
    a.spec = p->spec();
    const long long plane = (long long)p->m[1] * p->Sz;
    a.es = axis == 0 ? plane : p->Sz;
    a.os = axis == 0 ? p->Sz : plane;
    a.outer0 = outer0;
    a.n_in = n_in;
    a.M = p->m[axis];
    a.out_lo = out_lo;
    a.out_n = out_n;
    a.nzf = p->Nh;
    a.Wlog = p->colWlog[axis];
    a.mode = mode;
    a.scale = scale;
    a.tw = p->tw(axis);
    a.st = p->st[axis];
    const int W = 1 << a.Wlog;
    if (nouter <= 0) return PVD_OK;
    if (p->fastCols[axis] && p->fastCols[axis]->pipe[mode] && p->usePipe && p->pipeGrid[axis][mode] > 0) {
        const FastCols* f = p->fastCols[axis];
        ColPipeArgs pa;
        memset(&pa.tmap, 0, sizeof pa.tmap);
        pa.use_tma = 0;
        pa.error_flag = p->flag() + 1;
        if (axis == 1 && p->tmaCols && (mode == COL_FWD || mode == COL_INV) && in == p->buf() &&
            n_in == p->tmapRows[mode == COL_FWD ? 0 : 1]) {
            pa.tmap = p->tmapCols[mode == COL_FWD ? 0 : 1];
            pa.use_tma = 1;
        }
        pa.c = a;
        pa.ntz = (p->Nh + f->pipeW - 1) / f->pipeW;
        pa.ntiles = pa.ntz * nouter;
        pa.ntz_magic = (unsigned)((0x100000000ULL + pa.ntz - 1) / pa.ntz);  // exact for t * ntz < 2^32 (ntz == 1: magic wraps to 0)
        if (pa.ntz == 1) pa.ntz_magic = 0xFFFFFFFFu;
        const size_t smem = ((size_t)f->N * 2 * f->pipeW + 4 * f->N) * sizeof(float2);
        const int grid = std::min(pa.ntiles, p->pgrid(p->pipeGrid[axis][mode]));
        PVD_LAUNCH_PDL(p->pdl, f->pipe[mode], dim3((unsigned)grid), dim3(f->pipeNT[mode]), smem, stream, pa);
        PVD_CUDA_CHECK("cols_pipe_kernel");
        return PVD_OK;
    }
    if (p->fastCols[axis]) {
        const FastCols* f = p->fastCols[axis];
        const size_t smem = ((size_t)f->N * 16 + 4 * f->N) * sizeof(float2);
        if (mode == COL_CONV && p->fnGrid[axis] > 0) {
            // persistent walk over the tiles with the CTAs that are resident anyway (see cols_fast_kernel); launched
            // normally: as a programmatic dependent this pass measured +70 us per C3 volume
            a.loop_ntz = (p->Nh + 15) / 16;
            a.loop_ntiles = a.loop_ntz * nouter;
            const int grid = std::min(a.loop_ntiles, p->pgrid(p->fnGrid[axis]));
            PVD_LAUNCH_PDL(false, f->fn[mode], dim3((unsigned)grid), dim3(f->fnNT[mode]), smem, stream, a);
            PVD_CUDA_CHECK("cols_fast_kernel (looped)");
            return PVD_OK;
        }
        PVD_LAUNCH_PDL(false, f->fn[mode], dim3((unsigned)((p->Nh + 15) / 16), (unsigned)nouter), dim3(f->fnNT[mode]), smem, stream, a);
        PVD_CUDA_CHECK("cols_fast_kernel");
        return PVD_OK;
    }
    PVD_LAUNCH(cols_kernel, dim3((unsigned)((p->Nh + W - 1) / W), (unsigned)nouter), dim3(PVD_BLOCK), p->colSmem[axis],
               stream, a);
    PVD_CUDA_CHECK("cols_kernel");
    return PVD_OK;
}

int plan_finish(pvd_plan* p) {
    for (int i = 0; i < 3; ++i) {
        if (p->n[i] < 1 || p->k[i] < 1 || p->on[i] < 1 || p->olo[i] < 0)
            return fail(PVD_ERR_INVALID, "axis %d: extents must be positive (n=%d k=%d out_n=%d out_lo=%d)", i, p->n[i],
                        p->k[i], p->on[i], p->olo[i]);
        if (p->m[i] < p->n[i] || p->olo[i] + p->on[i] > p->m[i])
            return fail(PVD_ERR_INVALID, "axis %d: transform extent %d too small for n=%d / output [%d,%d)", i, p->m[i],
                        p->n[i], p->olo[i], p->olo[i] + p->on[i]);
        p->ke[i] = std::min(p->k[i], p->m[i]);
        if (!make_stages(p->m[i], p->st[i])) return fail(PVD_ERR_UNSUPPORTED, "axis %d: cannot factor length %d", i, p->m[i]);
    }
    p->Nh = p->m[2] / 2 + 1;
    p->Sz = (int)align_up((size_t)p->Nh, 16);
    // rows: 2^Llog complex lines, smem = 2 buffers * M2 * (L+1) float2
    {
        int Llog = 4;
        auto smem = [&](int l) { return 2 * (size_t)p->m[2] * ((1u << l) + 1) * sizeof(float2); };
        while (Llog > 0 && smem(Llog) > 114 * 1024) --Llog;
        if (smem(Llog) > kMaxSmem) return fail(PVD_ERR_UNSUPPORTED, "axis 2 length %d exceeds the shared-memory engine", p->m[2]);
        p->rowLlog = Llog;
        p->rowSmem = smem(Llog);
    }
    for (int a = 0; a < 2; ++a) {
        int Wlog = 4;
        auto smem = [&](int l) { return 2 * (size_t)p->m[a] * (1u << l) * sizeof(float2); };
        while (Wlog > 0 && smem(Wlog) > 72 * 1024) --Wlog;
        if (smem(Wlog) > kMaxSmem) return fail(PVD_ERR_UNSUPPORTED, "axis %d length %d exceeds the shared-memory engine", a, p->m[a]);
        p->colWlog[a] = Wlog;
        p->colSmem[a] = smem(Wlog);
    }
    size_t off = 0;
    for (int a = 0; a < 3; ++a) {
        p->off_tw[a] = off;
        off = align_up(off + (size_t)p->m[a] * sizeof(float2), 256);
    }
    const size_t vol = (size_t)p->m[0] * p->m[1] * p->Sz * sizeof(float2);
    p->off_buf = off;
    off = align_up(off + vol, 256);
    p->off_spec = off;
    off = align_up(off + vol, 256);
    p->off_flag = off;
    off = align_up(off + 256, 256);
    p->ws_bytes = off;
    p->algo = PVD_ALGO_FFT;
    if (!force_generic()) {
        for (int a = 0; a < 2; ++a) p->fastCols[a] = find_fast_cols(p->m[a]);
        p->fastRows = find_fast_rows(p->m[2]);
        // the specialised column kernels index the work buffer with 32-bit element offsets
        if ((double)p->m[0] * p->m[1] * p->Sz >= 2147483648.0) p->fastCols[0] = p->fastCols[1] = nullptr;
    }
    // ---- direct tiled convolution (TMA halo tiles), small kernels only.  Zero-boundary 'same' geometry for any K0, K1 <= 9,
    // K2 in {3, 5, 7}; the circular reference geometry (core/kernel_convolution.py:71-74) for cubic K in {3, 5, 7}.
    bool same_geom = true, ref_geom = true;
    for (int i = 0; i < 3; ++i) {
        same_geom = same_geom && p->on[i] == p->n[i] && p->olo[i] == p->k[i] / 2;
        ref_geom = ref_geom && p->on[i] == p->n[i] && p->olo[i] == 0 && p->m[i] == p->n[i] && p->k[i] <= p->n[i];
    }
    const bool kz_ok = (p->k[2] == 3 || p->k[2] == 5 || p->k[2] == 7) && p->n[2] % 4 == 0;
    const bool cubic = kz_ok && p->k[0] == p->k[2] && p->k[1] == p->k[2];
    const bool direct_ok = kz_ok && ((same_geom && p->k[0] <= 9 && p->k[1] <= 9) || (ref_geom && cubic));
    const long long taps = (long long)p->k[0] * p->k[1] * p->k[2];
    if (p->want_algo == PVD_ALGO_DIRECT && !direct_ok)
        return fail(PVD_ERR_UNSUPPORTED,
                    "direct algorithm needs K2 in {3,5,7}, n2 %% 4 == 0 and either zero-boundary 'same' geometry with K0,K1 <= 9 "
                    "or the circular reference geometry with a cubic kernel no larger than the volume");
    if (direct_ok && (p->want_algo == PVD_ALGO_DIRECT || (p->want_algo == PVD_ALGO_AUTO && taps <= 125))) {
        p->algo = PVD_ALGO_DIRECT;
        p->dcubic = cubic;
        p->dwrap = !same_geom;
        if (cubic) {
            p->dbox[0] = kCubTX + p->k[0] - 1;
            p->dbox[1] = kCubTY + p->k[1] - 1;
            p->dbox[2] = kCubTZ + 8;
            p->dsmem = 128 + (size_t)p->dbox[0] * p->dbox[1] * p->dbox[2] * sizeof(float) + 64;
        } else {
            p->dbox[0] = kDirTX + p->k[0] - 1;
            p->dbox[1] = kDirTY + p->k[1] - 1;
            p->dbox[2] = kDirTZ + 8;  // 4 lead-in (16-byte aligned TMA start) + 64 + up to 4 right reach
            p->dsmem = 128 + ((size_t)p->dbox[0] * p->dbox[1] * p->dbox[2] + (size_t)taps + 16) * sizeof(float) + 64;
        }
        p->off_taps = 0;
        p->off_flag = align_up((size_t)taps * sizeof(float), 256);
        p->ws_bytes = p->off_flag + 256;
    }
    return PVD_OK;
}


// Tensor maps of the y passes (axis 1) over the work buffer: dims (2*Sz floats, rows, m0 planes), one box = 16
// frequencies x BOXR rows of one plane.  Rows beyond `rows` are out of bounds = zero filled (the implicit padding).
void make_col_tensor_maps(pvd_plan* p) {
    p->tmaCols = false;
    p->tmaRows = false;
#ifndef PVD_EMULATE
    p->tmaRows = get_encode_tiled() != nullptr;
    const FastCols* f = p->fastCols[1];
    EncodeTiledFn enc = get_encode_tiled();
    if (!f || !f->pipe[COL_FWD] || !enc) return;
    const int rows[2] = {p->n[1], p->m[1]};
    for (int i = 0; i < 2; ++i) {
        memset(&p->tmapCols[i], 0, sizeof(CUtensorMap));
        const cuuint64_t gdim[3] = {(cuuint64_t)2 * p->Sz, (cuuint64_t)rows[i], (cuuint64_t)p->m[0]};
        const cuuint64_t gstr[2] = {(cuuint64_t)p->Sz * 8, (cuuint64_t)p->m[1] * p->Sz * 8};
        const cuuint32_t box[3] = {(cuuint32_t)(2 * f->pipeW), (cuuint32_t)tma_box_rows(f->N), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&p->tmapCols[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p->buf(), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return;
        p->tmapRows[i] = rows[i];
    }
    p->tmaCols = true;
#endif
}

template <int K, int ZS>
int launch_cubic(pvd_plan* p, const float* in, const CUtensorMap& tmap, const float* density, float rho_ref, float rho_min, float rho_cut,
                 float scale, float* dose, cudaStream_t stream) {
    CubicArgs<K> a;
    memset(&a, 0, sizeof a);
    memcpy(a.taps, p->h_taps.data(), sizeof a.taps);
    a.in = in;
    a.n0 = p->n[0];
    a.n1 = p->n[1];
    a.n2 = p->n[2];
    const int c = p->dwrap ? 0 : K / 2;  // the circular mode anchors the kernel at the origin: centre 0
    a.o0 = a.o1 = c - (K - 1);
    a.o2 = p->dwrap ? -4 * ((K - 1 + 3) / 4) : -4;
    a.out = dose;
    a.density = density;
    a.rho_ref = rho_ref;
    a.rho_min = rho_min;
    a.rho_cut = rho_cut;
    a.scale = scale;
    a.wrap = p->dwrap ? 1 : 0;
    a.error_flag = p->flag();
    const dim3 grid((unsigned)((p->n[2] + kCubTZ - 1) / kCubTZ), (unsigned)((p->n[1] + kCubTY - 1) / kCubTY),
                    (unsigned)((p->n[0] + kCubTX - 1) / kCubTX));
    PVD_LAUNCH((direct_conv_cubic_kernel<K, ZS>), grid, dim3(kCubThreads), p->dsmem, stream, tmap, a);
    PVD_CUDA_CHECK("direct_conv_cubic_kernel");
    return PVD_OK;
}

// ZS (first tap's index inside a staged 12-float run): zero boundary 4 - (K-1-K/2); circular 4*ceil((K-1)/4) - (K-1)
int execute_direct_cubic(pvd_plan* p, const float* in, const CUtensorMap& tmap, const float* density, float rho_ref, float rho_min,
                         float rho_cut, float scale, float* dose, cudaStream_t stream) {
    p->npass = 0;
    p->mark(stream, "D1 direct tiled conv (TMA halo tiles + density)", (density ? 12.0 : 8.0) * p->n[0] * p->n[1] * p->n[2]);
    int rc;
    if (p->dwrap) {
        rc = p->k[2] == 3 ? launch_cubic<3, 2>(p, in, tmap, density, rho_ref, rho_min, rho_cut, scale, dose, stream)
           : p->k[2] == 5 ? launch_cubic<5, 0>(p, in, tmap, density, rho_ref, rho_min, rho_cut, scale, dose, stream)
                          : launch_cubic<7, 2>(p, in, tmap, density, rho_ref, rho_min, rho_cut, scale, dose, stream);
    } else {
        rc = p->k[2] == 3 ? launch_cubic<3, 3>(p, in, tmap, density, rho_ref, rho_min, rho_cut, scale, dose, stream)
           : p->k[2] == 5 ? launch_cubic<5, 2>(p, in, tmap, density, rho_ref, rho_min, rho_cut, scale, dose, stream)
                          : launch_cubic<7, 1>(p, in, tmap, density, rho_ref, rho_min, rho_cut, scale, dose, stream);
    }
    p->mark_end(stream);
    return rc;
}

int execute_direct(pvd_plan* p, const float* const* h_act, const float* h_weights, int T, const float* density, float rho_ref,
                   float rho_min, float rho_cut, float scale, float* dose, cudaStream_t stream) {
    if (T != 1) return fail(PVD_ERR_INVALID, "direct algorithm takes one activity volume: pre-accumulate with pvd_weighted_sum");
    const float* in = h_act[0];
    if (((uintptr_t)in & 15) || ((uintptr_t)dose & 15) || (density && ((uintptr_t)density & 15)))
        return fail(PVD_ERR_INVALID, "direct algorithm needs 16-byte aligned volumes");
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
#ifndef PVD_EMULATE
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(PVD_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t gdim[3] = {(cuuint64_t)p->n[2], (cuuint64_t)p->n[1], (cuuint64_t)p->n[0]};  // innermost first
    const cuuint64_t gstr[2] = {(cuuint64_t)p->n[2] * 4, (cuuint64_t)p->n[2] * p->n[1] * 4};     // bytes, dims 1..2
    const cuuint32_t box[3] = {(cuuint32_t)p->dbox[2], (cuuint32_t)p->dbox[1], (cuuint32_t)p->dbox[0]};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PVD_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
#endif
    if (p->dcubic) return execute_direct_cubic(p, in, tmap, density, rho_ref, rho_min, rho_cut, scale * (h_weights ? h_weights[0] : 1.f), dose, stream);
    DirectArgs a;
    memset(&a, 0, sizeof a);
    a.in = in;
    a.n0 = p->n[0];
    a.n1 = p->n[1];
    a.n2 = p->n[2];
    a.taps = p->taps();
    a.K0 = p->k[0];
    a.K1 = p->k[1];
    a.o0 = p->k[0] / 2 - (p->k[0] - 1);
    a.o1 = p->k[1] / 2 - (p->k[1] - 1);
    a.o2 = -4;
    a.bx = p->dbox[0];
    a.by = p->dbox[1];
    a.bz = p->dbox[2];
    a.out = dose;
    a.density = density;
    a.rho_ref = rho_ref;
    a.rho_min = rho_min;
    a.rho_cut = rho_cut;
    a.scale = scale * (h_weights ? h_weights[0] : 1.f);
    a.error_flag = p->flag();
    const dim3 grid((unsigned)((p->n[2] + kDirTZ - 1) / kDirTZ), (unsigned)((p->n[1] + kDirTY - 1) / kDirTY),
                    (unsigned)((p->n[0] + kDirTX - 1) / kDirTX));
    p->npass = 0;
    p->mark(stream, "D1 direct tiled conv (TMA halo tiles + density)", (density ? 12.0 : 8.0) * p->n[0] * p->n[1] * p->n[2]);
    switch (p->k[2]) {
        case 3: PVD_LAUNCH((direct_conv_kernel<3, 0>), grid, dim3(kDirThreads), p->dsmem, stream, tmap, a); break;
        case 5: PVD_LAUNCH((direct_conv_kernel<5, 0>), grid, dim3(kDirThreads), p->dsmem, stream, tmap, a); break;
        default: PVD_LAUNCH((direct_conv_kernel<7, 0>), grid, dim3(kDirThreads), p->dsmem, stream, tmap, a); break;
    }
    PVD_CUDA_CHECK("direct_conv_kernel");
    p->mark_end(stream);
    return PVD_OK;
}

}  // namespace

extern "C" {

int pvd_version(void) { return PVD_VERSION; }
#ifndef PVD_BUILD_ID
#define PVD_BUILD_ID "unstamped"
#endif
const char* pvd_build_id(void) { return PVD_BUILD_ID; }
const char* pvd_last_error(void) { return g_err.c_str(); }
int pvd_good_fft_size(int n) { return good_size_axis(n, 0); }
int pvd_good_fft_size_axis(int n, int axis) { return good_size_axis(n, axis == 2 ? 2 : 0); }

int pvd_plan_create_ex(pvd_plan** out, const int n[3], const int m[3], const int out_lo[3], const int out_n[3],
                       const int k[3], int algo) {
    if (!out || !n || !out_lo || !out_n || !k) return fail(PVD_ERR_INVALID, "null argument");
    if (algo != PVD_ALGO_AUTO && algo != PVD_ALGO_FFT && algo != PVD_ALGO_DIRECT && algo != PVD_ALGO_FFT_UNPIPELINED)
        return fail(PVD_ERR_INVALID, "unknown algo %d", algo);
    pvd_plan* p = new pvd_plan();
    p->want_algo = algo == PVD_ALGO_FFT_UNPIPELINED ? PVD_ALGO_FFT : algo;
    p->usePipe = algo != PVD_ALGO_FFT_UNPIPELINED;
    for (int i = 0; i < 3; ++i) {
        p->n[i] = n[i];
        p->k[i] = k[i];
        p->olo[i] = out_lo[i];
        p->on[i] = out_n[i];
        p->m[i] = (m && m[i] > 0) ? m[i] : good_size_axis(std::max(n[i], out_lo[i] + out_n[i]), i);
    }
    int rc = plan_finish(p);
    if (rc != PVD_OK) {
        delete p;
        return rc;
    }
    *out = p;
    return PVD_OK;
}

int pvd_plan_create(pvd_plan** out, const int n[3], const int k[3], int boundary, int algo) {
    if (!out || !n || !k) return fail(PVD_ERR_INVALID, "null argument");
    int m[3], lo[3], on[3];
    for (int i = 0; i < 3; ++i) {
        if (n[i] < 1 || k[i] < 1) return fail(PVD_ERR_INVALID, "axis %d: n=%d k=%d must be positive", i, n[i], k[i]);
        on[i] = n[i];
        if (boundary == PVD_BOUNDARY_REFERENCE) {
            m[i] = n[i];
            lo[i] = 0;
        } else if (boundary == PVD_BOUNDARY_SAME) {
            const int c = k[i] / 2;
            lo[i] = c;
            m[i] = good_size_axis(std::max(std::max(n[i] + k[i] - 1 - c, k[i]), n[i] + c), i);
        } else {
            return fail(PVD_ERR_INVALID, "unknown boundary mode %d", boundary);
        }
    }
    return pvd_plan_create_ex(out, n, m, lo, on, k, algo);
}

int pvd_plan_get_info(const pvd_plan* p, pvd_plan_info* info) {
    if (!p || !info) return fail(PVD_ERR_INVALID, "null argument");
    for (int i = 0; i < 3; ++i) {
        info->n[i] = p->n[i];
        info->m[i] = p->m[i];
        info->out_lo[i] = p->olo[i];
        info->out_n[i] = p->on[i];
        info->k[i] = p->k[i];
    }
    info->algo = p->algo;
    info->passes = p->algo == PVD_ALGO_DIRECT ? 1 : 5;
    info->workspace_bytes = p->ws_bytes;
    if (p->algo == PVD_ALGO_DIRECT) {
        info->hbm_bytes_per_execute = 12.0 * p->n[0] * p->n[1] * p->n[2];
        return PVD_OK;
    }
    const double c = 8.0 * p->Nh;  // bytes of one half-spectrum row
    const double real_in = 4.0 * p->n[0] * p->n[1] * p->n[2];
    const double real_out = 4.0 * p->on[0] * p->on[1] * p->on[2];
    double b = real_in + c * p->n[0] * p->n[1];                                   // P1
    b += c * p->n[0] * (p->n[1] + p->m[1]);                                       // P2
    b += c * p->m[1] * ((double)p->n[0] + p->m[0] + p->on[0]);                    // P3 (+ spectrum)
    b += c * p->on[0] * ((double)p->m[1] + p->on[1]);                             // P4
    b += c * p->on[0] * p->on[1] + 2.0 * real_out;                                // P5 (+ density)
    info->hbm_bytes_per_execute = b;
    return PVD_OK;
}

int pvd_plan_workspace_bytes(const pvd_plan* p, size_t* bytes) {
    if (!p || !bytes) return fail(PVD_ERR_INVALID, "null argument");
    *bytes = p->ws_bytes;
    return PVD_OK;
}

int pvd_plan_set_workspace(pvd_plan* p, void* workspace, size_t bytes, void* stream_) {
    if (!p || !workspace) return fail(PVD_ERR_INVALID, "null argument");
    if (bytes < p->ws_bytes) return fail(PVD_ERR_INVALID, "workspace too small: %zu < %zu", bytes, p->ws_bytes);
    if ((uintptr_t)workspace % 256 != 0) return fail(PVD_ERR_INVALID, "workspace must be 256-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    p->ws = (char*)workspace;
    p->kernel_set = false;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (p->algo == PVD_ALGO_DIRECT) {
        if (PVD_SET_SMEM((direct_conv_kernel<3, 0>), kMaxSmem) != 0 || PVD_SET_SMEM((direct_conv_kernel<5, 0>), kMaxSmem) != 0 ||
            PVD_SET_SMEM((direct_conv_kernel<7, 0>), kMaxSmem) != 0 || PVD_SET_SMEM((direct_conv_cubic_kernel<3, 3>), kMaxSmem) != 0 ||
            PVD_SET_SMEM((direct_conv_cubic_kernel<3, 2>), kMaxSmem) != 0 || PVD_SET_SMEM((direct_conv_cubic_kernel<5, 2>), kMaxSmem) != 0 ||
            PVD_SET_SMEM((direct_conv_cubic_kernel<5, 0>), kMaxSmem) != 0 || PVD_SET_SMEM((direct_conv_cubic_kernel<7, 1>), kMaxSmem) != 0 ||
            PVD_SET_SMEM((direct_conv_cubic_kernel<7, 2>), kMaxSmem) != 0)
            return fail(PVD_ERR_CUDA, "cannot opt in to large shared memory (direct)");
        cudaMemsetAsync(p->flag(), 0, 256, stream);
        return PVD_OK;
    }
    for (int a = 0; a < 3; ++a) {
        PVD_LAUNCH(twiddle_kernel, dim3((unsigned)std::min(64, (p->m[a] + 127) / 128)), dim3(128), 0, stream, p->tw(a),
                   p->m[a]);
        PVD_CUDA_CHECK("twiddle_kernel");
    }
    // The specialised kernels build their shared-memory twiddle tables from these buffers BEFORE their
    // grid-dependency wait (programmatic dependent launch), so the tables must be complete before any of them can be
    // launched: plan set-up is a one-time call, a host-side wait here is the simplest guarantee.
    if (cudaStreamSynchronize(stream) != cudaSuccess) return fail(PVD_ERR_CUDA, "twiddle tables: stream synchronize failed");
    cudaMemsetAsync(p->flag(), 0, 256, stream);
    make_col_tensor_maps(p);
    for (int a = 0; a < 2; ++a)
        if (p->fastCols[a])
            for (int md = 0; md < 4; ++md)
            {
                const FastCols* f = p->fastCols[a];
                if (PVD_SET_SMEM(f->fn[md], kMaxSmem) != 0) return fail(PVD_ERR_CUDA, "cannot opt in to large shared memory (cols)");
                if (md == COL_CONV) {
                    int dev = 0, sms = 0, per = 0;
                    cudaGetDevice(&dev);
                    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, f->fn[md], f->fnNT[md], ((size_t)f->N * 16 + 4 * f->N) * sizeof(float2));
                    p->fnGrid[a] = sms * per;
                }
                if (f->pipe[md]) {
                    if (PVD_SET_SMEM(f->pipe[md], kMaxSmem) != 0) return fail(PVD_ERR_CUDA, "cannot opt in to large shared memory (pipe)");
                    int dev = 0, sms = 0, per = 0;
                    cudaGetDevice(&dev);
                    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, f->pipe[md], f->pipeNT[md],
                                                                  ((size_t)f->N * 2 * f->pipeW + 4 * f->N) * sizeof(float2));
                    p->pipeGrid[a][md] = sms * per;  // one resident wave: every CTA stays on its SM and loops
                }
            }
    if (p->fastRows && (PVD_SET_SMEM(p->fastRows->fwd, kMaxSmem) != 0 || PVD_SET_SMEM(p->fastRows->inv, kMaxSmem) != 0))
        return fail(PVD_ERR_CUDA, "cannot opt in to large shared memory (rows)");
    if (p->fastRows && p->fastRows->fwdPipe && p->fastRows->smemPipe <= kMaxSmem) {
        const FastRows* f = p->fastRows;
        // exact sizes: the kernels also hold a little static shared memory, so "the maximum" would overshoot
        if (PVD_SET_SMEM(f->fwdPipe, f->smemPipe) != 0 || PVD_SET_SMEM(f->invPipe, f->smemPipe) != 0)
            return fail(PVD_ERR_CUDA, "cannot opt in to large shared memory (rows pipe)");
        int dev = 0, sms = 0, per = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, f->fwdPipe, f->NT, f->smemPipe);
        p->rowPipeGrid[0] = sms * per;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, f->invPipe, f->NTinv, f->smemPipe);
        p->rowPipeGrid[1] = sms * per;
    }
    if (PVD_SET_SMEM(rows_fwd_kernel, kMaxSmem) != 0 || PVD_SET_SMEM(rows_inv_kernel, kMaxSmem) != 0 ||
        PVD_SET_SMEM(cols_kernel, kMaxSmem) != 0) {
        cudaGetLastError();
        return fail(PVD_ERR_CUDA, "cannot opt in to %zu bytes of dynamic shared memory", kMaxSmem);
    }
    return PVD_OK;
}

int pvd_plan_set_kernel(pvd_plan* p, const float* kernel, void* stream_) {
    if (!p || !kernel) return fail(PVD_ERR_INVALID, "null argument");
    if (!p->ws) return fail(PVD_ERR_STATE, "workspace not set");
    cudaStream_t stream = (cudaStream_t)stream_;
    p->kernel_set = false;
    const long long nk = (long long)p->k[0] * p->k[1] * p->k[2];
    cudaMemsetAsync(p->flag(), 0, sizeof(int), stream);
    PVD_LAUNCH(finite_check_kernel, dim3((unsigned)std::min<long long>(1024, (nk + 255) / 256)), dim3(256), 0, stream,
               kernel, nk, p->flag());
    PVD_CUDA_CHECK("finite_check_kernel");
    int bad = 0;
    cudaMemcpyAsync(&bad, p->flag(), sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) {
        PVD_CUDA_CHECK("finite check");
        return fail(PVD_ERR_CUDA, "finite check: stream synchronise failed");
    }
    if (bad) return fail(PVD_ERR_NONFINITE, "dose kernel contains non-finite values");
    if (p->algo == PVD_ALGO_DIRECT) {
        PVD_LAUNCH(flip_kernel_kernel, dim3(4), dim3(256), 0, stream, kernel, p->taps(), p->k[0], p->k[1], p->k[2]);
        PVD_CUDA_CHECK("flip_kernel_kernel");
        if (p->dcubic) {  // the cubic kernels take the flipped taps inside their argument struct: keep a host copy
            p->h_taps.assign((size_t)nk, 0.f);
            cudaMemcpyAsync(p->h_taps.data(), p->taps(), (size_t)nk * sizeof(float), cudaMemcpyDeviceToHost, stream);
            if (cudaStreamSynchronize(stream) != cudaSuccess) return fail(PVD_ERR_CUDA, "reading the flipped taps failed");
        }
        p->kernel_set = true;
        return PVD_OK;
    }
    const float* in[1] = {kernel};
    int rc = launch_rows_fwd(p, in, nullptr, 1, (long long)p->k[1] * p->k[2], p->k[2], p->ke, stream);
    if (rc) return rc;
    rc = launch_cols(p, 1, COL_FWD, p->buf(), p->buf(), 0, p->ke[0], p->ke[1], 0, p->m[1], 1.f, stream);
    if (rc) return rc;
    const float norm = (float)(1.0 / ((double)p->m[0] * p->m[1] * p->m[2]));
    rc = launch_cols(p, 0, COL_SPEC, p->buf(), p->spec(), 0, p->m[1], p->ke[0], 0, p->m[0], norm, stream);
    if (rc) return rc;
    p->kernel_set = true;
    return PVD_OK;
}

namespace {

// Plane-local forward passes (P1 z R2C + time-weighted sum, P2 y forward) of input planes [lo, hi) into the work buffer.
// `gain` (scale * rho_ref) rides on the input weights: the convolution is linear, so the last pass's epilogue is just
// dose = v / max(rho, rho_min) (3 instructions per voxel instead of 6).
int conv_forward(pvd_plan* p, const float* const* h_act, const float* h_weights, int T, float gain, int lo, int hi,
                 cudaStream_t stream, bool mark) {
    const double c = 8.0 * p->Nh;  // bytes of one half-spectrum row
    float wfold[PVD_MAX_T];
    const float* act[PVD_MAX_T];
    const long long plane = (long long)p->n[1] * p->n[2];
    for (int t = 0; t < T; ++t) {
        wfold[t] = (h_weights ? h_weights[t] : 1.f) * gain;
        act[t] = h_act[t] + (long long)lo * plane;
    }
    const int ext[3] = {hi - lo, p->n[1], p->n[2]};
    if (mark) p->mark(stream, "P1 rows_fwd (z R2C + time-weighted sum)", 4.0 * T * ext[0] * p->n[1] * p->n[2] + c * ext[0] * p->n[1]);
    int rc = launch_rows_fwd(p, act, wfold, T, plane, p->n[2], ext, stream, lo);
    if (rc) return rc;
    if (mark) p->mark(stream, "P2 cols y forward", c * ext[0] * ((double)p->n[1] + p->m[1]));
    return launch_cols(p, 1, COL_FWD, p->buf(), p->buf(), lo, hi - lo, p->n[1], 0, p->m[1], 1.f, stream);
}

int conv_finish(pvd_plan* p, const float* density, float rho_min, float rho_cut, float* dose, cudaStream_t stream, bool mark);
int conv_middle(pvd_plan* p, cudaStream_t stream, bool mark);
int conv_output(pvd_plan* p, const float* density, float rho_min, float rho_cut, float* dose, int lo, int hi, cudaStream_t stream, bool mark);

int check_execute_state(pvd_plan* p, const float* const* h_act, int T) {
    if (T < 1 || T > PVD_MAX_T) return fail(PVD_ERR_INVALID, "T=%d outside [1,%d]: pre-accumulate with pvd_weighted_sum", T, PVD_MAX_T);
    if (!p->ws) return fail(PVD_ERR_STATE, "workspace not set");
    if (!p->kernel_set) return fail(PVD_ERR_STATE, "dose kernel not set");
    for (int t = 0; t < T; ++t)
        if (!h_act[t]) return fail(PVD_ERR_INVALID, "activity pointer %d is null", t);
    return PVD_OK;
}

}  // namespace

int pvd_conv_execute(pvd_plan* p, const float* const* h_act, const float* h_weights, int T, const float* density,
                     float rho_ref, float rho_min, float rho_cut, float scale, float* dose, void* stream_) {
    if (!p || !h_act || !dose) return fail(PVD_ERR_INVALID, "null argument");
    if (int rc = check_execute_state(p, h_act, T)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (p->algo == PVD_ALGO_DIRECT) return execute_direct(p, h_act, h_weights, T, density, rho_ref, rho_min, rho_cut, scale, dose, stream);
    p->npass = 0;
    if (int rc = conv_forward(p, h_act, h_weights, T, scale * (density ? rho_ref : 1.f), 0, p->n[0], stream, true)) return rc;
    return conv_finish(p, density, rho_min, rho_cut, dose, stream, true);
}

int pvd_conv_execute_batch(pvd_plan* p, const float* const* h_act, const float* h_weights, int T, const float* const* h_density,
                           float rho_ref, float rho_min, float rho_cut, float scale, float* const* h_dose, int batch, void* stream_) {
    if (!p || !h_act || !h_dose) return fail(PVD_ERR_INVALID, "null argument");
    if (batch < 0) return fail(PVD_ERR_INVALID, "batch must not be negative");
    for (int b = 0; b < batch; ++b) {
        if (!h_dose[b]) return fail(PVD_ERR_INVALID, "null output pointer for volume %d", b);
        if (int rc = pvd_conv_execute(p, h_act + (size_t)b * T, h_weights, T, h_density ? h_density[b] : nullptr, rho_ref, rho_min, rho_cut,
                                      scale, h_dose[b], stream_))
            return rc;
    }
    return PVD_OK;
}

int pvd_conv_forward_planes(pvd_plan* p, const float* const* h_act, const float* h_weights, int T, float gain, int plane_lo,
                            int plane_hi, void* stream_) {
    if (!p || !h_act) return fail(PVD_ERR_INVALID, "null argument");
    if (int rc = check_execute_state(p, h_act, T)) return rc;
    if (p->algo != PVD_ALGO_FFT) return fail(PVD_ERR_UNSUPPORTED, "the split form exists for the FFT algorithm only");
    if (plane_lo < 0 || plane_hi > p->n[0] || plane_lo > plane_hi)
        return fail(PVD_ERR_INVALID, "plane range [%d, %d) outside [0, %d)", plane_lo, plane_hi, p->n[0]);
    if (plane_lo == plane_hi) return PVD_OK;
    return conv_forward(p, h_act, h_weights, T, gain, plane_lo, plane_hi, (cudaStream_t)stream_, false);
}

int pvd_conv_finish(pvd_plan* p, const float* density, float rho_min, float rho_cut, float* dose, void* stream_) {
    if (!p || !dose) return fail(PVD_ERR_INVALID, "null argument");
    if (!p->ws) return fail(PVD_ERR_STATE, "workspace not set");
    if (!p->kernel_set) return fail(PVD_ERR_STATE, "dose kernel not set");
    if (p->algo != PVD_ALGO_FFT) return fail(PVD_ERR_UNSUPPORTED, "the split form exists for the FFT algorithm only");
    return conv_finish(p, density, rho_min, rho_cut, dose, (cudaStream_t)stream_, false);
}

// ---- stream-ordered 32-bit flags (cuStreamWriteValue32 / cuStreamWaitValue32): cross-GPU signalling without a kernel
#ifndef PVD_EMULATE
namespace {
typedef CUresult (*StreamValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamValueFn get_stream_value_fn(const char* name) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
        return reinterpret_cast<StreamValueFn>(ptr);
    return nullptr;
}
}  // namespace
int pvd_stream_write_flag(void* d_flag, uint32_t value, void* stream) {
    static StreamValueFn fn = get_stream_value_fn("cuStreamWriteValue32");
    if (!d_flag) return fail(PVD_ERR_INVALID, "null argument");
    if (!fn) return fail(PVD_ERR_UNSUPPORTED, "cuStreamWriteValue32 not available from the driver");
    const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)d_flag, value, 0 /* CU_STREAM_WRITE_VALUE_DEFAULT */);
    if (r != CUDA_SUCCESS) return fail(PVD_ERR_CUDA, "cuStreamWriteValue32 failed with %d", (int)r);
    return PVD_OK;
}
int pvd_stream_wait_flag_geq(void* d_flag, uint32_t value, void* stream) {
    static StreamValueFn fn = get_stream_value_fn("cuStreamWaitValue32");
    if (!d_flag) return fail(PVD_ERR_INVALID, "null argument");
    if (!fn) return fail(PVD_ERR_UNSUPPORTED, "cuStreamWaitValue32 not available from the driver");
    const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)d_flag, value, 0 /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (r != CUDA_SUCCESS) return fail(PVD_ERR_CUDA, "cuStreamWaitValue32 failed with %d", (int)r);
    return PVD_OK;
}
int pvd_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
    if (!dst || !src) return fail(PVD_ERR_INVALID, "null argument");
    if (bytes == 0) return PVD_OK;
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(PVD_ERR_CUDA, "cudaMemcpyAsync (device to device): %s", cudaGetErrorString(e));
    return PVD_OK;
}
#else
int pvd_copy_async(void* dst, const void* src, size_t bytes, void*) {
    if (!dst || !src) return fail(PVD_ERR_INVALID, "null argument");
    memmove(dst, src, bytes);
    return PVD_OK;
}
int pvd_stream_write_flag(void* d_flag, uint32_t value, void*) {
    if (!d_flag) return fail(PVD_ERR_INVALID, "null argument");
    *reinterpret_cast<uint32_t*>(d_flag) = value;
    return PVD_OK;
}
int pvd_stream_wait_flag_geq(void* d_flag, uint32_t value, void*) {
    if (!d_flag) return fail(PVD_ERR_INVALID, "null argument");
    return *reinterpret_cast<uint32_t*>(d_flag) >= value ? PVD_OK : fail(PVD_ERR_STATE, "flag not reached (emulation has no concurrency)");
}
#endif

int pvd_plan_reserve_sms(pvd_plan* p, int n_sms) {
    if (!p) return fail(PVD_ERR_INVALID, "null argument");
    if (n_sms < 0 || (p->sms > 0 && n_sms >= p->sms)) return fail(PVD_ERR_INVALID, "cannot reserve %d of %d SMs", n_sms, p->sms);
    p->reserve_sms = n_sms;
    return PVD_OK;
}

int pvd_conv_middle(pvd_plan* p, void* stream_) {
    if (!p) return fail(PVD_ERR_INVALID, "null argument");
    if (!p->ws) return fail(PVD_ERR_STATE, "workspace not set");
    if (!p->kernel_set) return fail(PVD_ERR_STATE, "dose kernel not set");
    if (p->algo != PVD_ALGO_FFT) return fail(PVD_ERR_UNSUPPORTED, "the split form exists for the FFT algorithm only");
    return conv_middle(p, (cudaStream_t)stream_, false);
}

int pvd_conv_output_planes(pvd_plan* p, const float* density, float rho_min, float rho_cut, float* dose, int plane_lo, int plane_hi,
                           void* stream_) {
    if (!p || !dose) return fail(PVD_ERR_INVALID, "null argument");
    if (!p->ws) return fail(PVD_ERR_STATE, "workspace not set");
    if (p->algo != PVD_ALGO_FFT) return fail(PVD_ERR_UNSUPPORTED, "the split form exists for the FFT algorithm only");
    if (plane_lo < 0 || plane_hi > p->on[0] || plane_lo > plane_hi)
        return fail(PVD_ERR_INVALID, "output plane range [%d, %d) outside [0, %d)", plane_lo, plane_hi, p->on[0]);
    return conv_output(p, density, rho_min, rho_cut, dose, plane_lo, plane_hi, (cudaStream_t)stream_, false);
}

namespace {

// P3 (x forward * cached spectrum * x inverse) and P4 (y inverse): everything between the plane-local passes.
int conv_middle(pvd_plan* p, cudaStream_t stream, bool mark) {
    const double c = 8.0 * p->Nh;
    if (mark) p->mark(stream, "P3 cols x forward*spectrum*inverse", c * p->m[1] * ((double)p->n[0] + p->m[0] + p->on[0]));
    int rc = launch_cols(p, 0, COL_CONV, p->buf(), p->buf(), 0, p->m[1], p->n[0], p->olo[0], p->on[0], 1.f, stream);
    if (rc) return rc;
    if (mark) p->mark(stream, "P4 cols y inverse", c * p->on[0] * ((double)p->m[1] + p->on[1]));
    return launch_cols(p, 1, COL_INV, p->buf(), p->buf(), p->olo[0], p->on[0], p->m[1], p->olo[1], p->on[1], 1.f, stream);
}

// P5 (z complex-to-real, crop, density epilogue) of OUTPUT planes [lo, hi); density / dose point at plane 0 of the full volumes.
int conv_output(pvd_plan* p, const float* density, float rho_min, float rho_cut, float* dose, int lo, int hi, cudaStream_t stream,
                bool mark) {
    const double c = 8.0 * p->Nh;
    const float rho_ref = 1.f, scale = 1.f;  // folded into the input weights by conv_forward
    const int O0 = hi - lo;
    if (mark) p->mark(stream, "P5 rows_inv (z C2R + density + crop)",
                      c * O0 * p->on[1] + (density ? 8.0 : 4.0) * O0 * p->on[1] * p->on[2]);
    RowInvArgs a;
    memset(&a, 0, sizeof a);
    a.in_s0 = (long long)p->m[1] * p->Sz;
    a.in_s1 = p->Sz;
    const int xoff = p->olo[0] + lo;  // first transform plane of this range: folded into the base pointer
    float2* const in_base = p->buf() + (long long)xoff * a.in_s0;
    a.in = in_base;
    a.x_lo = 0;
    a.y_lo = p->olo[1];
    a.z_lo = p->olo[2];
    a.O0 = O0;
    a.O1 = p->on[1];
    a.O2 = p->on[2];
    a.out_s0 = (long long)p->on[1] * p->on[2];
    a.out_s1 = p->on[2];
    a.out = dose + (long long)lo * a.out_s0;
    a.den_s0 = a.out_s0;
    a.den_s1 = a.out_s1;
    a.density = density ? density + (long long)lo * a.den_s0 : nullptr;
    density = a.density;
    dose = a.out;
    a.rho_ref = rho_ref;
    a.rho_min = rho_min;
    a.rho_cut = rho_cut;
    a.scale = scale;
    a.M2 = p->m[2];
    a.Nh = p->Nh;
    a.Llog = p->rowLlog;
    a.tw = p->tw(2);
    a.st = p->st[2];
    a.dense = (a.x_lo == 0 && a.y_lo == 0 && a.in_s0 == (long long)a.O1 * a.in_s1 && a.out_s0 == (long long)a.O1 * a.out_s1 &&
               a.den_s0 == (long long)a.O1 * a.den_s1)
                  ? 1
                  : 0;
    a.plain_den = (density && rho_cut <= 0.f) ? 1 : 0;  // scale * rho_ref == 1 by the folding above

    a.vec4 = (p->on[2] % 4 == 0 && a.out_s0 % 4 == 0 && a.out_s1 % 4 == 0 && ((uintptr_t)dose & 15) == 0 &&
              (!density || (((uintptr_t)density & 15) == 0 && a.den_s0 % 4 == 0 && a.den_s1 % 4 == 0)))
                 ? 1
                 : 0;
    const long long nrows = (long long)O0 * p->on[1];
    const long long per = 2LL << p->rowLlog;
    if (nrows <= 0) return PVD_OK;
    if (p->fastRows && p->usePipe && p->rowPipeGrid[1] > 0 && nrows < 2000000000LL) {
        const FastRows* f = p->fastRows;
        const int grid = (int)std::min<long long>((nrows + 31) / 32, p->pgrid(p->rowPipeGrid[1]));
        a.use_tma = a.use_tma_den = 0;
        a.error_flag = p->flag() + 1;
#ifndef PVD_EMULATE
        // TMA staging of the half-spectrum rows (work buffer as 8-byte elements) and bulk L2 prefetch of the density rows
        const int lsc = (((f->N + 3) / 4) * 4 + 4) / 2;
        a.tma3d = 0;
        if (p->tmaRows && lsc <= 256 && lsc <= p->Sz && (a.dense || a.O1 % 32 == 0)) {
            const cuuint32_t estr[3] = {1, 1, 1};
            if (a.dense) {
                const cuuint64_t gdim[2] = {(cuuint64_t)p->Sz, (cuuint64_t)nrows};
                const cuuint64_t gstr[1] = {(cuuint64_t)a.in_s1 * 8};
                const cuuint32_t box[2] = {(cuuint32_t)lsc, 32};
                if (get_encode_tiled()(&a.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, in_base, gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                    a.use_tma = 1;
            } else {  // cropped output (x_lo / y_lo, M1 > O1): a tile never straddles x because 32 divides O1
                const cuuint64_t gdim[3] = {(cuuint64_t)p->Sz, (cuuint64_t)p->m[1], (cuuint64_t)(p->m[0] - xoff)};
                const cuuint64_t gstr[2] = {(cuuint64_t)a.in_s1 * 8, (cuuint64_t)a.in_s0 * 8};
                const cuuint32_t box[3] = {(cuuint32_t)lsc, 32, 1};
                if (get_encode_tiled()(&a.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, in_base, gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                    a.use_tma = 1;
                    a.tma3d = 1;
                }
            }
            if (density && a.vec4 && a.O2 / 2 <= 256) {
                const cuuint64_t gdim[2] = {(cuuint64_t)a.O2 / 2, (cuuint64_t)nrows};
                const cuuint64_t gstr[1] = {(cuuint64_t)a.den_s1 * 4};
                const cuuint32_t box[2] = {(cuuint32_t)(a.O2 / 2), 32};
                if (get_encode_tiled()(&a.tmap_den, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<float*>(density), gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                    a.use_tma_den = 1;
            }
        }
#endif
        PVD_LAUNCH_PDL(p->pdl, f->invPipe, dim3((unsigned)grid), dim3(f->NTinv), f->smemPipe, stream, a);
        PVD_CUDA_CHECK("rows_inv_pipe_kernel");
        if (mark) p->mark_end(stream);
        return PVD_OK;
    }
    if (p->fastRows) {
        const FastRows* f = p->fastRows;
        const size_t smem = ((size_t)f->N * 17 + 4 * f->N) * sizeof(float2);
        PVD_LAUNCH(f->inv, dim3((unsigned)((nrows + 31) / 32)), dim3(f->NTinv), smem, stream, a);
        PVD_CUDA_CHECK("rows_inv_fast_kernel");
        if (mark) p->mark_end(stream);
        return PVD_OK;
    }
    PVD_LAUNCH(rows_inv_kernel, dim3((unsigned)((nrows + per - 1) / per)), dim3(PVD_BLOCK), p->rowSmem, stream, a);
    PVD_CUDA_CHECK("rows_inv_kernel");
    if (mark) p->mark_end(stream);
    return PVD_OK;
}

int conv_finish(pvd_plan* p, const float* density, float rho_min, float rho_cut, float* dose, cudaStream_t stream, bool mark) {
    if (int rc = conv_middle(p, stream, mark)) return rc;
    return conv_output(p, density, rho_min, rho_cut, dose, 0, p->on[0], stream, mark);
}

}  // namespace

int pvd_plan_set_profiling(pvd_plan* p, int enable) {
    if (!p) return fail(PVD_ERR_INVALID, "null argument");
    if (enable && !p->ev_made) {
        for (int i = 0; i <= PVD_MAX_PASSES; ++i)
            if (cudaEventCreate(&p->ev[i]) != cudaSuccess) return fail(PVD_ERR_CUDA, "cudaEventCreate failed");
        p->ev_made = true;
    }
    p->prof = enable != 0;
    p->npass = 0;
    return PVD_OK;
}

int pvd_plan_get_pass_times(pvd_plan* p, float* ms, double* hbm_bytes, const char** names, int cap) {
    if (!p) return fail(PVD_ERR_INVALID, "null argument");
    if (!p->prof || p->npass == 0) return 0;
    if (cudaEventSynchronize(p->ev[p->npass]) != cudaSuccess) return fail(PVD_ERR_CUDA, "cudaEventSynchronize failed");
    const int n = std::min(p->npass, cap);
    for (int i = 0; i < n; ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, p->ev[i], p->ev[i + 1]);
        if (ms) ms[i] = t;
        if (hbm_bytes) hbm_bytes[i] = p->pass_bytes[i];
        if (names) names[i] = p->pass_names[i];
    }
    return n;
}

int pvd_plan_check_device_errors(pvd_plan* p, void* stream_) {
    if (!p) return fail(PVD_ERR_INVALID, "null argument");
    if (!p->ws) return fail(PVD_ERR_STATE, "workspace not set");
    cudaStream_t stream = (cudaStream_t)stream_;
    int flags[2] = {0, 0};
    cudaMemcpyAsync(flags, p->flag(), sizeof flags, cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) {
        PVD_CUDA_CHECK("device error flags");
        return fail(PVD_ERR_CUDA, "device error flags: stream synchronise failed");
    }
    const bool bad0 = p->algo == PVD_ALGO_DIRECT && flags[0] != 0, bad1 = flags[1] != 0;
    if (bad0 || bad1) {  // report once: the flag is sticky on the device, clear it so that the next execute starts clean
        cudaMemsetAsync(p->flag(), 0, sizeof flags, stream);
        cudaStreamSynchronize(stream);
        if (bad0) return fail(PVD_ERR_CUDA, "direct convolution: a TMA tile load never completed (flag %d); the dose map of that execute is invalid", flags[0]);
        return fail(PVD_ERR_CUDA, "FFT pass: a TMA tile load never completed (flag %d); the dose map of that execute is invalid", flags[1]);
    }
    return PVD_OK;
}

int pvd_plan_destroy(pvd_plan* p) {
    if (p && p->ev_made)
        for (int i = 0; i <= PVD_MAX_PASSES; ++i) cudaEventDestroy(p->ev[i]);
    delete p;
    return PVD_OK;
}

// ---------------------------------------------------------------------------------------------
static unsigned ew_grid(size_t n) { return (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16); }

int pvd_kernel_eval_radial(const pvd_radial_model* model, const double spacing[3], const int g[3], float* out, void* stream) {
    if (!model || !spacing || !g || !out) return fail(PVD_ERR_INVALID, "null argument");
    if (g[0] < 1 || g[1] < 1 || g[2] < 1) return fail(PVD_ERR_INVALID, "grid extents must be positive");
    if (model->n_beta < 0 || model->n_beta > PVD_RADIAL_MAX_TERMS || model->n_photon < 0 || model->n_photon > PVD_RADIAL_MAX_TERMS)
        return fail(PVD_ERR_INVALID, "at most %d beta and %d photon terms", PVD_RADIAL_MAX_TERMS, PVD_RADIAL_MAX_TERMS);
    RadialModel m;
    memset(&m, 0, sizeof m);
    m.nb = model->n_beta;
    m.np = model->n_photon;
    for (int i = 0; i < m.nb; ++i) {
        if (!(model->beta_range[i] > 0.0)) return fail(PVD_ERR_INVALID, "beta range %d must be positive", i);
        m.beta_range[i] = model->beta_range[i];
        m.beta_amp[i] = model->beta_amp[i];
    }
    for (int i = 0; i < m.np; ++i) {
        m.phot_mu[i] = model->phot_mu[i];
        m.phot_amp[i] = model->phot_amp[i];
    }
    m.scaling = model->scaling;
    for (int i = 0; i < 3; ++i) {
        if (!(spacing[i] > 0.0)) return fail(PVD_ERR_INVALID, "spacing must be positive");
        m.sp[i] = spacing[i];
        m.g[i] = g[i];
    }
    const size_t n = (size_t)g[0] * g[1] * g[2];
    PVD_LAUNCH(kernel_eval_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, m, out);
    PVD_CUDA_CHECK("kernel_eval_kernel");
    return PVD_OK;
}

static int fill_knots(const float* h_knots, int nk, Knots& k) {
    if (!h_knots || nk < 2 || nk > 32) return fail(PVD_ERR_INVALID, "need 2..32 (hu, rho) knots");
    k.nk = nk;
    for (int i = 0; i < nk; ++i) {
        k.hu[i] = h_knots[2 * i];
        k.rho[i] = h_knots[2 * i + 1];
        if (i && !(k.hu[i] > k.hu[i - 1])) return fail(PVD_ERR_INVALID, "HU knots must be strictly increasing");
    }
    return PVD_OK;
}

int pvd_hu_to_density_f32(const float* hu, const float* h_knots, int nk, float* rho, size_t n, void* stream) {
    Knots k;
    if (int rc = fill_knots(h_knots, nk, k)) return rc;
    if (!hu || !rho) return fail(PVD_ERR_INVALID, "null argument");
    if (n == 0) return PVD_OK;
    PVD_LAUNCH(hu_to_density_kernel<float>, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, hu, k, rho, n);
    PVD_CUDA_CHECK("hu_to_density_kernel");
    return PVD_OK;
}

int pvd_hu_to_density_i16(const int16_t* hu, const float* h_knots, int nk, float* rho, size_t n, void* stream) {
    Knots k;
    if (int rc = fill_knots(h_knots, nk, k)) return rc;
    if (!hu || !rho) return fail(PVD_ERR_INVALID, "null argument");
    if (n == 0) return PVD_OK;
    PVD_LAUNCH(hu_to_density_kernel<int16_t>, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, hu, k, rho, n);
    PVD_CUDA_CHECK("hu_to_density_kernel");
    return PVD_OK;
}

int pvd_i16_to_f32(const void* d_in, int is_unsigned, float slope, float intercept, float* d_out, size_t n, void* stream) {
    if (!d_in || !d_out) return fail(PVD_ERR_INVALID, "null argument");
    if (n == 0) return PVD_OK;
    if (is_unsigned)
        PVD_LAUNCH(rescale_to_f32_kernel<unsigned short>, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream,
                   (const unsigned short*)d_in, slope, intercept, d_out, n);
    else
        PVD_LAUNCH(rescale_to_f32_kernel<short>, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, (const short*)d_in, slope,
                   intercept, d_out, n);
    PVD_CUDA_CHECK("rescale_to_f32_kernel");
    return PVD_OK;
}

int pvd_weighted_combine(const float* const* h_vol, int T, const float* h_W, float* const* h_out, int J, size_t n, void* stream) {
    if (!h_vol || !h_out) return fail(PVD_ERR_INVALID, "null argument");
    if (T < 1 || T > PVD_MAX_T) return fail(PVD_ERR_INVALID, "T=%d outside [1,%d]", T, PVD_MAX_T);
    if (J < 1 || J > kMaxJ) return fail(PVD_ERR_INVALID, "J=%d outside [1,%d]", J, kMaxJ);
    WcombArgs a;
    memset(&a, 0, sizeof a);
    a.T = T;
    a.J = J;
    bool vec = true;
    for (int t = 0; t < T; ++t) {
        if (!h_vol[t]) return fail(PVD_ERR_INVALID, "volume pointer %d is null", t);
        a.v[t] = h_vol[t];
        vec = vec && ((uintptr_t)h_vol[t] % 16 == 0);
    }
    for (int j = 0; j < J; ++j) {
        if (!h_out[j]) return fail(PVD_ERR_INVALID, "output pointer %d is null", j);
        a.o[j] = h_out[j];
        vec = vec && ((uintptr_t)h_out[j] % 16 == 0);
        for (int t = 0; t < T; ++t) {
            const float w = h_W ? h_W[(size_t)j * T + t] : 1.f;
            a.w[j][t] = w;
            if (w != 0.f) a.used |= 1u << t;
        }
    }
    if (n == 0) return PVD_OK;
    const size_t ngroups = (n + 3) / 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (J == 1)
        PVD_LAUNCH(weighted_combine_kernel<1>, dim3(ew_grid(ngroups)), dim3(256), 0, st, a, n, (int)vec);
    else if (J <= 4)
        PVD_LAUNCH(weighted_combine_kernel<4>, dim3(ew_grid(ngroups)), dim3(256), 0, st, a, n, (int)vec);
    else if (J <= 8)
        PVD_LAUNCH(weighted_combine_kernel<8>, dim3(ew_grid(ngroups)), dim3(256), 0, st, a, n, (int)vec);
    else
        PVD_LAUNCH(weighted_combine_kernel<16>, dim3(ew_grid(ngroups)), dim3(256), 0, st, a, n, (int)vec);
    PVD_CUDA_CHECK("weighted_combine_kernel");
    return PVD_OK;
}

int pvd_weighted_sum(const float* const* h_vol, const float* h_weights, int T, float* out, size_t n, void* stream) {
    if (!out) return fail(PVD_ERR_INVALID, "null argument");
    float* const outs[1] = {out};
    return pvd_weighted_combine(h_vol, T, h_weights, outs, 1, n, stream);
}

int pvd_monoexp_integral(const float* A0, const float* lambda, float t_limit, float* out, size_t n, void* stream) {
    if (!A0 || !lambda || !out) return fail(PVD_ERR_INVALID, "null argument");
    if (n == 0) return PVD_OK;
    PVD_LAUNCH(monoexp_integral_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, A0, lambda, t_limit, out, n);
    PVD_CUDA_CHECK("monoexp_integral_kernel");
    return PVD_OK;
}

int pvd_density_scale(const float* dose, const float* density, float rho_ref, float rho_min, float rho_cut, float scale,
                      float* out, size_t n, void* stream) {
    if (!dose || !density || !out) return fail(PVD_ERR_INVALID, "null argument");
    if (n == 0) return PVD_OK;
    PVD_LAUNCH(density_scale_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, dose, density, rho_ref, rho_min,
               rho_cut, scale, out, n);
    PVD_CUDA_CHECK("density_scale_kernel");
    return PVD_OK;
}

// ---------------------------------------------------------------------------------------------
// steps either side of the convolution (analysis.cuh)
int pvd_monoexp_fit(const float* const* h_vol, const float* h_times, const float* h_weights, int T, float lambda0,
                    float t_limit, float* A0, float* lambda, float* accumulated, size_t n, void* stream) {
    if (!h_vol || !h_times) return fail(PVD_ERR_INVALID, "null argument");
    if (T < 2 || T > PVD_MAX_T) return fail(PVD_ERR_INVALID, "T=%d outside [2,%d]: a two-parameter fit needs two time points", T, PVD_MAX_T);
    if (!(lambda0 > 0.f)) return fail(PVD_ERR_INVALID, "lambda0 must be positive");
    FitArgs a;
    memset(&a, 0, sizeof a);
    a.T = T;
    a.lam0 = lambda0;
    a.tlim = t_limit;
    for (int t = 0; t < T; ++t) {
        if (!h_vol[t]) return fail(PVD_ERR_INVALID, "volume pointer %d is null", t);
        a.v[t] = h_vol[t];
        a.t[t] = h_times[t];
        a.w[t] = h_weights ? h_weights[t] : 1.f;
    }
    if (n == 0) return PVD_OK;
    const dim3 grid(ew_grid(n)), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    switch (T) {
        case 2: PVD_LAUNCH(monoexp_fit_kernel<2>, grid, block, 0, st, a, A0, lambda, accumulated, n); break;
        case 3: PVD_LAUNCH(monoexp_fit_kernel<3>, grid, block, 0, st, a, A0, lambda, accumulated, n); break;
        case 4: PVD_LAUNCH(monoexp_fit_kernel<4>, grid, block, 0, st, a, A0, lambda, accumulated, n); break;
        case 5: PVD_LAUNCH(monoexp_fit_kernel<5>, grid, block, 0, st, a, A0, lambda, accumulated, n); break;
        case 6: PVD_LAUNCH(monoexp_fit_kernel<6>, grid, block, 0, st, a, A0, lambda, accumulated, n); break;
        default: PVD_LAUNCH(monoexp_fit_kernel<0>, grid, block, 0, st, a, A0, lambda, accumulated, n); break;
    }
    PVD_CUDA_CHECK("monoexp_fit_kernel");
    return PVD_OK;
}

int pvd_ct_prepare(const float* hu, const int n[3], float metal_threshold, const float* h_knots, int nk,
                   const float* h_ranges, int nr, float* corrected, float* rho, unsigned char* labels, void* stream) {
    if (!hu || !n) return fail(PVD_ERR_INVALID, "null argument");
    if (n[0] < 1 || n[1] < 1 || n[2] < 1) return fail(PVD_ERR_INVALID, "extents must be positive");
    if (hu == corrected) return fail(PVD_ERR_INVALID, "the artifact fill reads neighbours: corrected must not alias hu");
    CtArgs a;
    memset(&a, 0, sizeof a);
    a.hu = hu;
    a.n0 = n[0];
    a.n1 = n[1];
    a.n2 = n[2];
    a.metal_thr = metal_threshold;
    {  // scipy.ndimage._filters._gaussian_kernel1d(sigma = 1, order = 0, radius = 4): exp(-x^2/2) normalised to sum 1
        double gsum = 0.0, gw[5];
        for (int d = 0; d <= 4; ++d) {
            gw[d] = std::exp(-0.5 * d * d);
            gsum += (d ? 2.0 : 1.0) * gw[d];
        }
        for (int d = 0; d <= 4; ++d) a.g[d] = (float)(gw[d] / gsum);
    }
    if (rho) {
        Knots k;
        if (int rc = fill_knots(h_knots, nk, k)) return rc;
        a.nseg = nk - 1;
        a.rho0 = k.rho[0];
        for (int j = 0; j + 1 < nk; ++j) {
            a.seg_hu[j] = k.hu[j];
            a.seg_len[j] = k.hu[j + 1] - k.hu[j];
            a.seg_slope[j] = (k.rho[j + 1] - k.rho[j]) / a.seg_len[j];
        }
    }
    for (int c = 0; c < 8; ++c) {  // empty ranges: never match
        a.lo[c] = INFINITY;
        a.hi[c] = -INFINITY;
    }
    if (labels) {
        if (!h_ranges || nr < 1 || nr > 8) return fail(PVD_ERR_INVALID, "labels need 1..8 (lo, hi) HU ranges");
        a.nr = nr;
        for (int c = 0; c < nr; ++c) {
            a.lo[c] = h_ranges[2 * c];
            a.hi[c] = h_ranges[2 * c + 1];
        }
    }
    a.corrected = corrected;
    a.rho = rho;
    a.labels = labels;
    const size_t nv = (size_t)n[0] * n[1] * n[2];
    const int vec = (((uintptr_t)hu | (uintptr_t)corrected | (uintptr_t)rho) & 15) == 0 && ((uintptr_t)labels & 3) == 0;
    const unsigned grid = (unsigned)std::min<size_t>((nv + 4 * kCtThreads - 1) / (4 * kCtThreads), 148 * 8);
    // merged interval table: cuts = density knots + range starts + the float after every range end
    {
        std::vector<float> cuts;
        Knots k;
        k.nk = 0;
        if (rho) {
            fill_knots(h_knots, nk, k);
            for (int j = 0; j < nk; ++j) cuts.push_back(k.hu[j]);
            a.knot_lo = k.hu[0];
            a.knot_hi = k.hu[nk - 1];
        }
        for (int c = 0; c < a.nr; ++c)
            if (a.lo[c] <= a.hi[c]) {
                cuts.push_back(a.lo[c]);
                if (a.hi[c] < INFINITY) cuts.push_back(std::nextafterf(a.hi[c], INFINITY));
            }
        std::sort(cuts.begin(), cuts.end());
        cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
        const int m = (int)cuts.size();
        a.lut_log = m <= 31 ? 5 : (m <= 63 ? 6 : 0);
        bool finite = true;
        for (float c : cuts) finite = finite && !std::isnan(c);
        if (!finite) a.lut_log = 0;
        if (a.lut_log) {
            const int slots = (1 << a.lut_log) - 1;
            for (int i = 0; i < slots; ++i) a.thr[i] = i < m ? cuts[i] : INFINITY;
            for (int e = 0; e <= m; ++e) {
                // a representative of the interval: its first value (entry 0 = below every cut / NaN: first knot, no class)
                const float r = e == 0 ? -INFINITY : cuts[e - 1];
                float a0 = 0.f, r0 = 0.f, sl = 0.f;
                if (k.nk) {
                    int j = 0;
                    while (j + 1 < k.nk && r >= k.hu[j + 1]) ++j;
                    a0 = k.hu[j];
                    r0 = k.rho[j];
                    if (r >= k.hu[0] && j + 1 < k.nk) sl = (k.rho[j + 1] - k.rho[j]) / (k.hu[j + 1] - k.hu[j]);
                }
                unsigned mask = 0;
                if (e > 0)
                    for (int c = 0; c < a.nr; ++c)
                        if (r >= a.lo[c] && r <= a.hi[c]) mask |= 1u << c;
                a.ent[e][0] = a0;
                a.ent[e][1] = r0;
                a.ent[e][2] = sl;
                memcpy(&a.ent[e][3], &mask, 4);
            }
            for (int e = m + 1; e < (1 << a.lut_log); ++e) memcpy(a.ent[e], a.ent[m], sizeof a.ent[e]);
        }
    }
    if (a.lut_log == 5)
        PVD_LAUNCH(ct_prepare_kernel<5>, dim3(grid), dim3(kCtThreads), 0, (cudaStream_t)stream, a, vec);
    else if (a.lut_log == 6)
        PVD_LAUNCH(ct_prepare_kernel<6>, dim3(grid), dim3(kCtThreads), 0, (cudaStream_t)stream, a, vec);
    else
        PVD_LAUNCH(ct_prepare_kernel<0>, dim3(grid), dim3(kCtThreads), 0, (cudaStream_t)stream, a, vec);
    PVD_CUDA_CHECK("ct_prepare_kernel");
    return PVD_OK;
}

// 4-voxel groups the DVH kernels may read with 128-bit (dose, float mask) / 32-bit (uint8 mask) loads
static size_t dvh_vec_groups(const float* dose, const void* mask, int mask_is_f32, size_t n) {
    const bool ok = ((uintptr_t)dose & 15) == 0 && ((uintptr_t)mask & (mask_is_f32 ? 15 : 3)) == 0;
    return ok ? n / 4 : 0;
}

int pvd_roi_minmax(const float* dose, const void* mask, int mask_is_f32, size_t n, void* d_scratch16, float* h_min,
                   float* h_max, unsigned long long* h_count, void* stream) {
    if (!dose || !mask || !d_scratch16 || !h_min || !h_max || !h_count) return fail(PVD_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    RoiStats init;
    init.min_key = 0xFFFFFFFFu;
    init.max_key = 0u;
    init.count = 0ull;
    RoiStats* d = reinterpret_cast<RoiStats*>(d_scratch16);
    if (cudaMemcpyAsync(d, &init, sizeof init, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(PVD_ERR_CUDA, "cudaMemcpyAsync failed");
    if (n) {
        const size_t n4 = dvh_vec_groups(dose, mask, mask_is_f32, n);
        if (mask_is_f32)
            PVD_LAUNCH(roi_minmax_kernel<float>, dim3(ew_grid(n / 4 + 1)), dim3(256), 0, st, dose, (const float*)mask, n, n4, d);
        else
            PVD_LAUNCH(roi_minmax_kernel<unsigned char>, dim3(ew_grid(n / 4 + 1)), dim3(256), 0, st, dose, (const unsigned char*)mask, n, n4, d);
        PVD_CUDA_CHECK("roi_minmax_kernel");
    }
    RoiStats h;
    if (cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
        return fail(PVD_ERR_CUDA, "reading the ROI statistics failed");
    auto key2f = [](unsigned k) {
        const unsigned u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
        float f;
        memcpy(&f, &u, 4);
        return f;
    };
    *h_count = h.count;
    *h_min = h.count ? key2f(h.min_key) : 0.f;
    *h_max = h.count ? key2f(h.max_key) : 0.f;
    return PVD_OK;
}

int pvd_dvh_histogram(const float* dose, const void* mask, int mask_is_f32, size_t n, const float* d_edges, int bins,
                      float first_edge, float last_edge, unsigned long long* d_hist, void* stream) {
    if (!dose || !mask || !d_edges || !d_hist) return fail(PVD_ERR_INVALID, "null argument");
    if (bins < 1) return fail(PVD_ERR_INVALID, "bins must be positive");
    if (!(last_edge > first_edge)) return fail(PVD_ERR_INVALID, "last_edge must exceed first_edge");
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(d_hist, 0, (size_t)bins * sizeof(unsigned long long), st) != cudaSuccess) return fail(PVD_ERR_CUDA, "cudaMemsetAsync failed");
    if (n == 0) return PVD_OK;
    const size_t n4 = dvh_vec_groups(dose, mask, mask_is_f32, n);
#ifndef PVD_EMULATE
    const size_t lanes_min_n = (size_t)1 << 22;  // lane-private bins: one 1024-thread CTA per SM (large volumes)
    const int lanes_threads = kDvhLaneThreads;
#else
    const size_t lanes_min_n = 2048;             // the emulator runs one host thread per CUDA thread
    const int lanes_threads = 64;
#endif
    if (bins <= kDvhLaneBins && n >= lanes_min_n) {
        const size_t smem = ((size_t)bins * 32 + bins + 1) * sizeof(unsigned);
        int sms = 2;
#ifndef PVD_EMULATE
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#endif
        if (mask_is_f32) {
            if (PVD_SET_SMEM(dvh_hist_lanes_kernel<float>, smem) != 0) return fail(PVD_ERR_CUDA, "cannot opt in to large shared memory (dvh)");
            PVD_LAUNCH(dvh_hist_lanes_kernel<float>, dim3((unsigned)sms), dim3(lanes_threads), smem, st, dose, (const float*)mask, n, n4,
                       d_edges, bins, first_edge, last_edge, d_hist);
        } else {
            if (PVD_SET_SMEM(dvh_hist_lanes_kernel<unsigned char>, smem) != 0)
                return fail(PVD_ERR_CUDA, "cannot opt in to large shared memory (dvh)");
            PVD_LAUNCH(dvh_hist_lanes_kernel<unsigned char>, dim3((unsigned)sms), dim3(lanes_threads), smem, st, dose,
                       (const unsigned char*)mask, n, n4, d_edges, bins, first_edge, last_edge, d_hist);
        }
        PVD_CUDA_CHECK("dvh_hist_lanes_kernel");
        return PVD_OK;
    }
    if (mask_is_f32)
        PVD_LAUNCH(dvh_hist_kernel<float>, dim3(ew_grid(n / 4 + 1)), dim3(256), 0, st, dose, (const float*)mask, n, n4, d_edges,
                   bins, first_edge, last_edge, d_hist);
    else
        PVD_LAUNCH(dvh_hist_kernel<unsigned char>, dim3(ew_grid(n / 4 + 1)), dim3(256), 0, st, dose, (const unsigned char*)mask, n,
                   n4, d_edges, bins, first_edge, last_edge, d_hist);
    PVD_CUDA_CHECK("dvh_hist_kernel");
    return PVD_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer staging (host_stage.cuh)
#ifdef PVD_EMULATE
struct pvd_stager { int threads; };
int pvd_stager_create(pvd_stager** out, int threads, size_t, int) {
    if (!out) return fail(PVD_ERR_INVALID, "null argument");
    *out = new pvd_stager{threads > 0 ? threads : 1};
    return PVD_OK;
}
int pvd_stager_destroy(pvd_stager* s) {
    delete s;
    return PVD_OK;
}
int pvd_stage_h2d(pvd_stager* s, const void* h_src, int dt, void* d_dst, size_t n, void*) {
    if (!s || !h_src || !d_dst) return fail(PVD_ERR_INVALID, "null argument");
    if (dt == PVD_DTYPE_F64) {
        for (size_t i = 0; i < n; ++i) ((float*)d_dst)[i] = (float)((const double*)h_src)[i];
    } else {
        memcpy(d_dst, h_src, n * (dt == PVD_DTYPE_F32 ? 4 : 2));
    }
    return PVD_OK;
}
int pvd_stage_d2h(pvd_stager* s, const float* d_src, void* h_dst, int dt, size_t n, void*) {
    if (!s || !d_src || !h_dst) return fail(PVD_ERR_INVALID, "null argument");
    if (dt == PVD_DTYPE_F64) {
        for (size_t i = 0; i < n; ++i) ((double*)h_dst)[i] = (double)d_src[i];
    } else if (dt == PVD_DTYPE_F32) {
        memcpy(h_dst, d_src, n * 4);
    } else {
        return fail(PVD_ERR_INVALID, "dose maps leave as float32 or float64");
    }
    return PVD_OK;
}
#else
struct pvd_stager {
    pvd::Stager impl;
    pvd_stager(int t, size_t c, int r) : impl(t, c, r) {}
};
int pvd_stager_create(pvd_stager** out, int threads, size_t chunk_bytes, int ring_chunks) {
    if (!out) return fail(PVD_ERR_INVALID, "null argument");
    if (threads <= 0) {
        const unsigned hw = std::thread::hardware_concurrency();
        threads = (int)std::min(8u, hw ? hw : 4u);
#ifdef __linux__
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0) threads = std::max(1, std::min(threads, CPU_COUNT(&set)));
#endif
    }
    if (chunk_bytes == 0) chunk_bytes = (size_t)4 << 20;
    if (ring_chunks <= 0) ring_chunks = 2 * threads + 2;
    if (threads > 64 || chunk_bytes < 4096 || chunk_bytes % 256 != 0 || ring_chunks < 2 || ring_chunks > 256)
        return fail(PVD_ERR_INVALID, "stager: threads <= 64, chunk a multiple of 256 bytes >= 4096, 2..256 ring chunks");
    pvd_stager* s = new pvd_stager(threads, chunk_bytes, ring_chunks);
    const cudaError_t e = s->impl.init();
    if (e != cudaSuccess) {
        delete s;
        cudaGetLastError();
        return fail(PVD_ERR_CUDA, "stager: %s", cudaGetErrorString(e));
    }
    *out = s;
    return PVD_OK;
}
int pvd_stager_destroy(pvd_stager* s) {
    delete s;
    return PVD_OK;
}
int pvd_stage_h2d(pvd_stager* s, const void* h_src, int dt, void* d_dst, size_t n, void* stream) {
    if (!s || !h_src || !d_dst) return fail(PVD_ERR_INVALID, "null argument");
    if (dt < PVD_DTYPE_F32 || dt > PVD_DTYPE_U16) return fail(PVD_ERR_INVALID, "unknown host dtype %d", dt);
    const cudaError_t e = s->impl.h2d(h_src, dt, d_dst, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(PVD_ERR_CUDA, "staged host-to-device copy: %s", cudaGetErrorString(e));
    return PVD_OK;
}
int pvd_stage_d2h(pvd_stager* s, const float* d_src, void* h_dst, int dt, size_t n, void* stream) {
    if (!s || !d_src || !h_dst) return fail(PVD_ERR_INVALID, "null argument");
    if (dt != PVD_DTYPE_F32 && dt != PVD_DTYPE_F64) return fail(PVD_ERR_INVALID, "dose maps leave as float32 or float64");
    const cudaError_t e = s->impl.d2h(d_src, h_dst, dt, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(PVD_ERR_CUDA, "staged device-to-host copy: %s", cudaGetErrorString(e));
    return PVD_OK;
}
#endif

}  // extern "C"
