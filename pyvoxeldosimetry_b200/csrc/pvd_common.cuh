// Common definitions for the pvdose CUDA sources.
//
// The product is compiled by nvcc for sm_100a.  The same sources also compile with plain g++
// when PVD_EMULATE is defined: a tiny SIMT emulator (one std::thread per CUDA thread, a
// std::barrier for __syncthreads) that exists ONLY so the index logic of every kernel can be
// exercised by the CPU test-suite in a container that has no GPU.  The emulated library is
// test infrastructure (built into tests/_emu/, never loaded by the package).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cmath>

#ifdef PVD_EMULATE
// ------------------------------------------------------------------ CPU emulation of SIMT
#include <algorithm>
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>
#include <cstdlib>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define PVD_UNROLL

namespace pvd_emu {
struct Ctx {
    dim3 tid, bid, bdim, gdim;
    std::barrier<>* bar = nullptr;
    char* smem = nullptr;
    std::barrier<>* warp_bar = nullptr;  // the 32 (or fewer) threads of this thread's warp
    unsigned* warp_buf = nullptr;        // 32-word exchange buffer of the warp (ballot / shuffle emulation)
    unsigned lane = 0, warp_lanes = 32;
};
inline thread_local Ctx ctx;
}  // namespace pvd_emu
#define threadIdx (pvd_emu::ctx.tid)
#define blockIdx (pvd_emu::ctx.bid)
#define blockDim (pvd_emu::ctx.bdim)
#define gridDim (pvd_emu::ctx.gdim)
static inline void __syncthreads() { pvd_emu::ctx.bar->arrive_and_wait(); }
// Warp-level primitives (full-mask, convergent use only): every lane publishes its word, the warp meets, reads, meets.
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { pvd_emu::ctx.warp_bar->arrive_and_wait(); }
static inline unsigned pvd_emu_exchange(unsigned mine, unsigned src_lane) {
    pvd_emu::Ctx& c = pvd_emu::ctx;
    c.warp_buf[c.lane] = mine;
    c.warp_bar->arrive_and_wait();
    const unsigned v = c.warp_buf[src_lane < c.warp_lanes ? src_lane : c.lane];
    c.warp_bar->arrive_and_wait();
    return v;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    pvd_emu::Ctx& c = pvd_emu::ctx;
    c.warp_buf[c.lane] = pred ? 1u : 0u;
    c.warp_bar->arrive_and_wait();
    unsigned m = 0;
    for (unsigned l = 0; l < c.warp_lanes; ++l) m |= c.warp_buf[l] << l;
    c.warp_bar->arrive_and_wait();
    return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    unsigned u;
    std::memcpy(&u, &v, 4);
    u = pvd_emu_exchange(u, pvd_emu::ctx.lane ^ (unsigned)lane_mask);
    std::memcpy(&v, &u, 4);
    return v;
}
static inline float __shfl_sync(unsigned, float v, int src) {
    unsigned u;
    std::memcpy(&u, &v, 4);
    u = pvd_emu_exchange(u, (unsigned)src);
    std::memcpy(&v, &u, 4);
    return v;
}
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __float_as_uint(float f) {
    unsigned u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float __uint_as_float(unsigned u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline void sincospi(double x, double* s, double* c) {
    *s = std::sin(M_PI * x);
    *c = std::cos(M_PI * x);
}
#define PVD_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(pvd_emu::ctx.smem)

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 3; return 0; }  // 3 "SMs": tiles > CTAs in tests
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 1; return 0; }
typedef void* cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t*) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }

namespace pvd_emu {
template <class K, class... Args>
void launch(K kernel, dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
    const unsigned T = block.x * block.y * block.z;
    std::vector<char> smem(smem_bytes + 64);
    std::barrier<> bar((std::ptrdiff_t)T);
    const unsigned nwarps = (T + 31) / 32;
    std::vector<std::unique_ptr<std::barrier<>>> wbars;
    for (unsigned w = 0; w < nwarps; ++w) wbars.emplace_back(new std::barrier<>((std::ptrdiff_t)std::min(32u, T - 32 * w)));
    std::vector<unsigned> wbuf(32 * nwarps);
    auto worker = [&](unsigned t) {
        Ctx& c = ctx;
        c.bdim = block;
        c.gdim = grid;
        c.bar = &bar;
        c.smem = smem.data();
        c.warp_bar = wbars[t / 32].get();
        c.warp_buf = wbuf.data() + 32 * (t / 32);
        c.lane = t % 32;
        c.warp_lanes = std::min(32u, T - 32 * (t / 32));
        c.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    c.bid = dim3(bx, by, bz);
                    kernel(args...);
                    bar.arrive_and_wait();  // block boundary: smem is reused by the next block
                }
    };
    std::vector<std::thread> th;
    th.reserve(T);
    for (unsigned t = 1; t < T; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& x : th) x.join();
}
}  // namespace pvd_emu
#define PVD_LAUNCH(kernel, grid, block, smem, stream, ...) pvd_emu::launch(kernel, grid, block, smem, __VA_ARGS__)
#define PVD_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, arg) pvd_emu::launch(kernel, grid, block, smem, arg)
static inline void grid_dep_wait() {}
static inline void grid_dep_launch() {}
#define PVD_SET_SMEM(kernel, bytes) (0)
static constexpr int PVD_BLOCK = 32;  // small blocks keep the emulator fast; kernels are block-size agnostic

#else
// ------------------------------------------------------------------ real CUDA
#include <cuda_runtime.h>
#define PVD_UNROLL _Pragma("unroll")
#define PVD_DYN_SMEM(T, name)                                         \
    extern __shared__ __align__(16) unsigned char pvd_dyn_smem_raw[]; \
    T* name = reinterpret_cast<T*>(pvd_dyn_smem_raw)
#define PVD_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
// Programmatic dependent launch (sm_90+): the kernel may be scheduled while its predecessor in the stream drains, so
// that launch latency, CTA start-up and the twiddle-table prologue overlap the predecessor's tail.  Every kernel
// launched this way calls grid_dep_wait() before it touches data the predecessor may have written (the wait returns
// once the predecessor grid has completed and its writes are visible) and grid_dep_launch() to let ITS successor in.
template <class Arg>
static inline cudaError_t pvd_launch_pdl(bool pdl, void (*kernel)(const Arg), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t stream, const Arg& arg) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, arg);
}
#define PVD_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, arg) pvd_launch_pdl(pdl, kernel, grid, block, smem, stream, arg)
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define PVD_SET_SMEM(kernel, bytes) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
static constexpr int PVD_BLOCK = 256;
#endif

namespace pvd {

static constexpr int kMaxStages = 16;
static constexpr int kMaxT = 16;  // activity volumes fused into one first-pass load

struct Stages {
    int n;
    int radix[kMaxStages];
};

// Complex add / subtract.  On sm_100a a float2 add is ONE packed instruction (add.f32x2 -> FADD2; the subtraction is
// fma.f32x2 with the constant pair (-1, -1), exactly rounded like a - b): the butterflies are mostly complex
// adds, so this removes a third of the floating-point issue slots.  Results are bit-identical to the scalar form.
#ifndef PVD_F32X2
#define PVD_F32X2 1
#endif
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && !defined(PVD_EMULATE) && PVD_F32X2
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && !defined(PVD_EMULATE) && PVD_F32X2
    return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// packed fused multiply-add: (a.x * b.x + c.x, a.y * b.y + c.y), each half rounded once like fmaf
__host__ __device__ __forceinline__ float2 cfma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && !defined(PVD_EMULATE) && PVD_F32X2
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

}  // namespace pvd
