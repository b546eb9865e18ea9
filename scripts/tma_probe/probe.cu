// Standalone TMA 3-D box-load probe: which descriptor / coordinate combinations does the B200 accept?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, int box, float* out, int fence_kind) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* tile = reinterpret_cast<float*>(raw);
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
        if (fence_kind == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(box * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"((unsigned)__cvta_generic_to_shared(tile)), "l"(&tmap), "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    unsigned ok = 0; long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
        if (clock64() - t0 > 2000000000LL) break;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < box; i += blockDim.x) out[i] = ok ? tile[i] : -777.f;
}
int main() {
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) { printf("no entry point\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)ptr;
    const int n0 = 16, n1 = 16, n2 = 64;
    std::vector<float> h(n0 * n1 * n2);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 1 << 20);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Case { int b2, b1, b0, c0, c1, c2, fence; const char* name; } cases[] = {
        {32, 4, 4, 0, 0, 0, 0, "box 32x4x4 inside, fence.mbarrier_init"},
        {32, 4, 4, 0, 0, 0, 1, "box 32x4x4 inside, fence.proxy.async"},
        {32, 4, 4, 0, -1, -1, 0, "inner 0, outer -1,-1"},
        {32, 4, 4, -4, -1, -1, 0, "inner -4 (aligned), outer -1,-1"},
        {72, 10, 10, -4, -1, -2, 0, "box 72x10x10 (> inner dim 64), inner -4"},
        {72, 14, 14, 60, 12, 13, 0, "box 72x14x14 overhanging the far corner"},
        {32, 4, 4, 2, 0, 0, 0, "inner +2 (8 bytes, misaligned?)"},
        {32, 4, 4, -1, 0, 0, 0, "inner -1 (misaligned)"},
    };
    for (auto& c : cases) {
        CUtensorMap tm; memset(&tm, 0, sizeof tm);
        cuuint64_t gdim[3] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0};
        cuuint64_t gstr[2] = {(cuuint64_t)n2 * 4, (cuuint64_t)n2 * n1 * 4};
        cuuint32_t box[3] = {(cuuint32_t)c.b2, (cuuint32_t)c.b1, (cuuint32_t)c.b0};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int nbox = c.b0 * c.b1 * c.b2;
        if (r != CUDA_SUCCESS) { printf("%-50s encode failed %d\n", c.name, (int)r); continue; }
        k<<<1, 128, nbox * 4 + 256>>>(tm, c.c0, c.c1, c.c2, nbox, out, c.fence);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-50s KERNEL ERROR: %s\n", c.name, cudaGetErrorString(e)); return 0; }
        std::vector<float> o(nbox); cudaMemcpy(o.data(), out, nbox * 4, cudaMemcpyDeviceToHost);
        // expected value at box element (i0,i1,i2): tensor[c2+i0][c1+i1][c0+i2] or 0
        int bad = 0;
        for (int i0 = 0; i0 < c.b0; ++i0) for (int i1 = 0; i1 < c.b1; ++i1) for (int i2 = 0; i2 < c.b2; ++i2) {
            int g0 = c.c2 + i0, g1 = c.c1 + i1, g2 = c.c0 + i2;
            float want = (g0 >= 0 && g0 < n0 && g1 >= 0 && g1 < n1 && g2 >= 0 && g2 < n2) ? h[(g0 * n1 + g1) * n2 + g2] : 0.f;
            if (o[(i0 * c.b1 + i1) * c.b2 + i2] != want) ++bad;
        }
        printf("%-50s ok, mismatches %d, first %.0f\n", c.name, bad, o[0]);
    }
    return 0;
}
