// Micro-benchmark: how fast can a warp-per-slice kernel stream a packed matrix (4 B per entry, slices of WIDTH x 32 words)
// without any gather?  Variants: words in flight per thread (U), 32-bit vs 128-bit loads, threads per SM, L2 prefetch,
// bulk copies into shared memory.  Prints GB/s.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a stream_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t ldcs(const uint32_t* p) { uint32_t v; asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint4 ldcs4(const uint4* p) { uint4 v; asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }

// V1: as the product kernel: lane-strided 32-bit loads, U in flight, then a dependent use
template <int U, int MAXT, int MINB, int PF>
__global__ void __launch_bounds__(MAXT, MINB) k_v1(const uint32_t* __restrict__ pk, long long nslices, int width, int spc, unsigned* out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long s0 = (long long)blockIdx.x * spc, s1 = min(s0 + spc, nslices);
    unsigned acc = 0;
    for (long long s = s0 + warp; s < s1; s += nwarp) {
        const uint32_t* pp = pk + s * width * 32 + lane;
        if (PF) { const long long sn = s + nwarp; if (sn < s1) for (int l = lane; l < width; l += 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(pk + sn * width * 32 + l * 32)); }
        int j = 0;
        for (; j + U <= width; j += U) {
            uint32_t w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) w[u] = ldcs(pp + (size_t)(j + u) * 32);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += w[u] * (u + 1);
        }
        for (; j < width; ++j) acc += ldcs(pp + (size_t)j * 32);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// V2: 128-bit loads: lane reads 16 B, a warp request is 512 B contiguous; U4 of them in flight (= 4*U4 words)
template <int U4, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_v2(const uint32_t* __restrict__ pk, long long nslices, int width, int spc, unsigned* out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long s0 = (long long)blockIdx.x * spc, s1 = min(s0 + spc, nslices);
    unsigned acc = 0;
    const int w4 = width / 4;
    for (long long s = s0 + warp; s < s1; s += nwarp) {
        const uint4* pp = reinterpret_cast<const uint4*>(pk + s * width * 32) + lane;
        int j = 0;
        for (; j + U4 <= w4; j += U4) {
            uint4 w[U4];
#pragma unroll
            for (int u = 0; u < U4; ++u) w[u] = ldcs4(pp + (size_t)(j + u) * 32);
#pragma unroll
            for (int u = 0; u < U4; ++u) acc += w[u].x + w[u].y * 2 + w[u].z * 3 + w[u].w * 4;
        }
        for (; j < w4; ++j) { uint4 v = ldcs4(pp + (size_t)j * 32); acc += v.x + v.y + v.z + v.w; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// V3: one bulk copy (cp.async.bulk) per slice into a per-warp shared-memory buffer, double buffered; lanes then read smem
template <int MAXT, int MINB, int WMAX>
__global__ void __launch_bounds__(MAXT, MINB) k_v3(const uint32_t* __restrict__ pk, long long nslices, int width, int spc, unsigned* out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem) + (size_t)warp * 2 * WMAX * 32;
    __shared__ uint64_t bar[2 * (MAXT / 32)];
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bar[warp * 2]);
    if (lane == 0) { for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared.b64 [%0], 1;" :: "r"(bar0 + 8 * b)); }
    __syncwarp();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long s0 = (long long)blockIdx.x * spc, s1 = min(s0 + spc, nslices);
    const uint32_t bytes = (uint32_t)width * 128;
    unsigned acc = 0;
    auto issue = [&](long long s, int b) {
        if (lane == 0) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf + (size_t)b * WMAX * 32);
            asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" :: "r"(bar0 + 8 * b), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(pk + s * width * 32), "r"(bytes), "r"(bar0 + 8 * b) : "memory");
        }
    };
    long long s = s0 + warp;
    if (s < s1) issue(s, 0);
    int it = 0;
    for (; s < s1; s += nwarp, ++it) {
        const int b = it & 1;
        if (s + nwarp < s1) issue(s + nwarp, b ^ 1);
        const uint32_t phase = (it >> 1) & 1;
        uint32_t done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar0 + 8 * b), "r"(phase) : "memory");
        const uint32_t* sp = buf + (size_t)b * WMAX * 32 + lane;
#pragma unroll 8
        for (int j = 0; j < width; ++j) acc += sp[j * 32] * (j + 1);
        __syncwarp();
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
static float timeit(F f, int reps = 5)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main()
{
    const int width = 28;
    const long long nslices = 1570176;
    const size_t words = (size_t)nslices * width * 32;
    uint32_t* pk; unsigned* out;
    CK(cudaMalloc(&pk, words * 4)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(pk, 1, words * 4));
    const double gb = words * 4 / 1e9;
    printf("stream %.2f GB, %lld slices of %d x 32 words\n", gb, nslices, width);
#define RUN1(U, T, B, PF, SPC) { const int spc = SPC; const int grid = (int)((nslices + spc - 1) / spc); \
        float ms = timeit([&] { k_v1<U, T, B, PF><<<grid, T>>>(pk, nslices, width, spc, out); }); \
        printf("v1 ld32  U=%2d threads=%d x %d pf=%d spc=%d : %.3f ms  %.0f GB/s\n", U, T, B, PF, spc, ms, gb / ms * 1e3); }
    RUN1(15, 384, 3, 0, 192) RUN1(15, 384, 3, 1, 192) RUN1(28, 384, 3, 0, 192) RUN1(28, 384, 3, 1, 192) RUN1(14, 384, 3, 0, 192)
    RUN1(8, 512, 3, 0, 128) RUN1(8, 512, 4, 0, 128) RUN1(7, 512, 4, 0, 128) RUN1(14, 512, 4, 0, 128) RUN1(28, 512, 4, 0, 128) RUN1(28, 256, 8, 0, 64)
    RUN1(28, 1024, 2, 0, 256) RUN1(14, 1024, 2, 1, 256)
#define RUN2(U4, T, B, SPC) { const int spc = SPC; const int grid = (int)((nslices + spc - 1) / spc); \
        float ms = timeit([&] { k_v2<U4, T, B><<<grid, T>>>(pk, nslices, width, spc, out); }); \
        printf("v2 ld128 U4=%d threads=%d x %d spc=%d : %.3f ms  %.0f GB/s\n", U4, T, B, spc, ms, gb / ms * 1e3); }
    RUN2(4, 384, 3, 192) RUN2(7, 384, 3, 192) RUN2(2, 384, 3, 192) RUN2(4, 512, 4, 128) RUN2(7, 512, 4, 128) RUN2(7, 1024, 2, 256)
#define RUN3(T, B, SPC) { const int spc = SPC; const int grid = (int)((nslices + spc - 1) / spc); const size_t sh = (size_t)(T / 32) * 2 * 32 * 32 * 4; \
        cudaFuncSetAttribute(k_v3<T, B, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); \
        float ms = timeit([&] { k_v3<T, B, 32><<<grid, T, sh>>>(pk, nslices, width, spc, out); }); \
        cudaError_t e = cudaGetLastError(); \
        printf("v3 bulk  threads=%d x %d spc=%d smem=%zu : %.3f ms  %.0f GB/s %s\n", T, B, spc, sh, ms, gb / ms * 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e)); }
    RUN3(384, 2, 192) RUN3(256, 3, 128) RUN3(128, 6, 64) RUN3(128, 4, 64) RUN3(256, 2, 128)
    return 0;
}
