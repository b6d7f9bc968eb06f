// Probe: streaming copy (read 268 MB + write 268 MB) as a function of load/store width and cache hints.  L2 flushed per launch.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/probes/probe_copy_bw.cu -o /tmp/probe.bin && /tmp/probe.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int LD> __device__ __forceinline__ uint4 ld16(const uint4 *p) {
    uint4 v;
    if (LD == 0) v = *p;
    else asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int ST> __device__ __forceinline__ void st16(uint4 *p, uint4 v) {
    if (ST == 0) *p = v;
    else if (ST == 1) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int LD, int ST, int UNROLL>
__global__ void __launch_bounds__(256) copy16(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ld16<LD>(x + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { v[u].x += 1; st16<ST>(y + i + u * stride, v[u]); }
    }
    for (; i < n16; i += stride) { uint4 v = ld16<LD>(x + i); v.x += 1; st16<ST>(y + i, v); }
}
struct U8 { uint4 a, b; };
template <int UNROLL>
__global__ void __launch_bounds__(256) copy32(const char *__restrict__ x, char *__restrict__ y, size_t n32) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n32; i += UNROLL * stride) {
        U8 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (i + u * stride < n32)
                asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[u].a.x), "=r"(v[u].a.y), "=r"(v[u].a.z), "=r"(v[u].a.w), "=r"(v[u].b.x), "=r"(v[u].b.y), "=r"(v[u].b.z), "=r"(v[u].b.w)
                             : "l"(x + (i + u * stride) * 32));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (i + u * stride < n32) {
                v[u].a.x += 1;
                asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(y + (i + u * stride) * 32),
                             "r"(v[u].a.x), "r"(v[u].a.y), "r"(v[u].a.z), "r"(v[u].a.w), "r"(v[u].b.x), "r"(v[u].b.y), "r"(v[u].b.z), "r"(v[u].b.w) : "memory");
            }
    }
}

int main() {
    const size_t bytes = 268435456;
    char *x, *y, *flush;
    cudaMalloc(&x, bytes); cudaMalloc(&y, bytes); cudaMalloc(&flush, 512u << 20);
    cudaMemset(x, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, auto launch) {
        float tot = 0.f;
        for (int r = 0; r < 7; ++r) {
            cudaMemsetAsync(flush, r, 512u << 20);
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 2) tot += ms;
        }
        printf("%-56s avg %.1f us  %.0f GB/s (read+write)\n", name, tot / 5 * 1e3, 2 * bytes / (tot / 5) / 1e6);
    };
    const size_t n16 = bytes / 16, n32 = bytes / 32;
    for (int ctas : {148 * 4, 148 * 8, 148 * 16}) {
        char nm[96];
        snprintf(nm, 96, "ld default  / st default,      4 in flight, %d CTAs", ctas); run(nm, [&] { copy16<0, 0, 4><<<ctas, 256>>>((const uint4 *)x, (uint4 *)y, n16); });
        snprintf(nm, 96, "ld no_alloc / st default,      4 in flight, %d CTAs", ctas); run(nm, [&] { copy16<1, 0, 4><<<ctas, 256>>>((const uint4 *)x, (uint4 *)y, n16); });
        snprintf(nm, 96, "ld no_alloc / st no_alloc,     4 in flight, %d CTAs", ctas); run(nm, [&] { copy16<1, 1, 4><<<ctas, 256>>>((const uint4 *)x, (uint4 *)y, n16); });
        snprintf(nm, 96, "ld no_alloc / st .cs,          4 in flight, %d CTAs", ctas); run(nm, [&] { copy16<1, 2, 4><<<ctas, 256>>>((const uint4 *)x, (uint4 *)y, n16); });
        snprintf(nm, 96, "ld no_alloc / st no_alloc,     8 in flight, %d CTAs", ctas); run(nm, [&] { copy16<1, 1, 8><<<ctas, 256>>>((const uint4 *)x, (uint4 *)y, n16); });
        snprintf(nm, 96, "256-bit NA evict_first ld+st,  2 in flight, %d CTAs", ctas); run(nm, [&] { copy32<2><<<ctas, 256>>>(x, y, n32); });
        snprintf(nm, 96, "256-bit NA evict_first ld+st,  4 in flight, %d CTAs", ctas); run(nm, [&] { copy32<4><<<ctas, 256>>>(x, y, n32); });
    }
    return 0;
}
