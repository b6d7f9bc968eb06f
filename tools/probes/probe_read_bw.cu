// Probe: what read bandwidth does a plain streaming kernel reach on 268 MB (the (65536, 2048) bf16 gradient tensor), as a
// function of load width / cache hints / loads in flight / CTAs per SM?  L2 flushed before every launch.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/probes/probe_read_bw.cu -o /tmp/probe && /tmp/probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int MODE> __device__ __forceinline__ void ld16(const uint4 *p, uint4 &v) {
    if (MODE == 0) v = *p;
    else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
}
struct U8 { uint4 a, b; };
__device__ __forceinline__ void ld32(const void *p, U8 &v) {
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.a.x), "=r"(v.a.y), "=r"(v.a.z), "=r"(v.a.w), "=r"(v.b.x), "=r"(v.b.y), "=r"(v.b.z), "=r"(v.b.w) : "l"(p));
}

// grid-stride over 16-byte words; UNROLL independent loads in flight per thread
template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) read16(const uint4 *__restrict__ x, size_t n16, uint32_t *sink) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) ld16<MODE>(x + i + u * stride, v[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n16; i += stride) { uint4 v; ld16<MODE>(x + i, v); acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}
template <int UNROLL>
__global__ void __launch_bounds__(256) read32(const char *__restrict__ x, size_t n32, uint32_t *sink) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n32; i += UNROLL * stride) {
        U8 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) ld32(x + (i + u * stride) * 32, v[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].a.x ^ v[u].a.y ^ v[u].b.z ^ v[u].b.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const size_t bytes = 268435456;
    char *x, *flush; uint32_t *sink;
    cudaMalloc(&x, bytes); cudaMalloc(&flush, 512u << 20); cudaMalloc(&sink, 4);
    cudaMemset(x, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, auto launch) {
        float best = 1e9f, tot = 0.f;
        for (int r = 0; r < 7; ++r) {
            cudaMemsetAsync(flush, r, 512u << 20);
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 2) { tot += ms; if (ms < best) best = ms; }
        }
        printf("%-48s avg %.1f us  best %.1f us  %.0f GB/s (avg)\n", name, tot / 5 * 1e3, best * 1e3, bytes / (tot / 5) / 1e6);
    };
    const size_t n16 = bytes / 16, n32 = bytes / 32;
    for (int ctas : {148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        char nm[96];
        snprintf(nm, 96, "LDG.128 default, 4 in flight, %d CTAs", ctas); run(nm, [&] { read16<0, 4><<<ctas, 256>>>((const uint4 *)x, n16, sink); });
        snprintf(nm, 96, "LDG.128 default, 8 in flight, %d CTAs", ctas); run(nm, [&] { read16<0, 8><<<ctas, 256>>>((const uint4 *)x, n16, sink); });
        snprintf(nm, 96, "LDG.128 nc no_allocate, 8 in flight, %d CTAs", ctas); run(nm, [&] { read16<1, 8><<<ctas, 256>>>((const uint4 *)x, n16, sink); });
        snprintf(nm, 96, "LDG.128 no_allocate (not nc), 8 in flight, %d CTAs", ctas); run(nm, [&] { read16<2, 8><<<ctas, 256>>>((const uint4 *)x, n16, sink); });
        snprintf(nm, 96, "LDG.256 NA evict_first, 4 in flight, %d CTAs", ctas); run(nm, [&] { read32<4><<<ctas, 256>>>(x, n32, sink); });
        snprintf(nm, 96, "LDG.256 NA evict_first, 8 in flight, %d CTAs", ctas); run(nm, [&] { read32<8><<<ctas, 256>>>(x, n32, sink); });
    }
    float ms;
    cudaMemsetAsync(flush, 0, 512u << 20);
    cudaEventRecord(e0); cudaMemcpyAsync(flush, x, bytes, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D 268 MB: %.1f us (%.0f GB/s read+write)\n", ms * 1e3, 2 * bytes / ms / 1e6);
    return 0;
}
