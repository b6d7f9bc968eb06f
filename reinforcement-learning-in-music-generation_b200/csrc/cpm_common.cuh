// Shared device/host helpers for libcpmusic (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include "../../include/cpmusic.h"

namespace cpm {

// ---------------------------------------------------------------- host error plumbing
extern thread_local char g_err[512];
int fail(int code, const char *fmt, ...);
inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return CPM_OK;
}
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
#define CPM_REQUIRE(cond, code, ...) do { if (!(cond)) return ::cpm::fail(code, __VA_ARGS__); } while (0)

// ---------------------------------------------------------------- dtype helpers
template <typename T> struct Vec8;           // 8 activations = one 16-byte (bf16) or 32-byte (f32) access
template <> struct Vec8<float> {
    float v[8];
    __device__ __forceinline__ void load(const float *p) {
        float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ __forceinline__ void store(float *p) const {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Vec8<__nv_bfloat16> {
    float v[8];
    __device__ __forceinline__ void load(const __nv_bfloat16 *p) {
        uint4 raw = *reinterpret_cast<const uint4 *>(p);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    __device__ __forceinline__ void store(__nv_bfloat16 *p) const {
        uint4 raw;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4 *>(p) = raw;
    }
};
// the raw bits of 8 activations: loads that stay in flight (no conversion, hence no use of the registers) until unpack()
template <typename T> struct Raw8;
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float *p) { a = *reinterpret_cast<const float4 *>(p); b = *reinterpret_cast<const float4 *>(p + 4); }
    __device__ __forceinline__ void unpack(float (&v)[8]) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
};
template <> struct Raw8<__nv_bfloat16> {
    uint4 r;
    __device__ __forceinline__ void load(const __nv_bfloat16 *p) { r = *reinterpret_cast<const uint4 *>(p); }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
    }
};
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// elu(x)+1 feature map and its derivative (ft default feature map)
__device__ __forceinline__ float phi(float x) { return x > 0.f ? x + 1.f : __expf(x); }
__device__ __forceinline__ float dphi(float x) { return x > 0.f ? 1.f : __expf(x); }

// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. 2011)
struct Philox {
    __device__ __forceinline__ static uint4 block(uint4 c, uint2 k) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
        }
        return c;
    }
};
// Dropout keep-mask for the 8 elements [8*g, 8*g+8): 16 random bits each, keep iff bits >= thr.
__device__ __forceinline__ void dropout_mask8(uint64_t seed, uint64_t offset, uint64_t g, uint32_t thr, bool keep[8]) {
    uint64_t ctr = offset + g;
    uint4 r = Philox::block(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x44524F50u /*"DROP"*/, 0u),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    keep[0] = (r.x & 0xFFFFu) >= thr; keep[1] = (r.x >> 16) >= thr;
    keep[2] = (r.y & 0xFFFFu) >= thr; keep[3] = (r.y >> 16) >= thr;
    keep[4] = (r.z & 0xFFFFu) >= thr; keep[5] = (r.z >> 16) >= thr;
    keep[6] = (r.w & 0xFFFFu) >= thr; keep[7] = (r.w >> 16) >= thr;
}
inline uint32_t dropout_threshold(float p) {           // host: P(drop) = thr / 65536
    if (p <= 0.f) return 0u;
    double t = (double)p * 65536.0 + 0.5;
    return t >= 65535.0 ? 65535u : (uint32_t)t;
}
inline float dropout_scale(float p) { uint32_t t = dropout_threshold(p); return t ? 65536.0f / (65536.0f - (float)t) : 1.0f; }

// device-side RNG base (cpm_set_rng_base, elementwise.cu): added to every dropout kernel's host-side offset
extern const unsigned long long *g_rng_base;
__device__ __forceinline__ uint64_t rng_off(uint64_t host_offset, const unsigned long long *base) {
    return host_offset + (base ? *base : 0ull);
}

// ---------------------------------------------------------------- exact GELU pieces + 16-element dropout groups
// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. exact at fp32-parity tolerance): branch-free,
// two MUFU ops; e = exp(-x^2/2) is shared with the Gaussian term of the derivative.  (CUDA's erff costs ~2x
// the instructions; this kernel is instruction-issue bound, ncu: 40 instr/element before, HBM needs <= 14.)
struct GeluParts { float cdf, e; };            // Phi(x) = 0.5 (1 + erf(x / sqrt 2)),  e = exp(-x^2 / 2)
__device__ __forceinline__ GeluParts gelu_parts(float x) {
    const float a = fabsf(x) * 0.70710678118654752f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170f));      // -0.5 * log2(e)
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float half_erfc = 0.5f * p * t * e;                   // 0.5 * erfc(|x| / sqrt 2)
    GeluParts r;
    r.cdf = x >= 0.f ? 1.f - half_erfc : half_erfc;
    r.e = e;
    return r;
}
// EXACT (the fp32 parity mode): libdevice erff / expf.
template <bool EXACT> __device__ __forceinline__ float gelu_f(float x) {
    if (EXACT) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    return x * gelu_parts(x).cdf;
}
template <bool EXACT> __device__ __forceinline__ float dgelu_f(float x) {
    if (EXACT) return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * expf(-0.5f * x * x);
    const GeluParts g = gelu_parts(x);
    return fmaf(x * 0.39894228040143268f, g.e, g.cdf);
}
// Dropout keep-bits for the 16 elements [16*g, 16*g+16): one Philox4x32-10 block, 8 random bits per element,
// keep iff bits >= thr8 (drop probability quantised to thr8 / 256).
__device__ __forceinline__ uint32_t dropout_keep16(uint64_t seed, uint64_t offset, uint64_t g, uint32_t thr8) {
    const uint64_t ctr = offset + g;
    const uint4 r = Philox::block(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x44523136u /*"DR16"*/, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t keep = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) keep |= (((w[i >> 2] >> (8 * (i & 3))) & 0xFFu) >= thr8 ? 1u : 0u) << i;
    return keep;
}
inline uint32_t dropout_threshold8(float p) {
    if (p <= 0.f) return 0u;
    double t = (double)p * 256.0 + 0.5;
    return t >= 255.0 ? 255u : (t < 1.0 ? 1u : (uint32_t)t);
}
inline float dropout_scale8(float p) { uint32_t t = dropout_threshold8(p); return t ? 256.0f / (256.0f - (float)t) : 1.0f; }


// ---------------------------------------------------------------- programmatic dependent launch (the rollout token step)
// The token step is a chain of small dependent kernels.  With cpm_set_chain_pdl(1) every kernel of the chain is launched with
// the programmatic-stream-serialization attribute: it may become resident while its predecessor still runs, executes its
// set-up, and blocks in griddep_wait() until the predecessor has completed and flushed.  RULE for every chain kernel: call
// griddep_wait() before the first access to anything another kernel of the chain produces OR still reads (buffers are recycled
// by the allocator, so writes are ordered too); only constant data (weights, tables) may be touched before it.  Keep the call
// UNCONDITIONAL and ahead of those accesses in program order: placed inside a branch (a grid-stride variant of the state kernel
// tried `if (first tile) griddep_wait();`), nvcc hoisted the read-only `const __restrict__` loads of q / k / v over it and the
// kernel consumed its predecessor's output too early (profiles/r02_summary.md, section S).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
extern int g_chain_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, dim3 cluster, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (g_chain_pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster.x * cluster.y * cluster.z > 1) {          // thread-block cluster (the K slices of one output tile)
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster.x;
        at[n].val.clusterDim.y = cluster.y;
        at[n].val.clusterDim.z = cluster.z;
        ++n;
    }
    cfg.attrs = n ? at : nullptr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    return launch_chain_cluster(kernel, grid, block, smem, st, dim3(1, 1, 1), static_cast<Args &&>(args)...);
}

inline int num_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

}  // namespace cpm
