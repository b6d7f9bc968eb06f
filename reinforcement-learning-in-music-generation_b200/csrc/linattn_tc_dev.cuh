// Device helpers of the chunk-parallel tcgen05 linear-attention kernels (linattn_cp.cu): tile constants, feature map on
// packed bf16, bf16 pack / unpack, per-thread tile geometry.
#pragma once
#include "cpm_common.cuh"
#include "tc_common.cuh"

namespace cpm {
namespace tcdev {
using namespace tc;

constexpr int CHUNK = 128;
constexpr int NTH = 256;                          // 8 warps: two threads per token row (column halves)
constexpr uint32_t TILE_BYTES = 128 * 128;       // [128 rows x 128 B] bf16 tile

// elu(x)+1 on 8 packed bf16, rounded back to bf16; returns the packed result and the fp32 values
__device__ __forceinline__ uint4 phi8(uint4 raw, float (&f)[8]) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
    uint4 o;
    uint32_t *po = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 x = __bfloat1622float2(h[i]);
        __nv_bfloat162 y = __floats2bfloat162_rn(phi(x.x), phi(x.y));
        float2 yr = __bfloat1622float2(y);
        f[2 * i] = yr.x;
        f[2 * i + 1] = yr.y;
        po[i] = *reinterpret_cast<uint32_t *>(&y);
    }
    return o;
}
__device__ __forceinline__ void unpack8(uint4 raw, float (&f)[8]) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 x = __bfloat1622float2(h[i]); f[2 * i] = x.x; f[2 * i + 1] = x.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ uint4 pack8u(const uint32_t *r, float scale) {
    return make_uint4(pack_bf16(__uint_as_float(r[0]) * scale, __uint_as_float(r[1]) * scale),
                      pack_bf16(__uint_as_float(r[2]) * scale, __uint_as_float(r[3]) * scale),
                      pack_bf16(__uint_as_float(r[4]) * scale, __uint_as_float(r[5]) * scale),
                      pack_bf16(__uint_as_float(r[6]) * scale, __uint_as_float(r[7]) * scale));
}

// Per-thread geometry: 8 warps; warp w reads TMEM lanes 32*(w&3).., i.e. token rows 32*(w&3)+lane, and
// owns column half (w>>2) of every 64-wide row (16-byte chunks 4*half .. 4*half+3).
struct Geo {
    int tid, warp, lane, row, half, erow;
    uint32_t t_lane;
    __device__ __forceinline__ Geo(uint32_t tmem) {
        tid = threadIdx.x; warp = tid >> 5; lane = tid & 31;
        row = ((warp & 3) << 5) + lane; half = warp >> 2;
        erow = 16 * (warp & 3) + (lane & 15);                 // row of a 64x64 state tile (M=64 TMEM layout)
        t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    }
};

}  // namespace tcdev
}  // namespace cpm
