// Device helpers shared by the tcgen05 linear-attention kernels (linattn_tc.cu: sequential-chunk
// kernels; linattn_cp.cu: chunk-parallel kernels): feature map on packed bf16, per-thread tile geometry,
// TMEM score tile -> masked bf16 shared-memory tile, swizzled-tile column sums.
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

// 64x64 fp32 state in TMEM (M=64 layout) -> bf16 smem tile rows e, this thread's 32 columns
__device__ __forceinline__ void state_half_to_smem(const Geo &g, uint32_t tm_col, uint8_t *sS) {
    uint32_t r[32];
    tmem_ld32(g.t_lane + tm_col + 32 * g.half, r);
    tmem_ld_wait();
    if (g.lane < 16) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4 *>(sS + sw128_off(g.erow, g.half * 4 + cc)) = pack8u(r + 8 * cc, 1.f);
    }
}
// initial state (fp32, row-major [e][m]) -> TMEM + bf16 smem tile
__device__ __forceinline__ void seed_state_half(const Geo &g, const float *init, uint32_t tm_col, uint8_t *sS) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(init[g.erow * 64 + g.half * 32 + i]);
    tmem_st32(g.t_lane + tm_col + 32 * g.half, r);
    if (g.lane < 16) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4 *>(sS + sw128_off(g.erow, g.half * 4 + cc)) = pack8u(r + 8 * cc, 1.f);
    }
    tmem_st_wait();
}

// TMEM [128 x 128] score tile -> (+row_add, +col_add[c]) -> triangular mask -> bf16 -> sX block `half`.
// LOWER keeps column c <= row (forward / dq);  otherwise keeps c >= row (dk/dv).  Returns the row sum of
// the bf16-rounded kept entries over this thread's 64 columns.
template <bool LOWER>
__device__ __forceinline__ float convert_scores(const Geo &g, uint32_t tm_col, uint8_t *sX, float row_add, const float *col_add) {
    float rowsum = 0.f;
    const int wq = g.warp & 3;
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
        const int p = 2 * g.half + pp;                      // 32-column piece
        uint32_t r[32];
        const bool live = LOWER ? (p <= wq) : (p >= wq);
        const bool diag = p == wq;
        if (live) {
            tmem_ld32(g.t_lane + tm_col + 32 * p, r);
            tmem_ld_wait();
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c0 = 32 * p + 8 * cc + 2 * i;
                float a = 0.f, b = 0.f;
                if (live) {
                    a = __uint_as_float(r[8 * cc + 2 * i]) + row_add + (col_add ? col_add[c0] : 0.f);
                    b = __uint_as_float(r[8 * cc + 2 * i + 1]) + row_add + (col_add ? col_add[c0 + 1] : 0.f);
                    if (diag) {
                        if (LOWER) { a = c0 <= g.row ? a : 0.f; b = c0 + 1 <= g.row ? b : 0.f; }
                        else { a = c0 >= g.row ? a : 0.f; b = c0 + 1 >= g.row ? b : 0.f; }
                    }
                }
                const __nv_bfloat162 hb = __floats2bfloat162_rn(a, b);
                const float2 fb = __bfloat1622float2(hb);
                rowsum += fb.x + fb.y;
                w[i] = *reinterpret_cast<const uint32_t *>(&hb);
            }
            *reinterpret_cast<uint4 *>(sX + g.half * TILE_BYTES + sw128_off(g.row, pp * 4 + cc)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    return rowsum;
}

// partial column sums over 32 rows (quarter) of a swizzled [128 x 64] bf16 tile, optionally row-weighted
__device__ __forceinline__ float colsum_quarter(const uint8_t *tile, int e, int quarter, const float *w) {
    float s = 0.f;
#pragma unroll 8
    for (int j = 32 * quarter; j < 32 * quarter + 32; ++j) {
        const __nv_bfloat16 x = *reinterpret_cast<const __nv_bfloat16 *>(tile + sw128_off(j, e >> 3) + (e & 7) * 2);
        s = fmaf(__bfloat162float(x), w ? w[j] : 1.f, s);
    }
    return s;
}

// G' = go/den in place over this thread's 4 chunks; returns the partial dot go.out
__device__ __forceinline__ float prep_grad_half(const Geo &g, uint8_t *sG, const uint8_t *sOt, float inv) {
    float dot = 0.f;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const uint32_t off = sw128_off(g.row, 4 * g.half + cc);
        float gg[8], o[8];
        unpack8(*reinterpret_cast<const uint4 *>(sG + off), gg);
        unpack8(*reinterpret_cast<const uint4 *>(sOt + off), o);
#pragma unroll
        for (int i = 0; i < 8; ++i) { dot = fmaf(gg[i], o[i], dot); gg[i] *= inv; }
        *reinterpret_cast<uint4 *>(sG + off) = pack8(gg);
    }
    return dot;
}

}  // namespace tcdev
}  // namespace cpm
