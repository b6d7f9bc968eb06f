// Chunk-parallel causal linear attention on tcgen05 / TMA (sm_100a), bf16 I/O, fp32 accumulation.
//
// One CTA per whole (batch, head) chain of chunks would expose only 128 independent chains of a 16 x 1024 x 8 call to
// 148 SMs (the round-1 design, since removed).  Here every 128-token chunk of every (batch, head) is its own tile of a
// persistent kernel; the inter-chunk dependency (the KV state) is broken out into a tensor-core pre-pass plus a prefix scan:
//
//   forward   F1  cp_state_fwd    per chunk c < C-1 : dS_c = Kf_c^T V_c (M64 N64 K128 UMMA), dz_c = colsum Kf_c   -> fp32 partials
//             F2  cp_scan         per (n,h)         : exclusive prefix over c  -> Sp_c (bf16 tile), zp_c (fp32)   [saved for backward]
//             F3  cp_out_fwd      per chunk         : P = Qf Kf^T (masked), O = Qf Sp_c + P V, den = rowsum P + Qf.zp_c
//   backward  B1  cp_state_bwd    per chunk         : G' = go/den, gd = -(go.out)/den, dR_c = Qf_c^T G'_c, drz_c = Qf_c^T gd
//             B2  cp_scan (reverse)                 : exclusive suffix over c  -> Rs_c (bf16 tile), rzs_c (fp32)
//             (F1+F2 and B1+B2 are replaced by the streaming kernels F1s / B1s - one CTA per (batch, head) accumulating the
//              state in TMEM across chunks - whenever there are >= 96 (batch, head) chains to fill the GPU)
//             B3  cp_bwd_main     per chunk         : dq, dk, dv of the chunk from (q,k,v,go) + Sp_c + Rs_c, all within the CTA
//
// HBM traffic per (chunk, head): forward reads q,k,v (48 KB) and writes out (16 KB) + den; k,v are read a
// second time by the state kernel and the prefix tile costs 8.4 KB each way; backward reads q,k,v,go (64 KB), both state tiles,
// writes dq,dk,dv (48 KB), and its state kernel reads q, go, out once more.  Measured at 128 x 1024 x 8: 2.33 GB of DRAM traffic
// for 1.476 GB algorithmic, every kernel at 0.70-0.80 of the copy bandwidth on its own bytes (profiles/r02_summary.md, K).
// Shared-memory tiles are SWIZZLE_128B as TMA writes them; the intra-chunk score tile goes TMEM -> registers (mask, bf16) ->
// shared memory -> second UMMA.  Output rows leave as whole lines: out / dq / dk through a dead score-tile block and one bulk
// tensor store per tile, dv and the state snapshots as 256-bit stores (128-bit stores at a 3 KB row stride were the longest
// phase of a tile).  Any sequence length: a short last chunk's tile runs on into the next sequence's rows (or TMA zero fill) and
// those rows are masked out (causality forward; G' = 0, gd = 0 backward) and never written.  Head width 64 or 128 (template D).
#include "cpm_common.cuh"
#include "linattn_plan.h"
#include "tc_common.cuh"
#include "linattn_tc_dev.cuh"

namespace cpm {
namespace {
using namespace tc;
using namespace tcdev;

constexpr uint32_t S_TILE_BYTES = 64 * 128;                           // [64 e rows x 64 m] bf16 state tile
constexpr uint32_t IDESC_KK128 = idesc_bf16(128, 128, false, false);   // A K-major, B K-major, N=128
constexpr uint32_t IDESC_KM64 = idesc_bf16(128, 64, false, true);      // A K-major, B MN-major, N=64
constexpr uint32_t IDESC_KK64 = idesc_bf16(128, 64, false, false);     // A K-major, B K-major,  N=64
constexpr uint32_t IDESC_MM64 = idesc_bf16(64, 64, true, true);        // A MN-major, B MN-major (M=64 state tile)


// M=64 accumulator in TMEM columns [col, col+64) -> fp32 rows dst[e*64 + m] (128-thread kernels, warps 0..3:
// warp w holds rows 16w..16w+15 in lanes 0..15 of its TMEM quarter).
__device__ __forceinline__ void state64_to_global(uint32_t tmem, uint32_t col, float *dst) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    const int erow = 16 * warp + (lane & 15);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        uint32_t r[32];
        tmem_ld32(t_lane + col + 32 * p, r);
        tmem_ld_wait();
        if (lane < 16) {
            float4 *d = reinterpret_cast<float4 *>(dst + erow * 64 + 32 * p);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                d[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                   __uint_as_float(r[4 * i + 3]));
        }
    }
}

// column sums over 64 rows (half) of a swizzled [128 x 64] bf16 tile, optionally row-weighted
__device__ __forceinline__ float colsum_half(const uint8_t *tile, int e, int half, const float *w) {
    float s = 0.f;
#pragma unroll 8
    for (int j = 64 * half; j < 64 * half + 64; ++j) {
        const __nv_bfloat16 x = *reinterpret_cast<const __nv_bfloat16 *>(tile + sw128_off(j, e >> 3) + (e & 7) * 2);
        s = fmaf(__bfloat162float(x), w ? w[j] : 1.f, s);
    }
    return s;
}

// ---------------------------------------------------------------- lean elementwise helpers
// These kernels are instruction-issue bound, not tensor- or HBM-bound (ncu: ~15 k warp instructions per
// 128-token tile before this diet), so the per-element code is kept to the minimum: bit-twiddled bf16
// unpack, branch-free feature map, row sums and normalisers pushed onto the tensor core (ones-column UMMA).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// elu(x)+1 = max(x,0) + exp(min(x,0)) on two packed bf16, rounded back to bf16
__device__ __forceinline__ uint32_t phi2(uint32_t u) {
    const float a = __uint_as_float(u << 16), b = __uint_as_float(u & 0xffff0000u);
    const float ya = fmaxf(a, 0.f) + ex2_approx(fminf(a * 1.4426950408889634f, 0.f));
    const float yb = fmaxf(b, 0.f) + ex2_approx(fminf(b * 1.4426950408889634f, 0.f));
    return pack_bf16(ya, yb);
}
__device__ __forceinline__ uint4 phi8_lean(uint4 r) { return make_uint4(phi2(r.x), phi2(r.y), phi2(r.z), phi2(r.w)); }
// same, also accumulating dot += phi(x) . z over the 8 elements (fp32 feature values, before rounding)
__device__ __forceinline__ uint32_t phi2_dot(uint32_t u, float z0, float z1, float &dot) {
    const float a = __uint_as_float(u << 16), b = __uint_as_float(u & 0xffff0000u);
    const float ya = fmaxf(a, 0.f) + ex2_approx(fminf(a * 1.4426950408889634f, 0.f));
    const float yb = fmaxf(b, 0.f) + ex2_approx(fminf(b * 1.4426950408889634f, 0.f));
    dot = fmaf(ya, z0, fmaf(yb, z1, dot));
    return pack_bf16(ya, yb);
}
__device__ __forceinline__ uint4 phi8_dot(uint4 r, const float *z, float &dot) {
    const float4 z0 = *reinterpret_cast<const float4 *>(z), z1 = *reinterpret_cast<const float4 *>(z + 4);
    return make_uint4(phi2_dot(r.x, z0.x, z0.y, dot), phi2_dot(r.y, z0.z, z0.w, dot), phi2_dot(r.z, z1.x, z1.y, dot),
                      phi2_dot(r.w, z1.z, z1.w, dot));
}
__device__ __forceinline__ uint32_t scale2(uint32_t u, float s) {
    return pack_bf16(__uint_as_float(u << 16) * s, __uint_as_float(u & 0xffff0000u) * s);
}

// TMEM [128 x 128] score tile -> (+row_add) (+col_add[c]) -> triangular mask -> bf16 -> sX block `half`.
// LOWER keeps column c <= row, otherwise c >= row.  No row sums here (they come from a ones-column UMMA).
template <bool LOWER, bool HAS_ROW, bool HAS_COL, bool ROWSUM = false>
__device__ __forceinline__ float convert_lean(const Geo &g, uint32_t tm_col, uint8_t *sX, float row_add, const float *col_add) {
    float rowsum = 0.f;                                      // fp32 sum of the kept entries (this thread's 64 columns)
    const int wq = g.warp & 3;
    const uint32_t keep = LOWER ? ((2u << g.lane) - 1u) : ~((1u << g.lane) - 1u);    // diagonal piece: bit i <=> column 32p+i kept
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
        const int p = 2 * g.half + pp;                      // 32-column piece (warp-uniform)
        const bool live = LOWER ? (p <= wq) : (p >= wq);
        uint8_t *dst = sX + g.half * TILE_BYTES;
        if (!live) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4 *>(dst + sw128_off(g.row, pp * 4 + cc)) = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        uint32_t r[32];
        tmem_ld32(g.t_lane + tm_col + 32 * p, r);
        tmem_ld_wait();
        if (HAS_ROW || HAS_COL) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float x = __uint_as_float(r[i]);
                if (HAS_ROW) x += row_add;
                if (HAS_COL) x += col_add[32 * p + i];
                r[i] = __float_as_uint(x);
            }
        }
        if (p == wq) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = (keep >> i) & 1u ? r[i] : 0u;
        }
        if (ROWSUM) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                s0 += __uint_as_float(r[i]); s1 += __uint_as_float(r[i + 1]); s2 += __uint_as_float(r[i + 2]); s3 += __uint_as_float(r[i + 3]);
            }
            rowsum += (s0 + s1) + (s2 + s3);
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
            *reinterpret_cast<uint4 *>(dst + sw128_off(g.row, pp * 4 + cc)) = pack8u(r + 8 * cc, 1.f);
    }
    return rowsum;
}

// 64 contiguous bytes (32 bf16 of one output row, or half a row of a state tile) as TWO 256-bit stores when the address is
// 32-byte aligned: with a 128-bit store per lane a warp instruction puts half a sector per row on the wire, and the rows of a
// warp are 128 B - 3 KB apart.  The stores, not the arithmetic, were the longest phase of the per-chunk kernels
// (tools/phase_timing_linattn_bwd.py: 20.4 k -> 15.6 k cycles per backward tile; 128 x 1024 x 8 fwd 200 -> 187 us, bwd 423 -> 346).
#ifndef CPM_STORE256
#define CPM_STORE256 1
#endif
__device__ __forceinline__ void store64(void *dst, const uint4 (&v)[4]) {
#if CPM_STORE256
    if ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(reinterpret_cast<uint8_t *>(dst) + 32 * i), "r"(v[2 * i].x),
                         "r"(v[2 * i].y), "r"(v[2 * i].z), "r"(v[2 * i].w), "r"(v[2 * i + 1].x), "r"(v[2 * i + 1].y), "r"(v[2 * i + 1].z),
                         "r"(v[2 * i + 1].w)
                         : "memory");
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<uint4 *>(dst)[i] = v[i];
}
__device__ __forceinline__ void store_row32(void *base, int64_t ld_elems, int64_t row, int col, const uint4 (&v)[4]) {
    store64(reinterpret_cast<__nv_bfloat16 *>(base) + row * ld_elems + col, v);
}

// =============================================================================================
// Head width.  D = number of 64-wide halves of a head (1: the reference's 64-wide heads; 2: 128-wide heads, SURVEY §8 a7 / cfg5
// read as 8 x 128).  A 128-wide head is D tiles of q / k (feature halves a) and D tiles of v / out (value halves b) per chunk;
// its state is D x D tiles S_ab = Kf_a^T V_b of 64 x 64 and D key-sum vectors z_a.  Tile (a, b) of chunk-slot s lives at
// ((s * D + a) * D + b) * 8 KB of a state region, z_a at (s * D + a) * 64 floats behind the tiles; an fp32 increment is
// [D*D tiles of 4096][D*64] floats.  The per-chunk kernels below are templates over D with every loop over a / b unrolled.
// =============================================================================================
template <int D> struct Wide {
    static constexpr int TILES = D * D;
    static constexpr int STATE_F = TILES * 4096 + D * 64;                       // floats per fp32 increment
    static constexpr uint32_t S_BYTES = TILES * S_TILE_BYTES;                   // bf16 state tiles per chunk-slot
};
template <int D> __host__ __device__ inline int64_t state_tiles_bytes_w(int64_t nhc) { return nhc * (int64_t)Wide<D>::S_BYTES; }
template <int D> __host__ __device__ inline int64_t state_region_bytes_w(int64_t nhc) { return nhc * (int64_t)(Wide<D>::S_BYTES + D * 256); }

// =============================================================================================
// F1: per-chunk state increments  dS_ab = Kf_a^T V_b, dz_a = colsum Kf_a        (128 threads, D*D*64 TMEM columns)
// =============================================================================================
template <int D> struct PreCfg {
    static constexpr uint32_t F_MISC = 2 * D * TILE_BYTES;                      // forward: K | V
    static constexpr uint32_t F_SMEM = F_MISC + D * 512 + 64;
    static constexpr uint32_t B_MISC = 3 * D * TILE_BYTES;                      // backward: Q | go | out
    static constexpr uint32_t B_SMEM = B_MISC + D * 512 + 512 + 64;
    static constexpr int TCOLS = D * D * 64;
};

template <int D>
__global__ void __launch_bounds__(128)
cp_state_fwd_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, float *__restrict__ part,
                    int L, int H, int nchunks) {
    using C = PreCfg<D>;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sK = sm, *sV = sm + D * TILE_BYTES;                               // D tiles each
    float *sdz = reinterpret_cast<float *>(sm + C::F_MISC);                    // [D][2][64]
    uint64_t *bar_load = reinterpret_cast<uint64_t *>(sm + C::F_MISC + D * 512), *bar_mma = bar_load + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_load + 2);
    const int tid = threadIdx.x;
    const int per = nchunks - 1;
    const int nh = blockIdx.x / per, c = blockIdx.x % per, n = nh / H, h = nh % H;
    const int grow = n * L + c * CHUNK, col0 = h * 64 * D;
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        mbar_expect_tx(bar_load, 2 * D * TILE_BYTES);
#pragma unroll
        for (int a = 0; a < D; ++a) {
            tma_load_2d(sK + a * TILE_BYTES, &tmK, bar_load, col0 + 64 * a, grow);
            tma_load_2d(sV + a * TILE_BYTES, &tmV, bar_load, col0 + 64 * a, grow);
        }
    }
    if ((tid >> 5) == 0) tmem_alloc<C::TCOLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    mbar_wait(bar_load, 0);
#pragma unroll
    for (int a = 0; a < D; ++a) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            const uint32_t off = a * TILE_BYTES + sw128_off(tid, ch);
            float f[8];
            *reinterpret_cast<uint4 *>(sK + off) = phi8(*reinterpret_cast<const uint4 *>(sK + off), f);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int a = 0; a < D; ++a) {
#pragma unroll
            for (int b = 0; b < D; ++b) {
                const uint64_t dK = smem_desc_sw128(smem_u32(sK + a * TILE_BYTES)), dV = smem_desc_sw128(smem_u32(sV + b * TILE_BYTES));
#pragma unroll
                for (int k = 0; k < 8; ++k) mma_ss(tmem + 64 * (a * D + b), dK + 128 * k, dV + 128 * k, IDESC_MM64, k > 0);
            }
        }
        mma_commit(bar_mma);
    }
#pragma unroll
    for (int a = 0; a < D; ++a) sdz[a * 128 + tid] = colsum_half(sK + a * TILE_BYTES, tid & 63, tid >> 6, nullptr);   // overlaps the MMA
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    float *dst = part + ((int64_t)nh * nchunks + c) * Wide<D>::STATE_F;
#pragma unroll
    for (int t = 0; t < D * D; ++t) state64_to_global(tmem, 64 * t, dst + 4096 * t);
    tc_fence_before();
    __syncthreads();
    if (tid < 64) {
#pragma unroll
        for (int a = 0; a < D; ++a) dst[D * D * 4096 + 64 * a + tid] = sdz[a * 128 + tid] + sdz[a * 128 + 64 + tid];
    }
    if ((tid >> 5) == 0) tmem_dealloc<C::TCOLS>(tmem);
}

// =============================================================================================
// B1: G' = go/den, gd = -(go.out)/den (stored per token), dR_ab = Qf_a^T G'_b, drz_a = Qf_a^T gd
// =============================================================================================
template <int D>
__global__ void __launch_bounds__(128)
cp_state_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmGo,
                    const __grid_constant__ CUtensorMap tmO, const float *__restrict__ den, float *__restrict__ gd_out,
                    float *__restrict__ part, int L, int H, int nchunks) {
    using C = PreCfg<D>;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sQ = sm, *sG = sm + D * TILE_BYTES, *sO = sm + 2 * D * TILE_BYTES;
    float *sdr = reinterpret_cast<float *>(sm + C::B_MISC);                    // [D][2][64]
    float *sgd = sdr + D * 128;                                                // [128]
    uint64_t *bar_load = reinterpret_cast<uint64_t *>(sm + C::B_MISC + D * 512 + 512), *bar_mma = bar_load + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_load + 2);
    const int tid = threadIdx.x;
    const int nh = blockIdx.x / nchunks, c = blockIdx.x % nchunks, n = nh / H, h = nh % H;
    const int grow = n * L + c * CHUNK, col0 = h * 64 * D;
    const bool need_state = c > 0;                     // chunk 0's increment is never consumed (suffix scan)
    const bool live = tid < L - c * CHUNK;             // rows past the end of the sequence (a short last chunk): G' = 0, gd = 0, nothing stored
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        mbar_expect_tx(bar_load, (need_state ? 3 : 2) * D * TILE_BYTES);
#pragma unroll
        for (int a = 0; a < D; ++a) {
            tma_load_2d(sG + a * TILE_BYTES, &tmGo, bar_load, col0 + 64 * a, grow);
            tma_load_2d(sO + a * TILE_BYTES, &tmO, bar_load, col0 + 64 * a, grow);
            if (need_state) tma_load_2d(sQ + a * TILE_BYTES, &tmQ, bar_load, col0 + 64 * a, grow);
        }
    }
    if ((tid >> 5) == 0) tmem_alloc<C::TCOLS>(tmem_slot);
    const float inv = live ? 1.f / den[(int64_t)(grow + tid) * H + h] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    mbar_wait(bar_load, 0);
    float dot = 0.f;
#pragma unroll
    for (int a = 0; a < D; ++a) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            const uint32_t off = a * TILE_BYTES + sw128_off(tid, ch);
            float gg[8], o[8];
            unpack8(*reinterpret_cast<const uint4 *>(sG + off), gg);
            unpack8(*reinterpret_cast<const uint4 *>(sO + off), o);
#pragma unroll
            for (int i = 0; i < 8; ++i) { dot = fmaf(gg[i], o[i], dot); gg[i] *= inv; }
            if (need_state) {
                *reinterpret_cast<uint4 *>(sG + off) = pack8(gg);
                float f[8];
                *reinterpret_cast<uint4 *>(sQ + off) = phi8(*reinterpret_cast<const uint4 *>(sQ + off), f);
            }
        }
    }
    const float gd = -inv * dot;
    if (live) gd_out[(int64_t)(grow + tid) * H + h] = gd;
    if (need_state) {                                   // uniform over the CTA
        sgd[tid] = gd;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int a = 0; a < D; ++a) {
#pragma unroll
                for (int b = 0; b < D; ++b) {
                    const uint64_t dQ = smem_desc_sw128(smem_u32(sQ + a * TILE_BYTES)), dG = smem_desc_sw128(smem_u32(sG + b * TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < 8; ++k) mma_ss(tmem + 64 * (a * D + b), dQ + 128 * k, dG + 128 * k, IDESC_MM64, k > 0);
                }
            }
            mma_commit(bar_mma);
        }
#pragma unroll
        for (int a = 0; a < D; ++a) sdr[a * 128 + tid] = colsum_half(sQ + a * TILE_BYTES, tid & 63, tid >> 6, sgd);
        mbar_wait(bar_mma, 0);
        tc_fence_after();
        float *dst = part + ((int64_t)nh * nchunks + c) * Wide<D>::STATE_F;
#pragma unroll
        for (int t = 0; t < D * D; ++t) state64_to_global(tmem, 64 * t, dst + 4096 * t);
        tc_fence_before();
        __syncthreads();
        if (tid < 64) {
#pragma unroll
            for (int a = 0; a < D; ++a) dst[D * D * 4096 + 64 * a + tid] = sdr[a * 128 + tid] + sdr[a * 128 + 64 + tid];
        }
    } else {
        tc_fence_before();
        __syncthreads();
    }
    if ((tid >> 5) == 0) tmem_dealloc<C::TCOLS>(tmem);
}

// =============================================================================================
// F1s: streaming prefix states.  One CTA per (batch, head) walks its chunks in order; the KV state
// accumulates in TMEM across chunks (UMMA accumulate flag) and is snapshotted after every chunk straight
// into the bf16 prefix tile of the NEXT chunk - no fp32 increments, no scan launch.  The loop carries no
// dependency through memory, so K/V tiles are prefetched ST_STAGES chunks ahead by TMA.
// Used when there are enough (batch, head) pairs to fill the GPU; otherwise F1 + F2.
// =============================================================================================
#ifndef CPM_ST_STAGES
#define CPM_ST_STAGES 2
#endif
constexpr int ST_STAGES = CPM_ST_STAGES;            // 2 stages: 66 KB of shared memory, three CTAs per SM
constexpr uint32_t ST_STAGE_BYTES = 2 * TILE_BYTES;
constexpr uint32_t ST_OFF_ONES = ST_STAGES * ST_STAGE_BYTES;                    // 2 KB of bf16 ones (layout-agnostic B operand)
constexpr uint32_t ST_OFF_BAR = ST_OFF_ONES + 2048, ST_SMEM = ST_OFF_BAR + 64;
constexpr uint32_t IDESC_Z8 = idesc_bf16(64, 8, true, true);                    // z[e] += sum_j Kf[j][e] * 1
constexpr uint32_t TS_S = 0, TS_Z = 64;

__global__ void __launch_bounds__(NTH, ST_STAGES == 2 ? 3 : 2)
cp_prefix_stream_fwd_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                            uint8_t *__restrict__ tiles, float *__restrict__ zs, int L, int H, int nchunks) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint32_t *sOnes = reinterpret_cast<uint32_t *>(sm + ST_OFF_ONES);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(sm + ST_OFF_BAR);        // [ST_STAGES]
    uint64_t *bar_mma = bar_full + ST_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nh = blockIdx.x, n = nh / H, h = nh % H;
    const int row0 = n * L, col0 = h * 64;
    const int nwork = nchunks - 1;                       // chunks 0 .. C-2 feed the prefixes of chunks 1 .. C-1
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < ST_STAGES; ++s) mbar_init(bar_full + s, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        for (int s = 0; s < ST_STAGES && s < nwork; ++s) {
            mbar_expect_tx(bar_full + s, ST_STAGE_BYTES);
            tma_load_2d(sm + s * ST_STAGE_BYTES, &tmK, bar_full + s, col0, row0 + s * CHUNK);
            tma_load_2d(sm + s * ST_STAGE_BYTES + TILE_BYTES, &tmV, bar_full + s, col0, row0 + s * CHUNK);
        }
    }
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    for (int i = tid; i < 512; i += NTH) sOnes[i] = 0x3F803F80u;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int row = ((warp & 3) << 5) + lane, half = warp >> 2;                 // phi: two threads per token row
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int erow = 16 * (warp & 3) + (lane & 15);                             // snapshot: M=64 accumulator rows
    const uint64_t dOnes = smem_desc_sw128(smem_u32(sOnes));
    auto phi_stage = [&](int c) {
        uint8_t *sK = sm + (c % ST_STAGES) * ST_STAGE_BYTES;
        mbar_wait(bar_full + (c % ST_STAGES), (c / ST_STAGES) & 1);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const uint32_t off = sw128_off(row, 4 * half + cc);
            *reinterpret_cast<uint4 *>(sK + off) = phi8_lean(*reinterpret_cast<const uint4 *>(sK + off));
        }
    };
    if (nwork > 0) phi_stage(0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    for (int c = 0; c < nwork; ++c) {
        const int s = c % ST_STAGES;
        if (tid == 0) {                                   // S += Kf^T V ; z += Kf^T 1   (accumulated in TMEM across chunks)
            tc_fence_after();
            const uint64_t dK = smem_desc_sw128(smem_u32(sm + s * ST_STAGE_BYTES)), dV = smem_desc_sw128(smem_u32(sm + s * ST_STAGE_BYTES + TILE_BYTES));
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_ss(tmem + TS_S, dK + 128 * k, dV + 128 * k, IDESC_MM64, (c > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_ss(tmem + TS_Z, dK + 128 * k, dOnes, IDESC_Z8, (c > 0 || k > 0) ? 1u : 0u);
            mma_commit(bar_mma);
        }
        if (c + 1 < nwork) phi_stage(c + 1);              // overlaps the MMA of chunk c
        mbar_wait(bar_mma, c & 1);
        tc_fence_after();
        const int64_t slot = (int64_t)nh * nchunks + c + 1;
        {   // snapshot: warp w reads rows 16*(w&3).. (lanes 0..15), columns 32*(w>>2)..+31
            uint32_t r[32];
            tmem_ld32(t_lane + TS_S + 32 * half, r);
            uint32_t z8[8];
            if (half == 0) tmem_ld8(t_lane + TS_Z, z8);
            tmem_ld_wait();
            if (lane < 16) {
                uint4 o[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) o[cc] = pack8u(r + 8 * cc, 1.f);
                store64(tiles + slot * S_TILE_BYTES + erow * 128 + 64 * half, o);
                if (half == 0) zs[slot * 64 + erow] = __uint_as_float(z8[0]);
            }
        }
        fence_proxy_async();                              // phi(c+1) writes -> visible to the next UMMA
        tc_fence_before();
        __syncthreads();                                  // TMEM rows read; stage s consumed
        if (tid == 0 && c + ST_STAGES < nwork) {          // refill the stage
            uint8_t *sK = sm + s * ST_STAGE_BYTES;
            mbar_expect_tx(bar_full + s, ST_STAGE_BYTES);
            tma_load_2d(sK, &tmK, bar_full + s, col0, row0 + (c + ST_STAGES) * CHUNK);
            tma_load_2d(sK + TILE_BYTES, &tmV, bar_full + s, col0, row0 + (c + ST_STAGES) * CHUNK);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

// =============================================================================================
// B1s: streaming suffix states (the backward twin of F1s).  One CTA per (batch, head) walks its chunks from the last to
// the first: G' = go/den, gd = -(go.out)/den (stored per token for B3), R += Qf^T G' and rz += Qf^T gd accumulate in TMEM
// and are snapshotted after every chunk into the bf16 suffix tile of the PREVIOUS chunk - no fp32 increments, no scan.
// Operand prep of chunk c-1 overlaps the UMMA of chunk c; q / go / out tiles are double-buffered by TMA.
// =============================================================================================
constexpr int SB_STAGES = 2;
constexpr uint32_t SB_STAGE_BYTES = 3 * TILE_BYTES;                             // q | go | out
constexpr uint32_t SB_OFF_GD = SB_STAGES * SB_STAGE_BYTES;                      // 2 x [8 rows x 128 tokens] K-major bf16: row 0 gd_hi, row 1 gd_lo
constexpr uint32_t SB_OFF_BAR = SB_OFF_GD + 2 * 2048, SB_SMEM = SB_OFF_BAR + 64;
constexpr uint32_t IDESC_RZ = idesc_bf16(64, 8, true, false);                   // A = Qf (MN-major), B = gd tile (K-major)
constexpr uint32_t TSB_R = 0, TSB_RZ = 64;

__global__ void __launch_bounds__(NTH)
cp_suffix_stream_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmGo,
                            const __grid_constant__ CUtensorMap tmO, const float *__restrict__ den, float *__restrict__ gd_out,
                            uint8_t *__restrict__ tiles, float *__restrict__ rzs, int L, int H, int nchunks) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sGd = sm + SB_OFF_GD;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(sm + SB_OFF_BAR);        // [SB_STAGES]
    uint64_t *bar_mma = bar_full + SB_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nh = blockIdx.x, n = nh / H, h = nh % H;
    const int row0 = n * L, col0 = h * 64;
    auto chunk_of = [&](int w) { return nchunks - 1 - w; };                    // work item w = 0 .. C-1 walks the chunks backwards
    auto issue = [&](int w) {                                                  // tid 0
        const int s = w % SB_STAGES, c = chunk_of(w);
        uint8_t *st = sm + s * SB_STAGE_BYTES;
        mbar_expect_tx(bar_full + s, (c > 0 ? 3 : 2) * TILE_BYTES);
        tma_load_2d(st + TILE_BYTES, &tmGo, bar_full + s, col0, row0 + c * CHUNK);
        tma_load_2d(st + 2 * TILE_BYTES, &tmO, bar_full + s, col0, row0 + c * CHUNK);
        if (c > 0) tma_load_2d(st, &tmQ, bar_full + s, col0, row0 + c * CHUNK);
    };
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < SB_STAGES; ++s) mbar_init(bar_full + s, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        for (int w = 0; w < SB_STAGES && w < nchunks; ++w) issue(w);
    }
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    for (int i = tid; i < 1024; i += NTH) reinterpret_cast<uint32_t *>(sGd)[i] = 0u;      // rows 2..7 of both gd tiles stay zero
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int prow = tid >> 1, phalf = tid & 1;                                 // operand prep: two ADJACENT threads per token row
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int erow = 16 * (warp & 3) + (lane & 15), shalf = warp >> 2;          // snapshot: M=64 accumulator rows, column halves
    // rows past the end of the sequence (short last chunk): den reads as +inf, so G' = go / den = 0 and gd = 0; nothing is stored
    auto den_of = [&](int c) { return c * CHUNK + prow < L ? den[(int64_t)(row0 + c * CHUNK + prow) * H + h] : __int_as_float(0x7f800000); };
    float den_reg = den_of(chunk_of(0));
    auto prep = [&](int w) {
        const int s = w % SB_STAGES, c = chunk_of(w);
        uint8_t *sQ = sm + s * SB_STAGE_BYTES, *sG = sQ + TILE_BYTES, *sO = sQ + 2 * TILE_BYTES;
        const float inv = 1.f / den_reg;
        if (w + 1 < nchunks) den_reg = den_of(chunk_of(w + 1));                 // in flight during this prep
        mbar_wait(bar_full + s, (w / SB_STAGES) & 1);
        float dot = 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const uint32_t off = sw128_off(prow, 4 * phalf + cc);
            float gg[8], o[8];
            const uint4 graw = *reinterpret_cast<const uint4 *>(sG + off);
            unpack8(graw, gg);
            unpack8(*reinterpret_cast<const uint4 *>(sO + off), o);
#pragma unroll
            for (int i = 0; i < 8; ++i) dot = fmaf(gg[i], o[i], dot);
            if (c > 0) {
                *reinterpret_cast<uint4 *>(sG + off) = make_uint4(scale2(graw.x, inv), scale2(graw.y, inv), scale2(graw.z, inv), scale2(graw.w, inv));
                *reinterpret_cast<uint4 *>(sQ + off) = phi8_lean(*reinterpret_cast<const uint4 *>(sQ + off));
            }
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        const float gd = -inv * dot;
        if (phalf == 0) {
            if (c * CHUNK + prow < L) gd_out[(int64_t)(row0 + c * CHUNK + prow) * H + h] = gd;
            if (c > 0) {
                const __nv_bfloat16 hi = __float2bfloat16_rn(gd), lo = __float2bfloat16_rn(gd - __bfloat162float(hi));
                uint8_t *gt = sGd + (w & 1) * 2048 + (prow >> 6) * 1024;
                const int i63 = prow & 63;
                *reinterpret_cast<__nv_bfloat16 *>(gt + 0 * 128 + (((i63 >> 3) ^ 0) << 4) + (i63 & 7) * 2) = hi;
                *reinterpret_cast<__nv_bfloat16 *>(gt + 1 * 128 + (((i63 >> 3) ^ 1) << 4) + (i63 & 7) * 2) = lo;
            }
        }
    };
    prep(0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    for (int w = 0; w + 1 < nchunks; ++w) {                                     // chunks C-1 .. 1 feed the suffixes of chunks C-2 .. 0
        const int s = w % SB_STAGES, c = chunk_of(w);
        if (tid == 0) {
            tc_fence_after();
            const uint64_t dQ = smem_desc_sw128(smem_u32(sm + s * SB_STAGE_BYTES)), dG = smem_desc_sw128(smem_u32(sm + s * SB_STAGE_BYTES + TILE_BYTES));
            const uint64_t dGd = smem_desc_sw128(smem_u32(sGd + (w & 1) * 2048));
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_ss(tmem + TSB_R, dQ + 128 * k, dG + 128 * k, IDESC_MM64, (w > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_ss(tmem + TSB_RZ, dQ + 128 * k, dGd + (k >> 2) * 64 + 2 * (k & 3), IDESC_RZ, (w > 0 || k > 0) ? 1u : 0u);
            mma_commit(bar_mma);
        }
        prep(w + 1);                                                            // overlaps the UMMA of chunk c
        mbar_wait(bar_mma, w & 1);
        tc_fence_after();
        const int64_t slot = (int64_t)nh * nchunks + c - 1;
        {
            uint32_t r[32];
            tmem_ld32(t_lane + TSB_R + 32 * shalf, r);
            uint32_t z8[8];
            if (shalf == 0) tmem_ld8(t_lane + TSB_RZ, z8);
            tmem_ld_wait();
            if (lane < 16) {
                uint4 o[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) o[cc] = pack8u(r + 8 * cc, 1.f);
                store64(tiles + slot * S_TILE_BYTES + erow * 128 + 64 * shalf, o);
                if (shalf == 0) rzs[slot * 64 + erow] = __uint_as_float(z8[0]) + __uint_as_float(z8[1]);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();                                                        // TMEM rows read; stage s consumed; prep(w+1) visible
        if (tid == 0 && w + SB_STAGES < nchunks) issue(w + SB_STAGES);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

// =============================================================================================
// F2 / B2: exclusive prefix (reverse = 0) or exclusive suffix (reverse = 1) of the per-chunk increments over
// the chunk axis; fp32 accumulation, bf16 state tile + fp32 z out.  One thread = 4 consecutive state floats.
// =============================================================================================
template <int D>
__global__ void __launch_bounds__(256)
cp_scan_kernel(const float *__restrict__ part, uint8_t *__restrict__ tiles, float *__restrict__ zs, int nchunks, int reverse) {
    constexpr int STATE_F = Wide<D>::STATE_F, TILE4 = Wide<D>::TILES * 1024;     // float4 groups of the tiles part
    const int nh = blockIdx.x, i4 = blockIdx.y * 256 + threadIdx.x;
    if (i4 >= STATE_F / 4) return;
    const float4 *src = reinterpret_cast<const float4 *>(part + (int64_t)nh * nchunks * STATE_F) + i4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int s = 0; s + 1 < nchunks; ++s) {
        const int c = reverse ? nchunks - 1 - s : s;        // increment consumed
        const int d = reverse ? c - 1 : c + 1;              // chunk that receives the running sum
        const float4 p = src[(int64_t)c * (STATE_F / 4)];
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        const int64_t slot = (int64_t)nh * nchunks + d;
        if (i4 < TILE4) {
            uint2 o = make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
            *reinterpret_cast<uint2 *>(tiles + slot * Wide<D>::S_BYTES + (int64_t)i4 * 8) = o;
        } else {
            *reinterpret_cast<float4 *>(zs + slot * (D * 64) + (i4 - TILE4) * 4) = acc;
        }
    }
}

// =============================================================================================
// F3: per-chunk outputs.  256 threads (two per token row), 128 TMEM columns; 3 CTAs / SM for 64-wide heads, 1 for 128-wide.
//   shared memory: sQ (D tiles) | sV (D) | sS (D*D state tiles) | sK (D) .. sP  (the 32 KB bf16 score tile starts on the dead K tiles)
//   P = sum_a Qf_a Kf_a^T (masked);  O_b = sum_a Qf_a Sp_ab + P V_b  in TMEM columns [64 b, 64 b + 64);  den = rowsum P + sum_a Qf_a . z_a
// =============================================================================================
template <int D> struct FwdCfg {
    static constexpr uint32_t OFF_Q = 0, OFF_V = D * TILE_BYTES, OFF_S = 2 * D * TILE_BYTES, OFF_K = OFF_S + D * D * S_TILE_BYTES, OFF_P = OFF_K;
    static constexpr uint32_t OFF_Z = OFF_K + 2 * TILE_BYTES /* D*64 floats */, OFF_DP = OFF_Z + D * 256 /* 2 x 128 floats */,
                              OFF_BAR = OFF_DP + 1024, SMEM = OFF_BAR + 32;
};
static_assert(FwdCfg<1>::OFF_Z == 73728 && FwdCfg<1>::SMEM == 75040, "64-wide layout: three CTAs per SM");

#ifndef CPM_TMA_STORE
#define CPM_TMA_STORE 1       // output rows leave through a swizzled shared-memory tile and ONE bulk tensor store (whole 128-byte lines)
#endif
template <int D>
__global__ void __launch_bounds__(NTH, D == 1 ? 3 : 1)
cp_out_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmS,
                  const __grid_constant__ CUtensorMap tmO, const float *__restrict__ zp,
                  void *__restrict__ out, float *__restrict__ den, int L, int H, int nchunks, int NH, int64_t ld_o, float eps,
                  long long *__restrict__ dbg) {
    // Persistent: CTA b handles tiles b, b + gridDim.x, ... in chunk-major order (tile t = chunk t / NH of pair t % NH),
    // so barrier / TMEM set-up is paid once and the next tile's TMA loads fly while this tile's epilogue runs.
    using C = FwdCfg<D>;
    // bulk-store staging: the score tile's second block - dead after the second UMMA round and, with 64-wide heads, not a target
    // of the next tile's loads (K's one tile lies under the first block).  128-wide heads have two K tiles under it: plain stores.
    constexpr bool TMA_OUT = CPM_TMA_STORE && D == 1;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sQ = sm + C::OFF_Q, *sK = sm + C::OFF_K, *sV = sm + C::OFF_V, *sS = sm + C::OFF_S, *sP = sm + C::OFF_P;
    uint8_t *sOut = sP + TILE_BYTES;
    float *sz = reinterpret_cast<float *>(sm + C::OFF_Z), *sdp = reinterpret_cast<float *>(sm + C::OFF_DP);
    uint64_t *bar_load = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR), *bar_mma = bar_load + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_load + 2);
    const int tid = threadIdx.x;
    const int ntiles = NH * nchunks;
    auto issue_loads = [&](int t) {                   // tid 0 only
        const int c = t / NH, nh = t % NH, n = nh / H, h = nh % H;
        const int grow = n * L + c * CHUNK, col0 = h * 64 * D;
        mbar_expect_tx(bar_load, 3 * D * TILE_BYTES + (c > 0 ? D * D * S_TILE_BYTES : 0));
#pragma unroll
        for (int a = 0; a < D; ++a) {
            tma_load_2d(sQ + a * TILE_BYTES, &tmQ, bar_load, col0 + 64 * a, grow);
            tma_load_2d(sK + a * TILE_BYTES, &tmK, bar_load, col0 + 64 * a, grow);
            tma_load_2d(sV + a * TILE_BYTES, &tmV, bar_load, col0 + 64 * a, grow);
        }
        if (c > 0) {
#pragma unroll
            for (int i = 0; i < D * D; ++i)
                tma_load_2d(sS + i * S_TILE_BYTES, &tmS, bar_load, 0, (int)((((int64_t)nh * nchunks + c) * (D * D) + i) * 64));
        }
    };
    auto fetch_z = [&](int t) -> float {              // tid < 64 D: the key-sum prefix of tile t (0 for a sequence's first chunk)
        const int c = t / NH, nh = t % NH;
        return c > 0 ? zp[((int64_t)nh * nchunks + c) * (64 * D) + tid] : 0.f;
    };
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        if ((int)blockIdx.x < ntiles) issue_loads(blockIdx.x);
    }
    if ((tid >> 5) == 0) tmem_alloc<128>(tmem_slot);
    if (tid < 64 * D && (int)blockIdx.x < ntiles) sz[tid] = fetch_z(blockIdx.x);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const Geo g(tmem);
    const uint64_t dQ = smem_desc_sw128(smem_u32(sQ)), dK = smem_desc_sw128(smem_u32(sK)), dV = smem_desc_sw128(smem_u32(sV));
    const uint64_t dP = smem_desc_sw128(smem_u32(sP)), dS = smem_desc_sw128(smem_u32(sS));
    constexpr uint32_t TSTEP = TILE_BYTES >> 4, SSTEP = S_TILE_BYTES >> 4;          // descriptor steps to the next operand / state tile
    uint32_t ph_load = 0, ph_mma = 0;
    int dbg_i = 0;
#define CPM_STAMP() do { if (dbg && tid == 0 && dbg_i < 64) dbg[(int64_t)blockIdx.x * 64 + dbg_i++] = clock64(); } while (0)
    CPM_STAMP();
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int c = t / NH, nh = t % NH, n = nh / H, h = nh % H;
        const int grow = n * L + c * CHUNK, col0 = h * 64 * D;
        const bool have_state = c > 0;
        // A sequence's last chunk may be short (L % 128 != 0): its tile then runs past the sequence into the rows of the next one
        // (or past the tensor: zero fill).  Causality keeps those rows out of every real row's result; they are never WRITTEN.
        const int valid = min(CHUNK, L - c * CHUNK);
        const int tn = t + gridDim.x;
        float z_next = 0.f;                               // prefetched now, parked in a register until the epilogue
        if (tid < 64 * D && tn < ntiles) z_next = fetch_z(tn);
        mbar_wait(bar_load, ph_load);
        ph_load ^= 1;
        CPM_STAMP();
        float den_part = 0.f;
#pragma unroll
        for (int a = 0; a < D; ++a) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int ch = 4 * g.half + cc;
                const uint32_t off = a * TILE_BYTES + sw128_off(g.row, ch);
                *reinterpret_cast<uint4 *>(sQ + off) = phi8_dot(*reinterpret_cast<const uint4 *>(sQ + off), sz + 64 * a + 8 * ch, den_part);
                *reinterpret_cast<uint4 *>(sK + off) = phi8_lean(*reinterpret_cast<const uint4 *>(sK + off));
            }
        }
        if (TMA_OUT && tid == 0) tma_store_wait_read0();  // the previous tile's output store has read its staging block (long ago)
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        CPM_STAMP();
        if (tid == 0) {                                   // P = sum_a Qf_a Kf_a^T
            tc_fence_after();
#pragma unroll
            for (int a = 0; a < D; ++a) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss(tmem, dQ + a * TSTEP + 2 * k, dK + a * TSTEP + 2 * k, IDESC_KK128, (a | k) > 0);
            }
            mma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        CPM_STAMP();
        den_part += convert_lean<true, false, false, true>(g, 0, sP, 0.f, nullptr);           // overwrites the dead K tiles
        sdp[g.half * 128 + g.row] = den_part;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        CPM_STAMP();
        if (tid == 0) {                                   // O_b = sum_a Qf_a Sp_ab + P V_b   (onto the score tile's columns [64 b, 64 b + 64))
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < D; ++b) {
                uint32_t acc = 0;
                if (have_state) {
#pragma unroll
                    for (int a = 0; a < D; ++a) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) { mma_ss(tmem + 64 * b, dQ + a * TSTEP + 2 * k, dS + (a * D + b) * SSTEP + 128 * k, IDESC_KM64, acc); acc = 1; }
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) { mma_ss(tmem + 64 * b, dP + (k >> 2) * TSTEP + 2 * (k & 3), dV + b * TSTEP + 128 * k, IDESC_KM64, acc); acc = 1; }
            }
            mma_commit(bar_mma);
        }
        const float dn = sdp[g.row] + sdp[128 + g.row] + eps;
        const float inv = 1.f / dn;
        if (den && g.half == 0 && g.row < valid) den[(int64_t)(grow + g.row) * H + h] = dn;
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        CPM_STAMP();
        if (tn < ntiles) {                                // every operand tile is dead: fetch the next tile under the epilogue
            if (tid == 0) issue_loads(tn);
            if (tid < 64 * D) sz[tid] = z_next;
        }
#pragma unroll
        for (int b = 0; b < D; ++b) {
            uint32_t r[32];
            tmem_ld32(g.t_lane + 64 * b + 32 * g.half, r);
            tmem_ld_wait();
            uint4 o[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) o[cc] = pack8u(r + 8 * cc, inv);
            if (TMA_OUT && valid == CHUNK) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4 *>(sOut + sw128_off(g.row, 4 * g.half + cc)) = o[cc];
                fence_proxy_async();
            } else if (g.row < valid) {
                store_row32(out, ld_o, grow + g.row, col0 + 64 * b + 32 * g.half, o);
            }
        }
        tc_fence_before();
        __syncthreads();                                  // TMEM columns, sz and sdp are reused by the next tile
        tc_fence_after();
        if (TMA_OUT && valid == CHUNK && tid == 0) { tma_store_2d(&tmO, sOut, col0, grow); tma_store_commit(); }
        CPM_STAMP();
    }
#undef CPM_STAMP
    if (TMA_OUT && tid == 0) tma_store_wait_all0();
    if ((tid >> 5) == 0) tmem_dealloc<128>(tmem);
}

// =============================================================================================
// B3: dq, dk, dv of one chunk.  256 threads, 256 TMEM columns, 2 CTAs / SM, persistent over tiles.
//   X[i][j]  = G'[i].v[j] + gd_i   (j <= i)        dQf = X Kf + G' Sp^T + gd z         dq = dQf * phi'(q)
//                                                  dKf = X^T Qf + v Rs^T + rz          dk = dKf * phi'(k)
//   PT[j][i] = Kf[j].Qf[i]         (i >= j)        dv  = PT G' + Kf Rs
// Three UMMA rounds: X; {dQf, dKf, PT}; dv.  X^T costs nothing: the stored score tile is read as an MN-major A operand.
// Loads are split over two barriers so the tiles that die first (Q, V, Sp after round 2) are refilled for the next
// tile a whole round before the rest (K, G', Rs after round 3).
// =============================================================================================
template <int D> struct BwdCfg {
    static constexpr uint32_t OFF_Q = 0, OFF_K = D * TILE_BYTES, OFF_V = 2 * D * TILE_BYTES, OFF_G = 3 * D * TILE_BYTES, OFF_S = 4 * D * TILE_BYTES,
                              OFF_R = OFF_S + D * D * S_TILE_BYTES, OFF_X = OFF_R + D * D * S_TILE_BYTES;
    static constexpr uint32_t OFF_GD = OFF_X + 2 * TILE_BYTES /* 128 floats: rz */, OFF_Z = OFF_GD + 512 /* D*64 floats */, OFF_BAR = OFF_Z + D * 256,
                              SMEM = OFF_BAR + 32;
    static constexpr uint32_t TB_X = 0, TB_A1 = 128, TB_A2 = 128 + 64 * D;        // score tile | dQf_a (D x 64 columns) | dKf_a
    static constexpr int TCOLS = D == 1 ? 256 : 512;
};
static_assert(BwdCfg<1>::SMEM == 115488 && BwdCfg<1>::TB_A2 == 192, "64-wide layout: two CTAs per SM");
static_assert(BwdCfg<2>::SMEM <= 232448, "128-wide layout: one CTA per SM, within the 227 KB a block may take");

struct BwdMainArgs {
    const float *den, *gd, *zp, *rzs;
    void *gq, *gk, *gv;
    int64_t ld_g;
    int L, H, nchunks, NH;
    long long *dbg;                                  // development aid: clock64 stamps of thread 0 at the phase boundaries (128 per CTA), or NULL
};

// acc (32 fp32 TMEM values) + add[c] (+ rowscale * vec[c]) then * phi'(f) -> 32 bf16
template <bool HAS_VEC>
__device__ __forceinline__ void grad_row_epilogue(const uint32_t (&r)[32], const uint32_t (&fr)[16], const float *vec, float rowscale,
                                                  uint4 (&o)[4]) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t u = fr[4 * cc + i];
            const float f0 = __uint_as_float(u << 16), f1 = __uint_as_float(u & 0xffff0000u);
            float a0 = __uint_as_float(r[8 * cc + 2 * i]), a1 = __uint_as_float(r[8 * cc + 2 * i + 1]);
            if (HAS_VEC) { a0 = fmaf(rowscale, vec[8 * cc + 2 * i], a0); a1 = fmaf(rowscale, vec[8 * cc + 2 * i + 1], a1); }
            w[i] = pack_bf16(a0 * fminf(f0, 1.f), a1 * fminf(f1, 1.f));
        }
        o[cc] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <int D>
__global__ void __launch_bounds__(NTH, D == 1 ? 2 : 1)
cp_bwd_main_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmGo,
                   const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmR,
                   const __grid_constant__ CUtensorMap tmGq, const __grid_constant__ CUtensorMap tmGk, BwdMainArgs a) {
    using C = BwdCfg<D>;
    constexpr uint32_t TB_X = C::TB_X, TB_A1 = C::TB_A1, TB_A2 = C::TB_A2;
    // dq / dk leave through the score buffer and two bulk tensor stores; 128-wide heads have four such tiles for the one 32 KB
    // buffer and keep the 256-bit stores
    constexpr bool TMA_OUT = CPM_TMA_STORE && D == 1;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sQ = sm + C::OFF_Q, *sK = sm + C::OFF_K, *sV = sm + C::OFF_V, *sG = sm + C::OFF_G, *sS = sm + C::OFF_S, *sR = sm + C::OFF_R;
    uint8_t *sX = sm + C::OFF_X;
    float *sgd = reinterpret_cast<float *>(sm + C::OFF_GD), *sz = reinterpret_cast<float *>(sm + C::OFF_Z);
    uint64_t *bar_a = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR), *bar_b = bar_a + 1, *bar_mma = bar_a + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_a + 3);
    const int tid = threadIdx.x;
    const int ntiles = a.NH * a.nchunks;
    auto tile_coords = [&](int t, int &c, int &nh, int &grow, int &col0, int &h) {
        c = t / a.NH; nh = t % a.NH;
        const int n = nh / a.H;
        h = nh % a.H;
        grow = n * a.L + c * CHUNK; col0 = h * 64 * D;
    };
    auto issue_a = [&](int t) {                       // tid 0: the tiles that die last (after round 3): K, go, Rs
        int c, nh, grow, col0, h;
        tile_coords(t, c, nh, grow, col0, h);
        const bool hr = c + 1 < a.nchunks;
        mbar_expect_tx(bar_a, 2 * D * TILE_BYTES + (hr ? D * D * S_TILE_BYTES : 0));
#pragma unroll
        for (int i = 0; i < D; ++i) {
            tma_load_2d(sK + i * TILE_BYTES, &tmK, bar_a, col0 + 64 * i, grow);
            tma_load_2d(sG + i * TILE_BYTES, &tmGo, bar_a, col0 + 64 * i, grow);
        }
        if (hr) {
#pragma unroll
            for (int i = 0; i < D * D; ++i)
                tma_load_2d(sR + i * S_TILE_BYTES, &tmR, bar_a, 0, (int)((((int64_t)nh * a.nchunks + c) * (D * D) + i) * 64));
        }
    };
    auto issue_b = [&](int t) {                       // tid 0: the tiles that die after round 2: Q, V, Sp
        int c, nh, grow, col0, h;
        tile_coords(t, c, nh, grow, col0, h);
        const bool hs = c > 0;
        mbar_expect_tx(bar_b, 2 * D * TILE_BYTES + (hs ? D * D * S_TILE_BYTES : 0));
#pragma unroll
        for (int i = 0; i < D; ++i) {
            tma_load_2d(sQ + i * TILE_BYTES, &tmQ, bar_b, col0 + 64 * i, grow);
            tma_load_2d(sV + i * TILE_BYTES, &tmV, bar_b, col0 + 64 * i, grow);
        }
        if (hs) {
#pragma unroll
            for (int i = 0; i < D * D; ++i)
                tma_load_2d(sS + i * S_TILE_BYTES, &tmS, bar_b, 0, (int)((((int64_t)nh * a.nchunks + c) * (D * D) + i) * 64));
        }
    };
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        mbar_init(bar_a, 1);
        mbar_init(bar_b, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        if ((int)blockIdx.x < ntiles) { issue_b(blockIdx.x); issue_a(blockIdx.x); }
    }
    if ((tid >> 5) == 0) tmem_alloc<C::TCOLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const Geo g(tmem);
    const uint64_t dQ = smem_desc_sw128(smem_u32(sQ)), dK = smem_desc_sw128(smem_u32(sK)), dV = smem_desc_sw128(smem_u32(sV));
    const uint64_t dG = smem_desc_sw128(smem_u32(sG)), dS = smem_desc_sw128(smem_u32(sS)), dR = smem_desc_sw128(smem_u32(sR));
    const uint64_t dX = smem_desc_sw128(smem_u32(sX));
    constexpr uint32_t TSTEP = TILE_BYTES >> 4, SSTEP = S_TILE_BYTES >> 4;          // descriptor steps to the next operand / state tile
    // The masked score tile X (rows i, 64 j contiguous per 128-byte row, two 16 KB blocks for j < 64 / j >= 64) read as an
    // MN-major A operand IS its transpose: M = j (two 64-wide atoms, LBO = one block apart), K = i.  WT = X^T (same values, same
    // mask) therefore needs no UMMA and no conversion pass of its own.
    const uint64_t dXT = smem_desc_sw128(smem_u32(sX), TILE_BYTES, 1024);
    constexpr uint32_t IDESC_MM128 = idesc_bf16(128, 64, true, true);
    // per-row scalars of the first tile; later tiles' are prefetched one tile ahead
    float inv_n = 0.f, gd_n = 0.f;
    if ((int)blockIdx.x < ntiles) {
        int c, nh, grow, col0, h;
        tile_coords(blockIdx.x, c, nh, grow, col0, h);
        const int64_t ri = (int64_t)(grow + g.row) * a.H + h;
        const bool live = c * CHUNK + g.row < a.L;          // rows past the end of the sequence: G' = 0, gd = 0, nothing stored
        inv_n = live ? 1.f / a.den[ri] : 0.f;
        gd_n = live ? a.gd[ri] : 0.f;
    }
    uint32_t ph_a = 0, ph_b = 0, ph_mma = 0;
    int dbg_i = 0;
#define CPM_STAMP() do { if (a.dbg && tid == 0 && dbg_i < 128) a.dbg[(int64_t)blockIdx.x * 128 + dbg_i++] = clock64(); } while (0)
    CPM_STAMP();
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int c, nh, grow, col0, h;
        tile_coords(t, c, nh, grow, col0, h);
        const bool have_s = c > 0, have_r = c + 1 < a.nchunks;
        const int valid = min(CHUNK, a.L - c * CHUNK);      // < 128 in a short last chunk: those rows are computed with G' = 0 and not stored
        const int64_t slot = (int64_t)nh * a.nchunks + c;
        const int tn = t + gridDim.x;
        const float inv = inv_n, gd = gd_n;
        if (tid < 64 * D) {                                // z (dq epilogue) and rz (dk epilogue) of this tile
            sz[tid] = have_s ? a.zp[slot * (64 * D) + tid] : 0.f;
            sgd[tid] = have_r ? a.rzs[slot * (64 * D) + tid] : 0.f;
        }
        if (tn < ntiles) {                                // next tile's per-row scalars: in flight for the whole tile
            int c2, nh2, grow2, col2, h2;
            tile_coords(tn, c2, nh2, grow2, col2, h2);
            const int64_t ri = (int64_t)(grow2 + g.row) * a.H + h2;
            const bool live = c2 * CHUNK + g.row < a.L;
            inv_n = live ? a.den[ri] : __int_as_float(0x7f800000);      // inverted below: 1 / inf = 0
            gd_n = live ? a.gd[ri] : 0.f;
        }
        uint32_t qfr[D][16], kfr[D][16];                   // this thread's Qf / Kf values (packed bf16) for phi'
        mbar_wait(bar_b, ph_b);
        ph_b ^= 1;
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const uint32_t off = i * TILE_BYTES + sw128_off(g.row, 4 * g.half + cc);
                const uint4 qv = phi8_lean(*reinterpret_cast<const uint4 *>(sQ + off));
                *reinterpret_cast<uint4 *>(sQ + off) = qv;
                qfr[i][4 * cc + 0] = qv.x; qfr[i][4 * cc + 1] = qv.y; qfr[i][4 * cc + 2] = qv.z; qfr[i][4 * cc + 3] = qv.w;
            }
        }
        CPM_STAMP();                                       // 1: Q, V, Sp landed; phi(Q) done
        mbar_wait(bar_a, ph_a);
        ph_a ^= 1;
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const uint32_t off = i * TILE_BYTES + sw128_off(g.row, 4 * g.half + cc);
                const uint4 gv = *reinterpret_cast<const uint4 *>(sG + off);
                *reinterpret_cast<uint4 *>(sG + off) = make_uint4(scale2(gv.x, inv), scale2(gv.y, inv), scale2(gv.z, inv), scale2(gv.w, inv));
                const uint4 kv = phi8_lean(*reinterpret_cast<const uint4 *>(sK + off));
                *reinterpret_cast<uint4 *>(sK + off) = kv;
                kfr[i][4 * cc + 0] = kv.x; kfr[i][4 * cc + 1] = kv.y; kfr[i][4 * cc + 2] = kv.z; kfr[i][4 * cc + 3] = kv.w;
            }
        }
        if (TMA_OUT && tid == 0) tma_store_wait_read0();   // the previous tile's dq / dk stores have read the score buffer (long ago)
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        CPM_STAMP();                                       // 2: K, go, Rs landed; G', phi(K) done; CTA in step
        // ---- round 1: X = sum_b G'_b V_b^T
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < D; ++b) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_X, dG + b * TSTEP + 2 * k, dV + b * TSTEP + 2 * k, IDESC_KK128, (b | k) > 0);
            }
            mma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        CPM_STAMP();                                       // 3: round 1 done
        convert_lean<true, true, false>(g, TB_X, sX, gd, nullptr);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        CPM_STAMP();                                       // 4: X converted; CTA in step
        // ---- round 2: dQf_a = X Kf_a (+ sum_b G'_b Sp_ab^T) ; dKf_a = X^T Qf_a (+ sum_b v_b Rs_ab^T) ; PT = sum_a Kf_a Qf_a^T
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    mma_ss(tmem + TB_A1 + 64 * i, dX + (k >> 2) * TSTEP + 2 * (k & 3), dK + i * TSTEP + 128 * k, IDESC_KM64, k > 0);
                if (have_s) {
#pragma unroll
                    for (int b = 0; b < D; ++b) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_A1 + 64 * i, dG + b * TSTEP + 2 * k, dS + (i * D + b) * SSTEP + 2 * k, IDESC_KK64, 1);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int k = 0; k < 8; ++k) mma_ss(tmem + TB_A2 + 64 * i, dXT + 128 * k, dQ + i * TSTEP + 128 * k, IDESC_MM128, k > 0);
                if (have_r) {
#pragma unroll
                    for (int b = 0; b < D; ++b) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_A2 + 64 * i, dV + b * TSTEP + 2 * k, dR + (i * D + b) * SSTEP + 2 * k, IDESC_KK64, 1);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_X, dK + i * TSTEP + 2 * k, dQ + i * TSTEP + 2 * k, IDESC_KK128, (i | k) > 0);
            }
            mma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        CPM_STAMP();                                       // 5: round 2 done
        if (tid == 0 && tn < ntiles) issue_b(tn);          // Q, V, Sp are dead: refill them a round early
        // The transposed score tile goes to shared memory FIRST, so that round 3 can start; the dq / dk rows (accumulators A1 / A2,
        // untouched by round 3, whose dv lands on the just-converted score columns) are then written out while its UMMAs run.
        convert_lean<false, false, false>(g, TB_X, sX, 0.f, nullptr);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        CPM_STAMP();                                       // 6: PT converted; CTA in step
        // ---- round 3: dv_b = PT G'_b (+ sum_a Kf_a Rs_ab)
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < D; ++b) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    mma_ss(tmem + TB_X + 64 * b, dX + (k >> 2) * TSTEP + 2 * (k & 3), dG + b * TSTEP + 128 * k, IDESC_KM64, k > 0);
                if (have_r) {
#pragma unroll
                    for (int i = 0; i < D; ++i) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_X + 64 * b, dK + i * TSTEP + 2 * k, dR + (i * D + b) * SSTEP + 128 * k, IDESC_KM64, 1);
                    }
                }
            }
            mma_commit(bar_mma);
        }
        // dq / dk rows.  64-wide heads: finished in registers while round 3 runs; once it is done the score buffer is free and
        // stages them (swizzled, block 0 = dq, block 1 = dk) for two bulk tensor stores - whole lines instead of 32-byte pieces per lane.
        uint4 oq[4], ok[4];
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < D; ++i) {
            uint32_t r[32];
            tmem_ld32(g.t_lane + TB_A1 + 64 * i + 32 * g.half, r);
            tmem_ld_wait();
            grad_row_epilogue<true>(r, qfr[i], sz + 64 * i + 32 * g.half, gd, oq);
            if (!(TMA_OUT && valid == CHUNK) && g.row < valid) store_row32(a.gq, a.ld_g, grow + g.row, col0 + 64 * i + 32 * g.half, oq);
            tmem_ld32(g.t_lane + TB_A2 + 64 * i + 32 * g.half, r);
            tmem_ld_wait();
            grad_row_epilogue<true>(r, kfr[i], sgd + 64 * i + 32 * g.half, 1.f, ok);
            if (!(TMA_OUT && valid == CHUNK) && g.row < valid) store_row32(a.gk, a.ld_g, grow + g.row, col0 + 64 * i + 32 * g.half, ok);
        }
        if (tn < ntiles) inv_n = 1.f / inv_n;              // (prefetched den of the next tile)
        CPM_STAMP();                                       // 7: dq, dk rows computed (stored, without the bulk store)
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        CPM_STAMP();                                       // 8: round 3 done
        if (tid == 0 && tn < ntiles) issue_a(tn);
        if (TMA_OUT && valid == CHUNK) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const uint32_t off = sw128_off(g.row, 4 * g.half + cc);
                *reinterpret_cast<uint4 *>(sX + off) = oq[cc];
                *reinterpret_cast<uint4 *>(sX + TILE_BYTES + off) = ok[cc];
            }
            fence_proxy_async();
        }
#pragma unroll
        for (int b = 0; b < D; ++b) {   // dv rows
            uint32_t r[32];
            tmem_ld32(g.t_lane + TB_X + 64 * b + 32 * g.half, r);
            tmem_ld_wait();
            uint4 o[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) o[cc] = pack8u(r + 8 * cc, 1.f);
            if (g.row < valid) store_row32(a.gv, a.ld_g, grow + g.row, col0 + 64 * b + 32 * g.half, o);
        }
        tc_fence_before();
        __syncthreads();                                  // TMEM accumulators, sgd, sz are reused by the next tile
        tc_fence_after();
        if (TMA_OUT && valid == CHUNK && tid == 0) {
            tma_store_2d(&tmGq, sX, col0, grow);
            tma_store_2d(&tmGk, sX + TILE_BYTES, col0, grow);
            tma_store_commit();
        }
        CPM_STAMP();                                       // 9: dv rows stored; CTA in step
    }
#undef CPM_STAMP
    if (TMA_OUT && tid == 0) tma_store_wait_all0();
    if ((tid >> 5) == 0) tmem_dealloc<C::TCOLS>(tmem);
}

}  // namespace

// ---------------------------------------------------------------- host side
// Optional per-CTA phase timestamps (clock64) of cp_out_fwd_kernel for tools/; 64 slots per CTA.  NULL = off.
static long long *g_cp_timing = nullptr;
void linattn_cp_set_timing_buffer(long long *p) { g_cp_timing = p; }

// workspace: [increments fp32 NHC x Wide<D>::STATE_F][Sp region][Rs region][gd fp32 N*L*H];  width = 64 D
int64_t linattn_cp_workspace_bytes(int N, int L, int H, int width) {
    if (L <= 0 || (width != 64 && width != 128)) return 0;
    const int64_t nhc = (int64_t)N * H * ((L + CHUNK - 1) / CHUNK);
    if (width == 128) return nhc * Wide<2>::STATE_F * 4 + 2 * state_region_bytes_w<2>(nhc) + (int64_t)N * L * H * 4;
    return nhc * Wide<1>::STATE_F * 4 + 2 * state_region_bytes_w<1>(nhc) + (int64_t)N * L * H * 4;
}
int64_t linattn_cp_saved_bytes(int N, int L, int H, int width) {
    if (L <= 0 || (width != 64 && width != 128)) return 0;
    const int64_t nhc = (int64_t)N * H * ((L + CHUNK - 1) / CHUNK);
    return width == 128 ? state_region_bytes_w<2>(nhc) : state_region_bytes_w<1>(nhc);
}

namespace {
template <typename K> int smem_attr(K kernel, uint32_t bytes, const char *what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "%s smem attribute: %s", what, cudaGetErrorString(e));
    return CPM_OK;
}
template <int D> int set_attrs_once() {               // opt-in shared-memory sizes of every kernel of this head width
    static bool done = false;
    if (done) return CPM_OK;
    int rc;
    if ((rc = smem_attr(cp_state_fwd_kernel<D>, PreCfg<D>::F_SMEM, "cp_state_fwd"))) return rc;
    if ((rc = smem_attr(cp_state_bwd_kernel<D>, PreCfg<D>::B_SMEM, "cp_state_bwd"))) return rc;
    if ((rc = smem_attr(cp_out_fwd_kernel<D>, FwdCfg<D>::SMEM, "cp_out_fwd"))) return rc;
    if ((rc = smem_attr(cp_bwd_main_kernel<D>, BwdCfg<D>::SMEM, "cp_bwd_main"))) return rc;
    if (D == 1) {
        if ((rc = smem_attr(cp_prefix_stream_fwd_kernel, ST_SMEM, "cp_prefix_stream_fwd"))) return rc;
        if ((rc = smem_attr(cp_suffix_stream_bwd_kernel, SB_SMEM, "cp_suffix_stream_bwd"))) return rc;
    }
    done = true;
    return CPM_OK;
}
// the streaming state kernels (one CTA per chain, state carried in TMEM) exist for 64-wide heads; they need enough chains to fill the GPU
template <int D> bool use_stream(int N, int H) { return D == 1 && N * H >= 96; }

// F1 + F2 (or F1s) into a state region (tiles | z)
template <int D>
int prefix_states(const void *k, const void *v, int N, int L, int H, int64_t ld_qkv, float *part, uint8_t *region, cudaStream_t st) {
    const int nchunks = (L + CHUNK - 1) / CHUNK;
    if (nchunks <= 1) return CPM_OK;
    const int64_t nhc = (int64_t)N * H * nchunks;
    CUtensorMap tk, tv;
    int rc;
    const uint64_t rows = (uint64_t)N * L, inner = (uint64_t)H * 64 * D;
    if ((rc = make_tmap_bf16_2d(&tk, k, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tv, v, inner, rows, ld_qkv, CHUNK))) return rc;
    float *zs = reinterpret_cast<float *>(region + state_tiles_bytes_w<D>(nhc));
    if (use_stream<D>(N, H)) {
        cp_prefix_stream_fwd_kernel<<<N * H, NTH, ST_SMEM, st>>>(tk, tv, region, zs, L, H, nchunks);
        return check_launch("linattn_cp prefix stream");
    }
    cp_state_fwd_kernel<D><<<N * H * (nchunks - 1), 128, PreCfg<D>::F_SMEM, st>>>(tk, tv, part, L, H, nchunks);
    cp_scan_kernel<D><<<dim3(N * H, (Wide<D>::STATE_F / 4 + 255) / 256), 256, 0, st>>>(part, region, zs, nchunks, 0);
    return check_launch("linattn_cp prefix states");
}

template <int D>
int fwd_launch(const void *q, const void *k, const void *v, void *out, float *den, int N, int L, int H, int64_t ld_qkv, int64_t ld_o,
               float eps, void *ws, void *saved, cudaStream_t st) {
    const int nchunks = (L + CHUNK - 1) / CHUNK;
    const int64_t nhc = (int64_t)N * H * nchunks;
    if (nhc * 64 * D * D > 0x7fffffffLL) return CPM_ERR_UNSUPPORTED;
    float *part = reinterpret_cast<float *>(ws);
    uint8_t *region = saved ? reinterpret_cast<uint8_t *>(saved) : reinterpret_cast<uint8_t *>(ws) + nhc * Wide<D>::STATE_F * 4;
    int rc;
    if ((rc = set_attrs_once<D>())) return rc;
    if ((rc = prefix_states<D>(k, v, N, L, H, ld_qkv, part, region, st))) return rc;
    CUtensorMap tq, tk, tv, ts, to;
    const uint64_t rows = (uint64_t)N * L, inner = (uint64_t)H * 64 * D;
    if ((rc = make_tmap_bf16_2d(&tq, q, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tk, k, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tv, v, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&to, out, inner, rows, ld_o, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&ts, region, 64, (uint64_t)nhc * 64 * D * D, 64, 64))) return rc;
    const int64_t slots = (D == 1 ? 3ll : 1ll) * num_sms();
    cp_out_fwd_kernel<D><<<(unsigned)(nhc < slots ? nhc : slots), NTH, FwdCfg<D>::SMEM, st>>>(
        tq, tk, tv, ts, to, reinterpret_cast<const float *>(region + state_tiles_bytes_w<D>(nhc)), out, den, L, H, nchunks, N * H, ld_o, eps,
        g_cp_timing);
    return check_launch("linattn_fwd_cp");
}

template <int D>
int bwd_launch(const void *q, const void *k, const void *v, const void *out, const float *den, const void *gout, void *gq, void *gk,
               void *gv, int N, int L, int H, int64_t ld_qkv, int64_t ld_o, int64_t ld_g, void *ws, const void *saved, cudaStream_t st) {
    const int nchunks = (L + CHUNK - 1) / CHUNK;
    const int64_t nhc = (int64_t)N * H * nchunks;
    if (nhc * 64 * D * D > 0x7fffffffLL) return CPM_ERR_UNSUPPORTED;
    uint8_t *w8 = reinterpret_cast<uint8_t *>(ws);
    float *part = reinterpret_cast<float *>(w8);
    uint8_t *sp_region = w8 + nhc * Wide<D>::STATE_F * 4;
    uint8_t *rs_region = sp_region + state_region_bytes_w<D>(nhc);
    float *gd = reinterpret_cast<float *>(rs_region + state_region_bytes_w<D>(nhc));
    int rc;
    if ((rc = set_attrs_once<D>())) return rc;
    if (!saved) {                                       // forward did not keep its prefix states: rebuild them
        if ((rc = prefix_states<D>(k, v, N, L, H, ld_qkv, part, sp_region, st))) return rc;
    } else {
        sp_region = const_cast<uint8_t *>(reinterpret_cast<const uint8_t *>(saved));
    }
    CUtensorMap tq, tk, tv, tgo, to, ts, tr, tgq, tgk;
    const uint64_t rows = (uint64_t)N * L, inner = (uint64_t)H * 64 * D;
    if ((rc = make_tmap_bf16_2d(&tq, q, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tk, k, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tv, v, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tgo, gout, inner, rows, ld_o, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&to, out, inner, rows, ld_o, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tgq, gq, inner, rows, ld_g, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tgk, gk, inner, rows, ld_g, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&ts, sp_region, 64, (uint64_t)nhc * 64 * D * D, 64, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&tr, rs_region, 64, (uint64_t)nhc * 64 * D * D, 64, 64))) return rc;
    float *rzs = reinterpret_cast<float *>(rs_region + state_tiles_bytes_w<D>(nhc));
    if (use_stream<D>(N, H)) {
        cp_suffix_stream_bwd_kernel<<<N * H, NTH, SB_SMEM, st>>>(tq, tgo, to, den, gd, rs_region, rzs, L, H, nchunks);
    } else {
        cp_state_bwd_kernel<D><<<(unsigned)nhc, 128, PreCfg<D>::B_SMEM, st>>>(tq, tgo, to, den, gd, part, L, H, nchunks);
        if (nchunks > 1)
            cp_scan_kernel<D><<<dim3(N * H, (Wide<D>::STATE_F / 4 + 255) / 256), 256, 0, st>>>(part, rs_region, rzs, nchunks, 1);
    }
    BwdMainArgs a;
    a.den = den; a.gd = gd;
    a.zp = reinterpret_cast<const float *>(sp_region + state_tiles_bytes_w<D>(nhc));
    a.rzs = rzs;
    a.gq = gq; a.gk = gk; a.gv = gv; a.ld_g = ld_g; a.L = L; a.H = H; a.nchunks = nchunks; a.NH = N * H;
    a.dbg = g_cp_timing;
    const int64_t slots = (D == 1 ? 2ll : 1ll) * num_sms();
    cp_bwd_main_kernel<D><<<(unsigned)(nhc < slots ? nhc : slots), NTH, BwdCfg<D>::SMEM, st>>>(tq, tk, tv, tgo, ts, tr, tgq, tgk, a);
    return check_launch("linattn_bwd_cp");
}
}  // namespace

bool linattn_cp_streams(int N, int H, int width) { return width == 64 && N * H >= 96; }

int linattn_fwd_cp_launch(const void *q, const void *k, const void *v, void *out, float *den, int N, int L, int H, int width,
                          int64_t ld_qkv, int64_t ld_o, float eps, void *ws, void *saved, cudaStream_t st) {
    if (L <= 0 || (width != 64 && width != 128)) return CPM_ERR_UNSUPPORTED;
    return width == 128 ? fwd_launch<2>(q, k, v, out, den, N, L, H, ld_qkv, ld_o, eps, ws, saved, st)
                        : fwd_launch<1>(q, k, v, out, den, N, L, H, ld_qkv, ld_o, eps, ws, saved, st);
}

int linattn_bwd_cp_launch(const void *q, const void *k, const void *v, const void *out, const float *den, const void *gout,
                          void *gq, void *gk, void *gv, int N, int L, int H, int width, int64_t ld_qkv, int64_t ld_o, int64_t ld_g,
                          void *ws, const void *saved, cudaStream_t st) {
    if (L <= 0 || (width != 64 && width != 128)) return CPM_ERR_UNSUPPORTED;
    return width == 128 ? bwd_launch<2>(q, k, v, out, den, gout, gq, gk, gv, N, L, H, ld_qkv, ld_o, ld_g, ws, saved, st)
                        : bwd_launch<1>(q, k, v, out, den, gout, gq, gk, gv, N, L, H, ld_qkv, ld_o, ld_g, ws, saved, st);
}

}  // namespace cpm
