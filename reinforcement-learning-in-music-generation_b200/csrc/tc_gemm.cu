// The Linear layers of the CP transformer (in_linear, q/k/v/out projections, linear1/linear2, the output heads:
// agent_pretrain.py:239,244-253,337,360-375 and ft's AttentionLayer / TransformerEncoderLayer) as
// hand-written tcgen05 GEMMs for sm_100a.  bf16 operands, fp32 accumulation in tensor memory.
//
//   cpm_gemm_nt :  D[M x N] = epilogue( A[M x K] . B[N x K]^T )        forward (B = W) and data gradient (B = W^T copy)
//   cpm_gemm_tn :  dW[N x K] += dY[T x N]^T . X[T x K]  (fp32),  db[N] += column sums of dY      weight / bias gradient
//
// Both run on CTA PAIRS (cta_group::2): a cluster of two CTAs owns a 256 x 256 output tile; each CTA stages its 128 rows of
// A and its 128 rows of B per 64-wide K block (32 KB per CTA and stage, 5-stage TMA ring), the leader issues 256 x 256 x 16
// UMMAs for both SMs, each CTA's tensor memory holds its 128 x 256 half of the accumulator.  Compared with one-CTA tiles the
// B operand crosses L2 -> shared memory once per pair instead of once per CTA and every UMMA reads half as much shared
// memory per SM (B300_MICROARCH: 2-CTA mode is what reaches the tensor floor).
//
// cpm_gemm_nt is persistent (one pair per 2 SMs, static round-robin over tiles, N-tile fastest so that the A rows of a
// tile row are re-read from L2) with TWO accumulator stages of 256 TMEM columns: the epilogue of tile i (tcgen05.ld ->
// registers -> bf16 -> swizzled shared-memory staging -> TMA store) overlaps the UMMAs of tile i+1.
// Warp roles: 0 TMA producer, 1 UMMA issuer (leader CTA only), 2 TMEM allocator, 4-7 epilogue (TMEM lane quarter = warp % 4).
// Epilogues: bias; bias + exact-erf GELU + dropout writing BOTH the pre-activation (kept for backward) and the activation
// (what the gelu kernel did in a second pass over HBM); GELU backward (dgrad of linear2 times gelu'(h) and the regenerated
// dropout mask, what the gelu backward kernel did).
//
// cpm_gemm_tn reads both operands MN-major straight from the row-major activations (no transposes): the contraction index
// is the token, so a [64 tokens x 64 columns] TMA box is one SWIZZLE_128B MN-major atom column.  The token range is split
// over clusters (one (tile, split) work item per cluster) and the partial tiles are accumulated into the fp32 gradient with
// vector red.global.add - which is also the gradient ACCUMULATION across micro-batches.  The bias gradient comes from the
// same UMMA stream: one extra N = 16 instruction per K step against a constant tile of ones (db = dY^T . 1).
#include <stdlib.h>
#include "cpm_common.cuh"
#include "tc_common.cuh"

namespace cpm {
namespace {
using namespace tc;

constexpr int GM_THREADS = 384, GM_NS = 5;                                  // warps 0-3: TMA / UMMA / TMEM allocator / idle; 4-11 epilogue
constexpr uint32_t GM_A_BYTES = 16384, GM_STAGE = 32768;                 // per CTA: A half 128 x 64, B half 128 x 64 (bf16)
constexpr uint32_t GM_OFF_STG = GM_NS * GM_STAGE;                        // epilogue staging: 8 warps x 2 buffers x [32 rows x 128 B]
constexpr uint32_t GM_OFF_BAR = GM_OFF_STG + 65536;
constexpr uint32_t GM_SMEM = GM_OFF_BAR + 256;                           // full[5] empty[5] tfull[2] tempty[2] + TMEM slot

constexpr uint32_t IDESC_NT = idesc_bf16(256, 256, false, false);
constexpr uint32_t IDESC_TN = idesc_bf16(256, 256, true, true);

struct GemmNtArgs {
    const float *bias;                 // [N] fp32 or NULL
    const __nv_bfloat16 *aux;          // CPM_GEMM_EPI_DGELU: pre-activation h (M x N)
    int64_t ld_aux;
    int M, N, K;
    uint32_t thr8;                     // dropout of the GELU epilogues: 8 random bits per element, 16-element Philox groups
    float scale;                       //   (the streams of elementwise.cu's gelu kernel: fused and unfused paths draw the same masks)
    uint64_t seed, rng_offset;
    const unsigned long long *rng_base;
};

struct GemmTnArgs {
    float *dW[CPM_GEMM_TN_MAX_DST];    // destination(s): output rows [i * rows_per_dst, (i + 1) * rows_per_dst) -> dW[i] (rows_per_dst x K fp32)
    int64_t ldw;
    int T, N, K, rows_per_dst, splits, kb_per;
};

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void gemm_setup(uint8_t *sm, uint64_t *bars, uint32_t *tmem_slot, int tid, int warp, const CUtensorMap *m0,
                                           const CUtensorMap *m1, const CUtensorMap *m2, const CUtensorMap *m3) {
    uint64_t *bar_full = bars, *bar_empty = bars + GM_NS, *bar_tfull = bar_empty + GM_NS, *bar_tempty = bar_tfull + 2;
    cluster_sync_all();                                   // both CTAs of the pair are resident before anything touches the peer
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < GM_NS; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }     // full: the leader's expect_tx is the only arrival
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + s, 1); mbar_init(bar_tempty + s, 16); }     // tempty: 8 warps x 2 CTAs
        fence_barrier_init();
        tma_prefetch_desc(m0);
        tma_prefetch_desc(m1);
        if (m2) tma_prefetch_desc(m2);
        if (m3) tma_prefetch_desc(m3);
    }
    if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
}

// ---------------------------------------------------------------- epilogue value transforms (32 accumulator columns of one row)
template <int EPI>
__device__ __forceinline__ void nt_epilogue_half(const uint32_t (&r)[32], const GemmNtArgs &a, int64_t row, int n0, uint64_t rng_offset,
                                                 uint32_t (&o0)[16], uint32_t (&o1)[16]) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (EPI != CPM_GEMM_EPI_DGELU && a.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (n0 + j + 4 <= a.N) {
                const float4 b = __ldg(reinterpret_cast<const float4 *>(a.bias + n0 + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
        }
    }
    if (EPI == CPM_GEMM_EPI_BIAS) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o0[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
        return;
    }
    // ---- GELU forward / backward: two 16-element dropout groups per half
    uint32_t keep[2] = {0xFFFFu, 0xFFFFu};
    if (a.thr8) {
        const uint64_t g = (uint64_t)(row * a.N + n0) >> 4;
        keep[0] = dropout_keep16(a.seed, rng_offset, g, a.thr8);
        keep[1] = dropout_keep16(a.seed, rng_offset, g + 1, a.thr8);
    }
    if (EPI == CPM_GEMM_EPI_GELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t hb = pack_bf16(v[2 * j], v[2 * j + 1]);          // the stored pre-activation; GELU acts on what backward will read
            o0[j] = hb;
            const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xFFFF0000u);
            const float g0 = (keep[j >> 3] >> ((2 * j) & 15)) & 1u ? gelu_f<false>(h0) * a.scale : 0.f;
            const float g1 = (keep[j >> 3] >> ((2 * j + 1) & 15)) & 1u ? gelu_f<false>(h1) * a.scale : 0.f;
            o1[j] = pack_bf16(g0, g1);
        }
    } else {                                                                 // CPM_GEMM_EPI_DGELU
        uint32_t hw[16];
        const __nv_bfloat16 *hp = a.aux + row * a.ld_aux + n0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (row < a.M && n0 + 8 * q + 8 <= a.N) t = __ldg(reinterpret_cast<const uint4 *>(hp + 8 * q));
            hw[4 * q] = t.x; hw[4 * q + 1] = t.y; hw[4 * q + 2] = t.z; hw[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float h0 = __uint_as_float(hw[j] << 16), h1 = __uint_as_float(hw[j] & 0xFFFF0000u);
            const float g0 = (keep[j >> 3] >> ((2 * j) & 15)) & 1u ? v[2 * j] * dgelu_f<false>(h0) * a.scale : 0.f;
            const float g1 = (keep[j >> 3] >> ((2 * j + 1) & 15)) & 1u ? v[2 * j + 1] * dgelu_f<false>(h1) * a.scale : 0.f;
            o0[j] = pack_bf16(g0, g1);
        }
    }
}

__device__ __forceinline__ void stage_half(uint8_t *buf, int lane, int half, const uint32_t (&o)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4 *>(buf + sw128_off(lane, 4 * half + q)) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}

// =====================================================================================================================
// D = epilogue(A . B^T), persistent over 256 x 256 tiles
// =====================================================================================================================
template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GM_THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
               const __grid_constant__ CUtensorMap tmD2, const GemmNtArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + GM_OFF_BAR);
    uint64_t *bar_full = bars, *bar_empty = bars + GM_NS, *bar_tfull = bar_empty + GM_NS, *bar_tempty = bar_tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_tempty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int NB = (a.N + 255) >> 8, MB = (a.M + 255) >> 8, KB = (a.K + 63) >> 6, tiles = MB * NB;
    gemm_setup(sm, bars, tmem_slot, tid, warp, &tmA, &tmB, &tmD, EPI == CPM_GEMM_EPI_GELU ? &tmD2 : nullptr);
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer (one per CTA; transaction bytes land on the leader's barrier)
            const uint32_t full0 = mapa_u32(smem_u32(bar_full), 0);
            uint32_t s = 0, ph = 0;
            for (int t = pair; t < tiles; t += npairs) {
                const int nb = t % NB, mb = t / NB;
                const int row_a = mb * 256 + (int)rank * 128, row_b = nb * 256 + (int)rank * 128;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(bar_empty + s, ph ^ 1);
                    // only the leader arrives (announcing the bytes of both CTAs): a remote arrive per K block is a release at cluster
                    // scope and stalled the peer's producer for longer than the K block's UMMAs take (tools/probes/tma_ingest.cu)
                    if (rank == 0) mbar_expect_tx(bar_full + s, 2 * GM_STAGE);
                    tma_load_2d_2sm(sm + s * GM_STAGE, &tmA, full0 + 8 * s, kb * 64, row_a);
                    tma_load_2d_2sm(sm + s * GM_STAGE + GM_A_BYTES, &tmB, full0 + 8 * s, kb * 64, row_b);
                    if (++s == GM_NS) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {                      // ---- UMMA issuer (leader CTA, one thread, both SMs' tensor cores)
            uint32_t s = 0, ph = 0, ti = 0;
            for (int t = pair; t < tiles; t += npairs, ++ti) {
                const uint32_t as = ti & 1, aph = (ti >> 1) & 1;
                mbar_wait(bar_tempty + as, aph ^ 1);       // the epilogues of both CTAs drained this accumulator stage
                tc_fence_after();
                const uint32_t d = tmem + as * 256;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(bar_full + s, ph);
                    tc_fence_after();
                    const uint64_t dA = smem_desc_sw128(smem_u32(sm + s * GM_STAGE)), dB = smem_desc_sw128(smem_u32(sm + s * GM_STAGE + GM_A_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k) mma_ss_2sm(d, dA + 2 * k, dB + 2 * k, IDESC_NT, (kb > 0 || k > 0) ? 1u : 0u);
                    mma_commit_2sm(bar_empty + s, 3);      // frees this stage in BOTH CTAs once the UMMAs retire
                    if (++s == GM_NS) { s = 0; ph ^= 1; }
                }
                mma_commit_2sm(bar_tfull + as, 3);
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: 8 warps, two per TMEM lane quarter: warp owns accumulator rows [32 w, 32 w + 32) of this CTA's half tile and
        // the column half `ch` (two 64-column chunks).  With 4 warps the epilogue of a K = 512 tile (4.8 us) was longer than its UMMAs.
        const int w = warp & 3, ch = (warp - 4) >> 2;
        uint8_t *stg = sm + GM_OFF_STG + (warp - 4) * 8192;
        const uint32_t tempty0 = mapa_u32(smem_u32(bar_tempty), 0);
        const uint64_t rng_offset = rng_off(a.rng_offset, a.rng_base);
        uint32_t ti = 0, buf = 0;
        for (int t = pair; t < tiles; t += npairs, ++ti) {
            const uint32_t as = ti & 1, aph = (ti >> 1) & 1;
            const int nb = t % NB, mb = t / NB;
            const int grow0 = mb * 256 + (int)rank * 128 + w * 32;
            mbar_wait(bar_tfull + as, aph);
            tc_fence_after();
            const uint32_t tbase = tmem + ((uint32_t)(w * 32) << 16) + as * 256;
#pragma unroll 1
            for (int c = 2 * ch; c < 2 * ch + 2; ++c) {
                const int n0 = nb * 256 + c * 64;
                uint32_t r[32], o0[16], o1[16];
                const bool live = n0 < a.N && grow0 < a.M;                    // warp-uniform: the chunk holds at least one real element
                if (live) {                                                      // the staging buffer(s) must have been read by their last store
                    if (lane == 0) { if (EPI == CPM_GEMM_EPI_GELU) tma_store_wait_read0(); else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
                    __syncwarp();
                }
                uint8_t *b0 = stg + (EPI == CPM_GEMM_EPI_GELU ? 0u : buf * 4096u), *b1 = stg + 4096;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    tmem_ld32(tbase + c * 64 + h * 32, r);
                    tmem_ld_wait();
                    if (c == 2 * ch + 1 && h == 1) {                             // this warp's last read of the accumulator stage: hand it back to the issuer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * as);
                    }
                    if (live && n0 + h * 32 < a.N) {
                        nt_epilogue_half<EPI>(r, a, (int64_t)grow0 + lane, n0 + h * 32, rng_offset, o0, o1);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { o0[j] = 0u; o1[j] = 0u; }
                    }
                    if (live) {
                        stage_half(b0, lane, h, o0);
                        if (EPI == CPM_GEMM_EPI_GELU) stage_half(b1, lane, h, o1);
                    }
                }
                if (live) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {                                             // TMA clips the box at the tensor edge (ragged M, N)
                        tma_store_2d(&tmD, b0, n0, grow0);
                        if (EPI == CPM_GEMM_EPI_GELU) tma_store_2d(&tmD2, b1, n0, grow0);
                        tma_store_commit();
                    }
                    buf ^= 1;
                }
            }
        }
        if (lane == 0) tma_store_wait_all0();
    }
    tc_fence_before();
    cluster_sync_all();                                    // nobody leaves while the peer may still signal into its shared memory
    if (warp == 2) tmem_dealloc_2sm<512>(tmem);
}

// =====================================================================================================================
// Operand-stationary schedules.  Measured (profiles/r02_gemm_microbench_v1_stream.jsonl): the kernel above runs every layer
// shape at 660-680 TFLOP/s = 5.3 TB/s of L2 -> shared-memory operand traffic at its 128 flop/B (64 KB per pair and K block
// for 256 x 256 x 64 MACs) - the aggregate L2 -> SM ingest is the bound, not the tensor pipe.  Two schedules raise the
// The WIDE schedule raises the flop/B without changing the UMMA shape: one 256 x 512 tile per pair - all 512 TMEM columns, two
// N = 256 UMMAs per K step, 96 KB per pair and K block -> 175 flop/B.  Measured (profiles/r02_gemm_microbench_v2_modes.jsonl):
// 1000-1050 TFLOP/s on the N = 512 layers against 590-680 streamed.  Its price is a single accumulator stage (the epilogue
// of a tile is not hidden behind the next tile's UMMAs), paid back many times over at these K.  8 epilogue warps (two per
// TMEM lane quarter, splitting the columns), one 4 KB staging buffer per warp.
// (Also measured and dropped: an A-stationary schedule for K <= 512 - A resident in shared memory, only B streamed, 256
// flop/B on paper - ran SLOWER than streaming, 585 vs 660 TFLOP/s: with 128 KB pinned by A only 64 KB of B fits in flight per
// CTA, too little to cover the loaded L2 latency.  And GELU in the epilogue: see ops.py / profiles.)
// =====================================================================================================================
constexpr int G2_THREADS = 384;                                              // warps 0-3 as above, warps 4-11 epilogue
constexpr uint32_t G2_OFF_STG = 196608, G2_OFF_BAR = 229376, G2_SMEM = G2_OFF_BAR + 256;
constexpr int WD_NS = 4;                                                     // weight gradient: ring 4 x (A 16 KB + B 32 KB)
constexpr uint32_t WD_STAGE = 49152;
// wide forward tile: ring 3 x 48 KB | staging 8 warps x 2 x 4 KB (the TMA store of one 64-column chunk reads its buffer while the
// next chunk is converted into the other one)
constexpr int WN_NS = 3;
constexpr uint32_t WN_OFF_STG = WN_NS * WD_STAGE, WN_OFF_BAR = WN_OFF_STG + 65536, WN_SMEM = WN_OFF_BAR + 256;

// one 64-column chunk of this warp's 32 accumulator rows: TMEM -> epilogue -> swizzled staging -> TMA store
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(uint32_t taddr, const GemmNtArgs &a, const CUtensorMap *tmD, const CUtensorMap *tmD2, int grow0,
                                               int n0, int lane, uint8_t *stg, uint64_t rng_offset, bool last_read, uint32_t tempty_addr) {
    // `stg` is this chunk's 4 KB staging buffer: the caller alternates between two, so only the store before the previous one
    // must have finished reading (the GELU epilogue stores twice from one buffer and drains it completely)
    const bool live = n0 < a.N && grow0 < a.M;
    uint32_t r[32], o0[16], o1[16], g0[16];
    if (live) {
        if (lane == 0) { if (EPI == CPM_GEMM_EPI_GELU) tma_store_wait_read0(); else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
        __syncwarp();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        tmem_ld32(taddr + h * 32, r);
        tmem_ld_wait();
        if (last_read && h == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_addr);
        }
        if (live && n0 + h * 32 < a.N) {
            nt_epilogue_half<EPI>(r, a, (int64_t)grow0 + lane, n0 + h * 32, rng_offset, o0, o1);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { o0[j] = 0u; o1[j] = 0u; }
        }
        if (live) stage_half(stg, lane, h, o0);
        if (EPI == CPM_GEMM_EPI_GELU && h == 0) {         // second output (the activation) of half 0: parked until the first store has read the buffer
#pragma unroll
            for (int j = 0; j < 16; ++j) g0[j] = o1[j];
        }
    }
    if (!live) return;
    fence_proxy_async();
    __syncwarp();
    if (EPI == CPM_GEMM_EPI_GELU) {
        if (lane == 0) { tma_store_2d(tmD, stg, n0, grow0); tma_store_commit(); tma_store_wait_read0(); }
        __syncwarp();
        stage_half(stg, lane, 0, g0);
        stage_half(stg, lane, 1, o1);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { tma_store_2d(tmD2, stg, n0, grow0); tma_store_commit(); }
    } else {
        if (lane == 0) { tma_store_2d(tmD, stg, n0, grow0); tma_store_commit(); }
    }
}

__device__ __forceinline__ void g2_setup(uint8_t *sm, uint64_t *bars, int n_full_a, int n_ring, int tempty_count, uint32_t *tmem_slot, int tid,
                                         int warp, const CUtensorMap *m0, const CUtensorMap *m1, const CUtensorMap *m2, const CUtensorMap *m3) {
    // layout of `bars`: a_full[8] a_empty[8] full[4] empty[4] tfull[2] tempty[2]
    cluster_sync_all();
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < n_full_a; ++s) { mbar_init(bars + s, 1); mbar_init(bars + 8 + s, 1); }
        for (int s = 0; s < n_ring; ++s) { mbar_init(bars + 16 + s, 1); mbar_init(bars + 20 + s, 1); }      // full: the leader's expect_tx only (see below)
        for (int s = 0; s < 2; ++s) { mbar_init(bars + 24 + s, 1); mbar_init(bars + 26 + s, tempty_count); }
        fence_barrier_init();
        tma_prefetch_desc(m0);
        tma_prefetch_desc(m1);
        tma_prefetch_desc(m2);
        if (m3) tma_prefetch_desc(m3);
    }
    if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_nt_wide_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                    const __grid_constant__ CUtensorMap tmD2, const GemmNtArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + WN_OFF_BAR);
    uint64_t *bar_full = bars + 16, *bar_empty = bars + 20, *bar_tfull = bars + 24, *bar_tempty = bars + 26;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 28);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int MB = (a.M + 255) >> 8, KB = (a.K + 63) >> 6, NB2 = (a.N + 511) >> 9, tiles = MB * NB2;     // 256 x 512 tiles, N tile fastest
    g2_setup(sm, bars, 0, WN_NS, 8, tmem_slot, tid, warp, &tmA, &tmB, &tmD, EPI == CPM_GEMM_EPI_GELU ? &tmD2 : nullptr);
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer: A half + both 128-row halves of B's two 256-row slabs
            const uint32_t full0 = mapa_u32(smem_u32(bar_full), 0);
            uint32_t s = 0, ph = 0;
            for (int t = pair; t < tiles; t += npairs) {
                const int row_a = (t / NB2) * 256 + (int)rank * 128, row_b = (t % NB2) * 512 + (int)rank * 128;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(bar_empty + s, ph ^ 1);
                    if (rank == 0) mbar_expect_tx(bar_full + s, 2 * WD_STAGE);          // the leader alone arrives, for both CTAs' bytes
                    uint8_t *st = sm + s * WD_STAGE;
                    tma_load_2d_2sm(st, &tmA, full0 + 8 * s, kb * 64, row_a);
                    tma_load_2d_2sm(st + GM_A_BYTES, &tmB, full0 + 8 * s, kb * 64, row_b);
                    tma_load_2d_2sm(st + 2 * GM_A_BYTES, &tmB, full0 + 8 * s, kb * 64, row_b + 256);
                    if (++s == WN_NS) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {                      // ---- UMMA issuer: two N = 256 instructions per K step
            uint32_t s = 0, ph = 0, ti = 0;
            for (int t = pair; t < tiles; t += npairs, ++ti) {
                mbar_wait(bar_tempty + 0, (ti & 1) ^ 1);   // both column halves drained by the previous tile's epilogue
                mbar_wait(bar_tempty + 1, (ti & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(bar_full + s, ph);
                    tc_fence_after();
                    const uint32_t base = smem_u32(sm + s * WD_STAGE);
                    const uint64_t dA = smem_desc_sw128(base), dB0 = smem_desc_sw128(base + GM_A_BYTES), dB1 = smem_desc_sw128(base + 2 * GM_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        mma_ss_2sm(tmem, dA + 2 * k, dB0 + 2 * k, IDESC_NT, acc);
                        mma_ss_2sm(tmem + 256, dA + 2 * k, dB1 + 2 * k, IDESC_NT, acc);
                    }
                    mma_commit_2sm(bar_empty + s, 3);
                    if (++s == WN_NS) { s = 0; ph ^= 1; }
                }
                mma_commit_2sm(bar_tfull, 3);
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: lane quarter q, slab ch (columns 256 ch ..): four 64-column chunks per warp and tile
        const int q = warp & 3, ch = (warp - 4) >> 2;
        uint8_t *stg = sm + WN_OFF_STG + (warp - 4) * 8192;
        uint32_t buf = 0;
        const uint32_t tempty0 = mapa_u32(smem_u32(bar_tempty), 0);
        const uint64_t rng_offset = rng_off(a.rng_offset, a.rng_base);
        uint32_t ti = 0;
        for (int t = pair; t < tiles; t += npairs, ++ti) {
            const int grow0 = (t / NB2) * 256 + (int)rank * 128 + q * 32, ncol0 = (t % NB2) * 512 + ch * 256;
            mbar_wait(bar_tfull, ti & 1);
            tc_fence_after();
            const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + ch * 256;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                epilogue_chunk<EPI>(tbase + c * 64, a, &tmD, &tmD2, grow0, ncol0 + c * 64, lane, stg + (EPI == CPM_GEMM_EPI_GELU ? 0u : buf * 4096u),
                                    rng_offset, c == 3, tempty0 + 8 * ch);
                if (ncol0 + c * 64 < a.N && grow0 < a.M) buf ^= 1;          // a chunk that stored: the next one takes the other buffer
            }
        }
        if (lane == 0) tma_store_wait_all0();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm<512>(tmem);
}

// =====================================================================================================================
// Wide tile, bias epilogue, EARLY accumulator release.  The kernel above keeps the single 512-column accumulator busy for the
// whole epilogue (about 4 us per tile against 7 us of UMMAs at K = 512): each of its 8 epilogue warps converts, stages and
// TMA-stores chunk after chunk and hands the tensor memory back after its last read.  Here 16 epilogue warps (four per TMEM
// lane quarter, 128 columns each) first drain their part of the accumulator into 64 registers of packed bf16 (bias added on the
// way) and release the tensor memory - the next tile's UMMAs start then - and only afterwards swizzle the parked values
// into the staging buffer and store them.
// =====================================================================================================================
constexpr int G3_THREADS = 640;                                              // warps 0-3 as above, warps 4-19 epilogue

// PAIRS = 2: a cluster of FOUR CTAs = two pairs on neighbouring 256-row M tiles of the same 512-column N tile.  The B operand is
// the same for both pairs, so every CTA fetches only ONE of its two 128-row B slab halves and multicasts it to its counterpart
// in the other pair: 64 KB instead of 96 KB leave L2 per pair and K block (262 flop per L2 byte; the 2-CTA kernels saturate at
// about 5.3 TB/s of L2 -> shared-memory traffic, profiles/r02_summary.md section B).  A ring stage is free when BOTH pairs'
// UMMAs have retired (their commits are multicast to all four CTAs), so the two pairs walk the K loop in lock step.
template <int PAIRS>
__global__ void __cluster_dims__(2 * PAIRS, 1, 1) __launch_bounds__(G3_THREADS, 1)
gemm_nt_wide_bias_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                         const GemmNtArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + WN_OFF_BAR);
    uint64_t *bar_full = bars + 16, *bar_empty = bars + 20, *bar_tfull = bars + 24, *bar_tempty = bars + 26;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 28);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank(), rank = crank & 1u, pr = crank >> 1, lead = crank & ~1u;      // rank inside the pair; pair; leader CTA
    const int cl = blockIdx.x / (2 * PAIRS), ncl = gridDim.x / (2 * PAIRS);
    const int MBS = (a.M + 256 * PAIRS - 1) / (256 * PAIRS), KB = (a.K + 63) >> 6, NB2 = (a.N + 511) >> 9, tiles = MBS * NB2;
    cluster_sync_all();
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < WN_NS; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, PAIRS); }      // full: the leader's expect_tx is the only arrival
        mbar_init(bar_tfull, 1);
        for (int s = 0; s < 2; ++s) mbar_init(bar_tempty + s, 16);            // 8 warps x 2 CTAs per 256-column slab
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmD);
    }
    if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer
            const uint32_t full_lead = mapa_u32(smem_u32(bar_full), lead);
            const uint16_t mc_mask = (uint16_t)((1u << rank) | (1u << (rank + 2)));    // this CTA and its counterpart in the other pair
            uint32_t s = 0, ph = 0;
            for (int t = cl; t < tiles; t += ncl) {
                const int row_a = ((t / NB2) * PAIRS + (int)pr) * 256 + (int)rank * 128, row_b = (t % NB2) * 512 + (int)rank * 128;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(bar_empty + s, ph ^ 1);
                    // The leader announces the bytes of BOTH CTAs; the peer does not arrive at all (its bytes may land first - the
                    // transaction count then runs negative until the leader's expect_tx, the phase cannot complete before that
                    // arrival).  A remote arrive per K block is a release at cluster scope: it stalled the peer's producer for
                    // longer than a K block's UMMAs take (tools/probes/tma_ingest.cu).
                    if (rank == 0) mbar_expect_tx(bar_full + s, 2 * WD_STAGE);
                    uint8_t *st = sm + s * WD_STAGE;
                    tma_load_2d_2sm(st, &tmA, full_lead + 8 * s, kb * 64, row_a);
                    if (PAIRS == 1) {
                        tma_load_2d_2sm(st + GM_A_BYTES, &tmB, full_lead + 8 * s, kb * 64, row_b);
                        tma_load_2d_2sm(st + 2 * GM_A_BYTES, &tmB, full_lead + 8 * s, kb * 64, row_b + 256);
                    } else {                                // slab `pr` for both pairs
                        tma_load_2d_2sm_mc(st + GM_A_BYTES + pr * GM_A_BYTES, &tmB, full_lead + 8 * s, mc_mask, kb * 64, row_b + 256 * (int)pr);
                    }
                    if (++s == WN_NS) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {                      // ---- UMMA issuer (the leader of each pair)
            const uint16_t pair_mask = (uint16_t)(3u << (2 * pr)), all_mask = (uint16_t)((1u << (2 * PAIRS)) - 1u);
            uint32_t s = 0, ph = 0, ti = 0;
            for (int t = cl; t < tiles; t += ncl, ++ti) {
                mbar_wait(bar_tempty + 0, (ti & 1) ^ 1);
                mbar_wait(bar_tempty + 1, (ti & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(bar_full + s, ph);
                    tc_fence_after();
                    const uint32_t base = smem_u32(sm + s * WD_STAGE);
                    const uint64_t dA = smem_desc_sw128(base), dB0 = smem_desc_sw128(base + GM_A_BYTES), dB1 = smem_desc_sw128(base + 2 * GM_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        mma_ss_2sm(tmem, dA + 2 * k, dB0 + 2 * k, IDESC_NT, acc);
                        mma_ss_2sm(tmem + 256, dA + 2 * k, dB1 + 2 * k, IDESC_NT, acc);
                    }
                    mma_commit_2sm(bar_empty + s, all_mask);   // every CTA that writes into this stage anywhere in the cluster
                    if (++s == WN_NS) { s = 0; ph ^= 1; }
                }
                mma_commit_2sm(bar_tfull, pair_mask);
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: lane quarter q, column group cg (columns 128 cg .. 128 cg + 127 of the tile)
        const int q = warp & 3, cg = (warp - 4) >> 2;
        uint8_t *stg = sm + WN_OFF_STG + (warp - 4) * 4096;
        const uint32_t tempty_lead = mapa_u32(smem_u32(bar_tempty), lead);
        uint32_t ti = 0;
        for (int t = cl; t < tiles; t += ncl, ++ti) {
            const int grow0 = ((t / NB2) * PAIRS + (int)pr) * 256 + (int)rank * 128 + q * 32, ncol0 = (t % NB2) * 512 + cg * 128;
            mbar_wait(bar_tfull, ti & 1);
            tc_fence_after();
            const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + cg * 128;
            uint32_t pk[64];                                // this thread's row: 128 columns of packed bf16
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t r[16];
                tmem_ld16(tbase + 16 * i, r);
                tmem_ld_wait();
                const int n0 = ncol0 + 16 * i;
                if (a.bias) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        if (n0 + j + 4 <= a.N) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(a.bias + n0 + j));
                            r[j] = __float_as_uint(__uint_as_float(r[j]) + b.x);
                            r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + b.y);
                            r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + b.z);
                            r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + b.w);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[8 * i + j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            }
            tc_fence_before();                               // the accumulator columns are read: the next tile's UMMAs may overwrite them
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_lead + 8 * (cg >> 1));
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int n0 = ncol0 + 64 * c;
                if (n0 < a.N && grow0 < a.M) {              // warp-uniform
                    if (lane == 0) tma_store_wait_read0();   // the previous store has read the staging buffer
                    __syncwarp();
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch)
                        *reinterpret_cast<uint4 *>(stg + sw128_off(lane, ch)) =
                            make_uint4(pk[32 * c + 4 * ch], pk[32 * c + 4 * ch + 1], pk[32 * c + 4 * ch + 2], pk[32 * c + 4 * ch + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(&tmD, stg, n0, grow0); tma_store_commit(); }
                }
            }
        }
        if (lane == 0) tma_store_wait_all0();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm<512>(tmem);
}

// =====================================================================================================================
// dW += dY^T . X over one token range; one (256 x 512 tile, split) per cluster
// =====================================================================================================================
// Both operands are read MN-major straight from the row-major activations: the contraction index is the token, a [64 tokens x
// 64 columns] TMA box is one SWIZZLE_128B MN-major atom column (atoms 8192 B apart = LBO, 8-token groups 1024 B apart = SBO).
// Per stage and CTA: 128 output-feature columns of dY (2 boxes) and 256 input-feature columns of X (4 boxes) = 48 KB; two
// N = 256 UMMAs per K step fill all 512 TMEM columns (175 flop per operand byte, like the wide forward tile).  The epilogue
// adds the partial tile into the fp32 gradient with vector red.global.add - which is also the accumulation across
// micro-batches.  Output rows may be routed to several destination matrices (the q / k / v masters behind one fused GEMM).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const GemmTnArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + G2_OFF_BAR);
    uint64_t *bar_full = bars + 16, *bar_empty = bars + 20, *bar_tfull = bars + 24;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 28);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int item = blockIdx.x >> 1;
    const int NB2 = (a.K + 511) >> 9;                      // output tile columns index the INPUT features
    const int split = item % a.splits, tile = item / a.splits, nb = tile % NB2, mb = tile / NB2;
    const int kb_all = (a.T + 63) >> 6, kb0 = split * a.kb_per, kb1 = min(kb_all, kb0 + a.kb_per), KB = max(kb1 - kb0, 0);
    g2_setup(sm, bars, 0, WD_NS, 8, tmem_slot, tid, warp, &tmY, &tmX, &tmX, nullptr);
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer
            const uint32_t full0 = mapa_u32(smem_u32(bar_full), 0);
            const int col_y = mb * 256 + (int)rank * 128, col_x = nb * 512 + (int)rank * 128;
            uint32_t s = 0, ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(bar_empty + s, ph ^ 1);
                if (rank == 0) mbar_expect_tx(bar_full + s, 2 * WD_STAGE);              // the leader alone arrives, for both CTAs' bytes
                uint8_t *st = sm + s * WD_STAGE;
                tma_load_2d_2sm(st, &tmY, full0 + 8 * s, col_y, kb * 64);
                tma_load_2d_2sm(st + 8192, &tmY, full0 + 8 * s, col_y + 64, kb * 64);
#pragma unroll
                for (int j = 0; j < 2; ++j) {              // slab j = tile columns [256 j, 256 j + 256): this CTA's 128 of them
                    tma_load_2d_2sm(st + GM_A_BYTES + j * 16384, &tmX, full0 + 8 * s, col_x + 256 * j, kb * 64);
                    tma_load_2d_2sm(st + GM_A_BYTES + j * 16384 + 8192, &tmX, full0 + 8 * s, col_x + 256 * j + 64, kb * 64);
                }
                if (++s == WD_NS) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0 && KB > 0) {            // ---- UMMA issuer
            uint32_t s = 0, ph = 0;
            for (int i = 0; i < KB; ++i) {
                mbar_wait(bar_full + s, ph);
                tc_fence_after();
                const uint32_t base = smem_u32(sm + s * WD_STAGE);
                const uint64_t dA = smem_desc_sw128(base, 8192, 1024), dB0 = smem_desc_sw128(base + GM_A_BYTES, 8192, 1024),
                               dB1 = smem_desc_sw128(base + GM_A_BYTES + 16384, 8192, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) {              // 16 tokens further along K = 2048 B = 128 descriptor units
                    const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
                    mma_ss_2sm(tmem, dA + 128 * k, dB0 + 128 * k, IDESC_TN, acc);
                    mma_ss_2sm(tmem + 256, dA + 128 * k, dB1 + 128 * k, IDESC_TN, acc);
                }
                mma_commit_2sm(bar_empty + s, 3);
                if (++s == WD_NS) { s = 0; ph ^= 1; }
            }
            mma_commit_2sm(bar_tfull, 3);
        }
    } else if (warp >= 4 && KB > 0) {
        // ---- epilogue: lane quarter q, slab ch; accumulate this split's partial tile into the fp32 gradient(s)
        const int q = warp & 3, ch = (warp - 4) >> 2;
        const int m = mb * 256 + (int)rank * 128 + q * 32 + lane;            // output row = output feature
        mbar_wait(bar_tfull, 0);
        tc_fence_after();
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + ch * 256;
        const int dsti = m / a.rows_per_dst;
        float *dst = (m < a.N) ? a.dW[dsti] + (int64_t)(m - dsti * a.rows_per_dst) * a.ldw : nullptr;
        const int ncol0 = nb * 512 + ch * 256;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            uint32_t r[32];
            tmem_ld32(tbase + c * 32, r);
            tmem_ld_wait();
            const int n0 = ncol0 + c * 32;
            if (dst) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (n0 + j + 4 <= a.K)
                        red_add_v4(dst + n0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm<512>(tmem);
}

template <typename K>
int set_smem(K kernel, const char *name) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM);
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "%s shared-memory attribute: %s", name, cudaGetErrorString(e));
    return CPM_OK;
}

int g_gemm_mode = 0;          // 0 auto | 1 stream | 3 wide   (cpm_gemm_set_mode: A/B measurements)
const bool g_gemm_late_release = getenv("CPM_GEMM_LATE_RELEASE") != nullptr;      // A/B: the 8-warp epilogue that holds the accumulator
const bool g_gemm_no_multicast = getenv("CPM_GEMM_NO_MULTICAST") != nullptr;      // A/B: pairs fetch all of B themselves

template <int EPI>
int launch_nt(const CUtensorMap &tA, const CUtensorMap &tB, const CUtensorMap &tD, const CUtensorMap &tD2, const GemmNtArgs &a, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        int rc = set_smem(gemm_nt_kernel<EPI>, "gemm_nt");
        if (rc) return rc;
        cudaError_t e = cudaFuncSetAttribute(gemm_nt_wide_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WN_SMEM);
        if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "gemm_nt shared-memory attribute: %s", cudaGetErrorString(e));
        attr = true;
    }
    const int MB = (a.M + 255) / 256, NB = (a.N + 255) / 256, NB2 = (a.N + 511) / 512;
    static const int pair_cap = getenv("CPM_GEMM_MAX_PAIRS") ? atoi(getenv("CPM_GEMM_MAX_PAIRS")) : 0;      // A/B: leave SMs idle
    const int max_pairs = pair_cap > 0 ? min(pair_cap, num_sms() / 2) : num_sms() / 2;
    int mode = g_gemm_mode;
    // auto (interleaved A/B at the update shapes, tools/probes/gemm_modes_ab.py, after the full barriers lost their remote arrival): the
    // two-stage 256 x 256 schedule is as fast as or faster than the 256 x 512 one (its epilogue hides behind the next tile) except
    // for very short K, where halving the number of tiles pays
    if (mode != 1 && mode != 3) mode = (NB > 1 && a.K <= 384) ? 3 : 1;
    if (mode == 3) {
        if (EPI == CPM_GEMM_EPI_BIAS && !g_gemm_late_release) {
            static bool attr3 = false;
            static int max_quads = 0;                       // co-resident 4-CTA clusters (a GPC may leave SMs over)
            if (!attr3) {
                cudaError_t e = cudaFuncSetAttribute(gemm_nt_wide_bias_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WN_SMEM);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_nt_wide_bias_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WN_SMEM);
                if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "gemm_nt shared-memory attribute: %s", cudaGetErrorString(e));
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(4 * (num_sms() / 4));
                cfg.blockDim = dim3(G3_THREADS);
                cfg.dynamicSmemBytes = WN_SMEM;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                if (cudaOccupancyMaxActiveClusters(&max_quads, gemm_nt_wide_bias_kernel<2>, &cfg) != cudaSuccess) { max_quads = 0; cudaGetLastError(); }
                if (getenv("CPM_GEMM_VERBOSE")) fprintf(stderr, "cpmusic: gemm_nt wide: %d co-resident 4-CTA clusters on %d SMs\n", max_quads, num_sms());
                attr3 = true;
            }
            const int MBS = (a.M + 511) / 512;
            // the 4-CTA cluster pays when two M tiles exist per N tile and (nearly) every SM can be part of a quad
            if (!g_gemm_no_multicast && a.M > 256 && max_quads > 0 && (max_quads >= (num_sms() / 4) - 2 || getenv("CPM_GEMM_FORCE_MULTICAST"))) {
                gemm_nt_wide_bias_kernel<2><<<4 * min(MBS * NB2, max_quads), G3_THREADS, WN_SMEM, st>>>(tA, tB, tD, a);
                return check_launch("gemm_nt (wide, B multicast over two pairs)");
            }
            gemm_nt_wide_bias_kernel<1><<<2 * min(MB * NB2, max_pairs), G3_THREADS, WN_SMEM, st>>>(tA, tB, tD, a);
            return check_launch("gemm_nt (wide, early release)");
        }
        gemm_nt_wide_kernel<EPI><<<2 * min(MB * NB2, max_pairs), G2_THREADS, WN_SMEM, st>>>(tA, tB, tD, tD2, a);
        return check_launch("gemm_nt (wide)");
    }
    gemm_nt_kernel<EPI><<<2 * min(MB * NB, max_pairs), GM_THREADS, GM_SMEM, st>>>(tA, tB, tD, tD2, a);
    return check_launch("gemm_nt");
}

}  // namespace
}  // namespace cpm

using namespace cpm;

extern "C" int cpm_gemm_set_mode(int mode) {
    CPM_REQUIRE(mode >= 0 && mode <= 3, CPM_ERR_BAD_SHAPE, "gemm_set_mode: %d", mode);
    g_gemm_mode = mode;
    return CPM_OK;
}

extern "C" int cpm_gemm_nt(const void *A, int64_t lda, const void *B, int64_t ldb, void *D, int64_t ldd, void *D2, int64_t ldd2, int M, int N,
                           int K, const float *bias, int epilogue, const void *aux, int64_t ld_aux, float p_drop, uint64_t seed,
                           uint64_t rng_offset, void *stream) {
    CPM_REQUIRE(A && B && D, CPM_ERR_NULL, "gemm_nt: A/B/D must be non-NULL");
    CPM_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 8 == 0, CPM_ERR_BAD_SHAPE, "gemm_nt: M=%d N=%d K=%d (N, K multiples of 8)", M, N, K);
    CPM_REQUIRE(lda >= K && ldb >= K && ldd >= N && lda % 8 == 0 && ldb % 8 == 0 && ldd % 8 == 0, CPM_ERR_BAD_SHAPE, "gemm_nt: row strides must be >= the row width and multiples of 8 elements");
    CPM_REQUIRE(aligned16(A) && aligned16(B) && aligned16(D) && (!bias || aligned16(bias)), CPM_ERR_BAD_ALIGN, "gemm_nt: operands must be 16-byte aligned");
    CPM_REQUIRE(epilogue >= CPM_GEMM_EPI_BIAS && epilogue <= CPM_GEMM_EPI_DGELU, CPM_ERR_BAD_SHAPE, "gemm_nt: epilogue %d", epilogue);
    if (epilogue == CPM_GEMM_EPI_GELU)
        CPM_REQUIRE(D2 && ldd2 >= N && ldd2 % 8 == 0 && aligned16(D2) && N % 16 == 0 && ldd == N, CPM_ERR_BAD_SHAPE, "gemm_nt: the GELU epilogue needs a second output, N %% 16 == 0 and a dense pre-activation");
    if (epilogue == CPM_GEMM_EPI_DGELU)
        CPM_REQUIRE(aux && ld_aux == N && aligned16(aux) && N % 16 == 0, CPM_ERR_BAD_SHAPE, "gemm_nt: the GELU-backward epilogue needs the dense pre-activation (M x N)");
    CUtensorMap tA, tB, tD, tD2;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&tB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&tD, D, (uint64_t)N, (uint64_t)M, (uint64_t)ldd, 32))) return rc;
    tD2 = tD;
    if (epilogue == CPM_GEMM_EPI_GELU && (rc = make_tmap_bf16_2d(&tD2, D2, (uint64_t)N, (uint64_t)M, (uint64_t)ldd2, 32))) return rc;
    GemmNtArgs a;
    a.bias = bias; a.aux = (const __nv_bfloat16 *)aux; a.ld_aux = ld_aux; a.M = M; a.N = N; a.K = K;
    const bool drop = epilogue != CPM_GEMM_EPI_BIAS && p_drop > 0.f;
    a.thr8 = drop ? dropout_threshold8(p_drop) : 0u;
    a.scale = drop ? dropout_scale8(p_drop) : 1.f;
    a.seed = seed; a.rng_offset = rng_offset; a.rng_base = g_rng_base;
    cudaStream_t st = (cudaStream_t)stream;
    switch (epilogue) {
    case CPM_GEMM_EPI_BIAS: return launch_nt<CPM_GEMM_EPI_BIAS>(tA, tB, tD, tD2, a, st);
    case CPM_GEMM_EPI_GELU: return launch_nt<CPM_GEMM_EPI_GELU>(tA, tB, tD, tD2, a, st);
    default: return launch_nt<CPM_GEMM_EPI_DGELU>(tA, tB, tD, tD2, a, st);
    }
}

extern "C" int cpm_gemm_tn(const void *dY, int64_t ldy, const void *X, int64_t ldx, float *const *dW_host, int n_dst, int rows_per_dst,
                           int64_t ldw, int T, int N, int K, void *stream) {
    CPM_REQUIRE(dY && X && dW_host, CPM_ERR_NULL, "gemm_tn: dY/X/dW must be non-NULL");
    CPM_REQUIRE(T > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, CPM_ERR_BAD_SHAPE, "gemm_tn: T=%d N=%d K=%d (N, K multiples of 8)", T, N, K);
    CPM_REQUIRE(n_dst >= 1 && n_dst <= CPM_GEMM_TN_MAX_DST && rows_per_dst > 0 && (int64_t)n_dst * rows_per_dst >= N, CPM_ERR_BAD_SHAPE,
                "gemm_tn: %d destinations of %d rows for N=%d", n_dst, rows_per_dst, N);
    CPM_REQUIRE(ldy >= N && ldx >= K && ldw >= K && ldy % 8 == 0 && ldx % 8 == 0 && ldw % 4 == 0, CPM_ERR_BAD_SHAPE, "gemm_tn: row strides");
    CPM_REQUIRE(aligned16(dY) && aligned16(X), CPM_ERR_BAD_ALIGN, "gemm_tn: operands must be 16-byte aligned");
    GemmTnArgs a;
    for (int i = 0; i < CPM_GEMM_TN_MAX_DST; ++i) {
        a.dW[i] = i < n_dst ? dW_host[i] : nullptr;
        CPM_REQUIRE(i >= n_dst || (a.dW[i] && aligned16(a.dW[i])), CPM_ERR_BAD_ALIGN, "gemm_tn: destination %d must be non-NULL and 16-byte aligned", i);
    }
    CUtensorMap tY, tX;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tY, dY, (uint64_t)N, (uint64_t)T, (uint64_t)ldy, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&tX, X, (uint64_t)K, (uint64_t)T, (uint64_t)ldx, 64))) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM);
        if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "gemm_tn shared-memory attribute: %s", cudaGetErrorString(e));
        attr = true;
    }
    a.ldw = ldw; a.T = T; a.N = N; a.K = K; a.rows_per_dst = rows_per_dst;
    const int tiles = ((N + 255) / 256) * ((K + 511) / 512), kb_all = (T + 63) / 64;
    int splits = (num_sms() / 2 * 2) / tiles;              // two waves of clusters when the token range is long enough
    if (splits < 1) splits = 1;
    if (splits > (kb_all + 7) / 8) splits = (kb_all + 7) / 8;   // at least 8 K blocks per work item
    if (splits < 1) splits = 1;
    a.kb_per = (kb_all + splits - 1) / splits;
    a.splits = (kb_all + a.kb_per - 1) / a.kb_per;
    gemm_tn_kernel<<<2 * tiles * a.splits, G2_THREADS, G2_SMEM, (cudaStream_t)stream>>>(tY, tX, a);
    return check_launch("gemm_tn");
}
