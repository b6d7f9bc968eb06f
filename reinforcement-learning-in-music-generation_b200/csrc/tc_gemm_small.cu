// The Linear layers of the RECURRENT token step (M = songs in flight, 256 in the bench) on tcgen05:
//
//     D[M x N] = epilogue( A[M x K] . W[N x K]^T + bias )            bf16 in / out, fp32 accumulation in TMEM
//
// One token step of the rollout (testing-no-type-cp.py:157-167 batched) is a chain of ~90 dependent small kernels; each
// GEMM in it has 0.3 us of math and is bound by (a) the launch / drain gap to its neighbours and (b) how fast ONE SM can
// pull its operand tiles out of L2 (measured ~48 B/clk per SM).  This kernel is shaped for that regime, not for FLOPs:
//
//  * small tiles, many CTAs: 64 rows x 32 columns per CTA (UMMA M = 64, N = 32), so M = 256 spreads over 4 x N/32 CTAs and
//    the activation tile a CTA must ingest is 64 x K;
//  * < 100 KB of shared memory and 32 TMEM columns per CTA: two CTAs per SM, and - with programmatic dependent launch - the
//    NEXT kernel's CTAs become resident while this one still runs;
//  * PDL: everything that does not depend on the previous kernel (barrier init, TMEM allocation, tensor-map prefetch and the
//    whole WEIGHT tile stream) is issued before griddepcontrol.wait; only the activation fetch, the UMMAs and the epilogue stay
//    on the critical path.  griddepcontrol.launch_dependents is issued right after the set-up;
//  * epilogues: bias, or bias + exact-erf GELU on the bf16-rounded pre-activation (bit-identical to the GEMM + gelu kernel
//    pair of the teacher-forced path).
//
//  * LayerNorm without LayerNorm kernels (cpm_gemm_nt_small_ln, the rollout chain): a Linear that consumes LayerNorm(y) runs on
//    the RAW pre-norm rows y with gamma folded into the weights,
//        LN(y) W^T + b  =  rstd (y (gamma o W)^T - mean c1) + c2,    c1_n = sum_k (gamma o W)_nk,  c2_n = sum_k beta_k W_nk + b_n,
//    the row statistics coming from the activation tile the CTA holds in shared memory anyway (K <= 512: every k-block has its
//    own ring stage), computed by the epilogue warps while the UMMAs run; a Linear whose output is added to the residual stream
//    writes the pre-norm sum y' = (acc + b) + residual, the residual being either a plain tensor or LayerNorm(y_prev) rebuilt
//    from y_prev and the statistics its consumer left in global memory.
//
// K is streamed in 64-wide blocks through an 8-deep ring (12 KB per stage); for K <= 512 every block has its own stage and the
// producer never waits.  Warps 0-3: epilogue (M = 64 accumulator layout: warp w holds rows 16w..16w+15 in lanes 0..15),
// warp 4: TMA producer, warp 5: UMMA issuer.
#include "cpm_common.cuh"
#include "tc_common.cuh"

namespace cpm {
namespace {
using namespace tc;

constexpr int SG_BM = 64, SG_BN = 32, SG_NS = 8, SG_THREADS = 192;
#ifndef SG_GRP
#define SG_GRP 4
#endif
constexpr int GRP = SG_GRP;                                      // k-blocks per barrier in resident mode (K <= 512)
constexpr uint32_t SG_A_BYTES = SG_BM * 128, SG_W_BYTES = SG_BN * 128, SG_STAGE = SG_A_BYTES + SG_W_BYTES;      // 8 KB + 4 KB
constexpr uint32_t SG_OFF_BAR = SG_NS * SG_STAGE;                                                                // 98304
constexpr uint32_t SG_SMEM = SG_OFF_BAR + 256;
constexpr uint32_t IDESC_SG = idesc_bf16(SG_BM, SG_BN, false, false);

// development aid (cpm_debug_small_timing): CTA (0,0) of every launch appends 6 %globaltimer stamps - kernel entry, set-up done,
// griddepcontrol.wait returned, first activation block landed, accumulator ready, epilogue stored - to a device log
struct SmallTiming { unsigned long long count; unsigned long long stamps[1]; };
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct SmallArgs {
    const float *bias;
    __nv_bfloat16 *D;
    int64_t ldd;
    int M, N, K;
    // ---- LayerNorm folding (cpm_gemm_nt_small_ln); all NULL / 0 for the plain entry point
    const float *c1;                   // [N] fold: sum_k (gamma o W)_nk; `bias` then holds c2
    float2 *stats_out;                 // [M] (mean, rstd) of the rows of A, written by the CTAs of the first column tile
    float ln_eps;
    const __nv_bfloat16 *R;            // [M x N] residual added to the output (row stride ldr) ...
    int64_t ldr;
    const float2 *r_stats;             // ... through LayerNorm(R) with these row statistics and r_gamma / r_beta when non-NULL
    const float *r_gamma, *r_beta;
    SmallTiming *timing;               // NULL unless cpm_debug_small_timing installed a log
    int timing_cap;
};

// SPLIT: the two halves of K as a cluster (1, 1, 2) per output tile; slice 1 drops its fp32 tile into a buffer behind slice 0's
// ring (st.shared::cluster), one cluster barrier, slice 0 adds it and runs the epilogue.
template <int EPI, bool FOLD = false, bool RESID = false, bool SPLIT = false>
__global__ void __launch_bounds__(SG_THREADS, 2)
gemm_small_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const SmallArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(sm + SG_OFF_BAR), *bar_empty = bar_full + SG_NS, *bar_done = bar_empty + SG_NS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * SG_BN, m0 = blockIdx.y * SG_BM, KB = SPLIT ? (((a.K + 63) >> 6) >> 1) : ((a.K + 63) >> 6);
    const int kb0 = SPLIT ? (int)blockIdx.z * KB : 0;                  // SPLIT requires an even number of k-blocks
    __shared__ unsigned long long *s_log;
    if (a.timing && tid == 0) {
        s_log = nullptr;
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            const unsigned long long slot = atomicAdd(&a.timing->count, 1ull);
            if (slot < (unsigned long long)a.timing_cap) { s_log = a.timing->stamps + slot * 8; s_log[0] = gtimer(); s_log[6] = (unsigned long long)a.N; s_log[7] = (unsigned long long)a.K; }
        }
    }
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < SG_NS; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
        mbar_init(bar_done, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 4) tmem_alloc<32>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) griddep_launch();                      // the next kernel may start its own set-up / weight prefetch
    unsigned long long *logp = a.timing ? s_log : nullptr;
    if (logp && tid == 0) logp[1] = gtimer();

    if (warp == 4) {
        if (lane == 0) {                                 // ---- TMA producer
            const int pre = KB < SG_NS ? KB : SG_NS;
            // K <= 512: every k-block has its own stage and nothing is recycled, so the blocks report in GROUPS of four on the
            // barrier of the group's first stage - the UMMA thread (and the statistics pass) pays two mbarrier round trips, not eight
            const bool grouped = KB <= SG_NS;
            for (int i = 0; i < pre; ++i) {              // weights first: they do not depend on the previous kernel
                uint64_t *bar = bar_full + (grouped ? (i / GRP) * GRP : i);
                if (!grouped) mbar_expect_tx(bar, SG_STAGE);
                else if (i % GRP == 0) mbar_expect_tx(bar, (uint32_t)min(GRP, KB - i) * SG_STAGE);
                tma_load_2d(sm + i * SG_STAGE + SG_A_BYTES, &tmW, bar, (kb0 + i) * 64, n0);
            }
            griddep_wait();                              // the activations (and every buffer this kernel writes) belong to the chain
            if (logp) logp[2] = gtimer();
            for (int i = 0; i < pre; ++i) tma_load_2d(sm + i * SG_STAGE, &tmA, bar_full + (grouped ? (i / GRP) * GRP : i), (kb0 + i) * 64, m0);
            for (int i = pre; i < KB; ++i) {
                const int s = i % SG_NS;
                mbar_wait(bar_empty + s, ((i / SG_NS) - 1) & 1);
                mbar_expect_tx(bar_full + s, SG_STAGE);
                tma_load_2d(sm + s * SG_STAGE + SG_A_BYTES, &tmW, bar_full + s, (kb0 + i) * 64, n0);
                tma_load_2d(sm + s * SG_STAGE, &tmA, bar_full + s, (kb0 + i) * 64, m0);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {                                 // ---- UMMA issuer
            const uint64_t dA0 = smem_desc_sw128(smem_u32(sm)), dW0 = smem_desc_sw128(smem_u32(sm + SG_A_BYTES));
            if (KB <= SG_NS) {                           // resident: groups of four k-blocks per barrier, no stage recycling
                for (int i = 0; i < KB; ++i) {
                    if (i % GRP == 0) { mbar_wait(bar_full + i, 0); tc_fence_after(); if (logp && i == 0) logp[3] = gtimer(); }
                    mma_ss_kblock(tmem, dA0 + (uint64_t)(i * (SG_STAGE >> 4)), dW0 + (uint64_t)(i * (SG_STAGE >> 4)), IDESC_SG, i > 0 ? 1u : 0u);
                }
            } else {
                int s = 0, ph = 0;
                for (int i = 0; i < KB; ++i) {
                    mbar_wait(bar_full + s, ph);
                    tc_fence_after();
                    mma_ss_kblock(tmem, dA0 + (uint64_t)(s * (SG_STAGE >> 4)), dW0 + (uint64_t)(s * (SG_STAGE >> 4)), IDESC_SG, i > 0 ? 1u : 0u);
                    mma_commit(bar_empty + s);
                    if (++s == SG_NS) { s = 0; ph ^= 1; }
                }
            }
            mma_commit(bar_done);
        }
    } else {
        // ---- epilogue: lanes 0..15 of warp w own accumulator rows 16w + lane, 32 columns each
        const int row = m0 + 16 * warp + (lane & 15);
        float bv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) bv[j] = (a.bias && n0 + j < a.N) ? __ldg(a.bias + n0 + j) : 0.f;       // weights: no dependency
        float mean = 0.f, rstd = 1.f;
        if (FOLD) {
            // Row statistics of the raw activation tile, straight from the ring (K <= 512: block i sits in stage i): lanes l and
            // l + 16 take the two 64-byte halves of row 16 w + (l & 15) in every k-block; one pass (sum, sum of squares, fp32).
            const int r = 16 * warp + (lane & 15), half = lane >> 4;
            float sx = 0.f, sxx = 0.f;
            for (int i = 0; i < KB; ++i) {
                if (i % GRP == 0) mbar_wait(bar_full + i, 0);         // FOLD implies K <= 512: grouped barriers
                const uint8_t *rp = sm + i * SG_STAGE + r * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(rp + (((4 * half + c) ^ (r & 7)) << 4));
                    const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float x0 = __uint_as_float(w4[j] << 16), x1 = __uint_as_float(w4[j] & 0xFFFF0000u);
                        sx += x0 + x1;
                        sxx = fmaf(x0, x0, fmaf(x1, x1, sxx));
                    }
                }
            }
            sx += __shfl_xor_sync(0xffffffffu, sx, 16);
            sxx += __shfl_xor_sync(0xffffffffu, sxx, 16);
            mean = sx / (float)a.K;
            rstd = rsqrtf(fmaxf(sxx / (float)a.K - mean * mean, 0.f) + a.ln_eps);
            if (a.stats_out && blockIdx.x == 0 && lane < 16 && row < a.M) a.stats_out[row] = make_float2(mean, rstd);
        }
        float c1v[FOLD ? 32 : 1];
        if (FOLD) {
#pragma unroll
            for (int j = 0; j < 32; ++j) c1v[j] = (n0 + j < a.N) ? __ldg(a.c1 + n0 + j) : 0.f;
        }
        // the residual tile is chain data: fetched behind this thread's own griddepcontrol.wait, while the UMMAs run
        uint4 rraw[RESID ? 4 : 1];
        float2 rs = make_float2(0.f, 1.f);
        if (RESID) {
            griddep_wait();
            if (lane < 16 && row < a.M) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    rraw[q] = (n0 + 8 * q + 8 <= a.N) ? *reinterpret_cast<const uint4 *>(a.R + (int64_t)row * a.ldr + n0 + 8 * q) : make_uint4(0u, 0u, 0u, 0u);
                if (a.r_stats) {
                    // LayerNorm(R) rebuilt NOW, while the UMMAs run; the value is rounded to bf16 (as the LayerNorm kernel's output
                    // is), so it goes back into the same 16 registers without loss
                    rs = a.r_stats[row];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (n0 + 8 * q + 8 > a.N) continue;
                        const float4 g0 = __ldg(reinterpret_cast<const float4 *>(a.r_gamma + n0 + 8 * q)), g1 = __ldg(reinterpret_cast<const float4 *>(a.r_gamma + n0 + 8 * q + 4));
                        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(a.r_beta + n0 + 8 * q)), b1 = __ldg(reinterpret_cast<const float4 *>(a.r_beta + n0 + 8 * q + 4));
                        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        uint32_t w4[4] = {rraw[q].x, rraw[q].y, rraw[q].z, rraw[q].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float lo = (__uint_as_float(w4[j] << 16) - rs.x) * rs.y * gm[2 * j] + bt[2 * j];
                            const float hi = (__uint_as_float(w4[j] & 0xFFFF0000u) - rs.x) * rs.y * gm[2 * j + 1] + bt[2 * j + 1];
                            w4[j] = pack_bf16(lo, hi);
                        }
                        rraw[q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                    }
                }
            }
        }
        mbar_wait(bar_done, 0);                          // the UMMAs ran after the producer's griddepcontrol.wait: ordered behind the chain
        tc_fence_after();
        if (logp && tid == 0) logp[4] = gtimer();
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait();
        if (SPLIT) {
            const uint32_t rank = cluster_ctarank();
            uint8_t *part = sm + SG_SMEM;                  // 8 KB behind the ring and the barriers: slice 1's fp32 tile
            if (rank == 1 && lane < 16) {
                const uint32_t dst0 = mapa_u32(smem_u32(part) + (uint32_t)(16 * warp + lane) * 128u, 0);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + 16 * q), "r"(r[4 * q]), "r"(r[4 * q + 1]), "r"(r[4 * q + 2]),
                                 "r"(r[4 * q + 3])
                                 : "memory");
            }
            cluster_sync_all();
            if (rank == 0 && lane < 16) {
                const uint4 *src = reinterpret_cast<const uint4 *>(part + (uint32_t)(16 * warp + lane) * 128u);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 t = src[q];
                    r[4 * q] = __float_as_uint(__uint_as_float(r[4 * q]) + __uint_as_float(t.x));
                    r[4 * q + 1] = __float_as_uint(__uint_as_float(r[4 * q + 1]) + __uint_as_float(t.y));
                    r[4 * q + 2] = __float_as_uint(__uint_as_float(r[4 * q + 2]) + __uint_as_float(t.z));
                    r[4 * q + 3] = __float_as_uint(__uint_as_float(r[4 * q + 3]) + __uint_as_float(t.w));
                }
            }
        }
        if ((!SPLIT || cluster_ctarank() == 0) && lane < 16 && row < a.M) {
            __nv_bfloat16 *dst = a.D + (int64_t)row * a.ldd + n0;
            uint4 ov[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t w[4];
                float res[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (RESID && n0 + 8 * q + 8 <= a.N) {    // the residual stream: a plain tensor, or LayerNorm(R) rebuilt (bf16, like the LayerNorm kernel's output)
                    const uint4 rq = rraw[q];
                    const uint32_t w4[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) { res[2 * j] = __uint_as_float(w4[j] << 16); res[2 * j + 1] = __uint_as_float(w4[j] & 0xFFFF0000u); }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x0 = __uint_as_float(r[8 * q + 2 * j]), x1 = __uint_as_float(r[8 * q + 2 * j + 1]);
                    if (FOLD) {
                        x0 = rstd * (x0 - mean * c1v[8 * q + 2 * j]);
                        x1 = rstd * (x1 - mean * c1v[8 * q + 2 * j + 1]);
                    }
                    x0 += bv[8 * q + 2 * j];
                    x1 += bv[8 * q + 2 * j + 1];
                    if (RESID) {
                        x0 = bf16_round(x0) + res[2 * j];
                        x1 = bf16_round(x1) + res[2 * j + 1];
                    }
                    if (EPI == CPM_GEMM_EPI_GELU) {      // GELU of the bf16-rounded pre-activation, like the GEMM + gelu kernel pair
                        x0 = gelu_f<false>(bf16_round(x0));
                        x1 = gelu_f<false>(bf16_round(x1));
                    }
                    w[j] = pack_bf16(x0, x1);
                }
                ov[q] = make_uint4(w[0], w[1], w[2], w[3]);
            }
#ifndef CPM_SMALL_STORE256
#define CPM_SMALL_STORE256 1
#endif
            // the row's 64 bytes as two 256-bit stores (whole sectors) where the tile is complete and the row 32-byte aligned
            if (CPM_SMALL_STORE256 && n0 + 32 <= a.N && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 16 * i), "r"(ov[2 * i].x), "r"(ov[2 * i].y),
                                 "r"(ov[2 * i].z), "r"(ov[2 * i].w), "r"(ov[2 * i + 1].x), "r"(ov[2 * i + 1].y), "r"(ov[2 * i + 1].z),
                                 "r"(ov[2 * i + 1].w)
                                 : "memory");
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (n0 + 8 * q + 8 <= a.N) *reinterpret_cast<uint4 *>(dst + 8 * q) = ov[q];
            }
        }
    }
    if (SPLIT && warp >= 4) {                            // the producer / UMMA warps take part in the cluster barrier of the reduction
        __syncwarp();
        cluster_sync_all();
    }
    tc_fence_before();
    __syncthreads();
    if (logp && tid == 0) logp[5] = gtimer();
    if (warp == 4) tmem_dealloc<32>(tmem);
}

SmallTiming *g_small_timing = nullptr;
int g_small_timing_cap = 0;

int g_small_split = 1;        // cpm_gemm_small_set_split(0): never split K (A/B runs)

template <int EPI, bool FOLD = false, bool RESID = false>
int launch_small(const CUtensorMap &tA, const CUtensorMap &tW, const SmallArgs &a, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_small_kernel<EPI, FOLD, RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM);
        if (e == cudaSuccess && !FOLD)
            e = cudaFuncSetAttribute(gemm_small_kernel<EPI, false, RESID, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM + 8192);
        if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "gemm_small shared-memory attribute: %s", cudaGetErrorString(e));
        attr = true;
    }
    const int gx = (a.N + SG_BN - 1) / SG_BN, gy = (a.M + SG_BM - 1) / SG_BM, KB_all = (a.K + 63) / 64;
    const dim3 grid(gx, gy);
    SmallArgs b = a;
    b.timing = g_small_timing;
    b.timing_cap = g_small_timing_cap;
    cudaError_t e;
    // long K on few output tiles (linear2: 64 CTAs x 128 small UMMAs x 384 KB of operands): two K slices per tile as a cluster
    if (!FOLD && g_small_split && KB_all >= 16 && KB_all % 2 == 0 && 2 * gx * gy <= 2 * num_sms())
        e = launch_chain_cluster(gemm_small_kernel<EPI, false, RESID, true>, dim3(gx, gy, 2), dim3(SG_THREADS), SG_SMEM + 8192, st, dim3(1, 1, 2), tA, tW, b);
    else
        e = launch_chain(gemm_small_kernel<EPI, FOLD, RESID>, grid, dim3(SG_THREADS), SG_SMEM, st, tA, tW, b);
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "gemm_small launch: %s", cudaGetErrorString(e));
    return CPM_OK;
}

}  // namespace
}  // namespace cpm

using namespace cpm;

extern "C" int cpm_gemm_nt_small(const void *A, int64_t lda, const void *W, int64_t ldw, void *D, int64_t ldd, int M, int N, int K,
                                 const float *bias, int epilogue, void *stream) {
    CPM_REQUIRE(A && W && D, CPM_ERR_NULL, "gemm_nt_small: A/W/D must be non-NULL");
    CPM_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 8 == 0, CPM_ERR_BAD_SHAPE, "gemm_nt_small: M=%d N=%d K=%d (N, K multiples of 8)", M, N, K);
    CPM_REQUIRE(lda >= K && ldw >= K && ldd >= N && lda % 8 == 0 && ldw % 8 == 0 && ldd % 8 == 0, CPM_ERR_BAD_SHAPE, "gemm_nt_small: row strides");
    CPM_REQUIRE(aligned16(A) && aligned16(W) && aligned16(D), CPM_ERR_BAD_ALIGN, "gemm_nt_small: operands must be 16-byte aligned");
    CPM_REQUIRE(epilogue == CPM_GEMM_EPI_BIAS || epilogue == CPM_GEMM_EPI_GELU, CPM_ERR_BAD_SHAPE, "gemm_nt_small: epilogue %d", epilogue);
    CUtensorMap tA, tW;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, SG_BM))) return rc;
    if ((rc = make_tmap_bf16_2d(&tW, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, SG_BN))) return rc;
    SmallArgs a = {};
    a.bias = bias; a.D = (__nv_bfloat16 *)D; a.ldd = ldd; a.M = M; a.N = N; a.K = K;
    cudaStream_t st = (cudaStream_t)stream;
    if (epilogue == CPM_GEMM_EPI_GELU) return launch_small<CPM_GEMM_EPI_GELU>(tA, tW, a, st);
    return launch_small<CPM_GEMM_EPI_BIAS>(tA, tW, a, st);
}

extern "C" int cpm_gemm_small_set_split(int on) {
    g_small_split = on ? 1 : 0;
    return CPM_OK;
}

extern "C" int cpm_debug_small_timing(void *device_log, int capacity) {
    g_small_timing = (SmallTiming *)device_log;
    g_small_timing_cap = device_log ? capacity : 0;
    return CPM_OK;
}

extern "C" int cpm_gemm_nt_small_ln(const void *A, int64_t lda, const void *W, int64_t ldw, void *D, int64_t ldd, int M, int N, int K,
                                    const float *bias, int epilogue, const float *fold_c1, float *stats_out, float ln_eps, const void *R, int64_t ldr,
                                    const float *r_stats, const float *r_gamma, const float *r_beta, void *stream) {
    CPM_REQUIRE(A && W && D, CPM_ERR_NULL, "gemm_nt_small_ln: A/W/D must be non-NULL");
    CPM_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 8 == 0, CPM_ERR_BAD_SHAPE, "gemm_nt_small_ln: M=%d N=%d K=%d (N, K multiples of 8)", M, N, K);
    CPM_REQUIRE(lda >= K && ldw >= K && ldd >= N && lda % 8 == 0 && ldw % 8 == 0 && ldd % 8 == 0, CPM_ERR_BAD_SHAPE, "gemm_nt_small_ln: row strides");
    CPM_REQUIRE(aligned16(A) && aligned16(W) && aligned16(D), CPM_ERR_BAD_ALIGN, "gemm_nt_small_ln: operands must be 16-byte aligned");
    CPM_REQUIRE(epilogue == CPM_GEMM_EPI_BIAS || epilogue == CPM_GEMM_EPI_GELU, CPM_ERR_BAD_SHAPE, "gemm_nt_small_ln: epilogue %d", epilogue);
    const bool fold = fold_c1 != nullptr, resid = R != nullptr;
    CPM_REQUIRE(fold != resid, CPM_ERR_BAD_SHAPE, "gemm_nt_small_ln: exactly one of the LayerNorm fold (fold_c1) and the residual (R) must be given");
    if (fold) CPM_REQUIRE(K <= 64 * SG_NS && bias, CPM_ERR_BAD_SHAPE, "gemm_nt_small_ln: the fold needs K <= %d and the folded bias c2", 64 * SG_NS);
    if (resid) {
        CPM_REQUIRE(epilogue == CPM_GEMM_EPI_BIAS && ldr >= N && ldr % 8 == 0 && aligned16(R), CPM_ERR_BAD_SHAPE, "gemm_nt_small_ln: residual layout");
        CPM_REQUIRE(!r_stats || (r_gamma && r_beta), CPM_ERR_NULL, "gemm_nt_small_ln: r_stats needs r_gamma and r_beta");
    }
    CUtensorMap tA, tW;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, SG_BM))) return rc;
    if ((rc = make_tmap_bf16_2d(&tW, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, SG_BN))) return rc;
    SmallArgs a = {};
    a.bias = bias; a.D = (__nv_bfloat16 *)D; a.ldd = ldd; a.M = M; a.N = N; a.K = K;
    a.c1 = fold_c1; a.stats_out = (float2 *)stats_out; a.ln_eps = ln_eps;
    a.R = (const __nv_bfloat16 *)R; a.ldr = ldr; a.r_stats = (const float2 *)r_stats; a.r_gamma = r_gamma; a.r_beta = r_beta;
    cudaStream_t st = (cudaStream_t)stream;
    if (resid) return launch_small<CPM_GEMM_EPI_BIAS, false, true>(tA, tW, a, st);
    if (epilogue == CPM_GEMM_EPI_GELU) return launch_small<CPM_GEMM_EPI_GELU, true, false>(tA, tW, a, st);
    return launch_small<CPM_GEMM_EPI_BIAS, true, false>(tA, tW, a, st);
}
