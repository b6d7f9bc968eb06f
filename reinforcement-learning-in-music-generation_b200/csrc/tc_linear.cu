// Small-M Linear layers of the rollout token step on tcgen05 / TMA (sm_100a):
//
//     Y[M x N] = epilogue( A[M x K] . W[N x K]^T )          bf16 in / out, fp32 accumulation in TMEM
//
// One token step of 256 songs is a chain of ~100 dependent small kernels (M = 256 rows); every LayerNorm,
// GELU and residual add between the GEMMs costs a launch of its own (~3-5 us each, pure latency).  This
// kernel removes them by folding them into the GEMM:
//
//  * LayerNorm of the INPUT is folded algebraically (no prologue pass over A):
//        LN(a) . W^T = rstd * ( a . (gamma (.) W)^T  -  mean * c1 ) + c2 ,   c1[n] = sum_k gamma_k W[n,k],
//                                                                            c2[n] = sum_k beta_k  W[n,k] + bias[n]
//    so the UMMA runs on the raw pre-LayerNorm sums with the pre-scaled weight W' = gamma (.) W, and the
//    epilogue applies the per-row (mean, rstd) and per-column (c1, c2) terms.
//  * The per-row statistics come from the PRODUCER: every kernel that writes a pre-LayerNorm sum also writes,
//    per N-tile, the partial (sum, sum of squares) of the bf16 values it stored; consumers add the partials in
//    a fixed order (deterministic, no atomics).
//  * A residual that is itself a LayerNorm output is recomputed on the fly from (pre-LN sum, stats, gamma, beta),
//    so the LayerNorm output is never materialised.
//  * bias / exact GELU / residual / positional encoding are epilogue variants.
//
// CTA = 128 rows x BN columns, K streamed in 64-wide blocks through an NS-stage TMA ring (the whole K extent is
// resident for K <= 512).  Warps 0-3: epilogue (one thread per row), warp 4: TMA producer, warp 5: UMMA issuer.
// With PDL (programmatic dependent launch) the weight tiles - which do not depend on the previous kernel - are
// fetched before griddepcontrol.wait, so only the activation fetch, the UMMAs and the epilogue remain on the
// token-step critical path.
#include "cpm_common.cuh"
#include "tc_common.cuh"

namespace cpm {
namespace {
using namespace tc;

constexpr int TL_BM = 128, TL_BK = 64, TL_NS = 8;
constexpr uint32_t TL_A_BYTES = TL_BM * 128;                     // [128 rows x 64 bf16] SW128 K-major block

struct TcLinearArgs {
    const float *c1, *c2;                   // per output column; c1 == NULL: no LayerNorm fold (c2 = bias, may be NULL)
    const float *stats_in;                  // [M][parts_in][2] partial (sum, sumsq) of A's rows
    const __nv_bfloat16 *R;                 // residual source rows (plain values, or pre-LN sums for RES_LN)
    const float *stats_r, *gamma_r, *beta_r;
    const float *pe;
    const int *pos_dev;
    __nv_bfloat16 *Y;
    float *stats_out;                       // [M][gridDim.x][2] or NULL
    int64_t ldr, ldy;
    int M, N, K, epi, parts_in, parts_r, pe_max, pos_offset, use_pdl;
    float eps;
};

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ void row_stats(const float *stats, int parts, int64_t m, int width, float eps, float &mean, float &rstd) {
    float s = 0.f, q = 0.f;
    const float2 *p = reinterpret_cast<const float2 *>(stats) + m * parts;
    for (int i = 0; i < parts; ++i) { const float2 v = p[i]; s += v.x; q += v.y; }
    mean = s / (float)width;
    rstd = rsqrtf(fmaxf(q / (float)width - mean * mean, 0.f) + eps);
}

template <int BN>
__global__ void __launch_bounds__(192)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, TcLinearArgs a) {
    constexpr uint32_t W_BYTES = BN * 128, STAGE = TL_A_BYTES + W_BYTES;
    constexpr uint32_t IDESC = idesc_bf16(128, BN, false, false);
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(sm + TL_NS * STAGE), *bar_empty = bar_full + TL_NS, *bar_done = bar_empty + TL_NS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_done + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TL_BM;
    const int KB = a.K / TL_BK;
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < TL_NS; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        if ((tid & 31) == 0) {                       // ---- TMA producer
            const int pre = KB < TL_NS ? KB : TL_NS;
            for (int kb = 0; kb < pre; ++kb) {       // weights first: they do not depend on the previous kernel
                mbar_expect_tx(bar_full + kb, STAGE);
                tma_load_2d(sm + kb * STAGE + TL_A_BYTES, &tmW, bar_full + kb, kb * TL_BK, n0);
            }
            if (a.use_pdl) griddep_wait();
            for (int kb = 0; kb < pre; ++kb) tma_load_2d(sm + kb * STAGE, &tmA, bar_full + kb, kb * TL_BK, m0);
            for (int kb = pre; kb < KB; ++kb) {
                const int s = kb % TL_NS;
                mbar_wait(bar_empty + s, ((kb / TL_NS) - 1) & 1);
                mbar_expect_tx(bar_full + s, STAGE);
                tma_load_2d(sm + s * STAGE + TL_A_BYTES, &tmW, bar_full + s, kb * TL_BK, n0);
                tma_load_2d(sm + s * STAGE, &tmA, bar_full + s, kb * TL_BK, m0);
            }
        }
    } else if (warp == 5) {
        if ((tid & 31) == 0) {                       // ---- UMMA issuer
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % TL_NS;
                mbar_wait(bar_full + s, (kb / TL_NS) & 1);
                tc_fence_after();
                const uint64_t dA = smem_desc_sw128(smem_u32(sm + s * STAGE)), dW = smem_desc_sw128(smem_u32(sm + s * STAGE + TL_A_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss(tmem, dA + 2 * k, dW + 2 * k, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                mma_commit(bar_empty + s);            // the stage is free once these UMMAs retire
            }
            mma_commit(bar_done);
        }
    } else {
        // ---- epilogue: thread = output row
        if (a.use_pdl) griddep_wait();
        const int m = m0 + tid;
        const bool row_ok = m < a.M;
        float mean = 0.f, rstd = 1.f, mean_r = 0.f, rstd_r = 1.f;
        if (row_ok && a.c1) row_stats(a.stats_in, a.parts_in, m, a.K, a.eps, mean, rstd);
        if (row_ok && a.epi == CPM_TL_EPI_RES_LN) row_stats(a.stats_r, a.parts_r, m, a.N, a.eps, mean_r, rstd_r);
        const float *pe_row = nullptr;
        if (a.epi == CPM_TL_EPI_PE) {
            int pos = a.pos_offset + (a.pos_dev ? *a.pos_dev : 0);
            pos = pos < a.pe_max ? pos : a.pe_max - 1;
            pe_row = a.pe + (int64_t)pos * a.N;
        }
        mbar_wait(bar_done, 0);
        tc_fence_after();
        if (a.use_pdl) griddep_launch();             // the next kernel may start fetching its weights
        const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
        float ssum = 0.f, ssq = 0.f;
#pragma unroll
        for (int p = 0; p < BN / 32; ++p) {
            uint32_t r[32];
            tmem_ld32(t_lane + 32 * p, r);
            tmem_ld_wait();
            const int nb = n0 + 32 * p;
            if (row_ok && nb < a.N) {
                uint4 rres[4];
                if (a.epi == CPM_TL_EPI_RES || a.epi == CPM_TL_EPI_RES_LN) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) rres[i] = *reinterpret_cast<const uint4 *>(a.R + (int64_t)m * a.ldr + nb + 8 * i);
                }
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float v[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int n = nb + i + j;
                        float x = __uint_as_float(r[i + j]);
                        const float b = a.c2 ? __ldg(a.c2 + n) : 0.f;
                        x = a.c1 ? fmaf(rstd, x - mean * __ldg(a.c1 + n), b) : x + b;
                        if (a.epi == CPM_TL_EPI_GELU) x = gelu_erf(x);
                        if (a.epi == CPM_TL_EPI_RES || a.epi == CPM_TL_EPI_RES_LN) {
                            const uint32_t u = reinterpret_cast<const uint32_t *>(rres)[(i + j) >> 1];
                            float rr = __uint_as_float(((i + j) & 1) ? (u & 0xffff0000u) : (u << 16));
                            if (a.epi == CPM_TL_EPI_RES_LN) rr = fmaf((rr - mean_r) * rstd_r, __ldg(a.gamma_r + n), __ldg(a.beta_r + n));
                            x += rr;
                        }
                        if (a.epi == CPM_TL_EPI_PE) x += __ldg(pe_row + n);
                        v[j] = x;
                    }
                    const __nv_bfloat162 hb = __floats2bfloat162_rn(v[0], v[1]);
                    const float2 fb = __bfloat1622float2(hb);          // statistics of what is actually stored
                    ssum += fb.x + fb.y;
                    ssq = fmaf(fb.x, fb.x, fmaf(fb.y, fb.y, ssq));
                    w[i >> 1] = *reinterpret_cast<const uint32_t *>(&hb);
                }
                uint4 *dst = reinterpret_cast<uint4 *>(a.Y + (int64_t)m * a.ldy + nb);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
            }
        }
        if (row_ok && a.stats_out) reinterpret_cast<float2 *>(a.stats_out)[(int64_t)m * gridDim.x + blockIdx.x] = make_float2(ssum, ssq);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem);
}

template <int BN>
int launch_tc_linear(const CUtensorMap &tA, const CUtensorMap &tW, const TcLinearArgs &a, cudaStream_t st) {
    constexpr uint32_t SMEM = TL_NS * (TL_A_BYTES + BN * 128) + (2 * TL_NS + 1) * 8 + 16;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "tc_linear smem attribute: %s", cudaGetErrorString(e));
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.N + BN - 1) / BN, (a.M + TL_BM - 1) / TL_BM, 1);
    cfg.blockDim = dim3(192, 1, 1);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = a.use_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_linear_kernel<BN>, tA, tW, a);
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "tc_linear launch: %s", cudaGetErrorString(e));
    return CPM_OK;
}

}  // namespace
}  // namespace cpm

using namespace cpm;

extern "C" int cpm_tc_linear(const void *A, int64_t lda, const void *W, int64_t w_rows, const float *c1, const float *c2, void *Y, int64_t ldy,
                             int M, int N, int K, int epilogue, const float *stats_in, int parts_in, float eps, const void *R, int64_t ldr,
                             const float *stats_r, int parts_r, const float *gamma_r, const float *beta_r, const float *pe, int pe_max,
                             int pos_offset, const int *pos_dev, float *stats_out, int block_n, int use_pdl, void *stream) {
    CPM_REQUIRE(A && W && Y, CPM_ERR_NULL, "tc_linear: A/W/Y must be non-NULL");
    CPM_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 32 == 0, CPM_ERR_BAD_SHAPE, "tc_linear: M=%d N=%d K=%d (K%%64, N%%32)", M, N, K);
    CPM_REQUIRE(block_n == 32 || block_n == 64, CPM_ERR_BAD_SHAPE, "tc_linear: block_n must be 32 or 64");
    CPM_REQUIRE(lda >= K && lda % 8 == 0 && ldy >= N && ldy % 8 == 0 && w_rows >= N, CPM_ERR_BAD_SHAPE, "tc_linear: strides");
    CPM_REQUIRE(aligned16(A) && aligned16(W) && aligned16(Y), CPM_ERR_BAD_ALIGN, "tc_linear: A/W/Y must be 16-byte aligned");
    CPM_REQUIRE(epilogue >= CPM_TL_EPI_BIAS && epilogue <= CPM_TL_EPI_PE, CPM_ERR_BAD_SHAPE, "tc_linear: epilogue %d", epilogue);
    CPM_REQUIRE(!c1 || (stats_in && parts_in > 0), CPM_ERR_NULL, "tc_linear: LayerNorm fold needs the input row statistics");
    CPM_REQUIRE((epilogue != CPM_TL_EPI_RES && epilogue != CPM_TL_EPI_RES_LN) || (R && ldr % 8 == 0 && aligned16(R)), CPM_ERR_NULL,
                "tc_linear: residual epilogue needs an aligned R");
    CPM_REQUIRE(epilogue != CPM_TL_EPI_RES_LN || (stats_r && parts_r > 0 && gamma_r && beta_r), CPM_ERR_NULL, "tc_linear: RES_LN needs stats/gamma/beta");
    CPM_REQUIRE(epilogue != CPM_TL_EPI_PE || (pe && pe_max > 0), CPM_ERR_NULL, "tc_linear: PE epilogue needs pe");
    CUtensorMap tA, tW;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TL_BM))) return rc;
    if ((rc = make_tmap_bf16_2d(&tW, W, (uint64_t)K, (uint64_t)w_rows, (uint64_t)K, (uint32_t)block_n))) return rc;
    TcLinearArgs a;
    a.c1 = c1; a.c2 = c2; a.stats_in = stats_in; a.R = (const __nv_bfloat16 *)R; a.stats_r = stats_r; a.gamma_r = gamma_r; a.beta_r = beta_r;
    a.pe = pe; a.pos_dev = pos_dev; a.Y = (__nv_bfloat16 *)Y; a.stats_out = stats_out; a.ldr = ldr; a.ldy = ldy;
    a.M = M; a.N = N; a.K = K; a.epi = epilogue; a.parts_in = parts_in; a.parts_r = parts_r; a.pe_max = pe_max; a.pos_offset = pos_offset;
    a.use_pdl = use_pdl; a.eps = eps;
    cudaStream_t st = (cudaStream_t)stream;
    return block_n == 32 ? launch_tc_linear<32>(tA, tW, a, st) : launch_tc_linear<64>(tA, tW, a, st);
}
