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

constexpr int TL_BM = 128, TL_BK = 64, TL_NS_MAX = 8, TL_THREADS = 256;
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
    int split_k, ns;                        // cluster size along K (grid.z), ring stages
    uint32_t data_bytes;                    // ring / partial-tile region
    float eps;
    long long *dbg;                         // optional clock64 stamps (development aid), 8 per CTA
};

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ float ld_dsmem(uint32_t addr) {      // not volatile: the 8 x split_k loads of a unit are batched
    float v;
    asm("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void ldg8(const float *p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <int S>
__device__ __forceinline__ void sum_partials(uint32_t part_s, uint32_t off, float (&acc)[8]) {
    float t[S][8];
#pragma unroll
    for (int p = 0; p < S; ++p) {
        const uint32_t base = dsmem_addr(part_s, (uint32_t)p) + off;
#pragma unroll
        for (int j = 0; j < 8; ++j) t[p][j] = ld_dsmem(base + j * 512u);
    }
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += t[p][j];
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ void row_stats(const float *stats, int parts, int64_t m, int width, float eps, float &mean, float &rstd) {
    float s = 0.f, q = 0.f;
    const float2 *p = reinterpret_cast<const float2 *>(stats) + m * parts;
    for (int i = 0; i < parts; ++i) { const float2 v = p[i]; s += v.x; q += v.y; }
    mean = s / (float)width;
    rstd = rsqrtf(fmaxf(q / (float)width - mean * mean, 0.f) + eps);
}

// CTA (x = N-tile, y = M-tile, z = K-slice); the split_k CTAs of one (x, y) form a cluster along z.
//   1. every CTA accumulates its K-slice of the [128 x BN] tile in TMEM (TMA ring -> UMMA);
//   2. the partial tiles are parked in shared memory (column-major fp32) and the cluster synchronises;
//   3. CTA z finalises rows [z*128/split_k, (z+1)*128/split_k): sums the split_k partials over distributed shared
//      memory in rank order (deterministic), applies the epilogue and stores bf16 (+ the row statistics partial).
// Splitting K keeps the per-CTA TMA traffic small (per-SM L2->smem throughput, not FLOPs, bounds these GEMMs: at
// M = 256 every N-tile re-reads the activations) and spreads the epilogue over the cluster.
template <int BN>
__global__ void __launch_bounds__(TL_THREADS)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmR,
                 TcLinearArgs a) {
    static_assert(BN == 64, "the residual tile and the constant staging assume 64-column (128-byte) tiles");
    constexpr uint32_t W_BYTES = BN * 128, STAGE = TL_A_BYTES + W_BYTES;
    constexpr uint32_t IDESC = idesc_bf16(128, BN, false, false);
    constexpr int CG = BN / 8;                                  // 8-column groups per row
    extern __shared__ __align__(1024) uint8_t sm[];
    float *part = reinterpret_cast<float *>(sm);               // [BN][128] fp32, aliases the (dead) ring
    uint8_t *sR = sm + a.data_bytes;                              // [128/split_k rows][128 B] SW128: this CTA's residual rows
    float *srow = reinterpret_cast<float *>(sR + 16384);         // [128][4]: mean, rstd, mean_r, rstd_r
    float *upart = srow + 512;                                   // [CG][128][2]
    float *cst = upart + CG * 256;                               // [5][64]: c1, c2, gamma_r, beta_r, pe of this tile's columns
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(cst + 5 * 64), *bar_empty = bar_full + TL_NS_MAX, *bar_done = bar_empty + TL_NS_MAX;
    uint64_t *bar_res = bar_done + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_res + 1);
    const bool has_res = a.epi == CPM_TL_EPI_RES || a.epi == CPM_TL_EPI_RES_LN;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TL_BM, z = blockIdx.z, S = a.split_k, NS = a.ns;
    const int KB_all = a.K / TL_BK, per = (KB_all + S - 1) / S;
    const int kb0 = z * per, kb1 = min(KB_all, kb0 + per), KB = max(kb1 - kb0, 0);
    long long *dbg = a.dbg ? a.dbg + (int64_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
    if (dbg && tid == 0) dbg[0] = clock64();
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < TL_NS_MAX; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
        mbar_init(bar_done, 1);
        mbar_init(bar_res, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        if (has_res) tma_prefetch_desc(&tmR);
    }
    if (warp == 4) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (dbg && tid == 0) dbg[1] = clock64();

    if (warp == 4) {
        if ((tid & 31) == 0) {                       // ---- TMA producer
            const int pre = KB < NS ? KB : NS;
            for (int i = 0; i < pre; ++i) {          // weights first: they do not depend on the previous kernel
                mbar_expect_tx(bar_full + i, STAGE);
                tma_load_2d(sm + i * STAGE + TL_A_BYTES, &tmW, bar_full + i, (kb0 + i) * TL_BK, n0);
            }
            if (a.use_pdl) griddep_wait();
            for (int i = 0; i < pre; ++i) tma_load_2d(sm + i * STAGE, &tmA, bar_full + i, (kb0 + i) * TL_BK, m0);
            if (has_res) {                           // the residual rows this CTA will finalise
                mbar_expect_tx(bar_res, (uint32_t)(TL_BM / S) * 128u);
                tma_load_2d(sR, &tmR, bar_res, n0, m0 + z * (TL_BM / S));
            }
            for (int i = pre; i < KB; ++i) {
                const int s = i % NS;
                mbar_wait(bar_empty + s, ((i / NS) - 1) & 1);
                mbar_expect_tx(bar_full + s, STAGE);
                tma_load_2d(sm + s * STAGE + TL_A_BYTES, &tmW, bar_full + s, (kb0 + i) * TL_BK, n0);
                tma_load_2d(sm + s * STAGE, &tmA, bar_full + s, (kb0 + i) * TL_BK, m0);
            }
        }
    } else if (warp == 5) {
        if ((tid & 31) == 0) {                       // ---- UMMA issuer
            for (int i = 0; i < KB; ++i) {
                const int s = i % NS;
                mbar_wait(bar_full + s, (i / NS) & 1);
                tc_fence_after();
                const uint64_t dA = smem_desc_sw128(smem_u32(sm + s * STAGE)), dW = smem_desc_sw128(smem_u32(sm + s * STAGE + TL_A_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss(tmem, dA + 2 * k, dW + 2 * k, IDESC, (i > 0 || k > 0) ? 1u : 0u);
                mma_commit(bar_empty + s);            // the stage is free once these UMMAs retire
            }
            mma_commit(bar_done);
        }
    } else if (warp < 4) {
        // ---- thread = tile row: row statistics while the main loop runs, then park the accumulator row in smem
        if (a.use_pdl) griddep_wait();
        const int m = m0 + tid;
        float4 st = make_float4(0.f, 1.f, 0.f, 1.f);
        if (m < a.M) {
            if (a.c1) row_stats(a.stats_in, a.parts_in, m, a.K, a.eps, st.x, st.y);
            if (a.epi == CPM_TL_EPI_RES_LN) row_stats(a.stats_r, a.parts_r, m, a.N, a.eps, st.z, st.w);
        }
        reinterpret_cast<float4 *>(srow)[tid] = st;
        if (dbg && tid == 0) dbg[5] = clock64();
        mbar_wait(bar_done, 0);
        tc_fence_after();
        if (dbg && tid == 0) dbg[6] = clock64();
        const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int p = 0; p < BN / 32; ++p) {
            uint32_t r[32];
            if (KB > 0) {
                tmem_ld32(t_lane + 32 * p, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) part[(32 * p + i) * 128 + tid] = __uint_as_float(r[i]);
        }
        if (dbg && tid == 0) dbg[2] = clock64();
    }
    else {
        // ---- warps 6-7: stage the per-column constants of this tile
        if (a.use_pdl) griddep_wait();
        const int t = tid - 192, n = n0 + t;
        const bool ok = n < a.N;
        cst[t] = (ok && a.c1) ? a.c1[n] : 0.f;
        cst[64 + t] = (ok && a.c2) ? a.c2[n] : 0.f;
        cst[128 + t] = (ok && a.epi == CPM_TL_EPI_RES_LN) ? a.gamma_r[n] : 1.f;
        cst[192 + t] = (ok && a.epi == CPM_TL_EPI_RES_LN) ? a.beta_r[n] : 0.f;
        float pe = 0.f;
        if (ok && a.epi == CPM_TL_EPI_PE) {
            int pos = a.pos_offset + (a.pos_dev ? *a.pos_dev : 0);
            pos = pos < a.pe_max ? pos : a.pe_max - 1;
            pe = a.pe[(int64_t)pos * a.N + n];
        }
        cst[256 + t] = pe;
    }
    tc_fence_before();
    __syncthreads();
    if (dbg && tid == 0) dbg[3] = clock64();
    if (a.use_pdl && tid == 0) griddep_launch();     // the next kernel may start fetching its weights
    if (S > 1) cluster_sync_all();
    if (dbg && tid == 0) dbg[4] = clock64();

    // ---- finalise rows [z*R, (z+1)*R) of the tile: all 256 threads, 8 columns per unit; everything it reads is
    // already on chip (partials in (distributed) shared memory, residual rows, constants, row statistics)
    const int R = TL_BM / S;
    const uint32_t part_s = smem_u32(part);
    if (has_res) mbar_wait(bar_res, 0);
    for (int u = tid; u < R * CG; u += TL_THREADS) {
        const int rl = u % R, cg = u / R, r = z * R + rl, m = m0 + r, n = n0 + 8 * cg;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        if (S > 1) {                                  // all split_k x 8 loads in flight at once, summed in fixed rank order
            const uint32_t off = (uint32_t)((8 * cg) * 128 + r) * 4u;
            if (S == 2) sum_partials<2>(part_s, off, acc);
            else if (S == 4) sum_partials<4>(part_s, off, acc);
            else sum_partials<8>(part_s, off, acc);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = part[(8 * cg + j) * 128 + r];
        }
        float2 sp = make_float2(0.f, 0.f);
        if (m < a.M && n < a.N) {
            const float4 st = reinterpret_cast<const float4 *>(srow)[r];
            uint4 rres = make_uint4(0u, 0u, 0u, 0u);
            if (has_res) rres = *reinterpret_cast<const uint4 *>(sR + rl * 128 + ((cg ^ (rl & 7)) << 4));
            const uint32_t rw[4] = {rres.x, rres.y, rres.z, rres.w};
            const float *c1v = cst + 8 * cg, *c2v = cst + 64 + 8 * cg, *gv = cst + 128 + 8 * cg, *bv = cst + 192 + 8 * cg, *pv = cst + 256 + 8 * cg;
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                float v[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float x = acc[j + e];
                    x = a.c1 ? fmaf(st.y, x - st.x * c1v[j + e], c2v[j + e]) : x + c2v[j + e];
                    if (a.epi == CPM_TL_EPI_GELU) x = gelu_erf(x);
                    if (has_res) {
                        const uint32_t uu = rw[j >> 1];
                        float rr = __uint_as_float(e ? (uu & 0xffff0000u) : (uu << 16));
                        if (a.epi == CPM_TL_EPI_RES_LN) rr = fmaf((rr - st.z) * st.w, gv[j + e], bv[j + e]);
                        x += rr;
                    }
                    if (a.epi == CPM_TL_EPI_PE) x += pv[j + e];
                    v[e] = x;
                }
                const __nv_bfloat162 hb = __floats2bfloat162_rn(v[0], v[1]);
                const float2 fb = __bfloat1622float2(hb);              // statistics of what is actually stored
                sp.x += fb.x + fb.y;
                sp.y = fmaf(fb.x, fb.x, fmaf(fb.y, fb.y, sp.y));
                w[j >> 1] = *reinterpret_cast<const uint32_t *>(&hb);
            }
            *reinterpret_cast<uint4 *>(a.Y + (int64_t)m * a.ldy + n) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        reinterpret_cast<float2 *>(upart)[cg * 128 + r] = sp;
    }
    __syncthreads();
    if (a.stats_out && tid < R) {
        const int r = z * R + tid, m = m0 + r;
        if (m < a.M) {
            float2 t = make_float2(0.f, 0.f);
#pragma unroll
            for (int cg = 0; cg < CG; ++cg) { const float2 v = reinterpret_cast<const float2 *>(upart)[cg * 128 + r]; t.x += v.x; t.y += v.y; }
            reinterpret_cast<float2 *>(a.stats_out)[(int64_t)m * gridDim.x + blockIdx.x] = t;
        }
    }
    if (dbg && tid == 0) dbg[7] = clock64();
    if (S > 1) cluster_sync_all();                   // peers may still be reading this CTA's partial tile
    if (warp == 4) tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem);
}

template <int BN>
int launch_tc_linear(const CUtensorMap &tA, const CUtensorMap &tW, const CUtensorMap &tR, TcLinearArgs a, cudaStream_t st) {
    const int kb_per = (a.K / TL_BK + a.split_k - 1) / a.split_k;
    a.ns = kb_per < TL_NS_MAX ? kb_per : TL_NS_MAX;
    uint32_t data = (uint32_t)a.ns * (TL_A_BYTES + BN * 128);
    if (data < (uint32_t)BN * 512u) data = BN * 512u;
    a.data_bytes = (data + 1023u) & ~1023u;
    const uint32_t smem = a.data_bytes + 16384 + 2048 + (BN / 8) * 1024 + 5 * 64 * 4 + (2 * TL_NS_MAX + 2) * 8 + 16;
    static uint32_t attr_set = 0;
    if (smem > attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "tc_linear smem attribute: %s", cudaGetErrorString(e));
        attr_set = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.N + BN - 1) / BN, (a.M + TL_BM - 1) / TL_BM, a.split_k);
    cfg.blockDim = dim3(TL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (a.split_k > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 1; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = a.split_k;
        ++na;
    }
    if (a.use_pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_linear_kernel<BN>, tA, tW, tR, a);
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "tc_linear launch: %s", cudaGetErrorString(e));
    return CPM_OK;
}

}  // namespace
}  // namespace cpm

using namespace cpm;

static long long *g_tl_timing = nullptr;
extern "C" int cpm_debug_tc_linear_timing(void *buf) {
    g_tl_timing = reinterpret_cast<long long *>(buf);
    return CPM_OK;
}

extern "C" int cpm_tc_linear(const void *A, int64_t lda, const void *W, int64_t w_rows, const float *c1, const float *c2, void *Y, int64_t ldy,
                             int M, int N, int K, int epilogue, const float *stats_in, int parts_in, float eps, const void *R, int64_t ldr,
                             const float *stats_r, int parts_r, const float *gamma_r, const float *beta_r, const float *pe, int pe_max,
                             int pos_offset, const int *pos_dev, float *stats_out, int block_n, int split_k, int use_pdl, void *stream) {
    CPM_REQUIRE(A && W && Y, CPM_ERR_NULL, "tc_linear: A/W/Y must be non-NULL");
    CPM_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 64 == 0, CPM_ERR_BAD_SHAPE, "tc_linear: M=%d N=%d K=%d (K%%64, N%%64)", M, N, K);
    CPM_REQUIRE(block_n == 64, CPM_ERR_BAD_SHAPE, "tc_linear: block_n must be 64");
    CPM_REQUIRE(split_k == 1 || split_k == 2 || split_k == 4 || split_k == 8, CPM_ERR_BAD_SHAPE, "tc_linear: split_k must be 1, 2, 4 or 8");
    CPM_REQUIRE(lda >= K && lda % 8 == 0 && ldy >= N && ldy % 8 == 0 && w_rows >= N, CPM_ERR_BAD_SHAPE, "tc_linear: strides");
    CPM_REQUIRE(aligned16(A) && aligned16(W) && aligned16(Y), CPM_ERR_BAD_ALIGN, "tc_linear: A/W/Y must be 16-byte aligned");
    CPM_REQUIRE(epilogue >= CPM_TL_EPI_BIAS && epilogue <= CPM_TL_EPI_PE, CPM_ERR_BAD_SHAPE, "tc_linear: epilogue %d", epilogue);
    CPM_REQUIRE(!c1 || (stats_in && parts_in > 0), CPM_ERR_NULL, "tc_linear: LayerNorm fold needs the input row statistics");
    CPM_REQUIRE((epilogue != CPM_TL_EPI_RES && epilogue != CPM_TL_EPI_RES_LN) || (R && ldr % 8 == 0 && aligned16(R)), CPM_ERR_NULL,
                "tc_linear: residual epilogue needs an aligned R");
    CPM_REQUIRE(epilogue != CPM_TL_EPI_RES_LN || (stats_r && parts_r > 0 && gamma_r && beta_r), CPM_ERR_NULL, "tc_linear: RES_LN needs stats/gamma/beta");
    CPM_REQUIRE(epilogue != CPM_TL_EPI_PE || (pe && pe_max > 0), CPM_ERR_NULL, "tc_linear: PE epilogue needs pe");
    CUtensorMap tA, tW, tR;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TL_BM))) return rc;
    if ((rc = make_tmap_bf16_2d(&tW, W, (uint64_t)K, (uint64_t)w_rows, (uint64_t)K, (uint32_t)block_n))) return rc;
    const bool has_res = epilogue == CPM_TL_EPI_RES || epilogue == CPM_TL_EPI_RES_LN;
    if (has_res) {                                  // residual rows arrive by TMA: one [128/split_k x 64] box per CTA
        if ((rc = make_tmap_bf16_2d(&tR, R, (uint64_t)N, (uint64_t)M, (uint64_t)ldr, (uint32_t)(TL_BM / split_k)))) return rc;
    } else {
        tR = tA;
    }
    TcLinearArgs a;
    a.c1 = c1; a.c2 = c2; a.stats_in = stats_in; a.R = (const __nv_bfloat16 *)R; a.stats_r = stats_r; a.gamma_r = gamma_r; a.beta_r = beta_r;
    a.pe = pe; a.pos_dev = pos_dev; a.Y = (__nv_bfloat16 *)Y; a.stats_out = stats_out; a.ldr = ldr; a.ldy = ldy;
    a.M = M; a.N = N; a.K = K; a.epi = epilogue; a.parts_in = parts_in; a.parts_r = parts_r; a.pe_max = pe_max; a.pos_offset = pos_offset;
    a.use_pdl = use_pdl; a.eps = eps; a.dbg = g_tl_timing; a.split_k = split_k; a.ns = 0; a.data_bytes = 0;
    cudaStream_t st = (cudaStream_t)stream;
    return launch_tc_linear<64>(tA, tW, tR, a, st);
}
