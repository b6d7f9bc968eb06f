// Chunked causal linear attention on the 5th-generation tensor cores (sm_100a): bf16 operands staged
// in shared memory by TMA (SWIZZLE_128B), fp32 accumulators in TMEM, tcgen05.mma issued by one
// elected thread, tcgen05.ld epilogues.  One CTA = one (batch, head, segment); chunks of 128 tokens
// are processed in order with the KV state carried in TMEM.
//
// Forward, per 128-token chunk (i, j index tokens of the chunk; e, m the 64 feature / value dims):
//   TMA      : raw q, k, v tiles [128 x 64] bf16 -> sQ, sK, sV
//   threads  : Qf = elu(q)+1, Kf = elu(k)+1 in place (thread r owns token row r);
//              den_inter_r = Qf_r . z   (z = running key sum, fp32 in smem)
//   MMA 1    : P[i][j]  = sum_e Qf[i][e] Kf[j][e]            (M128 N128 K64 , A,B K-major)      -> TMEM
//   threads  : P -> causal mask -> bf16 -> sP ; den_intra_r = rowsum
//   MMA 2    : O[i][m]  = sum_e Qf[i][e] S[e][m]             (M128 N64  K64 , B = sS MN-major)
//              O[i][m] += sum_j P[i][j]  v[j][m]             (M128 N64  K128, B = sV MN-major)
//   MMA 3    : S[e][m] += sum_j Kf[j][e] v[j][m]             (M64  N64  K128, A = sK, B = sV MN-major)
//              Z[e][*] += sum_j Kf[j][e] * 1                 (M64  N8   K128, B = all-ones tile)
//   threads  : out_r = O_r / (den_intra_r + den_inter_r + eps) -> bf16 -> sO -> TMA store;
//              S, Z -> bf16 sS / fp32 z for the next chunk.
// The normaliser and the numerator use the same bf16-rounded operands, the state never leaves the
// chip, and HBM traffic is exactly q,k,v in + out,den out (512 B per token-head + 4 B).
#include "cpm_common.cuh"
#include "linattn_plan.h"
#include "tc_common.cuh"
#include "linattn_tc_dev.cuh"

namespace cpm {

// ---------------------------------------------------------------- host: tensor maps
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}
}  // namespace

int make_tmap_bf16_2d(CUtensorMap *out, const void *base, uint64_t inner_elems, uint64_t rows, uint64_t row_stride_elems,
                      uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return fail(CPM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    // The encode call is a driver-API entry point and needs a current context on THIS thread; autograd
    // worker threads may not have touched the runtime yet, so bind the primary context once per thread.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(nullptr);
        ctx_bound = true;
    }
    cuuint64_t dims[2] = {inner_elems, rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CPM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CPM_OK;
}

namespace {
using namespace tc;
using namespace tcdev;


// =============================================================================================
// forward
// =============================================================================================
constexpr uint32_t OFF_Q = 0, OFF_K = 16384, OFF_V = 32768, OFF_P = 49152, OFF_O = 81920, OFF_S = 98304, OFF_ONES = 106496;
constexpr uint32_t OFF_Z = 108544, OFF_DP = OFF_Z + 256 /* den partials [2][2][128] */, OFF_BAR = OFF_DP + 2048, OFF_TMEM = OFF_BAR + 32;
constexpr uint32_t FWD_SMEM_BYTES = OFF_TMEM + 16;
constexpr uint32_t TM_P = 0, TM_S = 128, TM_Z = 192;
constexpr uint32_t IDESC_P = idesc_bf16(128, 128, false, false);
constexpr uint32_t IDESC_QS = idesc_bf16(128, 64, false, true);
constexpr uint32_t IDESC_PV = idesc_bf16(128, 64, false, true);
constexpr uint32_t IDESC_KV = idesc_bf16(64, 64, true, true);
constexpr uint32_t IDESC_Z = idesc_bf16(64, 8, true, true);

__global__ void __launch_bounds__(NTH, 2)
linattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, float *__restrict__ den,
                      int L, int H, int nseg, int seg_len, const float *__restrict__ ws_fwd, float eps) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sQ = sm + OFF_Q, *sK = sm + OFF_K, *sV = sm + OFF_V, *sP = sm + OFF_P, *sO = sm + OFF_O, *sS = sm + OFF_S;
    uint32_t *sOnes = reinterpret_cast<uint32_t *>(sm + OFF_ONES);
    float *sz = reinterpret_cast<float *>(sm + OFF_Z);
    float *sdp = reinterpret_cast<float *>(sm + OFF_DP);          // [0..255]: inter partials [half][row]; [256..511]: intra partials
    uint64_t *bar_load = reinterpret_cast<uint64_t *>(sm + OFF_BAR), *bar_mma = bar_load + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + OFF_TMEM);

    const int tid = threadIdx.x;
    const int seg = blockIdx.x % nseg, nh = blockIdx.x / nseg, n = nh / H, h = nh % H;
    const int col0 = h * 64;
    const int t_begin = seg * seg_len, t_end = min(L, t_begin + seg_len);
    const int nchunks = (t_end - t_begin) / CHUNK;
    const int row_base = n * L + t_begin;

    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmO);
    }
    if ((tid >> 5) == 0) tmem_alloc<256>(tmem_slot);
    for (int i = tid; i < 512; i += NTH) sOnes[i] = 0x3F803F80u;             // bf16 1.0 everywhere (layout-agnostic)
    const float *init = (nseg > 1 && seg > 0) ? ws_fwd + (int64_t)blockIdx.x * STATE_FLOATS : nullptr;
    if (tid < 64) sz[tid] = init ? init[4096 + tid] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const Geo g(tmem);

    bool have_state = init != nullptr;
    if (have_state) {
        seed_state_half(g, init, TM_S, sS);
        if (g.half == 0) {
            uint32_t z8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) z8[i] = __float_as_uint(init[4096 + g.erow]);
            tmem_st8(g.t_lane + TM_Z, z8);
            tmem_st_wait();
        }
        fence_proxy_async();
    }
    const uint64_t dQ = smem_desc_sw128(smem_u32(sQ)), dK = smem_desc_sw128(smem_u32(sK)), dV = smem_desc_sw128(smem_u32(sV));
    const uint64_t dP = smem_desc_sw128(smem_u32(sP)), dS = smem_desc_sw128(smem_u32(sS)), dOnes = smem_desc_sw128(smem_u32(sOnes));

    if (tid == 0 && nchunks > 0) {
        mbar_expect_tx(bar_load, 3 * TILE_BYTES);
        tma_load_2d(sQ, &tmQ, bar_load, col0, row_base);
        tma_load_2d(sK, &tmK, bar_load, col0, row_base);
        tma_load_2d(sV, &tmV, bar_load, col0, row_base);
    }
    tc_fence_before();
    __syncthreads();

    uint32_t ph_load = 0, ph_mma = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int grow = row_base + c * CHUNK;
        mbar_wait(bar_load, ph_load);
        ph_load ^= 1;
        // ---- feature map in place (this thread: 4 chunks of its row), inter-chunk normaliser partial
        float den_inter = 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int ch = 4 * g.half + cc;
            const uint32_t off = sw128_off(g.row, ch);
            float f[8];
            *reinterpret_cast<uint4 *>(sQ + off) = phi8(*reinterpret_cast<const uint4 *>(sQ + off), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) den_inter = fmaf(f[i], sz[ch * 8 + i], den_inter);
            *reinterpret_cast<uint4 *>(sK + off) = phi8(*reinterpret_cast<const uint4 *>(sK + off), f);
        }
        sdp[g.half * 128 + g.row] = den_inter;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 1: P = Qf Kf^T
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss(tmem + TM_P, dQ + 2 * k, dK + 2 * k, IDESC_P, k > 0);
            mma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        // ---- P -> mask -> bf16 -> sP ; intra-chunk normaliser partial
        sdp[256 + g.half * 128 + g.row] = convert_scores<true>(g, TM_P, sP, 0.f, nullptr);
        if (tid == 0) tma_store_wait_read0();             // the previous chunk's TMA store no longer reads sO
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 2 + 3
        if (tid == 0) {
            tc_fence_after();
            uint32_t acc = 0;
            if (have_state) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { mma_ss(tmem + TM_P, dQ + 2 * k, dS + 128 * k, IDESC_QS, acc); acc = 1; }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) { mma_ss(tmem + TM_P, dP + (k >> 2) * (TILE_BYTES >> 4) + 2 * (k & 3), dV + 128 * k, IDESC_PV, acc); acc = 1; }
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_ss(tmem + TM_S, dK + 128 * k, dV + 128 * k, IDESC_KV, (have_state || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_ss(tmem + TM_Z, dK + 128 * k, dOnes, IDESC_Z, (have_state || k > 0) ? 1u : 0u);
            mma_commit(bar_mma);
        }
        have_state = true;
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        if (tid == 0 && c + 1 < nchunks) {               // operand tiles are free: prefetch the next chunk
            mbar_expect_tx(bar_load, 3 * TILE_BYTES);
            tma_load_2d(sQ, &tmQ, bar_load, col0, grow + CHUNK);
            tma_load_2d(sK, &tmK, bar_load, col0, grow + CHUNK);
            tma_load_2d(sV, &tmV, bar_load, col0, grow + CHUNK);
        }
        // ---- output rows (this thread: 32 columns)
        const float dn = sdp[g.row] + sdp[128 + g.row] + sdp[256 + g.row] + sdp[384 + g.row] + eps;
        const float inv = 1.f / dn;
        if (den && g.half == 0) den[(int64_t)(grow + g.row) * H + h] = dn;
        {
            uint32_t r[32];
            tmem_ld32(g.t_lane + TM_P + 32 * g.half, r);
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4 *>(sO + sw128_off(g.row, g.half * 4 + cc)) = pack8u(r + 8 * cc, inv);
        }
        if (c + 1 < nchunks) {                            // state for the next chunk
            state_half_to_smem(g, TM_S, sS);
            if (g.half == 0) {
                uint32_t z8[8];
                tmem_ld8(g.t_lane + TM_Z, z8);
                tmem_ld_wait();
                if (g.lane < 16) sz[g.erow] = __uint_as_float(z8[0]);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tma_store_2d(&tmO, sO, col0, grow);
            tma_store_commit();
        }
    }
    if (tid == 0) tma_store_wait_all0();
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc<256>(tmem);
}

// =============================================================================================
// Backward.  Two kernels with the forward's structure (one CTA per (n, h, segment), 256 threads,
// 128-token chunks, 2 CTAs/SM, 256 TMEM columns):
//   dq pass  (chunks in forward order, carries S, z):
//     X[i][j]  = G'[i].v[j] (+ gd_i) masked j<=i           G' = go/den, gd_i = -(go_i.out_i)/den_i
//     dQf[i][e] = sum_j X[i][j] Kf[j][e] + sum_m G'[i][m] S[e][m] + gd_i z[e];  dq = dQf * phi'(q)
//   dk/dv pass (chunks in reverse order, carries R[e][m] = sum_{later i} Qf[i][e] G'[i][m], rz):
//     PT[j][i] = Kf[j].Qf[i] masked i>=j ;  dv[j][m] = sum_i PT[j][i] G'[i][m] + sum_e Kf[j][e] R[e][m]
//     WT[j][i] = v[j].G'[i] + gd_i masked ;  dKf[j][e] = sum_i WT[j][i] Qf[i][e] + sum_m v[j][m] R[e][m] + rz[e]
// The carried state tile (rows e, 64 m contiguous) is read MN-major where the contraction runs over
// e and K-major where it runs over m, so no transposed copy is ever built.
// =============================================================================================
constexpr uint32_t B_OFF_Q = 0, B_OFF_K = 16384, B_OFF_V = 32768, B_OFF_G = 49152, B_OFF_X = 65536, B_OFF_S = 98304;
constexpr uint32_t B_OFF_Z = 106496;                 // 2 x 64 floats (double-buffered z / rz)
constexpr uint32_t B_OFF_DZ = B_OFF_Z + 512;         // 4 x 64 floats partial column sums
constexpr uint32_t B_OFF_GD = B_OFF_DZ + 1024;       // 128 floats gd_i
constexpr uint32_t B_OFF_GP = B_OFF_GD + 512;        // 2 x 128 floats partial go.out dots
constexpr uint32_t B_OFF_BAR = B_OFF_GP + 1024, B_OFF_TMEM = B_OFF_BAR + 32;
constexpr uint32_t BWD_SMEM_BYTES = B_OFF_TMEM + 16;
constexpr uint32_t TB_X = 0, TB_ACC = 128, TB_ST = 192;
constexpr uint32_t IDESC_KK128 = idesc_bf16(128, 128, false, false);   // A K-major, B K-major, N=128
constexpr uint32_t IDESC_KM64 = idesc_bf16(128, 64, false, true);      // A K-major, B MN-major, N=64
constexpr uint32_t IDESC_KK64 = idesc_bf16(128, 64, false, false);     // A K-major, B K-major,  N=64
constexpr uint32_t IDESC_MM64 = idesc_bf16(64, 64, true, true);        // A MN-major, B MN-major (state update)


struct BwdArgs {
    const float *den;
    const float *ws_fwd, *ws_rev;
    int L, H, nseg, seg_len;
};

__global__ void __launch_bounds__(NTH, 2)
linattn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmGo,
                         const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmGq, BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sQ = sm + B_OFF_Q, *sK = sm + B_OFF_K, *sV = sm + B_OFF_V, *sG = sm + B_OFF_G, *sX = sm + B_OFF_X, *sS = sm + B_OFF_S;
    float *sz = reinterpret_cast<float *>(sm + B_OFF_Z), *sdz = reinterpret_cast<float *>(sm + B_OFF_DZ);
    float *sgp = reinterpret_cast<float *>(sm + B_OFF_GP);
    uint64_t *bar_load = reinterpret_cast<uint64_t *>(sm + B_OFF_BAR), *bar_mma = bar_load + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + B_OFF_TMEM);
    const int tid = threadIdx.x;
    const int seg = blockIdx.x % a.nseg, nh = blockIdx.x / a.nseg, n = nh / a.H, h = nh % a.H;
    const int col0 = h * 64;
    const int t_begin = seg * a.seg_len, t_end = min(a.L, t_begin + a.seg_len);
    const int nchunks = (t_end - t_begin) / CHUNK;
    const int row_base = n * a.L + t_begin;
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    if ((tid >> 5) == 0) tmem_alloc<256>(tmem_slot);
    const float *init = (a.nseg > 1 && seg > 0) ? a.ws_fwd + (int64_t)blockIdx.x * STATE_FLOATS : nullptr;
    if (tid < 64) sz[tid] = init ? init[4096 + tid] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const Geo g(tmem);
    bool have_state = init != nullptr;
    if (have_state) {
        seed_state_half(g, init, TB_ST, sS);
        fence_proxy_async();
    }
    const uint64_t dK = smem_desc_sw128(smem_u32(sK)), dV = smem_desc_sw128(smem_u32(sV)), dG = smem_desc_sw128(smem_u32(sG));
    const uint64_t dX = smem_desc_sw128(smem_u32(sX)), dS = smem_desc_sw128(smem_u32(sS));
    auto issue_loads = [&](int grow) {
        mbar_expect_tx(bar_load, 5 * TILE_BYTES);
        tma_load_2d(sQ, &tmQ, bar_load, col0, grow);
        tma_load_2d(sK, &tmK, bar_load, col0, grow);
        tma_load_2d(sV, &tmV, bar_load, col0, grow);
        tma_load_2d(sG, &tmGo, bar_load, col0, grow);
        tma_load_2d(sX, &tmO, bar_load, col0, grow);          // out tile parks in sX block 0 until X is written
    };
    if (tid == 0 && nchunks > 0) issue_loads(row_base);
    tc_fence_before();
    __syncthreads();
    uint32_t ph_load = 0, ph_mma = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int grow = row_base + c * CHUNK;
        const float *zc = sz + 64 * (c & 1);
        float *zn = sz + 64 * ((c + 1) & 1);
        mbar_wait(bar_load, ph_load);
        ph_load ^= 1;
        const float inv = 1.f / a.den[(int64_t)(grow + g.row) * a.H + h];
        sgp[g.half * 128 + g.row] = prep_grad_half(g, sG, sX, inv);
        uint32_t qraw[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const uint32_t off = sw128_off(g.row, 4 * g.half + cc);
            float f[8];
            *reinterpret_cast<uint4 *>(sK + off) = phi8(*reinterpret_cast<const uint4 *>(sK + off), f);
            const uint4 qv = *reinterpret_cast<const uint4 *>(sQ + off);
            qraw[4 * cc + 0] = qv.x; qraw[4 * cc + 1] = qv.y; qraw[4 * cc + 2] = qv.z; qraw[4 * cc + 3] = qv.w;
        }
        if (tid == 0) tma_store_wait_read0();      // previous chunk's dq store has finished reading sX block 1
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {                      // X = G' V^T
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_X, dG + 2 * k, dV + 2 * k, IDESC_KK128, k > 0);
            mma_commit(bar_mma);
        }
        const float gd = -inv * (sgp[g.row] + sgp[128 + g.row]);
        sdz[64 * (tid >> 6) + (tid & 63)] = colsum_quarter(sK, tid & 63, tid >> 6, nullptr);     // overlaps the MMA
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        convert_scores<true>(g, TB_X, sX, gd, nullptr);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)      // dQf = X Kf
                mma_ss(tmem + TB_ACC, dX + (k >> 2) * (TILE_BYTES >> 4) + 2 * (k & 3), dK + 128 * k, IDESC_KM64, k > 0);
            if (have_state) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_ACC, dG + 2 * k, dS + 2 * k, IDESC_KK64, 1);     // + G' S^T
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_ss(tmem + TB_ST, dK + 128 * k, dV + 128 * k, IDESC_MM64, (have_state || k > 0) ? 1u : 0u);
            mma_commit(bar_mma);
        }
        have_state = true;
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        if (tid == 0 && c + 1 < nchunks) issue_loads(grow + CHUNK);
        {   // dq rows -> staging (sX block 1), this thread's 32 columns
            uint32_t r[32];
            tmem_ld32(g.t_lane + TB_ACC + 32 * g.half, r);
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float f[8], qf[8];
                const int ch = g.half * 4 + cc;
                unpack8(make_uint4(qraw[4 * cc], qraw[4 * cc + 1], qraw[4 * cc + 2], qraw[4 * cc + 3]), qf);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = (__uint_as_float(r[8 * cc + i]) + gd * zc[8 * ch + i]) * dphi(qf[i]);
                *reinterpret_cast<uint4 *>(sX + TILE_BYTES + sw128_off(g.row, ch)) = pack8(f);
            }
        }
        if (c + 1 < nchunks) {
            state_half_to_smem(g, TB_ST, sS);
            if (tid < 64) zn[tid] = zc[tid] + sdz[tid] + sdz[64 + tid] + sdz[128 + tid] + sdz[192 + tid];
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tma_store_2d(&tmGq, sX + TILE_BYTES, col0, grow);
            tma_store_commit();
        }
    }
    if (tid == 0) tma_store_wait_all0();
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc<256>(tmem);
}

__global__ void __launch_bounds__(NTH, 2)
linattn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmGo,
                          const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmGk,
                          const __grid_constant__ CUtensorMap tmGv, BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sQ = sm + B_OFF_Q, *sK = sm + B_OFF_K, *sV = sm + B_OFF_V, *sG = sm + B_OFF_G, *sX = sm + B_OFF_X, *sR = sm + B_OFF_S;
    float *srz = reinterpret_cast<float *>(sm + B_OFF_Z), *sdr = reinterpret_cast<float *>(sm + B_OFF_DZ);
    float *sgd = reinterpret_cast<float *>(sm + B_OFF_GD), *sgp = reinterpret_cast<float *>(sm + B_OFF_GP);
    uint64_t *bar_load = reinterpret_cast<uint64_t *>(sm + B_OFF_BAR), *bar_mma = bar_load + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + B_OFF_TMEM);
    const int tid = threadIdx.x;
    const int seg = blockIdx.x % a.nseg, nh = blockIdx.x / a.nseg, n = nh / a.H, h = nh % a.H;
    const int col0 = h * 64;
    const int t_begin = seg * a.seg_len, t_end = min(a.L, t_begin + a.seg_len);
    const int nchunks = (t_end - t_begin) / CHUNK;
    const int row_base = n * a.L + t_begin;
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    if ((tid >> 5) == 0) tmem_alloc<256>(tmem_slot);
    const float *init = (a.nseg > 1 && seg < a.nseg - 1) ? a.ws_rev + (int64_t)blockIdx.x * STATE_FLOATS : nullptr;
    if (tid < 64) srz[tid] = init ? init[4096 + tid] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const Geo g(tmem);
    bool have_state = init != nullptr;
    if (have_state) {
        seed_state_half(g, init, TB_ST, sR);
        fence_proxy_async();
    }
    const uint64_t dQ = smem_desc_sw128(smem_u32(sQ)), dK = smem_desc_sw128(smem_u32(sK)), dV = smem_desc_sw128(smem_u32(sV));
    const uint64_t dG = smem_desc_sw128(smem_u32(sG)), dX = smem_desc_sw128(smem_u32(sX)), dR = smem_desc_sw128(smem_u32(sR));
    auto issue_loads = [&](int grow) {
        mbar_expect_tx(bar_load, 5 * TILE_BYTES);
        tma_load_2d(sQ, &tmQ, bar_load, col0, grow);
        tma_load_2d(sK, &tmK, bar_load, col0, grow);
        tma_load_2d(sV, &tmV, bar_load, col0, grow);
        tma_load_2d(sG, &tmGo, bar_load, col0, grow);
        tma_load_2d(sX, &tmO, bar_load, col0, grow);
    };
    if (tid == 0 && nchunks > 0) issue_loads(row_base + (nchunks - 1) * CHUNK);
    tc_fence_before();
    __syncthreads();
    uint32_t ph_load = 0, ph_mma = 0;
    for (int it = 0; it < nchunks; ++it) {
        const int c = nchunks - 1 - it;
        const int grow = row_base + c * CHUNK;
        const float *rzc = srz + 64 * (it & 1);
        float *rzn = srz + 64 * ((it + 1) & 1);
        mbar_wait(bar_load, ph_load);
        ph_load ^= 1;
        const float inv = 1.f / a.den[(int64_t)(grow + g.row) * a.H + h];
        sgp[g.half * 128 + g.row] = prep_grad_half(g, sG, sX, inv);
        uint32_t kfr[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const uint32_t off = sw128_off(g.row, 4 * g.half + cc);
            float f[8];
            *reinterpret_cast<uint4 *>(sQ + off) = phi8(*reinterpret_cast<const uint4 *>(sQ + off), f);
            const uint4 kv = phi8(*reinterpret_cast<const uint4 *>(sK + off), f);
            *reinterpret_cast<uint4 *>(sK + off) = kv;
            kfr[4 * cc + 0] = kv.x; kfr[4 * cc + 1] = kv.y; kfr[4 * cc + 2] = kv.z; kfr[4 * cc + 3] = kv.w;
        }
        __syncthreads();                     // both halves of every go.out dot are in sgp
        if (g.half == 0) sgd[g.row] = -inv * (sgp[g.row] + sgp[128 + g.row]);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {                      // X = PT = Kf Qf^T
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_X, dK + 2 * k, dQ + 2 * k, IDESC_KK128, k > 0);
            mma_commit(bar_mma);
        }
        sdr[64 * (tid >> 6) + (tid & 63)] = colsum_quarter(sQ, tid & 63, tid >> 6, sgd);          // sum_i Qf[i][e] gd_i
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        convert_scores<false>(g, TB_X, sX, 0.f, nullptr);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)      // dv = PT G'
                mma_ss(tmem + TB_ACC, dX + (k >> 2) * (TILE_BYTES >> 4) + 2 * (k & 3), dG + 128 * k, IDESC_KM64, k > 0);
            if (have_state) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_ACC, dK + 2 * k, dR + 128 * k, IDESC_KM64, 1);   // + Kf R
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_X, dV + 2 * k, dG + 2 * k, IDESC_KK128, k > 0);      // X = WT = v G'^T
            mma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        {   // dv rows -> staging in sK (Kf is no longer an operand; this thread keeps its part of the row in kfr)
            uint32_t r[32];
            tmem_ld32(g.t_lane + TB_ACC + 32 * g.half, r);
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4 *>(sK + sw128_off(g.row, g.half * 4 + cc)) = pack8u(r + 8 * cc, 1.f);
        }
        convert_scores<false>(g, TB_X, sX, 0.f, sgd);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tma_store_2d(&tmGv, sK, col0, grow);
            tma_store_commit();
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)      // dKf = WT Qf
                mma_ss(tmem + TB_ACC, dX + (k >> 2) * (TILE_BYTES >> 4) + 2 * (k & 3), dQ + 128 * k, IDESC_KM64, k > 0);
            if (have_state) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss(tmem + TB_ACC, dV + 2 * k, dR + 2 * k, IDESC_KK64, 1);      // + v R^T
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_ss(tmem + TB_ST, dQ + 128 * k, dG + 128 * k, IDESC_MM64, (have_state || k > 0) ? 1u : 0u);
            mma_commit(bar_mma);
        }
        have_state = true;
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        {   // dk rows -> staging (sX block 1)
            uint32_t r[32];
            tmem_ld32(g.t_lane + TB_ACC + 32 * g.half, r);
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float f[8], kf[8];
                const int ch = g.half * 4 + cc;
                unpack8(make_uint4(kfr[4 * cc], kfr[4 * cc + 1], kfr[4 * cc + 2], kfr[4 * cc + 3]), kf);
#pragma unroll
                for (int i = 0; i < 8; ++i)         // phi'(k) = phi(k) when phi(k) <= 1 (k <= 0), else 1
                    f[i] = (__uint_as_float(r[8 * cc + i]) + rzc[8 * ch + i]) * (kf[i] <= 1.f ? kf[i] : 1.f);
                *reinterpret_cast<uint4 *>(sX + TILE_BYTES + sw128_off(g.row, ch)) = pack8(f);
            }
        }
        if (it + 1 < nchunks) {
            state_half_to_smem(g, TB_ST, sR);
            if (tid < 64) rzn[tid] = rzc[tid] + sdr[tid] + sdr[64 + tid] + sdr[128 + tid] + sdr[192 + tid];
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tma_store_2d(&tmGk, sX + TILE_BYTES, col0, grow);
            tma_store_commit();
            if (it + 1 < nchunks) {
                tma_store_wait_read0();             // both staging tiles (sK, sX block 1) have been read out
                issue_loads(grow - CHUNK);
            }
        }
    }
    if (tid == 0) tma_store_wait_all0();
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc<256>(tmem);
}

}  // namespace

int linattn_fwd_tc_launch(const void *q, const void *k, const void *v, void *out, float *den, int N, int L, int H, int64_t ld_qkv,
                          int64_t ld_o, float eps, void *ws, cudaStream_t st) {
    if (L % CHUNK != 0) return CPM_ERR_UNSUPPORTED;
    int nseg, seg_len;
    plan_segments(N, H, L, &nseg, &seg_len);
    if (seg_len % CHUNK != 0) return CPM_ERR_UNSUPPORTED;
    CUtensorMap tq, tk, tv, to;
    int rc;
    const uint64_t rows = (uint64_t)N * L, inner = (uint64_t)H * 64;
    if ((rc = make_tmap_bf16_2d(&tq, q, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tk, k, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tv, v, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&to, out, inner, rows, ld_o, CHUNK))) return rc;
    if (nseg > 1) {
        rc = linattn_segment_states_launch(q, k, v, nullptr, nullptr, nullptr, N, L, H, ld_qkv, ld_o, CPM_BF16, ws, false, st);
        if (rc) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linattn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM_BYTES);
        if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "linattn_fwd_tc smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    linattn_fwd_tc_kernel<<<N * H * nseg, NTH, FWD_SMEM_BYTES, st>>>(tq, tk, tv, to, den, L, H, nseg, seg_len, (const float *)ws, eps);
    return check_launch("linattn_fwd_tc");
}

int linattn_bwd_tc_launch(const void *q, const void *k, const void *v, const void *out, const float *den, const void *gout,
                          void *gq, void *gk, void *gv, int N, int L, int H, int64_t ld_qkv, int64_t ld_o, int64_t ld_g, float eps,
                          void *ws, cudaStream_t st) {
    (void)eps;
    if (L % CHUNK != 0) return CPM_ERR_UNSUPPORTED;
    int nseg, seg_len;
    plan_segments(N, H, L, &nseg, &seg_len);
    if (seg_len % CHUNK != 0) return CPM_ERR_UNSUPPORTED;
    CUtensorMap tq, tk, tv, tgo, to, tgq, tgk, tgv;
    int rc;
    const uint64_t rows = (uint64_t)N * L, inner = (uint64_t)H * 64;
    if ((rc = make_tmap_bf16_2d(&tq, q, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tk, k, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tv, v, inner, rows, ld_qkv, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tgo, gout, inner, rows, ld_o, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&to, out, inner, rows, ld_o, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tgq, gq, inner, rows, ld_g, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tgk, gk, inner, rows, ld_g, CHUNK))) return rc;
    if ((rc = make_tmap_bf16_2d(&tgv, gv, inner, rows, ld_g, CHUNK))) return rc;
    if (nseg > 1) {
        rc = linattn_segment_states_launch(q, k, v, out, den, gout, N, L, H, ld_qkv, ld_o, CPM_BF16, ws, true, st);
        if (rc) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linattn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(linattn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM_BYTES);
        if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "linattn_bwd_tc smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    BwdArgs a;
    a.den = den;
    a.ws_fwd = (const float *)ws;
    a.ws_rev = a.ws_fwd + (int64_t)N * H * nseg * STATE_FLOATS;
    a.L = L; a.H = H; a.nseg = nseg; a.seg_len = seg_len;
    linattn_bwd_dq_tc_kernel<<<N * H * nseg, NTH, BWD_SMEM_BYTES, st>>>(tq, tk, tv, tgo, to, tgq, a);
    linattn_bwd_dkv_tc_kernel<<<N * H * nseg, NTH, BWD_SMEM_BYTES, st>>>(tq, tk, tv, tgo, to, tgk, tgv, a);
    return check_launch("linattn_bwd_tc");
}

}  // namespace cpm
