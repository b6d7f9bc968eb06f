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

constexpr int CHUNK = 128;
constexpr uint32_t TILE_BYTES = 128 * 128;       // [128 rows x 128 B] bf16 tile
// shared-memory carve-up (byte offsets from a 1024-aligned base)
constexpr uint32_t OFF_Q = 0, OFF_K = 16384, OFF_V = 32768, OFF_P = 49152, OFF_O = 81920, OFF_S = 98304, OFF_ONES = 106496;
constexpr uint32_t OFF_Z = 108544, OFF_BAR = OFF_Z + 256, OFF_TMEM = OFF_BAR + 32, FWD_SMEM_USED = OFF_TMEM + 16;
constexpr uint32_t FWD_SMEM_BYTES = FWD_SMEM_USED + 1024;     // + alignment slack
// TMEM columns (256 allocated): P / O at 0 (128 wide), S at 128 (64), Z at 192 (8)
constexpr uint32_t TM_P = 0, TM_S = 128, TM_Z = 192;
constexpr uint32_t IDESC_P = idesc_bf16(128, 128, false, false);
constexpr uint32_t IDESC_QS = idesc_bf16(128, 64, false, true);
constexpr uint32_t IDESC_PV = idesc_bf16(128, 64, false, true);
constexpr uint32_t IDESC_KV = idesc_bf16(64, 64, true, true);
constexpr uint32_t IDESC_Z = idesc_bf16(64, 8, true, true);

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

__global__ void __launch_bounds__(128, 2)
linattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, float *__restrict__ den,
                      int L, int H, int nseg, int seg_len, const float *__restrict__ ws_fwd, float eps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = sm + OFF_Q, *sK = sm + OFF_K, *sV = sm + OFF_V, *sP = sm + OFF_P, *sO = sm + OFF_O, *sS = sm + OFF_S;
    uint32_t *sOnes = reinterpret_cast<uint32_t *>(sm + OFF_ONES);
    float *sz = reinterpret_cast<float *>(sm + OFF_Z);
    uint64_t *bar_load = reinterpret_cast<uint64_t *>(sm + OFF_BAR), *bar_mma = bar_load + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int seg = blockIdx.x % nseg, nh = blockIdx.x / nseg, n = nh / H, h = nh % H;
    const int col0 = h * 64;
    const int t_begin = seg * seg_len, t_end = min(L, t_begin + seg_len);
    const int nchunks = (t_end - t_begin) / CHUNK;
    const int row_base = n * L + t_begin;

    if (tid == 0) {
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmO);
    }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    for (int i = tid; i < 512; i += 128) sOnes[i] = 0x3F803F80u;            // bf16 1.0 everywhere (layout-agnostic)
    const float *init = (nseg > 1 && seg > 0) ? ws_fwd + (int64_t)blockIdx.x * STATE_FLOATS : nullptr;
    if (tid < 64) sz[tid] = init ? init[4096 + tid] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    const int erow = 16 * warp + (lane & 15);          // state row owned by this lane (lanes 0..15 of each warp)

    bool have_state = init != nullptr;
    if (have_state) {      // seed TMEM (fp32) and sS (bf16) with the segment's initial state
        uint32_t r[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(init[erow * 64 + half * 32 + i]);
            tmem_st32(t_lane + TM_S + half * 32, r);
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 pk;
                    pk.x = pack_bf16(__uint_as_float(r[8 * c + 0]), __uint_as_float(r[8 * c + 1]));
                    pk.y = pack_bf16(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
                    pk.z = pack_bf16(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]));
                    pk.w = pack_bf16(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
                    *reinterpret_cast<uint4 *>(sS + sw128_off(erow, half * 4 + c)) = pk;
                }
            }
        }
        uint32_t z8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z8[i] = __float_as_uint(init[4096 + erow]);
        tmem_st8(t_lane + TM_Z, z8);
        tmem_st_wait();
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
        // ---- feature map in place (thread = token row), inter-chunk normaliser
        float den_inter = 0.f;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            const uint32_t off = sw128_off(tid, ch);
            float f[8];
            uint4 qv = phi8(*reinterpret_cast<const uint4 *>(sQ + off), f);
            *reinterpret_cast<uint4 *>(sQ + off) = qv;
#pragma unroll
            for (int i = 0; i < 8; ++i) den_inter = fmaf(f[i], sz[ch * 8 + i], den_inter);
            uint4 kv = phi8(*reinterpret_cast<const uint4 *>(sK + off), f);
            *reinterpret_cast<uint4 *>(sK + off) = kv;
        }
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
        // ---- P -> mask -> bf16 -> sP ; intra-chunk normaliser
        float den_intra = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t r[32];
            const bool live = (32 * p) <= (32 * warp + 31);        // warp-uniform: some j <= r in this block of columns
            if (live) {
                tmem_ld32(t_lane + TM_P + 32 * p, r);
                tmem_ld_wait();
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j0 = 32 * p + 8 * cc + 2 * i;
                    float a = (live && j0 <= tid) ? bf16_round(__uint_as_float(r[8 * cc + 2 * i])) : 0.f;
                    float b = (live && j0 + 1 <= tid) ? bf16_round(__uint_as_float(r[8 * cc + 2 * i + 1])) : 0.f;
                    den_intra += a + b;
                    w[i] = pack_bf16(a, b);
                }
                *reinterpret_cast<uint4 *>(sP + (p >> 1) * TILE_BYTES + sw128_off(tid, (p & 1) * 4 + cc)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        if (tid == 0) tma_store_wait_read0();            // the previous chunk's TMA store no longer reads sO
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
        // ---- all MMAs of this chunk are complete: refill the operand tiles for the next chunk
        if (tid == 0 && c + 1 < nchunks) {
            mbar_expect_tx(bar_load, 3 * TILE_BYTES);
            tma_load_2d(sQ, &tmQ, bar_load, col0, grow + CHUNK);
            tma_load_2d(sK, &tmK, bar_load, col0, grow + CHUNK);
            tma_load_2d(sV, &tmV, bar_load, col0, grow + CHUNK);
        }
        // ---- output rows
        const float dn = den_intra + den_inter + eps;
        const float inv = 1.f / dn;
        if (den) den[(int64_t)(grow + tid) * H + h] = dn;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld32(t_lane + TM_P + 32 * half, r);
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint4 pk;
                pk.x = pack_bf16(__uint_as_float(r[8 * cc + 0]) * inv, __uint_as_float(r[8 * cc + 1]) * inv);
                pk.y = pack_bf16(__uint_as_float(r[8 * cc + 2]) * inv, __uint_as_float(r[8 * cc + 3]) * inv);
                pk.z = pack_bf16(__uint_as_float(r[8 * cc + 4]) * inv, __uint_as_float(r[8 * cc + 5]) * inv);
                pk.w = pack_bf16(__uint_as_float(r[8 * cc + 6]) * inv, __uint_as_float(r[8 * cc + 7]) * inv);
                *reinterpret_cast<uint4 *>(sO + sw128_off(tid, half * 4 + cc)) = pk;
            }
        }
        // ---- state for the next chunk: S -> bf16 sS, Z -> fp32 sz
        if (c + 1 < nchunks) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tmem_ld32(t_lane + TM_S + 32 * half, r);
                tmem_ld_wait();
                if (lane < 16) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        uint4 pk;
                        pk.x = pack_bf16(__uint_as_float(r[8 * cc + 0]), __uint_as_float(r[8 * cc + 1]));
                        pk.y = pack_bf16(__uint_as_float(r[8 * cc + 2]), __uint_as_float(r[8 * cc + 3]));
                        pk.z = pack_bf16(__uint_as_float(r[8 * cc + 4]), __uint_as_float(r[8 * cc + 5]));
                        pk.w = pack_bf16(__uint_as_float(r[8 * cc + 6]), __uint_as_float(r[8 * cc + 7]));
                        *reinterpret_cast<uint4 *>(sS + sw128_off(erow, half * 4 + cc)) = pk;
                    }
                }
            }
            uint32_t z8[8];
            tmem_ld8(t_lane + TM_Z, z8);
            tmem_ld_wait();
            if (lane < 16) sz[erow] = __uint_as_float(z8[0]);
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
    if (warp == 0) tmem_dealloc<256>(tmem);
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
    linattn_fwd_tc_kernel<<<N * H * nseg, 128, FWD_SMEM_BYTES, st>>>(tq, tk, tv, to, den, L, H, nseg, seg_len, (const float *)ws, eps);
    return check_launch("linattn_fwd_tc");
}

int linattn_bwd_tc_launch(const void *, const void *, const void *, const void *, const float *, const void *, void *, void *, void *,
                          int, int, int, int64_t, int64_t, int64_t, float, void *, cudaStream_t) {
    return CPM_ERR_UNSUPPORTED;      // backward runs on the SIMT kernels until the tcgen05 version lands
}

}  // namespace cpm
