// B1, persistent form: the recurrent step with the state tiles staged through shared memory by the bulk-copy engine.
//
// cpm_linattn_step launches one CTA per (sequence, head): 2048 CTAs at the rollout shape = 1.73 waves of 8 CTAs/SM, every CTA
// paying its own load latency before it can do anything.  Here k CTAs per SM stay resident and walk their tiles; the 16 KB
// state tile of the tile `STAGES` iterations ahead is always in flight (cp.async.bulk global -> shared, completion on an
// mbarrier), so the HBM read stream never drains between tiles and there is no second, partially filled wave.  Arithmetic,
// thread <-> element mapping and summation order are those of linattn_step_kernel (api.cu): results are bit-identical.
// The write-back goes straight from registers with 256-bit evict-first stores, as before.
#include "cpm_common.cuh"
#include "tc_common.cuh"

namespace cpm {
namespace {
using namespace tc;

constexpr int STEP_STAGES = 4;                       // 4 tiles in flight per CTA
constexpr int STEP_TILE_BYTES = 64 * 64 * 4;
constexpr int STEP_VEC_BYTES = 256;                  // slot for one 64-element vector (fp32: 256 B, bf16: 128 B used)
constexpr int STEP_STAGE_BYTES = STEP_TILE_BYTES + 4 * STEP_VEC_BYTES;      // S tile | z | q | k | v

struct F8s { float4 a, b; };
__device__ __forceinline__ void st_stream256(float *p, const F8s &v) {
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.a.x), "f"(v.a.y),
                 "f"(v.a.z), "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w)
                 : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// everything tile `nh` needs, in one transaction group: the state tile, its key sum, and the token's q / k / v head slices
template <typename T>
__device__ __forceinline__ void load_stage(unsigned char *stage, uint64_t *bar, const T *q, const T *k, const T *v, const float *S, const float *Z,
                                           int nh, int H, int64_t ld_qkv) {
    constexpr uint32_t VB = 64 * sizeof(T);
    const int64_t qoff = (int64_t)(nh / H) * ld_qkv + (nh % H) * 64;
    mbar_expect_tx(bar, STEP_TILE_BYTES + 256 + 3 * VB);
    bulk_load(stage, S + (int64_t)nh * 4096, STEP_TILE_BYTES, bar);
    bulk_load(stage + STEP_TILE_BYTES, Z + (int64_t)nh * 64, 256, bar);
    bulk_load(stage + STEP_TILE_BYTES + STEP_VEC_BYTES, q + qoff, VB, bar);
    bulk_load(stage + STEP_TILE_BYTES + 2 * STEP_VEC_BYTES, k + qoff, VB, bar);
    bulk_load(stage + STEP_TILE_BYTES + 3 * STEP_VEC_BYTES, v + qoff, VB, bar);
}

template <typename T>
__global__ void __launch_bounds__(256) linattn_step_tma_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                                                               float *__restrict__ S, float *__restrict__ Z, T *__restrict__ out, int H, int NH,
                                                               int64_t ld_qkv, int64_t ld_o, float eps) {
    extern __shared__ __align__(128) unsigned char stage_mem[];      // STEP_STAGES tiles of 64 x 64 fp32
    __shared__ __align__(8) uint64_t full[STEP_STAGES];
    __shared__ float part[2][8][68];                                  // double-buffered: one __syncthreads per tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = tid >> 2, m0 = (tid & 3) * 16;
    const int my_tiles = ((int)blockIdx.x < NH) ? (NH - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STEP_STAGES; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int s = 0; s < STEP_STAGES && s < my_tiles; ++s)
            load_stage<T>(stage_mem + s * STEP_STAGE_BYTES, &full[s], q, k, v, S, Z, blockIdx.x + s * gridDim.x, H, ld_qkv);
    }
    __syncthreads();
    for (int it = 0; it < my_tiles; ++it) {
        const int nh = blockIdx.x + it * gridDim.x, n = nh / H, h = nh % H;
        const int stage = it % STEP_STAGES;
        const uint32_t parity = (it / STEP_STAGES) & 1;
        unsigned char *stg = stage_mem + stage * STEP_STAGE_BYTES;
        mbar_wait(&full[stage], parity);
        const float *sz = reinterpret_cast<const float *>(stg + STEP_TILE_BYTES);
        const T *sq = reinterpret_cast<const T *>(stg + STEP_TILE_BYTES + STEP_VEC_BYTES);
        const T *sk = reinterpret_cast<const T *>(stg + STEP_TILE_BYTES + 2 * STEP_VEC_BYTES);
        const T *sv = reinterpret_cast<const T *>(stg + STEP_TILE_BYTES + 3 * STEP_VEC_BYTES);
        const float ke = phi(to_f(sk[e])), qe = phi(to_f(sq[e]));
        float vv[16];
        {
            Vec8<T> v0, v1;
            v0.load(sv + m0);
            v1.load(sv + m0 + 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) { vv[i] = v0.v[i]; vv[8 + i] = v1.v[i]; }
        }
        float dpart = 0.f;
        if ((tid & 3) == 0) {                              // normaliser: Z += Kf ; den = Qf.Z + eps
            const float zn = sz[e] + ke;
            Z[(int64_t)nh * 64 + e] = zn;
            dpart = qe * zn;
        }
        const float4 *srow_s = reinterpret_cast<const float4 *>(stg) + (e * 64 + m0) / 4;
        float4 s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = srow_s[i];
        float acc[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s[i].x = fmaf(ke, vv[4 * i + 0], s[i].x); s[i].y = fmaf(ke, vv[4 * i + 1], s[i].y);
            s[i].z = fmaf(ke, vv[4 * i + 2], s[i].z); s[i].w = fmaf(ke, vv[4 * i + 3], s[i].w);
            acc[4 * i + 0] = qe * s[i].x; acc[4 * i + 1] = qe * s[i].y;
            acc[4 * i + 2] = qe * s[i].z; acc[4 * i + 3] = qe * s[i].w;
        }
        {
            float *srow = S + (int64_t)nh * 4096 + e * 64 + m0;
            F8s lo, hi;
            lo.a = s[0]; lo.b = s[1]; hi.a = s[2]; hi.b = s[3];
            st_stream256(srow, lo);
            st_stream256(srow + 8, hi);
        }
        // reduce over the 8 rows held by this warp (lanes with equal lane%4), then over the warps
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 4);
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
        }
        dpart += __shfl_xor_sync(0xffffffffu, dpart, 4);
        dpart += __shfl_xor_sync(0xffffffffu, dpart, 8);
        dpart += __shfl_xor_sync(0xffffffffu, dpart, 16);
        float(*pb)[68] = part[it & 1];
        if (lane < 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pb[warp][lane * 16 + i] = acc[i];
            if (lane == 0) pb[warp][64] = dpart;
        }
        __syncthreads();                                   // partials visible; every thread is done reading this stage
        if (tid == 0 && it + STEP_STAGES < my_tiles)
            load_stage<T>(stg, &full[stage], q, k, v, S, Z, blockIdx.x + (it + STEP_STAGES) * gridDim.x, H, ld_qkv);
        if (tid < 64) {
            float o = 0.f, d = eps;
#pragma unroll
            for (int w = 0; w < 8; ++w) { o += pb[w][tid]; d += pb[w][64]; }
            out[(int64_t)n * ld_o + h * 64 + tid] = from_f<T>(o / d);
        }
    }
}

}  // namespace
}  // namespace cpm

using namespace cpm;

extern "C" int cpm_linattn_step_tma(const void *q, const void *k, const void *v, float *S, float *Z, void *out, int N, int H, int64_t ld_qkv,
                                    int64_t ld_o, int dtype, float eps, int ctas_per_sm, void *stream) {
    CPM_REQUIRE(q && k && v && S && Z && out, CPM_ERR_NULL, "linattn_step_tma: NULL pointer");
    CPM_REQUIRE(N > 0 && H > 0, CPM_ERR_BAD_SHAPE, "linattn_step_tma: N=%d H=%d", N, H);
    CPM_REQUIRE(ld_qkv >= (int64_t)H * 64 && ld_o >= (int64_t)H * 64, CPM_ERR_BAD_SHAPE, "linattn_step_tma: strides");
    const int esz = dtype == CPM_F32 ? 4 : 2;
    CPM_REQUIRE(aligned16(S) && aligned16(Z) && aligned16(q) && aligned16(k) && aligned16(v) && (ld_qkv * esz) % 16 == 0, CPM_ERR_BAD_ALIGN,
                "linattn_step_tma: S, Z, q, k, v and the q/k/v row pitch must be 16-byte aligned (bulk copies)");
    CPM_REQUIRE(ctas_per_sm >= 1 && ctas_per_sm <= 3, CPM_ERR_BAD_SHAPE, "linattn_step_tma: ctas_per_sm=%d (1..3: 68 KB of tile stages per CTA)", ctas_per_sm);
    CPM_REQUIRE(dtype == CPM_F32 || dtype == CPM_BF16, CPM_ERR_BAD_DTYPE, "linattn_step_tma: dtype %d", dtype);
    const int NH = N * H;
    const int grid = NH < num_sms() * ctas_per_sm ? NH : num_sms() * ctas_per_sm;
    const size_t smem = (size_t)STEP_STAGES * STEP_STAGE_BYTES;
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(linattn_step_tma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(linattn_step_tma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (dtype == CPM_F32)
        linattn_step_tma_kernel<float><<<grid, 256, smem, st>>>((const float *)q, (const float *)k, (const float *)v, S, Z, (float *)out, H, NH,
                                                                 ld_qkv, ld_o, eps);
    else
        linattn_step_tma_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16 *)q, (const __nv_bfloat16 *)k, (const __nv_bfloat16 *)v, S,
                                                                         Z, (__nv_bfloat16 *)out, H, NH, ld_qkv, ld_o, eps);
    return check_launch("linattn_step_tma");
}
