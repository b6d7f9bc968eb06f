// Fused "skinny" linear layer for the recurrent rollout step: Y[M x N] = epi( pro(A)[M x K] . W[N x K]^T + b )
// with M = number of sequences (<= 64).  At this size a Linear layer is a weight-streaming,
// latency-bound op, so everything around the GEMM is folded into one launch:
//   prologue : none | LayerNorm(A) (affine) computed by every CTA in shared memory (A is only M x K);
//              the normalised activations are also written out (xout) for the residual of a later layer
//   epilogue : + bias | + bias, exact-erf GELU | + bias + residual | + bias + positional encoding
// One CTA owns 8 or 16 output columns; its 8 warps split K; weights go global -> registers with
// 128-bit loads issued before anything else (they do not depend on the previous kernel), the
// activation tile is staged once in shared memory, products run on mma.sync m16n8k16 (bf16, fp32
// accumulate; K is consumed through a fixed within-64 permutation shared by both operands so that
// both are read with 128-bit accesses), partial sums are reduced through shared memory.
#include "cpm_common.cuh"

namespace cpm {
namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_MAXB = 4;          // 64-wide K blocks per warp (K <= 2048 with 8 K-slices)

struct SkinnyParams {
    const __nv_bfloat16 *A; int64_t lda;
    const __nv_bfloat16 *W;            // [N_pad x K], K contiguous
    const __nv_bfloat16 *bias;         // [N_pad] or null
    __nv_bfloat16 *Y; int64_t ldy;
    int M, N, K;
    int pro;                           // 0 none, 1 LayerNorm
    const float *gamma, *beta; float eps;
    __nv_bfloat16 *xout;               // [M x K] normalised A (pro == 1), optional
    int epi;                           // 0 bias, 1 bias+gelu, 2 bias+residual, 3 bias+pe
    const __nv_bfloat16 *R; int64_t ldr;
    const float *pe; int pe_max; int pos_offset; const int32_t *pos_dev;
    int ntc;                           // n-tiles (8 columns) per CTA: 1 or 2
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

template <int MT>      // number of 16-row tiles: M <= 16*MT
__global__ void __launch_bounds__(SK_THREADS) skinny_linear_kernel(SkinnyParams p) {
    extern __shared__ __align__(16) uint8_t sk_smem[];
    constexpr int ROWS = 16 * MT;
    const int lds = p.K + 8;                                     // padded row stride (elements)
    __nv_bfloat16 *sA = reinterpret_cast<__nv_bfloat16 *>(sk_smem);
    float *spart = reinterpret_cast<float *>(sk_smem + (size_t)ROWS * lds * 2);     // [8 warps][ROWS*8]
    float *sGamma = spart + 8 * ROWS * 8, *sBeta = sGamma + p.K;                    // LayerNorm parameters (pro == 1)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int ntc = p.ntc, ks_n = 8 / ntc;
    const int nt = warp % ntc, ks = warp / ntc;
    const int n0 = (blockIdx.x * ntc + nt) * 8;
    const int nblk = p.K >> 6;
    const int b_begin = (nblk * ks) / ks_n, b_end = (nblk * (ks + 1)) / ks_n;

    // ---- 1. weight prefetch (independent of the producer kernel)
    uint4 wreg[SK_MAXB][2];
    const bool n_ok = (n0 + g) < p.N;
#pragma unroll
    for (int b = 0; b < SK_MAXB; ++b) {
        if (b_begin + b < b_end && n_ok) {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.W + (int64_t)(n0 + g) * p.K + (b_begin + b) * 64 + 16 * q);
            wreg[b][0] = __ldg(src);
            wreg[b][1] = __ldg(src + 1);
        } else {
            wreg[b][0] = make_uint4(0, 0, 0, 0);
            wreg[b][1] = make_uint4(0, 0, 0, 0);
        }
    }

    // ---- 2. activation tile (+ LayerNorm parameters) -> shared memory, all copies in flight at once
    {
        const int vpr = p.K >> 3;                                 // 16-byte vectors per row
        const int total = p.M * vpr;
        for (int i = tid; i < total; i += SK_THREADS) {
            const int r = i / vpr, c = (i - r * vpr) * 8;
            cp_async16(sA + r * lds + c, p.A + (int64_t)r * p.lda + c);
        }
        for (int i = p.M * vpr + tid; i < ROWS * vpr; i += SK_THREADS) {      // zero the padding rows
            const int r = i / vpr, c = (i - r * vpr) * 8;
            *reinterpret_cast<uint4 *>(sA + r * lds + c) = make_uint4(0, 0, 0, 0);
        }
        if (p.pro == 1) {
            for (int i = tid; i < (p.K >> 2); i += SK_THREADS) {
                cp_async16(sGamma + 4 * i, p.gamma + 4 * i);
                cp_async16(sBeta + 4 * i, p.beta + 4 * i);
            }
        }
        cp_async_wait_all();
    }
    __syncthreads();
    if (p.pro == 1) {
        // LayerNorm in place: TPR threads per row, every row handled concurrently
        constexpr int TPR = SK_THREADS / ROWS;                    // 16, 8 or 4
        const int r = tid / TPR, j = tid % TPR;
        const int vpr = p.K >> 3;
        float sum = 0.f, sq = 0.f;
        for (int v = j; v < vpr; v += TPR) {
            Vec8<__nv_bfloat16> x;
            x.load(sA + r * lds + v * 8);
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += x.v[e];
        }
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)p.K;
        for (int v = j; v < vpr; v += TPR) {
            Vec8<__nv_bfloat16> x;
            x.load(sA + r * lds + v * 8);
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = x.v[e] - mean; sq += d * d; }
        }
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq / (float)p.K + p.eps);
        if (r < p.M) {
            for (int v = j; v < vpr; v += TPR) {
                Vec8<__nv_bfloat16> x;
                x.load(sA + r * lds + v * 8);
#pragma unroll
                for (int e = 0; e < 8; ++e) x.v[e] = (x.v[e] - mean) * rstd * sGamma[v * 8 + e] + sBeta[v * 8 + e];
                x.store(sA + r * lds + v * 8);
            }
        }
        __syncthreads();
        if (p.xout) {                    // publish this CTA's column slice of the normalised activations
            const int cw = (((p.K + (int)gridDim.x - 1) / (int)gridDim.x) + 7) & ~7;
            const int c_begin = blockIdx.x * cw, c_end = min(p.K, c_begin + cw);
            const int vs = cw >> 3;
            for (int i = tid; i < p.M * vs; i += SK_THREADS) {
                const int rr = i / vs, c = c_begin + (i % vs) * 8;
                if (c < c_end) *reinterpret_cast<uint4 *>(p.xout + (int64_t)rr * p.K + c) = *reinterpret_cast<const uint4 *>(sA + rr * lds + c);
            }
        }
    }

    // ---- 3. tensor-core products over this warp's K slice
    float acc[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][i] = 0.f;
#pragma unroll
    for (int b = 0; b < SK_MAXB; ++b) {
        if (b_begin + b < b_end) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(&wreg[b][0]);
            const int kb = (b_begin + b) * 64 + 16 * q;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const uint4 x0 = *reinterpret_cast<const uint4 *>(sA + (16 * m + g) * lds + kb);
                const uint4 x1 = *reinterpret_cast<const uint4 *>(sA + (16 * m + g) * lds + kb + 8);
                const uint4 y0 = *reinterpret_cast<const uint4 *>(sA + (16 * m + g + 8) * lds + kb);
                const uint4 y1 = *reinterpret_cast<const uint4 *>(sA + (16 * m + g + 8) * lds + kb + 8);
                const uint32_t xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const uint32_t yr[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const uint32_t a[4] = {xr[2 * s], yr[2 * s], xr[2 * s + 1], yr[2 * s + 1]};
                    mma_bf16_16816(acc[m], a, w[2 * s], w[2 * s + 1]);
                }
            }
        }
    }
    // ---- 4. partial sums -> shared memory
    float *mine = spart + warp * (ROWS * 8);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        mine[(16 * m + g) * 8 + 2 * q] = acc[m][0];
        mine[(16 * m + g) * 8 + 2 * q + 1] = acc[m][1];
        mine[(16 * m + g + 8) * 8 + 2 * q] = acc[m][2];
        mine[(16 * m + g + 8) * 8 + 2 * q + 1] = acc[m][3];
    }
    __syncthreads();
    // ---- 5. reduce over K slices + epilogue
    const int cols = ntc * 8;
    const int pos = (p.epi == 3) ? min(p.pe_max - 1, (p.pos_dev ? p.pos_dev[0] : p.pos_offset)) : 0;
    for (int o = tid; o < p.M * cols; o += SK_THREADS) {
        const int r = o / cols, cc = o % cols, t = cc >> 3, c = cc & 7;
        const int n = (blockIdx.x * ntc + t) * 8 + c;
        if (n >= p.N) continue;
        float v = 0.f;
        for (int s = 0; s < ks_n; ++s) v += spart[(s * ntc + t) * (ROWS * 8) + r * 8 + c];
        if (p.bias) v += __bfloat162float(p.bias[n]);
        if (p.epi == 1) v = gelu_exact(v);
        else if (p.epi == 2) v += __bfloat162float(p.R[(int64_t)r * p.ldr + n]);
        else if (p.epi == 3) v += p.pe[(int64_t)pos * p.N + n];
        p.Y[(int64_t)r * p.ldy + n] = __float2bfloat16_rn(v);
    }
}

}  // namespace
}  // namespace cpm

using namespace cpm;

extern "C" int cpm_skinny_linear(const void *A, int64_t lda, const void *W, const void *bias, void *Y, int64_t ldy, int M, int N, int K,
                                 int prologue, const float *gamma, const float *beta, float eps, void *xout, int epilogue,
                                 const void *residual, int64_t ldr, const float *pe, int pe_max_len, int pos_offset,
                                 const int32_t *pos_dev, void *stream) {
    CPM_REQUIRE(A && W && Y, CPM_ERR_NULL, "skinny_linear: NULL pointer");
    CPM_REQUIRE(M >= 1 && M <= 64, CPM_ERR_BAD_SHAPE, "skinny_linear: M=%d must be in [1,64] (use a dense GEMM beyond that)", M);
    CPM_REQUIRE(N >= 1 && K >= 64 && K % 64 == 0 && K <= 2048, CPM_ERR_BAD_SHAPE, "skinny_linear: N=%d K=%d (K must be a multiple of 64, <= 2048)", N, K);
    CPM_REQUIRE(prologue == 0 || (prologue == 1 && gamma && beta), CPM_ERR_NULL, "skinny_linear: LayerNorm prologue needs gamma/beta");
    CPM_REQUIRE(epilogue >= 0 && epilogue <= 3, CPM_ERR_BAD_SHAPE, "skinny_linear: epilogue %d", epilogue);
    CPM_REQUIRE(epilogue != 2 || residual, CPM_ERR_NULL, "skinny_linear: residual is NULL");
    CPM_REQUIRE(epilogue != 3 || (pe && pe_max_len > 0), CPM_ERR_NULL, "skinny_linear: pe is NULL");
    CPM_REQUIRE(aligned16(A) && aligned16(W) && lda % 8 == 0 && (!xout || aligned16(xout)) && (!gamma || (aligned16(gamma) && aligned16(beta))), CPM_ERR_BAD_ALIGN, "skinny_linear: alignment");
    SkinnyParams p{};
    p.A = (const __nv_bfloat16 *)A; p.lda = lda; p.W = (const __nv_bfloat16 *)W; p.bias = (const __nv_bfloat16 *)bias;
    p.Y = (__nv_bfloat16 *)Y; p.ldy = ldy; p.M = M; p.N = N; p.K = K; p.pro = prologue; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.xout = (__nv_bfloat16 *)xout; p.epi = epilogue; p.R = (const __nv_bfloat16 *)residual; p.ldr = ldr; p.pe = pe; p.pe_max = pe_max_len;
    p.pos_offset = pos_offset; p.pos_dev = pos_dev;
    const int ntiles = (N + 7) / 8;
    p.ntc = (K <= 1024 && ntiles >= 128) ? 2 : 1;       // 8 K-slices when K is long or the layer is narrow
    const int grid = (ntiles + p.ntc - 1) / p.ntc;
    const int MT = M <= 16 ? 1 : (M <= 32 ? 2 : 4);
    const size_t smem = (size_t)(16 * MT) * (K + 8) * 2 + (size_t)8 * (16 * MT) * 8 * sizeof(float) + (size_t)2 * K * sizeof(float);
    CPM_REQUIRE(smem <= 200 * 1024, CPM_ERR_UNSUPPORTED, "skinny_linear: M=%d K=%d needs %zu bytes of shared memory (> 200 KB)", M, K, smem);
    cudaStream_t st = (cudaStream_t)stream;
#define SK_LAUNCH(MTV)                                                                                                             \
    {                                                                                                                              \
        static size_t attr = 0;                                                                                                    \
        if (smem > 48 * 1024 && smem > attr) {                                                                                     \
            cudaError_t e = cudaFuncSetAttribute(skinny_linear_kernel<MTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
            if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "skinny_linear smem attribute: %s", cudaGetErrorString(e));            \
            attr = 200 * 1024;                                                                                                     \
        }                                                                                                                          \
        skinny_linear_kernel<MTV><<<grid, SK_THREADS, smem, st>>>(p);                                                              \
    }
    if (MT == 1) SK_LAUNCH(1) else if (MT == 2) SK_LAUNCH(2) else SK_LAUNCH(4)
#undef SK_LAUNCH
    return check_launch("skinny_linear");
}
