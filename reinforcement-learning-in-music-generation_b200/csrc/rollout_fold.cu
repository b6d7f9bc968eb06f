// Rollout token step with the LayerNorms folded away (no LayerNorm launches between the library GEMMs).
//
// ft's RecurrentTransformerEncoderLayer is post-norm:  x = norm1(x + attn(x)),  y = norm2(x + ffn(x))  (SURVEY App. A.2),
// which costs two LayerNorm launches per layer per token; at 256 songs each is ~4 us of pure latency.  Both are
// removed by (i) running the consumer GEMM on the RAW pre-LayerNorm sums with the pre-scaled weight W' = gamma (.) W
//     LN(s) . W^T + b  =  rstd * ( s . W'^T  -  mean * c1 ) + c2 ,      c1[n] = sum_k W'[n,k],  c2[n] = sum_k beta_k W[n,k] + b[n]
// and applying the per-row / per-column correction in the kernel that consumes the GEMM output anyway (the recurrent
// attention step for q,k,v; the GELU for the FFN hidden), and (ii) letting that same kernel write the LayerNorm output
// (plus the NEXT GEMM's bias) that the next GEMM takes as its accumulate-into operand (cuBLAS beta = 1 residual).
// Per layer: QKV GEMM -> step_fold -> out-proj GEMM(+=) -> FFN1 GEMM -> gelu_fold -> FFN2 GEMM(+=): 6 launches, not 8.
#include "cpm_common.cuh"

namespace cpm {
namespace {

struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ld_stream(const float *p) {          // 256-bit streaming load (see api.cu)
    F8 v;
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float *p, const F8 &v) {
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.a.x), "f"(v.a.y),
                 "f"(v.a.z), "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w)
                 : "memory");
}

// mean / rstd of one bf16 row of width d (d <= 8 * blockDim.x, d % 8 == 0) by the whole 256-thread block
__device__ __forceinline__ void block_row_stats(const __nv_bfloat16 *row, int d, float eps, float *red /* [18] */, float &mean, float &rstd) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float s = 0.f, q = 0.f;
    if (tid * 8 < d) {
        Vec8<__nv_bfloat16> v;
        v.load(row + tid * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s += v.v[i]; q = fmaf(v.v[i], v.v[i], q); }
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) { red[warp] = s; red[8 + warp] = q; }
    __syncthreads();
    if (tid == 0) {
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { ts += red[w]; tq += red[8 + w]; }
        const float m = ts / (float)d;
        red[16] = m;
        red[17] = rsqrtf(fmaxf(tq / (float)d - m * m, 0.f) + eps);
    }
    __syncthreads();
    mean = red[16];
    rstd = red[17];
}

struct StepFoldArgs {
    const __nv_bfloat16 *raw;      // (N, 3*H*64): s . W'^T without bias (q | k | v)
    const __nv_bfloat16 *s_prev;   // (N, d): pre-LayerNorm sums (fold) or the plain layer input (no fold)
    const float *c1, *c2;          // (3*H*64): fold terms (c1 unused without fold; c2 = bias then)
    const float *gamma, *beta;     // (d): the folded LayerNorm's affine parameters (fold only)
    const float *bias_next;        // (d): bias of the out-projection, added to the residual operand
    float *S, *Z;                  // recurrent state, updated in place
    __nv_bfloat16 *out;            // (N, H*64): attention output
    __nv_bfloat16 *xres;           // (N, d): LayerNorm(s_prev) (or s_prev) + bias_next  -> accumulate operand of the out-projection
    int H, d, fold;
    float eps_ln, eps_attn;
};

// One CTA per (sequence, head), 256 threads; same state-tile mapping as linattn_step_kernel (api.cu).  Every global
// load the CTA needs is issued before the first barrier (state tile, raw q/k/v, fold constants, the s_prev row), so the
// fold adds barriers but no extra round trip to the HBM-bound critical path.
__global__ void __launch_bounds__(256) linattn_step_fold_kernel(StepFoldArgs a) {
    __shared__ float part[8][68];
    __shared__ float red[18];
    __shared__ __align__(16) float sqkv[192];                       // feature-mapped q | feature-mapped k | v of this (sequence, head)
    const int nh = blockIdx.x, n = nh / a.H, h = nh % a.H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = tid >> 2, m0 = (tid & 3) * 16;
    const int HW = a.H * 64;
    float *srow = a.S + (int64_t)nh * 4096 + e * 64 + m0;
    float4 s[4];
    {   // the 16 KB state tile first: its latency overlaps everything below
        const F8 lo = ld_stream(srow), hi = ld_stream(srow + 8);
        s[0] = lo.a; s[1] = lo.b; s[2] = hi.a; s[3] = hi.b;
    }
    // thread t < 192 owns element t of [q | k | v]; thread t < 64 also owns column 64h + t of the residual operand
    float raw = 0.f, k1 = 0.f, k2 = 0.f, xin = 0.f, g = 1.f, b = 0.f, bn = 0.f;
    const __nv_bfloat16 *sp = a.s_prev + (int64_t)n * a.d;
    if (tid < 192) {
        const int col = (tid >> 6) * HW + h * 64 + (tid & 63);
        raw = __bfloat162float(a.raw[(int64_t)n * 3 * HW + col]);
        k2 = __ldg(a.c2 + col);
        if (a.fold) k1 = __ldg(a.c1 + col);
    }
    if (tid < 64) {
        const int col = h * 64 + tid;
        xin = __bfloat162float(sp[col]);
        bn = __ldg(a.bias_next + col);
        if (a.fold) { g = __ldg(a.gamma + col); b = __ldg(a.beta + col); }
    }
    float zold = 0.f;
    float *zp = a.Z + (int64_t)nh * 64 + e;
    if ((tid & 3) == 0) zold = *zp;
    float mean = 0.f, rstd = 1.f;
    if (a.fold) block_row_stats(sp, a.d, a.eps_ln, red, mean, rstd);
    if (tid < 192) {
        const float x = a.fold ? fmaf(rstd, raw - mean * k1, k2) : raw + k2;
        sqkv[tid] = tid < 128 ? phi(x) : x;
    }
    if (tid < 64) {
        const float x = a.fold ? fmaf((xin - mean) * rstd, g, b) : xin;
        a.xres[(int64_t)n * a.d + h * 64 + tid] = __float2bfloat16_rn(x + bn);
    }
    __syncthreads();
    const float qe = sqkv[e], ke = sqkv[64 + e];
    float dpart = 0.f;
    if ((tid & 3) == 0) {
        const float zn = zold + ke;
        *zp = zn;
        dpart = qe * zn;
    }
    float acc[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 vv = *reinterpret_cast<const float4 *>(sqkv + 128 + m0 + 4 * i);
        s[i].x = fmaf(ke, vv.x, s[i].x); s[i].y = fmaf(ke, vv.y, s[i].y);
        s[i].z = fmaf(ke, vv.z, s[i].z); s[i].w = fmaf(ke, vv.w, s[i].w);
        acc[4 * i + 0] = qe * s[i].x; acc[4 * i + 1] = qe * s[i].y;
        acc[4 * i + 2] = qe * s[i].z; acc[4 * i + 3] = qe * s[i].w;
    }
    { F8 lo, hi; lo.a = s[0]; lo.b = s[1]; hi.a = s[2]; hi.b = s[3]; st_stream(srow, lo); st_stream(srow + 8, hi); }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 4);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 4);
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 8);
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 16);
    if (lane < 4) {
#pragma unroll
        for (int i = 0; i < 16; ++i) part[warp][lane * 16 + i] = acc[i];
        if (lane == 0) part[warp][64] = dpart;
    }
    __syncthreads();
    if (tid < 64) {
        float o = 0.f, dd = a.eps_attn;
#pragma unroll
        for (int w = 0; w < 8; ++w) { o += part[w][tid]; dd += part[w][64]; }
        a.out[(int64_t)n * HW + h * 64 + tid] = __float2bfloat16_rn(o / dd);
    }
}

struct GeluFoldArgs {
    const __nv_bfloat16 *raw;      // (N, dff): s . W1'^T without bias
    const __nv_bfloat16 *s;        // (N, d): pre-LayerNorm sums
    const float *c1, *c2;          // (dff)
    const float *gamma, *beta;     // (d)
    const float *bias_next;        // (d): bias of linear2
    __nv_bfloat16 *h;              // (N, dff): gelu(LN(s) W1^T + b1)
    __nv_bfloat16 *xres;           // (N, d): LN(s) + bias_next  -> accumulate operand of linear2
    int d, dff;
    float eps;
};

// One CTA per row (sequence): statistics of the row, then dff / 8 output groups and d / 8 residual groups.
__global__ void __launch_bounds__(256) gelu_fold_kernel(GeluFoldArgs a) {
    __shared__ float red[18];
    const int n = blockIdx.x, tid = threadIdx.x;
    float mean, rstd;
    const __nv_bfloat16 *sp = a.s + (int64_t)n * a.d;
    block_row_stats(sp, a.d, a.eps, red, mean, rstd);
    for (int g = tid; g < a.dff / 8; g += 256) {
        Vec8<__nv_bfloat16> v, o;
        v.load(a.raw + (int64_t)n * a.dff + g * 8);
        const float4 c1a = __ldg(reinterpret_cast<const float4 *>(a.c1 + g * 8)), c1b = __ldg(reinterpret_cast<const float4 *>(a.c1 + g * 8) + 1);
        const float4 c2a = __ldg(reinterpret_cast<const float4 *>(a.c2 + g * 8)), c2b = __ldg(reinterpret_cast<const float4 *>(a.c2 + g * 8) + 1);
        const float c1v[8] = {c1a.x, c1a.y, c1a.z, c1a.w, c1b.x, c1b.y, c1b.z, c1b.w};
        const float c2v[8] = {c2a.x, c2a.y, c2a.z, c2a.w, c2b.x, c2b.y, c2b.z, c2b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float x = fmaf(rstd, v.v[i] - mean * c1v[i], c2v[i]);
            o.v[i] = 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
        }
        o.store(a.h + (int64_t)n * a.dff + g * 8);
    }
    if (tid * 8 < a.d) {
        Vec8<__nv_bfloat16> v, o;
        v.load(sp + tid * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int col = tid * 8 + i;
            o.v[i] = fmaf((v.v[i] - mean) * rstd, __ldg(a.gamma + col), __ldg(a.beta + col)) + __ldg(a.bias_next + col);
        }
        o.store(a.xres + (int64_t)n * a.d + tid * 8);
    }
}

}  // namespace
}  // namespace cpm

using namespace cpm;

extern "C" int cpm_linattn_step_fold(const void *raw_qkv, const void *s_prev, const float *c1, const float *c2, const float *gamma,
                                     const float *beta, const float *bias_next, float *S, float *Z, void *out, void *xres, int N, int H,
                                     int d, int fold, float eps_ln, float eps_attn, void *stream) {
    CPM_REQUIRE(raw_qkv && s_prev && c2 && bias_next && S && Z && out && xres, CPM_ERR_NULL, "linattn_step_fold: NULL pointer");
    CPM_REQUIRE(!fold || (c1 && gamma && beta), CPM_ERR_NULL, "linattn_step_fold: fold needs c1 / gamma / beta");
    CPM_REQUIRE(N > 0 && H > 0 && d == H * 64 && d % 8 == 0 && d <= 2048, CPM_ERR_BAD_SHAPE, "linattn_step_fold: N=%d H=%d d=%d (d = 64 H <= 2048)", N, H, d);
    CPM_REQUIRE(aligned16(raw_qkv) && aligned16(s_prev) && aligned16(S), CPM_ERR_BAD_ALIGN, "linattn_step_fold: alignment");
    StepFoldArgs a;
    a.raw = (const __nv_bfloat16 *)raw_qkv; a.s_prev = (const __nv_bfloat16 *)s_prev; a.c1 = c1; a.c2 = c2; a.gamma = gamma; a.beta = beta;
    a.bias_next = bias_next; a.S = S; a.Z = Z; a.out = (__nv_bfloat16 *)out; a.xres = (__nv_bfloat16 *)xres; a.H = H; a.d = d; a.fold = fold;
    a.eps_ln = eps_ln; a.eps_attn = eps_attn;
    linattn_step_fold_kernel<<<N * H, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("linattn_step_fold");
}

extern "C" int cpm_gelu_fold(const void *raw, const void *s, const float *c1, const float *c2, const float *gamma, const float *beta,
                             const float *bias_next, void *h, void *xres, int N, int d, int dff, float eps, void *stream) {
    CPM_REQUIRE(raw && s && c1 && c2 && gamma && beta && bias_next && h && xres, CPM_ERR_NULL, "gelu_fold: NULL pointer");
    CPM_REQUIRE(N > 0 && d > 0 && d % 8 == 0 && d <= 2048 && dff > 0 && dff % 8 == 0, CPM_ERR_BAD_SHAPE, "gelu_fold: N=%d d=%d dff=%d", N, d, dff);
    CPM_REQUIRE(aligned16(raw) && aligned16(s) && aligned16(h) && aligned16(xres) && aligned16(c1) && aligned16(c2), CPM_ERR_BAD_ALIGN,
                "gelu_fold: alignment");
    GeluFoldArgs a;
    a.raw = (const __nv_bfloat16 *)raw; a.s = (const __nv_bfloat16 *)s; a.c1 = c1; a.c2 = c2; a.gamma = gamma; a.beta = beta;
    a.bias_next = bias_next; a.h = (__nv_bfloat16 *)h; a.xres = (__nv_bfloat16 *)xres; a.d = d; a.dff = dff; a.eps = eps;
    gelu_fold_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("gelu_fold");
}
