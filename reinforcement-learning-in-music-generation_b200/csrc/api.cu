// C-ABI entry points: version / errors / linear-attention dispatch / recurrent step.
#include "cpm_common.cuh"
#include "linattn_plan.h"

namespace cpm {
thread_local char g_err[512] = "";
thread_local const char *g_linattn_impl = "none";
int g_chain_pdl = 0;

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace {
// ------------------------------------------------------------------------------------------
// B1: recurrent step. One CTA per (sequence, head); 256 threads; thread t owns row e = t/4 and
// the 16 columns m = 16*(t%4)..+16 of the 64x64 fp32 state (4 x 128-bit loads in flight, fully
// coalesced: a warp covers 8 consecutive rows = 2 KB).  out_m = sum_e Qf_e S_em / (Qf.Z + eps)
// is reduced with warp shuffles (over the 8 rows of a warp) and one shared-memory pass.
// ------------------------------------------------------------------------------------------
// streaming access to the recurrent state: it is touched exactly once per token step and is far larger than
// L2 at rollout batch sizes (256 sequences x 12 layers x 128 KB), so keep it from evicting the weights.
struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ld_stream(const float *p) {          // 256-bit load, 32-byte aligned
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

// PF: 0 = plain; 1 / 2 = additionally pull the same (sequence, head) tile of ANOTHER state tensor (the next layer's)
// into L2 with one bulk prefetch per CTA, issued after this tile's write-back (1) or before its loads (2).  The next
// layer's step runs ~25 us of latency-bound GEMM / LayerNorm kernels later and then finds its 16 KB tiles in L2.
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// stand-alone prefetch: thread i asks for bytes [16384 i, 16384 (i+1)) of the range (no SM resources beyond the launch)
__global__ void l2_prefetch_kernel(const char *p, int64_t bytes) {
    const int64_t off = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16384;
    if (off < bytes) prefetch_l2_bulk(p + off, (uint32_t)min((int64_t)16384, bytes - off));
}

template <typename T, int PF = 0>
__global__ void __launch_bounds__(256) linattn_step_kernel(const T *__restrict__ q, const T *__restrict__ k,
                                                           const T *__restrict__ v, float *__restrict__ S,
                                                           float *__restrict__ Z, T *__restrict__ out, int H,
                                                           int64_t ld_qkv, int64_t ld_o, float eps,
                                                           const float *__restrict__ S_next = nullptr) {
    __shared__ float part[8][68];                      // per warp: 64 output partials + the normaliser partial
    const int nh = blockIdx.x, n = nh / H, h = nh % H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = tid >> 2, m0 = (tid & 3) * 16;
    if (PF == 2 && tid == 0) prefetch_l2_bulk(S_next + (int64_t)nh * 4096, 16384u);
    // the 16 KB state tile first: its latency overlaps the q/k/v loads below (nothing here depends on them)
    float *srow = S + (int64_t)nh * 4096 + e * 64 + m0;
    float4 s[4];
    {
        const F8 lo = ld_stream(srow), hi = ld_stream(srow + 8);
        s[0] = lo.a; s[1] = lo.b; s[2] = hi.a; s[3] = hi.b;
    }
    // Chain kernel (cpm_common.cuh): the state tile above belongs to this layer alone (last touched a whole token step ago), so
    // its HBM latency may overlap the tail of the q/k/v projection; q, k, v are the predecessor's output.
    griddep_launch();
    griddep_wait();
    const int64_t qoff = (int64_t)n * ld_qkv + h * 64;
    const float ke = phi(to_f(k[qoff + e])), qe = phi(to_f(q[qoff + e]));
    float vv[16];
    {
        Vec8<T> v0, v1;
        v0.load(v + qoff + m0);
        v1.load(v + qoff + m0 + 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) { vv[i] = v0.v[i]; vv[8 + i] = v1.v[i]; }
    }
    float dpart = 0.f;
    if ((tid & 3) == 0) {                              // normaliser: Z += Kf ; den = Qf.Z + eps
        float *z = Z + (int64_t)nh * 64 + e;
        const float zn = *z + ke;
        *z = zn;
        dpart = qe * zn;
    }
    float acc[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s[i].x = fmaf(ke, vv[4 * i + 0], s[i].x); s[i].y = fmaf(ke, vv[4 * i + 1], s[i].y);
        s[i].z = fmaf(ke, vv[4 * i + 2], s[i].z); s[i].w = fmaf(ke, vv[4 * i + 3], s[i].w);
        acc[4 * i + 0] = qe * s[i].x; acc[4 * i + 1] = qe * s[i].y;
        acc[4 * i + 2] = qe * s[i].z; acc[4 * i + 3] = qe * s[i].w;
    }
    { F8 lo, hi; lo.a = s[0]; lo.b = s[1]; hi.a = s[2]; hi.b = s[3]; st_stream(srow, lo); st_stream(srow + 8, hi); }
    if (PF == 1 && tid == 0) prefetch_l2_bulk(S_next + (int64_t)nh * 4096, 16384u);
    // reduce over the 8 rows held by this warp (lanes with equal lane%4) by recursive halving: at every level a lane keeps half
    // of its columns and hands the other half to its partner, 8 + 4 + 2 shuffles instead of 16 x 3.  The additions pair the
    // same operands in the same tree as the plain butterfly (x + y is commutative bit for bit), so every sum is unchanged.
    float a8[8], a4[4], a2[2];
    {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float send = up ? acc[j] : acc[8 + j], keep = up ? acc[8 + j] : acc[j];
            a8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float send = up ? a8[j] : a8[4 + j], keep = up ? a8[4 + j] : a8[j];
            a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float send = up ? a4[j] : a4[2 + j], keep = up ? a4[2 + j] : a4[j];
            a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 4);
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 8);
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 16);
    {   // this lane now owns columns m0 + 8*[lane bit 2] + 4*[bit 3] + 2*[bit 4] + {0, 1}
        const int col = m0 + ((lane & 4) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 16) ? 2 : 0);
        *reinterpret_cast<float2 *>(&part[warp][col]) = make_float2(a2[0], a2[1]);
        if (lane == 0) part[warp][64] = dpart;
    }
    __syncthreads();
    if (tid < 64) {
        float o = 0.f, d = eps;
#pragma unroll
        for (int w = 0; w < 8; ++w) { o += part[w][tid]; d += part[w][64]; }
        out[(int64_t)n * ld_o + h * 64 + tid] = from_f<T>(o / d);
    }
}

// ------------------------------------------------------------------------------------------
// B1, other head widths (E rows <= 256, M in {32, 64, 128} columns; SURVEY §8 a7 lists 8 heads x 128 for cfg5).  One CTA per
// (sequence, head); M/4 threads span a state row with 128-bit accesses, the 256/(M/4) row groups stride over the rows; the
// output is reduced over row groups through shared memory.  Same arithmetic order per element as the 64-wide kernel
// (fma(Kf_e, v_m, S_em), then Qf_e * S_em), different summation order over e.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) linattn_step_wide_kernel(const T *__restrict__ q, const T *__restrict__ k,
                                                                const T *__restrict__ v, float *__restrict__ S,
                                                                float *__restrict__ Z, T *__restrict__ out, int H, int E, int M,
                                                                int64_t ld_qkv, int64_t ld_o, float eps) {
    extern __shared__ __align__(16) float sm_wide[];
    float *qf = sm_wide, *kf = qf + E, *vv = kf + E, *part = vv + M;       // part: (row groups, M)
    __shared__ float dred[8];
    const int nh = blockIdx.x, n = nh / H, h = nh % H, tid = threadIdx.x;
    float dpart = 0.f;
    for (int i = tid; i < E; i += 256) {
        const float ke = phi(to_f(k[(int64_t)n * ld_qkv + h * E + i])), qe = phi(to_f(q[(int64_t)n * ld_qkv + h * E + i]));
        float *z = Z + (int64_t)nh * E + i;
        const float zn = *z + ke;
        *z = zn;
        kf[i] = ke; qf[i] = qe;
        dpart += qe * zn;
    }
    for (int i = tid; i < M; i += 256) vv[i] = to_f(v[(int64_t)n * ld_qkv + h * M + i]);
    dpart = warp_sum(dpart);
    if ((tid & 31) == 0) dred[tid >> 5] = dpart;
    __syncthreads();
    const int tpr = M >> 2, RG = 256 / tpr, c4 = (tid % tpr) * 4, rg = tid / tpr;
    const float4 v4 = *reinterpret_cast<const float4 *>(vv + c4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float *base = S + (int64_t)nh * E * M + c4;
#pragma unroll 4
    for (int e = rg; e < E; e += RG) {
        float4 s = *reinterpret_cast<const float4 *>(base + (int64_t)e * M);
        const float ke = kf[e], qe = qf[e];
        s.x = fmaf(ke, v4.x, s.x); s.y = fmaf(ke, v4.y, s.y); s.z = fmaf(ke, v4.z, s.z); s.w = fmaf(ke, v4.w, s.w);
        *reinterpret_cast<float4 *>(base + (int64_t)e * M) = s;
        acc.x = fmaf(qe, s.x, acc.x); acc.y = fmaf(qe, s.y, acc.y); acc.z = fmaf(qe, s.z, acc.z); acc.w = fmaf(qe, s.w, acc.w);
    }
    *reinterpret_cast<float4 *>(part + rg * M + c4) = acc;
    __syncthreads();
    if (tid < M) {
        float o = 0.f, d = eps;
        for (int g = 0; g < RG; ++g) o += part[g * M + tid];
#pragma unroll
        for (int w = 0; w < 8; ++w) d += dred[w];
        out[(int64_t)n * ld_o + h * M + tid] = from_f<T>(o / d);
    }
}

// ------------------------------------------------------------------------------------------
// B1 (split): the step as two kernels so that the state write-back leaves the token step's critical path.
//   linattn_step_out_kernel     reads S, forms S + Kf (x) v in registers (same FMA as the fused kernel: bit-identical
//                               output), writes the attention output, Z, and parks [Kf | v] (512 B) for the second half;
//   linattn_state_update_kernel re-reads S and the parked [Kf | v], stores S + Kf (x) v.  The rollout engine launches it
//                               on a side branch of the step graph, where it overlaps the latency-bound GEMM / LayerNorm
//                               launches that follow; it only has to finish before the same layer's next token.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) linattn_step_out_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                                                               const float *__restrict__ S, float *__restrict__ Z, float *__restrict__ kvp,
                                                               T *__restrict__ out, int H, int64_t ld_qkv, int64_t ld_o, float eps) {
    __shared__ float part[8][68];
    const int nh = blockIdx.x, n = nh / H, h = nh % H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = tid >> 2, m0 = (tid & 3) * 16;
    const float *srow = S + (int64_t)nh * 4096 + e * 64 + m0;
    float4 s[4];
    {
        const F8 lo = ld_stream(srow), hi = ld_stream(srow + 8);
        s[0] = lo.a; s[1] = lo.b; s[2] = hi.a; s[3] = hi.b;
    }
    const int64_t qoff = (int64_t)n * ld_qkv + h * 64;
    const float ke = phi(to_f(k[qoff + e])), qe = phi(to_f(q[qoff + e]));
    float vv[16];
    {
        Vec8<T> v0, v1;
        v0.load(v + qoff + m0);
        v1.load(v + qoff + m0 + 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) { vv[i] = v0.v[i]; vv[8 + i] = v1.v[i]; }
    }
    float dpart = 0.f;
    if ((tid & 3) == 0) {
        float *z = Z + (int64_t)nh * 64 + e;
        const float zn = *z + ke;
        *z = zn;
        dpart = qe * zn;
        kvp[(int64_t)nh * 128 + e] = ke;                  // park Kf
    }
    if (e == 0) {                                          // threads 0..3 park v (16 floats each)
        float4 *dst = reinterpret_cast<float4 *>(kvp + (int64_t)nh * 128 + 64 + m0);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
    }
    float acc[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s[i].x = fmaf(ke, vv[4 * i + 0], s[i].x); s[i].y = fmaf(ke, vv[4 * i + 1], s[i].y);
        s[i].z = fmaf(ke, vv[4 * i + 2], s[i].z); s[i].w = fmaf(ke, vv[4 * i + 3], s[i].w);
        acc[4 * i + 0] = qe * s[i].x; acc[4 * i + 1] = qe * s[i].y;
        acc[4 * i + 2] = qe * s[i].z; acc[4 * i + 3] = qe * s[i].w;
    }
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
        float o = 0.f, d = eps;
#pragma unroll
        for (int w = 0; w < 8; ++w) { o += part[w][tid]; d += part[w][64]; }
        out[(int64_t)n * ld_o + h * 64 + tid] = from_f<T>(o / d);
    }
}

__global__ void __launch_bounds__(256) linattn_state_update_kernel(float *__restrict__ S, const float *__restrict__ kvp) {
    const int nh = blockIdx.x, tid = threadIdx.x;
    const int e = tid >> 2, m0 = (tid & 3) * 16;
    float *srow = S + (int64_t)nh * 4096 + e * 64 + m0;
    const F8 lo = ld_stream(srow), hi = ld_stream(srow + 8);
    const float ke = kvp[(int64_t)nh * 128 + e];
    const float4 *vp = reinterpret_cast<const float4 *>(kvp + (int64_t)nh * 128 + 64 + m0);
    const float4 v0 = vp[0], v1 = vp[1], v2 = vp[2], v3 = vp[3];
    F8 a, b;
    a.a = make_float4(fmaf(ke, v0.x, lo.a.x), fmaf(ke, v0.y, lo.a.y), fmaf(ke, v0.z, lo.a.z), fmaf(ke, v0.w, lo.a.w));
    a.b = make_float4(fmaf(ke, v1.x, lo.b.x), fmaf(ke, v1.y, lo.b.y), fmaf(ke, v1.z, lo.b.z), fmaf(ke, v1.w, lo.b.w));
    b.a = make_float4(fmaf(ke, v2.x, hi.a.x), fmaf(ke, v2.y, hi.a.y), fmaf(ke, v2.z, hi.a.z), fmaf(ke, v2.w, hi.a.w));
    b.b = make_float4(fmaf(ke, v3.x, hi.b.x), fmaf(ke, v3.y, hi.b.y), fmaf(ke, v3.z, hi.b.z), fmaf(ke, v3.w, hi.b.w));
    st_stream(srow, a);
    st_stream(srow + 8, b);
}

// ------------------------------------------------------------------------------------------
// B1 (lazy): the same step with the state write-back deferred.  The rank-1 updates of the last p = step % C tokens
// are kept in a small ring (C x [Kf | v] fp32 per (sequence, head)); every step rebuilds S_eff = S + sum_j Kf_j (x) v_j
// in registers in the original order (bit-identical to the eager kernel), and only every C-th step writes S back.
// HBM traffic per (sequence, head, step): 16 KB read + 16 KB / C written + <= C x 512 B of ring, instead of 32 KB.
// ------------------------------------------------------------------------------------------
template <typename T, int C>
__global__ void __launch_bounds__(256) linattn_step_lazy_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                                                                   float *__restrict__ S, float *__restrict__ Z, float *__restrict__ ring,
                                                                   T *__restrict__ out, const int *__restrict__ step_dev, int flush_only, int H,
                                                                   int64_t ld_qkv, int64_t ld_o, float eps) {
    __shared__ float part[8][68];
    __shared__ __align__(16) float spend[C][128];        // pending [Kf | v] entries, the newest last
    const int nh = blockIdx.x, n = nh / H, h = nh % H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = tid >> 2, m0 = (tid & 3) * 16;
    // Every load is issued up front and unconditionally (the whole ring, not just the p live entries), so that the
    // device-side step counter does not add a dependent memory round trip in front of them.
    float *srow = S + (int64_t)nh * 4096 + e * 64 + m0;
    float4 s[4];
    {
        const F8 lo = ld_stream(srow), hi = ld_stream(srow + 8);
        s[0] = lo.a; s[1] = lo.b; s[2] = hi.a; s[3] = hi.b;
    }
    float *rg = ring + (int64_t)nh * C * 128;
    float tmp[C / 2];
#pragma unroll
    for (int it = 0; it < C / 2; ++it) tmp[it] = rg[tid + 256 * it];
    const int64_t qoff = (int64_t)n * ld_qkv + h * 64;
    float newv = 0.f, qraw = 0.f, zold = 0.f;
    float *zp = Z + (int64_t)nh * 64 + e;
    if (!flush_only) {
        if (tid < 64) newv = to_f(k[qoff + tid]);
        else if (tid < 128) newv = to_f(v[qoff + tid - 64]);
        qraw = to_f(q[qoff + e]);
        if ((tid & 3) == 0) zold = *zp;
    }
    const int p = *step_dev % C;                          // pending entries already in the ring
    if (flush_only && p == 0) return;
#pragma unroll
    for (int it = 0; it < C / 2; ++it) { const int i = tid + 256 * it; if (i < p * 128) spend[i >> 7][i & 127] = tmp[it]; }
    int np = p;                                           // entries to apply
    float qe = 0.f;
    if (!flush_only) {
        if (tid < 128) {
            const float x = tid < 64 ? phi(newv) : newv;
            spend[p][tid] = x;
            if (p + 1 < C) rg[p * 128 + tid] = x;         // keep it for the following steps (not needed when flushing now)
        }
        qe = phi(qraw);
        np = p + 1;
    }
    __syncthreads();
    float dpart = 0.f;
    if (!flush_only && (tid & 3) == 0) {                  // normaliser: Z += Kf (every step); den = Qf.Z + eps
        const float zn = zold + spend[p][e];
        *zp = zn;
        dpart = qe * zn;
    }
    for (int j = 0; j < np; ++j) {                        // oldest first: the eager kernel's summation order
        const float ke = spend[j][e];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 vv = *reinterpret_cast<const float4 *>(&spend[j][64 + m0 + 4 * i]);
            s[i].x = fmaf(ke, vv.x, s[i].x); s[i].y = fmaf(ke, vv.y, s[i].y);
            s[i].z = fmaf(ke, vv.z, s[i].z); s[i].w = fmaf(ke, vv.w, s[i].w);
        }
    }
    if (flush_only || np == C) {                          // write the state back once per C steps
        F8 lo, hi; lo.a = s[0]; lo.b = s[1]; hi.a = s[2]; hi.b = s[3];
        st_stream(srow, lo); st_stream(srow + 8, hi);
    }
    if (flush_only) return;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[4 * i + 0] = qe * s[i].x; acc[4 * i + 1] = qe * s[i].y;
        acc[4 * i + 2] = qe * s[i].z; acc[4 * i + 3] = qe * s[i].w;
    }
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
        float o = 0.f, d = eps;
#pragma unroll
        for (int w = 0; w < 8; ++w) { o += part[w][tid]; d += part[w][64]; }
        out[(int64_t)n * ld_o + h * 64 + tid] = from_f<T>(o / d);
    }
}
}  // namespace
}  // namespace cpm

using namespace cpm;

extern "C" {

int cpm_version(void) { return CPM_VERSION; }
int cpm_set_chain_pdl(int on) { g_chain_pdl = on ? 1 : 0; return CPM_OK; }
const char *cpm_last_error_string(void) { return g_err; }
const char *cpm_linattn_last_impl(void) { return g_linattn_impl; }
const char *cpm_error_name(int code) {
    switch (code) {
        case CPM_OK: return "CPM_OK";
        case CPM_ERR_BAD_SHAPE: return "CPM_ERR_BAD_SHAPE";
        case CPM_ERR_BAD_ALIGN: return "CPM_ERR_BAD_ALIGN";
        case CPM_ERR_BAD_DTYPE: return "CPM_ERR_BAD_DTYPE";
        case CPM_ERR_NULL: return "CPM_ERR_NULL";
        case CPM_ERR_WORKSPACE: return "CPM_ERR_WORKSPACE";
        case CPM_ERR_CUDA: return "CPM_ERR_CUDA";
        case CPM_ERR_UNSUPPORTED: return "CPM_ERR_UNSUPPORTED";
        default: return "CPM_ERR_UNKNOWN";
    }
}

int64_t cpm_linattn_workspace_bytes(int N, int L, int H) {
    if (N <= 0 || L <= 0 || H <= 0) return 0;
    int nseg, seg_len;
    plan_segments(N, H, L, &nseg, &seg_len);
    const int64_t seg = 2ll * N * H * nseg * STATE_FLOATS * (int64_t)sizeof(float), cp = linattn_cp_workspace_bytes(N, L, H);
    return seg > cp ? seg : cp;
}
int cpm_debug_linattn_timing(void *buf) {
    linattn_cp_set_timing_buffer(reinterpret_cast<long long *>(buf));
    return CPM_OK;
}
int64_t cpm_linattn_saved_bytes(int N, int L, int H) {
    if (N <= 0 || L <= 0 || H <= 0) return 0;
    return linattn_cp_saved_bytes(N, L, H);
}

static int linattn_check(const void *a, const void *b, const void *c, const void *d, int N, int L, int H, int E, int M,
                         int64_t ld_qkv, int64_t ld_o, int dtype, void *ws, int64_t ws_bytes) {
    CPM_REQUIRE(a && b && c && d, CPM_ERR_NULL, "linattn: q/k/v/out must be non-NULL");
    CPM_REQUIRE(N > 0 && L > 0 && H > 0, CPM_ERR_BAD_SHAPE, "linattn: N=%d L=%d H=%d must be positive", N, L, H);
    CPM_REQUIRE(E == 64 && M == 64, CPM_ERR_BAD_SHAPE, "linattn: only E=M=64 is supported (got E=%d M=%d)", E, M);
    CPM_REQUIRE(dtype == CPM_F32 || dtype == CPM_BF16, CPM_ERR_BAD_DTYPE, "linattn: dtype %d", dtype);
    CPM_REQUIRE(ld_qkv >= (int64_t)H * E && ld_o >= (int64_t)H * M && ld_qkv % 8 == 0 && ld_o % 8 == 0, CPM_ERR_BAD_SHAPE,
                "linattn: token strides (%lld,%lld) must be >= H*64 and multiples of 8", (long long)ld_qkv, (long long)ld_o);
    CPM_REQUIRE(aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d), CPM_ERR_BAD_ALIGN,
                "linattn: q/k/v/out must be 16-byte aligned");
    CPM_REQUIRE(ws_bytes >= cpm_linattn_workspace_bytes(N, L, H) && (ws || ws_bytes == 0), CPM_ERR_WORKSPACE,
                "linattn: workspace %lld < required %lld", (long long)ws_bytes, (long long)cpm_linattn_workspace_bytes(N, L, H));
    return CPM_OK;
}

int cpm_linattn_fwd(const void *q, const void *k, const void *v, void *out, float *den, int N, int L, int H, int E, int M,
                    int64_t ld_qkv, int64_t ld_o, int dtype, float eps, int impl, void *workspace, int64_t workspace_bytes,
                    void *saved, int64_t saved_bytes, void *stream) {
    int rc = linattn_check(q, k, v, out, N, L, H, E, M, ld_qkv, ld_o, dtype, workspace, workspace_bytes);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    bool tc_ok = dtype == CPM_BF16 && L % 128 == 0;
    CPM_REQUIRE((impl != 2 && impl != 3) || tc_ok, CPM_ERR_UNSUPPORTED, "linattn_fwd: tcgen05 path needs bf16 and L%%128==0");
    CPM_REQUIRE(!saved || saved_bytes >= cpm_linattn_saved_bytes(N, L, H), CPM_ERR_WORKSPACE, "linattn_fwd: saved-state buffer %lld < %lld",
                (long long)saved_bytes, (long long)cpm_linattn_saved_bytes(N, L, H));
    if (impl == 3 || (impl == 0 && tc_ok)) {
        rc = linattn_fwd_cp_launch(q, k, v, out, den, N, L, H, ld_qkv, ld_o, eps, workspace, saved, st);
        // "-stream": one CTA per (batch, head) chain carries S / z in tensor memory across chunks (N*H >= 96); otherwise the
        // per-chunk state kernels + scan
        if (rc != CPM_ERR_UNSUPPORTED || impl == 3) { g_linattn_impl = (N * H >= 96 && L > 128) ? "tcgen05-cp-stream" : "tcgen05-cp"; return rc; }
    }
    if (impl == 2 || (impl == 0 && tc_ok)) {
        rc = linattn_fwd_tc_launch(q, k, v, out, den, N, L, H, ld_qkv, ld_o, eps, workspace, st);
        if (rc != CPM_ERR_UNSUPPORTED || impl == 2) { g_linattn_impl = "tcgen05"; return rc; }
    }
    g_linattn_impl = "simt";
    return linattn_fwd_simt_launch(q, k, v, out, den, N, L, H, ld_qkv, ld_o, dtype, eps, workspace, st);
}

int cpm_linattn_bwd(const void *q, const void *k, const void *v, const void *out, const float *den, const void *gout,
                    void *gq, void *gk, void *gv, int N, int L, int H, int E, int M, int64_t ld_qkv, int64_t ld_o,
                    int64_t ld_g, int dtype, float eps, int impl, void *workspace, int64_t workspace_bytes, const void *saved,
                    int64_t saved_bytes, void *stream) {
    int rc = linattn_check(q, k, v, out, N, L, H, E, M, ld_qkv, ld_o, dtype, workspace, workspace_bytes);
    if (rc) return rc;
    CPM_REQUIRE(den && gout && gq && gk && gv, CPM_ERR_NULL, "linattn_bwd: den/gout/gq/gk/gv must be non-NULL");
    CPM_REQUIRE(ld_g >= (int64_t)H * 64 && ld_g % 8 == 0, CPM_ERR_BAD_SHAPE, "linattn_bwd: ld_g=%lld", (long long)ld_g);
    CPM_REQUIRE(aligned16(gout) && aligned16(gq) && aligned16(gk) && aligned16(gv), CPM_ERR_BAD_ALIGN,
                "linattn_bwd: gradient buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    bool tc_ok = dtype == CPM_BF16 && L % 128 == 0;
    CPM_REQUIRE((impl != 2 && impl != 3) || tc_ok, CPM_ERR_UNSUPPORTED, "linattn_bwd: tcgen05 path needs bf16 and L%%128==0");
    CPM_REQUIRE(!saved || saved_bytes >= cpm_linattn_saved_bytes(N, L, H), CPM_ERR_WORKSPACE, "linattn_bwd: saved-state buffer %lld < %lld",
                (long long)saved_bytes, (long long)cpm_linattn_saved_bytes(N, L, H));
    if (impl == 3 || (impl == 0 && tc_ok)) {
        rc = linattn_bwd_cp_launch(q, k, v, out, den, gout, gq, gk, gv, N, L, H, ld_qkv, ld_o, ld_g, workspace, saved, st);
        if (rc != CPM_ERR_UNSUPPORTED || impl == 3) { g_linattn_impl = N * H >= 96 ? "tcgen05-cp-stream" : "tcgen05-cp"; return rc; }
    }
    if (impl == 2 || (impl == 0 && tc_ok)) {
        rc = linattn_bwd_tc_launch(q, k, v, out, den, gout, gq, gk, gv, N, L, H, ld_qkv, ld_o, ld_g, eps, workspace, st);
        if (rc != CPM_ERR_UNSUPPORTED || impl == 2) { g_linattn_impl = "tcgen05"; return rc; }
    }
    g_linattn_impl = "simt";
    return linattn_bwd_simt_launch(q, k, v, out, den, gout, gq, gk, gv, N, L, H, ld_qkv, ld_o, ld_g, dtype, eps, workspace, st);
}

int cpm_linattn_step(const void *q, const void *k, const void *v, float *S, float *Z, void *out, int N, int H, int E, int M,
                     int64_t ld_qkv, int64_t ld_o, int dtype, float eps, void *stream) {
    CPM_REQUIRE(q && k && v && S && Z && out, CPM_ERR_NULL, "linattn_step: NULL pointer");
    CPM_REQUIRE(N > 0 && H > 0, CPM_ERR_BAD_SHAPE, "linattn_step: N=%d H=%d", N, H);
    CPM_REQUIRE(E > 0 && E <= 256 && (M == 32 || M == 64 || M == 128), CPM_ERR_BAD_SHAPE,
                "linattn_step: E=%d M=%d (supported: E <= 256, M in {32, 64, 128}; E = M = 64 takes the streaming kernel)", E, M);
    CPM_REQUIRE(ld_qkv >= (int64_t)H * (E > M ? E : M) && ld_o >= (int64_t)H * M, CPM_ERR_BAD_SHAPE, "linattn_step: strides");
    CPM_REQUIRE(aligned16(S), CPM_ERR_BAD_ALIGN, "linattn_step: S must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (E != 64 || M != 64) {
        CPM_REQUIRE(dtype == CPM_F32 || dtype == CPM_BF16, CPM_ERR_BAD_DTYPE, "linattn_step: dtype %d", dtype);
        const size_t smem = (size_t)(2 * E + M + (256 / (M / 4)) * M) * sizeof(float);
        if (dtype == CPM_F32)
            linattn_step_wide_kernel<float><<<N * H, 256, smem, st>>>((const float *)q, (const float *)k, (const float *)v, S, Z,
                                                                      (float *)out, H, E, M, ld_qkv, ld_o, eps);
        else
            linattn_step_wide_kernel<__nv_bfloat16><<<N * H, 256, smem, st>>>((const __nv_bfloat16 *)q, (const __nv_bfloat16 *)k,
                                                                              (const __nv_bfloat16 *)v, S, Z, (__nv_bfloat16 *)out,
                                                                              H, E, M, ld_qkv, ld_o, eps);
        return check_launch("linattn_step (wide)");
    }
    if (dtype == CPM_F32)
        launch_chain(linattn_step_kernel<float, 0>, dim3(N * H), dim3(256), 0, st, (const float *)q, (const float *)k, (const float *)v, S, Z,
                     (float *)out, H, ld_qkv, ld_o, eps, (const float *)nullptr);
    else if (dtype == CPM_BF16)
        launch_chain(linattn_step_kernel<__nv_bfloat16, 0>, dim3(N * H), dim3(256), 0, st, (const __nv_bfloat16 *)q, (const __nv_bfloat16 *)k,
                     (const __nv_bfloat16 *)v, S, Z, (__nv_bfloat16 *)out, H, ld_qkv, ld_o, eps, (const float *)nullptr);
    else
        return fail(CPM_ERR_BAD_DTYPE, "linattn_step: dtype %d", dtype);
    return check_launch("linattn_step");
}

int cpm_l2_prefetch(const void *p, int64_t bytes, void *stream) {
    CPM_REQUIRE(p, CPM_ERR_NULL, "l2_prefetch: NULL pointer");
    CPM_REQUIRE(bytes > 0 && bytes % 16 == 0 && aligned16(p), CPM_ERR_BAD_ALIGN, "l2_prefetch: pointer and size must be multiples of 16 bytes");
    const int64_t pieces = (bytes + 16383) / 16384;
    l2_prefetch_kernel<<<(unsigned)((pieces + 31) / 32), 32, 0, (cudaStream_t)stream>>>((const char *)p, bytes);
    return check_launch("l2_prefetch");
}

int cpm_linattn_step_prefetch(const void *q, const void *k, const void *v, float *S, float *Z, void *out, const float *S_next, int when,
                              int N, int H, int64_t ld_qkv, int64_t ld_o, int dtype, float eps, void *stream) {
    CPM_REQUIRE(q && k && v && S && Z && out && S_next, CPM_ERR_NULL, "linattn_step_prefetch: NULL pointer");
    CPM_REQUIRE(N > 0 && H > 0, CPM_ERR_BAD_SHAPE, "linattn_step_prefetch: N=%d H=%d", N, H);
    CPM_REQUIRE(when == 1 || when == 2, CPM_ERR_BAD_SHAPE, "linattn_step_prefetch: when=%d (1: after the write-back, 2: first)", when);
    CPM_REQUIRE(ld_qkv >= (int64_t)H * 64 && ld_o >= (int64_t)H * 64, CPM_ERR_BAD_SHAPE, "linattn_step_prefetch: strides");
    CPM_REQUIRE(aligned16(S) && aligned16(S_next), CPM_ERR_BAD_ALIGN, "linattn_step_prefetch: S / S_next must be 16-byte aligned");
    CPM_REQUIRE(dtype == CPM_BF16 || dtype == CPM_F32, CPM_ERR_BAD_DTYPE, "linattn_step_prefetch: dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
#define CPM_STEP_PF(T, PF)                                                                                                         \
    linattn_step_kernel<T, PF><<<N * H, 256, 0, st>>>((const T *)q, (const T *)k, (const T *)v, S, Z, (T *)out, H, ld_qkv, ld_o, \
                                                      eps, S_next)
    if (dtype == CPM_F32) { if (when == 1) CPM_STEP_PF(float, 1); else CPM_STEP_PF(float, 2); }
    else { if (when == 1) CPM_STEP_PF(__nv_bfloat16, 1); else CPM_STEP_PF(__nv_bfloat16, 2); }
#undef CPM_STEP_PF
    return check_launch("linattn_step_prefetch");
}

int cpm_linattn_step_out(const void *q, const void *k, const void *v, const float *S, float *Z, float *kv_pending, void *out, int N, int H,
                         int64_t ld_qkv, int64_t ld_o, int dtype, float eps, void *stream) {
    CPM_REQUIRE(q && k && v && S && Z && kv_pending && out, CPM_ERR_NULL, "linattn_step_out: NULL pointer");
    CPM_REQUIRE(N > 0 && H > 0 && ld_qkv >= (int64_t)H * 64 && ld_o >= (int64_t)H * 64, CPM_ERR_BAD_SHAPE, "linattn_step_out: N=%d H=%d / strides", N, H);
    CPM_REQUIRE(aligned16(S) && aligned16(kv_pending), CPM_ERR_BAD_ALIGN, "linattn_step_out: S / kv_pending must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CPM_F32)
        linattn_step_out_kernel<float><<<N * H, 256, 0, st>>>((const float *)q, (const float *)k, (const float *)v, S, Z, kv_pending, (float *)out,
                                                              H, ld_qkv, ld_o, eps);
    else if (dtype == CPM_BF16)
        linattn_step_out_kernel<__nv_bfloat16><<<N * H, 256, 0, st>>>((const __nv_bfloat16 *)q, (const __nv_bfloat16 *)k, (const __nv_bfloat16 *)v, S,
                                                                      Z, kv_pending, (__nv_bfloat16 *)out, H, ld_qkv, ld_o, eps);
    else
        return fail(CPM_ERR_BAD_DTYPE, "linattn_step_out: dtype %d", dtype);
    return check_launch("linattn_step_out");
}

int cpm_linattn_state_update(float *S, const float *kv_pending, int N, int H, void *stream) {
    CPM_REQUIRE(S && kv_pending, CPM_ERR_NULL, "linattn_state_update: NULL pointer");
    CPM_REQUIRE(N > 0 && H > 0, CPM_ERR_BAD_SHAPE, "linattn_state_update: N=%d H=%d", N, H);
    CPM_REQUIRE(aligned16(S) && aligned16(kv_pending), CPM_ERR_BAD_ALIGN, "linattn_state_update: alignment");
    linattn_state_update_kernel<<<N * H, 256, 0, (cudaStream_t)stream>>>(S, kv_pending);
    return check_launch("linattn_state_update");
}

int cpm_linattn_step_lazy(const void *q, const void *k, const void *v, float *S, float *Z, float *ring, void *out, int N, int H,
                          int64_t ld_qkv, int64_t ld_o, int dtype, float eps, const int32_t *step_dev, int flush_only, void *stream) {
    CPM_REQUIRE(S && Z && ring && step_dev && (flush_only || (q && k && v && out)), CPM_ERR_NULL, "linattn_step_lazy: NULL pointer");
    CPM_REQUIRE(N > 0 && H > 0, CPM_ERR_BAD_SHAPE, "linattn_step_lazy: N=%d H=%d", N, H);
    CPM_REQUIRE(flush_only || (ld_qkv >= (int64_t)H * 64 && ld_o >= (int64_t)H * 64), CPM_ERR_BAD_SHAPE, "linattn_step_lazy: strides");
    CPM_REQUIRE(aligned16(S) && aligned16(ring), CPM_ERR_BAD_ALIGN, "linattn_step_lazy: S / ring must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int C = CPM_LAZY_STATE_PERIOD;
    if (dtype == CPM_F32)
        linattn_step_lazy_kernel<float, C><<<N * H, 256, 0, st>>>((const float *)q, (const float *)k, (const float *)v, S, Z, ring, (float *)out,
                                                                  step_dev, flush_only, H, ld_qkv, ld_o, eps);
    else if (dtype == CPM_BF16)
        linattn_step_lazy_kernel<__nv_bfloat16, C><<<N * H, 256, 0, st>>>((const __nv_bfloat16 *)q, (const __nv_bfloat16 *)k,
                                                                          (const __nv_bfloat16 *)v, S, Z, ring, (__nv_bfloat16 *)out, step_dev,
                                                                          flush_only, H, ld_qkv, ld_o, eps);
    else
        return fail(CPM_ERR_BAD_DTYPE, "linattn_step_lazy: dtype %d", dtype);
    return check_launch("linattn_step_lazy");
}

}  // extern "C"
