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

template <typename T>
__global__ void __launch_bounds__(256) linattn_step_kernel(const T *__restrict__ q, const T *__restrict__ k,
                                                           const T *__restrict__ v, float *__restrict__ S,
                                                           float *__restrict__ Z, T *__restrict__ out, int H,
                                                           int64_t ld_qkv, int64_t ld_o, float eps) {
    __shared__ float part[8][68];                      // per warp: 64 output partials + the normaliser partial
    const int nh = blockIdx.x, n = nh / H, h = nh % H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = tid >> 2, m0 = (tid & 3) * 16;
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
    const int64_t seg = 2ll * N * H * nseg * STATE_FLOATS * (int64_t)sizeof(float), cp = linattn_cp_workspace_bytes(N, L, H, 64);
    return seg > cp ? seg : cp;
}
int64_t cpm_linattn_workspace_bytes_wide(int N, int L, int H, int E) {
    if (E == 64) return cpm_linattn_workspace_bytes(N, L, H);
    if (N <= 0 || L <= 0 || H <= 0 || E != 128) return 0;
    return linattn_cp_workspace_bytes(N, L, H, 128);    // tensor-core path only
}
int64_t cpm_linattn_saved_bytes_wide(int N, int L, int H, int E) {
    if (N <= 0 || L <= 0 || H <= 0 || (E != 64 && E != 128)) return 0;
    return linattn_cp_saved_bytes(N, L, H, E);
}
int cpm_debug_linattn_timing(void *buf) {
    linattn_cp_set_timing_buffer(reinterpret_cast<long long *>(buf));
    return CPM_OK;
}
int64_t cpm_linattn_saved_bytes(int N, int L, int H) {
    if (N <= 0 || L <= 0 || H <= 0) return 0;
    return linattn_cp_saved_bytes(N, L, H, 64);
}

static int linattn_check(const void *a, const void *b, const void *c, const void *d, int N, int L, int H, int E, int M,
                         int64_t ld_qkv, int64_t ld_o, int dtype, void *ws, int64_t ws_bytes) {
    CPM_REQUIRE(a && b && c && d, CPM_ERR_NULL, "linattn: q/k/v/out must be non-NULL");
    CPM_REQUIRE(N > 0 && L > 0 && H > 0, CPM_ERR_BAD_SHAPE, "linattn: N=%d L=%d H=%d must be positive", N, L, H);
    CPM_REQUIRE((E == 64 || E == 128) && M == E, CPM_ERR_BAD_SHAPE, "linattn: head widths E = M = 64 or 128 are supported (got E=%d M=%d)", E, M);
    CPM_REQUIRE(dtype == CPM_F32 || dtype == CPM_BF16, CPM_ERR_BAD_DTYPE, "linattn: dtype %d", dtype);
    CPM_REQUIRE(E == 64 || dtype == CPM_BF16, CPM_ERR_UNSUPPORTED,
                "linattn: 128-wide heads run on the tensor-core kernels only (bf16; got dtype %d)", dtype);
    CPM_REQUIRE(ld_qkv >= (int64_t)H * E && ld_o >= (int64_t)H * M && ld_qkv % 8 == 0 && ld_o % 8 == 0, CPM_ERR_BAD_SHAPE,
                "linattn: token strides (%lld,%lld) must be >= H*E and multiples of 8", (long long)ld_qkv, (long long)ld_o);
    CPM_REQUIRE(aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d), CPM_ERR_BAD_ALIGN,
                "linattn: q/k/v/out must be 16-byte aligned");
    CPM_REQUIRE(ws_bytes >= cpm_linattn_workspace_bytes_wide(N, L, H, E) && (ws || ws_bytes == 0), CPM_ERR_WORKSPACE,
                "linattn: workspace %lld < required %lld", (long long)ws_bytes, (long long)cpm_linattn_workspace_bytes_wide(N, L, H, E));
    return CPM_OK;
}

int cpm_linattn_fwd(const void *q, const void *k, const void *v, void *out, float *den, int N, int L, int H, int E, int M,
                    int64_t ld_qkv, int64_t ld_o, int dtype, float eps, int impl, void *workspace, int64_t workspace_bytes,
                    void *saved, int64_t saved_bytes, void *stream) {
    int rc = linattn_check(q, k, v, out, N, L, H, E, M, ld_qkv, ld_o, dtype, workspace, workspace_bytes);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    bool tc_ok = dtype == CPM_BF16;                     // any length: a short last chunk runs with its surplus rows masked out
    CPM_REQUIRE(impl == 0 || impl == 1 || impl == 3, CPM_ERR_BAD_SHAPE, "linattn_fwd: impl %d (0 auto | 1 simt | 3 tcgen05 chunk-parallel)", impl);
    CPM_REQUIRE(impl != 3 || tc_ok, CPM_ERR_UNSUPPORTED, "linattn_fwd: the tcgen05 path needs bf16");
    CPM_REQUIRE(E == 64 || impl != 1, CPM_ERR_UNSUPPORTED, "linattn_fwd: the CUDA-core kernels are 64-wide");
    CPM_REQUIRE(!saved || saved_bytes >= cpm_linattn_saved_bytes_wide(N, L, H, E), CPM_ERR_WORKSPACE, "linattn_fwd: saved-state buffer %lld < %lld",
                (long long)saved_bytes, (long long)cpm_linattn_saved_bytes_wide(N, L, H, E));
    if (impl == 3 || (impl == 0 && tc_ok)) {
        rc = linattn_fwd_cp_launch(q, k, v, out, den, N, L, H, E, ld_qkv, ld_o, eps, workspace, saved, st);
        // "-stream": one CTA per (batch, head) chain carries S / z in tensor memory across chunks (64-wide heads, N*H >= 96);
        // otherwise the per-chunk state kernels + scan
        if (rc != CPM_ERR_UNSUPPORTED || impl == 3 || E != 64) {
            g_linattn_impl = (linattn_cp_streams(N, H, E) && L > 128) ? "tcgen05-cp-stream" : "tcgen05-cp";
            return rc;
        }
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
    CPM_REQUIRE(ld_g >= (int64_t)H * E && ld_g % 8 == 0, CPM_ERR_BAD_SHAPE, "linattn_bwd: ld_g=%lld", (long long)ld_g);
    CPM_REQUIRE(aligned16(gout) && aligned16(gq) && aligned16(gk) && aligned16(gv), CPM_ERR_BAD_ALIGN,
                "linattn_bwd: gradient buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    bool tc_ok = dtype == CPM_BF16;
    CPM_REQUIRE(impl == 0 || impl == 1 || impl == 3, CPM_ERR_BAD_SHAPE, "linattn_bwd: impl %d (0 auto | 1 simt | 3 tcgen05 chunk-parallel)", impl);
    CPM_REQUIRE(impl != 3 || tc_ok, CPM_ERR_UNSUPPORTED, "linattn_bwd: the tcgen05 path needs bf16");
    CPM_REQUIRE(E == 64 || impl != 1, CPM_ERR_UNSUPPORTED, "linattn_bwd: the CUDA-core kernels are 64-wide");
    CPM_REQUIRE(!saved || saved_bytes >= cpm_linattn_saved_bytes_wide(N, L, H, E), CPM_ERR_WORKSPACE, "linattn_bwd: saved-state buffer %lld < %lld",
                (long long)saved_bytes, (long long)cpm_linattn_saved_bytes_wide(N, L, H, E));
    if (impl == 3 || (impl == 0 && tc_ok)) {
        rc = linattn_bwd_cp_launch(q, k, v, out, den, gout, gq, gk, gv, N, L, H, E, ld_qkv, ld_o, ld_g, workspace, saved, st);
        if (rc != CPM_ERR_UNSUPPORTED || impl == 3 || E != 64) { g_linattn_impl = linattn_cp_streams(N, H, E) ? "tcgen05-cp-stream" : "tcgen05-cp"; return rc; }
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
        launch_chain(linattn_step_kernel<float>, dim3(N * H), dim3(256), 0, st, (const float *)q, (const float *)k, (const float *)v, S, Z,
                     (float *)out, H, ld_qkv, ld_o, eps);
    else if (dtype == CPM_BF16)
        launch_chain(linattn_step_kernel<__nv_bfloat16>, dim3(N * H), dim3(256), 0, st, (const __nv_bfloat16 *)q, (const __nv_bfloat16 *)k,
                     (const __nv_bfloat16 *)v, S, Z, (__nv_bfloat16 *)out, H, ld_qkv, ld_o, eps);
    else
        return fail(CPM_ERR_BAD_DTYPE, "linattn_step: dtype %d", dtype);
    return check_launch("linattn_step");
}

}  // extern "C"
