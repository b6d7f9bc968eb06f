// Persistent "megakernel" for one recurrent rollout token step (all sequences, all layers).
//
// At 32 sequences a token step moves ~77 MB of weights and ~50 MB of recurrent state through a
// chain of ~65 tiny dependent operations; as separate launches each costs 4-6 us of launch and
// dependent-load latency (profiles/r01_summary.md).  Here the whole step is ONE cooperative launch:
// 148 CTAs (one per SM, 256 threads) walk a host-built list of phases, separated by a software grid
// barrier; while a CTA waits at the barrier its weight fragments for the next phase are already in
// flight.  Phase kinds:
//   GEMM   Y[M,N] = epi( pro(A)[M,K] . W[N,K]^T + b )   pro: none | LayerNorm | LayerNorm∘LayerNorm |
//          CP-embedding gather;  epi: bias | +GELU | +residual | +positional encoding.
//          A CTA owns 8-column tiles; 8 warps split K; mma.sync m16n8k16 bf16, fp32 accumulate.
//   ATTN   recurrent linear-attention step per (sequence, head): Z += Kf, S += Kf v^T, out = Qf^T S / (Qf.Z+eps)
//   SAMPLE per-attribute temperature / nucleus / greedy decode with Philox, token + log-prob history
// Data produced by other CTAs inside the launch is read through L2 (cp.async.cg / ld.global.cg).
#include "cpm_common.cuh"

namespace cpm {

// ---- layouts shared with the Python side (ctypes.Structure mirrors in rollout.py) ---------------
struct MegaPhase {
    int32_t type;            // 0 gemm, 1 attn, 2 sample
    int32_t M, N, K;
    int32_t lda, ldy, ldr;
    int32_t pro;             // 0 none, 1 LN, 2 LN then LN2, 3 embedding gather
    int32_t epi;             // 0 bias, 1 bias+gelu, 2 bias+residual, 3 bias+pe
    int32_t H;               // attn: heads
    float eps;               // LayerNorm eps / attention eps
    int32_t pad0;
    const void *A, *W, *bias, *R;
    void *Y, *xout;
    const float *gamma, *beta, *gamma2, *beta2;
    float *S, *Z;            // attn state of this layer
};

struct MegaGlobals {
    int32_t batch, n_attr, emb_total, logits_ld;
    int32_t n_tokens[8], emb[8], emb_off[9], seg[9];
    float emb_scale[8], temperature[8], top_p[8];
    int32_t greedy, true_positions, max_steps, pe_max;
    uint64_t seed;
    int64_t seq_base;
    const float *tables[8];
    const float *pe;
    int64_t *cur, *hist_tok;
    float *logp, *hist_logp;
    int32_t *step_dev;
    uint32_t *barrier;
    int32_t n_phases, pad1;
};

namespace {

constexpr int MG_THREADS = 256;
constexpr int MG_ROWS = 32;              // sequences per launch (<= 32)
constexpr int MG_MAXB = 4;
constexpr int MG_MAXK = 2048;

__device__ __forceinline__ void cp_async16_cg(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ldcg_bf16(const __nv_bfloat16 *p) {
    unsigned short u;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(u) : "l"(p));
    return __bfloat162float(__ushort_as_bfloat16(u));
}

// monotonic-counter grid barrier; `phase` counts barriers passed in this launch (counter starts at 0)
__device__ __forceinline__ void grid_sync(uint32_t *bar, uint32_t &phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        const uint32_t target = (++phase) * gridDim.x;
        uint32_t spins = 0;
        while (ld_acquire(bar) < target) {
            if (++spins > (1u << 26)) {
                printf("cpmusic: rollout megakernel grid barrier timed out (block %d, phase %u)\n", (int)blockIdx.x, phase);
                __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}

struct Smem {
    __nv_bfloat16 *sA;      // [32][K+8]
    float *spart;           // [8][32*8]
    float *sGamma, *sBeta;  // [K] each (LN)
    float *misc;            // attention / sampler scratch
};

// weight fragments of one 8-column tile for this warp's K slice
struct WFrag { uint4 w[MG_MAXB][2]; };

__device__ __forceinline__ void load_wfrag(WFrag &f, const MegaPhase &ph, int tile, int warp, int g, int q) {
    const int nblk = ph.K >> 6;
    const int b_begin = (nblk * warp) >> 3, b_end = (nblk * (warp + 1)) >> 3;
    const int n = tile * 8 + g;
    const bool ok = n < ph.N;
#pragma unroll
    for (int b = 0; b < MG_MAXB; ++b) {
        if (b_begin + b < b_end && ok) {
            const uint4 *src = reinterpret_cast<const uint4 *>((const __nv_bfloat16 *)ph.W + (int64_t)n * ph.K + (b_begin + b) * 64 + 16 * q);
            f.w[b][0] = __ldg(src);
            f.w[b][1] = __ldg(src + 1);
        } else {
            f.w[b][0] = make_uint4(0, 0, 0, 0);
            f.w[b][1] = make_uint4(0, 0, 0, 0);
        }
    }
}

__device__ __forceinline__ void layernorm_inplace(const Smem &sm, int M, int K, int lds, const float *gamma, const float *beta, float eps) {
    // stage gamma/beta, then 8 threads per row, all rows concurrently
    const int tid = threadIdx.x;
    for (int i = tid; i < (K >> 2); i += MG_THREADS) {
        *reinterpret_cast<float4 *>(sm.sGamma + 4 * i) = __ldg(reinterpret_cast<const float4 *>(gamma) + i);
        *reinterpret_cast<float4 *>(sm.sBeta + 4 * i) = __ldg(reinterpret_cast<const float4 *>(beta) + i);
    }
    __syncthreads();
    constexpr int TPR = MG_THREADS / MG_ROWS;     // 8
    const int r = tid / TPR, j = tid % TPR, vpr = K >> 3;
    float sum = 0.f, sq = 0.f;
    for (int v = j; v < vpr; v += TPR) {
        Vec8<__nv_bfloat16> x;
        x.load(sm.sA + r * lds + v * 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += x.v[e];
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)K;
    for (int v = j; v < vpr; v += TPR) {
        Vec8<__nv_bfloat16> x;
        x.load(sm.sA + r * lds + v * 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = x.v[e] - mean; sq += d * d; }
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)K + eps);
    if (r < M) {
        for (int v = j; v < vpr; v += TPR) {
            Vec8<__nv_bfloat16> x;
            x.load(sm.sA + r * lds + v * 8);
#pragma unroll
            for (int e = 0; e < 8; ++e) x.v[e] = (x.v[e] - mean) * rstd * sm.sGamma[v * 8 + e] + sm.sBeta[v * 8 + e];
            x.store(sm.sA + r * lds + v * 8);
        }
    }
    __syncthreads();
}

__device__ void gemm_phase(const MegaPhase &ph, const MegaGlobals &G, const Smem &sm, WFrag &wf, bool wf_ready, int pos) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int ntiles = (ph.N + 7) >> 3;
    if ((int)blockIdx.x >= ntiles) return;
    const int lds = ph.K + 8, vpr = ph.K >> 3, M = ph.M;
    if (!wf_ready) load_wfrag(wf, ph, blockIdx.x, warp, g, q);
    // ---- stage the activation tile
    if (ph.pro == 3) {          // CP embedding gather: A[r] = concat_a table_a[cur[r,a]] * sqrt(emb_a)
        for (int i = tid; i < MG_ROWS * vpr; i += MG_THREADS) {
            const int r = i / vpr, c = (i - r * vpr) * 8;
            Vec8<__nv_bfloat16> o;
#pragma unroll
            for (int e = 0; e < 8; ++e) o.v[e] = 0.f;
            if (r < M) {
                int a = 0;
                while (c >= G.emb_off[a + 1]) ++a;
                int64_t id = G.cur[(int64_t)r * G.n_attr + a];
                id = id < 0 ? 0 : (id >= G.n_tokens[a] ? G.n_tokens[a] - 1 : id);
                const float *src = G.tables[a] + id * G.emb[a] + (c - G.emb_off[a]);
                const float4 x = __ldg(reinterpret_cast<const float4 *>(src)), y = __ldg(reinterpret_cast<const float4 *>(src) + 1);
                const float s = G.emb_scale[a];
                o.v[0] = x.x * s; o.v[1] = x.y * s; o.v[2] = x.z * s; o.v[3] = x.w * s;
                o.v[4] = y.x * s; o.v[5] = y.y * s; o.v[6] = y.z * s; o.v[7] = y.w * s;
            }
            o.store(sm.sA + r * lds + c);
        }
        __syncthreads();
    } else {
        for (int i = tid; i < M * vpr; i += MG_THREADS) {
            const int r = i / vpr, c = (i - r * vpr) * 8;
            cp_async16_cg(sm.sA + r * lds + c, (const __nv_bfloat16 *)ph.A + (int64_t)r * ph.lda + c);
        }
        for (int i = M * vpr + tid; i < MG_ROWS * vpr; i += MG_THREADS) {
            const int r = i / vpr, c = (i - r * vpr) * 8;
            *reinterpret_cast<uint4 *>(sm.sA + r * lds + c) = make_uint4(0, 0, 0, 0);
        }
        cp_async_wait();
        __syncthreads();
        if (ph.pro >= 1) layernorm_inplace(sm, M, ph.K, lds, ph.gamma, ph.beta, ph.eps);
        if (ph.pro == 2) layernorm_inplace(sm, M, ph.K, lds, ph.gamma2, ph.beta2, ph.eps);
        if (ph.pro >= 1 && ph.xout) {       // publish a column slice of the normalised activations
            const int cw = (((ph.K + ntiles - 1) / ntiles) + 7) & ~7;
            const int c_begin = blockIdx.x * cw, c_end = min(ph.K, c_begin + cw), vs = cw >> 3;
            for (int i = tid; i < M * vs; i += MG_THREADS) {
                const int rr = i / vs, c = c_begin + (i % vs) * 8;
                if (c < c_end)
                    *reinterpret_cast<uint4 *>((__nv_bfloat16 *)ph.xout + (int64_t)rr * ph.K + c) = *reinterpret_cast<const uint4 *>(sm.sA + rr * lds + c);
            }
        }
    }
    // ---- tiles of this CTA
    const int nblk = ph.K >> 6;
    const int b_begin = (nblk * warp) >> 3, b_end = (nblk * (warp + 1)) >> 3;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (tile != (int)blockIdx.x) load_wfrag(wf, ph, tile, warp, g, q);
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int b = 0; b < MG_MAXB; ++b) {
            if (b_begin + b < b_end) {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(&wf.w[b][0]);
                const int kb = (b_begin + b) * 64 + 16 * q;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const uint4 x0 = *reinterpret_cast<const uint4 *>(sm.sA + (16 * m + g) * lds + kb);
                    const uint4 x1 = *reinterpret_cast<const uint4 *>(sm.sA + (16 * m + g) * lds + kb + 8);
                    const uint4 y0 = *reinterpret_cast<const uint4 *>(sm.sA + (16 * m + g + 8) * lds + kb);
                    const uint4 y1 = *reinterpret_cast<const uint4 *>(sm.sA + (16 * m + g + 8) * lds + kb + 8);
                    const uint32_t xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                    const uint32_t yr[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const uint32_t a[4] = {xr[2 * s], yr[2 * s], xr[2 * s + 1], yr[2 * s + 1]};
                        mma16816(acc[m], a, w[2 * s], w[2 * s + 1]);
                    }
                }
            }
        }
        float *mine = sm.spart + warp * (MG_ROWS * 8);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            mine[(16 * m + g) * 8 + 2 * q] = acc[m][0];
            mine[(16 * m + g) * 8 + 2 * q + 1] = acc[m][1];
            mine[(16 * m + g + 8) * 8 + 2 * q] = acc[m][2];
            mine[(16 * m + g + 8) * 8 + 2 * q + 1] = acc[m][3];
        }
        __syncthreads();
        {   // 256 threads <-> 32 rows x 8 columns
            const int r = tid >> 3, c = tid & 7, n = tile * 8 + c;
            if (r < M && n < ph.N) {
                float v = 0.f;
#pragma unroll
                for (int s = 0; s < 8; ++s) v += sm.spart[s * (MG_ROWS * 8) + r * 8 + c];
                if (ph.bias) v += __bfloat162float(__ldg((const __nv_bfloat16 *)ph.bias + n));
                if (ph.epi == 1) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
                else if (ph.epi == 2) v += ldcg_bf16((const __nv_bfloat16 *)ph.R + (int64_t)r * ph.ldr + n);
                else if (ph.epi == 3) v += __ldg(G.pe + (int64_t)pos * ph.N + n);
                ((__nv_bfloat16 *)ph.Y)[(int64_t)r * ph.ldy + n] = __float2bfloat16_rn(v);
            }
        }
        __syncthreads();
    }
}

// recurrent linear-attention step for the (sequence, head) pairs of this CTA
__device__ void attn_phase(const MegaPhase &ph, const Smem &sm) {
    float *sq = sm.misc, *sk = sq + 64, *sv = sk + 64, *part = sv + 64, *sden = part + 8 * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = ph.H, pairs = ph.M * H;
    const __nv_bfloat16 *qkv = (const __nv_bfloat16 *)ph.A;
    for (int nh = blockIdx.x; nh < pairs; nh += gridDim.x) {
        const int n = nh / H, h = nh % H;
        const int64_t base = (int64_t)n * ph.lda + h * 64;
        if (tid < 64) sq[tid] = phi(ldcg_bf16(qkv + base + tid));
        else if (tid < 128) sk[tid - 64] = phi(ldcg_bf16(qkv + base + H * 64 + tid - 64));
        else if (tid < 192) sv[tid - 128] = ldcg_bf16(qkv + base + 2 * H * 64 + tid - 128);
        __syncthreads();
        if (warp == 7) {
            float *z = ph.Z + (int64_t)nh * 64;
            const float z0 = z[lane] + sk[lane], z1 = z[lane + 32] + sk[lane + 32];
            z[lane] = z0; z[lane + 32] = z1;
            const float d = warp_sum(sq[lane] * z0 + sq[lane + 32] * z1);
            if (lane == 0) *sden = d + ph.eps;
        }
        const int e = tid >> 2, m0 = (tid & 3) * 16;
        float4 *srow = reinterpret_cast<float4 *>(ph.S + (int64_t)nh * 4096 + e * 64 + m0);
        float4 s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = srow[i];
        const float ke = sk[e], qe = sq[e];
        float acc[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s[i].x = fmaf(ke, sv[m0 + 4 * i + 0], s[i].x); s[i].y = fmaf(ke, sv[m0 + 4 * i + 1], s[i].y);
            s[i].z = fmaf(ke, sv[m0 + 4 * i + 2], s[i].z); s[i].w = fmaf(ke, sv[m0 + 4 * i + 3], s[i].w);
            srow[i] = s[i];
            acc[4 * i + 0] = qe * s[i].x; acc[4 * i + 1] = qe * s[i].y; acc[4 * i + 2] = qe * s[i].z; acc[4 * i + 3] = qe * s[i].w;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 4);
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
        }
        if (lane < 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) part[warp * 64 + lane * 16 + i] = acc[i];
        }
        __syncthreads();
        if (tid < 64) {
            float o = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) o += part[w * 64 + tid];
            ((__nv_bfloat16 *)ph.Y)[(int64_t)n * ph.ldy + h * 64 + tid] = __float2bfloat16_rn(o / *sden);
        }
        __syncthreads();
    }
}

struct AM { float v; int i; };
__device__ __forceinline__ AM warp_argmax2(AM a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, a.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
        if (ov > a.v || (ov == a.v && oi < a.i)) { a.v = ov; a.i = oi; }
    }
    return a;
}

// decode: one warp per (row, attribute); same arithmetic as heads_sample_kernel (heads.cu)
__device__ void sample_phase(const MegaPhase &ph, const MegaGlobals &G, const Smem &sm, int step) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *buf = sm.misc + warp * 2048, *pr = buf + 1024;
    const int items = ph.M * G.n_attr;
    const __nv_bfloat16 *logits = (const __nv_bfloat16 *)ph.A;
    for (int it = blockIdx.x * 8 + warp; it < items; it += gridDim.x * 8) {
        const int r = it / G.n_attr, a = it % G.n_attr;
        const int w = G.seg[a + 1] - G.seg[a];
        const __nv_bfloat16 *row = logits + (int64_t)r * ph.lda + G.seg[a];
        AM am{-INFINITY, 0x7fffffff};
        for (int i = lane; i < w; i += 32) {
            const float x = ldcg_bf16(row + i);
            buf[i] = x;
            if (x > am.v) { am.v = x; am.i = i; }
        }
        am = warp_argmax2(am);
        float s = 0.f;
        for (int i = lane; i < w; i += 32) s += __expf(buf[i] - am.v);
        const float lse = am.v + __logf(warp_sum(s));
        __syncwarp();
        int tok = am.i;
        if (!G.greedy) {
            const float invt = 1.f / G.temperature[a];
            float ss = 0.f;
            for (int i = lane; i < w; i += 32) { const float e = __expf((buf[i] - am.v) * invt); pr[i] = e; ss += e; }
            ss = warp_sum(ss);
            const float top_p = G.top_p[a];
            const bool nuc = top_p > 0.f && top_p < 1.f;
            const float norm = nuc ? 1.f / (ss * (1.f + 1e-5f)) : 1.f / ss;
            __syncwarp();
            for (int i = lane; i < w; i += 32) pr[i] *= norm;
            __syncwarp();
            float zkeep = 0.f, ex[32];
#pragma unroll 1
            for (int c = 0; c * 32 + lane < w; ++c) {
                const int i = c * 32 + lane;
                const float pi = pr[i];
                float e = 0.f;
                for (int j = 0; j < w; ++j) { const float pj = pr[j]; e += (pj > pi || (pj == pi && j < i)) ? pj : 0.f; }
                const bool keep = !(nuc && e > top_p);
                ex[c] = keep ? e : -1.f;
                zkeep += keep ? pi : 0.f;
            }
            zkeep = warp_sum(zkeep);
            const uint64_t sid = (uint64_t)(G.seq_base + r);
            const uint4 rnd = Philox::block(make_uint4((uint32_t)sid, (uint32_t)step, (uint32_t)a, (uint32_t)(sid >> 32)),
                                            make_uint2((uint32_t)G.seed, (uint32_t)(G.seed >> 32)));
            const float target = (float)(rnd.x >> 8) * (1.0f / 16777216.0f) * zkeep;
            AM best{-1.f, 0x7fffffff};
#pragma unroll 1
            for (int c = 0; c * 32 + lane < w; ++c) {
                const float e = ex[c];
                if (e >= 0.f && e <= target && e > best.v) { best.v = e; best.i = c * 32 + lane; }
            }
            best = warp_argmax2(best);
            tok = best.i == 0x7fffffff ? am.i : best.i;
        }
        if (lane == 0) {
            const float lp = buf[tok] - lse;
            G.cur[it] = tok;
            G.logp[it] = lp;
            if (step < G.max_steps) {
                G.hist_tok[(int64_t)step * items + it] = tok;
                G.hist_logp[(int64_t)step * items + it] = lp;
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(MG_THREADS, 1) rollout_step_megakernel(const MegaGlobals *__restrict__ Gp, const MegaPhase *__restrict__ phases) {
    extern __shared__ __align__(16) uint8_t mg_smem[];
    __shared__ MegaGlobals G;
    __shared__ MegaPhase ph, nx;
    for (int i = threadIdx.x; i < (int)(sizeof(MegaGlobals) / 4); i += MG_THREADS) reinterpret_cast<uint32_t *>(&G)[i] = reinterpret_cast<const uint32_t *>(Gp)[i];
    __syncthreads();
    Smem sm;
    sm.sA = reinterpret_cast<__nv_bfloat16 *>(mg_smem);
    sm.spart = reinterpret_cast<float *>(mg_smem + (size_t)MG_ROWS * (MG_MAXK + 8) * 2);
    sm.sGamma = sm.spart + 8 * MG_ROWS * 8;
    sm.sBeta = sm.sGamma + MG_MAXK;
    sm.misc = sm.sBeta + MG_MAXK;          // 8 warps x 2048 floats
    const int step = G.step_dev[0];
    const int pos = G.true_positions ? min(step, G.pe_max - 1) : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t bphase = 0;
    WFrag wf;
    bool wf_ready = false;
    for (int p = 0; p < G.n_phases; ++p) {
        for (int i = threadIdx.x; i < (int)(sizeof(MegaPhase) / 4); i += MG_THREADS) reinterpret_cast<uint32_t *>(&ph)[i] = reinterpret_cast<const uint32_t *>(phases + p)[i];
        __syncthreads();
        if (ph.type == 0) gemm_phase(ph, G, sm, wf, wf_ready, pos);
        else if (ph.type == 1) attn_phase(ph, sm);
        else sample_phase(ph, G, sm, step);
        wf_ready = false;
        if (p + 1 < G.n_phases) {
            // start fetching the next phase's weight fragments before waiting at the barrier
            __syncthreads();
            for (int i = threadIdx.x; i < (int)(sizeof(MegaPhase) / 4); i += MG_THREADS) reinterpret_cast<uint32_t *>(&nx)[i] = reinterpret_cast<const uint32_t *>(phases + p + 1)[i];
            __syncthreads();
            if (nx.type == 0 && (int)blockIdx.x < ((nx.N + 7) >> 3)) {
                load_wfrag(wf, nx, blockIdx.x, warp, lane >> 2, lane & 3);
                wf_ready = true;
            }
            grid_sync(G.barrier, bphase);
        }
    }
    // final arrive-only barrier: the last CTA resets the counter and advances the step counter
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t old = atomicAdd(G.barrier, 1u);
        if (old == (bphase + 1) * gridDim.x - 1) {
            G.step_dev[0] = step + 1;
            __threadfence();
            atomicExch(G.barrier, 0u);
        }
    }
}

}  // namespace
}  // namespace cpm

using namespace cpm;

extern "C" {

int cpm_mega_sizes(int *globals_bytes, int *phase_bytes) {
    if (globals_bytes) *globals_bytes = (int)sizeof(MegaGlobals);
    if (phase_bytes) *phase_bytes = (int)sizeof(MegaPhase);
    return CPM_OK;
}

int64_t cpm_mega_smem_bytes(void) {
    return (int64_t)MG_ROWS * (MG_MAXK + 8) * 2 + (int64_t)8 * MG_ROWS * 8 * 4 + (int64_t)2 * MG_MAXK * 4 + (int64_t)8 * 2048 * 4;
}

int cpm_rollout_step_mega(const void *globals_dev, const void *phases_dev, void *stream) {
    CPM_REQUIRE(globals_dev && phases_dev, CPM_ERR_NULL, "rollout_step_mega: NULL pointer");
    const size_t smem = (size_t)cpm_mega_smem_bytes();
    static int grid = 0;
    if (!grid) {
        cudaError_t e = cudaFuncSetAttribute(rollout_step_megakernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "rollout megakernel smem attribute: %s", cudaGetErrorString(e));
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_step_megakernel, MG_THREADS, smem);
        if (e != cudaSuccess || per_sm < 1) return fail(CPM_ERR_CUDA, "rollout megakernel does not fit on an SM (%s)", cudaGetErrorString(e));
        grid = num_sms();
    }
    void *args[2] = {(void *)&globals_dev, (void *)&phases_dev};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)rollout_step_megakernel, dim3(grid), dim3(MG_THREADS), args, smem, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "rollout megakernel launch: %s", cudaGetErrorString(e));
    return CPM_OK;
}

}  // extern "C"
