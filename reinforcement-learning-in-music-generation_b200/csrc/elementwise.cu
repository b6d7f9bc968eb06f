// Memory-bound fused kernels around the encoder GEMMs: CP embedding gather/scatter,
// positional-encoding add, dropout, residual+dropout+LayerNorm (fwd/bwd), bias+GELU+dropout.
// All use 128-bit accesses (8 bf16 / 2x4 fp32 per thread access) and grid-stride loops.
#include "cpm_common.cuh"

namespace cpm {
// ---------------------------------------------------------------- device-side RNG base
// Dropout masks are a pure function of (seed, offset, element).  The host passes `rng_offset` by value, which a CUDA graph
// would freeze; cpm_set_rng_base installs a device counter that every dropout kernel adds to its offset, so a captured
// training step draws fresh masks on every replay (the graph itself advances the counter; see graphs.py).
const unsigned long long *g_rng_base = nullptr;      // declared in cpm_common.cuh (the GEMM epilogues draw from the same streams)
namespace {



struct EmbedParams {
    const float *tables[CPM_MAX_ATTR];
    float *gtables[CPM_MAX_ATTR];
    int n_tokens[CPM_MAX_ATTR];
    int emb[CPM_MAX_ATTR];
    int off[CPM_MAX_ATTR + 1];
    float scale[CPM_MAX_ATTR];
    int n_attr;
};

// ---------------------------------------------------------------- C1 embedding forward
template <typename T>
__global__ void __launch_bounds__(256) embed_fwd_kernel(const int64_t *__restrict__ idx, EmbedParams p, int64_t T_, T *__restrict__ out,
                                                        int *err_flag) {
    griddep_launch();
    griddep_wait();                                     // chain kernel: idx is the previous token step's sample
    const int width = p.off[p.n_attr], G = width >> 3;
    const int64_t total = T_ * G;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / G;
        const int col = (int)(i % G) * 8;
        int a = 0;
        while (col >= p.off[a + 1]) ++a;
        const int64_t id = idx[t * p.n_attr + a];
        Vec8<T> o;
        if (id < 0 || id >= p.n_tokens[a]) {
            if (err_flag) atomicExch(err_flag, 1);
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
        } else {
            const float *src = p.tables[a] + id * p.emb[a] + (col - p.off[a]);
            float4 x = *reinterpret_cast<const float4 *>(src), y = *reinterpret_cast<const float4 *>(src + 4);
            const float s = p.scale[a];
            o.v[0] = x.x * s; o.v[1] = x.y * s; o.v[2] = x.z * s; o.v[3] = x.w * s;
            o.v[4] = y.x * s; o.v[5] = y.y * s; o.v[6] = y.z * s; o.v[7] = y.w * s;
        }
        o.store(out + t * width + col);
    }
}

// ---------------------------------------------------------------- C1 embedding backward
// grid.x = column tiles of 32 over the concatenated width, grid.y = token chunks.  Each block keeps
// a [n_tokens_a x 32] fp32 accumulator in shared memory (shared atomics), then flushes the non-zero
// rows with one global atomic per element.
template <typename T>
__global__ void __launch_bounds__(256) embed_bwd_kernel(const int64_t *__restrict__ idx, const T *__restrict__ gout, EmbedParams p,
                                                        int64_t T_, int64_t tokens_per_block) {
    extern __shared__ float acc[];
    const int width = p.off[p.n_attr];
    const int col0 = blockIdx.x * 32;
    int a = 0;
    while (col0 >= p.off[a + 1]) ++a;
    const int nrow = p.n_tokens[a];
    for (int i = threadIdx.x; i < nrow * 32; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int64_t t_begin = (int64_t)blockIdx.y * tokens_per_block;
    const int64_t t_end = min(T_, t_begin + tokens_per_block);
    for (int64_t t = t_begin + warp; t < t_end; t += nwarp) {
        const int64_t id = idx[t * p.n_attr + a];
        if (id < 0 || id >= nrow) continue;
        atomicAdd(&acc[id * 32 + lane], to_f(gout[t * width + col0 + lane]));
    }
    __syncthreads();
    float *dst = p.gtables[a] + (col0 - p.off[a]);
    const float s = p.scale[a];
    for (int i = threadIdx.x; i < nrow * 32; i += blockDim.x) {
        float g = acc[i];
        if (g != 0.f) atomicAdd(dst + (int64_t)(i >> 5) * p.emb[a] + (i & 31), g * s);
    }
}

// ---------------------------------------------------------------- PE add (+dropout)
template <typename T>
__global__ void __launch_bounds__(256) add_pe_kernel(const T *__restrict__ x, const float *__restrict__ pe, T *__restrict__ y, int64_t rows,
                                                     int L, int d, int pos_offset, const int32_t *__restrict__ pos_dev, int max_len,
                                                     uint32_t thr, float scale, uint64_t seed, uint64_t rng_offset_h, const unsigned long long *rng_base) {
    griddep_launch();
    griddep_wait();
    const uint64_t rng_offset = rng_off(rng_offset_h, rng_base);
    const int G = d >> 3;
    const int64_t total = rows * G;
    const int base = pos_dev ? pos_dev[0] : pos_offset;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / G;
        const int c = (int)(i % G) * 8;
        int pos = base + (int)(r % L);
        pos = pos < max_len ? pos : max_len - 1;
        Vec8<T> v;
        v.load(x + r * d + c);
        const float *ps = pe + (int64_t)pos * d + c;
        float4 p0 = *reinterpret_cast<const float4 *>(ps), p1 = *reinterpret_cast<const float4 *>(ps + 4);
        v.v[0] += p0.x; v.v[1] += p0.y; v.v[2] += p0.z; v.v[3] += p0.w;
        v.v[4] += p1.x; v.v[5] += p1.y; v.v[6] += p1.z; v.v[7] += p1.w;
        if (thr) {
            bool keep[8];
            dropout_mask8(seed, rng_offset, (uint64_t)i, thr, keep);
#pragma unroll
            for (int j = 0; j < 8; ++j) v.v[j] = keep[j] ? v.v[j] * scale : 0.f;
        }
        v.store(y + r * d + c);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T *__restrict__ x, T *__restrict__ y, int64_t n8, uint32_t thr, float scale,
                                                      uint64_t seed, uint64_t rng_offset_h, const unsigned long long *rng_base) {
    const uint64_t rng_offset = rng_off(rng_offset_h, rng_base);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        Vec8<T> v;
        v.load(x + i * 8);
        if (thr) {
            bool keep[8];
            dropout_mask8(seed, rng_offset, (uint64_t)i, thr, keep);
#pragma unroll
            for (int j = 0; j < 8; ++j) v.v[j] = keep[j] ? v.v[j] * scale : 0.f;
        }
        v.store(y + i * 8);
    }
}

// ---------------------------------------------------------------- C2 residual + dropout + LayerNorm
// One warp per row; the row lives in registers (MAXV vectors of 8 per lane).
template <typename T, int MAXV>
__global__ void __launch_bounds__(128) ln_residual_fwd_kernel(const T *__restrict__ x, const T *__restrict__ res, const float *__restrict__ res_bias,
                                                              const float *__restrict__ gamma,
                                                              const float *__restrict__ beta, T *__restrict__ y, T *__restrict__ s_out,
                                                              float *__restrict__ mean_out, float *__restrict__ rstd_out, int64_t rows, int d,
                                                              float eps, uint32_t thr, float scale, uint64_t seed, uint64_t rng_offset_h, const unsigned long long *rng_base) {
    griddep_launch();
    griddep_wait();
    const uint64_t rng_offset = rng_off(rng_offset_h, rng_base);
    const int lane = threadIdx.x & 31;
    const int G = d >> 3;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp_global; r < rows; r += nwarps) {
        Vec8<T> v[MAXV];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int g = lane + 32 * i;
            if (g < G) {
                v[i].load(x + r * d + g * 8);
                if (res) {
                    Vec8<T> rr;
                    rr.load(res + r * d + g * 8);
                    if (res_bias) {                   // bias of the Linear that produced `res` (its GEMM runs bias-less)
#pragma unroll
                        for (int j = 0; j < 8; ++j) rr.v[j] += res_bias[g * 8 + j];
                    }
                    if (thr) {
                        bool keep[8];
                        dropout_mask8(seed, rng_offset, (uint64_t)(r * G + g), thr, keep);
#pragma unroll
                        for (int j = 0; j < 8; ++j) rr.v[j] = keep[j] ? rr.v[j] * scale : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[i].v[j] += rr.v[j];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) sum += v[i].v[j];
            }
        }
        const float mean = warp_sum(sum) / (float)d;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (lane + 32 * i < G) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { float c = v[i].v[j] - mean; sq += c * c; }
            }
        const float rstd = rsqrtf(warp_sum(sq) / (float)d + eps);
        if (lane == 0) {
            if (mean_out) mean_out[r] = mean;
            if (rstd_out) rstd_out[r] = rstd;
        }
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int g = lane + 32 * i;
            if (g < G) {
                if (s_out) v[i].store(s_out + r * d + g * 8);
                float4 g0 = *reinterpret_cast<const float4 *>(gamma + g * 8), g1 = *reinterpret_cast<const float4 *>(gamma + g * 8 + 4);
                float4 b0 = *reinterpret_cast<const float4 *>(beta + g * 8), b1 = *reinterpret_cast<const float4 *>(beta + g * 8 + 4);
                const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                Vec8<T> o;
#pragma unroll
                for (int j = 0; j < 8; ++j) o.v[j] = (v[i].v[j] - mean) * rstd * gm[j] + bt[j];
                o.store(y + r * d + g * 8);
            }
        }
    }
}

constexpr int LN_BWD_BLOCKS = 296;   // 2 CTAs per SM on 148 SMs (126 registers at d = 512; 3 per SM spills and is slower)
inline int ln_bwd_blocks(int) { return LN_BWD_BLOCKS; }
constexpr int GELU_BWD_BLOCKS = 1184; // 8 CTAs per SM: the fused-bias-gradient variant keeps its partial matrix small
constexpr int LN_BWD_THREADS = 256;

template <typename T, int MAXV, bool WANT_DRES>
__global__ void __launch_bounds__(LN_BWD_THREADS, MAXV <= 2 ? 2 : 1) ln_residual_bwd_kernel(const T *__restrict__ gy, const T *__restrict__ s, const float *__restrict__ mean,
                                                                         const float *__restrict__ rstd, const float *__restrict__ gamma, T *__restrict__ gs,
                                                                         T *__restrict__ gres, float *__restrict__ partials, int64_t rows, int d, uint32_t thr,
                                                                         float scale, uint64_t seed, uint64_t rng_offset_h, const unsigned long long *rng_base) {
    const uint64_t rng_offset = rng_off(rng_offset_h, rng_base);
    extern __shared__ __align__(16) float red[];   // [3][d] block partial sums: dgamma | dbeta | column sums of the residual-branch gradient; [d] gamma
    float *sgam = red + 3 * d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = d >> 3;
    for (int i = threadIdx.x; i < 3 * d; i += blockDim.x) red[i] = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) sgam[i] = gamma[i];
    __syncthreads();
    // d <= 512: the NEXT row's gy / s (raw 16-byte words), mean and rstd are fetched before the current row's math, so a
    // warp always has one row in flight (16 warps per SM at 126 registers: one row per warp is not enough to cover HBM latency)
    constexpr bool PIPE = MAXV <= 2;
    float dg[MAXV][8], db[MAXV][8], dr[MAXV][8];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { dg[i][j] = 0.f; db[i][j] = 0.f; dr[i][j] = 0.f; }
    }
    const int64_t warp_global = (int64_t)blockIdx.x * (LN_BWD_THREADS / 32) + warp;
    const int64_t nwarps = (int64_t)gridDim.x * (LN_BWD_THREADS / 32);
    Raw8<T> ng[PIPE ? MAXV : 1], nx[PIPE ? MAXV : 1];
    float nmu = 0.f, nrs = 0.f;
    auto fetch = [&](int64_t r) {
        if (r < rows) {
            nmu = mean[r];
            nrs = rstd[r];
#pragma unroll
            for (int i = 0; i < (PIPE ? MAXV : 1); ++i) {
                const int g = lane + 32 * i;
                if (g < G) { ng[i].load(gy + r * d + g * 8); nx[i].load(s + r * d + g * 8); }
            }
        }
    };
    if (PIPE) fetch(warp_global);
    for (int64_t r = warp_global; r < rows; r += nwarps) {
        const float mu = PIPE ? nmu : mean[r], rs = PIPE ? nrs : rstd[r];
        Vec8<T> g_[MAXV], x_[MAXV];
        if (PIPE) {
#pragma unroll
            for (int i = 0; i < MAXV; ++i)
                if (lane + 32 * i < G) { ng[i < (PIPE ? MAXV : 1) ? i : 0].unpack(g_[i].v); nx[i < (PIPE ? MAXV : 1) ? i : 0].unpack(x_[i].v); }
            fetch(r + nwarps);
        }
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int g = lane + 32 * i;
            if (g < G) {
                if (!PIPE) {
                    g_[i].load(gy + r * d + g * 8);
                    x_[i].load(s + r * d + g * 8);
                }
                const float4 gm0 = *reinterpret_cast<const float4 *>(sgam + g * 8), gm1 = *reinterpret_cast<const float4 *>(sgam + g * 8 + 4);
                const float gm[8] = {gm0.x, gm0.y, gm0.z, gm0.w, gm1.x, gm1.y, gm1.z, gm1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (x_[i].v[j] - mu) * rs;
                    const float gg = g_[i].v[j];
                    dg[i][j] += gg * xh;
                    db[i][j] += gg;
                    const float w = gg * gm[j];
                    c1 += w * xh;
                    c2 += w;
                    x_[i].v[j] = xh;
                    g_[i].v[j] = w;
                }
            }
        }
        c1 = warp_sum(c1) / (float)d;
        c2 = warp_sum(c2) / (float)d;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int g = lane + 32 * i;
            if (g < G) {
                Vec8<T> o;
#pragma unroll
                for (int j = 0; j < 8; ++j) o.v[j] = rs * (g_[i].v[j] - c2 - x_[i].v[j] * c1);
                o.store(gs + r * d + g * 8);
                if (gres && thr) {
                    bool keep[8];
                    dropout_mask8(seed, rng_offset, (uint64_t)(r * G + g), thr, keep);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o.v[j] = keep[j] ? o.v[j] * scale : 0.f;
                    o.store(gres + r * d + g * 8);
                }
                if (WANT_DRES) {                      // d(loss)/d(res_bias) = column sums of the residual-branch gradient
                    Vec8<T> rounded;                  // as stored (what a reduction over the stored tensor would see)
                    rounded = o;
#pragma unroll
                    for (int j = 0; j < 8; ++j) dr[i][j] += to_f(from_f<T>(rounded.v[j]));
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int g = lane + 32 * i;
        if (g < G) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&red[g * 8 + j], dg[i][j]);
                atomicAdd(&red[d + g * 8 + j], db[i][j]);
                if (WANT_DRES) atomicAdd(&red[2 * d + g * 8 + j], dr[i][j]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * d; i += blockDim.x) partials[(int64_t)blockIdx.x * 3 * d + i] = red[i];
}

// out_k[c] (+)= sum over `nrows` partial rows of width `width` (k = c / d selects out0 | out1 | out2; NULL = skip).
// One CTA per 32-column slab, 8 row groups, 8 independent loads in flight per thread; fixed order (deterministic).
template <bool ACCUM>
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float *__restrict__ partials, int nrows, int width, int d, float *__restrict__ out0,
                                                              float *__restrict__ out1, float *__restrict__ out2) {
    __shared__ float sm[8][32];
    const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;          // 32 columns x 8 row groups
    const int c = blockIdx.x * 32 + cl;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < width) {
        int r = rg;
        for (; r + 56 < nrows; r += 64) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += partials[(int64_t)(r + 8 * i) * width + c];
        }
        for (; r < nrows; r += 8) a[0] += partials[(int64_t)r * width + c];
    }
    sm[rg][cl] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    __syncthreads();
    if (rg == 0 && c < width) {
        const float tot = ((sm[0][cl] + sm[1][cl]) + (sm[2][cl] + sm[3][cl])) + ((sm[4][cl] + sm[5][cl]) + (sm[6][cl] + sm[7][cl]));
        float *out = c < d ? out0 : (c < 2 * d ? out1 : out2);
        if (out) out[c % d] = ACCUM ? out[c % d] + tot : tot;
    }
}

// ---------------------------------------------------------------- row dot: out[r] = h[r,:] . u + c   (collapsed critic value head)
// Critic_Transformer.value_produce (ppo_policy/model.py:345-394) applies six Linear(512 -> n_a) heads and six Linear(n_a -> 1)
// value heads and averages: linear in h, so value[r] = h[r,:] . u + c with u = sum_a W_a^T w_a / 6 (model.py builds u, c with
// autograd through that tiny product).  One warp per row, 128-bit loads; fp32 result.  Backward: dh[r,:] = g[r] u (compute
// dtype) and du = sum_r g[r] h[r,:] through per-CTA partial rows + reduce_partials_kernel (deterministic).
constexpr int ROWDOT_BWD_BLOCKS = 592;
template <typename T, int MAXV>
__global__ void __launch_bounds__(256) rowdot_fwd_kernel(const T *__restrict__ h, const float *__restrict__ u, const float *__restrict__ c,
                                                         float *__restrict__ out, int64_t rows, int d) {
    const int lane = threadIdx.x & 31, G = d >> 3;
    float uv[MAXV][8];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int g = lane + 32 * i;
#pragma unroll
        for (int j = 0; j < 8; ++j) uv[i][j] = g < G ? u[g * 8 + j] : 0.f;
    }
    const float c0 = c ? *c : 0.f;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp_global; r < rows; r += nwarps) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int g = lane + 32 * i;
            if (g < G) {
                Vec8<T> v;
                v.load(h + r * d + g * 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc = fmaf(v.v[j], uv[i][j], acc);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) out[r] = acc + c0;
    }
}
template <typename T, int MAXV>
__global__ void __launch_bounds__(256) rowdot_bwd_kernel(const T *__restrict__ h, const float *__restrict__ g, const float *__restrict__ u,
                                                         T *__restrict__ dh, float *__restrict__ partials, int64_t rows, int d) {
    __shared__ float red[8][32 * MAXV * 8 + 8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, G = d >> 3;
    float uv[MAXV][8], du[MAXV][8];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int gi = lane + 32 * i;
#pragma unroll
        for (int j = 0; j < 8; ++j) { uv[i][j] = gi < G ? u[gi * 8 + j] : 0.f; du[i][j] = 0.f; }
    }
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp_global; r < rows; r += nwarps) {
        const float gr = g[r];
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int gi = lane + 32 * i;
            if (gi < G) {
                Vec8<T> v, o;
                v.load(h + r * d + gi * 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) { du[i][j] = fmaf(gr, v.v[j], du[i][j]); o.v[j] = gr * uv[i][j]; }
                if (dh) o.store(dh + r * d + gi * 8);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) red[wib][(lane + 32 * i) * 8 + j] = du[i][j];
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += 256) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][c];
        partials[(int64_t)blockIdx.x * d + c] = t;
    }
}

// ---------------------------------------------------------------- column sums (bias gradients of the Linear layers)
// out[c] = sum_r x[r][c], x (rows, width) bf16 / fp32 row-major with row stride ld.  Stage 1: CTA (slab, rs) sums its rows of a
// 1 KB-wide column slab: one warp spans the slab with 256-bit streaming loads (L1 no-allocate, L2 evict-first: measured 5.5-5.9
// TB/s on a pure read against 4.1-4.6 TB/s for plain 128-bit loads, tools/probes/probe_read_bw.cu), 8 row lanes, 4 loads in
// flight per thread; it writes one partial row.  Stage 2 is reduce_partials_kernel.  Deterministic; 2 launches.
// Rows whose address is not 32-byte aligned (odd ld / column offset) take 128-bit loads.
template <typename T> struct ColsumGeo { static constexpr int CPT = 32 / (int)sizeof(T), SLAB = 32 * CPT; };
inline int colsum_row_slabs(int width, int slab) {          // ~8 CTAs per SM in total whatever the width
    const int slabs = (width + slab - 1) / slab, r = (1184 + slabs - 1) / slabs;
    return r < 1 ? 1 : (r > 256 ? 256 : r);
}
template <typename T, bool WIDE>
__device__ __forceinline__ void colsum_load(const T *p, float (&v)[32 / sizeof(T)]) {
    constexpr int CPT = 32 / (int)sizeof(T);
    uint32_t w[8];
    if (WIDE) {
        asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
    } else {
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "l"(reinterpret_cast<const char *>(p) + 16));
    }
    if (sizeof(T) == 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j % CPT] = __uint_as_float(w[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[(2 * j) % CPT] = __uint_as_float(w[j] << 16); v[(2 * j + 1) % CPT] = __uint_as_float(w[j] & 0xFFFF0000u); }
    }
}
template <typename T, bool WIDE>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T *__restrict__ x, int64_t rows, int width, int64_t ld, float *__restrict__ partials) {
    constexpr int CPT = ColsumGeo<T>::CPT, SLAB = ColsumGeo<T>::SLAB;
    __shared__ float sm[8][SLAB + 4];
    const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c0 = blockIdx.x * SLAB + cg * CPT;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = min(rows, r0 + per);
    float acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
    if (c0 < width) {                                  // width % CPT == 0: a thread's columns are all in or all out
        int64_t r = r0 + rl;
        for (; r + 24 < r1; r += 32) {
            float v[4][CPT];
#pragma unroll
            for (int i = 0; i < 4; ++i) colsum_load<T, WIDE>(x + (r + 8 * i) * ld + c0, v[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < CPT; ++j) acc[j] += v[i][j];
        }
        for (; r < r1; r += 8) {
            float v[CPT];
            colsum_load<T, WIDE>(x + r * ld + c0, v);
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] += v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) sm[rl][cg * CPT + j] = acc[j];
    __syncthreads();
    for (int cl = threadIdx.x; cl < SLAB; cl += 256) {
        const int c = blockIdx.x * SLAB + cl;
        if (c < width)
            partials[(int64_t)blockIdx.y * width + c] = ((sm[0][cl] + sm[1][cl]) + (sm[2][cl] + sm[3][cl])) + ((sm[4][cl] + sm[5][cl]) + (sm[6][cl] + sm[7][cl]));
    }
}

// ---------------------------------------------------------------- bias + exact GELU + dropout (helpers: cpm_common.cuh)
// 16 elements per thread and iteration (two 128-bit loads in flight per operand, one Philox block)
// dbias_partials (BWD only, optional): row (blockIdx * (4096 / d) + (tid * 16) / d) of a [gridDim * 4096 / d][d] fp32 matrix receives
// this thread's column sums of the stored gx — valid because 4096 % d == 0 makes a thread's 16 columns the same in every iteration.
template <typename T, bool BWD, bool DBIAS = false>
__global__ void __launch_bounds__(256) gelu_kernel(const T *__restrict__ x, const float *__restrict__ bias, const T *__restrict__ gy, T *__restrict__ out,
                                                   int64_t n_groups, int d, uint32_t thr8, float scale, uint64_t seed, uint64_t rng_offset_h,
                                                   float *__restrict__ dbias_partials, const unsigned long long *rng_base) {
    const uint64_t rng_offset = rng_off(rng_offset_h, rng_base);
    float cs[16], bv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { cs[j] = 0.f; bv[j] = 0.f; }
    // 4096 % d == 0: the grid stride (a multiple of 256 threads x 16 elements) is a multiple of the row width, so a thread
    // sees the same 16 columns in every iteration and its bias values are loaded once.
    const bool fixed_cols = bias && (4096 % d == 0);
    if (fixed_cols) {
        const int c = (threadIdx.x * 16) % d;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4 *>(bias + c + j);
            bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
        }
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_groups; i += (int64_t)gridDim.x * blockDim.x) {
        Vec8<T> v0, v1, g0, g1;
        v0.load(x + i * 16);
        v1.load(x + i * 16 + 8);
        if (BWD) { g0.load(gy + i * 16); g1.load(gy + i * 16 + 8); }
        if (bias && !fixed_cols) {
            const int c = (int)((i * 16) % d);
#pragma unroll
            for (int j = 0; j < 16; ++j) bv[j] = bias[c + j];
        }
        if (bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { v0.v[j] += bv[j]; v1.v[j] += bv[8 + j]; }
        }
        const uint32_t keep = thr8 ? dropout_keep16(seed, rng_offset, (uint64_t)i, thr8) : 0xFFFFu;
        Vec8<T> o0, o1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            constexpr bool EX = sizeof(T) == 4;
            const float r0 = BWD ? g0.v[j] * dgelu_f<EX>(v0.v[j]) : gelu_f<EX>(v0.v[j]);
            const float r1 = BWD ? g1.v[j] * dgelu_f<EX>(v1.v[j]) : gelu_f<EX>(v1.v[j]);
            o0.v[j] = (keep >> j) & 1u ? r0 * scale : 0.f;
            o1.v[j] = (keep >> (8 + j)) & 1u ? r1 * scale : 0.f;
        }
        o0.store(out + i * 16);
        o1.store(out + i * 16 + 8);
        if (DBIAS) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { cs[j] += to_f(from_f<T>(o0.v[j])); cs[8 + j] += to_f(from_f<T>(o1.v[j])); }
        }
    }
    if (DBIAS) {
        const int per = 4096 / d;
        float *dst = dbias_partials + ((int64_t)blockIdx.x * per + (threadIdx.x * 16) / d) * d + (threadIdx.x * 16) % d;
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(cs[j], cs[j + 1], cs[j + 2], cs[j + 3]);
    }
}

// ---------------------------------------------------------------- packed compute copies of the fp32 masters, ONE launch
// A model's Linear layers keep bf16 packings of their fp32 master weights (row-concatenated q / k / v, heads, ...): W (N x K) for
// the forward GEMM, its transpose (K x N) for the data-gradient GEMM, the bias in bf16 and fp32.  After an optimizer step all of
// them are stale at once; refreshing them Linear by Linear took five to eight tiny copy kernels each (~240 launches per model).
// Here every 32 x 32 tile of every master is one thread block of a single launch: coalesced fp32 read, bf16 row-major write,
// transposed write through shared memory; the blocks of a tile column 0 also copy the bias rows.
struct PackItem {
    const float *w, *b;          // master weight (rows x cols, contiguous) and bias (rows) or NULL
    __nv_bfloat16 *wc, *wt, *bc; // destinations: wc + r0 * cols (row-major), wt + r0 (K x ld_t, transposed), bc + r0; wt / bc may be NULL
    float *b32;                  // fp32 bias destination (+ r0) or NULL
    int rows, cols, ld_t, tile0; // ld_t: row stride of wt (the packing's padded row count); tile0: first block of this item
};
__global__ void __launch_bounds__(256) pack_weights_kernel(const PackItem *__restrict__ items, int n_items) {
    __shared__ float tile[32][33];
    __shared__ int s_item;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        int lo = 0, hi = n_items - 1;                          // last item whose tile0 <= blockIdx.x
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (items[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1; }
        s_item = lo;
    }
    __syncthreads();
    const PackItem it = items[s_item];
    const int tcols = (it.cols + 31) >> 5, t = blockIdx.x - it.tile0, tr = t / tcols, tc = t % tcols;
    const int r0 = tr * 32, c0 = tc * 32, tx = threadIdx.x, ty = threadIdx.y;      // block (32, 8)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = r0 + ty + 8 * j, c = c0 + tx;
        float v = 0.f;
        if (r < it.rows && c < it.cols) {
            v = it.w[(int64_t)r * it.cols + c];
            it.wc[(int64_t)r * it.cols + c] = __float2bfloat16_rn(v);
        }
        tile[ty + 8 * j][tx] = v;
    }
    __syncthreads();
    if (it.wt) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + ty + 8 * j, r = r0 + tx;             // wt[c][r] = w[r][c]
            if (r < it.rows && c < it.cols) it.wt[(int64_t)c * it.ld_t + r] = __float2bfloat16_rn(tile[tx][ty + 8 * j]);
        }
    }
    if (tc == 0 && ty == 0 && it.b && r0 + tx < it.rows) {
        const float bv = it.b[r0 + tx];
        if (it.bc) it.bc[r0 + tx] = __float2bfloat16_rn(bv);
        if (it.b32) it.b32[r0 + tx] = bv;
    }
}

inline int grid_for(int64_t work_items, int threads) {
    int64_t b = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

int fill_embed_params(EmbedParams &p, const float *const *tables, float *const *gtables, const int *n_tokens, const int *emb,
                      int n_attr, int mult) {
    CPM_REQUIRE(n_attr >= 1 && n_attr <= CPM_MAX_ATTR, CPM_ERR_BAD_SHAPE, "embed: n_attr=%d out of [1,%d]", n_attr, CPM_MAX_ATTR);
    p.n_attr = n_attr;
    p.off[0] = 0;
    for (int a = 0; a < n_attr; ++a) {
        CPM_REQUIRE(emb[a] > 0 && emb[a] % mult == 0, CPM_ERR_BAD_SHAPE, "embed: emb_sizes[%d]=%d must be a multiple of %d", a, emb[a], mult);
        CPM_REQUIRE(n_tokens[a] > 0, CPM_ERR_BAD_SHAPE, "embed: n_tokens[%d]=%d", a, n_tokens[a]);
        p.tables[a] = tables ? tables[a] : nullptr;
        p.gtables[a] = gtables ? gtables[a] : nullptr;
        CPM_REQUIRE(p.tables[a] || p.gtables[a], CPM_ERR_NULL, "embed: table %d is NULL", a);
        p.n_tokens[a] = n_tokens[a];
        p.emb[a] = emb[a];
        p.off[a + 1] = p.off[a] + emb[a];
        p.scale[a] = sqrtf((float)emb[a]);
    }
    return CPM_OK;
}

}  // namespace
}  // namespace cpm

using namespace cpm;

#define DISPATCH_DTYPE(dtype, ...)                                              \
    if ((dtype) == CPM_F32) { using T = float; __VA_ARGS__; }                   \
    else if ((dtype) == CPM_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }     \
    else return fail(CPM_ERR_BAD_DTYPE, "unsupported dtype %d", (int)(dtype));

#define DISPATCH_MAXV(d, ...)                                                   \
    if ((d) <= 256) { constexpr int MAXV = 1; __VA_ARGS__; }                    \
    else if ((d) <= 512) { constexpr int MAXV = 2; __VA_ARGS__; }               \
    else if ((d) <= 1024) { constexpr int MAXV = 4; __VA_ARGS__; }              \
    else { constexpr int MAXV = 8; __VA_ARGS__; }

extern "C" {

int cpm_embed_fwd(const int64_t *idx, const float *const *tables_host, const int *n_tokens_host, const int *emb_sizes_host,
                  int n_attr, int64_t T_, void *out, int dtype, int *err_flag, void *stream) {
    CPM_REQUIRE(T_ >= 0, CPM_ERR_BAD_SHAPE, "embed_fwd: T=%lld", (long long)T_);
    if (T_ == 0) return CPM_OK;                 // empty batch: nothing to read or write
    CPM_REQUIRE(idx && tables_host && n_tokens_host && emb_sizes_host && out, CPM_ERR_NULL, "embed_fwd: NULL pointer");
    EmbedParams p{};
    int rc = fill_embed_params(p, tables_host, nullptr, n_tokens_host, emb_sizes_host, n_attr, 8);
    if (rc) return rc;
    CPM_REQUIRE(aligned16(out), CPM_ERR_BAD_ALIGN, "embed_fwd: out not 16-byte aligned");
    const int64_t items = T_ * (p.off[n_attr] / 8);
    DISPATCH_DTYPE(dtype, launch_chain(embed_fwd_kernel<T>, dim3(grid_for(items, 256)), dim3(256), 0, (cudaStream_t)stream, idx, p, T_, (T *)out, err_flag));
    return check_launch("embed_fwd");
}

int cpm_embed_bwd(const int64_t *idx, const void *gout, float *const *gtables_host, const int *n_tokens_host,
                  const int *emb_sizes_host, int n_attr, int64_t T_, int dtype, void *stream) {
    if (T_ == 0) return CPM_OK;
    CPM_REQUIRE(idx && gout && gtables_host && n_tokens_host && emb_sizes_host, CPM_ERR_NULL, "embed_bwd: NULL pointer");
    EmbedParams p{};
    int rc = fill_embed_params(p, nullptr, gtables_host, n_tokens_host, emb_sizes_host, n_attr, 32);
    if (rc) return rc;
    int max_rows = 0;
    for (int a = 0; a < n_attr; ++a) max_rows = n_tokens_host[a] > max_rows ? n_tokens_host[a] : max_rows;
    const size_t smem = (size_t)max_rows * 32 * sizeof(float);
    CPM_REQUIRE(smem <= 160 * 1024, CPM_ERR_UNSUPPORTED, "embed_bwd: vocabulary of %d rows exceeds the shared-memory accumulator", max_rows);
    int ychunks = (int)((T_ + 1023) / 1024);
    if (ychunks > 64) ychunks = 64;
    const int64_t tpb = (T_ + ychunks - 1) / ychunks;
    dim3 grid(p.off[n_attr] / 32, ychunks);
    DISPATCH_DTYPE(dtype, {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(embed_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "embed_bwd smem: %s", cudaGetErrorString(e));
        }
        embed_bwd_kernel<T><<<grid, 256, smem, (cudaStream_t)stream>>>(idx, (const T *)gout, p, T_, tpb);
    });
    return check_launch("embed_bwd");
}

int cpm_add_pe(const void *x, const float *pe, void *y, int64_t rows, int L, int d, int pos_offset, const int32_t *pos_dev,
               int max_len, float p_drop, uint64_t seed, uint64_t rng_offset, int dtype, void *stream) {
    CPM_REQUIRE(x && pe && y, CPM_ERR_NULL, "add_pe: NULL pointer");
    CPM_REQUIRE(rows >= 0 && L > 0 && d > 0 && d % 8 == 0 && max_len > 0, CPM_ERR_BAD_SHAPE, "add_pe: rows=%lld L=%d d=%d", (long long)rows, L, d);
    CPM_REQUIRE(pos_dev || pos_offset + L <= max_len, CPM_ERR_BAD_SHAPE, "add_pe: positions %d..%d exceed max_len %d", pos_offset, pos_offset + L, max_len);
    CPM_REQUIRE(aligned16(x) && aligned16(y) && aligned16(pe), CPM_ERR_BAD_ALIGN, "add_pe: alignment");
    if (rows == 0) return CPM_OK;
    const uint32_t thr = dropout_threshold(p_drop);
    DISPATCH_DTYPE(dtype, launch_chain(add_pe_kernel<T>, dim3(grid_for(rows * (d / 8), 256)), dim3(256), 0, (cudaStream_t)stream,
                                       (const T *)x, pe, (T *)y, rows, L, d, pos_offset, pos_dev, max_len, thr, dropout_scale(p_drop), seed, rng_offset, g_rng_base));
    return check_launch("add_pe");
}

int cpm_dropout(const void *x, void *y, int64_t n, float p_drop, uint64_t seed, uint64_t rng_offset, int dtype, void *stream) {
    CPM_REQUIRE(x && y, CPM_ERR_NULL, "dropout: NULL pointer");
    CPM_REQUIRE(n >= 0 && n % 8 == 0, CPM_ERR_BAD_SHAPE, "dropout: n=%lld must be a multiple of 8", (long long)n);
    CPM_REQUIRE(aligned16(x) && aligned16(y), CPM_ERR_BAD_ALIGN, "dropout: alignment");
    if (n == 0) return CPM_OK;
    DISPATCH_DTYPE(dtype, dropout_kernel<T><<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
                              (const T *)x, (T *)y, n / 8, dropout_threshold(p_drop), dropout_scale(p_drop), seed, rng_offset, g_rng_base));
    return check_launch("dropout");
}

int cpm_ln_residual_fwd(const void *x, const void *res, const float *res_bias, const float *gamma, const float *beta, void *y, void *s_out,
                        float *mean, float *rstd, int64_t rows, int d, float eps, float p_drop, uint64_t seed, uint64_t rng_offset, int dtype,
                        void *stream) {
    CPM_REQUIRE(x && gamma && beta && y, CPM_ERR_NULL, "ln_residual_fwd: NULL pointer");
    CPM_REQUIRE(rows >= 0 && d > 0 && d % 8 == 0 && d <= 2048, CPM_ERR_BAD_SHAPE, "ln_residual_fwd: d=%d must be a multiple of 8 and <= 2048", d);
    CPM_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) && (!res || aligned16(res)) && (!s_out || aligned16(s_out)),
                CPM_ERR_BAD_ALIGN, "ln_residual_fwd: alignment");
    if (rows == 0) return CPM_OK;
    const uint32_t thr = res ? dropout_threshold(p_drop) : 0u;
    const int grid = grid_for(rows * 32, 128);
    DISPATCH_DTYPE(dtype, DISPATCH_MAXV(d, launch_chain(ln_residual_fwd_kernel<T, MAXV>, dim3(grid), dim3(128), 0, (cudaStream_t)stream,
                                               (const T *)x, (const T *)res, res ? res_bias : nullptr, gamma, beta, (T *)y, (T *)s_out, mean, rstd, rows, d, eps, thr,
                                               dropout_scale(p_drop), seed, rng_offset, g_rng_base)));
    return check_launch("ln_residual_fwd");
}

int cpm_ln_partials_rows(void) { return LN_BWD_BLOCKS; }

int cpm_rowdot_partials_rows(void) { return ROWDOT_BWD_BLOCKS; }

int cpm_rowdot_fwd(const void *h, const float *u, const float *c, float *out, int64_t rows, int d, int dtype, void *stream) {
    CPM_REQUIRE(h && u && out, CPM_ERR_NULL, "rowdot_fwd: NULL pointer");
    CPM_REQUIRE(rows >= 0 && d > 0 && d % 8 == 0 && d <= 1024, CPM_ERR_BAD_SHAPE, "rowdot_fwd: d=%d must be a multiple of 8 and <= 1024", d);
    CPM_REQUIRE(aligned16(h), CPM_ERR_BAD_ALIGN, "rowdot_fwd: h must be 16-byte aligned");
    if (rows == 0) return CPM_OK;
    const int grid = grid_for(rows * 32, 256);
    if (d <= 512) { DISPATCH_DTYPE(dtype, (rowdot_fwd_kernel<T, 2><<<grid, 256, 0, (cudaStream_t)stream>>>((const T *)h, u, c, out, rows, d))); }
    else { DISPATCH_DTYPE(dtype, (rowdot_fwd_kernel<T, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const T *)h, u, c, out, rows, d))); }
    return check_launch("rowdot_fwd");
}

int cpm_rowdot_bwd(const void *h, const float *g, const float *u, void *dh, float *du, float *partials, int64_t rows, int d, int dtype,
                   void *stream) {
    CPM_REQUIRE(h && g && u && du && partials, CPM_ERR_NULL, "rowdot_bwd: NULL pointer");
    CPM_REQUIRE(rows >= 0 && d > 0 && d % 8 == 0 && d <= 1024, CPM_ERR_BAD_SHAPE, "rowdot_bwd: d=%d must be a multiple of 8 and <= 1024", d);
    CPM_REQUIRE(aligned16(h) && (!dh || aligned16(dh)), CPM_ERR_BAD_ALIGN, "rowdot_bwd: alignment");
    if (d <= 512) { DISPATCH_DTYPE(dtype, (rowdot_bwd_kernel<T, 2><<<ROWDOT_BWD_BLOCKS, 256, 0, (cudaStream_t)stream>>>((const T *)h, g, u, (T *)dh, partials, rows, d))); }
    else { DISPATCH_DTYPE(dtype, (rowdot_bwd_kernel<T, 4><<<ROWDOT_BWD_BLOCKS, 256, 0, (cudaStream_t)stream>>>((const T *)h, g, u, (T *)dh, partials, rows, d))); }
    reduce_partials_kernel<false><<<(d + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partials, ROWDOT_BWD_BLOCKS, d, d, du, nullptr, nullptr);
    return check_launch("rowdot_bwd");
}

int cpm_set_rng_base(const uint64_t *device_counter) {
    g_rng_base = reinterpret_cast<const unsigned long long *>(device_counter);
    return CPM_OK;
}

int cpm_ln_residual_bwd(const void *gy, const void *s, const float *mean, const float *rstd, const float *gamma, void *gs, void *gres,
                        float *dgamma, float *dbeta, float *dres_bias, float *partials, int64_t rows, int d, float p_drop, uint64_t seed,
                        uint64_t rng_offset, int dtype, void *stream) {
    CPM_REQUIRE(gy && s && mean && rstd && gamma && gs && dgamma && dbeta && partials, CPM_ERR_NULL, "ln_residual_bwd: NULL pointer");
    CPM_REQUIRE(rows >= 0 && d > 0 && d % 8 == 0 && d <= 2048, CPM_ERR_BAD_SHAPE, "ln_residual_bwd: d=%d", d);
    CPM_REQUIRE(aligned16(gy) && aligned16(s) && aligned16(gs) && (!gres || aligned16(gres)), CPM_ERR_BAD_ALIGN, "ln_residual_bwd: alignment");
    const uint32_t thr = dropout_threshold(p_drop);
    CPM_REQUIRE(!thr || gres, CPM_ERR_NULL, "ln_residual_bwd: gres required when p_drop > 0");
    if (rows == 0) return CPM_OK;
    const size_t smem = 4 * (size_t)d * sizeof(float);
    if (dres_bias) {
        DISPATCH_DTYPE(dtype, DISPATCH_MAXV(d, ln_residual_bwd_kernel<T, MAXV, true><<<ln_bwd_blocks(d), LN_BWD_THREADS, smem, (cudaStream_t)stream>>>(
                                                   (const T *)gy, (const T *)s, mean, rstd, gamma, (T *)gs, (T *)gres, partials, rows, d, thr,
                                                   dropout_scale(p_drop), seed, rng_offset, g_rng_base)));
    } else {
        DISPATCH_DTYPE(dtype, DISPATCH_MAXV(d, ln_residual_bwd_kernel<T, MAXV, false><<<ln_bwd_blocks(d), LN_BWD_THREADS, smem, (cudaStream_t)stream>>>(
                                                   (const T *)gy, (const T *)s, mean, rstd, gamma, (T *)gs, (T *)gres, partials, rows, d, thr,
                                                   dropout_scale(p_drop), seed, rng_offset, g_rng_base)));
    }
    const int width = (dres_bias ? 3 : 2) * d;
    reduce_partials_kernel<true><<<(width + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partials, ln_bwd_blocks(d), 3 * d, d, dgamma, dbeta, dres_bias);
    return check_launch("ln_residual_bwd");
}

int cpm_pack_item_bytes(void) { return (int)sizeof(PackItem); }

int cpm_pack_weights(const void *items_device, int n_items, int n_tiles, void *stream) {
    CPM_REQUIRE(items_device && n_items > 0 && n_tiles > 0, CPM_ERR_NULL, "pack_weights: empty item table");
    pack_weights_kernel<<<n_tiles, dim3(32, 8), 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackItem *>(items_device), n_items);
    return check_launch("pack_weights");
}

int cpm_colsum_partials_rows(int width) { return width > 0 ? 256 : 0; }      // upper bound of the row slabs for any dtype

int cpm_colsum(const void *x, int64_t rows, int width, int64_t ld, float *out, float *partials, int dtype, void *stream) {
    CPM_REQUIRE(x && out && partials, CPM_ERR_NULL, "colsum: NULL pointer");
    CPM_REQUIRE(dtype == CPM_F32 || dtype == CPM_BF16, CPM_ERR_BAD_DTYPE, "colsum: dtype %d", dtype);
    const int cpt = dtype == CPM_F32 ? 8 : 16, slab = 32 * cpt, esz = dtype == CPM_F32 ? 4 : 2;
    CPM_REQUIRE(rows >= 0 && width > 0 && width % cpt == 0 && ld >= width && ld % 8 == 0, CPM_ERR_BAD_SHAPE,
                "colsum: width=%d must be a multiple of %d, ld=%lld of 8", width, cpt, (long long)ld);
    CPM_REQUIRE(aligned16(x), CPM_ERR_BAD_ALIGN, "colsum: x must be 16-byte aligned");
    const int R = colsum_row_slabs(width, slab);
    const dim3 grid((width + slab - 1) / slab, R);
    const bool wide = (reinterpret_cast<uintptr_t>(x) % 32 == 0) && ((ld * esz) % 32 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CPM_F32) {
        if (wide) colsum_partial_kernel<float, true><<<grid, 256, 0, st>>>((const float *)x, rows, width, ld, partials);
        else colsum_partial_kernel<float, false><<<grid, 256, 0, st>>>((const float *)x, rows, width, ld, partials);
    } else {
        if (wide) colsum_partial_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, rows, width, ld, partials);
        else colsum_partial_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, rows, width, ld, partials);
    }
    reduce_partials_kernel<false><<<(width + 31) / 32, 256, 0, st>>>(partials, R, width, width, out, nullptr, nullptr);
    return check_launch("colsum");
}

int cpm_gelu_fwd(const void *x, const float *bias, void *y, int64_t rows, int d, float p_drop, uint64_t seed, uint64_t rng_offset,
                 int dtype, void *stream) {
    CPM_REQUIRE(x && y, CPM_ERR_NULL, "gelu_fwd: NULL pointer");
    CPM_REQUIRE(rows >= 0 && d > 0 && d % 16 == 0, CPM_ERR_BAD_SHAPE, "gelu_fwd: d=%d must be a multiple of 16", d);
    CPM_REQUIRE(aligned16(x) && aligned16(y) && (!bias || aligned16(bias)), CPM_ERR_BAD_ALIGN, "gelu_fwd: alignment");
    if (rows == 0) return CPM_OK;
    DISPATCH_DTYPE(dtype, gelu_kernel<T, false><<<grid_for(rows * (d / 16), 256), 256, 0, (cudaStream_t)stream>>>(
                              (const T *)x, bias, nullptr, (T *)y, rows * (d / 16), d, dropout_threshold8(p_drop), dropout_scale8(p_drop), seed,
                              rng_offset, nullptr, g_rng_base));
    return check_launch("gelu_fwd");
}

int cpm_gelu_bwd_partials_rows(int d) { return (d > 0 && 4096 % d == 0) ? GELU_BWD_BLOCKS * (4096 / d) : 0; }

int cpm_gelu_bwd(const void *x, const float *bias, const void *gy, void *gx, float *dbias, float *partials, int64_t rows, int d, float p_drop,
                 uint64_t seed, uint64_t rng_offset, int dtype, void *stream) {
    CPM_REQUIRE(x && gy && gx, CPM_ERR_NULL, "gelu_bwd: NULL pointer");
    CPM_REQUIRE(rows >= 0 && d > 0 && d % 16 == 0, CPM_ERR_BAD_SHAPE, "gelu_bwd: d=%d must be a multiple of 16", d);
    CPM_REQUIRE(aligned16(x) && aligned16(gy) && aligned16(gx) && (!bias || aligned16(bias)), CPM_ERR_BAD_ALIGN, "gelu_bwd: alignment");
    if (rows == 0) return CPM_OK;
    if (dbias) {                                    // fused bias gradient: per-thread column sums -> partial rows -> one reduction
        CPM_REQUIRE(partials && 4096 % d == 0, CPM_ERR_BAD_SHAPE, "gelu_bwd: the fused bias gradient needs d | 4096 (d=%d) and a partials buffer", d);
        DISPATCH_DTYPE(dtype, gelu_kernel<T, true, true><<<GELU_BWD_BLOCKS, 256, 0, (cudaStream_t)stream>>>(
                                  (const T *)x, bias, (const T *)gy, (T *)gx, rows * (d / 16), d, dropout_threshold8(p_drop), dropout_scale8(p_drop),
                                  seed, rng_offset, partials, g_rng_base));
        reduce_partials_kernel<true><<<(d + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partials, cpm_gelu_bwd_partials_rows(d), d, d, dbias, nullptr, nullptr);
        return check_launch("gelu_bwd");
    }
    DISPATCH_DTYPE(dtype, gelu_kernel<T, true><<<grid_for(rows * (d / 16), 256), 256, 0, (cudaStream_t)stream>>>(
                              (const T *)x, bias, (const T *)gy, (T *)gx, rows * (d / 16), d, dropout_threshold8(p_drop), dropout_scale8(p_drop), seed,
                              rng_offset, nullptr, g_rng_base));
    return check_launch("gelu_bwd");
}

}  // extern "C"
