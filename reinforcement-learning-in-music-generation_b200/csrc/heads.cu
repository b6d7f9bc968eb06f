// Per-attribute head kernels over the concatenated CP logits: decode (greedy / temperature /
// nucleus with Philox inverse-CDF draws), log-prob + entropy (fwd/bwd) and masked cross-entropy
// (fwd/bwd).  One warp owns one (row, attribute) segment (decode, generic layouts) or one whole row of concatenated
// logits (log-prob / cross-entropy when the row is <= 1024 wide and 16-byte aligned); segments are <= 1024 wide.
#include "cpm_common.cuh"
#include "heads_dev.cuh"

namespace cpm {
namespace {

constexpr int MAX_SEG = 1024;
constexpr int WARPS = 4;

// loads the segment into buf (fp32), returns first-index argmax and logsumexp (T=1)
template <typename T>
__device__ __forceinline__ void load_segment(const T *__restrict__ row, int w, float *buf, int lane, ArgMax &am, float &lse) {
    for (int i = lane; i < w; i += 32) buf[i] = to_f(row[i]);
    segment_stats(buf, w, lane, am, lse);          // each lane reads back only what it wrote
}

// ---------------------------------------------------------------- C3 decode
template <typename T>
__global__ void __launch_bounds__(WARPS * 32) heads_sample_kernel(const T *__restrict__ logits, int64_t rows, int64_t ld, SegParams sp, int mode,
                                                                  uint64_t seed, int64_t seq_base, int step, const int32_t *__restrict__ step_dev,
                                                                  int64_t *__restrict__ tokens, float *__restrict__ logp, float *__restrict__ entropy) {
    __shared__ float sbuf[WARPS][MAX_SEG];
    griddep_launch();
    griddep_wait();                                     // chain kernel: the logits are the heads GEMM's output
    __shared__ float sprob[WARPS][MAX_SEG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *buf = sbuf[warp], *pr = sprob[warp];
    const int64_t items = rows * sp.n_attr;
    const int cur_step = step_dev ? step_dev[0] : step;
    for (int64_t it = (int64_t)blockIdx.x * WARPS + warp; it < items; it += (int64_t)gridDim.x * WARPS) {
        const int64_t r = it / sp.n_attr;
        const int a = (int)(it % sp.n_attr);
        const int w = sp.seg[a + 1] - sp.seg[a];
        ArgMax am;
        float lse;
        __syncwarp();
        load_segment(logits + r * ld + sp.seg[a], w, buf, lane, am, lse);
        __syncwarp();
        int tok = am.i;
        if (mode == 1)
            tok = sample_segment<MAX_SEG / 32>(buf, pr, w, lane, am, sp.temperature[a], sp.top_p[a], seed, (uint64_t)(seq_base + r), cur_step, a);
        if (lane == 0) {
            tokens[it] = tok;
            if (logp) logp[it] = buf[tok] - lse;
        }
        if (entropy) {
            float hx = 0.f;
            for (int i = lane; i < w; i += 32) { float lp = buf[i] - lse; hx -= __expf(lp) * lp; }
            hx = warp_sum(hx);
            if (lane == 0) entropy[it] = hx;
        }
    }
}

// ---------------------------------------------------------------- log-prob / entropy of given tokens
template <typename T>
__global__ void __launch_bounds__(WARPS * 32) heads_logp_kernel(const T *__restrict__ logits, int64_t rows, int64_t ld, SegParams sp,
                                                                const int64_t *__restrict__ tokens, float *__restrict__ logp,
                                                                float *__restrict__ entropy) {
    __shared__ float sbuf[WARPS][MAX_SEG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *buf = sbuf[warp];
    const int64_t items = rows * sp.n_attr;
    for (int64_t it = (int64_t)blockIdx.x * WARPS + warp; it < items; it += (int64_t)gridDim.x * WARPS) {
        const int64_t r = it / sp.n_attr;
        const int a = (int)(it % sp.n_attr);
        const int w = sp.seg[a + 1] - sp.seg[a];
        ArgMax am;
        float lse;
        __syncwarp();
        load_segment(logits + r * ld + sp.seg[a], w, buf, lane, am, lse);
        __syncwarp();
        int64_t tok = tokens[it];
        tok = tok < 0 ? 0 : (tok >= w ? w - 1 : tok);
        if (lane == 0 && logp) logp[it] = buf[tok] - lse;
        if (entropy) {
            float hx = 0.f;
            for (int i = lane; i < w; i += 32) { float lp = buf[i] - lse; hx -= __expf(lp) * lp; }
            hx = warp_sum(hx);
            if (lane == 0) entropy[it] = hx;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32) heads_logp_bwd_kernel(const T *__restrict__ logits, int64_t rows, int64_t ld, SegParams sp,
                                                                    const int64_t *__restrict__ tokens, const float *__restrict__ glogp,
                                                                    const float *__restrict__ gent, T *__restrict__ dlogits) {
    __shared__ float sbuf[WARPS][MAX_SEG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *buf = sbuf[warp];
    const int64_t items = rows * sp.n_attr;
    for (int64_t it = (int64_t)blockIdx.x * WARPS + warp; it < items; it += (int64_t)gridDim.x * WARPS) {
        const int64_t r = it / sp.n_attr;
        const int a = (int)(it % sp.n_attr);
        const int w = sp.seg[a + 1] - sp.seg[a];
        ArgMax am;
        float lse;
        __syncwarp();
        load_segment(logits + r * ld + sp.seg[a], w, buf, lane, am, lse);
        __syncwarp();
        int64_t tok = tokens[it];
        tok = tok < 0 ? 0 : (tok >= w ? w - 1 : tok);
        const float gl = glogp ? glogp[it] : 0.f;
        const float ge = gent ? gent[it] : 0.f;
        float hx = 0.f;
        if (gent) {
            for (int i = lane; i < w; i += 32) { float lp = buf[i] - lse; hx -= __expf(lp) * lp; }
            hx = warp_sum(hx);
        }
        T *dst = dlogits + r * ld + sp.seg[a];
        for (int i = lane; i < w; i += 32) {
            const float lp = buf[i] - lse, p = __expf(lp);
            float g = gl * ((i == tok ? 1.f : 0.f) - p) - ge * p * (lp + hx);
            dst[i] = from_f<T>(g);
        }
        if (a == sp.n_attr - 1)
            for (int64_t i = sp.seg[sp.n_attr] + lane; i < ld; i += 32) dlogits[r * ld + i] = from_f<T>(0.f);
    }
}

// ---------------------------------------------------------------- C4 masked cross-entropy
template <typename T>
__global__ void __launch_bounds__(WARPS * 32) masked_ce_fwd_kernel(const T *__restrict__ logits, int64_t T_, int64_t ld, SegParams sp,
                                                                   const int64_t *__restrict__ targets, const float *__restrict__ mask,
                                                                   float *__restrict__ loss_num, float *__restrict__ mask_sum,
                                                                   float *__restrict__ lse_out) {
    __shared__ float sbuf[WARPS][MAX_SEG];
    __shared__ float sacc[CPM_MAX_ATTR + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x <= CPM_MAX_ATTR) sacc[threadIdx.x] = 0.f;
    __syncthreads();
    float *buf = sbuf[warp];
    const int64_t items = T_ * sp.n_attr;
    for (int64_t it = (int64_t)blockIdx.x * WARPS + warp; it < items; it += (int64_t)gridDim.x * WARPS) {
        const int64_t t = it / sp.n_attr;
        const int a = (int)(it % sp.n_attr);
        const int w = sp.seg[a + 1] - sp.seg[a];
        ArgMax am;
        float lse;
        __syncwarp();
        load_segment(logits + t * ld + sp.seg[a], w, buf, lane, am, lse);
        __syncwarp();
        if (lane == 0) {
            int64_t tg = targets[it];
            tg = tg < 0 ? 0 : (tg >= w ? w - 1 : tg);
            const float m = mask[t];
            if (lse_out) lse_out[it] = lse;
            atomicAdd(&sacc[a], m * (lse - buf[tg]));
            if (a == 0) atomicAdd(&sacc[CPM_MAX_ATTR], m);
        }
    }
    __syncthreads();
    if (threadIdx.x < sp.n_attr) atomicAdd(&loss_num[threadIdx.x], sacc[threadIdx.x]);
    if (threadIdx.x == CPM_MAX_ATTR && mask_sum) atomicAdd(mask_sum, sacc[CPM_MAX_ATTR]);
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32) masked_ce_bwd_kernel(const T *__restrict__ logits, int64_t T_, int64_t ld, SegParams sp,
                                                                   const int64_t *__restrict__ targets, const float *__restrict__ mask,
                                                                   const float *__restrict__ lse_in, const float *__restrict__ gscale,
                                                                   const float *__restrict__ denom, T *__restrict__ dlogits) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t items = T_ * sp.n_attr;
    const float inv_denom = 1.f / denom[0];
    for (int64_t it = (int64_t)blockIdx.x * WARPS + warp; it < items; it += (int64_t)gridDim.x * WARPS) {
        const int64_t t = it / sp.n_attr;
        const int a = (int)(it % sp.n_attr);
        const int w = sp.seg[a + 1] - sp.seg[a];
        int64_t tg = targets[it];
        tg = tg < 0 ? 0 : (tg >= w ? w - 1 : tg);
        const float coef = gscale[a] * mask[t] * inv_denom;
        const float lse = lse_in[it];
        const T *src = logits + t * ld + sp.seg[a];
        T *dst = dlogits + t * ld + sp.seg[a];
        for (int i = lane; i < w; i += 32) {
            float g = coef == 0.f ? 0.f : coef * (__expf(to_f(src[i]) - lse) - (i == tg ? 1.f : 0.f));
            dst[i] = from_f<T>(g);
        }
        if (a == sp.n_attr - 1)
            for (int64_t i = sp.seg[sp.n_attr] + lane; i < ld; i += 32) dlogits[t * ld + i] = from_f<T>(0.f);
    }
}

// ---------------------------------------------------------------- row-per-warp variants (rows <= MAX_SEG wide, ld % 8 == 0)
// The (row, attribute)-per-warp kernels above keep one 36-270 byte segment in flight per warp and pay a memory round
// trip per segment; at 65 536 rows they are latency bound (4-14 % of HBM).  Here a warp fetches the whole row of concatenated
// logits with 16-byte loads (one round trip), works on all attribute segments out of shared memory, and writes gradient
// rows back with 16-byte stores.
template <typename T>
__device__ __forceinline__ void load_row(const T *__restrict__ row, int width, float *buf, int lane) {
    const int G = width >> 3;
    for (int g = lane; g < G; g += 32) {
        Vec8<T> v;
        v.load(row + g * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) buf[g * 8 + j] = v.v[j];
    }
    for (int i = G * 8 + lane; i < width; i += 32) buf[i] = to_f(row[i]);
}
template <typename T>
__device__ __forceinline__ void store_row(T *__restrict__ row, int ld, const float *buf, int width, int lane) {   // zero-pads [width, ld)
    for (int g = lane; g < (ld >> 3); g += 32) {
        Vec8<T> v;
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] = (g * 8 + j < width) ? buf[g * 8 + j] : 0.f;
        v.store(row + g * 8);
    }
}
__device__ __forceinline__ void seg_stats(const float *b, int w, int lane, ArgMax &am, float &lse) {
    ArgMax a{-INFINITY, 0x7fffffff};
    for (int i = lane; i < w; i += 32) { const float x = b[i]; if (x > a.v) { a.v = x; a.i = i; } }
    am = warp_argmax(a);
    float s = 0.f;
    for (int i = lane; i < w; i += 32) s += __expf(b[i] - am.v);
    lse = am.v + __logf(warp_sum(s));
}
__device__ __forceinline__ int clamp_tok(int64_t t, int w) { return t < 0 ? 0 : (t >= w ? w - 1 : (int)t); }

template <typename T>
__global__ void __launch_bounds__(WARPS * 32) heads_logp_row_kernel(const T *__restrict__ logits, int64_t rows, int64_t ld, SegParams sp,
                                                                    const int64_t *__restrict__ tokens, float *__restrict__ logp,
                                                                    float *__restrict__ entropy) {
    __shared__ __align__(16) float sbuf[WARPS][MAX_SEG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, A = sp.n_attr, width = sp.seg[A];
    float *buf = sbuf[warp];
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < rows; r += (int64_t)gridDim.x * WARPS) {
        const int64_t mytok = lane < A ? tokens[r * A + lane] : 0;
        __syncwarp();
        load_row(logits + r * ld, width, buf, lane);
        __syncwarp();
        for (int a = 0; a < A; ++a) {
            const float *b = buf + sp.seg[a];
            const int w = sp.seg[a + 1] - sp.seg[a];
            ArgMax am;
            float lse;
            seg_stats(b, w, lane, am, lse);
            const int tok = clamp_tok(__shfl_sync(0xffffffffu, mytok, a), w);
            if (lane == 0 && logp) logp[r * A + a] = b[tok] - lse;
            if (entropy) {
                float hx = 0.f;
                for (int i = lane; i < w; i += 32) { const float lp = b[i] - lse; hx -= __expf(lp) * lp; }
                hx = warp_sum(hx);
                if (lane == 0) entropy[r * A + a] = hx;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32) heads_logp_bwd_row_kernel(const T *__restrict__ logits, int64_t rows, int64_t ld, SegParams sp,
                                                                        const int64_t *__restrict__ tokens, const float *__restrict__ glogp,
                                                                        const float *__restrict__ gent, T *__restrict__ dlogits) {
    __shared__ __align__(16) float sbuf[WARPS][MAX_SEG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, A = sp.n_attr, width = sp.seg[A];
    float *buf = sbuf[warp];
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < rows; r += (int64_t)gridDim.x * WARPS) {
        const int64_t mytok = lane < A ? tokens[r * A + lane] : 0;
        const float mygl = (glogp && lane < A) ? glogp[r * A + lane] : 0.f;
        const float myge = (gent && lane < A) ? gent[r * A + lane] : 0.f;
        __syncwarp();
        load_row(logits + r * ld, width, buf, lane);
        __syncwarp();
        for (int a = 0; a < A; ++a) {
            float *b = buf + sp.seg[a];
            const int w = sp.seg[a + 1] - sp.seg[a];
            ArgMax am;
            float lse;
            seg_stats(b, w, lane, am, lse);
            const int tok = clamp_tok(__shfl_sync(0xffffffffu, mytok, a), w);
            const float gl = __shfl_sync(0xffffffffu, mygl, a), ge = __shfl_sync(0xffffffffu, myge, a);
            float hx = 0.f;
            if (gent) {
                for (int i = lane; i < w; i += 32) { const float lp = b[i] - lse; hx -= __expf(lp) * lp; }
                hx = warp_sum(hx);
            }
            for (int i = lane; i < w; i += 32) {
                const float lp = b[i] - lse, p = __expf(lp);
                b[i] = gl * ((i == tok ? 1.f : 0.f) - p) - ge * p * (lp + hx);
            }
        }
        __syncwarp();
        store_row(dlogits + r * ld, (int)ld, buf, width, lane);
    }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32) masked_ce_fwd_row_kernel(const T *__restrict__ logits, int64_t T_, int64_t ld, SegParams sp,
                                                                       const int64_t *__restrict__ targets, const float *__restrict__ mask,
                                                                       float *__restrict__ loss_num, float *__restrict__ mask_sum,
                                                                       float *__restrict__ lse_out) {
    __shared__ __align__(16) float sbuf[WARPS][MAX_SEG];
    __shared__ float sacc[CPM_MAX_ATTR + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, A = sp.n_attr, width = sp.seg[A];
    if (threadIdx.x <= CPM_MAX_ATTR) sacc[threadIdx.x] = 0.f;
    __syncthreads();
    float *buf = sbuf[warp];
    float acc = 0.f, macc = 0.f;                       // lane a < A accumulates attribute a; lane 0 the mask sum
    for (int64_t t = (int64_t)blockIdx.x * WARPS + warp; t < T_; t += (int64_t)gridDim.x * WARPS) {
        const int64_t mytg = lane < A ? targets[t * A + lane] : 0;
        const float m = mask[t];
        __syncwarp();
        load_row(logits + t * ld, width, buf, lane);
        __syncwarp();
        for (int a = 0; a < A; ++a) {
            const float *b = buf + sp.seg[a];
            const int w = sp.seg[a + 1] - sp.seg[a];
            ArgMax am;
            float lse;
            seg_stats(b, w, lane, am, lse);
            const int tg = clamp_tok(__shfl_sync(0xffffffffu, mytg, a), w);
            if (lane == a) {
                if (lse_out) lse_out[t * A + a] = lse;
                acc += m * (lse - b[tg]);
            }
        }
        if (lane == 0) macc += m;
    }
    if (lane < A) atomicAdd(&sacc[lane], acc);
    if (lane == 0) atomicAdd(&sacc[CPM_MAX_ATTR], macc);
    __syncthreads();
    if (threadIdx.x < A) atomicAdd(&loss_num[threadIdx.x], sacc[threadIdx.x]);
    if (threadIdx.x == CPM_MAX_ATTR && mask_sum) atomicAdd(mask_sum, sacc[CPM_MAX_ATTR]);
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32) masked_ce_bwd_row_kernel(const T *__restrict__ logits, int64_t T_, int64_t ld, SegParams sp,
                                                                       const int64_t *__restrict__ targets, const float *__restrict__ mask,
                                                                       const float *__restrict__ lse_in, const float *__restrict__ gscale,
                                                                       const float *__restrict__ denom, T *__restrict__ dlogits) {
    __shared__ __align__(16) float sbuf[WARPS][MAX_SEG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, A = sp.n_attr, width = sp.seg[A];
    float *buf = sbuf[warp];
    const float inv_denom = 1.f / denom[0];
    const float mygs = lane < A ? gscale[lane] : 0.f;
    for (int64_t t = (int64_t)blockIdx.x * WARPS + warp; t < T_; t += (int64_t)gridDim.x * WARPS) {
        const int64_t mytg = lane < A ? targets[t * A + lane] : 0;
        const float mylse = lane < A ? lse_in[t * A + lane] : 0.f;
        const float m = mask[t];
        __syncwarp();
        load_row(logits + t * ld, width, buf, lane);
        __syncwarp();
        for (int a = 0; a < A; ++a) {
            float *b = buf + sp.seg[a];
            const int w = sp.seg[a + 1] - sp.seg[a];
            const int tg = clamp_tok(__shfl_sync(0xffffffffu, mytg, a), w);
            const float lse = __shfl_sync(0xffffffffu, mylse, a);
            const float coef = __shfl_sync(0xffffffffu, mygs, a) * m * inv_denom;
            for (int i = lane; i < w; i += 32) b[i] = coef == 0.f ? 0.f : coef * (__expf(b[i] - lse) - (i == tg ? 1.f : 0.f));
        }
        __syncwarp();
        store_row(dlogits + t * ld, (int)ld, buf, width, lane);
    }
}

// ---------------------------------------------------------------- one THREAD per row (ld % 8 == 0, 16-byte aligned rows)
// The warp-per-row kernels above spend ~2000 warp instructions per row and an eight-lanes-per-row variant still 440: narrow
// segments leave most lanes of a cross-lane reduction idle; staging 32 rows per warp in shared memory removes the reductions
// but leaves 8 warps per SM (profiles/r02_summary.md, section I).  Here every THREAD streams its own row from global memory in
// 16-byte pieces (the 128-byte lines it touches stay in L1 for the 8 consecutive reads that consume them) and keeps ONE open
// segment's running (max, sum, sum x) in registers with the online-softmax update - the segment layout is the same for every
// row, so all 32 threads take the same branches at the same column: no shuffles, no divergence, no shared memory, full
// occupancy.  About 90 warp instructions per row.
template <typename V> __device__ __forceinline__ V tpr_pick(const V (&v)[CPM_MAX_ATTR], int a) {   // v[a], a warp-uniform
    V r = v[0];
#pragma unroll
    for (int j = 1; j < CPM_MAX_ATTR; ++j) r = (a == j) ? v[j] : r;
    return r;
}
template <typename V> __device__ __forceinline__ void tpr_put(V (&v)[CPM_MAX_ATTR], int a, V x) {
#pragma unroll
    for (int j = 0; j < CPM_MAX_ATTR; ++j) v[j] = (a == j) ? x : v[j];
}

template <typename T>
__global__ void __launch_bounds__(128) heads_logp_tpr_kernel(const T *__restrict__ logits, int64_t rows, int64_t ld, SegParams sp,
                                                             const int64_t *__restrict__ tokens, float *__restrict__ logp,
                                                             float *__restrict__ entropy) {
    const int A = sp.n_attr;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const T *row = logits + r * ld;
    auto emit = [&](int a, float m, float s, float sx) {
        const float lse = m + __logf(s);
        if (logp) {
            const int tok = clamp_tok(tokens[r * A + a], sp.seg[a + 1] - sp.seg[a]);
            logp[r * A + a] = to_f(row[sp.seg[a] + tok]) - lse;
        }
        if (entropy) entropy[r * A + a] = lse - sx / s;
    };
    if (entropy) tpr_row_stats<T, 2>(row, (int)ld, sp, A, emit); else tpr_row_stats<T, 1>(row, (int)ld, sp, A, emit);
}

// gradient chunks: out[k] = f(x[k], column) for k in [k0, k1) of the open segment
template <typename F> __device__ __forceinline__ void tpr_apply(const float (&x)[8], float (&out)[8], int c0, int k0, int k1, F f) {
    if (k0 == 0 && k1 == 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) out[k] = f(x[k], c0 + k);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k >= k0 && k < k1) out[k] = f(x[k], c0 + k);
    }
}
// second walk: writes all ld columns of the gradient row (zeros behind the last segment); open(a) loads segment a's parameters
template <typename T, typename Open, typename F>
__device__ __forceinline__ void tpr_row_grad(const T *__restrict__ row, T *__restrict__ drow, int ld, const SegParams &sp, int A, Open open, F f) {
    const int width = sp.seg[A];
    int a = 0, hi = sp.seg[1];
    open(0);
    RowStream<T> rs(row, ld);
    constexpr int CPS = RowStream<T>::CPS;
    for (int span = 0; span * CPS * 8 < ld; ++span) {
        rs.advance(span);
        auto body = [&](int c0, const float (&x)[8]) {
            float out[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] = 0.f;
            int k0 = 0;
            while (k0 < 8 && a < A && c0 < width) {
                const int k1 = min(8, hi - c0);
                tpr_apply(x, out, c0, k0, k1, f);
                if (hi <= c0 + 8) {
                    ++a;
                    hi = a < A ? sp.seg[a + 1] : 0x7fffffff;
                    if (a < A) open(a);
                    k0 = k1;
                } else {
                    k0 = 8;
                }
            }
            tpr_store8(drow + c0, out);
        };
        float x[8];
#define CPM_TPR_CHUNK(J)                                                   \
        if (J < CPS && (span * CPS + J) * 8 < ld) {                        \
            rs.template chunk<(J < CPS ? J : 0)>(x);                       \
            body((span * CPS + J) * 8, x);                                 \
        }
        CPM_TPR_CHUNK(0) CPM_TPR_CHUNK(1) CPM_TPR_CHUNK(2) CPM_TPR_CHUNK(3)
        CPM_TPR_CHUNK(4) CPM_TPR_CHUNK(5) CPM_TPR_CHUNK(6) CPM_TPR_CHUNK(7)
#undef CPM_TPR_CHUNK
    }
}

template <typename T>
__global__ void __launch_bounds__(128) heads_logp_bwd_tpr_kernel(const T *__restrict__ logits, int64_t rows, int64_t ld, SegParams sp,
                                                                 const int64_t *__restrict__ tokens, const float *__restrict__ glogp,
                                                                 const float *__restrict__ gent, T *__restrict__ dlogits) {
    const int A = sp.n_attr;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const T *row = logits + r * ld;
    float lse[CPM_MAX_ATTR], hx[CPM_MAX_ATTR];
#pragma unroll
    for (int j = 0; j < CPM_MAX_ATTR; ++j) { lse[j] = 0.f; hx[j] = 0.f; }
    auto emit = [&](int a, float m, float s, float sx) {
        const float l = m + __logf(s);
        tpr_put(lse, a, l);
        tpr_put(hx, a, l - sx / s);
    };
    if (gent) tpr_row_stats<T, 2>(row, (int)ld, sp, A, emit); else tpr_row_stats<T, 1>(row, (int)ld, sp, A, emit);
    float c_lse = 0.f, c_hx = 0.f, c_gl = 0.f, c_ge = 0.f;
    int c_tok = 0;
    auto open = [&](int a) {
        c_lse = tpr_pick(lse, a);
        c_hx = tpr_pick(hx, a);
        c_gl = glogp ? glogp[r * A + a] : 0.f;
        c_ge = gent ? gent[r * A + a] : 0.f;
        c_tok = sp.seg[a] + clamp_tok(tokens[r * A + a], sp.seg[a + 1] - sp.seg[a]);
    };
    tpr_row_grad(row, dlogits + r * ld, (int)ld, sp, A, open, [&](float x, int i) {
        const float lp = x - c_lse, p = tpr_ex2(lp * TPR_LOG2E);
        return c_gl * ((i == c_tok ? 1.f : 0.f) - p) - c_ge * p * (lp + c_hx);
    });
}

template <typename T>
__global__ void __launch_bounds__(128) masked_ce_fwd_tpr_kernel(const T *__restrict__ logits, int64_t T_, int64_t ld, SegParams sp,
                                                                const int64_t *__restrict__ targets, const float *__restrict__ mask,
                                                                float *__restrict__ loss_num, float *__restrict__ mask_sum,
                                                                float *__restrict__ lse_out) {
    __shared__ float sacc[CPM_MAX_ATTR + 1];
    const int A = sp.n_attr, lane = threadIdx.x & 31;
    if (threadIdx.x <= CPM_MAX_ATTR) sacc[threadIdx.x] = 0.f;
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < T_;
    const int64_t tt = live ? t : T_ - 1;                  // dead threads walk the last row (uniform control flow) and add nothing
    const T *row = logits + tt * ld;
    const float m_ = live ? mask[tt] : 0.f;
    float acc[CPM_MAX_ATTR];
#pragma unroll
    for (int j = 0; j < CPM_MAX_ATTR; ++j) acc[j] = 0.f;
    tpr_row_stats<T, 1>(row, (int)ld, sp, A, [&](int a, float m, float s, float) {
        const float lse = m + __logf(s);
        if (live && lse_out) lse_out[tt * A + a] = lse;
        const int tg = clamp_tok(targets[tt * A + a], sp.seg[a + 1] - sp.seg[a]);
        tpr_put(acc, a, m_ * (lse - to_f(row[sp.seg[a] + tg])));
    });
#pragma unroll
    for (int j = 0; j < CPM_MAX_ATTR; ++j) {
        if (j < A) {
            const float v = warp_sum(acc[j]);
            if (lane == 0) atomicAdd(&sacc[j], v);
        }
    }
    const float mm = warp_sum(m_);
    if (lane == 0) atomicAdd(&sacc[CPM_MAX_ATTR], mm);
    __syncthreads();
    if (threadIdx.x < A) atomicAdd(&loss_num[threadIdx.x], sacc[threadIdx.x]);
    if (threadIdx.x == CPM_MAX_ATTR && mask_sum) atomicAdd(mask_sum, sacc[CPM_MAX_ATTR]);
}

template <typename T>
__global__ void __launch_bounds__(128) masked_ce_bwd_tpr_kernel(const T *__restrict__ logits, int64_t T_, int64_t ld, SegParams sp,
                                                                const int64_t *__restrict__ targets, const float *__restrict__ mask,
                                                                const float *__restrict__ lse_in, const float *__restrict__ gscale,
                                                                const float *__restrict__ denom, T *__restrict__ dlogits) {
    const int A = sp.n_attr;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T_) return;
    const T *row = logits + t * ld;
    const float mk = mask[t] * (1.f / denom[0]);
    float c_lse = 0.f, c_coef = 0.f;
    int c_tg = 0;
    auto open = [&](int a) {
        c_lse = lse_in[t * A + a];
        c_coef = gscale[a] * mk;
        c_tg = sp.seg[a] + clamp_tok(targets[t * A + a], sp.seg[a + 1] - sp.seg[a]);
    };
    tpr_row_grad(row, dlogits + t * ld, (int)ld, sp, A, open, [&](float x, int i) {
        return c_coef == 0.f ? 0.f : c_coef * (tpr_ex2((x - c_lse) * TPR_LOG2E) - (i == c_tg ? 1.f : 0.f));
    });
}

inline bool tpr_path_ok(const SegParams &sp, int64_t ld, const void *p0, const void *p1) {
    return sp.seg[0] == 0 && ld % 8 == 0 && aligned16(p0) && (!p1 || aligned16(p1));
}
inline int tpr_grid(int64_t rows) { return (int)((rows + 127) / 128); }

inline bool row_path_ok(const SegParams &sp, int64_t ld, const void *p0, const void *p1) {
    return sp.seg[0] == 0 && sp.seg[sp.n_attr] <= MAX_SEG && ld % 8 == 0 && ld <= MAX_SEG && aligned16(p0) && (!p1 || aligned16(p1));
}

int fill_seg(SegParams &sp, const int *seg, int n_attr, int64_t ld, const float *temperature, const float *top_p) {
    CPM_REQUIRE(seg, CPM_ERR_NULL, "heads: seg_host is NULL");
    CPM_REQUIRE(n_attr >= 1 && n_attr <= CPM_MAX_ATTR, CPM_ERR_BAD_SHAPE, "heads: n_attr=%d out of [1,%d]", n_attr, CPM_MAX_ATTR);
    sp.n_attr = n_attr;
    for (int a = 0; a <= n_attr; ++a) sp.seg[a] = seg[a];
    for (int a = 0; a < n_attr; ++a) {
        const int w = seg[a + 1] - seg[a];
        CPM_REQUIRE(w >= 1 && w <= MAX_SEG, CPM_ERR_BAD_SHAPE, "heads: segment %d width %d out of [1,%d]", a, w, MAX_SEG);
        sp.temperature[a] = temperature ? temperature[a] : 1.f;
        sp.top_p[a] = top_p ? top_p[a] : 0.f;
        CPM_REQUIRE(sp.temperature[a] > 0.f, CPM_ERR_BAD_SHAPE, "heads: temperature[%d]=%f must be > 0", a, sp.temperature[a]);
    }
    CPM_REQUIRE(seg[0] >= 0 && (int64_t)seg[n_attr] <= ld, CPM_ERR_BAD_SHAPE, "heads: segments [%d,%d) exceed row stride %lld", seg[0], seg[n_attr], (long long)ld);
    return CPM_OK;
}

inline int warp_grid(int64_t items) {
    int64_t b = (items + WARPS - 1) / WARPS;
    int64_t cap = (int64_t)num_sms() * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace cpm

using namespace cpm;

#define DISPATCH_DTYPE(dtype, ...)                                              \
    if ((dtype) == CPM_F32) { using T = float; __VA_ARGS__; }                   \
    else if ((dtype) == CPM_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }     \
    else return fail(CPM_ERR_BAD_DTYPE, "unsupported dtype %d", (int)(dtype));

extern "C" {

int cpm_heads_sample(const void *logits, int64_t rows, int64_t ld_logits, const int *seg_host, int n_attr,
                     const float *temperature_host, const float *top_p_host, int mode, uint64_t seed, int64_t seq_base, int step,
                     const int32_t *step_dev, int64_t *tokens, float *logp, float *entropy, int dtype, void *stream) {
    CPM_REQUIRE(logits && tokens, CPM_ERR_NULL, "heads_sample: NULL pointer");
    CPM_REQUIRE(mode == 0 || mode == 1, CPM_ERR_BAD_SHAPE, "heads_sample: mode %d", mode);
    SegParams sp{};
    int rc = fill_seg(sp, seg_host, n_attr, ld_logits, temperature_host, top_p_host);
    if (rc) return rc;
    if (rows <= 0) return rows == 0 ? CPM_OK : fail(CPM_ERR_BAD_SHAPE, "heads_sample: rows=%lld", (long long)rows);
    DISPATCH_DTYPE(dtype, launch_chain(heads_sample_kernel<T>, dim3(warp_grid(rows * n_attr)), dim3(WARPS * 32), 0, (cudaStream_t)stream,
                                       (const T *)logits, rows, ld_logits, sp, mode, seed, seq_base, step, step_dev, tokens, logp, entropy));
    return check_launch("heads_sample");
}

int cpm_heads_logp(const void *logits, int64_t rows, int64_t ld_logits, const int *seg_host, int n_attr, const int64_t *tokens,
                   float *logp, float *entropy, int dtype, void *stream) {
    CPM_REQUIRE(logits && tokens && (logp || entropy), CPM_ERR_NULL, "heads_logp: NULL pointer");
    SegParams sp{};
    int rc = fill_seg(sp, seg_host, n_attr, ld_logits, nullptr, nullptr);
    if (rc) return rc;
    if (rows <= 0) return rows == 0 ? CPM_OK : fail(CPM_ERR_BAD_SHAPE, "heads_logp: rows=%lld", (long long)rows);
    if (tpr_path_ok(sp, ld_logits, logits, nullptr)) {
        DISPATCH_DTYPE(dtype, heads_logp_tpr_kernel<T><<<tpr_grid(rows), 128, 0, (cudaStream_t)stream>>>((const T *)logits, rows, ld_logits, sp, tokens, logp, entropy));
        return check_launch("heads_logp");
    }
    if (row_path_ok(sp, ld_logits, logits, nullptr)) {
        DISPATCH_DTYPE(dtype, heads_logp_row_kernel<T><<<warp_grid(rows), WARPS * 32, 0, (cudaStream_t)stream>>>(
                                  (const T *)logits, rows, ld_logits, sp, tokens, logp, entropy));
        return check_launch("heads_logp");
    }
    DISPATCH_DTYPE(dtype, heads_logp_kernel<T><<<warp_grid(rows * n_attr), WARPS * 32, 0, (cudaStream_t)stream>>>(
                              (const T *)logits, rows, ld_logits, sp, tokens, logp, entropy));
    return check_launch("heads_logp");
}

int cpm_heads_logp_bwd(const void *logits, int64_t rows, int64_t ld_logits, const int *seg_host, int n_attr, const int64_t *tokens,
                       const float *glogp, const float *gentropy, void *dlogits, int dtype, void *stream) {
    CPM_REQUIRE(logits && tokens && dlogits, CPM_ERR_NULL, "heads_logp_bwd: NULL pointer");
    SegParams sp{};
    int rc = fill_seg(sp, seg_host, n_attr, ld_logits, nullptr, nullptr);
    if (rc) return rc;
    if (rows <= 0) return rows == 0 ? CPM_OK : fail(CPM_ERR_BAD_SHAPE, "heads_logp_bwd: rows=%lld", (long long)rows);
    if (tpr_path_ok(sp, ld_logits, logits, dlogits)) {
        DISPATCH_DTYPE(dtype, heads_logp_bwd_tpr_kernel<T><<<tpr_grid(rows), 128, 0, (cudaStream_t)stream>>>((const T *)logits, rows, ld_logits, sp, tokens, glogp, gentropy, (T *)dlogits));
        return check_launch("heads_logp_bwd");
    }
    if (row_path_ok(sp, ld_logits, logits, dlogits)) {
        DISPATCH_DTYPE(dtype, heads_logp_bwd_row_kernel<T><<<warp_grid(rows), WARPS * 32, 0, (cudaStream_t)stream>>>(
                                  (const T *)logits, rows, ld_logits, sp, tokens, glogp, gentropy, (T *)dlogits));
        return check_launch("heads_logp_bwd");
    }
    DISPATCH_DTYPE(dtype, heads_logp_bwd_kernel<T><<<warp_grid(rows * n_attr), WARPS * 32, 0, (cudaStream_t)stream>>>(
                              (const T *)logits, rows, ld_logits, sp, tokens, glogp, gentropy, (T *)dlogits));
    return check_launch("heads_logp_bwd");
}

int cpm_masked_ce_fwd(const void *logits, int64_t T_, int64_t ld_logits, const int *seg_host, int n_attr, const int64_t *targets,
                      const float *mask, float *loss_num, float *mask_sum, float *lse, int dtype, void *stream) {
    CPM_REQUIRE(logits && targets && mask && loss_num, CPM_ERR_NULL, "masked_ce_fwd: NULL pointer");
    SegParams sp{};
    int rc = fill_seg(sp, seg_host, n_attr, ld_logits, nullptr, nullptr);
    if (rc) return rc;
    if (T_ <= 0) return T_ == 0 ? CPM_OK : fail(CPM_ERR_BAD_SHAPE, "masked_ce_fwd: T=%lld", (long long)T_);
    if (tpr_path_ok(sp, ld_logits, logits, nullptr)) {
        DISPATCH_DTYPE(dtype, masked_ce_fwd_tpr_kernel<T><<<tpr_grid(T_), 128, 0, (cudaStream_t)stream>>>((const T *)logits, T_, ld_logits, sp, targets, mask, loss_num, mask_sum, lse));
        return check_launch("masked_ce_fwd");
    }
    if (row_path_ok(sp, ld_logits, logits, nullptr)) {
        DISPATCH_DTYPE(dtype, masked_ce_fwd_row_kernel<T><<<warp_grid(T_), WARPS * 32, 0, (cudaStream_t)stream>>>(
                                  (const T *)logits, T_, ld_logits, sp, targets, mask, loss_num, mask_sum, lse));
        return check_launch("masked_ce_fwd");
    }
    DISPATCH_DTYPE(dtype, masked_ce_fwd_kernel<T><<<warp_grid(T_ * n_attr), WARPS * 32, 0, (cudaStream_t)stream>>>(
                              (const T *)logits, T_, ld_logits, sp, targets, mask, loss_num, mask_sum, lse));
    return check_launch("masked_ce_fwd");
}

int cpm_masked_ce_bwd(const void *logits, int64_t T_, int64_t ld_logits, const int *seg_host, int n_attr, const int64_t *targets,
                      const float *mask, const float *lse, const float *gscale, const float *denom, void *dlogits, int dtype,
                      void *stream) {
    CPM_REQUIRE(logits && targets && mask && lse && gscale && denom && dlogits, CPM_ERR_NULL, "masked_ce_bwd: NULL pointer");
    SegParams sp{};
    int rc = fill_seg(sp, seg_host, n_attr, ld_logits, nullptr, nullptr);
    if (rc) return rc;
    if (T_ <= 0) return T_ == 0 ? CPM_OK : fail(CPM_ERR_BAD_SHAPE, "masked_ce_bwd: T=%lld", (long long)T_);
    if (tpr_path_ok(sp, ld_logits, logits, dlogits)) {
        DISPATCH_DTYPE(dtype, masked_ce_bwd_tpr_kernel<T><<<tpr_grid(T_), 128, 0, (cudaStream_t)stream>>>((const T *)logits, T_, ld_logits, sp, targets, mask, lse, gscale, denom, (T *)dlogits));
        return check_launch("masked_ce_bwd");
    }
    if (row_path_ok(sp, ld_logits, logits, dlogits)) {
        DISPATCH_DTYPE(dtype, masked_ce_bwd_row_kernel<T><<<warp_grid(T_), WARPS * 32, 0, (cudaStream_t)stream>>>(
                                  (const T *)logits, T_, ld_logits, sp, targets, mask, lse, gscale, denom, (T *)dlogits));
        return check_launch("masked_ce_bwd");
    }
    DISPATCH_DTYPE(dtype, masked_ce_bwd_kernel<T><<<warp_grid(T_ * n_attr), WARPS * 32, 0, (cudaStream_t)stream>>>(
                              (const T *)logits, T_, ld_logits, sp, targets, mask, lse, gscale, denom, (T *)dlogits));
    return check_launch("masked_ce_bwd");
}

}  // extern "C"
