// Device pieces of the CP head decode shared by the stand-alone sampler (heads.cu) and the persistent rollout step
// (rollout_step.cu): both draw the same token from the same logits, bit for bit.
// Reference behaviour: dqn_policy/model.py:19-55 (temperature softmax, nucleus cut, one draw per attribute).
#pragma once
#include "cpm_common.cuh"

namespace cpm {

struct SegParams {
    int seg[CPM_MAX_ATTR + 1];
    float temperature[CPM_MAX_ATTR];
    float top_p[CPM_MAX_ATTR];
    int n_attr;
};

// ---- one thread per row of concatenated logits (heads.cu log-prob / cross-entropy kernels, rl.cu TD kernel): the row streamed
// from global memory in 128-byte spans, one open segment's running statistics in registers, warp-uniform control flow
constexpr float TPR_NEG = -3.0e38f;                       // "minus infinity" that stays finite under subtraction
constexpr float TPR_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float tpr_ex2(float x) {        // 2^x, flush-to-zero: one MUFU, no denormal fix-up code around it
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// This thread's row as a stream of 128-byte spans (8 x 16 bytes, all eight loads in flight together, the next span fetched
// while the current one is consumed) cut into chunks of 8 elements.
template <typename T>
struct RowStream {
    static constexpr int PPC = 8 * (int)sizeof(T) / 16;     // 16-byte pieces per 8-element chunk: 1 (bf16) or 2 (fp32)
    static constexpr int CPS = 8 / PPC;                     // chunks per span
    uint4 cur[8], nxt[8];
    const uint4 *base;
    int npieces;                                            // readable 16-byte pieces of the row (ld elements)
    __device__ __forceinline__ RowStream(const T *row, int ld) : base(reinterpret_cast<const uint4 *>(row)), npieces(ld * (int)sizeof(T) / 16) {
        fetch(0);
    }
    __device__ __forceinline__ void fetch(int span) {
#pragma unroll
        for (int p = 0; p < 8; ++p)
            if (span * 8 + p < npieces) nxt[p] = base[span * 8 + p];
    }
    __device__ __forceinline__ void advance(int span) {    // make `span` current, start fetching span + 1
#pragma unroll
        for (int p = 0; p < 8; ++p) cur[p] = nxt[p];
        fetch(span + 1);
    }
    template <int J> __device__ __forceinline__ void chunk(float (&x)[8]) const;
};
template <> template <int J> __device__ __forceinline__ void RowStream<__nv_bfloat16>::chunk(float (&x)[8]) const {
    const uint32_t w[4] = {cur[J].x, cur[J].y, cur[J].z, cur[J].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[2 * i] = __uint_as_float(w[i] << 16); x[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}
template <> template <int J> __device__ __forceinline__ void RowStream<float>::chunk(float (&x)[8]) const {
    const uint4 a = cur[2 * J], b = cur[2 * J + 1];
    x[0] = __uint_as_float(a.x); x[1] = __uint_as_float(a.y); x[2] = __uint_as_float(a.z); x[3] = __uint_as_float(a.w);
    x[4] = __uint_as_float(b.x); x[5] = __uint_as_float(b.y); x[6] = __uint_as_float(b.z); x[7] = __uint_as_float(b.w);
}
template <typename T> __device__ __forceinline__ void tpr_store8(T *p, const float (&x)[8]) {
    Vec8<T> v;
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = x[k];
    v.store(p);
}
// (m, s, sx) <- online-softmax update with x[k], k in [k0, k1) (warp-uniform bounds); s and sx are relative to exp(m)
// MODE: 0 running max only, 1 + sum of exp, 2 + sum of exp * x
template <int MODE>
__device__ __forceinline__ void tpr_accumulate(const float (&x)[8], int k0, int k1, float &m, float &s, float &sx) {
    constexpr bool ENT = MODE == 2;
    const bool full = k0 == 0 && k1 == 8;
    float cm = TPR_NEG;
    if (full) {
#pragma unroll
        for (int k = 0; k < 8; ++k) cm = fmaxf(cm, x[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) cm = (k >= k0 && k < k1) ? fmaxf(cm, x[k]) : cm;
    }
    if (MODE == 0) { m = fmaxf(m, cm); return; }
    const float mn = fmaxf(m, cm), sc = tpr_ex2((m - mn) * TPR_LOG2E), nb = -mn * TPR_LOG2E;
    float e0 = 0.f, e1 = 0.f, y0 = 0.f, y1 = 0.f;
    if (full) {
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const float a = tpr_ex2(fmaf(x[k], TPR_LOG2E, nb)), b = tpr_ex2(fmaf(x[k + 1], TPR_LOG2E, nb));
            e0 += a; e1 += b;
            if (ENT) { y0 = fmaf(a, x[k], y0); y1 = fmaf(b, x[k + 1], y1); }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k >= k0 && k < k1) {
                const float a = tpr_ex2(fmaf(x[k], TPR_LOG2E, nb));
                e0 += a;
                if (ENT) y0 = fmaf(a, x[k], y0);
            }
        }
    }
    m = mn;
    s = fmaf(s, sc, e0 + e1);
    if (ENT) sx = fmaf(sx, sc, y0 + y1);
}
// Walks one row chunk by chunk; emit(a, max, sum, sumx) is called when segment a closes (warp-uniform control flow).
template <typename T, int MODE, typename Emit>
__device__ __forceinline__ void tpr_row_stats(const T *__restrict__ row, int ld, const SegParams &sp, int A, Emit emit) {
    const int width = sp.seg[A];
    int a = 0, hi = sp.seg[1];
    float m = TPR_NEG, s = 0.f, sx = 0.f;
    RowStream<T> rs(row, ld);
    constexpr int CPS = RowStream<T>::CPS;
    for (int span = 0; span * CPS * 8 < width; ++span) {
        rs.advance(span);
        auto body = [&](int c0, const float (&x)[8]) {
            int k0 = 0;
            while (k0 < 8 && a < A) {
                const int k1 = min(8, hi - c0);
                tpr_accumulate<MODE>(x, k0, k1, m, s, sx);
                if (hi <= c0 + 8) {
                    emit(a, m, s, sx);
                    ++a;
                    hi = a < A ? sp.seg[a + 1] : 0x7fffffff;
                    m = TPR_NEG; s = 0.f; sx = 0.f;
                    k0 = k1;
                } else {
                    k0 = 8;
                }
            }
        };
        float x[8];
#define CPM_TPR_CHUNK(J)                                                   \
        if (J < CPS && (span * CPS + J) * 8 < width) {                     \
            rs.template chunk<(J < CPS ? J : 0)>(x);                       \
            body((span * CPS + J) * 8, x);                                 \
        }
        CPM_TPR_CHUNK(0) CPM_TPR_CHUNK(1) CPM_TPR_CHUNK(2) CPM_TPR_CHUNK(3)
        CPM_TPR_CHUNK(4) CPM_TPR_CHUNK(5) CPM_TPR_CHUNK(6) CPM_TPR_CHUNK(7)
#undef CPM_TPR_CHUNK
    }
}
struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax warp_argmax(ArgMax a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, a.v, o);
        int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
        if (ov > a.v || (ov == a.v && oi < a.i)) { a.v = ov; a.i = oi; }
    }
    return a;
}

// first-index argmax and logsumexp (T = 1) of the fp32 segment already staged in buf[0..w)
__device__ __forceinline__ void segment_stats(const float *buf, int w, int lane, ArgMax &am, float &lse) {
    ArgMax a{-INFINITY, 0x7fffffff};
    for (int i = lane; i < w; i += 32) {
        const float x = buf[i];
        if (x > a.v) { a.v = x; a.i = i; }
    }
    am = warp_argmax(a);
    float s = 0.f;
    for (int i = lane; i < w; i += 32) s += __expf(buf[i] - am.v);
    lse = am.v + __logf(warp_sum(s));
}

// One warp draws the token of one (sequence, attribute): temperature softmax of buf into pr, nucleus cut by exclusive
// descending-order mass, inverse-CDF draw with the Philox uniform of (seed, sequence id, step, attribute).
// MAXC = ceil(max segment width / 32).  Returns the token (all lanes).
template <int MAXC>
__device__ __forceinline__ int sample_segment(const float *buf, float *pr, int w, int lane, const ArgMax &am, float temperature, float top_p,
                                              uint64_t seed, uint64_t sid, int cur_step, int a) {
    const float invt = 1.f / temperature;
    float s = 0.f;
    for (int i = lane; i < w; i += 32) { float e = __expf((buf[i] - am.v) * invt); pr[i] = e; s += e; }
    s = warp_sum(s);
    const bool use_nucleus = top_p > 0.f && top_p < 1.f;
    // reference: softmax, then nucleus divides by (sum + 1e-5); weighted_sampling by sum.
    const float norm = use_nucleus ? 1.f / (s * (1.f + 1e-5f)) : 1.f / s;
    __syncwarp();
    for (int i = lane; i < w; i += 32) pr[i] *= norm;
    __syncwarp();
    // exclusive mass of everything ranked before element i in descending order
    float zkeep = 0.f;
    float ex[MAXC];
#pragma unroll 1
    for (int c = 0; c * 32 + lane < w; ++c) {
        const int i = c * 32 + lane;
        const float pi = pr[i];
        float e = 0.f;
        for (int j = 0; j < w; ++j) {
            const float pj = pr[j];
            e += (pj > pi || (pj == pi && j < i)) ? pj : 0.f;
        }
        const bool keep = !(use_nucleus && e > top_p);
        ex[c] = keep ? e : -1.f;
        zkeep += keep ? pi : 0.f;
    }
    zkeep = warp_sum(zkeep);
    // Philox uniform for (sequence, step, attribute)
    uint4 rnd = Philox::block(make_uint4((uint32_t)sid, (uint32_t)cur_step, (uint32_t)a, (uint32_t)(sid >> 32)),
                              make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float u = (float)(rnd.x >> 8) * (1.0f / 16777216.0f);
    const float target = u * zkeep;
    ArgMax best{-1.f, 0x7fffffff};     // largest exclusive mass <= target among kept
#pragma unroll 1
    for (int c = 0; c * 32 + lane < w; ++c) {
        const float e = ex[c];
        if (e >= 0.f && e <= target && e > best.v) { best.v = e; best.i = c * 32 + lane; }
    }
    best = warp_argmax(best);
    return best.i == 0x7fffffff ? am.i : best.i;
}

}  // namespace cpm
