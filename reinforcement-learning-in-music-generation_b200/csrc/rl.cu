// RL arithmetic: returns / GAE scans, moments + z-score, PPO losses (fwd+bwd), DQN TD (fwd+bwd).
// Reference formulas: ppo_policy/ppo_train.py:348-402, dqn_policy/IRL_dqn_train.py:285-330
// (compat modes reproduce their quirks).
#include "cpm_common.cuh"

namespace cpm {
namespace {

// ---------------------------------------------------------------- D1 scans
// One warp per trajectory.  Every mode is a first-order linear recurrence X_p = B_p + A_p X_{p-1}
// over a scan position p; 32 positions are combined per step with a warp-level affine scan.
__global__ void __launch_bounds__(128) returns_scan_kernel(const float *__restrict__ rewards, const float *__restrict__ values,
                                                           const float *__restrict__ dones, const float *__restrict__ last_value,
                                                           float *__restrict__ ret, float *__restrict__ adv, int B, int T, float gamma,
                                                           float lam, int mode) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const float *r = rewards + (int64_t)b * T;
    const float *v = values ? values + (int64_t)b * T : nullptr;
    const float *d = dones ? dones + (int64_t)b * T : nullptr;
    float carry = 0.f;
    for (int p0 = 0; p0 < T; p0 += 32) {
        const int p = p0 + lane;
        const bool valid = p < T;
        const int t = (mode == CPM_RET_COMPAT) ? p : T - 1 - p;     // time index read at scan position p
        float A = 1.f, Bv = 0.f, vt = 0.f;
        if (valid) {
            const float nd = (mode == CPM_RET_COMPAT || !d) ? 1.f : 1.f - d[t];
            if (mode == CPM_RET_GAE) {
                vt = v[t];
                const float vnext = (t == T - 1) ? (last_value ? last_value[b] : 0.f) : v[t + 1];
                Bv = r[t] + gamma * nd * vnext - vt;
                A = gamma * lam * nd;
            } else {
                Bv = r[t];
                A = gamma * nd;
            }
        } else {
            A = 1.f; Bv = 0.f;       // identity map
        }
        // inclusive scan of affine maps: (A,B) o (A',B') where primed = earlier positions
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float Ap = __shfl_up_sync(0xffffffffu, A, o);
            const float Bp = __shfl_up_sync(0xffffffffu, Bv, o);
            if (lane >= o) { Bv = fmaf(A, Bp, Bv); A = A * Ap; }
        }
        const float x = fmaf(A, carry, Bv);
        if (valid) {
            const int out_t = (mode == CPM_RET_COMPAT) ? T - 1 - p : t;
            if (mode == CPM_RET_GAE) {
                if (adv) adv[(int64_t)b * T + out_t] = x;
                if (ret) ret[(int64_t)b * T + out_t] = x + vt;
            } else {
                ret[(int64_t)b * T + out_t] = x;
            }
        }
        carry = __shfl_sync(0xffffffffu, x, 31);
    }
}

__global__ void __launch_bounds__(256) moments_kernel(const float *__restrict__ x, const float *__restrict__ sub, int64_t n, double *__restrict__ out3) {
    __shared__ double s1[8], s2[8];
    double a = 0.0, b = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)x[i] - (sub ? (double)sub[i] : 0.0);
        a += v; b += v * v;
    }
    a = warp_sum(a); b = warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s1[warp] = a; s2[warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < 8; ++w) { ta += s1[w]; tb += s2[w]; }
        atomicAdd(&out3[1], ta);
        atomicAdd(&out3[2], tb);
        if (blockIdx.x == 0) atomicAdd(&out3[0], (double)n);
    }
}

__global__ void __launch_bounds__(256) zscore_kernel(const float *__restrict__ x, const float *__restrict__ sub, float *__restrict__ out, int64_t n,
                                                     const double *__restrict__ m3, int unbiased, float eps) {
    const double cnt = m3[0], mean = m3[1] / cnt;
    double var = (m3[2] - cnt * mean * mean) / (unbiased ? cnt - 1.0 : cnt);
    var = var > 0.0 ? var : 0.0;
    const float fm = (float)mean, inv = (float)(1.0 / (sqrt(var) + (double)eps));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = ((x[i] - (sub ? sub[i] : 0.f)) - fm) * inv;
}

// ---------------------------------------------------------------- D2 PPO losses
__device__ __forceinline__ float block_sum(float v, float *scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? scratch[threadIdx.x] : 0.f;
        t = warp_sum(t);
    }
    return t;      // valid in warp 0
}

// compat: new (C) broadcast over T rows of old (T,C); adv (T)
__global__ void __launch_bounds__(256) ppo_compat_kernel(const float *__restrict__ newlp, const float *__restrict__ oldlp, const float *__restrict__ adv,
                                                         float *__restrict__ out, float *__restrict__ dnew, int64_t T, int64_t C, float clip,
                                                         float gscale) {
    __shared__ float scratch[8];
    const float inv_n = 1.f / (float)(T * C);
    float local = 0.f;
    for (int64_t c = threadIdx.x; c < C; c += blockDim.x) {
        const float nl = newlp[c];
        float g = 0.f;
        for (int64_t t = 0; t < T; ++t) {
            const float A = adv[t];
            const float ratio = __expf(nl - oldlp[t * C + c]);
            const float rc = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
            const float arm1 = 0.2f * A, arm2 = rc * A;
            local += fminf(arm1, arm2);
            // d/dnew of min(arm1, arm2): only arm2 depends on new, and only inside the clip range
            const bool inside = ratio > 1.f - clip && ratio < 1.f + clip;
            float w = arm2 < arm1 ? 1.f : (arm2 == arm1 ? 0.5f : 0.f);
            if (inside) g += w * ratio * A;
        }
        if (dnew) dnew[c] = -g * inv_n * gscale;
    }
    const float tot = block_sum(local, scratch);
    if (threadIdx.x == 0) atomicAdd(&out[0], -tot * inv_n);
}

__global__ void __launch_bounds__(256) ppo_standard_kernel(const float *__restrict__ newlp, const float *__restrict__ oldlp, const float *__restrict__ adv,
                                                           const float *__restrict__ entropy, const float *__restrict__ value, const float *__restrict__ ret,
                                                           float *__restrict__ out, float *__restrict__ dnew, float *__restrict__ dent,
                                                           float *__restrict__ dvalue, int64_t n, int64_t nv, float clip, float vf, float entc,
                                                           float gscale) {
    __shared__ float scratch[8];
    float sp = 0.f, se = 0.f, sv = 0.f;
    const float inv_n = 1.f / (float)n, inv_nv = nv > 0 ? 1.f / (float)nv : 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float A = adv[i];
        const float ratio = __expf(newlp[i] - oldlp[i]);
        const float rc = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
        const float a1 = ratio * A, a2 = rc * A;
        sp += fminf(a1, a2);
        const bool inside = ratio > 1.f - clip && ratio < 1.f + clip;
        float g;
        if (a1 < a2) g = a1;                         // d(ratio*A)/dnew = ratio*A
        else if (a1 > a2) g = inside ? a1 : 0.f;
        else g = 0.5f * a1 + (inside ? 0.5f * a1 : 0.f);
        if (dnew) dnew[i] = -g * inv_n * gscale;
        if (entropy) { se += entropy[i]; if (dent) dent[i] = -entc * inv_n * gscale; }
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        const float dlt = value[i] - ret[i];
        sv += dlt * dlt;
        if (dvalue) dvalue[i] = vf * 2.f * dlt * inv_nv * gscale;
    }
    const float tp = block_sum(sp, scratch), te = block_sum(se, scratch), tv = block_sum(sv, scratch);
    if (threadIdx.x == 0) {
        const float pl = -tp * inv_n, el = te * inv_n, vl = tv * inv_nv;
        atomicAdd(&out[0], pl + vf * vl - entc * el);
        atomicAdd(&out[1], pl);
        atomicAdd(&out[2], vl);
        atomicAdd(&out[3], el);
    }
}

// ---------------------------------------------------------------- D3 DQN TD
// row staging of the TD kernel: rows of at most R8_MAXW logits, 16 bytes longer than that in shared memory so that the four
// rows a warp works on sit on different banks
constexpr int R8_MAXW = 512, R8_STRIDE = R8_MAXW + 8;
template <typename T>
__device__ __forceinline__ void r8_load_row(const T *__restrict__ row, int width, T *buf, int sl) {
    constexpr int N = 16 / (int)sizeof(T), MAXC = R8_MAXW / N / 8;     // elements per 16-byte piece; pieces per lane
    const int C = width / N;
    uint4 raw[MAXC];
#pragma unroll
    for (int j = 0; j < MAXC; ++j)
        if (sl + 8 * j < C) raw[j] = *reinterpret_cast<const uint4 *>(row + (sl + 8 * j) * N);
#pragma unroll
    for (int j = 0; j < MAXC; ++j)
        if (sl + 8 * j < C) *reinterpret_cast<uint4 *>(buf + (sl + 8 * j) * N) = raw[j];
    for (int i = C * N + sl; i < width; i += 8) buf[i] = row[i];
}

struct TdParams { int seg[CPM_MAX_ATTR + 1]; int n_attr; };

__device__ __forceinline__ void atomic_add_t(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_t(__nv_bfloat16 *p, float v) { atomicAdd(p, __float2bfloat16_rn(v)); }

// One CTA per replay sample b.  Phase 1: per attribute, max over vocabulary at every position of
// the target net's logits (each logit read once).  Phase 2: top-A over positions (rank by
// counting).  Phase 3: gather Q(s,a), squared error, scatter the gradient.
template <typename T>
__global__ void __launch_bounds__(256) dqn_td_kernel(const T *__restrict__ q_logits, const T *__restrict__ next_logits, const int64_t *__restrict__ action,
                                                     const float *__restrict__ reward, const float *__restrict__ done, float *__restrict__ out,
                                                     T *__restrict__ dq, float *__restrict__ targets_out, int B, int L, int64_t ld, TdParams tp, int A,
                                                     float gamma, float gscale, int mode, int stream_rows) {
    extern __shared__ __align__(16) float sm[];
    float *mx = sm;                       // [n_attr][L]
    float *tg = mx + tp.n_attr * L;       // [n_attr][A]
    __shared__ float scratch[8];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (stream_rows) {
        // eight lanes per position: the whole row of concatenated logits arrives with 16-byte loads (one round trip, all loads
        // issued before the first use) and is staged in its own dtype; the per-attribute maxima are taken out of shared memory
        // with 3-step shuffles, four positions per warp at a time.  (One THREAD per position, as in heads.cu, is slower here:
        // 50 positions per sample leave 11 warps per SM.)
        const int grp = threadIdx.x >> 3, sl = threadIdx.x & 7, ngrp = blockDim.x >> 3, width = tp.seg[tp.n_attr];
        T *stage = reinterpret_cast<T *>(sm + ((tp.n_attr * (L + A) + 3) & ~3)) + grp * R8_STRIDE;
        for (int l0 = 0; l0 < L; l0 += ngrp) {
            const int l = l0 + grp;
            const bool live = l < L;
            __syncwarp();
            r8_load_row(next_logits + ((int64_t)b * L + (live ? l : L - 1)) * ld, width, stage, sl);
            __syncwarp();
            for (int a = 0; a < tp.n_attr; ++a) {
                const int w = tp.seg[a + 1] - tp.seg[a];
                float m = -INFINITY;
                for (int i = sl; i < w; i += 8) m = fmaxf(m, to_f(stage[tp.seg[a] + i]));
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                if (live && sl == 0) mx[a * L + l] = m;
            }
        }
    } else {
        for (int it = warp; it < tp.n_attr * L; it += nwarp) {
            const int a = it / L, l = it % L;
            const T *row = next_logits + ((int64_t)b * L + l) * ld + tp.seg[a];
            const int w = tp.seg[a + 1] - tp.seg[a];
            float m = -INFINITY;
            for (int i = lane; i < w; i += 32) m = fmaxf(m, to_f(row[i]));
            m = warp_max(m);
            if (lane == 0) mx[a * L + l] = m;
        }
    }
    __syncthreads();
    for (int it = threadIdx.x; it < tp.n_attr * L; it += blockDim.x) {
        const int a = it / L, l = it % L;
        if (mode == CPM_TD_COMPAT) {
            const float x = mx[it];
            int rank = 0;
            for (int j = 0; j < L; ++j) { const float y = mx[a * L + j]; rank += (y > x || (y == x && j < l)) ? 1 : 0; }
            if (rank < A) tg[a * A + rank] = x;
        } else {
            const int k = L - 1 - l;
            if (k < A) tg[a * A + k] = mx[it];
        }
    }
    __syncthreads();
    const float rb = reward[b], nd = 1.f - done[b];
    const float inv = 1.f / ((float)B * (float)A * (float)tp.n_attr);
    float local = 0.f;
    for (int it = threadIdx.x; it < tp.n_attr * A; it += blockDim.x) {
        const int a = it / A, k = it % A;
        const float target = rb + gamma * nd * tg[it];
        int64_t act = action[((int64_t)b * A + k) * tp.n_attr + a];
        const int w = tp.seg[a + 1] - tp.seg[a];
        act = act < 0 ? 0 : (act >= w ? w - 1 : act);
        const int64_t loc = (mode == CPM_TD_COMPAT) ? ((int64_t)b * ld + tp.seg[a] + act)                       // batch 0, position b
                                                     : (((int64_t)b * L + (L - 1 - k)) * ld + tp.seg[a] + act);
        const float diff = to_f(q_logits[loc]) - target;
        local += diff * diff;
        if (dq) atomic_add_t(&dq[loc], 2.f * diff * inv * gscale);
        if (targets_out) targets_out[((int64_t)b * A + k) * tp.n_attr + a] = target;
    }
    const float tot = block_sum(local, scratch);
    if (threadIdx.x == 0) atomicAdd(&out[0], tot * inv);
}

__global__ void __launch_bounds__(256) rollout_advance_kernel(const int64_t *__restrict__ tokens, int64_t *__restrict__ htok, int64_t n_tok,
                                                              const float *__restrict__ vals, float *__restrict__ hf, int64_t n_f,
                                                              int32_t *step_dev, int32_t max_steps) {
    griddep_launch();
    griddep_wait();
    const int32_t step = *step_dev;
    if (step < max_steps) {
        if (htok) for (int64_t i = threadIdx.x; i < n_tok; i += blockDim.x) htok[(int64_t)step * n_tok + i] = tokens[i];
        if (hf) for (int64_t i = threadIdx.x; i < n_f; i += blockDim.x) hf[(int64_t)step * n_f + i] = vals[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) *step_dev = step + 1;
}

inline int blocks_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = (int64_t)num_sms() * 8;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace cpm

using namespace cpm;

namespace cpm {
namespace {
// D4 reward head.  eval_a(proj_a(h)).mean(L) is linear in h, so per attribute it collapses to one d-vector:
//   score_a[n] = sigmoid( mean_l h[n,l,:] . u_a + c_a ),  u_a = W_a^T w_a,  c_a = w_a . b_a + bias_a     (host-built, fp32)
// One CTA per sequence: column means of h over L (h is read once), A dot products, sigmoid, average.
template <typename T>
__global__ void __launch_bounds__(256) reward_head_kernel(const T *__restrict__ h, const float *__restrict__ u, const float *__restrict__ c,
                                                          float *__restrict__ reward, float *__restrict__ scores, int L, int d, int A) {
    __shared__ float red[8][CPM_MAX_ATTR];
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T *hn = h + (int64_t)n * L * d;
    float dot[CPM_MAX_ATTR];
#pragma unroll
    for (int a = 0; a < CPM_MAX_ATTR; ++a) dot[a] = 0.f;
    for (int g = tid; g < d / 8; g += 256) {              // this thread's 8 columns
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int l = 0; l < L; ++l) {
            Vec8<T> v;
            v.load(hn + (int64_t)l * d + g * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] += v.v[i];
        }
#pragma unroll
        for (int a = 0; a < CPM_MAX_ATTR; ++a) {
            if (a < A) {
#pragma unroll
                for (int i = 0; i < 8; ++i) dot[a] = fmaf(s[i], u[a * d + g * 8 + i], dot[a]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < CPM_MAX_ATTR; ++a) {
        const float w = warp_sum(dot[a]);
        if (lane == 0) red[warp][a] = w;
    }
    __syncthreads();
    if (tid == 0) {
        float tot = 0.f;
        for (int a = 0; a < A; ++a) {
            float x = 0.f;
            for (int w = 0; w < 8; ++w) x += red[w][a];
            const float sc = 1.f / (1.f + __expf(-(x / (float)L + c[a])));
            if (scores) scores[n * A + a] = sc;
            tot += sc;
        }
        reward[n] = tot / (float)A;
    }
}
}  // namespace
}  // namespace cpm

extern "C" {

int cpm_returns_scan(const float *rewards, const float *values, const float *dones, const float *last_value, float *ret, float *adv,
                     int B, int T, float gamma, float lam, int mode, void *stream) {
    CPM_REQUIRE(rewards, CPM_ERR_NULL, "returns_scan: rewards is NULL");
    CPM_REQUIRE(B >= 0 && T >= 0, CPM_ERR_BAD_SHAPE, "returns_scan: B=%d T=%d", B, T);
    CPM_REQUIRE(mode == CPM_RET_COMPAT || mode == CPM_RET_TOGO || mode == CPM_RET_GAE, CPM_ERR_BAD_SHAPE, "returns_scan: mode %d", mode);
    if (mode == CPM_RET_GAE) CPM_REQUIRE(values && (adv || ret), CPM_ERR_NULL, "returns_scan: GAE needs values and adv/ret");
    else CPM_REQUIRE(ret, CPM_ERR_NULL, "returns_scan: ret is NULL");
    if (B == 0 || T == 0) return CPM_OK;
    returns_scan_kernel<<<(B * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, values, dones, last_value, ret, adv, B, T, gamma, lam, mode);
    return check_launch("returns_scan");
}

int cpm_moments(const float *x, const float *sub, int64_t n, double *out3, void *stream) {
    CPM_REQUIRE(x && out3, CPM_ERR_NULL, "moments: NULL pointer");
    CPM_REQUIRE(n >= 0, CPM_ERR_BAD_SHAPE, "moments: n=%lld", (long long)n);
    if (n == 0) return CPM_OK;
    moments_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, sub, n, out3);
    return check_launch("moments");
}

int cpm_zscore(const float *x, const float *sub, float *out, int64_t n, const double *moments3, int unbiased, float eps, void *stream) {
    CPM_REQUIRE(x && out && moments3, CPM_ERR_NULL, "zscore: NULL pointer");
    CPM_REQUIRE(n >= 0, CPM_ERR_BAD_SHAPE, "zscore: n=%lld", (long long)n);
    if (n == 0) return CPM_OK;
    zscore_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, sub, out, n, moments3, unbiased, eps);
    return check_launch("zscore");
}

int cpm_ppo_loss_fwd_bwd(const float *new_logp, const float *old_logp, const float *adv, const float *entropy, const float *value,
                         const float *ret, float *out, float *dnew, float *dentropy, float *dvalue, int64_t T, int64_t C, int64_t nv,
                         float clip, float vf_coef, float ent_coef, float grad_scale, int mode, void *stream) {
    CPM_REQUIRE(new_logp && old_logp && adv && out, CPM_ERR_NULL, "ppo_loss: NULL pointer");
    CPM_REQUIRE(T > 0 && C > 0, CPM_ERR_BAD_SHAPE, "ppo_loss: T=%lld C=%lld", (long long)T, (long long)C);
    if (mode == CPM_PPO_COMPAT) {
        ppo_compat_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(new_logp, old_logp, adv, out, dnew, T, C, clip, grad_scale);
    } else if (mode == CPM_PPO_STANDARD) {
        CPM_REQUIRE(nv == 0 || (value && ret), CPM_ERR_NULL, "ppo_loss: value/ret required when nv > 0");
        const int64_t n = T * C;
        ppo_standard_kernel<<<blocks_for(n > nv ? n : nv, 256), 256, 0, (cudaStream_t)stream>>>(
            new_logp, old_logp, adv, entropy, value, ret, out, dnew, dentropy, dvalue, n, nv, clip, vf_coef, ent_coef, grad_scale);
    } else {
        return fail(CPM_ERR_BAD_SHAPE, "ppo_loss: mode %d", mode);
    }
    return check_launch("ppo_loss");
}

int cpm_dqn_td_fwd_bwd(const void *q_logits, const void *next_logits, const int64_t *action, const float *reward, const float *done,
                       float *out, void *dq, float *targets_out, int B, int L, int64_t ld, const int *seg_host, int n_attr, int A,
                       float gamma, float grad_scale, int mode, int dtype, void *stream) {
    CPM_REQUIRE(q_logits && next_logits && action && reward && done && out && seg_host, CPM_ERR_NULL, "dqn_td: NULL pointer");
    CPM_REQUIRE(B > 0 && L > 0 && A > 0 && A <= L, CPM_ERR_BAD_SHAPE, "dqn_td: B=%d L=%d A=%d (need 0 < A <= L)", B, L, A);
    CPM_REQUIRE(n_attr >= 1 && n_attr <= CPM_MAX_ATTR, CPM_ERR_BAD_SHAPE, "dqn_td: n_attr=%d", n_attr);
    CPM_REQUIRE(mode == CPM_TD_COMPAT || mode == CPM_TD_STANDARD, CPM_ERR_BAD_SHAPE, "dqn_td: mode %d", mode);
    CPM_REQUIRE(mode != CPM_TD_COMPAT || B <= L, CPM_ERR_BAD_SHAPE,
                "dqn_td: compat gather reads q_logits[0, b, .] and needs B (%d) <= L (%d), like the reference's torch.gather", B, L);
    CPM_REQUIRE((int64_t)seg_host[n_attr] <= ld, CPM_ERR_BAD_SHAPE, "dqn_td: segments exceed ld");
    TdParams tp{};
    tp.n_attr = n_attr;
    for (int a = 0; a <= n_attr; ++a) tp.seg[a] = seg_host[a];
    const size_t esz = dtype == CPM_F32 ? 4 : 2;
    CPM_REQUIRE(dtype == CPM_F32 || dtype == CPM_BF16, CPM_ERR_BAD_DTYPE, "dqn_td: dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    if (dq) {
        cudaError_t e = cudaMemsetAsync(dq, 0, (size_t)B * L * ld * esz, st);
        if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "dqn_td memset: %s", cudaGetErrorString(e));
    }
    size_t smem = (size_t)n_attr * (L + A) * sizeof(float);
    CPM_REQUIRE(smem <= 40 * 1024, CPM_ERR_BAD_SHAPE, "dqn_td: L=%d too long for the shared-memory staging", L);
    // row staging for the 8-lanes-per-position maxima: 32 rows of R8_STRIDE elements behind the [n_attr][L + A] arrays
    const int stream_rows = seg_host[0] == 0 && seg_host[n_attr] <= R8_MAXW && ld % 8 == 0 && aligned16(next_logits);
    if (stream_rows) smem = ((size_t)((n_attr * (L + A) + 3) & ~3)) * sizeof(float) + (size_t)32 * R8_STRIDE * esz;
    static bool attr = false;
    if (!attr) {
        cudaError_t e1 = cudaFuncSetAttribute(dqn_td_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        cudaError_t e2 = cudaFuncSetAttribute(dqn_td_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(CPM_ERR_CUDA, "dqn_td shared-memory attribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        attr = true;
    }
    if (dtype == CPM_F32)
        dqn_td_kernel<float><<<B, 256, smem, st>>>((const float *)q_logits, (const float *)next_logits, action, reward, done, out, (float *)dq,
                                                   targets_out, B, L, ld, tp, A, gamma, grad_scale, mode, stream_rows);
    else
        dqn_td_kernel<__nv_bfloat16><<<B, 256, smem, st>>>((const __nv_bfloat16 *)q_logits, (const __nv_bfloat16 *)next_logits, action, reward,
                                                           done, out, (__nv_bfloat16 *)dq, targets_out, B, L, ld, tp, A, gamma, grad_scale, mode, stream_rows);
    return check_launch("dqn_td");
}

int cpm_rollout_advance(const int64_t *tokens, int64_t *history_tok, int64_t n_tok, const float *vals, float *history_f, int64_t n_f,
                        int32_t *step_dev, int32_t max_steps, void *stream) {
    CPM_REQUIRE(step_dev, CPM_ERR_NULL, "rollout_advance: step_dev is NULL");
    CPM_REQUIRE(!history_tok || tokens, CPM_ERR_NULL, "rollout_advance: tokens is NULL");
    CPM_REQUIRE(!history_f || vals, CPM_ERR_NULL, "rollout_advance: vals is NULL");
    launch_chain(rollout_advance_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, tokens, history_tok, n_tok, vals, history_f, n_f, step_dev, max_steps);
    return check_launch("rollout_advance");
}

int cpm_reward_head(const void *h, const float *u, const float *c, float *reward, float *scores, int N, int L, int d, int n_attr,
                    int dtype, void *stream) {
    CPM_REQUIRE(h && u && c && reward, CPM_ERR_NULL, "reward_head: NULL pointer");
    CPM_REQUIRE(N > 0 && L > 0 && d > 0 && d % 8 == 0 && n_attr >= 1 && n_attr <= CPM_MAX_ATTR, CPM_ERR_BAD_SHAPE,
                "reward_head: N=%d L=%d d=%d n_attr=%d", N, L, d, n_attr);
    CPM_REQUIRE(aligned16(h), CPM_ERR_BAD_ALIGN, "reward_head: h must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CPM_F32) cpm::reward_head_kernel<float><<<N, 256, 0, st>>>((const float *)h, u, c, reward, scores, L, d, n_attr);
    else if (dtype == CPM_BF16)
        cpm::reward_head_kernel<__nv_bfloat16><<<N, 256, 0, st>>>((const __nv_bfloat16 *)h, u, c, reward, scores, L, d, n_attr);
    else return cpm::fail(CPM_ERR_BAD_DTYPE, "reward_head: dtype %d", dtype);
    return cpm::check_launch("reward_head");
}

}  // extern "C"
