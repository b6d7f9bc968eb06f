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
