// Chunked causal linear attention, CUDA-core implementation (fp32 math, fp32 or bf16 I/O).
//
// This is the fp32-exact parity path (the tcgen05 kernels of linattn_cp.cu take every bf16 call).  One CTA owns one (batch, head, segment) and walks its
// 64-token chunks sequentially, carrying the E×M KV state and the E-vector key sum in
// shared memory; long sequences with few (batch, head) pairs are split into segments whose
// initial states come from a segment-total pass + a prefix scan (workspace).
//
// Maths (ft CausalLinearAttention + causal_product fwd/bwd):
//   Qf=elu(q)+1, Kf=elu(k)+1, A_ij = Qf_i.Kf_j (j<=i)
//   den_i = sum_j A_ij + eps ; out_i = sum_j A_ij v_j / den_i
//   G'_i = g_i/den_i ; gd_i = -(g_i.out_i)/den_i ; W_ij = G'_i.v_j + gd_i (j<=i)
//   dQf_i = sum_j W_ij Kf_j ; dKf_j = sum_i W_ij Qf_i ; dv_j = sum_i A_ij G'_i
#include "cpm_common.cuh"
#include "linattn_plan.h"

namespace cpm {

namespace {

constexpr int CH = 64;        // chunk length (tokens)
constexpr int DH = 64;        // head dim (E = M = 64)
constexpr int LDS = 68;       // padded smem row stride (floats), keeps float4 alignment
constexpr int TILE = CH * LDS;
constexpr int NT = 256;

struct Params {
    const void *q, *k, *v, *o, *go;
    void *out, *gq, *gk, *gv;
    float *den;
    int N, L, H;
    int64_t ld_qkv, ld_o, ld_g;
    float eps;
    int nseg, seg_len;
    float *ws_fwd;   // (N*H*nseg, STATE_FLOATS) forward-direction prefix states  [S | z]
    float *ws_rev;   // same size: reverse-direction suffix states               [R | rz]
};

// acc[r][c] += sum_k A[4ty+r][k] * B[k][4tx+c]
__device__ __forceinline__ void mm64(const float *__restrict__ A, const float *__restrict__ B,
                                     float (&acc)[4][4], int ty, int tx) {
#pragma unroll 2
    for (int k = 0; k < 64; k += 4) {
        float a[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float4 t = *reinterpret_cast<const float4 *>(&A[(4 * ty + r) * LDS + k]);
            a[r][0] = t.x; a[r][1] = t.y; a[r][2] = t.z; a[r][3] = t.w;
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 b = *reinterpret_cast<const float4 *>(&B[(k + kk) * LDS + 4 * tx]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[r][0] = fmaf(a[r][kk], b.x, acc[r][0]);
                acc[r][1] = fmaf(a[r][kk], b.y, acc[r][1]);
                acc[r][2] = fmaf(a[r][kk], b.z, acc[r][2]);
                acc[r][3] = fmaf(a[r][kk], b.w, acc[r][3]);
            }
        }
    }
}

__device__ __forceinline__ void zero(float (&acc)[4][4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
}

// Load a [64 tokens x 64] tile starting at token t0 (rows >= L are zero) into smem, optionally
// applying the feature map, writing row-major (dst) and/or transposed (dstT).
template <typename T, bool PHI>
__device__ __forceinline__ void load_tile(const T *__restrict__ base, int64_t ld, int t0, int L,
                                          float *dst, float *dstT) {
    for (int idx = threadIdx.x; idx < CH * 8; idx += NT) {
        int row = idx >> 3, c8 = (idx & 7) * 8;
        Vec8<T> x;
        bool valid = (t0 + row) < L;
        if (valid) x.load(base + (int64_t)(t0 + row) * ld + c8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float f = valid ? (PHI ? phi(x.v[i]) : x.v[i]) : 0.f;
            if (dst) dst[row * LDS + c8 + i] = f;
            if (dstT) dstT[(c8 + i) * LDS + row] = f;
        }
    }
}

// Load go, out, den for a chunk -> G' = go/den (row-major sG and/or transposed sGT), gd_i.
template <typename T>
__device__ __forceinline__ void load_grad_tile(const T *__restrict__ go, const T *__restrict__ o, int64_t ld,
                                               const float *__restrict__ den, int den_stride, int t0, int L,
                                               float *sG, float *sGT, float *sgd) {
    for (int idx = threadIdx.x; idx < CH * 8; idx += NT) {
        int row = idx >> 3, c8 = (idx & 7) * 8;
        bool valid = (t0 + row) < L;
        Vec8<T> g, y;
        float inv = 0.f, dot = 0.f;
        if (valid) {
            g.load(go + (int64_t)(t0 + row) * ld + c8);
            y.load(o + (int64_t)(t0 + row) * ld + c8);
            inv = 1.f / den[(int64_t)(t0 + row) * den_stride];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float gg = valid ? g.v[i] : 0.f;
            dot += valid ? gg * y.v[i] : 0.f;
            float f = gg * inv;
            if (sG) sG[row * LDS + c8 + i] = f;
            if (sGT) sGT[(c8 + i) * LDS + row] = f;
        }
        // the 8 threads of one row are 8 consecutive lanes
        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        if ((idx & 7) == 0) sgd[row] = -dot * inv;
    }
}

__device__ __forceinline__ void load_state(const float *src, float *sS, float *sz, bool transpose) {
    for (int idx = threadIdx.x; idx < DH * DH; idx += NT) {
        int e = idx >> 6, m = idx & 63;
        float f = src ? src[idx] : 0.f;
        if (transpose) sS[m * LDS + e] = f; else sS[e * LDS + m] = f;
    }
    if (threadIdx.x < DH) sz[threadIdx.x] = src ? src[DH * DH + threadIdx.x] : 0.f;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) linattn_fwd_simt(Params p) {
    extern __shared__ __align__(16) float sm[];
    float *sQ = sm, *sKT = sQ + TILE, *sV = sKT + TILE, *sP = sV + TILE, *sS = sP + TILE;
    float *sz = sS + TILE, *sden = sz + 64;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int seg = blockIdx.x % p.nseg, nh = blockIdx.x / p.nseg, n = nh / p.H, h = nh % p.H;
    const T *q = (const T *)p.q + (int64_t)n * p.L * p.ld_qkv + h * DH;
    const T *k = (const T *)p.k + (int64_t)n * p.L * p.ld_qkv + h * DH;
    const T *v = (const T *)p.v + (int64_t)n * p.L * p.ld_qkv + h * DH;
    T *out = (T *)p.out + (int64_t)n * p.L * p.ld_o + h * DH;
    float *den = p.den ? p.den + (int64_t)n * p.L * p.H + h : nullptr;

    load_state((p.nseg > 1 && seg > 0) ? p.ws_fwd + (int64_t)blockIdx.x * STATE_FLOATS : nullptr, sS, sz, false);
    const int t_begin = seg * p.seg_len, t_end = min(p.L, t_begin + p.seg_len);
    for (int t0 = t_begin; t0 < t_end; t0 += CH) {
        load_tile<T, true>(q, p.ld_qkv, t0, p.L, sQ, nullptr);
        load_tile<T, true>(k, p.ld_qkv, t0, p.L, nullptr, sKT);
        load_tile<T, false>(v, p.ld_qkv, t0, p.L, sV, nullptr);
        __syncthreads();
        float acc[4][4];
        zero(acc);
        mm64(sQ, sKT, acc, ty, tx);
        float rs[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i = 4 * ty + r;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int j = 4 * tx + c;
                float a = (j <= i) ? acc[r][c] : 0.f;
                sP[i * LDS + j] = a;
                s += a + sQ[i * LDS + j] * sz[j];      // inter-chunk part of the normaliser
            }
            rs[r] = s;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], o);
            if (tx == 0) sden[4 * ty + r] = rs[r] + p.eps;
        }
        __syncthreads();
        zero(acc);
        mm64(sP, sV, acc, ty, tx);
        mm64(sQ, sS, acc, ty, tx);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i = 4 * ty + r;
            if (t0 + i < p.L) {
                float inv = 1.f / sden[i];
                T *dst = out + (int64_t)(t0 + i) * p.ld_o + 4 * tx;
#pragma unroll
                for (int c = 0; c < 4; ++c) dst[c] = from_f<T>(acc[r][c] * inv);
                if (tx == 0 && den) den[(int64_t)(t0 + i) * p.H] = sden[i];
            }
        }
        __syncthreads();
        // state update: S += KfT . V ; z += rowsum(KfT)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = sS[(4 * ty + r) * LDS + 4 * tx + c];
        mm64(sKT, sV, acc, ty, tx);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) sS[(4 * ty + r) * LDS + 4 * tx + c] = acc[r][c];
        if (tid < DH) {
            float s = sz[tid];
            for (int j = 0; j < CH; ++j) s += sKT[tid * LDS + j];
            sz[tid] = s;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// segment totals: mode 0 -> [S|z] = sum Kf^T v, sum Kf ; mode 1 -> [R|rz] = sum Qf^T G', sum Qf gd
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) linattn_seg_total_simt(Params p, int mode) {
    extern __shared__ __align__(16) float sm[];
    float *sAT = sm, *sB = sAT + TILE, *sgd = sB + TILE;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int seg = blockIdx.x % p.nseg, nh = blockIdx.x / p.nseg, n = nh / p.H, h = nh % p.H;
    const int64_t off = (int64_t)n * p.L * p.ld_qkv + h * DH;
    const int64_t offo = (int64_t)n * p.L * p.ld_o + h * DH;
    float acc[4][4];
    zero(acc);
    float zs = 0.f;
    const int t_begin = seg * p.seg_len, t_end = min(p.L, t_begin + p.seg_len);
    for (int t0 = t_begin; t0 < t_end; t0 += CH) {
        if (mode == 0) {
            load_tile<T, true>((const T *)p.k + off, p.ld_qkv, t0, p.L, nullptr, sAT);
            load_tile<T, false>((const T *)p.v + off, p.ld_qkv, t0, p.L, sB, nullptr);
        } else {
            load_tile<T, true>((const T *)p.q + off, p.ld_qkv, t0, p.L, nullptr, sAT);
            load_grad_tile<T>((const T *)p.go + offo, (const T *)p.o + offo, p.ld_o,
                              p.den + (int64_t)n * p.L * p.H + h, p.H, t0, p.L, sB, nullptr, sgd);
        }
        __syncthreads();
        mm64(sAT, sB, acc, ty, tx);
        if (tid < DH) {
            if (mode == 0) for (int j = 0; j < CH; ++j) zs += sAT[tid * LDS + j];
            else for (int j = 0; j < CH; ++j) zs += sAT[tid * LDS + j] * sgd[j];
        }
        __syncthreads();
    }
    float *dst = (mode == 0 ? p.ws_fwd : p.ws_rev) + (int64_t)blockIdx.x * STATE_FLOATS;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[(4 * ty + r) * DH + 4 * tx + c] = acc[r][c];
    if (tid < DH) dst[DH * DH + tid] = zs;
}

}  // namespace

// exclusive prefix (reverse=0) or exclusive suffix (reverse=1) over the segment axis, in place
__global__ void linattn_seg_scan(float *ws, int nseg, int reverse) {
    float *base = ws + (int64_t)blockIdx.x * nseg * STATE_FLOATS;
    for (int i = threadIdx.x; i < STATE_FLOATS; i += blockDim.x) {
        float run = 0.f;
        for (int s = 0; s < nseg; ++s) {
            int ss = reverse ? nseg - 1 - s : s;
            float t = base[(int64_t)ss * STATE_FLOATS + i];
            base[(int64_t)ss * STATE_FLOATS + i] = run;
            run += t;
        }
    }
}

namespace {

// ------------------------------------------------------------------------------------------
// backward, forward-direction phase: dq
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) linattn_bwd_dq_simt(Params p) {
    extern __shared__ __align__(16) float sm[];
    float *sG = sm, *sVT = sG + TILE, *sK = sVT + TILE, *sW = sK + TILE, *sST = sW + TILE;
    float *sz = sST + TILE, *sgd = sz + 64;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int seg = blockIdx.x % p.nseg, nh = blockIdx.x / p.nseg, n = nh / p.H, h = nh % p.H;
    const int64_t off = (int64_t)n * p.L * p.ld_qkv + h * DH;
    const int64_t offo = (int64_t)n * p.L * p.ld_o + h * DH;
    const T *q = (const T *)p.q + off;
    T *gq = (T *)p.gq + (int64_t)n * p.L * p.ld_g + h * DH;
    load_state((p.nseg > 1 && seg > 0) ? p.ws_fwd + (int64_t)blockIdx.x * STATE_FLOATS : nullptr, sST, sz, true);
    const int t_begin = seg * p.seg_len, t_end = min(p.L, t_begin + p.seg_len);
    for (int t0 = t_begin; t0 < t_end; t0 += CH) {
        load_grad_tile<T>((const T *)p.go + offo, (const T *)p.o + offo, p.ld_o,
                          p.den + (int64_t)n * p.L * p.H + h, p.H, t0, p.L, sG, nullptr, sgd);
        load_tile<T, false>((const T *)p.v + off, p.ld_qkv, t0, p.L, nullptr, sVT);
        load_tile<T, true>((const T *)p.k + off, p.ld_qkv, t0, p.L, sK, nullptr);
        __syncthreads();
        float acc[4][4];
        zero(acc);
        mm64(sG, sVT, acc, ty, tx);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i = 4 * ty + r;
            float gd = sgd[i];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int j = 4 * tx + c;
                sW[i * LDS + j] = (j <= i) ? acc[r][c] + gd : 0.f;
            }
        }
        __syncthreads();
        zero(acc);
        mm64(sW, sK, acc, ty, tx);
        mm64(sG, sST, acc, ty, tx);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i = 4 * ty + r;
            if (t0 + i < p.L) {
                float gd = sgd[i];
                const T *qs = q + (int64_t)(t0 + i) * p.ld_qkv + 4 * tx;
                T *dst = gq + (int64_t)(t0 + i) * p.ld_g + 4 * tx;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    dst[c] = from_f<T>((acc[r][c] + gd * sz[4 * tx + c]) * dphi(to_f(qs[c])));
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = sST[(4 * ty + r) * LDS + 4 * tx + c];
        mm64(sVT, sK, acc, ty, tx);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) sST[(4 * ty + r) * LDS + 4 * tx + c] = acc[r][c];
        if (tid < DH) {
            float s = sz[tid];
            for (int j = 0; j < CH; ++j) s += sK[j * LDS + tid];
            sz[tid] = s;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// backward, reverse-direction phase: dk, dv
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) linattn_bwd_dkv_simt(Params p) {
    extern __shared__ __align__(16) float sm[];
    float *sK = sm, *sQT = sK + TILE, *sQ = sQT + TILE, *sV = sQ + TILE, *sG = sV + TILE, *sGT = sG + TILE;
    float *sPT = sGT + TILE, *sWT = sPT + TILE, *sR = sWT + TILE, *sRT = sR + TILE;
    float *srz = sRT + TILE, *sgd = srz + 64;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int seg = blockIdx.x % p.nseg, nh = blockIdx.x / p.nseg, n = nh / p.H, h = nh % p.H;
    const int64_t off = (int64_t)n * p.L * p.ld_qkv + h * DH;
    const int64_t offo = (int64_t)n * p.L * p.ld_o + h * DH;
    const T *k = (const T *)p.k + off;
    T *gk = (T *)p.gk + (int64_t)n * p.L * p.ld_g + h * DH;
    T *gv = (T *)p.gv + (int64_t)n * p.L * p.ld_g + h * DH;
    {
        const float *src = (p.nseg > 1 && seg < p.nseg - 1) ? p.ws_rev + (int64_t)blockIdx.x * STATE_FLOATS : nullptr;
        for (int idx = tid; idx < DH * DH; idx += NT) {
            int e = idx >> 6, m = idx & 63;
            float f = src ? src[idx] : 0.f;
            sR[e * LDS + m] = f;
            sRT[m * LDS + e] = f;
        }
        if (tid < DH) srz[tid] = src ? src[DH * DH + tid] : 0.f;
    }
    const int t_begin = seg * p.seg_len, t_end = min(p.L, t_begin + p.seg_len);
    const int nchunk = (t_end - t_begin + CH - 1) / CH;
    for (int c = nchunk - 1; c >= 0; --c) {
        const int t0 = t_begin + c * CH;
        load_tile<T, true>((const T *)p.q + off, p.ld_qkv, t0, p.L, sQ, sQT);
        load_tile<T, true>(k, p.ld_qkv, t0, p.L, sK, nullptr);
        load_tile<T, false>((const T *)p.v + off, p.ld_qkv, t0, p.L, sV, nullptr);
        load_grad_tile<T>((const T *)p.go + offo, (const T *)p.o + offo, p.ld_o,
                          p.den + (int64_t)n * p.L * p.H + h, p.H, t0, p.L, sG, sGT, sgd);
        __syncthreads();
        float acc[4][4];
        zero(acc);
        mm64(sK, sQT, acc, ty, tx);                       // PT[j][i]
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                int j = 4 * ty + r, i = 4 * tx + cc;
                sPT[j * LDS + i] = (i >= j) ? acc[r][cc] : 0.f;
            }
        zero(acc);
        mm64(sV, sGT, acc, ty, tx);                       // WT[j][i]
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                int j = 4 * ty + r, i = 4 * tx + cc;
                sWT[j * LDS + i] = (i >= j) ? acc[r][cc] + sgd[i] : 0.f;
            }
        __syncthreads();
        zero(acc);
        mm64(sPT, sG, acc, ty, tx);
        mm64(sK, sR, acc, ty, tx);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int j = 4 * ty + r;
            if (t0 + j < p.L) {
                T *dst = gv + (int64_t)(t0 + j) * p.ld_g + 4 * tx;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) dst[cc] = from_f<T>(acc[r][cc]);
            }
        }
        zero(acc);
        mm64(sWT, sQ, acc, ty, tx);
        mm64(sV, sRT, acc, ty, tx);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int j = 4 * ty + r;
            if (t0 + j < p.L) {
                const T *ks = k + (int64_t)(t0 + j) * p.ld_qkv + 4 * tx;
                T *dst = gk + (int64_t)(t0 + j) * p.ld_g + 4 * tx;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    dst[cc] = from_f<T>((acc[r][cc] + srz[4 * tx + cc]) * dphi(to_f(ks[cc])));
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[r][cc] = sR[(4 * ty + r) * LDS + 4 * tx + cc];
        mm64(sQT, sG, acc, ty, tx);
        float acc2[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc2[r][cc] = sRT[(4 * ty + r) * LDS + 4 * tx + cc];
        mm64(sGT, sQ, acc2, ty, tx);
        float rz = 0.f;
        if (tid < DH) {
            rz = srz[tid];
            for (int i = 0; i < CH; ++i) rz += sQT[tid * LDS + i] * sgd[i];
        }
        __syncthreads();     // all reads of sR/sRT/srz by the dv/dk GEMMs above are complete
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                sR[(4 * ty + r) * LDS + 4 * tx + cc] = acc[r][cc];
                sRT[(4 * ty + r) * LDS + 4 * tx + cc] = acc2[r][cc];
            }
        if (tid < DH) srz[tid] = rz;
        __syncthreads();
    }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
    return CPM_OK;
}

template <typename T>
int fwd_impl(Params p, cudaStream_t st) {
    const int NH = p.N * p.H;
    const size_t smem_main = (5 * TILE + 128) * sizeof(float);
    const size_t smem_tot = (2 * TILE + 64) * sizeof(float);
    int rc;
    if ((rc = set_smem(linattn_fwd_simt<T>, smem_main))) return rc;
    if (p.nseg > 1) {
        if ((rc = set_smem(linattn_seg_total_simt<T>, smem_tot))) return rc;
        linattn_seg_total_simt<T><<<NH * p.nseg, NT, smem_tot, st>>>(p, 0);
        linattn_seg_scan<<<NH, 256, 0, st>>>(p.ws_fwd, p.nseg, 0);
    }
    linattn_fwd_simt<T><<<NH * p.nseg, NT, smem_main, st>>>(p);
    return check_launch("linattn_fwd_simt");
}

template <typename T>
int bwd_impl(Params p, cudaStream_t st) {
    const int NH = p.N * p.H;
    const size_t smem_dq = (5 * TILE + 128) * sizeof(float);
    const size_t smem_dkv = (10 * TILE + 128) * sizeof(float);
    const size_t smem_tot = (2 * TILE + 64) * sizeof(float);
    int rc;
    if ((rc = set_smem(linattn_bwd_dq_simt<T>, smem_dq))) return rc;
    if ((rc = set_smem(linattn_bwd_dkv_simt<T>, smem_dkv))) return rc;
    if (p.nseg > 1) {
        if ((rc = set_smem(linattn_seg_total_simt<T>, smem_tot))) return rc;
        linattn_seg_total_simt<T><<<NH * p.nseg, NT, smem_tot, st>>>(p, 0);
        linattn_seg_total_simt<T><<<NH * p.nseg, NT, smem_tot, st>>>(p, 1);
        linattn_seg_scan<<<NH, 256, 0, st>>>(p.ws_fwd, p.nseg, 0);
        linattn_seg_scan<<<NH, 256, 0, st>>>(p.ws_rev, p.nseg, 1);
    }
    linattn_bwd_dq_simt<T><<<NH * p.nseg, NT, smem_dq, st>>>(p);
    linattn_bwd_dkv_simt<T><<<NH * p.nseg, NT, smem_dkv, st>>>(p);
    return check_launch("linattn_bwd_simt");
}

}  // namespace

int linattn_fwd_simt_launch(const void *q, const void *k, const void *v, void *out, float *den, int N, int L, int H,
                            int64_t ld_qkv, int64_t ld_o, int dtype, float eps, void *ws, cudaStream_t st) {
    Params p{};
    p.q = q; p.k = k; p.v = v; p.out = out; p.den = den; p.N = N; p.L = L; p.H = H;
    p.ld_qkv = ld_qkv; p.ld_o = ld_o; p.eps = eps;
    plan_segments(N, H, L, &p.nseg, &p.seg_len);
    p.ws_fwd = (float *)ws;
    p.ws_rev = p.ws_fwd + (int64_t)N * H * p.nseg * STATE_FLOATS;
    return dtype == CPM_F32 ? fwd_impl<float>(p, st) : fwd_impl<__nv_bfloat16>(p, st);
}

int linattn_bwd_simt_launch(const void *q, const void *k, const void *v, const void *out, const float *den,
                            const void *gout, void *gq, void *gk, void *gv, int N, int L, int H, int64_t ld_qkv,
                            int64_t ld_o, int64_t ld_g, int dtype, float eps, void *ws, cudaStream_t st) {
    Params p{};
    p.q = q; p.k = k; p.v = v; p.o = out; p.go = gout; p.den = const_cast<float *>(den);
    p.gq = gq; p.gk = gk; p.gv = gv; p.N = N; p.L = L; p.H = H;
    p.ld_qkv = ld_qkv; p.ld_o = ld_o; p.ld_g = ld_g; p.eps = eps;
    plan_segments(N, H, L, &p.nseg, &p.seg_len);
    p.ws_fwd = (float *)ws;
    p.ws_rev = p.ws_fwd + (int64_t)N * H * p.nseg * STATE_FLOATS;
    return dtype == CPM_F32 ? bwd_impl<float>(p, st) : bwd_impl<__nv_bfloat16>(p, st);
}

// Segment prefix (and optionally suffix) states for callers that run their own main kernels (the
// tcgen05 path): fills ws_fwd with exclusive forward prefixes of [S|z] and, when `reverse_too`,
// ws_rev with exclusive suffixes of [R|rz].
int linattn_segment_states_launch(const void *q, const void *k, const void *v, const void *out, const float *den,
                                  const void *gout, int N, int L, int H, int64_t ld_qkv, int64_t ld_o, int dtype, void *ws,
                                  bool reverse_too, cudaStream_t st) {
    Params p{};
    p.q = q; p.k = k; p.v = v; p.o = out; p.go = gout; p.den = const_cast<float *>(den); p.N = N; p.L = L; p.H = H;
    p.ld_qkv = ld_qkv; p.ld_o = ld_o;
    plan_segments(N, H, L, &p.nseg, &p.seg_len);
    p.ws_fwd = (float *)ws;
    p.ws_rev = p.ws_fwd + (int64_t)N * H * p.nseg * STATE_FLOATS;
    if (p.nseg <= 1) return CPM_OK;
    const int NH = N * H;
    const size_t smem_tot = (2 * TILE + 64) * sizeof(float);
    int rc;
    if (dtype == CPM_F32) {
        if ((rc = set_smem(linattn_seg_total_simt<float>, smem_tot))) return rc;
        linattn_seg_total_simt<float><<<NH * p.nseg, NT, smem_tot, st>>>(p, 0);
        if (reverse_too) linattn_seg_total_simt<float><<<NH * p.nseg, NT, smem_tot, st>>>(p, 1);
    } else {
        if ((rc = set_smem(linattn_seg_total_simt<__nv_bfloat16>, smem_tot))) return rc;
        linattn_seg_total_simt<__nv_bfloat16><<<NH * p.nseg, NT, smem_tot, st>>>(p, 0);
        if (reverse_too) linattn_seg_total_simt<__nv_bfloat16><<<NH * p.nseg, NT, smem_tot, st>>>(p, 1);
    }
    linattn_seg_scan<<<NH, 256, 0, st>>>(p.ws_fwd, p.nseg, 0);
    if (reverse_too) linattn_seg_scan<<<NH, 256, 0, st>>>(p.ws_rev, p.nseg, 1);
    return check_launch("linattn_segment_states");
}

}  // namespace cpm
