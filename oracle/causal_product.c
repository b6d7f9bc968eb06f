/* Oracle: plain-C restatement of pytorch-fast-transformers 0.4.0
 * `causal_product/causal_product_cpu.cpp` (causal_dot_product / _backward).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the real
 * package is not in /root/reference nor installable here; this follows its
 * published algorithm: per (n,h) a sequential scan
 * over L that keeps the E×M running sum kv += k⊗v and emits out_l = q_lᵀ kv;
 * backward = the same forward scan for gQ plus a reverse scan r += q⊗g for
 * gK, gV.  Tensors are (N,H,L,E)/(N,H,L,M) fp32 contiguous and the outputs are
 * accumulated into caller-zeroed buffers, exactly like the ft binding.
 * It doubles as "the reference CPU path" timed by bench.py's cpu_baseline.
 */
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_causal_dot_product(const float *Q, const float *K, const float *V, float *out,
                               int N, int H, int L, int E, int M) {
    long nh_total = (long)N * H;
#pragma omp parallel for schedule(static)
    for (long nh = 0; nh < nh_total; ++nh) {
        float *kv = (float *)calloc((size_t)E * M, sizeof(float));
        const float *q = Q + nh * (long)L * E, *k = K + nh * (long)L * E;
        const float *v = V + nh * (long)L * M;
        float *o = out + nh * (long)L * M;
        for (int l = 0; l < L; ++l) {
            for (int e = 0; e < E; ++e) {
                float ke = k[(long)l * E + e];
                for (int m = 0; m < M; ++m) kv[e * M + m] += ke * v[(long)l * M + m];
            }
            for (int e = 0; e < E; ++e) {
                float qe = q[(long)l * E + e];
                for (int m = 0; m < M; ++m) o[(long)l * M + m] += qe * kv[e * M + m];
            }
        }
        free(kv);
    }
}

void oracle_causal_dot_product_backward(const float *Q, const float *K, const float *V, const float *G,
                                        float *gQ, float *gK, float *gV,
                                        int N, int H, int L, int E, int M) {
    long nh_total = (long)N * H;
#pragma omp parallel for schedule(static)
    for (long nh = 0; nh < nh_total; ++nh) {
        float *kv = (float *)calloc((size_t)E * M, sizeof(float));
        const float *q = Q + nh * (long)L * E, *k = K + nh * (long)L * E;
        const float *v = V + nh * (long)L * M, *g = G + nh * (long)L * M;
        float *gq = gQ + nh * (long)L * E, *gk = gK + nh * (long)L * E, *gv = gV + nh * (long)L * M;
        for (int l = 0; l < L; ++l) {                       /* forward scan: gQ */
            for (int e = 0; e < E; ++e) {
                float ke = k[(long)l * E + e], acc = 0.f;
                for (int m = 0; m < M; ++m) {
                    kv[e * M + m] += ke * v[(long)l * M + m];
                    acc += kv[e * M + m] * g[(long)l * M + m];
                }
                gq[(long)l * E + e] += acc;
            }
        }
        memset(kv, 0, (size_t)E * M * sizeof(float));
        for (int l = L - 1; l >= 0; --l) {                  /* reverse scan: gK, gV */
            for (int e = 0; e < E; ++e) {
                float qe = q[(long)l * E + e], ke = k[(long)l * E + e], acc = 0.f;
                for (int m = 0; m < M; ++m) {
                    kv[e * M + m] += qe * g[(long)l * M + m];
                    acc += kv[e * M + m] * v[(long)l * M + m];
                    gv[(long)l * M + m] += kv[e * M + m] * ke;
                }
                gk[(long)l * E + e] += acc;
            }
        }
        free(kv);
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
