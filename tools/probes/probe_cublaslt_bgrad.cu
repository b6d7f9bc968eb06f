// Probe: weight-gradient GEMM dW(N,K) fp32 = gy(T,N)^T . x(T,K) (bf16 inputs) with and without cuBLASLt's fused bias-gradient
// epilogue (CUBLASLT_EPILOGUE_BGRADB: column sums of gy), against the separate column-sum pass it would replace.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/probes/probe_cublaslt_bgrad.cu -lcublasLt -o /tmp/probe && /tmp/probe
#include <cublasLt.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { auto _e = (x); if (_e != 0) { printf("error %d at %s:%d\n", (int)_e, __FILE__, __LINE__); return 1; } } while (0)

int main() {
    cublasLtHandle_t lt;
    CK(cublasLtCreate(&lt));
    const int shapes[][3] = {{65536, 1536, 512}, {65536, 512, 512}, {65536, 2048, 512}, {65536, 512, 2048}};   // T, N, K
    size_t ws_bytes = 256u << 20;
    void *ws; CK(cudaMalloc(&ws, ws_bytes));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    for (auto &sh : shapes) {
        const int T = sh[0], N = sh[1], K = sh[2];
        __nv_bfloat16 *x, *gy; float *dw, *db;
        CK(cudaMalloc(&x, (size_t)T * K * 2)); CK(cudaMalloc(&gy, (size_t)T * N * 2)); CK(cudaMalloc(&dw, (size_t)N * K * 4)); CK(cudaMalloc(&db, N * 4));
        CK(cudaMemset(x, 0, (size_t)T * K * 2)); CK(cudaMemset(gy, 0, (size_t)T * N * 2));
        for (int with_bgrad = 0; with_bgrad < 2; ++with_bgrad) {
            // column-major: dW^T-view C (K x N) = A (K x T: x row-major) . B (T x N: gy row-major viewed (N x T), transposed)
            cublasLtMatmulDesc_t op; CK(cublasLtMatmulDescCreate(&op, CUBLAS_COMPUTE_32F, CUDA_R_32F));
            cublasOperation_t ta = CUBLAS_OP_N, tb = CUBLAS_OP_T;
            CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)));
            CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)));
            if (with_bgrad) {
                cublasLtEpilogue_t epi = CUBLASLT_EPILOGUE_BGRADB;
                CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE, &epi, sizeof(epi)));
                CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &db, sizeof(db)));
                cudaDataType_t bt = CUDA_R_32F;
                CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &bt, sizeof(bt)));
            }
            cublasLtMatrixLayout_t la, lb, lc;
            CK(cublasLtMatrixLayoutCreate(&la, CUDA_R_16BF, K, T, K));
            CK(cublasLtMatrixLayoutCreate(&lb, CUDA_R_16BF, N, T, N));
            CK(cublasLtMatrixLayoutCreate(&lc, CUDA_R_32F, K, N, K));
            cublasLtMatmulPreference_t pref; CK(cublasLtMatmulPreferenceCreate(&pref));
            CK(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes)));
            std::vector<cublasLtMatmulHeuristicResult_t> res(16);
            int found = 0;
            cublasStatus_t hs = cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, lc, lc, pref, 16, res.data(), &found);
            if (hs != CUBLAS_STATUS_SUCCESS || found == 0) { printf("T=%d N=%d K=%d bgrad=%d: no algorithm (status %d)\n", T, N, K, with_bgrad, (int)hs); continue; }
            const float alpha = 1.f, beta = 0.f;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float first = 0.f, best = 1e9f; int besti = -1;
            for (int i = 0; i < found; ++i) {
                bool ok = true;
                for (int r = 0; r < 2 && ok; ++r)
                    ok = cublasLtMatmul(lt, op, &alpha, x, la, gy, lb, &beta, dw, lc, dw, lc, &res[i].algo, ws, ws_bytes, st) == CUBLAS_STATUS_SUCCESS;
                if (!ok) continue;
                cudaStreamSynchronize(st);
                cudaEventRecord(e0, st);
                for (int r = 0; r < 10; ++r) cublasLtMatmul(lt, op, &alpha, x, la, gy, lb, &beta, dw, lc, dw, lc, &res[i].algo, ws, ws_bytes, st);
                cudaEventRecord(e1, st);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const float us = ms * 1e2f;
                if (i == 0) first = us;
                if (us < best) { best = us; besti = i; }
            }
            printf("T=%d N=%d K=%d bgrad=%d: %d candidates, first %.1f us, best #%d %.1f us\n", T, N, K, with_bgrad, found, first, besti, best);
        }
        cudaFree(x); cudaFree(gy); cudaFree(dw); cudaFree(db);
    }
    return 0;
}
