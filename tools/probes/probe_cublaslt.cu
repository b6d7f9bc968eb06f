// Probe: how much do the rollout-step GEMMs (M = 256 sequences) gain from picking among cuBLASLt's heuristic candidates
// instead of taking the first one?  y(M,N) = x(M,K) . W(N,K)^T + b, bf16 in/out, fp32 accumulate, bias epilogue.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/probes/probe_cublaslt.cu -lcublasLt -o /tmp/probe && /tmp/probe
#include <cublasLt.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { auto _e = (x); if (_e != 0) { printf("error %d at %s:%d\n", (int)_e, __FILE__, __LINE__); return 1; } } while (0)

int main() {
    cublasLtHandle_t lt;
    CK(cublasLtCreate(&lt));
    const int shapes[][3] = {{256, 1536, 512}, {256, 512, 512}, {256, 2048, 512}, {256, 512, 2048}, {256, 512, 1216}, {256, 344, 512},
                             {65536, 1536, 512}, {65536, 2048, 512}, {65536, 512, 2048}};
    size_t ws_bytes = 64u << 20;
    void *ws; CK(cudaMalloc(&ws, ws_bytes));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    for (auto &sh : shapes) {
        const int M = sh[0], N = sh[1], K = sh[2];
        __nv_bfloat16 *x, *w, *y, *b;
        CK(cudaMalloc(&x, (size_t)M * K * 2)); CK(cudaMalloc(&w, (size_t)N * K * 2)); CK(cudaMalloc(&y, (size_t)M * N * 2)); CK(cudaMalloc(&b, N * 2));
        CK(cudaMemset(x, 0, (size_t)M * K * 2)); CK(cudaMemset(w, 0, (size_t)N * K * 2)); CK(cudaMemset(b, 0, N * 2));
        // column-major view: y^T (N x M) = W (N x K, as op(A) = A^T of a K x N col-major matrix) . x^T (K x M)
        cublasLtMatmulDesc_t op; CK(cublasLtMatmulDescCreate(&op, CUBLAS_COMPUTE_32F, CUDA_R_32F));
        cublasOperation_t ta = CUBLAS_OP_T, tb = CUBLAS_OP_N;
        CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)));
        CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)));
        cublasLtEpilogue_t epi = CUBLASLT_EPILOGUE_BIAS;
        CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE, &epi, sizeof(epi)));
        CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &b, sizeof(b)));
        cudaDataType_t bt = CUDA_R_16BF;
        CK(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &bt, sizeof(bt)));
        cublasLtMatrixLayout_t la, lb, lc;
        CK(cublasLtMatrixLayoutCreate(&la, CUDA_R_16BF, K, N, K));     // W stored (N,K) row-major = (K,N) col-major
        CK(cublasLtMatrixLayoutCreate(&lb, CUDA_R_16BF, K, M, K));     // x stored (M,K) row-major = (K,M) col-major
        CK(cublasLtMatrixLayoutCreate(&lc, CUDA_R_16BF, N, M, N));     // y stored (M,N) row-major = (N,M) col-major
        cublasLtMatmulPreference_t pref; CK(cublasLtMatmulPreferenceCreate(&pref));
        CK(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes)));
        std::vector<cublasLtMatmulHeuristicResult_t> res(64);
        int found = 0;
        CK(cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, lc, lc, pref, 64, res.data(), &found));
        const float alpha = 1.f, beta = 0.f;
        const int iters = M > 1000 ? 20 : 200;
        float first = 0.f, best = 1e9f; int besti = -1;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < found; ++i) {
            bool ok = true;
            for (int r = 0; r < 3 && ok; ++r)
                ok = cublasLtMatmul(lt, op, &alpha, w, la, x, lb, &beta, y, lc, y, lc, &res[i].algo, ws, ws_bytes, st) == CUBLAS_STATUS_SUCCESS;
            if (!ok) continue;
            // GPU time only: the loop is captured into a CUDA graph (back-to-back host calls are bound by the ~4.7 us call cost)
            cudaGraph_t graph; cudaGraphExec_t exec;
            cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
            for (int r = 0; r < iters; ++r) cublasLtMatmul(lt, op, &alpha, w, la, x, lb, &beta, y, lc, y, lc, &res[i].algo, ws, ws_bytes, st);
            cudaStreamEndCapture(st, &graph);
            cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphLaunch(exec, st);
            cudaStreamSynchronize(st);
            cudaEventRecord(e0, st);
            cudaGraphLaunch(exec, st);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const float us = ms * 1e3f / iters;
            if (i == 0) first = us;
            if (us < best) { best = us; besti = i; }
            if (M < 1000) printf("   #%d %.2f us (workspace %zu, waves %.2f)\n", i, us, res[i].workspaceSize, res[i].wavesCount);
        }
        printf("M=%d N=%d K=%d: %d candidates, heuristic first %.2f us, best #%d %.2f us (%.0f%%)\n", M, N, K, found, first, besti, best,
               100.f * best / first);
        cudaFree(x); cudaFree(w); cudaFree(y); cudaFree(b);
    }
    return 0;
}
