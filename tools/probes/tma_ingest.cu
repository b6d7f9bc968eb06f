// Probe: shared-memory ingest rate of one SM through TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of 64 bf16 x R rows), as the
// GEMM kernels use it - no tensor-core work at all.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_ingest tma_ingest.cu -lcuda && ./tma_ingest
// Every CTA (one per SM, `ctas` of them) walks its own 256-row slab of a [rows x K] bf16 matrix k-block after k-block (the A
// operand of a GEMM) through a ring of `stages` slots of `boxes` boxes each; one thread issues, the same thread waits for the slot
// to land before re-using it.  Prints bytes per clock and SM.
// Variants: plain loads on a local barrier | pairs (cluster of 2) with cta_group::2 loads that count on the LEADER's barrier, as
// tc_gemm.cu does (the leader then releases both CTAs' slot through a remote arrive).
#include "../../reinforcement-learning-in-music-generation_b200/csrc/tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace cpm::tc;

constexpr uint32_t BOX_BYTES = 16384;        // 64 x 128 rows

__global__ void __launch_bounds__(128, 1) probe_local(const __grid_constant__ CUtensorMap tm, int K, int stages, int boxes, int iters, int wrap, long long *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(full + s, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int KB = K / 64;
        int row = (blockIdx.x * 256) % wrap, kb = 0, s = 0;
        uint32_t ph = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters + stages; ++it) {
            if (it >= stages) mbar_wait(full + s, ph ^ 1);
            if (it < iters) {
                mbar_expect_tx(full + s, boxes * BOX_BYTES);
                for (int b = 0; b < boxes; ++b) tma_load_2d(sm + (s * boxes + b) * BOX_BYTES, &tm, full + s, kb * 64, row + b * 128);
                if (++kb == KB) { kb = 0; row += 37 * 256; if (row >= wrap) row -= wrap; }
            }
            if (++s == stages) { s = 0; ph ^= 1; }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

// the same ring driven by TWO issuing threads (warps 0 and 1): slots of even index by one, odd by the other
__global__ void __launch_bounds__(128, 1) probe_local2(const __grid_constant__ CUtensorMap tm, int K, int stages, int boxes, int iters, int wrap, long long *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t full[8];
    __shared__ long long tt[2];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(full + s, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && threadIdx.x < 64) {
        const int who = threadIdx.x >> 5;
        const int KB = K / 64, row0 = blockIdx.x * 256;
        const long long t0 = clock64();
        for (int it = who; it < iters + stages; it += 2) {
            const int s = it % stages;
            if (it >= stages) mbar_wait(full + s, ((it / stages) - 1) & 1);
            if (it < iters) {
                mbar_expect_tx(full + s, boxes * BOX_BYTES);
                for (int b = 0; b < boxes; ++b)
                    tma_load_2d(sm + (s * boxes + b) * BOX_BYTES, &tm, full + s, (it % KB) * 64, (row0 + b * 128 + (it / KB) * 37 * 256) % wrap);
            }
        }
        tt[who] = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = tt[0] > tt[1] ? tt[0] : tt[1];
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_pair(const __grid_constant__ CUtensorMap tm, int K, int stages, int boxes, int iters, int wrap, long long *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t full[8], empty[8];
    const uint32_t rank = cluster_ctarank();
    cluster_sync_all();
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full + s, 2); mbar_init(empty + s, 1); }
        fence_barrier_init();
    }
    cluster_sync_all();
    const uint32_t full0 = mapa_u32(smem_u32(full), 0), empty_peer = mapa_u32(smem_u32(empty), 1);
    const int KB = K / 64, row0 = blockIdx.x * 256;
    long long t0 = 0;
    if (threadIdx.x == 0) {                       // producer of each CTA
        int row = (blockIdx.x * 256) % wrap, kb = 0, s = 0;
        uint32_t ph = 0;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (it >= stages) mbar_wait(empty + s, ph ^ 1);
            if (rank == 0) mbar_expect_tx(full + s, 2 * boxes * BOX_BYTES);
            else mbar_arrive_cluster(full0 + 8 * s);
            for (int b = 0; b < boxes; ++b) tma_load_2d_2sm(sm + (s * boxes + b) * BOX_BYTES, &tm, full0 + 8 * s, kb * 64, row + b * 128);
            if (++kb == KB) { kb = 0; row += 37 * 256; if (row >= wrap) row -= wrap; }
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32 && rank == 0) {  // the leader's consumer: slot landed in both CTAs -> free it in both
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full + s, ph);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + s)) : "memory");
            mbar_arrive_cluster(empty_peer + 8 * s);
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_pair_local(const __grid_constant__ CUtensorMap tm, int K, int stages, int boxes, int iters, int wrap, int use_2sm, long long *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t full[8], empty[8], ready[8];
    const uint32_t rank = cluster_ctarank();
    cluster_sync_all();
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); mbar_init(ready + s, 2); }
        fence_barrier_init();
    }
    cluster_sync_all();
    const uint32_t ready0 = mapa_u32(smem_u32(ready), 0), empty_peer = mapa_u32(smem_u32(empty), 1), full_own = mapa_u32(smem_u32(full), rank);
    const int KB = K / 64;
    long long t0 = 0;
    if (threadIdx.x == 0) {                       // producer of each CTA: own barrier
        int row = (blockIdx.x * 256) % wrap, kb = 0, s = 0;
        uint32_t ph = 0;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (it >= stages) mbar_wait(empty + s, ph ^ 1);
            mbar_expect_tx(full + s, boxes * BOX_BYTES);
            for (int b = 0; b < boxes; ++b) {
                if (use_2sm) tma_load_2d_2sm(sm + (s * boxes + b) * BOX_BYTES, &tm, full_own + 8 * s, kb * 64, row + b * 128);
                else tma_load_2d(sm + (s * boxes + b) * BOX_BYTES, &tm, full + s, kb * 64, row + b * 128);
            }
            if (++kb == KB) { kb = 0; row += 37 * 256; if (row >= wrap) row -= wrap; }
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {               // relay of each CTA: my slot landed -> tell the leader
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full + s, ph);
            mbar_arrive_cluster(ready0 + 8 * s);
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 64 && rank == 0) {  // the leader's consumer: both landed -> free the slot in both CTAs
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(ready + s, ph);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + s)) : "memory");
            mbar_arrive_cluster(empty_peer + 8 * s);
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int rows = 1 << 17;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    long long *out;
    cudaMalloc(&out, 256 * sizeof(long long));
    cudaFuncSetAttribute(probe_local, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(probe_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(probe_local2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(probe_pair_local, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int K : {512}) {
        void *A;
        cudaMalloc(&A, (size_t)rows * K * 2);
        cudaMemset(A, 0, (size_t)rows * K * 2);
        for (int promo = 0; promo < 1; ++promo) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)K * 2};
            cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
            enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            for (int wrap : {8192})
            for (int pair : {0, 1, 3, 4})
                for (int ctas : {148})
                    for (int boxes : {2, 3})
                        for (int stages : {2, 4}) {
                            if ((size_t)stages * boxes * BOX_BYTES > 200 * 1024 || (pair == 2 && stages < 2)) continue;
                            const int iters = 2000;
                            const size_t smem = (size_t)stages * boxes * BOX_BYTES;
                            if (pair == 1) probe_pair<<<ctas, 128, smem>>>(tm, K, stages, boxes, iters, wrap, out);
                            else if (pair == 2) probe_local2<<<ctas, 128, smem>>>(tm, K, stages, boxes, iters, wrap, out);
                            else if (pair >= 3) probe_pair_local<<<ctas, 128, smem>>>(tm, K, stages, boxes, iters, wrap, pair == 4, out);
                            else probe_local<<<ctas, 128, smem>>>(tm, K, stages, boxes, iters, wrap, out);
                            cudaError_t e = cudaDeviceSynchronize();
                            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                            std::vector<long long> h(ctas);
                            cudaMemcpy(h.data(), out, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
                            long long mx = 0;
                            for (long long v : h) mx = v > mx ? v : mx;
                            printf("K %4d wrap %6d promo %s %s ctas %3d boxes/slot %d stages %d (%3zu KB in flight): %.1f B/clk/SM\n", K, wrap, promo ? "256B" : "128B",
                                   pair == 1 ? "pair(cta_group::2, leader barrier)" : pair == 3 ? "pair, plain loads, own barrier+relay" : pair == 4 ? "pair, cta_group::2, own barrier+relay" : pair == 2 ? "local barrier, two issuing threads " : "local barrier                     ", ctas, boxes, stages, smem >> 10,
                                   (double)iters * boxes * BOX_BYTES / (double)mx);
                        }
        }
        cudaFree(A);
    }
    return 0;
}
