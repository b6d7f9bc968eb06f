// Probe: cost of small-N tcgen05.mma (kind::f16, SS operands, SWIZZLE_128B K-major) per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_small_n umma_small_n.cu && ./umma_small_n
// One CTA, one issuing thread, 256 UMMAs of shape M x N x 16 on fixed shared-memory operands, cycles per instruction from clock64
// (issue -> commit -> mbarrier).  Variants: one accumulator vs two alternating accumulators; K slices of one 64-wide block
// (4 per block, as the kernels do) vs always the same slice.
#include "../../reinforcement-learning-in-music-generation_b200/csrc/tc_common.cuh"
#include <cstdio>
using namespace cpm::tc;

__global__ void __launch_bounds__(128, 1) probe(int M, int N, int n_mma, int n_acc, int blocks, long long *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = idesc_bf16(M, N, false, false);
        // A: `blocks` k-blocks of [M x 128 B] from offset 0; B: k-blocks of [N x 128 B] from offset 128 KB
        const uint64_t dA0 = smem_desc_sw128(smem_u32(sm)), dB0 = smem_desc_sw128(smem_u32(sm + 131072));
        const uint64_t stepA = (uint64_t)(blocks > 1 ? (M * 128) >> 4 : 0), stepB = (uint64_t)(blocks > 1 ? (N * 128) >> 4 : 0);
        const uint32_t acc2 = n_acc == 2 ? 256u : 0u;
        long long t0 = clock64();
        for (int it = 0; it < n_mma / 8; ++it) {                  // two k-blocks (0 and 1) per iteration, 4 K slices each
            uint64_t dA = dA0, dB = dB0;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss(tmem + ((k & 1) ? acc2 : 0u), dA + 2 * k, dB + 2 * k, idesc, (it > 0 || b > 0 || k > 1) ? 1u : 0u);
                dA += stepA;
                dB += stepB;
            }
        }
        mma_commit(&bar);
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
    const int shapes[][2] = {{64, 16}, {64, 32}, {64, 64}, {64, 128}, {64, 256}, {128, 16}, {128, 32}, {128, 64}, {128, 128}, {128, 256}};
    for (auto &s : shapes)
        for (int n_acc : {1, 2})
            for (int blocks : {1, 2}) {
                if (s[1] * n_acc > 512 && n_acc == 2 && s[1] > 256) continue;
                for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 196608>>>(s[0], s[1], 256, n_acc, blocks, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("M=%3d N=%3d accumulators=%d k-blocks=%d : issue %6.1f cyc/mma, complete %6.1f cyc/mma\n", s[0], s[1], n_acc, blocks, h[0] / 256.0, h[1] / 256.0);
            }
    return 0;
}
