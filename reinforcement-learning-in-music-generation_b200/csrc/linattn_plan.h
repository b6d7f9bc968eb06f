// Segment plan shared by the SIMT and tcgen05 linear-attention kernels.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace cpm {

constexpr int STATE_FLOATS = 64 * 64 + 64;   // [S (E x M) | z (E)] per (batch, head, segment)

// A sequence is cut into `nseg` segments of `seg_len` tokens (a multiple of 128) when there are
// too few (batch, head) pairs to fill the GPU; each segment becomes one CTA.
inline void plan_segments(int N, int H, int L, int *nseg, int *seg_len) {
    const int NH = N * H;
    int want = 1;
    if (NH < 148) {
        want = (296 + NH - 1) / NH;
        int cap = L / 256;
        if (cap < 1) cap = 1;
        if (want > cap) want = cap;
    }
    int sl = (L + want - 1) / want;
    sl = ((sl + 127) / 128) * 128;
    if (sl < 128) sl = 128;
    *seg_len = sl;
    *nseg = (L + sl - 1) / sl;
}

__global__ void linattn_seg_scan(float *ws, int nseg, int reverse);

int linattn_fwd_simt_launch(const void *q, const void *k, const void *v, void *out, float *den, int N, int L, int H,
                            int64_t ld_qkv, int64_t ld_o, int dtype, float eps, void *ws, cudaStream_t st);
int linattn_bwd_simt_launch(const void *q, const void *k, const void *v, const void *out, const float *den,
                            const void *gout, void *gq, void *gk, void *gv, int N, int L, int H, int64_t ld_qkv,
                            int64_t ld_o, int64_t ld_g, int dtype, float eps, void *ws, cudaStream_t st);
int linattn_segment_states_launch(const void *q, const void *k, const void *v, const void *out, const float *den,
                                  const void *gout, int N, int L, int H, int64_t ld_qkv, int64_t ld_o, int dtype, void *ws,
                                  bool reverse_too, cudaStream_t st);
// chunk-parallel tcgen05 path (linattn_cp.cu)
// width = head width: 64 (the reference's heads) or 128
int64_t linattn_cp_workspace_bytes(int N, int L, int H, int width);
int64_t linattn_cp_saved_bytes(int N, int L, int H, int width);
bool linattn_cp_streams(int N, int H, int width);       // does a call of this shape take the streaming state kernels?
void linattn_cp_set_timing_buffer(long long *p);
int linattn_fwd_cp_launch(const void *q, const void *k, const void *v, void *out, float *den, int N, int L, int H, int width,
                          int64_t ld_qkv, int64_t ld_o, float eps, void *ws, void *saved, cudaStream_t st);
int linattn_bwd_cp_launch(const void *q, const void *k, const void *v, const void *out, const float *den,
                          const void *gout, void *gq, void *gk, void *gv, int N, int L, int H, int width, int64_t ld_qkv,
                          int64_t ld_o, int64_t ld_g, void *ws, const void *saved, cudaStream_t st);

}  // namespace cpm
