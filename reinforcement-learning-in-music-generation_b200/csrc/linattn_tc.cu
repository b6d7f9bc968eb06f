// tcgen05/TMEM/TMA chunked causal linear attention (bf16).  Placeholder until the kernels land:
// reports "unsupported" so the dispatcher uses the SIMT path.
#include "cpm_common.cuh"
#include "linattn_plan.h"
namespace cpm {
int linattn_fwd_tc_launch(const void *, const void *, const void *, void *, float *, int, int, int, int64_t, int64_t,
                          float, void *, cudaStream_t) { return CPM_ERR_UNSUPPORTED; }
int linattn_bwd_tc_launch(const void *, const void *, const void *, const void *, const float *, const void *, void *,
                          void *, void *, int, int, int, int64_t, int64_t, int64_t, float, void *, cudaStream_t) {
    return CPM_ERR_UNSUPPORTED;
}
}  // namespace cpm
