// Host side of TMA: 2-D bf16 tensor maps (SWIZZLE_128B, 64-element = 128-byte inner box) for every tcgen05 kernel of the library
// (linear attention, GEMMs).  cuTensorMapEncodeTiled is resolved through the runtime's driver entry point query, so the library
// links against cudart only.
#include "cpm_common.cuh"
#include "tc_common.cuh"

namespace cpm {

// ---------------------------------------------------------------- host: tensor maps
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}
}  // namespace

int make_tmap_bf16_2d(CUtensorMap *out, const void *base, uint64_t inner_elems, uint64_t rows, uint64_t row_stride_elems,
                      uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return fail(CPM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    // The encode call is a driver-API entry point and needs a current context on THIS thread; autograd
    // worker threads may not have touched the runtime yet, so bind the primary context once per thread.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(nullptr);
        ctx_bound = true;
    }
    cuuint64_t dims[2] = {inner_elems, rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CPM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CPM_OK;
}

}  // namespace cpm
