// TEMPORARY bring-up stubs: the tensor-core paths report "unsupported" so AUTO falls to the small-shape kernels.
#include "common.cuh"
namespace csm {
bool attn_mma_supported(int, int64_t, int64_t, int64_t, int64_t) { return false; }
int attn_fwd_mma_launch(const void*, const void*, const void*, void*, float*, int, int, int, int, int, int64_t,
                        int64_t, int64_t, int64_t, float, cudaStream_t) { set_error("attn_mma: not built"); return CSM_ERR_SHAPE; }
int attn_bwd_mma_launch(const void*, const void*, const void*, const void*, const float*, const void*, void*,
                        void*, void*, float*, int, int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t,
                        int64_t, int64_t, float, cudaStream_t) { set_error("attn_mma: not built"); return CSM_ERR_SHAPE; }
}
