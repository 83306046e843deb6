// placeholder until the tcgen05 path lands
#include "common.cuh"
int vqb_conv2d_fwd_tc(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int, int, int, int,
                      int, float, float, cudaStream_t) {
    vqb_set_error("conv2d_fwd(tcgen05): not built");
    return VQB_ERR_UNSUPPORTED;
}
int vqb_conv2d_wgrad_tc(const void*, const void*, float*, int, int, int, int, int, int, int, int, cudaStream_t) {
    vqb_set_error("conv2d_wgrad(tcgen05): not built");
    return VQB_ERR_UNSUPPORTED;
}
