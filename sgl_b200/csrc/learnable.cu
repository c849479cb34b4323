// learnable.cu -- fused forward/backward of LearnableWeightedMessageOp (placeholder until the kernels land).
#include "common.cuh"
using namespace sglb200;
extern "C" {
int sglb200_lw_forward(int, const float *const *, int, int, int, int64_t, int, const float *, const float *, float *,
                       float *, void *)
{
    set_error("lw_forward: not built yet");
    return SGLB200_ERR_INVALID;
}
int sglb200_lw_backward(int, const float *const *, int, int, int, int64_t, int, const float *, const float *,
                        const float *, const float *, float *const *, float *, float *, float *, void *)
{
    set_error("lw_backward: not built yet");
    return SGLB200_ERR_INVALID;
}
}
