// hbf_fast.cuh -- shared-memory tiled HBF cascade kernels.
#pragma once
#include "common.cuh"
#define IDSP_HBF_FAST_NOT_APPLICABLE 12346
namespace idsp {
static int hbf_dec_fast_try(idsp_ctx *, int, float *, const float *, float *, size_t, size_t,
                            size_t, int) {
    return IDSP_HBF_FAST_NOT_APPLICABLE;
}
}  // namespace idsp
