// placeholder until the tcgen05 engine lands
#pragma once
#include "common.cuh"
namespace cf {
struct PwTcState {};
inline int pw_tc_init(PwTcState&, int) { return fail(CF_EINVAL, "tcgen05 point-wise engine not built yet"); }
inline void pw_tc_destroy(PwTcState&) {}
inline cudaError_t launch_pw_tc(PwTcState&, int, int, const float*, const float*, float*, int, int, int, EpiArgs, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace cf
