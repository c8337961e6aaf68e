// tc_engine.cu -- tensor-core (tcgen05) engine; placeholder until the kernels land.
#include "common.cuh"
namespace nmfb200 {
bool tc_supported(const nmfb200_handle*, const SolveArgs&) { return false; }
void tc_solve(nmfb200_handle*, const SolveArgs&, float*, int64_t, float*, int64_t, nmfb200_result*) {
    throw Error{NMFB200_ENOTSUP, "tensor-core engine not built"};
}
void tc_release(nmfb200_handle*) {}
}  // namespace nmfb200
