// lto_kernels_fast.cu -- dispatch of the throughput kernels (filled in per configuration).
#include "lto_internal.h"

namespace lto {

cudaError_t launch_direct_cw(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_cw(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch);

cudaError_t launch_direct_fast(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch) {
    return launch_direct_cw(a, nstate, st, n_launch);
}

cudaError_t launch_indirect_fast(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch) {
    return launch_indirect_cw(a, ndim, st, n_launch);
}

}  // namespace lto
