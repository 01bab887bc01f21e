// lto_kernels_fast.cu -- dispatch of the throughput kernels (filled in per configuration).
#include "lto_internal.h"
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace lto {

cudaError_t launch_direct_cw(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_cw(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_cw14(const IndirectArgs& a, cudaStream_t st, int* n_launch);
size_t indirect_cw14_scratch_bytes(int n_sm);
size_t indirect_cwv2_scratch_bytes(int n_sm);

cudaError_t launch_direct_fast(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch) {
    return launch_direct_cw(a, nstate, st, n_launch);
}

cudaError_t launch_indirect_fast(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch) {
    if (ndim == 14) return launch_indirect_cw14(a, st, n_launch);
    return launch_indirect_cw(a, ndim, st, n_launch);
}

size_t indirect_cw_scratch_bytes(int n_sm) {
    return std::max(indirect_cwv2_scratch_bytes(n_sm), indirect_cw14_scratch_bytes(n_sm));
}

}  // namespace lto
