// lto_kernels_fast.cu -- dispatch of the throughput kernels (filled in per configuration).
#include "lto_internal.h"
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace lto {

cudaError_t launch_direct_cw(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_cw(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_cw14(const IndirectArgs& a, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_hc(const IndirectArgs& a, cudaStream_t st, int* n_launch);
size_t indirect_hc_scratch_bytes(int n_sm);
size_t indirect_cw14_scratch_bytes(int n_sm);
size_t indirect_cwv2_scratch_bytes(int n_sm);

cudaError_t launch_direct_fast(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch) {
    return launch_direct_cw(a, nstate, st, n_launch);
}

// (development switch of round 2, removed once the half-column kernel has replaced the column-per-thread one: LTO_K3=cw)
static bool use_old_k3() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LTO_K3"); v = (e && strcmp(e, "cw") == 0) ? 1 : 0; }
    return v == 1;
}

cudaError_t launch_indirect_fast(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch) {
    if (ndim == 14) return launch_indirect_cw14(a, st, n_launch);
    if (a.phi != nullptr && !use_old_k3()) return launch_indirect_hc(a, st, n_launch);
    return launch_indirect_cw(a, ndim, st, n_launch);
}

size_t indirect_cw_scratch_bytes(int n_sm) {
    return std::max(indirect_hc_scratch_bytes(n_sm), std::max(indirect_cwv2_scratch_bytes(n_sm), indirect_cw14_scratch_bytes(n_sm)));
}

}  // namespace lto
