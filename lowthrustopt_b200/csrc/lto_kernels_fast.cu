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

// LTO_K3=hc selects the half-column layout of round 2 (lto_indirect_hc.cu: second-order variables, one thread per 3-vector Nystrom
// half-column, 12 warps with setmaxnreg, three tiles in flight).  Bit-for-bit parity with the default is NOT expected (different
// variables, same controller); it passes the same parity tests.  Default stays the column-per-thread kernel until the half-column
// one is faster on the bench (DESIGN.md section 4 has both sets of numbers).
static bool use_old_k3() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LTO_K3"); v = (e && strcmp(e, "hc") == 0) ? 0 : 1; }
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
