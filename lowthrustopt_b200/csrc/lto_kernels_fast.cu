// lto_kernels_fast.cu -- dispatch of the throughput kernels (filled in per configuration).
#include "lto_internal.h"
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace lto {

cudaError_t launch_direct_cw(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_cw(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_q3(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_cw14(const IndirectArgs& a, cudaStream_t st, int* n_launch);
size_t indirect_cw14_scratch_bytes(int n_sm);
size_t indirect_cwv2_scratch_bytes(int n_sm);
size_t indirect_q3_scratch_bytes(int n_sm);

cudaError_t launch_direct_fast(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch) {
    return launch_direct_cw(a, nstate, st, n_launch);
}

// LTO_K3=q3 selects the experimental q-split layout (lto_indirect_q3.cu) for the STM kernel; the default is the
// one-thread-per-column layout (lto_indirect_cw.cu), which measures faster.
static bool use_q3() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LTO_K3"); v = (e && strcmp(e, "q3") == 0) ? 1 : 0; }
    return v == 1;
}

cudaError_t launch_indirect_fast(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch) {
    if (ndim == 14) return launch_indirect_cw14(a, st, n_launch);
    if (a.phi != nullptr && use_q3()) return a.progress ? cudaErrorNotSupported : launch_indirect_q3(a, ndim, st, n_launch);   // no completion counters in the experiment
    return launch_indirect_cw(a, ndim, st, n_launch);
}

size_t indirect_cw_scratch_bytes(int n_sm) {
    const size_t x = indirect_cwv2_scratch_bytes(n_sm), y = indirect_q3_scratch_bytes(n_sm), z = indirect_cw14_scratch_bytes(n_sm);
    return std::max(x, std::max(y, z));
}

}  // namespace lto
