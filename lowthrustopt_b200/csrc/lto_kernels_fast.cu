// lto_kernels_fast.cu -- dispatch of the throughput kernels (filled in per configuration).
#include "lto_internal.h"
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace lto {

cudaError_t launch_direct_cw(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_cw(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch);
cudaError_t launch_indirect_cw14(const IndirectArgs& a, cudaStream_t st, int* n_launch);
#ifdef LTO_K3_EXPERIMENTS        // tools/experiments/build_variant.sh: the measured-and-not-adopted layouts of the 12-dim STM kernel
cudaError_t launch_indirect_hc(const IndirectArgs& a, cudaStream_t st, int* n_launch);
size_t indirect_hc_scratch_bytes(int n_sm);
cudaError_t launch_indirect_wl(const IndirectArgs& a, cudaStream_t st, int* n_launch);
size_t indirect_wl_scratch_bytes(int n_sm);
cudaError_t launch_indirect_hc2(const IndirectArgs& a, cudaStream_t st, int* n_launch);
size_t indirect_hc2_scratch_bytes(int n_sm);
#endif
size_t indirect_cw14_scratch_bytes(int n_sm);
size_t indirect_cwv2_scratch_bytes(int n_sm);

cudaError_t launch_direct_fast(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch) {
    return launch_direct_cw(a, nstate, st, n_launch);
}

// The product has ONE layout of the 12-dim STM kernel: cw (lto_indirect_cw.cu; one thread per STM column, state warps + column warps, two
// tiles).  Two round-2 rebuilds were measured and not adopted (DESIGN.md section 4 has the numbers; sources in tools/experiments/):
//   hc  second-order variables, one thread per half-column, setmaxnreg, three tiles in flight
//   wl  half-columns, every warp owns 8 segment slots outright: no roles, no inter-warp protocol
// A library built by tools/experiments/build_variant.sh (-DLTO_K3_EXPERIMENTS) contains them and selects one with LTO_K3=hc|wl; both pass
// the same parity tests as cw (bit-for-bit equality is not expected: different variables, same controller).
cudaError_t launch_indirect_fast(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch) {
    if (ndim == 14) return launch_indirect_cw14(a, st, n_launch);
#ifdef LTO_K3_EXPERIMENTS
    if (a.phi != nullptr && a.cfg.controller == 0) {
        static int v = -1;
        if (v < 0) { const char* e = getenv("LTO_K3"); v = !e ? 0 : strcmp(e, "hc") == 0 ? 1 : strcmp(e, "wl") == 0 ? 2 : strcmp(e, "hc2") == 0 ? 3 : 0; }
        if (v == 1) return launch_indirect_hc(a, st, n_launch);
        if (v == 2 && a.cfg.err_norm != 0) return launch_indirect_wl(a, st, n_launch);
        if (v == 3) return launch_indirect_hc2(a, st, n_launch);
    }
#endif
    return launch_indirect_cw(a, ndim, st, n_launch);
}

size_t indirect_cw_scratch_bytes(int n_sm) {
    size_t b = std::max(indirect_cwv2_scratch_bytes(n_sm), indirect_cw14_scratch_bytes(n_sm));
#ifdef LTO_K3_EXPERIMENTS
    b = std::max(b, std::max(indirect_hc_scratch_bytes(n_sm), std::max(indirect_wl_scratch_bytes(n_sm), indirect_hc2_scratch_bytes(n_sm))));
#endif
    return b;
}

}  // namespace lto
