// Fused samplers for separable (iso / diagonal Gaussian) model plugins:
// one launch = n_draws x C chains of HMCDiag / MALA / random-walk Metropolis.
// theta, rho and the gradient live in registers across all L leapfrog steps;
// momentum comes from in-kernel Philox (or an injected stream), the
// Hamiltonian, the per-chain reductions (group shuffles) and the Metropolis
// test are evaluated in-kernel; only the draw (+ logp, accept flag) goes to HBM.
//
// Reference semantics: hmc.py:36-63, mala.py:40-79, metropolis.py:12-135.
#include "sampler_sep_kernel.cuh"

namespace bk {

// BK_SEP_WIDE=0 (diagnostic): one element block per lane for every D
static bool wide_groups() {
    static int w = -1;
    if (w < 0) { const char* e = getenv("BK_SEP_WIDE"); w = (e && e[0] == '0') ? 0 : 1; }
    return w != 0;
}

template <typename T, int MK>
static int launch_mk(const SepArgs<T>& a, cudaStream_t st) {
    const int D = a.D;
    if (D <= 4) return launch_gj<T, 1, 1, MK>(a, st);
    if (D <= 16) return launch_gj<T, 4, 1, MK>(a, st);
    if (D <= 32) return launch_gj<T, 8, 1, MK>(a, st);
    // fp32 (timed mode): fewer lanes per chain with 16 elements each -- the per-lane overhead of a draw (Philox
    // set-up, reductions, accept, addressing) is paid by 8 lanes instead of 32.  G * J is unchanged, so the
    // Philox blocks and the accept-uniform rule (spare block G * J - 1) are the same as for the <32, 1> layout.
    if (sizeof(T) == 4 && D > 64 && D <= 128 && wide_groups()) return launch_gj<T, 8, 4, MK>(a, st);
    if (sizeof(T) == 4 && D > 32 && D <= 64 && wide_groups()) return launch_gj<T, 4, 4, MK>(a, st);
    if (D <= 64) return launch_gj<T, 16, 1, MK>(a, st);
    if (D <= 128) return launch_gj<T, 32, 1, MK>(a, st);
    if (D <= 256) return launch_gj<T, 32, 2, MK>(a, st);
    if (D <= 512) return launch_gj<T, 32, 4, MK>(a, st);
    set_error("fused separable sampler supports D <= %d (got %d)", SEP_MAX_D, D);
    return BK_E_UNSUPPORTED;
}

int launch_sep_narrow(const SepArgs<float>& a, cudaStream_t st, int* handled);   // sampler_sep_narrow.cu

template <typename T>
int launch_sep_sampler(const SepArgs<T>& a, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        int handled = 0;
        const int rc = launch_sep_narrow(a, st, &handled);
        if (rc || handled) return rc;
    }
    const bool iso = a.model.mu == nullptr && a.model.prec == nullptr && a.model.metric == nullptr;
    return iso ? launch_mk<T, MK_ISO>(a, st) : launch_mk<T, MK_DIAG>(a, st);
}

template int launch_sep_sampler<float>(const SepArgs<float>&, cudaStream_t);
template int launch_sep_sampler<double>(const SepArgs<double>&, cudaStream_t);

}  // namespace bk
