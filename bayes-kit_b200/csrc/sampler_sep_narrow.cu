// Narrow lane layouts of the fused samplers (fp32, isotropic plugin): 4 lanes per chain with J = 5..7 element blocks
// per lane, so that the 4 J lane slots match the ceil(D / 4) + 1 blocks a draw needs (elements + the accept-uniform
// block) instead of the next power of two.  D = 100 needs 26 blocks: 32 slots in the 8 x 4 layout (19 % of the
// Philox / Box-Muller / leapfrog instructions spent on padding), 28 in 4 x 7: c1 4.16 -> 5.06 G chain-steps/s.
// Measured and dropped: 2 lanes per chain (2 x 13 = 26 slots, 224 registers: 3.73 G; 2 x 7 at D = 50: 5.7 vs 7.8 G
// for 4 x 4) and 4 x 8 (same 32 slots as 8 x 4, MALA 14 % slower).
// Separate translation unit: every (G, J) is six fully unrolled kernels.
#include "sampler_sep_kernel.cuh"

namespace bk {

// returns BK_OK with *handled = 1 when a narrow layout served the launch
int launch_sep_narrow(const SepArgs<float>& a, cudaStream_t st, int* handled) {
    *handled = 0;
    const bool iso = a.model.mu == nullptr && a.model.prec == nullptr && a.model.metric == nullptr;
    if (!iso) return BK_OK;
    static int lay = -1;   // BK_SEP_LAYOUT=0 (diagnostic): power-of-two layouts only
    if (lay < 0) { const char* e = getenv("BK_SEP_LAYOUT"); lay = (e && e[0] == '0') ? 0 : 1; }
    if (lay == 0) return BK_OK;
    const int need = (a.D + 3) / 4 + 1;   // element blocks + the accept-uniform block
    if (need <= 16 || need > 28) return BK_OK;
    *handled = 1;
    switch ((need + 3) / 4) {
        case 5: return launch_gj<float, 4, 5, MK_ISO>(a, st);
        case 6: return launch_gj<float, 4, 6, MK_ISO>(a, st);
        default: return launch_gj<float, 4, 7, MK_ISO>(a, st);
    }
}

}  // namespace bk
