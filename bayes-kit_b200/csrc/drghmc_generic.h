#pragma once
#include "model.h"

namespace bk {

constexpr int DRG_KMAX = 6;   // 2^k leapfrog trajectories for proposal k (drghmc.py:424-436)

// DrGhmcDiag for any model plugin: lockstep chains, per-chain predication (drghmc_generic.cu)
size_t drghmc_generic_ws_bytes(const Model& m, int64_t C, int K);
int drghmc_generic(const Model& m, void* theta, void* rho, int64_t C, int K, const double* sizes, const int32_t* counts,
                   double damping, int prob_retry, const void* metric, int64_t n, const bk_rng* rng,
                   const bk_draw_out& out, int32_t* n_used, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace bk
