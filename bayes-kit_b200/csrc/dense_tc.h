// Tensor-core (tcgen05) path of the dense-precision Gaussian plugin.
#pragma once
#include "model.h"

namespace bk {

enum { TC_MODE_STEP = 0, TC_MODE_GRAD = 1 };

bool dense_tc_enabled(const Model& m);
size_t dense_tc_model_ws_bytes(const bk_model_desc& d);
// builds bf16 hi/lo splits of P (zero padded to Dp) and P*mu in the model workspace
int dense_tc_prepare(Model& m, void* ws, size_t ws_bytes, cudaStream_t st);
size_t dense_tc_hmc_ws_bytes(const Model& m, int64_t C);
int dense_tc_grad(const Model& m, const float* theta, int64_t C, float* grad, void* ws, size_t ws_bytes,
                  cudaStream_t st);
// report_lp: the logp output is log p(theta) instead of the joint log density -- MALA(eps) is this
// pipeline with L = 1 and step sqrt(2 eps) (mala.py:41-45 == one leapfrog step, test_equivalencies.py:12-32)
int dense_tc_hmc(const Model& m, float* theta, float* lp, float* grad, int32_t* cache_valid, int64_t C,
                 double eps, int L, const float* metric, int64_t n_draws, const bk_rng* rng,
                 const bk_draw_out& out, void* ws, size_t ws_bytes, cudaStream_t st, bool report_lp = false);

}  // namespace bk
