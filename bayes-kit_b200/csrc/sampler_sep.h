// Host-visible launch interface of the fused separable-model samplers.
#pragma once
#include "sep_common.cuh"

namespace bk {

enum { ALGO_HMC = 0, ALGO_MALA = 1, ALGO_MHRW = 2 };
constexpr int SEP_MAX_D = 512;

template <typename T>
struct SepArgs {
    T* theta;  // [C, D] in/out
    int64_t C;
    int D;
    int vec;   // 1: rows are 16B-aligned and D % 4 == 0 -> vector access
    SepModel<T> model;
    int algo;
    // HMC
    T eps, half_eps;
    int L;
    // MALA: sd = sqrt(2 eps), coef = -0.25 / eps
    T sd, coef;
    // random-walk Metropolis
    T scale, s2;
    int hastings;
    int64_t n_draws;
    bk_rng rng;
    T* draws;         // [n, C, D] or NULL
    T* logp;          // [n, C] or NULL
    int32_t* accept;  // [n, C] or NULL
    int layout;       // BK_DRAWS_NCD / BK_DRAWS_CDN (series-major [C, D, n], staged through shared memory)
    double* mom_mean; // [C, D] running mean (NULL: no streaming moments)
    double* mom_m2;   // [C, D] running sum of squared deviations
    int64_t mom_n0;   // draws the running moments already cover
};

template <typename T>
int launch_sep_sampler(const SepArgs<T>& a, cudaStream_t st);

}  // namespace bk
