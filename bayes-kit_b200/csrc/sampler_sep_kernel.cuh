// Kernel template of the fused separable-model samplers (instantiated by sampler_sep.cu and
// sampler_sep_narrow.cu, which split the lane layouts between them to keep compile times down).
#pragma once
#include <stdlib.h>

#include "sampler_sep.h"

namespace bk {

constexpr int SERIES_TILE = 8;   // draws staged per series before a flush: 8 floats = one 32-byte sector

// MOM: streaming moments -- per element the shifted sums of the launch's draws stay in registers
// and are merged ONCE into the running fp64 (mean, M2) at the end; no pass over stored draws.
template <typename T, int G, int J, int MK, int ALGO, bool MOM>
__global__ void __launch_bounds__(128) k_sep_sampler(SepArgs<T> a) {
    using A = Ar<T>;
    constexpr int NE = 4 * J;
    constexpr int CPB = 128 / G;     // chains per CTA
    extern __shared__ __align__(16) unsigned char sep_smem[];
    T* stage = reinterpret_cast<T*>(sep_smem);   // [SERIES_TILE][CPB][D] when the series-major layout is on
    const bool series = a.layout == BK_DRAWS_CDN && a.draws != nullptr;
    // Lanes past the last chain shadow chain C-1 (they must stay in the warp
    // for the full-mask group shuffles) and skip every store.
    const int64_t chain_raw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool active = chain_raw < a.C;
    const int64_t chain = active ? chain_raw : a.C - 1;
    Lanes<T, G, J> ln;
    ln.lane = threadIdx.x % G;
    ln.D = a.D;
    ln.vec = a.vec != 0;
    SepGauss<T, G, J, MK> md;
    md.init(a.model, ln);

    T th[NE];
    ln.load(a.theta + chain * (int64_t)a.D, th, T(0));
    T lp = md.logp(th);  // cached log p(theta) (mala.py:31, metropolis.py:99)
    T m_x0[MOM ? NE : 1], m_s1[MOM ? NE : 1], m_s2[MOM ? NE : 1];
    if constexpr (MOM) {
#pragma unroll
        for (int k = 0; k < NE; ++k) { m_x0[k] = T(0); m_s1[k] = T(0); m_s2[k] = T(0); }
    }

    const int ub = accept_block(a.D);
    for (int64_t t = 0; t < a.n_draws; ++t) {
        T z[NE];
        uint32_t raw2[2] = {0u, 0u};
        ln.normals(a.rng, a.C, chain, t, z, raw2, ub);
        // accept uniform (common.cuh: philox_accept_uniform): words of element block ub = ceil(D / 4) of the normal
        // stream -- made by the pass above when the lane layout has a slot for that block
        T u;
        if (a.rng.mode == BK_RNG_PHILOX) {
            if (ub < G * J) {
                const uint32_t w0 = __shfl_sync(0xffffffffu, raw2[0], ub % G, G);
                const uint32_t w1 = sizeof(T) == 8 ? __shfl_sync(0xffffffffu, raw2[1], ub % G, G) : 0u;
                u = uniform_of_words<T>(w0, w1);
            } else {
                u = philox_accept_uniform<T>(a.rng.seed, (uint32_t)(a.rng.chain_offset + (uint64_t)chain),
                                             (uint32_t)(a.rng.draw_offset + (uint64_t)t), a.D);
            }
        } else {
            u = ln.uniform(a.rng, a.C, chain, t, 0);
        }
        const T logu = log_u(u);
        bool acc;
        T out_lp;
        if constexpr (ALGO == ALGO_HMC && MK == MK_ISO && sizeof(T) == 4) {
            // fp32 timed mode, isotropic target.  The plugin's gradient -prec x is linear, so the kick and drift of
            // hmc.py:47-50 fold into the position (Stormer-Verlet) form of the SAME leapfrog,
            //     x_{n+1} = c x_n - x_{n-1},   c = 2 - eps^2 prec,   x_1 = eps rho + (c / 2) x_0,
            // ONE FMA per element and leapfrog step; the momentum never exists and is recovered where the
            // Hamiltonian needs it: eps rho_L = (c / 2) x_L - x_{L-1} (last drift undone, final half kick of
            // hmc.py:52 applied).  log p - kinetic is reduced in ONE group shuffle per Hamiltonian: the kernel is
            // issue-bound, not HBM-bound.
            const T prec = md.prec_scalar;
            const T c = T(2) - a.eps * a.eps * prec, hc = T(0.5) * c, inv_eps2 = T(1) / (a.eps * a.eps);
            T sx = T(0), sz = T(0);
#pragma unroll
            for (int k = 0; k < NE; ++k) { sx = fmaf(th[k], th[k], sx); sz = fmaf(z[k], z[k], sz); }
            const T h0 = T(-0.5) * group_sum<G>(fmaf(prec, sx, sz));
            if (!(a.L > 0 && a.eps * a.eps * prec >= T(0x1p-10))) {
                // eps^2 prec below 2^-10 (c would round towards 2 and lose the force) or L = 0: velocity form,
                // kick and drift as two FMAs
                const T kick = -a.eps * prec, hkick = -a.half_eps * prec;
                T q[NE];
#pragma unroll
                for (int k = 0; k < NE; ++k) {
                    q[k] = th[k];
                    z[k] = fmaf(-hkick, th[k], z[k]);   // backward half kick (hmc.py:46)
                }
                for (int s = 0; s < a.L; ++s) {
#pragma unroll
                    for (int k = 0; k < NE; ++k) {
                        z[k] = fmaf(kick, q[k], z[k]);
                        q[k] = fmaf(a.eps, z[k], q[k]);
                    }
                }
                T s1 = T(0), sr = T(0);
#pragma unroll
                for (int k = 0; k < NE; ++k) {
                    z[k] = fmaf(hkick, q[k], z[k]);     // forward half kick (hmc.py:52)
                    s1 = fmaf(q[k], q[k], s1);
                    sr = fmaf(z[k], z[k], sr);
                }
                const T h1 = T(-0.5) * group_sum<G>(fmaf(prec, s1, sr));
                acc = logu < h1 - h0;
                if (acc) {
#pragma unroll
                    for (int k = 0; k < NE; ++k) th[k] = q[k];
                }
                out_lp = acc ? h1 : h0;
            } else {
                T xa[NE], xb[NE];
#pragma unroll
                for (int k = 0; k < NE; ++k) xb[k] = fmaf(a.eps, z[k], hc * th[k]);   // x_1
                auto finish = [&](const T (&cur)[NE], const T (&prev)[NE]) {
                    T s1 = T(0), se = T(0);
#pragma unroll
                    for (int k = 0; k < NE; ++k) {
                        const T e = fmaf(hc, cur[k], -prev[k]);   // eps * rho_L
                        s1 = fmaf(cur[k], cur[k], s1);
                        se = fmaf(e, e, se);
                    }
                    const T h1 = T(-0.5) * group_sum<G>(fmaf(prec, s1, inv_eps2 * se));
                    acc = logu < h1 - h0;
                    if (acc) {
#pragma unroll
                        for (int k = 0; k < NE; ++k) th[k] = cur[k];
                    }
                    out_lp = acc ? h1 : h0;
                };
                if (a.L == 1) {
                    finish(xb, th);
                } else {
                    // x_s in A, x_{s-1} in B: pair steps, then ONE single step, so that both parities of L run the
                    // same tail (B = c A - B; finish(B, A)).  Odd L: the first recurrence step is peeled and reads th
                    // itself (no copy of the state); even L: th is copied and the recurrence starts one step earlier.
                    auto run = [&](T (&A_)[NE], T (&B_)[NE], int s) {
                        for (; s + 2 < a.L; s += 2) {
#pragma unroll
                            for (int k = 0; k < NE; ++k) {
                                B_[k] = fmaf(c, A_[k], -B_[k]);
                                A_[k] = fmaf(c, B_[k], -A_[k]);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < NE; ++k) B_[k] = fmaf(c, A_[k], -B_[k]);
                        finish(B_, A_);
                    };
                    if (a.L & 1) {
#pragma unroll
                        for (int k = 0; k < NE; ++k) xa[k] = fmaf(c, xb[k], -th[k]);   // x_2
                        run(xa, xb, 2);
                    } else {
#pragma unroll
                        for (int k = 0; k < NE; ++k) xa[k] = th[k];
                        run(xb, xa, 1);
                    }
                }
            }
        } else if constexpr (ALGO == ALGO_HMC) {
            // hmc.py:55-63
            const T h0 = A::sub(md.logp(th), md.kinetic(z));
            T q[NE];
#pragma unroll
            for (int k = 0; k < NE; ++k) {  // backward half kick (hmc.py:46)
                q[k] = th[k];
                z[k] = A::sub(z[k], A::mul(a.half_eps, md.mgrad(th, k)));
            }
            for (int s = 0; s < a.L; ++s) {  // hmc.py:47-50
#pragma unroll
                for (int k = 0; k < NE; ++k) {
                    z[k] = A::add(z[k], A::mul(a.eps, md.mgrad(q, k)));
                    q[k] = A::add(q[k], A::mul(a.eps, z[k]));
                }
            }
#pragma unroll
            for (int k = 0; k < NE; ++k)  // forward half kick (hmc.py:52)
                z[k] = A::add(z[k], A::mul(a.half_eps, md.mgrad(q, k)));
            const T h1 = A::sub(md.logp(q), md.kinetic(z));
            acc = logu < A::sub(h1, h0);
            if (acc) {
#pragma unroll
                for (int k = 0; k < NE; ++k) th[k] = q[k];
            }
            out_lp = acc ? h1 : h0;
        } else if constexpr (ALGO == ALGO_MALA) {
            // mala.py:40-66
            T q[NE];
#pragma unroll
            for (int k = 0; k < NE; ++k)
                q[k] = A::add(A::add(th[k], A::mul(a.eps, md.grad(th, k))), A::mul(a.sd, z[k]));
            const T lp_p = md.logp(q);
            T sf = T(0), sr = T(0);
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                if (ln.valid(k)) {
                    T df = A::sub(A::sub(q[k], th[k]), A::mul(a.eps, md.grad(th, k)));
                    T dr = A::sub(A::sub(th[k], q[k]), A::mul(a.eps, md.grad(q, k)));
                    sf = A::add(sf, A::mul(df, df));
                    sr = A::add(sr, A::mul(dr, dr));
                }
            }
            const T fwd = A::mul(a.coef, group_sum<G>(sf));
            const T rev = A::mul(a.coef, group_sum<G>(sr));
            acc = logu < A::add(A::sub(lp_p, lp), A::sub(rev, fwd));
            if (acc) {
#pragma unroll
                for (int k = 0; k < NE; ++k) th[k] = q[k];
                lp = lp_p;
            }
            out_lp = lp;
        } else {
            // metropolis.py:107-135 with proposal normal(loc=theta, scale)
            T q[NE];
#pragma unroll
            for (int k = 0; k < NE; ++k) q[k] = A::add(th[k], A::mul(a.scale, z[k]));
            const T lp_p = md.logp(q);
            T ratio = A::sub(lp_p, lp);
            if (a.hastings) {
                T sf = T(0), sr = T(0);
#pragma unroll
                for (int k = 0; k < NE; ++k) {
                    T df = A::sub(q[k], th[k]), dr = A::sub(th[k], q[k]);
                    sf = A::add(sf, A::mul(df, df));
                    sr = A::add(sr, A::mul(dr, dr));
                }
                const T fwd = A::mul(T(-0.5), group_sum<G>(sf)) / a.s2;
                const T rev = A::mul(T(-0.5), group_sum<G>(sr)) / a.s2;
                ratio = A::add(ratio, A::sub(rev, fwd));
            }
            acc = logu < ratio;
            if (acc) {
#pragma unroll
                for (int k = 0; k < NE; ++k) th[k] = q[k];
                lp = lp_p;
            }
            out_lp = lp;
        }
        if constexpr (MOM) {
            if (t == 0) {
#pragma unroll
                for (int k = 0; k < NE; ++k) m_x0[k] = th[k];     // shift: the first draw of the launch
            }
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                const T d = th[k] - m_x0[k];
                m_s1[k] += d;
                m_s2[k] = fma(d, d, m_s2[k]);
            }
        }
        if (series) {
            // series-major output [C, D, n]: stage SERIES_TILE draws in shared memory, then every (chain, dim)
            // leaves with its 8 consecutive draws as one full 32-byte sector
            const int slot = (int)(t % SERIES_TILE);
            T* row = stage + ((size_t)slot * CPB + (threadIdx.x / G)) * a.D;
#pragma unroll
            for (int k = 0; k < NE; ++k)
                if (ln.valid(k)) row[ln.elem(k)] = th[k];
            if (slot == SERIES_TILE - 1 || t == a.n_draws - 1) {
                __syncthreads();
                const int64_t t0 = t - slot;
                const int64_t chain0 = (int64_t)blockIdx.x * CPB;
                for (int c = 0; c < CPB && chain0 + c < a.C; ++c) {
                    for (int e = threadIdx.x; e < a.D; e += 128) {
                        T* dst = a.draws + ((chain0 + c) * (int64_t)a.D + e) * a.n_draws + t0;
                        const T* src = stage + (size_t)c * a.D + e;
                        if (slot == SERIES_TILE - 1 && sizeof(T) == 4 && (a.n_draws & 3) == 0) {
                            float4 lo4 = make_float4(src[0], src[(size_t)CPB * a.D], src[(size_t)2 * CPB * a.D], src[(size_t)3 * CPB * a.D]);
                            float4 hi4 = make_float4(src[(size_t)4 * CPB * a.D], src[(size_t)5 * CPB * a.D], src[(size_t)6 * CPB * a.D],
                                                     src[(size_t)7 * CPB * a.D]);
                            reinterpret_cast<float4*>(dst)[0] = lo4;
                            reinterpret_cast<float4*>(dst)[1] = hi4;
                        } else {
                            for (int q = 0; q <= slot; ++q) dst[q] = src[(size_t)q * CPB * a.D];
                        }
                    }
                }
                __syncthreads();
            }
        } else if (a.draws && active) {
            ln.store(a.draws + (t * a.C + chain) * (int64_t)a.D, th);
        }
        if (ln.lane == 0 && active) {
            if (a.logp) a.logp[t * a.C + chain] = out_lp;
            if (a.accept) a.accept[t * a.C + chain] = acc ? 1 : 0;
        }
    }
    if (active) ln.store(a.theta + chain * (int64_t)a.D, th);
    if constexpr (MOM) {
        // fold the launch's batch (n_b draws, shifted sums in registers) into the running moments: Chan et al.
        if (active && a.n_draws > 0) {
            const double nb = (double)a.n_draws, n0 = (double)a.mom_n0, n = n0 + nb;
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                if (!ln.valid(k)) continue;
                const int64_t i = chain * (int64_t)a.D + ln.elem(k);
                const double s1 = (double)m_s1[k], s2 = (double)m_s2[k];
                const double mean_b = (double)m_x0[k] + s1 / nb, m2_b = s2 - s1 * s1 / nb;
                if (a.mom_n0 == 0) {
                    a.mom_mean[i] = mean_b;
                    a.mom_m2[i] = m2_b;
                } else {
                    const double mean_a = a.mom_mean[i], delta = mean_b - mean_a;
                    a.mom_mean[i] = mean_a + delta * (nb / n);
                    a.mom_m2[i] = a.mom_m2[i] + m2_b + delta * delta * (n0 * nb / n);
                }
            }
        }
    }
}

template <typename T, int G, int J, int MK, int ALGO, bool MOM>
static int launch_one(const SepArgs<T>& a, unsigned blocks, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024)
        BK_CUDA(cudaFuncSetAttribute(k_sep_sampler<T, G, J, MK, ALGO, MOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sep_sampler<T, G, J, MK, ALGO, MOM><<<blocks, 128, smem, st>>>(a);
    return BK_OK;
}

template <typename T, int G, int J, int MK, int ALGO>
static int launch_mom(const SepArgs<T>& a, unsigned blocks, size_t smem, cudaStream_t st) {
    if (a.mom_mean) {
        if constexpr (sizeof(T) == 4) return launch_one<T, G, J, MK, ALGO, true>(a, blocks, smem, st);
        set_error("streaming moments are fused into the fp32 samplers only");
        return BK_E_UNSUPPORTED;
    }
    return launch_one<T, G, J, MK, ALGO, false>(a, blocks, smem, st);
}

template <typename T, int G, int J, int MK>
static int launch_gj(const SepArgs<T>& a, cudaStream_t st) {
    const int threads = 128;
    const int64_t chains_per_block = threads / G;
    const int64_t blocks = (a.C + chains_per_block - 1) / chains_per_block;
    if (blocks == 0) return BK_OK;
    const size_t smem = (a.layout == BK_DRAWS_CDN && a.draws) ? (size_t)SERIES_TILE * chains_per_block * a.D * sizeof(T) : 0;
    if (smem > 200 * 1024) {
        set_error("series-major draws: staging tile of %zu bytes exceeds shared memory", smem);
        return BK_E_UNSUPPORTED;
    }
    prof_begin(BK_PROF_SAMPLER, st);
    int rc;
    switch (a.algo) {
        case ALGO_HMC: rc = launch_mom<T, G, J, MK, ALGO_HMC>(a, (unsigned)blocks, smem, st); break;
        case ALGO_MALA: rc = launch_mom<T, G, J, MK, ALGO_MALA>(a, (unsigned)blocks, smem, st); break;
        case ALGO_MHRW: rc = launch_mom<T, G, J, MK, ALGO_MHRW>(a, (unsigned)blocks, smem, st); break;
        default: set_error("unknown fused algo %d", a.algo); return BK_E_INVALID;
    }
    if (rc) return rc;
    prof_end(BK_PROF_SAMPLER, st);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

}  // namespace bk
