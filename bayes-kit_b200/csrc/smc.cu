// TemperedLikelihoodSMC hot path (smc.py:12-89):
//   bk_smc_move_weight       -- RW-Metropolis move at time(n-1) fused with the
//                               importance log-weights lp_n - lp_{n-1}
//   bk_smc_weight_stats      -- max / sum exp / sum exp^2 (log-sum-exp normaliser, ESS)
//   bk_smc_resample_indices  -- normalise, block scan -> CDF, binary search
//                               (multinomial = np.random.choice; systematic)
//   bk_gather_rows           -- thetas[idx]
#include <stdlib.h>

#include "model.h"
#include "sep_common.cuh"

namespace bk {

template <typename T>
struct SmcArgs {
    T* thetas;                 // [M, D] moved particles (out)
    const T* src;              // particles to move: row src_idx[m] (the pending resample) or row m
    const int64_t* src_idx;    // [M] or NULL
    int64_t M;
    int D, vec, vec2;
    const T *mu, *pl, *m0, *p0;
    T t0, t1, scale;
    bk_rng rng;
    T* logw;
    const T* logw_prev;        // [M] log-weights carried from temperatures without resampling, or NULL
    int32_t* accept;
};

// tempered density lp_t(x) = ll(x) * t + prior(x)  (smc.py:47-51); returns ll and prior
template <typename T, int G, int J>
__device__ __forceinline__ void gpl_terms(const T (&x)[4 * J], const T (&mu)[4 * J], const T (&pl)[4 * J],
                                          const T (&m0)[4 * J], const T (&p0)[4 * J], T& ll, T& pr) {
    using A = Ar<T>;
    T s = T(0), s2 = T(0);
#pragma unroll
    for (int k = 0; k < 4 * J; ++k) {
        T dl = A::sub(x[k], mu[k]), dp = A::sub(x[k], m0[k]);
        s = A::add(s, A::mul(A::mul(pl[k], dl), dl));
        s2 = A::add(s2, A::mul(A::mul(p0[k], dp), dp));
    }
    ll = A::mul(T(-0.5), group_sum<G>(s));
    pr = A::mul(T(-0.5), group_sum<G>(s2));
}

// Persistent groups: a group of G lanes owns elements 4*lane..4*lane+3 (per j) of EVERY particle it
// visits, so the model's per-dimension parameters are loaded once and stay in registers.  The
// resample index and the particle row of the NEXT visit are requested before the current particle is
// processed (two dependent global loads deep), the accept uniform is drawn by the group's first lane only.
template <typename T, int G, int J>
__global__ void __launch_bounds__(128) k_smc_move_weight(SmcArgs<T> a) {
    using A = Ar<T>;
    constexpr int NE = 4 * J;
    const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t n_it = (a.M + n_groups - 1) / n_groups;       // same trip count for every lane of a warp
    Lanes<T, G, J> ln;
    ln.lane = threadIdx.x % G;
    ln.D = a.D;
    ln.vec = a.vec != 0;
    ln.vec2 = a.vec2 != 0;
    T mu[NE], pl[NE], m0[NE], p0[NE];
    ln.load(a.mu, mu, T(0));
    ln.load(a.pl, pl, T(0));
    ln.load(a.m0, m0, T(0));
    ln.load(a.p0, p0, T(0));
    const bool spare_block = a.rng.mode == BK_RNG_PHILOX && 4 * (G * J - 1) >= a.D;
    auto particle = [&](int64_t it) { const int64_t r = g0 + it * n_groups; return r < a.M ? r : a.M - 1; };
    auto row_of = [&](int64_t m) { return a.src_idx ? a.src_idx[m] : m; };
    T nx[NE];                                  // particle of the next visit
    int64_t row_nn = 0;                        // resample index of the visit after that
    if (n_it > 0) ln.load(a.src + row_of(particle(0)) * (int64_t)a.D, nx, T(0));
    if (n_it > 1) row_nn = row_of(particle(1));
    for (int64_t it = 0; it < n_it; ++it) {
        const int64_t raw = g0 + it * n_groups;
        const bool active = raw < a.M;
        const int64_t m = active ? raw : a.M - 1;
        T th[NE], z[NE], st[NE];
#pragma unroll
        for (int k = 0; k < NE; ++k) th[k] = nx[k];
        // thetas[idxs] of the previous importance_resample (smc.py:75) is folded into this read
        if (it + 1 < n_it) ln.load(a.src + row_nn * (int64_t)a.D, nx, T(0));
        if (it + 2 < n_it) row_nn = row_of(particle(it + 2));
        uint32_t raw2[2] = {0u, 0u};
        ln.normals(a.rng, a.M, m, 0, z, raw2);
#pragma unroll
        for (int k = 0; k < NE; ++k) st[k] = A::add(th[k], A::mul(a.scale, z[k]));  // smc.py:81
        T ll_c, pr_c, ll_s, pr_s;
        gpl_terms<T, G, J>(th, mu, pl, m0, p0, ll_c, pr_c);
        gpl_terms<T, G, J>(st, mu, pl, m0, p0, ll_s, pr_s);
        const T lp_c = A::add(A::mul(ll_c, a.t0), pr_c), lp_s = A::add(A::mul(ll_s, a.t0), pr_s);
        // accept uniform: when the group's last element block is pure padding (4 (G J - 1) >= D), its
        // Philox words are unused by the proposal and serve as the uniform -- one counter-mode call per
        // lane and particle instead of two.  Otherwise (and for injected streams) the dedicated stream.
        T lu = T(0);
        if (spare_block) {
            if (ln.lane == G - 1) {
                if constexpr (sizeof(T) == 4) lu = log_u(u01(raw2[0]));
                else lu = log_u(u01d(raw2[0], raw2[1]));
            }
            lu = __shfl_sync(0xffffffffu, lu, G - 1, G);
        } else {
            if (ln.lane == 0) lu = log_u(ln.uniform(a.rng, a.M, m, 0, 0));
            lu = __shfl_sync(0xffffffffu, lu, 0, G);
        }
        const bool acc = lu < A::sub(lp_s, lp_c);  // smc.py:85
        if (acc) {
#pragma unroll
            for (int k = 0; k < NE; ++k) th[k] = st[k];
            ll_c = ll_s;
            pr_c = pr_s;
        }
        // importance log-weight, both tempered densities in full (smc.py:67-70)
        const T lw = A::sub(A::add(A::mul(ll_c, a.t1), pr_c), A::add(A::mul(ll_c, a.t0), pr_c));
        if (active) {
            ln.store(a.thetas + m * (int64_t)a.D, th);
            if (ln.lane == 0) {
                a.logw[m] = a.logw_prev ? A::add(a.logw_prev[m], lw) : lw;
                if (a.accept) a.accept[m] = acc ? 1 : 0;
            }
        }
    }
}

template <typename T, int G, int J>
static int launch_smc(const SmcArgs<T>& a, cudaStream_t st) {
    const int64_t per_block = 128 / G;
    const int64_t need = (a.M + per_block - 1) / per_block;
    // persistent: exactly one wave -- as many CTAs of 128 threads as are resident at once (8 per SM at 62
    // registers, 3 for the 16-elements-per-lane layout)
    static int64_t cap = 0;
    if (!cap) {
        int dev = 0, sms = 148, nb = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_smc_move_weight<T, G, J>, 128, 0) != cudaSuccess || nb < 1) {
            cudaGetLastError();
            nb = 1;
        }
        cap = (int64_t)sms * (nb > 8 ? 8 : nb);
    }
    k_smc_move_weight<T, G, J><<<(unsigned)(need < cap ? need : cap), 128, 0, st>>>(a);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

template <typename T>
static int smc_move_t(const Model& m, const void* src, const int64_t* src_idx, void* thetas, int64_t M,
                      int n, int Tn, double scale, const bk_rng* rng, void* logw, const void* logw_prev,
                      int32_t* accept, cudaStream_t st) {
    SmcArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.thetas = (T*)thetas;
    a.src = (const T*)src;
    a.src_idx = src_idx;
    a.M = M;
    a.D = (int)m.d.dims;
    a.mu = (const T*)m.d.mu; a.pl = (const T*)m.d.prec; a.m0 = (const T*)m.d.m0; a.p0 = (const T*)m.d.p0;
    a.t0 = (T)((double)(n - 1) / Tn);  // smc.py:43-44
    a.t1 = (T)((double)n / Tn);
    a.scale = (T)scale;
    a.rng = *rng;
    a.logw = (T*)logw;
    a.logw_prev = (const T*)logw_prev;
    a.accept = accept;
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    a.vec = (a.D % 4 == 0 && al(thetas) && al(src) && al(a.mu) && al(a.pl) && al(a.m0) && al(a.p0) &&
             (rng->mode != BK_RNG_INJECTED || al(rng->normals))) ? 1 : 0;
    auto al8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; };
    // D even but not a multiple of 4 (c4: D = 50): rows are 8-byte aligned -> 64-bit accesses
    a.vec2 = (!a.vec && sizeof(T) == 4 && a.D % 2 == 0 && al8(thetas) && al8(src) && al8(a.mu) && al8(a.pl) &&
              al8(a.m0) && al8(a.p0) && (rng->mode != BK_RNG_INJECTED || al8(rng->normals))) ? 1 : 0;
    const int D = a.D;
    if (D <= 4) return launch_smc<T, 1, 1>(a, st);
    if (D <= 16) return launch_smc<T, 4, 1>(a, st);
    if (D <= 32) return launch_smc<T, 8, 1>(a, st);
    // fp32 (timed mode): 4 lanes x 16 elements instead of 16 x 4 -- the per-lane overhead of a particle visit is
    // paid by a quarter of the lanes; G * J (Philox blocks, spare-block uniform rule) is unchanged.  BK_SEP_WIDE=0: off
    static int wide = -1;
    if (wide < 0) { const char* e = getenv("BK_SEP_WIDE"); wide = (e && e[0] == '0') ? 0 : 1; }
    if (sizeof(T) == 4 && wide && D > 32 && D <= 64) return launch_smc<T, 4, 4>(a, st);
    if (D <= 64) return launch_smc<T, 16, 1>(a, st);
    if (D <= 128) return launch_smc<T, 32, 1>(a, st);
    if (D <= 256) return launch_smc<T, 32, 2>(a, st);
    set_error("bk_smc_move_weight supports D <= 256 (got %d)", D);
    return BK_E_UNSUPPORTED;
}

// ---- weight statistics ---------------------------------------------------------
constexpr int RED_THREADS = 256;

template <typename T>
__global__ void k_max_partial(const T* __restrict__ x, int64_t n, double* __restrict__ part) {
    __shared__ double sm[33];
    double v = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v = fmax(v, (double)x[i]);
    v = warp_max(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < (blockDim.x >> 5) ? sm[lane] : -INFINITY;
        v = warp_max(v);
        if (lane == 0) part[blockIdx.x] = v;
    }
}

// single block: out[0] = max over the block partials (0 when the weights are not shifted)
__global__ void k_max_final(const double* __restrict__ maxpart, int nb, int use_shift, double* __restrict__ out) {
    __shared__ double sm[33];
    double v = -INFINITY;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) v = fmax(v, maxpart[i]);
    v = warp_max(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < (blockDim.x >> 5) ? sm[lane] : -INFINITY;
        v = warp_max(v);
        if (lane == 0) out[0] = use_shift ? v : 0.0;
    }
}

// per-block sums of exp(x - shift) and its square, shift = stats[0]
template <typename T>
__global__ void k_sum_partial(const T* __restrict__ x, int64_t n, const double* __restrict__ stats,
                              double* __restrict__ part) {
    __shared__ double sm[33];
    const double shift = stats[0];
    double s = 0, s2 = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double w = exp((double)x[i] - shift);
        s += w;
        s2 += w * w;
    }
    s = block_sum(s, sm);
    s2 = block_sum(s2, sm);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = s;
        part[2 * blockIdx.x + 1] = s2;
    }
}

// single block, fixed order: out[1] = sum w, out[2] = sum w^2
__global__ void k_stats_final(const double* __restrict__ part, int nb, double* __restrict__ out) {
    __shared__ double sm[33];
    double s = 0, s2 = 0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        s += part[2 * i];
        s2 += part[2 * i + 1];
    }
    s = block_sum(s, sm);
    s2 = block_sum(s2, sm);
    if (threadIdx.x == 0) {
        out[1] = s;
        out[2] = s2;
    }
}

// ---- CDF: p_i = exp(logw_i - shift) / total ; inclusive scan ; cdf /= cdf[-1] ----
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 4, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_local(const T* __restrict__ logw, int64_t n,
                                                             double shift, double total,
                                                             const double* __restrict__ stats, int mode,
                                                             double* __restrict__ cdf,
                                                             double* __restrict__ block_tot) {
    __shared__ double wsum[SCAN_THREADS / 32];
    if (stats) {  // normaliser produced on device by bk_smc_weight_stats
        shift = stats[0];
        total = mode == BK_RESAMPLE_MULTINOMIAL ? stats[1] : 1.0;
    }
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    double v[SCAN_ITEMS];
    double run = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t j = base + i;
        double w = j < n ? exp((double)logw[j] - shift) / total : 0.0;
        run += w;
        v[i] = run;
    }
    // inclusive scan of per-thread totals
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    double woff = 0;
    for (int i = 0; i < w; ++i) woff += wsum[i];
    const double excl = woff + inc - run;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t j = base + i;
        if (j < n) cdf[j] = excl + v[i];
    }
    if (threadIdx.x == SCAN_THREADS - 1) block_tot[blockIdx.x] = excl + run;
}

// exclusive scan of the block totals: one CTA of 1024 threads walks chunks of 1024 totals
__global__ void __launch_bounds__(1024) k_scan_blocks(double* __restrict__ block_tot, int nb,
                                                      double* __restrict__ grand) {
    __shared__ double wsum[32];
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const double v = i < nb ? block_tot[i] : 0.0;
        double inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        double woff = 0;
        for (int k = 0; k < w; ++k) woff += wsum[k];
        const double carry = carry_s;
        if (i < nb) block_tot[i] = carry + woff + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand = carry_s;
}

__global__ void k_scan_finish(double* __restrict__ cdf, int64_t n, const double* __restrict__ block_off,
                              const double* __restrict__ grand) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    // cdf /= cdf[-1]  (legacy choice)
    cdf[j] = (cdf[j] + block_off[j / SCAN_TILE]) / *grand;
}

template <typename T>
__global__ void k_search(const double* __restrict__ cdf, int64_t M, int mode, const T* __restrict__ u_in,
                         bk_rng rng, int64_t n_points, int64_t point_offset,
                         int64_t* __restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    double u;
    if (mode == BK_RESAMPLE_MULTINOMIAL) {
        u = u_in ? (double)u_in[i]
                 : (double)philox_uniform<T>(rng.seed, 0, (uint32_t)(point_offset + i),
                                             (uint32_t)rng.draw_offset, TAG_RESAMPLE);
    } else {
        double u0 = u_in ? (double)u_in[0]
                         : (double)philox_uniform<T>(rng.seed, 0, 0u, (uint32_t)rng.draw_offset, TAG_RESAMPLE);
        u = ((double)(point_offset + i) + u0) / (double)M;
    }
    // first index with cdf[idx] > u  (searchsorted side='right')
    int64_t lo = 0, hi = M;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1;
        else hi = mid;
    }
    idx[i] = lo < M ? lo : M - 1;
}

template <typename T>
__global__ void k_gather(const T* __restrict__ src, const int64_t* __restrict__ idx, int64_t M, int D,
                         T* __restrict__ out) {
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= M) return;
    const T* s = src + idx[r] * (int64_t)D;
    T* o = out + r * (int64_t)D;
    for (int e = lane; e < D; e += 32) o[e] = s[e];
}

// Adaptive resampling: resample only when the importance-weight ESS (sum w)^2 / sum w^2 falls below
// the threshold; otherwise the particles stay (identity indices) and keep their log-weights.  The
// decision is taken on device from the stats bk_smc_weight_stats wrote -- no host round trip.
template <typename T>
__global__ void k_smc_adaptive_select(const double* __restrict__ stats, double thr, int64_t n, int64_t point_offset,
                                      int64_t* __restrict__ idx, T* __restrict__ logw, int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool resample = stats[1] * stats[1] / stats[2] < thr;
    if (i == 0 && flag) *flag = resample ? 1 : 0;
    if (i >= n) return;
    if (resample) logw[i] = T(0);            // equal weights after resampling
    else idx[i] = point_offset + i;          // keep particle i and its accumulated log-weight
}

static int stat_blocks(int64_t M) {
    int64_t b = (M + RED_THREADS - 1) / RED_THREADS;
    return (int)(b < 1 ? 1 : (b > 1184 ? 1184 : b));  // 8 x 148 SMs
}

}  // namespace bk

using namespace bk;

extern "C" {

int bk_smc_move_weight(uint64_t handle, void* thetas, int64_t M, int32_t n, int32_t T, double scale,
                       const bk_rng* rng, void* logw_out, int32_t* accept_out, void* stream) {
    return bk_smc_gather_move_weight(handle, thetas, nullptr, thetas, M, n, T, scale, rng, logw_out,
                                     accept_out, stream);
}

int bk_smc_gather_move_weight(uint64_t handle, const void* src, const int64_t* src_idx, void* thetas,
                              int64_t M, int32_t n, int32_t T, double scale, const bk_rng* rng,
                              void* logw_out, int32_t* accept_out, void* stream) {
    return bk_smc_gather_move_weight_acc(handle, src, src_idx, thetas, M, n, T, scale, rng, nullptr, logw_out,
                                         accept_out, stream);
}

int bk_smc_gather_move_weight_acc(uint64_t handle, const void* src, const int64_t* src_idx, void* thetas,
                                  int64_t M, int32_t n, int32_t T, double scale, const bk_rng* rng,
                                  const void* logw_prev, void* logw_out, int32_t* accept_out, void* stream) {
    const Model* m = get_model(handle);
    if (!m) return BK_E_HANDLE;
    BK_CHECK_ARG(m->d.kind == BK_MODEL_GAUSS_PRIOR_LIK,
                 "bk_smc_move_weight: model must be GAUSS_PRIOR_LIK (log_prior/log_likelihood)");
    BK_CHECK_ARG(src && thetas && logw_out && rng && M >= 0, "bk_smc_move_weight: bad argument");
    BK_CHECK_ARG(!src_idx || src != thetas,
                 "bk_smc_gather_move_weight: a gathered move cannot run in place (src == thetas)");
    BK_CHECK_ARG(T >= 1 && n >= 1 && n <= T, "bk_smc_move_weight: need 1 <= n <= T (n=%d, T=%d)", n, T);
    BK_CHECK_ARG(rng->mode != BK_RNG_INJECTED || (rng->normals && rng->uniforms && rng->n_uniform >= 1),
                 "bk_smc_move_weight: injected rng needs normals/uniforms");
    if (M == 0) return BK_OK;
    if (m->d.dtype == BK_F64)
        return smc_move_t<double>(*m, src, src_idx, thetas, M, n, T, scale, rng, logw_out, logw_prev, accept_out,
                                  (cudaStream_t)stream);
    return smc_move_t<float>(*m, src, src_idx, thetas, M, n, T, scale, rng, logw_out, logw_prev, accept_out,
                             (cudaStream_t)stream);
}

size_t bk_smc_resample_workspace_bytes(int64_t M) {
    if (M <= 0) return 256;
    size_t nb_scan = (size_t)((M + SCAN_TILE - 1) / SCAN_TILE);
    // max partials, sum partials, block totals + grand, cdf
    return align_up(1184 * 8, 256) + align_up(2 * 1184 * 8, 256) + align_up((nb_scan + 1) * 8, 256) +
           align_up((size_t)M * 8, 256) + 1024;
}

int bk_smc_weight_stats(const void* logw, int64_t M, int32_t dtype, int32_t mode, double* stats_out,
                        void* ws, size_t ws_bytes, void* stream) {
    BK_CHECK_ARG(logw && stats_out && M >= 1, "bk_smc_weight_stats: bad argument");
    Arena ar(ws, ws_bytes);
    double* maxp = ar.take<double>(1184);
    double* sump = ar.take<double>(2 * 1184);
    if (!ar.ok()) { set_error("bk_smc_weight_stats: workspace too small"); return BK_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = stat_blocks(M);
    const int shift = mode == BK_RESAMPLE_SYSTEMATIC;
    if (dtype == BK_F64) k_max_partial<double><<<nb, RED_THREADS, 0, st>>>((const double*)logw, M, maxp);
    else k_max_partial<float><<<nb, RED_THREADS, 0, st>>>((const float*)logw, M, maxp);
    BK_LAUNCH_CHECK();
    k_max_final<<<1, 256, 0, st>>>(maxp, nb, shift, stats_out);
    BK_LAUNCH_CHECK();
    if (dtype == BK_F64) k_sum_partial<double><<<nb, RED_THREADS, 0, st>>>((const double*)logw, M, stats_out, sump);
    else k_sum_partial<float><<<nb, RED_THREADS, 0, st>>>((const float*)logw, M, stats_out, sump);
    BK_LAUNCH_CHECK();
    k_stats_final<<<1, 256, 0, st>>>(sump, nb, stats_out);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

static int resample_impl(const void* logw, int64_t M, int32_t dtype, int32_t mode, double shift, double total,
                         const double* stats, const void* uniforms, const bk_rng* rng, int64_t n_points,
                         int64_t point_offset, int64_t* idx_out, void* cdf_out, void* ws, size_t ws_bytes,
                         void* stream) {
    BK_CHECK_ARG(logw && idx_out && M >= 1, "bk_smc_resample_indices: bad argument");
    BK_CHECK_ARG(mode == BK_RESAMPLE_MULTINOMIAL || mode == BK_RESAMPLE_SYSTEMATIC,
                 "bk_smc_resample_indices: bad mode %d", mode);
    BK_CHECK_ARG(uniforms || rng, "bk_smc_resample_indices: need uniforms or rng");
    BK_CHECK_ARG(stats || total > 0, "bk_smc_resample_indices: total weight must be > 0 (got %g)", total);
    BK_CHECK_ARG(n_points >= 0 && point_offset >= 0 && point_offset + n_points <= M,
                 "bk_smc_resample_indices: point range [%lld, %lld) outside [0, %lld)",
                 (long long)point_offset, (long long)(point_offset + n_points), (long long)M);
    Arena ar(ws, ws_bytes);
    ar.take<double>(1184);
    ar.take<double>(2 * 1184);
    const int nb = (int)((M + SCAN_TILE - 1) / SCAN_TILE);
    double* btot = ar.take<double>(nb + 1);
    double* cdf = cdf_out ? (double*)cdf_out : ar.take<double>(M);
    if (!ar.ok()) { set_error("bk_smc_resample_indices: workspace too small"); return BK_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == BK_F64)
        k_scan_local<double><<<nb, SCAN_THREADS, 0, st>>>((const double*)logw, M, shift, total, stats, mode,
                                                           cdf, btot);
    else
        k_scan_local<float><<<nb, SCAN_THREADS, 0, st>>>((const float*)logw, M, shift, total, stats, mode, cdf,
                                                          btot);
    BK_LAUNCH_CHECK();
    k_scan_blocks<<<1, 1024, 0, st>>>(btot, nb, btot + nb);
    BK_LAUNCH_CHECK();
    k_scan_finish<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(cdf, M, btot, btot + nb);
    BK_LAUNCH_CHECK();
    bk_rng r;
    memset(&r, 0, sizeof(r));
    if (rng) r = *rng;
    if (n_points > 0) {
        const unsigned sb = (unsigned)((n_points + 255) / 256);
        if (dtype == BK_F64)
            k_search<double><<<sb, 256, 0, st>>>(cdf, M, mode, (const double*)uniforms, r, n_points,
                                                  point_offset, idx_out);
        else
            k_search<float><<<sb, 256, 0, st>>>(cdf, M, mode, (const float*)uniforms, r, n_points,
                                                 point_offset, idx_out);
        BK_LAUNCH_CHECK();
    }
    return BK_OK;
}

int bk_smc_resample_indices(const void* logw, int64_t M, int32_t dtype, int32_t mode, double shift,
                            double total, const void* uniforms, const bk_rng* rng, int64_t n_points,
                            int64_t point_offset, int64_t* idx_out, void* cdf_out, void* ws,
                            size_t ws_bytes, void* stream) {
    return resample_impl(logw, M, dtype, mode, shift, total, nullptr, uniforms, rng, n_points, point_offset,
                         idx_out, cdf_out, ws, ws_bytes, stream);
}

int bk_smc_resample_indices_dev(const void* logw, int64_t M, int32_t dtype, int32_t mode, const double* stats,
                                const void* uniforms, const bk_rng* rng, int64_t n_points,
                                int64_t point_offset, int64_t* idx_out, void* cdf_out, void* ws,
                                size_t ws_bytes, void* stream) {
    BK_CHECK_ARG(stats, "bk_smc_resample_indices_dev: stats is required");
    return resample_impl(logw, M, dtype, mode, 0.0, 1.0, stats, uniforms, rng, n_points, point_offset, idx_out,
                         cdf_out, ws, ws_bytes, stream);
}

int bk_smc_adaptive_select(const double* stats, double ess_threshold, int64_t n_points, int64_t point_offset,
                           int64_t* idx_inout, void* logw_inout, int32_t dtype, int32_t* resampled_out, void* stream) {
    BK_CHECK_ARG(stats && idx_inout && logw_inout && n_points >= 0, "bk_smc_adaptive_select: bad argument");
    const unsigned blocks = (unsigned)((n_points + 255) / 256 > 0 ? (n_points + 255) / 256 : 1);
    if (dtype == BK_F64)
        k_smc_adaptive_select<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(stats, ess_threshold, n_points, point_offset,
                                                                               idx_inout, (double*)logw_inout, resampled_out);
    else
        k_smc_adaptive_select<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(stats, ess_threshold, n_points, point_offset,
                                                                              idx_inout, (float*)logw_inout, resampled_out);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

int bk_gather_rows(const void* src, const int64_t* idx, int64_t M, int64_t D, int32_t dtype, void* out,
                   void* stream) {
    BK_CHECK_ARG(src && idx && out && M >= 0 && D >= 1, "bk_gather_rows: bad argument");
    if (M == 0) return BK_OK;
    const unsigned blocks = (unsigned)((M * 32 + 255) / 256);
    if (dtype == BK_F64)
        k_gather<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double*)src, idx, M, (int)D, (double*)out);
    else
        k_gather<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)src, idx, M, (int)D, (float*)out);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

}  // extern "C"
