// Posterior diagnostics on device (fp64 accumulation throughout -- the 1e-9
// contract on ess/rhat cannot be met by an fp32 pipeline):
//   bk_chain_moments      per-series mean / ddof=1 variance      (rhat.py:165-166)
//   bk_rhat_from_moments  potential scale reduction              (rhat.py:111-171)
//   bk_iat_ess            Geyer IPSE / IMSE with on-demand lags  (iat.py:7-135, ess.py:5-69)
//   bk_autocorr           all-lag biased autocorrelation         (autocorr.py:6-33)
//
// The reference evaluates autocorr with a zero-padded FFT of length
// 2^ceil(log2(2N-1)); padding >= 2N-1 makes the circular correlation equal the
// linear one, so acor[k] = sum_t d_t d_{t+k} / var / N exactly.  ess/iat only
// consume lags up to the first negative pair (iat.py:37-43), so the fast path
// computes lag pairs on demand and stops there; bk_autocorr uses the direct
// sum for short series and the four-step FFT (fft_autocorr.cu) for long ones.
#include <stdlib.h>

#include "diag.h"

namespace bk {

constexpr int ACF_THREADS = 256;
constexpr int ACF_LAGS = 8;  // lags evaluated per block-wide round
constexpr int ACF_TT = 4;    // consecutive draws per thread (register tile)
constexpr int ACF_PAD = 16;  // zero padding behind the series in shared memory

// ---- moments -------------------------------------------------------------------
// one-pass shifted sums (shift = first draw) in fp64.
// Neighbouring series adjacent in memory (the samplers' [draws, chains, params] layout): a block
// owns 32 adjacent series; warp w walks draws w, w+8, ... (each load = one 128-byte segment, four
// in flight per thread) and the 8 slices are combined in shared memory in a fixed order.
constexpr int MOM_SLICES = 8;
__global__ void __launch_bounds__(32 * MOM_SLICES) k_moments_thread(SeriesView v, double* __restrict__ mean,
                                                                     double* __restrict__ var) {
    __shared__ double p1[MOM_SLICES][32], p2[MOM_SLICES][32];
    const int sx = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int64_t s = (int64_t)blockIdx.x * 32 + sx;
    const bool ok = s < v.n_series;
    const int64_t ss = ok ? s : v.n_series - 1;
    const double sh = v.at(ss, 0);
    double s1 = 0, s2 = 0;
    int64_t t = sl;
    for (; t + 3 * MOM_SLICES < v.N; t += 4 * MOM_SLICES) {
        const double d0 = v.at(ss, t) - sh, d1 = v.at(ss, t + MOM_SLICES) - sh,
                     d2 = v.at(ss, t + 2 * MOM_SLICES) - sh, d3 = v.at(ss, t + 3 * MOM_SLICES) - sh;
        s1 += (d0 + d1) + (d2 + d3);
        s2 = fma(d0, d0, s2); s2 = fma(d1, d1, s2); s2 = fma(d2, d2, s2); s2 = fma(d3, d3, s2);
    }
    for (; t < v.N; t += MOM_SLICES) {
        const double d = v.at(ss, t) - sh;
        s1 += d;
        s2 = fma(d, d, s2);
    }
    p1[sl][sx] = s1;
    p2[sl][sx] = s2;
    __syncthreads();
    if (sl == 0 && ok) {
        double a = 0, b = 0;
#pragma unroll
        for (int i = 0; i < MOM_SLICES; ++i) { a += p1[i][sx]; b += p2[i][sx]; }
        const double n = (double)v.N;
        if (mean) mean[s] = sh + a / n;
        if (var) var[s] = (b - a * a / n) / (n - 1.0);
    }
}

__global__ void k_moments_warp(SeriesView v, double* __restrict__ mean, double* __restrict__ var) {
    int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (s >= v.n_series) return;
    const double sh = v.at(s, 0);
    double s1 = 0, s2 = 0;
    for (int64_t t = lane; t < v.N; t += 32) {
        double d = v.at(s, t) - sh;
        s1 += d;
        s2 = fma(d, d, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const double n = (double)v.N;
    if (lane == 0) {
        if (mean) mean[s] = sh + s1 / n;
        if (var) var[s] = (s2 - s1 * s1 / n) / (n - 1.0);
    }
}

// Streaming moments: fold a batch of n draws [n, E] (E = chains x dims, the samplers'
// output) into running per-element (mean, M2) over n0 earlier draws -- Chan et al.'s pairwise merge,
// fp64.  One thread per element, coalesced across elements; the accumulators are touched once per
// BATCH (32 B per element), the draws once (s bytes per element-draw).
template <typename T>
__global__ void k_moments_accumulate(const T* __restrict__ draws, int64_t n, int64_t E, int64_t n0,
                                     double* __restrict__ mean, double* __restrict__ m2) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const double sh = (double)draws[e];            // shift by the batch's first draw
    double s1 = 0, s2 = 0;
    for (int64_t t = 0; t < n; ++t) {
        const double d = (double)draws[t * E + e] - sh;
        s1 += d;
        s2 = fma(d, d, s2);
    }
    const double nb = (double)n, mb = sh + s1 / nb, m2b = s2 - s1 * s1 / nb;
    if (n0 == 0) {
        mean[e] = mb;
        m2[e] = m2b;
    } else {
        const double na = (double)n0, nt = na + nb, delta = mb - mean[e];
        mean[e] += delta * (nb / nt);
        m2[e] += m2b + delta * delta * (na * nb / nt);
    }
}

// one warp per parameter; moments laid out [n_chains, n_params]
__global__ void k_rhat(const double* __restrict__ mean, const double* __restrict__ var,
                       const int64_t* __restrict__ lengths, int64_t N, int64_t n_chains,
                       int64_t n_params, double* __restrict__ out) {
    int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (p >= n_params) return;
    double sm = 0, sv = 0, sl = 0;
    for (int64_t m = lane; m < n_chains; m += 32) {
        sm += mean[m * n_params + p];
        sv += var[m * n_params + p];
        sl += lengths ? (double)lengths[m] : (double)N;
    }
    sm = warp_sum(sm); sv = warp_sum(sv); sl = warp_sum(sl);
    const double mm = sm / (double)n_chains;
    double sq = 0;
    for (int64_t m = lane; m < n_chains; m += 32) {
        double d = mean[m * n_params + p] - mm;
        sq = fma(d, d, sq);
    }
    sq = warp_sum(sq);
    if (lane == 0) {
        const double nbar = sl / (double)n_chains;
        const double between = sq / (double)(n_chains - 1);  // var(means, ddof=1)
        const double within = sv / (double)n_chains;         // mean(vars)
        out[p] = sqrt((nbar - 1.0) / nbar + between / within);
    }
}

// ---- direct-lag autocorrelation ---------------------------------------------------
// d: centred series in shared memory; computes raw sums S_k = sum_t d_t d_{t+k}
// for k = k0 .. k0+ACF_LAGS-1 into res[0..ACF_LAGS) (valid in every thread).
__device__ __forceinline__ void lag_round(const double* __restrict__ d, int64_t N, int64_t k0,
                                          double (*part)[ACF_LAGS], double* res) {
    // register tile: ACF_TT consecutive t per thread x ACF_LAGS lags -> 32 FMAs per 15
    // shared loads; d is zero-padded by ACF_PAD entries so no per-element bounds checks
    double acc[ACF_LAGS];
#pragma unroll
    for (int j = 0; j < ACF_LAGS; ++j) acc[j] = 0;
    for (int64_t t0 = (int64_t)threadIdx.x * ACF_TT; t0 + k0 < N; t0 += (int64_t)ACF_THREADS * ACF_TT) {
        double a[ACF_TT], w[ACF_TT + ACF_LAGS - 1];
#pragma unroll
        for (int i = 0; i < ACF_TT; ++i) a[i] = d[t0 + i];
#pragma unroll
        for (int i = 0; i < ACF_TT + ACF_LAGS - 1; ++i) w[i] = d[t0 + k0 + i];
#pragma unroll
        for (int i = 0; i < ACF_TT; ++i)
#pragma unroll
            for (int j = 0; j < ACF_LAGS; ++j) acc[j] = fma(a[i], w[i + j], acc[j]);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < ACF_LAGS; ++j) {
        double s = warp_sum(acc[j]);
        if (lane == 0) part[w][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < ACF_LAGS) {
        double s = 0;
#pragma unroll
        for (int i = 0; i < ACF_THREADS / 32; ++i) s += part[i][threadIdx.x];
        res[threadIdx.x] = s;
    }
    __syncthreads();
}

// mode 0: all lags -> acf_out [S, N];  mode 1: iat/ess
__global__ void __launch_bounds__(ACF_THREADS) k_acf_direct(SeriesView v, int mode, int estimator,
                                                            double* __restrict__ acf_out,
                                                            double* __restrict__ iat_out,
                                                            double* __restrict__ ess_out) {
    extern __shared__ double d[];  // [N]
    __shared__ double red[33];
    __shared__ double part[ACF_THREADS / 32][ACF_LAGS];
    __shared__ double res[ACF_LAGS];
    __shared__ int done;
    const int64_t N = v.N;
    for (int64_t s = blockIdx.x; s < v.n_series; s += gridDim.x) {
        double loc = 0;
        for (int64_t t = threadIdx.x; t < N; t += ACF_THREADS) {
            double x = v.at(s, t);
            d[t] = x;
            loc += x;
        }
        const double mean = block_sum(loc, red) / (double)N;
        loc = 0;
        for (int64_t t = threadIdx.x; t < N; t += ACF_THREADS) {
            double c = d[t] - mean;
            d[t] = c;
            loc = fma(c, c, loc);
        }
        const double var = block_sum(loc, red) / (double)N;  // ddof = 0 (autocorr.py:27)
        const double dn = (double)N;
        if (threadIdx.x < ACF_PAD) d[N + threadIdx.x] = 0.0;
        __syncthreads();
        if (mode == 0) {
            for (int64_t k0 = 0; k0 < N; k0 += ACF_LAGS) {
                lag_round(d, N, k0, part, res);
                if (threadIdx.x < ACF_LAGS && k0 + threadIdx.x < N)
                    acf_out[s * N + k0 + threadIdx.x] = res[threadIdx.x] / var / dn;
                __syncthreads();
            }
        } else {
            // Geyer truncation, pairs (2j, 2j+1) in order (iat.py:37-43, 127-135)
            double total = 0, low = 0;
            const int64_t n_pairs = N / 2;
            if (threadIdx.x == 0) done = 0;
            __syncthreads();
            for (int64_t k0 = 0; k0 < 2 * n_pairs && !done; k0 += ACF_LAGS) {
                lag_round(d, N, k0, part, res);
                if (threadIdx.x == 0) {
                    for (int j = 0; j < ACF_LAGS / 2; ++j) {
                        int64_t pj = k0 / 2 + j;
                        if (pj >= n_pairs) { done = 1; break; }
                        const double pair = res[2 * j] / var / dn + res[2 * j + 1] / var / dn;
                        if (estimator == BK_IAT_IPSE) {
                            if (pair < 0) { done = 1; break; }
                            total += pair;
                        } else {
                            if (pj == 0) { low = pair; total = pair; if (pair < 0) { done = 1; break; } }
                            else {
                                if (pair < 0) { done = 1; break; }
                                low = fmin(low, pair);
                                total += low;
                            }
                        }
                    }
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                const double iat = 2.0 * total - 1.0;
                if (iat_out) iat_out[s] = iat;
                if (ess_out) ess_out[s] = dn / iat;
            }
            __syncthreads();
        }
    }
}

}  // namespace bk

using namespace bk;

static int check_series(const void* x, int32_t dtype, const bk_series_layout* l, const char* who) {
    BK_CHECK_ARG(x && l, "%s: null argument", who);
    BK_CHECK_ARG(l->n_series >= 0 && l->n_inner >= 1, "%s: bad layout", who);
    BK_CHECK_ARG(dtype == BK_F32 || dtype == BK_F64, "%s: bad dtype %d", who, dtype);
    return BK_OK;
}

static SeriesView view_of(const void* x, int32_t dtype, const bk_series_layout* l) {
    return SeriesView{x, dtype, l->n_series, l->n_draws, l->n_inner, l->outer_stride, l->inner_stride,
                      l->draw_stride};
}

namespace bk {
// ---- cross-rank R-hat as an all-reduce of [P, 4] sums (SURVEY 8(e) diagnostics) ---------------
// pass 1 (ref == NULL): out[p] = {n_chains, sum_c mean_cp, sum_c var_cp}               ([P, 3])
// pass 2 (ref = grand means after the first all-reduce): out[p] = sum_c (mean_cp - ref_p)^2   ([P])
// One warp per parameter, chains strided over the lanes, fixed combination order.
__global__ void k_rhat_partial(const double* __restrict__ mean, const double* __restrict__ var, int64_t n_chains,
                               int64_t n_params, const double* __restrict__ ref, double* __restrict__ out) {
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= n_params) return;
    double s1 = 0, s2 = 0, s3 = 0;
    const double r = ref ? ref[p] : 0.0;
    for (int64_t c = lane; c < n_chains; c += 32) {
        const double m = mean[c * n_params + p];
        if (ref) { const double d = m - r; s2 += d * d; }
        else { s1 += m; s3 += var[c * n_params + p]; }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    if (lane == 0) {
        if (ref) out[p] = s2;
        else { out[3 * p] = (double)n_chains; out[3 * p + 1] = s1; out[3 * p + 2] = s3; }
    }
}
// grand mean of the chain means from the reduced sums: ref[p] = sums[p][1] / sums[p][0]
__global__ void k_rhat_ref(const double* __restrict__ sums, int64_t n_params, double* __restrict__ ref) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_params) ref[p] = sums[3 * p + 1] / sums[3 * p];
}
// rhat.py:163-170 from the reduced sums {m, sum mean, sum var} and sum (mean - gm)^2, common length N
// (fewer than two chains in total: NaN -- the per-rank call cannot know the global count up front)
__global__ void k_rhat_from_sums(const double* __restrict__ sums, const double* __restrict__ sqdev, int64_t n_params,
                                 double N, double* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_params) return;
    const double m = sums[3 * p];
    const double var_means = sqdev[p] / (m - 1.0), mean_vars = sums[3 * p + 2] / m;
    out[p] = m >= 2.0 ? sqrt((N - 1.0) / N + var_means / mean_vars) : nan("");
}
}  // namespace bk

extern "C" {

size_t bk_iat_ess_workspace_bytes(int32_t dtype, const bk_series_layout* layout) {
    if (!layout) return 0;
    return ess_stream_ws_bytes(view_of(nullptr, dtype, layout));
}

size_t bk_autocorr_workspace_bytes(int64_t n_series, int64_t N) {
    if (n_series <= 0 || N < 2) return 256;
    return acf_fft_ws_bytes(n_series, N);
}

int bk_chain_moments(const void* x, int32_t dtype, const bk_series_layout* layout, double* mean_out,
                     double* var_out, void* stream) {
    int rc = check_series(x, dtype, layout, "bk_chain_moments");
    if (rc) return rc;
    // rhat.py:159-162: every chain needs >= 2 draws
    BK_CHECK_ARG(layout->n_draws >= 2, "rhat requires len(chain) >= 2 for every chain in chains");
    const int64_t n_series = layout->n_series;
    if (n_series == 0) return BK_OK;
    SeriesView v = view_of(x, dtype, layout);
    cudaStream_t st = (cudaStream_t)stream;
    // neighbouring series adjacent in memory -> one thread per series is coalesced;
    // otherwise a warp walks one series along the draw axis
    if (layout->n_inner > 1 && layout->inner_stride == 1 && layout->draw_stride != 1)
        k_moments_thread<<<(unsigned)((n_series + 31) / 32), 32 * MOM_SLICES, 0, st>>>(v, mean_out, var_out);
    else
        k_moments_warp<<<(unsigned)((n_series * 32 + 255) / 256), 256, 0, st>>>(v, mean_out, var_out);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

int bk_moments_accumulate(const void* draws, int32_t dtype, int64_t n_draws, int64_t n_elems, int64_t n0,
                          double* mean_inout, double* m2_inout, void* stream) {
    BK_CHECK_ARG(draws && mean_inout && m2_inout, "bk_moments_accumulate: null argument");
    BK_CHECK_ARG(dtype == BK_F32 || dtype == BK_F64, "bk_moments_accumulate: bad dtype %d", dtype);
    BK_CHECK_ARG(n_draws >= 0 && n_elems >= 0 && n0 >= 0, "bk_moments_accumulate: negative size");
    if (n_draws == 0 || n_elems == 0) return BK_OK;
    const unsigned blocks = (unsigned)((n_elems + 255) / 256);
    if (dtype == BK_F64)
        k_moments_accumulate<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double*)draws, n_draws, n_elems,
                                                                             n0, mean_inout, m2_inout);
    else
        k_moments_accumulate<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)draws, n_draws, n_elems,
                                                                            n0, mean_inout, m2_inout);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

int bk_rhat_from_moments(const double* mean, const double* var, const int64_t* lengths, int64_t N,
                         int64_t n_chains, int64_t n_params, double* out, void* stream) {
    BK_CHECK_ARG(mean && var && out, "bk_rhat_from_moments: null argument");
    BK_CHECK_ARG(n_chains >= 2, "rhat requires len(chains) >= 2, but len(chains) = %lld",
                 (long long)n_chains);  // rhat.py:157-158
    if (n_params == 0) return BK_OK;
    k_rhat<<<(unsigned)((n_params * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        mean, var, lengths, N, n_chains, n_params, out);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

int bk_rhat_partial_sums(const double* mean, const double* var, int64_t n_chains, int64_t n_params,
                         const double* ref, double* sums, void* stream) {
    BK_CHECK_ARG(mean && sums && (ref || var) && n_chains >= 0, "bk_rhat_partial_sums: bad argument");
    if (n_params == 0) return BK_OK;
    k_rhat_partial<<<(unsigned)((n_params * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mean, var, n_chains,
                                                                                           n_params, ref, sums);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

int bk_rhat_from_sums(const double* sums, const double* sqdev, int64_t n_params, int64_t N, double* ref_out,
                      double* out, void* stream) {
    BK_CHECK_ARG(sums && (ref_out || (out && sqdev)), "bk_rhat_from_sums: bad argument");
    if (n_params == 0) return BK_OK;
    const unsigned blocks = (unsigned)((n_params + 255) / 256);
    if (ref_out) {
        k_rhat_ref<<<blocks, 256, 0, (cudaStream_t)stream>>>(sums, n_params, ref_out);
        BK_LAUNCH_CHECK();
    }
    if (out) {
        BK_CHECK_ARG(N >= 2, "rhat requires len(chain) >= 2 for every chain in chains");
        k_rhat_from_sums<<<blocks, 256, 0, (cudaStream_t)stream>>>(sums, sqdev, n_params, (double)N, out);
        BK_LAUNCH_CHECK();
    }
    return BK_OK;
}

static int acf_launch(const SeriesView& v, int mode, int estimator, double* acf, double* iat, double* ess,
                      cudaStream_t st) {
    const size_t smem = (size_t)(v.N + ACF_PAD) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("series length %lld exceeds the in-SM limit of %d draws", (long long)v.N, 200 * 1024 / 8);
        return BK_E_UNSUPPORTED;
    }
    static bool attr_set = false;
    if (!attr_set) {
        BK_CUDA(cudaFuncSetAttribute(k_acf_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    const unsigned blocks = (unsigned)(v.n_series < 148 * 8 ? v.n_series : 148 * 8);
    k_acf_direct<<<blocks, ACF_THREADS, smem, st>>>(v, mode, estimator, acf, iat, ess);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

int bk_autocorr(const void* x, int32_t dtype, const bk_series_layout* layout, double* out, void* ws,
                size_t ws_bytes, void* stream) {
    int rc = check_series(x, dtype, layout, "bk_autocorr");
    if (rc) return rc;
    BK_CHECK_ARG(layout->n_draws >= 2, "autocorr requires len(chain) >= 2, but len(chain)=%lld",
                 (long long)layout->n_draws);
    BK_CHECK_ARG(out, "bk_autocorr: out is required");
    if (layout->n_series == 0) return BK_OK;
    // short series: O(N^2) direct sums are cheaper than the transform; BK_ACF=direct|fft overrides
    const char* e = getenv("BK_ACF");
    const bool use_fft = e ? (e[0] == 'f') : layout->n_draws > 256;
    if (use_fft)
        return acf_fft_launch(view_of(x, dtype, layout), out, ws, ws_bytes, (cudaStream_t)stream);
    return acf_launch(view_of(x, dtype, layout), 0, 0, out, nullptr, nullptr, (cudaStream_t)stream);
}

int bk_iat_ess(const void* x, int32_t dtype, const bk_series_layout* layout, int32_t estimator,
               double* iat_out, double* ess_out, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_series(x, dtype, layout, "bk_iat_ess");
    if (rc) return rc;
    BK_CHECK_ARG(layout->n_draws >= 4, "iat/ess require len(chain) >= 4, but len(chain)=%lld",
                 (long long)layout->n_draws);
    BK_CHECK_ARG(estimator == BK_IAT_IPSE || estimator == BK_IAT_IMSE, "bk_iat_ess: bad estimator %d",
                 estimator);
    if (layout->n_series == 0) return BK_OK;
    // BK_ESS=block selects the block-per-series kernel (whole series staged in shared memory)
    const char* e = getenv("BK_ESS");
    if (e && e[0] == 'b')
        return acf_launch(view_of(x, dtype, layout), 1, estimator, nullptr, iat_out, ess_out,
                          (cudaStream_t)stream);
    return ess_stream_launch(view_of(x, dtype, layout), estimator, iat_out, ess_out, ws, ws_bytes,
                             (cudaStream_t)stream);
}

}  // extern "C"
