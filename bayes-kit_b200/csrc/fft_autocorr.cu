// Hand-written (cuFFT-free) FFT autocorrelation, fp64 -- autocorr.py:6-33.
//
// Per series: demean, zero-pad to S = 2^ceil(log2(2N-1)) (autocorr.py:26),
// forward FFT, |.|^2, inverse FFT, scale by 1/(var*N) (autocorr.py:27-32).
//
// Design
//  * TWO real series share one complex transform: z = d1 + i d2.  After the forward
//    FFT the spectra separate as F1[k] = (Z[k] + conj Z[S-k]) / 2 and
//    F2[k] = (Z[k] - conj Z[S-k]) / (2i); the power spectra are real and even, so the
//    inverse FFT of |F1|^2 + i |F2|^2 returns acf1 in its real and acf2 in its
//    imaginary part.  One forward + one inverse complex FFT per PAIR of series.
//  * RADIX-32 passes, IN PLACE on one S-element line of per-CTA scratch (512 KB at
//    S = 32768; 148 lines = 74 MB stay L2-resident): decimation in frequency forward
//    (natural in, digit-reversed out), the power spectrum is formed in the digit-reversed
//    domain (partner of position i is pos(S - k(i)), a three-digit index computation),
//    decimation in time back (digit-reversed in, natural out) -- no reordering pass.
//    S = 32768 is three passes each way.  A thread owns one radix-32 butterfly: 32 loads,
//    a 32-point DFT in registers (five fully unrolled radix-2 stages, constant
//    twiddles), twiddles exp(-2 pi i p k / n) from an L2-resident table, 32 stores.
//  * The first forward pass reads the series directly (demean, pack, zero-pad on the
//    fly); the last inverse pass writes only the N wanted lags, scaled.
// A CTA (256 threads, persistent) processes one pair at a time.
#include "diag.h"

namespace bk {

constexpr int FFT_THREADS = 256;

struct __align__(16) cplx { double x, y; };
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cplx{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cplx{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cplx{fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)};
}

// W[j] = exp(-2 pi i j / S), j < S
__global__ void k_fft_twiddles(cplx* __restrict__ W, int64_t S) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S) return;
    double s, c;
    sincospi(-2.0 * (double)j / (double)S, &s, &c);
    W[j] = cplx{c, s};
}

// exp(-2 pi i j / 32), j < 16 (forward); the inverse conjugates
__device__ __constant__ double C32[16] = {1.0,
                                          0.98078528040323044913,
                                          0.92387953251128675613,
                                          0.83146961230254523708,
                                          0.70710678118654752440,
                                          0.55557023301960222474,
                                          0.38268343236508977173,
                                          0.19509032201612826785,
                                          0.0,
                                          -0.19509032201612826785,
                                          -0.38268343236508977173,
                                          -0.55557023301960222474,
                                          -0.70710678118654752440,
                                          -0.83146961230254523708,
                                          -0.92387953251128675613,
                                          -0.98078528040323044913};
__device__ __constant__ double S32[16] = {0.0,
                                          0.19509032201612826785,
                                          0.38268343236508977173,
                                          0.55557023301960222474,
                                          0.70710678118654752440,
                                          0.83146961230254523708,
                                          0.92387953251128675613,
                                          0.98078528040323044913,
                                          1.0,
                                          0.98078528040323044913,
                                          0.92387953251128675613,
                                          0.83146961230254523708,
                                          0.70710678118654752440,
                                          0.55557023301960222474,
                                          0.38268343236508977173,
                                          0.19509032201612826785};

// In-register R-point DFT (R = 2..32, power of two): decimation in frequency, output element k
// ends up in a[bitrev(k)].  INV conjugates the twiddles (no scaling).
template <int R, bool INV>
__device__ __forceinline__ void dft_reg(cplx (&a)[R]) {
#pragma unroll
    for (int h = R / 2; h >= 1; h >>= 1) {
#pragma unroll
        for (int b = 0; b < R; b += 2 * h) {
#pragma unroll
            for (int j = 0; j < h; ++j) {
                const cplx u = a[b + j], v = a[b + j + h];
                a[b + j] = cadd(u, v);
                const cplx d = csub(u, v);
                const int tw = j * (16 / h);          // exponent of exp(-2 pi i / 32): j * (32 / (2h))
                if (tw == 0) {
                    a[b + j + h] = d;
                } else if (tw == 8) {                  // -i (forward) / +i (inverse)
                    a[b + j + h] = INV ? cplx{-d.y, d.x} : cplx{d.y, -d.x};
                } else {
                    const double c = C32[tw], s = INV ? S32[tw] : -S32[tw];
                    a[b + j + h] = cplx{fma(d.x, c, -d.y * s), fma(d.x, s, d.y * c)};
                }
            }
        }
    }
}
template <int R>
__host__ __device__ constexpr int bitrev(int k) {
    int r = 0;
    for (int b = 1, m = R >> 1; b < R; b <<= 1, m >>= 1)
        if (k & b) r |= m;
    return r;
}

// One in-place radix-R pass over S elements with current block length n (m = n / R):
//   forward (decimation in frequency):  x[b + p + k m] = w_n^{p k} sum_j x[b + p + j m] e^{-2 pi i j k / R}
//   inverse (decimation in time):       x[b + p + k m] = sum_j conj(w_n^{p j}) x[b + p + j m] e^{+2 pi i j k / R}
// for every block start b (multiple of n) and p < m; thread t: p = t mod m, b = (t / m) n.
// `ld(i)` yields input element i (only the first forward pass reads elsewhere), `st(i, v)` consumes
// output element i (only the last inverse pass writes elsewhere).
template <int R, bool INV, class Load, class Store>
__device__ __forceinline__ void radix_pass(int64_t S, int log2m, const cplx* __restrict__ W, Load ld, Store st) {
    const int64_t cnt = S / R, m = (int64_t)1 << log2m;
    int log2n = log2m, r = R;
    while (r > 1) { ++log2n; r >>= 1; }
    int log2S = 0;
    while (((int64_t)1 << log2S) < S) ++log2S;
    const int tshift = log2S - log2n;                  // twiddle index of w_n^{e} is e << tshift
    for (int64_t t = threadIdx.x; t < cnt; t += FFT_THREADS) {
        const int64_t p = t & (m - 1), base = ((t >> log2m) << log2n) + p;
        // w_n^p from the table, its powers by a running product (one table sector per butterfly
        // instead of R - 1 scattered ones; the rounding error grows by ~R ulp, far inside 1e-9)
        cplx w1 = W[p << tshift];
        if (INV) w1.y = -w1.y;
        cplx a[R];
        cplx wk = w1;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            a[j] = ld(base + ((int64_t)j << log2m));
            if (INV && j > 0) {
                a[j] = cmul(a[j], wk);
                wk = cmul(wk, w1);
            }
        }
        dft_reg<R, INV>(a);
        wk = w1;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            cplx v = a[bitrev<R>(k)];
            if (!INV && k > 0) {
                v = cmul(v, wk);
                wk = cmul(wk, w1);
            }
            st(base + ((int64_t)k << log2m), v);
        }
    }
    __syncthreads();
}

template <bool INV, class Load, class Store>
__device__ __forceinline__ void pass_r(int log2r, int64_t S, int log2m, const cplx* __restrict__ W, Load ld, Store st) {
    switch (log2r) {
        case 5: radix_pass<32, INV>(S, log2m, W, ld, st); break;
        case 4: radix_pass<16, INV>(S, log2m, W, ld, st); break;
        case 3: radix_pass<8, INV>(S, log2m, W, ld, st); break;
        case 2: radix_pass<4, INV>(S, log2m, W, ld, st); break;
        default: radix_pass<2, INV>(S, log2m, W, ld, st); break;
    }
}

struct FftPlan {
    int n_pass, log2S;
    int log2r[4];     // radix of forward pass i is 1 << log2r[i]
};

// position of frequency k after the forward (DIF) passes: digit reversal in the plan's radices
__device__ __forceinline__ int64_t pos_of_freq(const FftPlan& pl, int64_t k) {
    int64_t i = 0;
    int lm = pl.log2S;
    for (int ps = 0; ps < pl.n_pass; ++ps) {
        lm -= pl.log2r[ps];
        i += (k & (((int64_t)1 << pl.log2r[ps]) - 1)) << lm;
        k >>= pl.log2r[ps];
    }
    return i;
}

__global__ void __launch_bounds__(FFT_THREADS, 1)
k_acf_fft(SeriesView v, int64_t S, FftPlan plan, const cplx* __restrict__ W, cplx* __restrict__ scratch_all,
          double* __restrict__ out) {
    __shared__ double red[33];
    cplx* buf = scratch_all + (int64_t)blockIdx.x * S;
    const int64_t N = v.N, n_pairs = (v.n_series + 1) / 2;
    const double dn = (double)N;

    for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
        const int64_t s1 = 2 * pr, s2 = 2 * pr + 1;
        const bool two = s2 < v.n_series;
        // means / variances (ddof = 0, autocorr.py:27-28)
        double l1 = 0, l2 = 0;
        for (int64_t t = threadIdx.x; t < N; t += FFT_THREADS) {
            l1 += v.at(s1, t);
            if (two) l2 += v.at(s2, t);
        }
        const double m1 = block_sum(l1, red) / dn, m2 = block_sum(l2, red) / dn;
        l1 = l2 = 0;
        for (int64_t t = threadIdx.x; t < N; t += FFT_THREADS) {
            const double c1 = v.at(s1, t) - m1;
            l1 = fma(c1, c1, l1);
            if (two) {
                const double c2 = v.at(s2, t) - m2;
                l2 = fma(c2, c2, l2);
            }
        }
        const double var1 = block_sum(l1, red) / dn, var2 = block_sum(l2, red) / dn;

        auto ld_buf = [&](int64_t i) { return buf[i]; };
        auto st_buf = [&](int64_t i, cplx val) { buf[i] = val; };
        // ---- forward (DIF): z = d1 + i d2, natural -> digit-reversed ----------------------
        int log2m = plan.log2S;
        for (int ps = 0; ps < plan.n_pass; ++ps) {
            log2m -= plan.log2r[ps];
            if (ps == 0) {
                auto ld = [&](int64_t i) {
                    return i < N ? cplx{v.at(s1, i) - m1, two ? v.at(s2, i) - m2 : 0.0} : cplx{0.0, 0.0};
                };
                pass_r<false>(plan.log2r[ps], S, log2m, W, ld, st_buf);
            } else {
                pass_r<false>(plan.log2r[ps], S, log2m, W, ld_buf, st_buf);
            }
        }
        // ---- power spectra |F1|^2 + i |F2|^2, in place: each {k, S-k} pair by one thread ----
        for (int64_t k = threadIdx.x; k <= S / 2; k += FFT_THREADS) {
            const int64_t i = pos_of_freq(plan, k), i2 = pos_of_freq(plan, (S - k) & (S - 1));
            const cplx a = buf[i], b = buf[i2];
            const double f1x = 0.5 * (a.x + b.x), f1y = 0.5 * (a.y - b.y);     // (Z[k] + conj Z[S-k]) / 2
            const double f2x = 0.5 * (a.y + b.y), f2y = 0.5 * (b.x - a.x);     // (Z[k] - conj Z[S-k]) / 2i
            const cplx pw = cplx{fma(f1x, f1x, f1y * f1y), fma(f2x, f2x, f2y * f2y)};
            buf[i] = pw;
            buf[i2] = pw;
        }
        __syncthreads();
        // ---- inverse (DIT): digit-reversed -> natural; only the first N lags leave, scaled ---
        const double sc1 = 1.0 / (double)S / var1 / dn, sc2 = two ? 1.0 / (double)S / var2 / dn : 0.0;
        for (int ps = plan.n_pass - 1; ps >= 0; --ps) {
            if (ps == 0) {
                auto st_out = [&](int64_t i, cplx val) {
                    if (i < N) {
                        out[s1 * N + i] = val.x * sc1;
                        if (two) out[s2 * N + i] = val.y * sc2;
                    }
                };
                pass_r<true>(plan.log2r[ps], S, log2m, W, ld_buf, st_out);
            } else {
                pass_r<true>(plan.log2r[ps], S, log2m, W, ld_buf, st_buf);
            }
            log2m += plan.log2r[ps];
        }
    }
}

static int64_t fft_size(int64_t N) {
    int64_t S = 1;
    while (S < 2 * N - 1) S <<= 1;   // 2 ** ceil(log2(2N - 1))
    return S;
}
static int fft_blocks(int64_t n_series) {
    const int64_t pairs = (n_series + 1) / 2;
    return (int)(pairs < 148 ? pairs : 148);
}
static FftPlan fft_plan(int64_t S) {
    int m = 0;
    while (((int64_t)1 << m) < S) ++m;
    FftPlan p;
    p.n_pass = 0;
    p.log2S = m;
    for (int i = 0; i < 4; ++i) p.log2r[i] = 0;
    while (m > 0) {
        const int r = m >= 5 ? 5 : m;
        p.log2r[p.n_pass++] = r;
        m -= r;
    }
    return p;
}

size_t acf_fft_ws_bytes(int64_t n_series, int64_t N) {
    const int64_t S = fft_size(N);
    return align_up((size_t)S * sizeof(cplx), 256) + (size_t)fft_blocks(n_series) * S * sizeof(cplx) + 512;
}

int acf_fft_launch(const SeriesView& v, double* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int64_t S = fft_size(v.N);
    if (S > ((int64_t)1 << 20)) {
        set_error("bk_autocorr: series of %lld draws need a transform of %lld points (limit 2^20)", (long long)v.N,
                  (long long)S);
        return BK_E_UNSUPPORTED;
    }
    const int nb = fft_blocks(v.n_series);
    Arena ar(ws, ws_bytes);
    cplx* W = ar.take<cplx>((size_t)S);
    cplx* scratch = ar.take<cplx>((size_t)nb * S);
    if (!ar.ok()) {
        set_error("bk_autocorr: workspace too small (need %zu bytes, got %zu)", ar.off, ws_bytes);
        return BK_E_WORKSPACE;
    }
    k_fft_twiddles<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(W, S);
    BK_LAUNCH_CHECK();
    k_acf_fft<<<nb, FFT_THREADS, 0, st>>>(v, S, fft_plan(S), W, scratch, out);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

}  // namespace bk
