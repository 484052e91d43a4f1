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
#include <stdlib.h>

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

// =====================================================================================
// Shared-memory path: one REAL series per CTA, the whole transform on chip
// =====================================================================================
// The reference pads to S = 2^ceil(log2(2N-1)) (autocorr.py:26); any S >= 2N - 1 gives the same linear
// correlation, so this path picks the smallest S = c 2^a (c in {1, 3, 5}) whose HALF-length complex
// transform fits in shared memory: N = 10,000 -> S = 20,480, M = S / 2 = 10,240 complex doubles = 160 KB.
//   z_n = d_2n + i d_2n+1 (even / odd samples packed), Z = DFT_M(z) in place (decimation in frequency,
//   mixed radix: the odd factor first, then radix 16 / 8 / 4 / 2), natural -> digit-reversed;
//   pointwise, for every pair (k, M - k):  X[k] = E[k] + w_S^k O[k]  (E, O = spectra of the even / odd
//   samples),  P = |X|^2, and straight on to the packed spectrum of the (real, even) autocorrelation
//   Z'[k] = (P[k] + P[M-k]) + i conj(w_S^k) (P[k] - P[M-k]);  r_0 = sum_k P[k] is reduced on the way;
//   z' = IDFT_M(Z') (decimation in time, digit-reversed -> natural): acf_2n + i acf_2n+1 = z'_n / r_0.
// The series is read twice (mean, then the packed load), the N lags are written once; nothing else touches
// global memory except the twiddle tables (L2).  acf_k = r_k / r_0 equals autocorr.py:29-32's
// ifft(|fft|^2).real / var / N because r_0 = N var.
constexpr int RF_THREADS = 512;
constexpr int RF_MAX_M_CONST = 11264;     // (M + M / 8) complex doubles = 198 KB of shared memory for the skewed M-point line

template <int R, bool INV>
__device__ __forceinline__ void dft_small(cplx (&a)[R]) {
    if constexpr (R == 3) {
        const double h = 0.86602540378443864676;                  // sqrt(3) / 2
        const cplx t1 = cadd(a[1], a[2]);
        const cplx t2 = cplx{a[0].x - 0.5 * t1.x, a[0].y - 0.5 * t1.y};
        const cplx t3 = cplx{h * (a[1].x - a[2].x), h * (a[1].y - a[2].y)};
        a[0] = cadd(a[0], t1);
        const cplx m = cplx{t2.x + t3.y, t2.y - t3.x}, p = cplx{t2.x - t3.y, t2.y + t3.x};   // t2 -+ i t3
        a[1] = INV ? p : m;
        a[2] = INV ? m : p;
    } else if constexpr (R == 5) {
        const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;      // cos(2 pi/5), cos(4 pi/5)
        const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;       // sin(2 pi/5), sin(4 pi/5)
        const cplx t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]), t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
        const cplx m1 = cplx{a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y};
        const cplx m2 = cplx{a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y};
        const cplx n1 = cplx{s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y};
        const cplx n2 = cplx{s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y};
        a[0] = cadd(a[0], cadd(t1, t2));
        const cplx y1 = cplx{m1.x + n1.y, m1.y - n1.x}, y4 = cplx{m1.x - n1.y, m1.y + n1.x};   // m1 -+ i n1
        const cplx y2 = cplx{m2.x + n2.y, m2.y - n2.x}, y3 = cplx{m2.x - n2.y, m2.y + n2.x};
        a[1] = INV ? y4 : y1; a[4] = INV ? y1 : y4;
        a[2] = INV ? y3 : y2; a[3] = INV ? y2 : y3;
    } else {
        dft_reg<R, INV>(a);
    }
}
template <int R>
__host__ __device__ constexpr int out_slot(int k) { return (R == 3 || R == 5) ? k : bitrev<R>(k); }

struct RfftPlan {
    int M, n_pass;
    int radix[6];       // forward pass i has radix radix[i]; block length after pass i: m[i]
    int m[6];
};

// Twiddles w_M^e = exp(-2 pi i e / M) from two small shared-memory tables, w^e = A[e >> 6] * B[e & 63]: one
// complex multiply per twiddle and NO dependency between the R - 1 twiddles of a butterfly (a running product
// w^k = w^(k-1) w is a serial chain of fp64 multiplies -- with 8 warps per SM the kernel was latency-bound on it).
constexpr int RF_TW_LO = 64, RF_TW_HI_MAX = RF_MAX_M_CONST / RF_TW_LO + 1;
struct RfTw {
    const cplx* hi;   // [ceil(M / 64)]  w^(64 j)
    const cplx* lo;   // [64]            w^j
    __device__ __forceinline__ cplx at(int e, bool inv) const {
        cplx w = cmul(hi[e >> 6], lo[e & 63]);
        if (inv) w.y = -w.y;
        return w;
    }
};

// w^1 .. w^(R-1) from w^1 by a product tree of depth log2 R (w^k = w^hb w^(k-hb), hb = the highest power of two in k;
// squares for the powers of two): R - 2 complex multiplies, no table look-ups and no index arithmetic per twiddle,
// and no serial chain (a running product w^k = w^(k-1) w would be one).
__host__ __device__ constexpr int rf_hb(int k) { int h = 1; while (2 * h <= k) h *= 2; return h; }
template <int R>
__device__ __forceinline__ void rf_powers(cplx (&w)[R]) {
#pragma unroll
    for (int k = 2; k < R; ++k) {
        const int h = rf_hb(k);
        if (k == h) {
            const cplx b = w[k / 2];
            w[k] = cplx{fma(b.x, b.x, -b.y * b.y), 2.0 * b.x * b.y};
        } else {
            w[k] = cmul(w[h], w[k - h]);
        }
    }
}

// one in-place radix-R pass over the M-point line in shared memory; block length n = R m (see radix_pass).
// LIN: the R elements of a butterfly sit at a constant stride in the SKEWED line (element i lives at i + i / 8):
// for m % 8 == 0, SK(base + j m) = SK(base) + j (m + m / 8); for R m <= 8 a butterfly stays inside one group of 8.
// GL / GS: the load / store side goes through the callback (first forward pass reads the series, last inverse pass
// writes the lags) instead of the line.
template <int R, bool INV, bool LIN, bool GL, bool GS, class Load, class Store>
__device__ __forceinline__ void rf_pass(cplx* __restrict__ buf, int M, int m, const RfTw& tw, Load ld, Store st) {
    const int cnt = M / R, n = R * m, tw_stride = M / n;
    const int log2m = 31 - __clz(m);            // m is a power of two in every pass (the odd factor goes first)
    const int es = m >= 8 ? m + (m >> 3) : m;   // element stride in the skewed line (LIN)
    for (int t = threadIdx.x; t < cnt; t += RF_THREADS) {
        const int p = t & (m - 1), base = (t >> log2m) * n + p;
        cplx* pb = buf + base + (base >> 3);
        cplx a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if constexpr (GL) a[j] = ld(base + j * m);
            else if constexpr (LIN) a[j] = pb[j * es];
            else { const int i = base + j * m; a[j] = buf[i + (i >> 3)]; }
        }
        cplx w[R];
        w[1] = tw.at(p * tw_stride, INV);       // w_n^p = w_M^(p M / n); the other twiddles are its powers
        rf_powers<R>(w);
        if constexpr (INV) {
#pragma unroll
            for (int j = 1; j < R; ++j) a[j] = cmul(a[j], w[j]);
        }
        dft_small<R, INV>(a);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            cplx v = a[out_slot<R>(k)];
            if (!INV && k > 0) v = cmul(v, w[k]);
            if constexpr (GS) st(base + k * m, v);
            else if constexpr (LIN) pb[k * es] = v;
            else { const int i = base + k * m; buf[i + (i >> 3)] = v; }
        }
    }
    __syncthreads();
}
template <bool INV, bool GL, bool GS, class Load, class Store>
__device__ __forceinline__ void rf_pass_r(int R, cplx* __restrict__ buf, int M, int m, const RfTw& tw, Load ld, Store st) {
    const bool lin = m >= 8 || R * m <= 8;
#define BK_RF_CASE(RR)                                                                     \
    case RR:                                                                               \
        if (lin) rf_pass<RR, INV, true, GL, GS>(buf, M, m, tw, ld, st);                    \
        else rf_pass<RR, INV, false, GL, GS>(buf, M, m, tw, ld, st);                       \
        break;
    switch (R) {
        BK_RF_CASE(16)
        BK_RF_CASE(8)
        BK_RF_CASE(5)
        BK_RF_CASE(4)
        BK_RF_CASE(3)
        default:
            if (lin) rf_pass<2, INV, true, GL, GS>(buf, M, m, tw, ld, st);
            else rf_pass<2, INV, false, GL, GS>(buf, M, m, tw, ld, st);
            break;
    }
#undef BK_RF_CASE
}
// position of frequency k after the forward passes (mixed-radix digit reversal)
__device__ __forceinline__ int rf_pos(const RfftPlan& pl, int k) {
    int i = 0;
    for (int ps = 0; ps < pl.n_pass; ++ps) {
        int d;
        switch (pl.radix[ps]) {          // constant divisors: shifts / multiply-shift instead of a runtime division
            case 16: d = k & 15; k >>= 4; break;
            case 8: d = k & 7; k >>= 3; break;
            case 4: d = k & 3; k >>= 2; break;
            case 2: d = k & 1; k >>= 1; break;
            case 5: d = k % 5; k /= 5; break;
            default: d = k % 3; k /= 3; break;
        }
        i += d * pl.m[ps];
    }
    return i;
}

// TIN = element type of the draws; the per-series base offset (a 64-bit division / modulo in SeriesView::at) is
// formed once per series, the draw stride is 32-bit.
template <typename TIN>
__global__ void __launch_bounds__(RF_THREADS, 1)
k_acf_rfft(SeriesView v, RfftPlan plan, const cplx* __restrict__ WM, const cplx* __restrict__ WS,
           const uint16_t* __restrict__ POS_g, int pos_in_smem, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    cplx* buf = reinterpret_cast<cplx*>(rf_smem);          // [M + M / 8 + 1] skewed line
    __shared__ double red[33];
    __shared__ cplx tw_hi[RF_TW_HI_MAX], tw_lo[RF_TW_LO];
    __shared__ int s_rad[6], s_m[6];                       // the plan, out of local memory (it is indexed by the pass)
    const int M = plan.M, n_pass = plan.n_pass;
    for (int j = threadIdx.x; j * RF_TW_LO < M; j += RF_THREADS) tw_hi[j] = WM[j * RF_TW_LO];
    for (int j = threadIdx.x; j < RF_TW_LO; j += RF_THREADS) tw_lo[j] = WM[j < M ? j : 0];
    if (threadIdx.x < 6) { s_rad[threadIdx.x] = plan.radix[threadIdx.x]; s_m[threadIdx.x] = plan.m[threadIdx.x]; }
    // digit-reversed (skewed) positions of the pointwise pass: in shared memory behind the line when they fit
    const uint16_t* POS = POS_g;
    if (pos_in_smem) {
        uint16_t* ps = reinterpret_cast<uint16_t*>(rf_smem + (size_t)(M + M / 8 + 1) * sizeof(cplx));
        for (int j = threadIdx.x; j < M; j += RF_THREADS) ps[j] = POS_g[j];
        POS = ps;
    }
    const cplx w_s1 = WS[1];                               // exp(-2 pi i / S), S = 2 M
    __syncthreads();
    const RfTw WMt{tw_hi, tw_lo};
    const int N = (int)v.N;
    const int ds = (int)v.dstride;
    // element i lives at i + i / 8: the butterflies of the late passes walk the line with strides of 8 and 128
    // elements (16 B each) -- unskewed that is a 32-way / 4-way bank conflict, skewed every access pattern of the
    // transform costs the minimum of four wavefronts per 512-byte warp request
    auto SK = [](int i) { return i + (i >> 3); };
    auto ld_buf = [&](int i) { return buf[SK(i)]; };
    auto st_buf = [&](int i, cplx val) { buf[SK(i)] = val; };
    for (int64_t s = blockIdx.x; s < v.n_series; s += gridDim.x) {
        const TIN* __restrict__ xs = reinterpret_cast<const TIN*>(v.x) + (s / v.n_inner) * v.ostride + (s % v.n_inner) * v.istride;
        double* __restrict__ os = out + s * (int64_t)N;
        // ONE pass over the draws: the sum for the mean, and the raw samples parked in the (free) line, packed
        // z_n = d_2n + i d_2n+1 -- the first forward pass reads them from shared memory and demeans on the fly
        double l1 = 0;
        double* lined = reinterpret_cast<double*>(buf);
        for (int t = threadIdx.x; t < N; t += RF_THREADS) {
            const double xv = (double)xs[(int64_t)t * ds];
            l1 += xv;
            lined[2 * SK(t >> 1) + (t & 1)] = xv;
        }
        // the NEXT series of this CTA: pull its lines into L2 now, its mean pass and packed load start one transform later
        if (ds == 1 && s + gridDim.x < v.n_series) {
            const int64_t s2 = s + gridDim.x;
            const char* nx = reinterpret_cast<const char*>(reinterpret_cast<const TIN*>(v.x) + (s2 / v.n_inner) * v.ostride +
                                                           (s2 % v.n_inner) * v.istride);
            for (int64_t b = (int64_t)threadIdx.x * 128; b < (int64_t)N * (int64_t)sizeof(TIN); b += (int64_t)RF_THREADS * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + b));
        }
        const double mean = block_sum(l1, red) / (double)N;
        // ---- forward: packed load (even + i odd, demeaned, zero padded) fused into the first pass ----
        for (int ps = 0; ps < n_pass; ++ps) {
            if (ps == 0) {
                auto ld = [&](int i) {             // beyond the series the line still holds the previous transform: masked
                    const int e = 2 * i;
                    const cplx r = buf[SK(i)];
                    return cplx{e < N ? r.x - mean : 0.0, e + 1 < N ? r.y - mean : 0.0};
                };
                rf_pass_r<false, true, false>(s_rad[ps], buf, M, s_m[ps], WMt, ld, st_buf);
            } else {
                rf_pass_r<false, false, false>(s_rad[ps], buf, M, s_m[ps], WMt, ld_buf, st_buf);
            }
        }
        // ---- pointwise: power spectrum of the real series -> packed spectrum of its autocorrelation ----
        double r0 = 0;
#pragma unroll 4
        for (int k = threadIdx.x; k <= M / 2; k += RF_THREADS) {
            const int k2 = k ? M - k : 0;                  // partner (k = 0 pairs with itself and carries X[M])
            const int i = POS[k], i2 = POS[k2];     // skewed positions of the two frequencies (digit reversal: table)
            const cplx a = buf[i], b = buf[i2];
            if (k == 0) {
                const double x0 = 2.0 * (a.x + a.y), xm = 2.0 * (a.x - a.y);   // X[0] = E + O, X[M] = E - O (real; x2 like the k > 0 terms)
                const double p0 = x0 * x0, pm = xm * xm;
                r0 += p0 + pm;
                buf[i] = cplx{p0 + pm, p0 - pm};
            } else {
                // E = (Z[k] + conj Z[M-k]) / 2, O = (Z[k] - conj Z[M-k]) / 2i  (the factors 1/2 cancel in acf = r / r0)
                const cplx E = cplx{a.x + b.x, a.y - b.y}, O = cplx{a.y + b.y, b.x - a.x};
                // exp(-2 pi i k / S) = w_M^(k >> 1) (w_S^1 if k is odd): from the shared-memory tables, no global load
                cplx w = WMt.at(k >> 1, false);
                if (k & 1) w = cmul(w, w_s1);
                const cplx wo = cmul(w, O);
                const cplx Xk = cadd(E, wo);                            // X[k]
                const cplx Xm = cplx{E.x - wo.x, -(E.y - wo.y)};        // X[M-k] = conj(E - w O)
                const double pk = fma(Xk.x, Xk.x, Xk.y * Xk.y), pm = fma(Xm.x, Xm.x, Xm.y * Xm.y);
                r0 += (k == k2) ? 2.0 * pk : 2.0 * (pk + pm);
                // Z'[k] = (P[k] + P[M-k]) + i conj(w^k) (P[k] - P[M-k]);  Z'[M-k] = (P[k] + P[M-k]) + i w^k (P[k] - P[M-k]) ... sign below
                const double sm = pk + pm, df = pk - pm;
                // i * conj(w) * df = i (w.x - i w.y) df = (w.y df, w.x df)
                buf[i] = cplx{sm + w.y * df, w.x * df};
                if (k != k2) {
                    // at M-k: exp(+2 pi i (M-k) / S) = -conj(exp(+2 pi i k / S)) = -w  ->  i * (-w) * (P[M-k] - P[k]) = i w df
                    buf[i2] = cplx{sm - w.y * df, w.x * df};
                }
            }
        }
        r0 = block_sum(r0, red);
        __syncthreads();
        // ---- inverse: digit-reversed -> natural; the last pass writes acf_2n, acf_2n+1 = z'_n / r0 ----
        const double sc = 1.0 / r0;
        for (int ps = n_pass - 1; ps >= 0; --ps) {
            if (ps == 0) {
                auto st_out = [&](int i, cplx val) {
                    const int e = 2 * i;
                    if (e < N) os[e] = val.x * sc;
                    if (e + 1 < N) os[e + 1] = val.y * sc;
                };
                rf_pass_r<true, false, true>(s_rad[ps], buf, M, s_m[ps], WMt, ld_buf, st_out);
            } else {
                rf_pass_r<true, false, false>(s_rad[ps], buf, M, s_m[ps], WMt, ld_buf, st_buf);
            }
        }
    }
}

// WM[j] = exp(-2 pi i j / M) (j < M), WS[k] = exp(-2 pi i k / (2 M)) (k <= M / 2)
// POS[k] = skewed position of frequency k after the forward passes (k < M)
__global__ void k_rfft_twiddles(cplx* __restrict__ WM, cplx* __restrict__ WS, uint16_t* __restrict__ POS, RfftPlan plan, int M) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < M) { const int i = rf_pos(plan, j); POS[j] = (uint16_t)(i + (i >> 3)); }
    double sn, cs;
    if (j < M) {
        sincospi(-2.0 * (double)j / (double)M, &sn, &cs);
        WM[j] = cplx{cs, sn};
    }
    if (j <= M / 2) {
        sincospi(-(double)j / (double)M, &sn, &cs);
        WS[j] = cplx{cs, sn};
    }
}

constexpr int RF_MAX_M = RF_MAX_M_CONST;
// smallest S = c 2^a >= 2N - 1 (c in {1, 3, 5}, S even) whose half-length transform fits on chip; 0: none
static int64_t rfft_size(int64_t N) {
    int64_t best = 0;
    for (int c : {1, 3, 5})
        for (int64_t S = 2 * c; S / 2 <= RF_MAX_M; S *= 2)
            if (S >= 2 * N - 1) { if (!best || S < best) best = S; break; }
    return best;
}
static RfftPlan rfft_plan(int M) {
    RfftPlan p;
    memset(&p, 0, sizeof(p));
    p.M = M;
    int rest = M;
    auto push = [&](int r) { rest /= r; p.radix[p.n_pass] = r; p.m[p.n_pass] = rest; ++p.n_pass; };
    if (rest % 5 == 0) push(5);
    else if (rest % 3 == 0) push(3);
    while (rest % 16 == 0 && rest > 16) push(16);
    while (rest > 1) {
        if (rest % 16 == 0) push(16);
        else if (rest % 8 == 0) push(8);
        else if (rest % 4 == 0) push(4);
        else push(2);
    }
    return p;
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

static bool use_rfft(int64_t N) {
    const char* e = getenv("BK_ACF_RFFT");      // diagnostic: BK_ACF_RFFT=0 forces the global-scratch transform
    return !(e && e[0] == '0') && rfft_size(N) > 0;
}

size_t acf_fft_ws_bytes(int64_t n_series, int64_t N) {
    if (use_rfft(N)) {
        const int64_t M = rfft_size(N) / 2;
        return align_up((size_t)M * sizeof(cplx), 256) + align_up((size_t)(M / 2 + 1) * sizeof(cplx), 256) +
               align_up((size_t)M * sizeof(uint16_t), 256) + 512;
    }
    const int64_t S = fft_size(N);
    return align_up((size_t)S * sizeof(cplx), 256) + (size_t)fft_blocks(n_series) * S * sizeof(cplx) + 512;
}

int acf_fft_launch(const SeriesView& v, double* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    BK_CHECK_ARG(v.dstride < ((int64_t)1 << 31), "bk_autocorr: draw stride %lld too large", (long long)v.dstride);
    if (use_rfft(v.N)) {
        const int M = (int)(rfft_size(v.N) / 2);
        Arena ar(ws, ws_bytes);
        cplx* WM = ar.take<cplx>((size_t)M);
        cplx* WS = ar.take<cplx>((size_t)(M / 2 + 1));
        uint16_t* POS = ar.take<uint16_t>((size_t)M);
        if (!ar.ok()) {
            set_error("bk_autocorr: workspace too small (need %zu bytes, got %zu)", ar.off, ws_bytes);
            return BK_E_WORKSPACE;
        }
        const RfftPlan plan = rfft_plan(M);
        k_rfft_twiddles<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(WM, WS, POS, plan, M);
        BK_LAUNCH_CHECK();
        size_t smem = (size_t)(M + M / 8 + 1) * sizeof(cplx);
        const int pos_in_smem = smem + (size_t)M * sizeof(uint16_t) + 8 * 1024 <= 227 * 1024 ? 1 : 0;   // + static tables
        if (pos_in_smem) smem += (size_t)M * sizeof(uint16_t);
        BK_CUDA(cudaFuncSetAttribute(k_acf_rfft<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BK_CUDA(cudaFuncSetAttribute(k_acf_rfft<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int per_sm = smem <= 100 * 1024 ? 2 : 1;
        const int64_t nb = v.n_series < (int64_t)sms * per_sm ? v.n_series : (int64_t)sms * per_sm;
        if (v.dtype == BK_F64) k_acf_rfft<double><<<(unsigned)nb, RF_THREADS, smem, st>>>(v, plan, WM, WS, POS, pos_in_smem, out);
        else k_acf_rfft<float><<<(unsigned)nb, RF_THREADS, smem, st>>>(v, plan, WM, WS, POS, pos_in_smem, out);
        BK_LAUNCH_CHECK();
        return BK_OK;
    }
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
