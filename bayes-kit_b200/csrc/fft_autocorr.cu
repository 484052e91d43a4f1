// Hand-written (cuFFT-free) FFT autocorrelation, fp64 -- autocorr.py:6-33.
//
// Per series: demean, zero-pad to S = 2^ceil(log2(2N-1)) (autocorr.py:26),
// forward FFT, |.|^2, inverse FFT, scale by 1/(var*N) (autocorr.py:27-32).
// The forward transform is an in-place radix-2 decimation-in-FREQUENCY FFT
// (natural in, bit-reversed out), the power spectrum is order-agnostic, and the
// inverse is an in-place radix-2 decimation-in-TIME FFT (bit-reversed in,
// natural out) -- so no bit-reversal pass exists at all.
//
// One CTA per series (persistent, grid-stride).  Sub-transforms of up to
// B = 8192 points (128 KB of complex fp64) plus their 64 KB twiddle table live
// in shared memory; for S > B (N = 10,000 -> S = 32,768) the log2(S/B) widest
// DIF stages and the matching last DIT stages run on an L2-resident per-CTA
// scratch line of S complex values, everything else in shared memory.
#include "diag.h"

namespace bk {

constexpr int FFT_THREADS = 512;
constexpr int FFT_B = 8192;  // largest in-smem sub-transform

struct cplx { double x, y; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cplx{fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)};
}
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {  // a * conj(b)
    return cplx{fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y)};
}

// W[j] = exp(-2 pi i j / S), j < S/2
__global__ void k_fft_twiddles(cplx* __restrict__ W, int64_t S) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S / 2) return;
    double s, c;
    sincospi(-2.0 * (double)j / (double)S, &s, &c);
    W[j] = cplx{c, s};
}

// in-place DIF stages h = h0, h0/2, ..., h1 on `a` (len elements), twiddle table tw for length `tl`
// (tw[j] = exp(-2 pi i j / tl)); W_{2h}^j = tw[j * tl / (2h)]
__device__ __forceinline__ void dif_stages(cplx* a, int64_t len, int64_t h0, int64_t h1, const cplx* tw,
                                           int64_t tl) {
    for (int64_t h = h0; h >= h1; h >>= 1) {
        const int64_t tstride = tl / (2 * h);
        for (int64_t i = threadIdx.x; i < len / 2; i += FFT_THREADS) {
            const int64_t j = i & (h - 1), b = (i - j) << 1;
            cplx u = a[b + j], v = a[b + j + h];
            a[b + j] = cplx{u.x + v.x, u.y + v.y};
            a[b + j + h] = cmul(cplx{u.x - v.x, u.y - v.y}, tw[j * tstride]);
        }
        __syncthreads();
    }
}
// in-place DIT inverse stages h = h0, 2 h0, ..., h1
__device__ __forceinline__ void dit_inv_stages(cplx* a, int64_t len, int64_t h0, int64_t h1, const cplx* tw,
                                               int64_t tl) {
    for (int64_t h = h0; h <= h1; h <<= 1) {
        const int64_t tstride = tl / (2 * h);
        for (int64_t i = threadIdx.x; i < len / 2; i += FFT_THREADS) {
            const int64_t j = i & (h - 1), b = (i - j) << 1;
            cplx u = a[b + j], v = cmulc(a[b + j + h], tw[j * tstride]);
            a[b + j] = cplx{u.x + v.x, u.y + v.y};
            a[b + j + h] = cplx{u.x - v.x, u.y - v.y};
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FFT_THREADS) k_acf_fft(SeriesView v, int64_t S, const cplx* __restrict__ W,
                                                         cplx* __restrict__ scratch_all,
                                                         double* __restrict__ out) {
    extern __shared__ double smem_d[];
    __shared__ double red[33];
    const int64_t B = S < FFT_B ? S : FFT_B;
    cplx* sm = reinterpret_cast<cplx*>(smem_d);          // [B]
    cplx* tws = sm + B;                                   // [B/2] twiddles of the length-B transform
    cplx* scratch = S > B ? scratch_all + (int64_t)blockIdx.x * S : nullptr;
    const int64_t N = v.N;
    for (int64_t j = threadIdx.x; j < B / 2; j += FFT_THREADS) tws[j] = W[j * (S / B)];
    __syncthreads();

    for (int64_t s = blockIdx.x; s < v.n_series; s += gridDim.x) {
        // mean / variance (ddof = 0, autocorr.py:27-28)
        double loc = 0;
        for (int64_t t = threadIdx.x; t < N; t += FFT_THREADS) loc += v.at(s, t);
        const double mean = block_sum(loc, red) / (double)N;
        loc = 0;
        for (int64_t t = threadIdx.x; t < N; t += FFT_THREADS) {
            double c = v.at(s, t) - mean;
            loc = fma(c, c, loc);
        }
        const double var = block_sum(loc, red) / (double)N;
        const double scale = 1.0 / (double)S;

        if (S <= B) {
            for (int64_t i = threadIdx.x; i < S; i += FFT_THREADS)
                sm[i] = cplx{i < N ? v.at(s, i) - mean : 0.0, 0.0};
            __syncthreads();
            dif_stages(sm, S, S / 2, 1, tws, B);
            for (int64_t i = threadIdx.x; i < S; i += FFT_THREADS) {
                cplx z = sm[i];
                sm[i] = cplx{fma(z.x, z.x, z.y * z.y), 0.0};
            }
            __syncthreads();
            dit_inv_stages(sm, S, 1, S / 2, tws, B);
            for (int64_t k = threadIdx.x; k < N; k += FFT_THREADS)
                out[s * N + k] = sm[k].x * scale / var / (double)N;
            __syncthreads();
        } else {
            // widest DIF stages on the scratch line; the first one reads the (real, padded) input
            for (int64_t i = threadIdx.x; i < S; i += FFT_THREADS)
                scratch[i] = cplx{i < N ? v.at(s, i) - mean : 0.0, 0.0};
            __syncthreads();
            dif_stages(scratch, S, S / 2, B, W, S);
            // each contiguous length-B block: remaining DIF stages, |.|^2, first DIT stages -- in smem
            for (int64_t blk = 0; blk < S / B; ++blk) {
                cplx* g = scratch + blk * B;
                for (int64_t i = threadIdx.x; i < B; i += FFT_THREADS) sm[i] = g[i];
                __syncthreads();
                dif_stages(sm, B, B / 2, 1, tws, B);
                for (int64_t i = threadIdx.x; i < B; i += FFT_THREADS) {
                    cplx z = sm[i];
                    sm[i] = cplx{fma(z.x, z.x, z.y * z.y), 0.0};
                }
                __syncthreads();
                dit_inv_stages(sm, B, 1, B / 2, tws, B);
                for (int64_t i = threadIdx.x; i < B; i += FFT_THREADS) g[i] = sm[i];
                __syncthreads();
            }
            dit_inv_stages(scratch, S, B, S / 2, W, S);
            for (int64_t k = threadIdx.x; k < N; k += FFT_THREADS)
                out[s * N + k] = scratch[k].x * scale / var / (double)N;
            __syncthreads();
        }
    }
}

static int64_t fft_size(int64_t N) {
    int64_t S = 1;
    while (S < 2 * N - 1) S <<= 1;   // 2 ** ceil(log2(2N - 1))
    return S;
}
static int fft_blocks(int64_t n_series) { return (int)(n_series < 148 ? n_series : 148); }

size_t acf_fft_ws_bytes(int64_t n_series, int64_t N) {
    const int64_t S = fft_size(N);
    size_t b = align_up((size_t)(S / 2) * sizeof(cplx), 256) + 512;
    if (S > FFT_B) b += (size_t)fft_blocks(n_series) * S * sizeof(cplx);
    return b;
}

int acf_fft_launch(const SeriesView& v, double* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int64_t S = fft_size(v.N);
    const int nb = fft_blocks(v.n_series);
    Arena ar(ws, ws_bytes);
    cplx* W = ar.take<cplx>(S / 2 > 0 ? S / 2 : 1);
    cplx* scratch = S > FFT_B ? ar.take<cplx>((size_t)nb * S) : nullptr;
    if (!ar.ok()) {
        set_error("bk_autocorr: workspace too small (need %zu bytes, got %zu)", ar.off, ws_bytes);
        return BK_E_WORKSPACE;
    }
    k_fft_twiddles<<<(unsigned)((S / 2 + 255) / 256), 256, 0, st>>>(W, S);
    BK_LAUNCH_CHECK();
    const int64_t B = S < FFT_B ? S : FFT_B;
    const size_t smem = (size_t)(B + B / 2) * sizeof(cplx);
    static bool attr = false;
    if (!attr) {
        BK_CUDA(cudaFuncSetAttribute(k_acf_fft, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (FFT_B + FFT_B / 2) * (int)sizeof(cplx)));
        attr = true;
    }
    k_acf_fft<<<nb, FFT_THREADS, smem, st>>>(v, S, W, scratch, out);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

}  // namespace bk
