// Shared by the diagnostics kernels.
#pragma once
#include "common.cuh"

namespace bk {

struct SeriesView {
    const void* x;
    int dtype;
    int64_t n_series, N, n_inner, ostride, istride, dstride;
    __device__ __forceinline__ double at(int64_t s, int64_t t) const {
        int64_t i = (s / n_inner) * ostride + (s % n_inner) * istride + t * dstride;
        return dtype == BK_F64 ? reinterpret_cast<const double*>(x)[i]
                               : (double)reinterpret_cast<const float*>(x)[i];
    }
};


// streaming Geyer IAT / ESS, one warp per series (ess_stream.cu)
size_t ess_stream_ws_bytes(const SeriesView& v);
int ess_stream_launch(const SeriesView& v, int estimator, double* iat, double* ess, void* ws, size_t ws_bytes,
                      cudaStream_t st);

// hand-written FFT autocorrelation (fft_autocorr.cu)
size_t acf_fft_ws_bytes(int64_t n_series, int64_t N);
int acf_fft_launch(const SeriesView& v, double* out, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace bk
