// Hierarchical logistic regression plugin (BASELINE config c3; density defined in
// oracle/models.py::HierLogReg -- there is no upstream definition).
//
//   theta = (beta[Dx], mu, lam),  tau = exp(lam)
//   log p = sum_n [y_n z_n - softplus(z_n)]            z = X beta
//           - Dx*lam - 0.5 exp(-2 lam) sum_j (beta_j - mu)^2 - 0.5 mu^2 - 0.5 exp(2 lam) + lam
//
// CUDA-core evaluator, batched over chains (fp32 timed / fp64 parity):
//   k_hlr_partial: a CTA owns 64 chains x one slice of the observations; per 64-row
//     X tile it forms the 64x64 logit tile in registers (never written to HBM),
//     turns it into residuals r = y - sigmoid(z) in shared memory and accumulates
//     G[c, j] += sum_n r[n, c] X[n, j]  -- the FlashAttention-shaped
//     "GEMM -> pointwise -> GEMM" fusion of SURVEY.md section 7 step 5.
//   k_hlr_finish: fixed-order sum over the observation slices (deterministic: no
//     atomics) + the hierarchical prior terms.
// The tcgen05 version of the two contractions is the next step for this plugin.
#include <stdlib.h>

#include "model.h"

namespace bk {

constexpr int HLR_CT = 64;   // chains per CTA
constexpr int HLR_NT = 64;   // observations per tile
constexpr int HLR_MAXJ = 8;  // Dx <= 16 * HLR_MAXJ = 128

template <typename T>
__device__ __forceinline__ T softplus_(T z) {
    T az = z < T(0) ? -z : z;
    return (z > T(0) ? z : T(0)) + Ar<T>::log1p_(Ar<T>::exp_(-az));
}
template <typename T>
__device__ __forceinline__ T sigmoid_(T z) {
    return T(0.5) * (T(1) + tanh(T(0.5) * z));
}

template <typename T, bool WITH_GRAD>
__global__ void __launch_bounds__(256) k_hlr_partial(const T* __restrict__ X, const T* __restrict__ y,
                                                     const T* __restrict__ theta, int64_t C, int64_t N,
                                                     int Dx, int D, int64_t rows_per_split,
                                                     T* __restrict__ part_g, T* __restrict__ part_ll) {
    extern __shared__ unsigned char smem_raw[];
    const int S = Dx + 1;                       // odd-ish row stride: conflict-free column walks
    T* Bs = reinterpret_cast<T*>(smem_raw);     // [CT][S]  beta tile
    T* Xs = Bs + HLR_CT * S;                    // [NT][S]  X tile
    T* Rs = Xs + HLR_NT * S;                    // [NT][CT] residual tile
    T* Ys = Rs + HLR_NT * HLR_CT;               // [NT]
    T* Ls = Ys + HLR_NT;                        // [16][CT] log-lik partials
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t c0 = (int64_t)blockIdx.x * HLR_CT;
    const int split = blockIdx.y;
    const int64_t n_begin = split * rows_per_split;
    const int64_t n_end = n_begin + rows_per_split < N ? n_begin + rows_per_split : N;

    for (int i = threadIdx.x; i < HLR_CT * Dx; i += 256) {
        int c = i / Dx, j = i % Dx;
        Bs[c * S + j] = (c0 + c < C) ? theta[(c0 + c) * D + j] : T(0);
    }
    T accg[4][HLR_MAXJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < HLR_MAXJ; ++j) accg[i][j] = T(0);
    T ll[4] = {T(0), T(0), T(0), T(0)};   // chains c = tx + 16 j

    for (int64_t n0 = n_begin; n0 < n_end; n0 += HLR_NT) {
        __syncthreads();
        for (int i = threadIdx.x; i < HLR_NT * Dx; i += 256) {
            int n = i / Dx, j = i % Dx;
            Xs[n * S + j] = (n0 + n < n_end) ? X[(n0 + n) * Dx + j] : T(0);
        }
        if (threadIdx.x < HLR_NT) Ys[threadIdx.x] = (n0 + threadIdx.x < n_end) ? y[n0 + threadIdx.x] : T(0);
        __syncthreads();
        // logits z[n = ty + 16 i][c = tx + 16 j]
        T z[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) z[i][j] = T(0);
        for (int k = 0; k < Dx; ++k) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Xs[(ty + 16 * i) * S + k];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(tx + 16 * j) * S + k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) z[i][j] = fma(a[i], b[j], z[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = ty + 16 * i;
            const bool live = n0 + n < n_end;
            const T yn = Ys[n];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const T zz = z[i][j];
                if (live) ll[j] += yn * zz - softplus_(zz);
                if (WITH_GRAD) Rs[n * HLR_CT + tx + 16 * j] = live ? yn - sigmoid_(zz) : T(0);
            }
        }
        if (!WITH_GRAD) continue;   // log-density only: no second contraction
        __syncthreads();
        // G[c = ty + 16 i][j = tx + 16 jj] += sum_n R[n][c] X[n][j]
        for (int n = 0; n < HLR_NT; ++n) {
            T rv[4], xv[HLR_MAXJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) rv[i] = Rs[n * HLR_CT + ty + 16 * i];
#pragma unroll
            for (int jj = 0; jj < HLR_MAXJ; ++jj) {
                const int j = tx + 16 * jj;
                xv[jj] = j < Dx ? Xs[n * S + j] : T(0);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < HLR_MAXJ; ++jj) accg[i][jj] = fma(rv[i], xv[jj], accg[i][jj]);
        }
    }
    // write this slice's partials
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t c = c0 + ty + 16 * i;
        if (c >= C || !WITH_GRAD) continue;
#pragma unroll
        for (int jj = 0; jj < HLR_MAXJ; ++jj) {
            const int j = tx + 16 * jj;
            if (j < Dx) part_g[((int64_t)split * C + c) * Dx + j] = accg[i][jj];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) Ls[ty * HLR_CT + tx + 16 * j] = ll[j];
    __syncthreads();
    if (threadIdx.x < HLR_CT) {
        T s = T(0);
        for (int t = 0; t < 16; ++t) s += Ls[t * HLR_CT + threadIdx.x];
        if (c0 + threadIdx.x < C) part_ll[(int64_t)split * C + c0 + threadIdx.x] = s;
    }
}

// one warp per chain: fixed-order reduction over slices + prior terms
template <typename T>
__global__ void k_hlr_finish(const T* __restrict__ theta, const T* __restrict__ part_g,
                             const T* __restrict__ part_ll, int64_t C, int Dx, int D, int n_split,
                             T* __restrict__ lp, T* __restrict__ grad) {
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    const T* th = theta + c * D;
    const T mu = th[Dx], lam = th[Dx + 1];
    const T e2 = Ar<T>::exp_(T(-2) * lam), ep2 = Ar<T>::exp_(T(2) * lam);
    T sr = T(0), ss = T(0);
    for (int j = lane; j < Dx; j += 32) {
        T r = th[j] - mu;
        sr += r;
        ss = fma(r, r, ss);
        if (grad) {
            // four interleaved partial sums (independent loads in flight), combined in a fixed order
            T g0 = T(0), g1 = T(0), g2 = T(0), g3 = T(0);
            const T* pg = part_g + c * Dx + j;
            const int64_t step = C * (int64_t)Dx;
            int s = 0;
            for (; s + 4 <= n_split; s += 4) {
                const T a0 = pg[(s + 0) * step], a1 = pg[(s + 1) * step], a2 = pg[(s + 2) * step],
                        a3 = pg[(s + 3) * step];
                g0 += a0; g1 += a1; g2 += a2; g3 += a3;
            }
            for (; s < n_split; ++s) g0 += pg[s * step];
            grad[c * D + j] = ((g0 + g1) + (g2 + g3)) - e2 * r;
        }
    }
    sr = warp_sum(sr);
    ss = warp_sum(ss);
    T ll = T(0);
    if (lp) {   // slices over lanes, then the (deterministic) butterfly
        for (int s = lane; s < n_split; s += 32) ll += part_ll[(int64_t)s * C + c];
        ll = warp_sum(ll);
    }
    if (lane == 0) {
        if (lp) {
            lp[c] = ll - T(Dx) * lam - T(0.5) * e2 * ss - T(0.5) * mu * mu - T(0.5) * ep2 + lam;
        }
        if (grad) {
            grad[c * D + Dx] = e2 * sr - mu;
            grad[c * D + Dx + 1] = -T(Dx) + e2 * ss - ep2 + T(1);
        }
    }
}

static int hlr_splits(int64_t C, int64_t N) {
    const int64_t ctiles = (C + HLR_CT - 1) / HLR_CT;
    int64_t want = (148 * 4 + ctiles - 1) / ctiles;          // ~4 CTAs per SM in flight
    const int64_t max_split = (N + HLR_NT - 1) / HLR_NT;
    if (want > max_split) want = max_split;
    if (want < 1) want = 1;
    return (int)want;
}

size_t hlr_eval_ws_bytes(const Model& m, int64_t C) {
    const size_t es = m.d.dtype == BK_F64 ? 8 : 4;
    const int Dx = (int)m.d.dims - 2;
    const int ns = hlr_splits(C, m.d.n_obs);
    size_t b = align_up((size_t)ns * C * Dx * es, 256) + align_up((size_t)ns * C * es, 256) + 512;
    if (hlr_tc_enabled(m)) {
        size_t t = hlr_tc_eval_ws_bytes(m, C);
        if (t > b) b = t;
    }
    return b;
}

template <typename T>
static int hlr_eval_t(const Model& m, const T* theta, int64_t C, T* lp, T* grad, void* ws, size_t ws_bytes,
                      cudaStream_t st, bool precise) {
    const int D = (int)m.d.dims, Dx = D - 2;
    const int64_t N = m.d.n_obs;
    if constexpr (sizeof(T) == 4) {
        // tcgen05 paths (logreg_tc.cu): leapfrog gradients with bf16 operands (precise = false),
        // and the density on its own -- what a Metropolis test consumes -- with split-precision logits
        if (hlr_tc_enabled(m) && (!precise || !grad)) {
            float *pg, *pl;
            int ns;
            const char* dbg = getenv("BK_HLR_DEBUG");   // bit 2: time the gradient-only kernel through the public call
            const bool no_ll = !lp || (dbg && (atoi(dbg) & 2));
            const int mode = precise ? HLR_TC_LP : (no_ll ? HLR_TC_GRAD : HLR_TC_GRAD_LL);
            int rc = hlr_tc_partial(m, theta, C, ws, ws_bytes, &pg, &pl, &ns, st, mode);
            if (rc) return rc;
            k_hlr_finish<float><<<(unsigned)((C * 32 + 63) / 64), 64, 0, st>>>(theta, pg, pl, C, Dx, D, ns, lp,
                                                                                 precise ? nullptr : grad);
            BK_LAUNCH_CHECK();
            return BK_OK;
        }
    }
    if (Dx > 16 * HLR_MAXJ) {
        set_error("HIER_LOGREG supports up to %d regressors (got %d)", 16 * HLR_MAXJ, Dx);
        return BK_E_UNSUPPORTED;
    }
    const int ns = hlr_splits(C, N);
    Arena ar(ws, ws_bytes);
    T* pg = ar.take<T>((size_t)ns * C * Dx);
    T* pl = ar.take<T>((size_t)ns * C);
    if (!ar.ok()) {
        set_error("model eval workspace too small (%zu < %zu)", ws_bytes, ar.off);
        return BK_E_WORKSPACE;
    }
    int64_t rows = (N + ns - 1) / ns;
    rows = (rows + HLR_NT - 1) / HLR_NT * HLR_NT;
    const size_t smem = ((size_t)(HLR_CT + HLR_NT) * (Dx + 1) + HLR_NT * HLR_CT + HLR_NT + 16 * HLR_CT) * sizeof(T);
    static bool attr32 = false, attr64 = false;
    bool& attr = sizeof(T) == 8 ? attr64 : attr32;
    if (!attr) {
        BK_CUDA(cudaFuncSetAttribute(k_hlr_partial<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        BK_CUDA(cudaFuncSetAttribute(k_hlr_partial<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    if (smem > 200 * 1024) { set_error("HIER_LOGREG tile does not fit shared memory"); return BK_E_UNSUPPORTED; }
    dim3 grid((unsigned)((C + HLR_CT - 1) / HLR_CT), (unsigned)ns);
    prof_begin(BK_PROF_GRAD, st);
    if (grad)
        k_hlr_partial<T, true><<<grid, 256, smem, st>>>((const T*)m.d.X, (const T*)m.d.y, theta, C, N, Dx, D, rows, pg, pl);
    else
        k_hlr_partial<T, false><<<grid, 256, smem, st>>>((const T*)m.d.X, (const T*)m.d.y, theta, C, N, Dx, D, rows, pg, pl);
    prof_end(BK_PROF_GRAD, st);
    BK_LAUNCH_CHECK();
    k_hlr_finish<T><<<(unsigned)((C * 32 + 63) / 64), 64, 0, st>>>(theta, pg, pl, C, Dx, D, ns, lp, grad);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

int hlr_eval(const Model& m, const void* theta, int64_t C, void* lp, void* grad, void* ws, size_t ws_bytes,
             cudaStream_t st, bool precise) {
    if (m.d.dtype == BK_F64)
        return hlr_eval_t<double>(m, (const double*)theta, C, (double*)lp, (double*)grad, ws, ws_bytes, st, true);
    return hlr_eval_t<float>(m, (const float*)theta, C, (float*)lp, (float*)grad, ws, ws_bytes, st, precise);
}

}  // namespace bk
