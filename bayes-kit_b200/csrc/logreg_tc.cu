// Hierarchical logistic regression gradient on tcgen05 (fp32 timed mode, interior
// leapfrog steps): a FlashAttention-shaped fusion -- the N x C logits never leave
// the SM.
//
// A CTA owns 128 chains (TMEM lanes) and a slice of the observations; per tile of
// 128 observations:
//   GEMM1  Z[c, n]  = sum_j beta[c, j] X[n, j]          (tcgen05, D1 in TMEM, double-buffered)
//   epi    r[c, n]  = y_n - sigmoid(Z);  ll[c] += y_n Z - softplus(Z)   (registers; thread = chain)
//          r -> bf16 -> shared memory as the K-major A operand of GEMM2
//   GEMM2  G[c, j] += sum_n r[c, n] X[n, j]              (tcgen05, D2 persistent in TMEM)
// X is streamed twice per tile by TMA (as [n][j] for GEMM1 and as the pre-transposed
// [j][n] copy for GEMM2, both K-major / 128B swizzle, both L2-resident: 2 x 26 MB at
// c3).  Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = epilogue
// (lane quarter = warp % 4, observation half = (warp - 2) / 4).
// Operands are bf16: like the dense plugin this path serves the INTERIOR leapfrog
// gradients; the endpoint gradient / log density that enter the Metropolis test
// come from the fp32 CUDA-core evaluator (logreg.cu).
#include <stdlib.h>

#include "model.h"
#include "tc_ptx.cuh"

namespace bk {

using namespace ptx;

namespace hlrtc {
constexpr int CT = 128, NT = 128, KJ = 128;
constexpr int ATOM = 128 * 64 * 2;                    // one [128 rows x 64 k] swizzle block = 16 KB
constexpr int STAGE_BYTES = 4 * ATOM + 1024;          // X (2 atoms) + X^T (2 atoms) + y tile
constexpr int OFF_A1 = 0, OFF_STAGE = 2 * ATOM, OFF_A2 = OFF_STAGE + 2 * STAGE_BYTES;
constexpr int OFF_BARS = OFF_A2 + 2 * ATOM, OFF_LL = OFF_BARS + 256;
constexpr int SMEM_BYTES = OFF_LL + 2 * CT * 4 + 1024;
constexpr int THREADS = 320;
constexpr uint32_t IDESC = idesc_bf16(128, 128);

struct Args {
    int64_t C, N;
    int Dx;
    int64_t rows_per_split;     // multiple of NT
    const float* y;             // [Np] zero padded
    float* part_g;              // [n_split, C, Dx]
    float* part_ll;             // [n_split, C]
};

__global__ void __launch_bounds__(THREADS, 1)
k_hlr_tc(const __grid_constant__ CUtensorMap mapBeta, const __grid_constant__ CUtensorMap mapX,
         const __grid_constant__ CUtensorMap mapXT, const Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA1 = base + OFF_A1, sA2 = base + OFF_A2, bars = base + OFF_BARS;
    auto stage = [&](int s) { return base + OFF_STAGE + (uint32_t)s * STAGE_BYTES; };
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (2 + s); };
    auto d1_full = [&](int b) { return bars + 8u * (4 + b); };
    const uint32_t a2_full = bars + 8u * 6, a2_free = bars + 8u * 7, d2_full = bars + 8u * 8,
                   beta_full = bars + 8u * 9, tmem_slot = bars + 8u * 10;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + OFF_BARS + 80);
    float* lls = reinterpret_cast<float*>(gbase + OFF_LL);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c0 = (int64_t)blockIdx.x * CT;
    const int split = blockIdx.y;
    const int64_t n_begin = split * a.rows_per_split;
    const int64_t n_end = n_begin + a.rows_per_split < a.N ? n_begin + a.rows_per_split : a.N;
    const int T = n_end > n_begin ? (int)((n_end - n_begin + NT - 1) / NT) : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); mbar_init(d1_full(s), 1); }
        mbar_init(a2_full, 256); mbar_init(a2_free, 1); mbar_init(d2_full, 1); mbar_init(beta_full, 1);
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;
    const uint32_t tD1 = tmem, tD2 = tmem + 256;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(beta_full, 2 * ATOM);
            tma_load_2d(sA1, &mapBeta, beta_full, 0, (int)c0);
            tma_load_2d(sA1 + ATOM, &mapBeta, beta_full, 64, (int)c0);
            for (int i = 0; i < T; ++i) {
                const int s = i & 1;
                const uint32_t ph = (uint32_t)(i >> 1) & 1u;
                const int n0 = (int)(n_begin + (int64_t)i * NT);
                mbar_wait(empty(s), ph ^ 1u);
                mbar_expect_tx(full(s), 4 * ATOM + NT * 4);
                const uint32_t st = stage(s);
                tma_load_2d(st, &mapX, full(s), 0, n0);                 // X[n0.., j 0..63]
                tma_load_2d(st + ATOM, &mapX, full(s), 64, n0);         // X[n0.., j 64..127]
                tma_load_2d(st + 2 * ATOM, &mapXT, full(s), n0, 0);     // X^T[j, n0..n0+63]
                tma_load_2d(st + 3 * ATOM, &mapXT, full(s), n0 + 64, 0);
                bulk_load(st + 4 * ATOM, a.y + n0, NT * 4, full(s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && T > 0) {
            auto kdesc = [](uint32_t tile, int kk) { return umma_desc<128>(tile + (kk >> 2) * ATOM + (kk & 3) * 32); };
            auto mma2 = [&](int t) {   // G += R_t . X_t   (A = R in smem, B = X^T tile)
                const int s = t & 1;
                mbar_wait(a2_full, (uint32_t)t & 1u);
                tc_fence_after();
                const uint32_t xt = stage(s) + 2 * ATOM;
#pragma unroll
                for (int kk = 0; kk < NT / 16; ++kk) umma(tD2, kdesc(sA2, kk), kdesc(xt, kk), IDESC, (t | kk) != 0);
                umma_commit(empty(s));     // stage reusable once GEMM2 has read X^T
                umma_commit(a2_free);      // R buffer reusable
            };
            mbar_wait(beta_full, 0);
            for (int i = 0; i < T; ++i) {
                const int s = i & 1;
                mbar_wait(full(s), (uint32_t)(i >> 1) & 1u);
                tc_fence_after();
                const uint32_t d1 = tD1 + (uint32_t)(i & 1) * 128;
#pragma unroll
                for (int kk = 0; kk < KJ / 16; ++kk) umma(d1, kdesc(sA1, kk), kdesc(stage(s), kk), IDESC, kk != 0);
                umma_commit(d1_full(i & 1));
                if (i >= 1) mma2(i - 1);
            }
            mma2(T - 1);
            umma_commit(d2_full);
        }
    } else {
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int cl = quarter * 32 + lane;                    // chain within the tile = TMEM lane
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        float ll = 0.f;
        for (int i = 0; i < T; ++i) {
            const int64_t n0 = n_begin + (int64_t)i * NT;
            mbar_wait(full(i & 1), (uint32_t)(i >> 1) & 1u);   // acquire the TMA-written y tile
            mbar_wait(d1_full(i & 1), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const float* ys = reinterpret_cast<const float*>(gbase + OFF_STAGE + (i & 1) * STAGE_BYTES + 4 * ATOM);
            const uint32_t d1 = tD1 + (uint32_t)(i & 1) * 128 + lane_sel + (uint32_t)(half * 64);
            const uint32_t rrow = sA2 + (uint32_t)half * ATOM + (uint32_t)(cl >> 3) * 1024 + (uint32_t)(cl & 7) * 128;
            uint32_t packed[32];     // this chain's 64 residuals of the tile, bf16x2
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t zv[16];
                tmem_ld16(d1 + ch * 16, zv);
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float rr[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int nl = half * 64 + ch * 16 + j + u;
                        const float z = __uint_as_float(zv[j + u]);
                        const float yv = ys[nl];
                        const float e = __expf(-fabsf(z));                 // shared by sigmoid and softplus
                        const float inv = __fdividef(1.0f, 1.0f + e);
                        const float sig = z >= 0.f ? inv : e * inv;
                        const bool live = n0 + nl < n_end;
                        if (live) ll += yv * z - (fmaxf(z, 0.f) + __logf(1.0f + e));
                        rr[u] = live ? yv - sig : 0.f;
                    }
                    __nv_bfloat162 p2 = __floats2bfloat162_rn(rr[0], rr[1]);
                    packed[ch * 8 + (j >> 1)] = *reinterpret_cast<uint32_t*>(&p2);
                }
            }
            // the math above overlapped GEMM2 of tile i-1; only now does R have to be free
            mbar_wait(a2_free, (uint32_t)(i + 1) & 1u);
            // 64 k's = eight 16-byte chunks of this chain's row, XOR-swizzled like the TMA would
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) {
                const uint32_t addr = rrow + (((uint32_t)kc ^ (uint32_t)(cl & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[4 * kc]),
                             "r"(packed[4 * kc + 1]), "r"(packed[4 * kc + 2]), "r"(packed[4 * kc + 3])
                             : "memory");
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(a2_full);
        }
        // partial gradient of this slice: D2[c, j]
        const int64_t c = c0 + cl;
        if (T > 0) {
            mbar_wait(d2_full, 0);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t gv[16];
                tmem_ld16(tD2 + lane_sel + (uint32_t)(half * 64 + ch * 16), gv);
                if (c < a.C) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int jj = half * 64 + ch * 16 + j;
                        if (jj < a.Dx) a.part_g[((int64_t)split * a.C + c) * a.Dx + jj] = __uint_as_float(gv[j]);
                    }
                }
            }
        } else if (c < a.C) {
            for (int jj = half * 64; jj < half * 64 + 64; ++jj)
                if (jj < a.Dx) a.part_g[((int64_t)split * a.C + c) * a.Dx + jj] = 0.f;
        }
        lls[half * CT + cl] = ll;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0 && c < a.C) a.part_ll[(int64_t)split * a.C + c] = lls[cl] + lls[CT + cl];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ---- operand preparation ------------------------------------------------------------
// X [N, Dx] fp32 -> Xb [Np, 128] bf16, XbT [128, Np] bf16, yp [Np] fp32 (zero padded)
__global__ void k_hlr_prep_x(const float* __restrict__ X, const float* __restrict__ y, int64_t N, int64_t Np, int Dx,
                             __nv_bfloat16* __restrict__ Xb, __nv_bfloat16* __restrict__ XbT,
                             float* __restrict__ yp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np * KJ) return;
    const int64_t n = i / KJ;
    const int j = (int)(i % KJ);
    const float v = (n < N && j < Dx) ? X[n * Dx + j] : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    Xb[i] = b;
    XbT[(int64_t)j * Np + n] = b;
    if (j == 0) yp[n] = n < N ? y[n] : 0.f;
}
// theta [C, D] fp32 -> beta_b [Cp, 128] bf16 (regressor columns only, zero padded)
__global__ void k_hlr_prep_beta(const float* __restrict__ theta, int64_t C, int64_t Cp, int Dx, int D,
                                __nv_bfloat16* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cp * KJ) return;
    const int64_t c = i / KJ;
    const int j = (int)(i % KJ);
    out[i] = __float2bfloat16_rn((c < C && j < Dx) ? theta[c * D + j] : 0.f);
}

}  // namespace hlrtc

using namespace hlrtc;

static int64_t pad128(int64_t x) { return (x + 127) / 128 * 128; }

size_t hlr_tc_model_ws_bytes(const bk_model_desc& d) {
    if (d.kind != BK_MODEL_HIER_LOGREG || d.dtype != BK_F32 || d.dims - 2 > KJ) return 0;
    const size_t Np = (size_t)pad128(d.n_obs);
    return 2 * align_up(Np * KJ * 2, 256) + align_up(Np * 4, 256) + 1024;
}

int hlr_tc_prepare(Model& m, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (hlr_tc_model_ws_bytes(m.d) == 0) return BK_OK;
    const int64_t Np = pad128(m.d.n_obs);
    Arena ar(ws, ws_bytes);
    m.Xb = ar.take<__nv_bfloat16>((size_t)Np * KJ);
    m.XbT = ar.take<__nv_bfloat16>((size_t)Np * KJ);
    m.yp = ar.take<float>((size_t)Np);
    if (!ar.ok()) {
        m.Xb = m.XbT = nullptr; m.yp = nullptr;
        set_error("bk_model_create: workspace too small for the tensor-core operands");
        return BK_E_WORKSPACE;
    }
    const int64_t n = Np * KJ;
    k_hlr_prep_x<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)m.d.X, (const float*)m.d.y, m.d.n_obs, Np,
                                                              (int)m.d.dims - 2, m.Xb, m.XbT, m.yp);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

bool hlr_tc_enabled(const Model& m) {
    const char* e = getenv("BK_DISABLE_TC");
    if (e && e[0] == '1') return false;
    return m.d.kind == BK_MODEL_HIER_LOGREG && m.d.dtype == BK_F32 && m.Xb != nullptr;
}

int hlr_tc_splits(int64_t C, int64_t N) {
    const int64_t ctiles = (C + CT - 1) / CT;
    int64_t want = 148 / ctiles;   // one CTA per SM (200 KB of shared memory): stay within ONE wave
    const int64_t max_split = (N + NT - 1) / NT;
    if (want > max_split) want = max_split;
    return (int)(want < 1 ? 1 : want);
}

size_t hlr_tc_eval_ws_bytes(const Model& m, int64_t C) {
    const int Dx = (int)m.d.dims - 2;
    const int ns = hlr_tc_splits(C, m.d.n_obs);
    return align_up((size_t)pad128(C) * KJ * 2, 256) + align_up((size_t)ns * C * Dx * 4, 256) +
           align_up((size_t)ns * C * 4, 256) + 1024;
}

// partial gradients / log-likelihoods of every observation slice -> part_g, part_ll
int hlr_tc_partial(const Model& m, const float* theta, int64_t C, void* ws, size_t ws_bytes, float** part_g,
                   float** part_ll, int* n_split, cudaStream_t st) {
    const int D = (int)m.d.dims, Dx = D - 2;
    const int64_t N = m.d.n_obs, Np = pad128(N), Cp = pad128(C);
    const int ns = hlr_tc_splits(C, N);
    Arena ar(ws, ws_bytes);
    __nv_bfloat16* bb = ar.take<__nv_bfloat16>((size_t)Cp * KJ);
    float* pg = ar.take<float>((size_t)ns * C * Dx);
    float* pl = ar.take<float>((size_t)ns * C);
    if (!ar.ok()) { set_error("model eval workspace too small (%zu < %zu)", ws_bytes, ar.off); return BK_E_WORKSPACE; }
    k_hlr_prep_beta<<<(unsigned)((Cp * KJ + 255) / 256), 256, 0, st>>>(theta, C, Cp, Dx, D, bb);
    BK_LAUNCH_CHECK();
    CUtensorMap mB, mX, mXT;
    int rc;
    if ((rc = make_map_bf16(&mB, bb, Cp, KJ, KJ, CT))) return rc;
    if ((rc = make_map_bf16(&mX, m.Xb, Np, KJ, KJ, NT))) return rc;
    if ((rc = make_map_bf16(&mXT, m.XbT, KJ, Np, Np, KJ))) return rc;
    static bool attr = false;
    if (!attr) {
        BK_CUDA(cudaFuncSetAttribute(k_hlr_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr = true;
    }
    Args a;
    a.C = C; a.N = N; a.Dx = Dx; a.y = m.yp; a.part_g = pg; a.part_ll = pl;
    int64_t rows = (N + ns - 1) / ns;
    a.rows_per_split = (rows + NT - 1) / NT * NT;
    dim3 grid((unsigned)(Cp / CT), (unsigned)ns);
    prof_begin(BK_PROF_GRAD, st);
    k_hlr_tc<<<grid, THREADS, SMEM_BYTES, st>>>(mB, mX, mXT, a);
    prof_end(BK_PROF_GRAD, st);
    BK_LAUNCH_CHECK();
    *part_g = pg; *part_ll = pl; *n_split = ns;
    return BK_OK;
}

}  // namespace bk
