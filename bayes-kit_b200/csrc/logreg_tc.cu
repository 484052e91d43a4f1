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
// X is streamed ONCE per tile by TMA ([n][j], 128B swizzle, L2-resident: 26 MB at c3)
// through a 4-stage ring: the same shared-memory tile is GEMM1's K-major B operand
// (N = n, K = j) and GEMM2's MN-major B operand (N = j, K = n) -- no transposed copy.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = epilogue
// (k_hlr_tc: 16 warps, lane quarter = warp % 4, observation part = (warp - 2) / 4).
// Operands are bf16: like the dense plugin this kernel serves the leapfrog GRADIENTS
// (any deterministic gradient function keeps the leapfrog map reversible and volume
// preserving).  Without the log-likelihood the pointwise stage is ONE MUFU op per
// logit (sigmoid = 0.5 + 0.5 tanh(z/2), tanh.approx) so it keeps pace with the two
// 128x128x128 MMAs of a tile.
//
// The log density that enters the Metropolis test comes from k_hlr_lp_tc below: the
// same tiling, logits from a 3-pass bf16 split (X_hi b_hi + X_hi b_lo + X_lo b_hi,
// ~2^-16 relative, all three passes into one TMEM accumulator), pointwise
// y z - softplus(z) in fp32 with per-tile sums carried in fp64, no second GEMM.
#include <stdlib.h>

#include "model.h"
#include "tc_ptx.cuh"

namespace bk {

using namespace ptx;

namespace hlrtc {
constexpr int CT = 128, NT = 128, KJ = 128;
constexpr int ATOM = 128 * 64 * 2;                    // one [128 rows x 64 k] swizzle block = 16 KB
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = 2 * ATOM + 1024;          // X tile (2 atoms) + y tile
constexpr int OFF_A1 = 0, OFF_STAGE = 2 * ATOM, OFF_A2 = OFF_STAGE + NSTAGE * STAGE_BYTES;
constexpr int OFF_BARS = OFF_A2 + 4 * ATOM, OFF_LL = OFF_BARS + 256;   // R is double-buffered
constexpr int THREADS = 320;                          // k_hlr_lp_tc: 8 epilogue warps
constexpr int EWARPS = 16, EPARTS = EWARPS / 4;       // k_hlr_tc: 16 epilogue warps (4 observation parts)
#ifndef BK_HLR_EP_ALL
#define BK_HLR_EP_ALL 1
#endif
constexpr bool EP_ALL = BK_HLR_EP_ALL != 0;           // every epilogue warp on every tile (vs two groups on alternate tiles)
constexpr int GTHREADS = 64 + 32 * EWARPS;
constexpr int SMEM_BYTES = OFF_LL + EPARTS * CT * 4 + 1024;
constexpr uint32_t IDESC = idesc_bf16(128, 128);
constexpr uint32_t IDESC_BMN = IDESC | (1u << 16);     // B operand MN-major (bit 16 = b_major)
// MN-major SWIZZLE_128B operand: 64 MN-elements (128 B) contiguous, 8 K-rows 128 B apart form one
// 1024-byte swizzle atom; LBO = byte distance between 64-element MN blocks, SBO = between 8-row K groups
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct Args {
    int64_t C, N;
    int Dx;
    int64_t rows_per_split;     // multiple of NT
    const float* y;             // [Np] zero padded
    float* part_g;              // [n_split, C, Dx]
    float* part_ll;             // [n_split, C]
    int need_ll;                // 0: gradient only (one MUFU op per logit)
    float g_scale;              // 1 (gradient needs no rescaling; kept for experiments)
    int debug;                  // BK_HLR_DEBUG bits (timing experiments): 1 = skip the pointwise math
};

__device__ __forceinline__ float tanh_approx(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}

template <bool NEED_LL>   // compile-time: a per-logit branch would serialise the MUFU latencies
__global__ void __launch_bounds__(GTHREADS, 1)
k_hlr_tc(const __grid_constant__ CUtensorMap mapBeta, const __grid_constant__ CUtensorMap mapX, const Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA1 = base + OFF_A1, sA2 = base + OFF_A2, bars = base + OFF_BARS;
    auto stage = [&](int s) { return base + OFF_STAGE + (uint32_t)s * STAGE_BYTES; };
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (NSTAGE + s); };
    auto d1_full = [&](int b) { return bars + 8u * (2 * NSTAGE + b); };
    auto a2_full = [&](int b) { return bars + 8u * (2 * NSTAGE + 2 + b); };
    auto a2_free = [&](int b) { return bars + 8u * (2 * NSTAGE + 4 + b); };
    const uint32_t d2_full = bars + 8u * (2 * NSTAGE + 6), beta_full = bars + 8u * (2 * NSTAGE + 7),
                   tmem_slot = bars + 8u * (2 * NSTAGE + 8);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(gbase + OFF_BARS + 8 * (2 * NSTAGE + 8));
    float* lls = reinterpret_cast<float*>(gbase + OFF_LL);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c0 = (int64_t)blockIdx.x * CT;
    const int split = blockIdx.y;
    const int64_t n_begin = split * a.rows_per_split;
    const int64_t n_end = n_begin + a.rows_per_split < a.N ? n_begin + a.rows_per_split : a.N;
    const int T = n_end > n_begin ? (int)((n_end - n_begin + NT - 1) / NT) : 0;
    cudaTriggerProgrammaticLaunchCompletion();   // a dependent launch may be scheduled as SMs free up (it still waits for our completion)

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(d1_full(s), 1); mbar_init(a2_full(s), (EP_ALL ? 32 : 16) * EWARPS); mbar_init(a2_free(s), 1); }
        mbar_init(d2_full, 1); mbar_init(beta_full, 1);
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
            // Programmatic dependent launch (the interior-step launches set the attribute): this grid may start
            // while its predecessor in the stream (the finish kernel that writes the beta operand) is still
            // running.  The X / y tiles do not depend on it -- the ring is filled first -- and only the operand
            // load waits for the predecessor's completion.  Everything this kernel writes (part_g / part_ll) is
            // written after MMAs that consumed the operand, i.e. after the wait.  Without the attribute the wait
            // returns at once.
            auto issue_x = [&](int i) {
                const int s = i % NSTAGE;
                const uint32_t ph = (uint32_t)(i / NSTAGE) & 1u;
                const int n0 = (int)(n_begin + (int64_t)i * NT);
                mbar_wait(empty(s), ph ^ 1u);
                mbar_expect_tx(full(s), 2 * ATOM + NT * 4);
                const uint32_t st = stage(s);
                tma_load_2d(st, &mapX, full(s), 0, n0);                 // X[n0.., j 0..63]
                tma_load_2d(st + ATOM, &mapX, full(s), 64, n0);         // X[n0.., j 64..127]
                bulk_load(st + 2 * ATOM, a.y + n0, NT * 4, full(s));
            };
            const int pre = T < NSTAGE ? T : NSTAGE;
            for (int i = 0; i < pre; ++i) issue_x(i);
            cudaGridDependencySynchronize();
            mbar_expect_tx(beta_full, 2 * ATOM);
            tma_load_2d(sA1, &mapBeta, beta_full, 0, (int)c0);
            tma_load_2d(sA1 + ATOM, &mapBeta, beta_full, 64, (int)c0);
            for (int i = pre; i < T; ++i) issue_x(i);
        }
    } else if (warp == 1) {
        if (lane == 0 && T > 0) {
            auto kdesc = [](uint32_t tile, int kk) { return umma_desc<128>(tile + (kk >> 2) * ATOM + (kk & 3) * 32); };
            auto mma2 = [&](int t) {   // G += R_t . X_t   (A = R in smem, B = the X tile read MN-major)
                const int s = t % NSTAGE, b = t & 1;
                mbar_wait(a2_full(b), (uint32_t)(t >> 1) & 1u);
                tc_fence_after();
                const uint32_t xt = stage(s), ra = sA2 + (uint32_t)b * 2 * ATOM;
#pragma unroll
                for (int kk = 0; kk < NT / 16; ++kk)   // 16 observations = two 8-row swizzle atoms = 2048 B
                    umma(tD2, kdesc(ra, kk), umma_desc_mn128(xt + kk * 2048, ATOM, 1024), IDESC_BMN, (t | kk) != 0);
                umma_commit(empty(s));     // stage reusable once GEMM2 has read the tile
                umma_commit(a2_free(b));   // R buffer reusable
            };
            mbar_wait(beta_full, 0);
            for (int i = 0; i < T; ++i) {
                const int s = i % NSTAGE;
                mbar_wait(full(s), (uint32_t)(i / NSTAGE) & 1u);
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
        // 16 epilogue warps = two groups of 8 that take alternate tiles (group g owns D1 buffer g and R
        // buffer g), so the fixed latencies of a tile's hand-offs (TMEM load, proxy fence, mbarrier round
        // trips) overlap with the other group's tile.  TMEM lane quarter = warp % 4 (hardware rule),
        // observation half = bit 2 of (warp - 2).
        // EP_ALL (default): all 16 warps work on EVERY tile, 32 observations per thread -- the hand-over latency of a
        // tile (d1_full -> residuals in shared memory -> a2_full) halves and stays below the 0.9 us the tensor pipe
        // has queued behind it (GEMM2 of the previous tile + GEMM1 of the next).  !EP_ALL: the r1 arrangement.
        const int quarter = warp & 3;
        const int grp = EP_ALL ? 0 : (warp - 2) >> 3, part = EP_ALL ? (warp - 2) >> 2 : ((warp - 2) >> 2) & 1;
        const int part4 = EP_ALL ? part : grp * 2 + part;      // 0..3: slice of the final gradient read
        constexpr int PW = EP_ALL ? NT / 4 : NT / 2;           // observations per thread per tile (32 / 64)
        const int cl = quarter * 32 + lane;                    // chain within the tile = TMEM lane
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        // this chain's row of R: PW k's = PW/8 16-byte chunks, inside atom (part * PW) / 64
        const uint32_t rrow0 = sA2 + (uint32_t)((part * PW) >> 6) * ATOM + (uint32_t)(cl >> 3) * 1024 +
                               (uint32_t)(cl & 7) * 128;
        const int kc0 = ((part * PW) & 63) >> 3;
        float ll = 0.f;
        for (int i = grp; i < T; i += EP_ALL ? 1 : 2) {
            const int64_t n0 = n_begin + (int64_t)i * NT;
            mbar_wait(full(i % NSTAGE), (uint32_t)(i / NSTAGE) & 1u);   // acquire the TMA-written y tile
            mbar_wait(d1_full(i & 1), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const float* ys = reinterpret_cast<const float*>(gbase + OFF_STAGE + (i % NSTAGE) * STAGE_BYTES +
                                                             2 * ATOM) + part * PW;
            const uint32_t d1 = tD1 + (uint32_t)(i & 1) * 128 + lane_sel + (uint32_t)(part * PW);
            uint32_t packed[PW / 2];     // this chain's residuals of the tile, bf16x2
#pragma unroll
            for (int ch = 0; ch < PW / 16; ++ch) {
                uint32_t zv[16];
                tmem_ld16(d1 + ch * 16, zv);
                if (a.debug & 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) packed[ch * 8 + j] = zv[2 * j] & 0x3f803f80u;
                    continue;
                }
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4) {
                    const float4 y4 = *reinterpret_cast<const float4*>(ys + ch * 16 + j4);   // broadcast
                    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
                    float rr[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float z = __uint_as_float(zv[j4 + u]);
                        if constexpr (NEED_LL) {
                            const bool live = n0 + part * PW + ch * 16 + j4 + u < n_end;
                            const float e = __expf(-fabsf(z));             // shared by sigmoid and softplus
                            const float inv = __fdividef(1.0f, 1.0f + e);
                            const float sig = z >= 0.f ? inv : e * inv;
                            if (live) ll += yv[u] * z - (fmaxf(z, 0.f) + __logf(1.0f + e));
                            rr[u] = live ? yv[u] - sig : 0.f;
                        } else {
                            // z holds X.beta / 2 and ys holds y - 1/2 (0 on padded rows, where z = 0 too):
                            // y - sigmoid(2z) = (y - 1/2) - tanh(z)/2 -- one MUFU op and one FMA per logit
                            rr[u] = fmaf(-0.5f, tanh_approx(z), yv[u]);
                        }
                    }
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(rr[0], rr[1]), p1 = __floats2bfloat162_rn(rr[2], rr[3]);
                    packed[ch * 8 + (j4 >> 1)] = *reinterpret_cast<uint32_t*>(&p0);
                    packed[ch * 8 + (j4 >> 1) + 1] = *reinterpret_cast<uint32_t*>(&p1);
                }
            }
            // R is double-buffered: this buffer was last read by GEMM2 of tile i-2
            mbar_wait(a2_free(i & 1), ((uint32_t)(i >> 1) + 1u) & 1u);
            const uint32_t rrow = rrow0 + (uint32_t)(i & 1) * 2 * ATOM;
            // 16-byte chunks of this chain's row, XOR-swizzled like the TMA would
#pragma unroll
            for (int kc = 0; kc < PW / 8; ++kc) {
                const uint32_t addr = rrow + (((uint32_t)(kc0 + kc) ^ (uint32_t)(cl & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[4 * kc]),
                             "r"(packed[4 * kc + 1]), "r"(packed[4 * kc + 2]), "r"(packed[4 * kc + 3])
                             : "memory");
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(a2_full(i & 1));
        }
        // partial gradient of this slice: D2[c, j]
        const int64_t c = c0 + cl;
        if (T > 0) {
            mbar_wait(d2_full, 0);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t gv[16];
                tmem_ld16(tD2 + lane_sel + (uint32_t)(part4 * 32 + ch * 16), gv);
                if (c < a.C) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int jj = part4 * 32 + ch * 16 + j;
                        if (jj < a.Dx) a.part_g[((int64_t)split * a.C + c) * a.Dx + jj] = __uint_as_float(gv[j]) * a.g_scale;
                    }
                }
            }
        } else if (c < a.C) {
            for (int jj = part4 * 32; jj < part4 * 32 + 32; ++jj)
                if (jj < a.Dx) a.part_g[((int64_t)split * a.C + c) * a.Dx + jj] = 0.f;
        }
        if constexpr (NEED_LL) {
            lls[part4 * CT + cl] = ll;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EWARPS) : "memory");
            if (part4 == 0 && c < a.C) {
                float t = 0.f;
#pragma unroll
                for (int q = 0; q < EPARTS; ++q) t += lls[q * CT + cl];
                a.part_ll[(int64_t)split * a.C + c] = t;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ---- split-precision log-likelihood (the density that enters the Metropolis test) --------
constexpr int LP_STAGE_BYTES = 4 * ATOM + 1024;       // X_hi (2 atoms) + X_lo (2 atoms) + y tile
constexpr int LP_OFF_STAGE = 4 * ATOM;                // after beta_hi, beta_lo
constexpr int LP_OFF_BARS = LP_OFF_STAGE + 2 * LP_STAGE_BYTES;
constexpr int LP_OFF_LL = LP_OFF_BARS + 256;
constexpr int LP_SMEM_BYTES = LP_OFF_LL + 2 * CT * 8 + 1024;

__global__ void __launch_bounds__(THREADS, 1)
k_hlr_lp_tc(const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
            const __grid_constant__ CUtensorMap mapXhi, const __grid_constant__ CUtensorMap mapXlo, const Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sBhi = base, sBlo = base + 2 * ATOM, bars = base + LP_OFF_BARS;
    auto stage = [&](int s) { return base + LP_OFF_STAGE + (uint32_t)s * LP_STAGE_BYTES; };
    auto full = [&](int s) { return bars + 8u * s; };
    auto slot_free = [&](int s) { return bars + 8u * (2 + s); };   // epilogue drained stage s AND accumulator s
    auto d1_full = [&](int b) { return bars + 8u * (4 + b); };
    const uint32_t beta_full = bars + 8u * 6, tmem_slot = bars + 8u * 7;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + LP_OFF_BARS + 56);
    double* lls = reinterpret_cast<double*>(gbase + LP_OFF_LL);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c0 = (int64_t)blockIdx.x * CT;
    const int split = blockIdx.y;
    const int64_t n_begin = split * a.rows_per_split;
    const int64_t n_end = n_begin + a.rows_per_split < a.N ? n_begin + a.rows_per_split : a.N;
    const int T = n_end > n_begin ? (int)((n_end - n_begin + NT - 1) / NT) : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(full(s), 1); mbar_init(slot_free(s), 256); mbar_init(d1_full(s), 1); }
        mbar_init(beta_full, 1);
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(beta_full, 4 * ATOM);
            tma_load_2d(sBhi, &mapBhi, beta_full, 0, (int)c0);
            tma_load_2d(sBhi + ATOM, &mapBhi, beta_full, 64, (int)c0);
            tma_load_2d(sBlo, &mapBlo, beta_full, 0, (int)c0);
            tma_load_2d(sBlo + ATOM, &mapBlo, beta_full, 64, (int)c0);
            for (int i = 0; i < T; ++i) {
                const int s = i & 1;
                const int n0 = (int)(n_begin + (int64_t)i * NT);
                mbar_wait(slot_free(s), ((uint32_t)(i >> 1) & 1u) ^ 1u);
                mbar_expect_tx(full(s), 4 * ATOM + NT * 4);
                const uint32_t st = stage(s);
                tma_load_2d(st, &mapXhi, full(s), 0, n0);
                tma_load_2d(st + ATOM, &mapXhi, full(s), 64, n0);
                tma_load_2d(st + 2 * ATOM, &mapXlo, full(s), 0, n0);
                tma_load_2d(st + 3 * ATOM, &mapXlo, full(s), 64, n0);
                bulk_load(st + 4 * ATOM, a.y + n0, NT * 4, full(s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && T > 0) {
            auto kdesc = [](uint32_t tile, int kk) { return umma_desc<128>(tile + (kk >> 2) * ATOM + (kk & 3) * 32); };
            mbar_wait(beta_full, 0);
            for (int i = 0; i < T; ++i) {
                const int s = i & 1;
                mbar_wait(full(s), (uint32_t)(i >> 1) & 1u);    // implies slot_free(s): the producer waited for it
                tc_fence_after();
                const uint32_t d1 = tmem + (uint32_t)s * 128;
                const uint32_t xhi = stage(s), xlo = stage(s) + 2 * ATOM;
                // Z = b_hi X_hi + b_lo X_hi + b_hi X_lo, one accumulator
#pragma unroll
                for (int kk = 0; kk < KJ / 16; ++kk) umma(d1, kdesc(sBhi, kk), kdesc(xhi, kk), IDESC, kk != 0);
#pragma unroll
                for (int kk = 0; kk < KJ / 16; ++kk) umma(d1, kdesc(sBlo, kk), kdesc(xhi, kk), IDESC, 1u);
#pragma unroll
                for (int kk = 0; kk < KJ / 16; ++kk) umma(d1, kdesc(sBhi, kk), kdesc(xlo, kk), IDESC, 1u);
                umma_commit(d1_full(s));
            }
        }
    } else {
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int cl = quarter * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        double ll = 0.0;
        for (int i = 0; i < T; ++i) {
            const int s = i & 1;
            const int64_t n0 = n_begin + (int64_t)i * NT;
            mbar_wait(full(s), (uint32_t)(i >> 1) & 1u);       // acquire the TMA-written y tile
            mbar_wait(d1_full(s), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const float* ys = reinterpret_cast<const float*>(gbase + LP_OFF_STAGE + s * LP_STAGE_BYTES + 4 * ATOM);
            const uint32_t d1 = tmem + (uint32_t)s * 128 + lane_sel + (uint32_t)(half * 64);
            float lt = 0.f;                                     // this tile: 64 terms in fp32
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t zv[16];
                tmem_ld16(d1 + ch * 16, zv);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int nl = half * 64 + ch * 16 + j;
                    const float z = __uint_as_float(zv[j]);
                    const float e = __expf(-fabsf(z));
                    const float t = ys[nl] * z - (fmaxf(z, 0.f) + __logf(1.0f + e));
                    lt += (n0 + nl < n_end) ? t : 0.f;
                }
            }
            ll += (double)lt;
            tc_fence_before();
            mbar_arrive(slot_free(s));                          // stage s and accumulator s may be refilled
        }
        lls[half * CT + cl] = ll;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int64_t c = c0 + cl;
        if (half == 0 && c < a.C) a.part_ll[(int64_t)split * a.C + c] = (float)(lls[cl] + lls[CT + cl]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

// ---- operand preparation ------------------------------------------------------------
// X [N, Dx] fp32 -> Xb [Np, 128] bf16 (+ the residual Xlo), yp [Np] fp32 (zero padded)
__global__ void k_hlr_prep_x(const float* __restrict__ X, const float* __restrict__ y, int64_t N, int64_t Np, int Dx,
                             __nv_bfloat16* __restrict__ Xb, __nv_bfloat16* __restrict__ Xlo,
                             float* __restrict__ yp, float* __restrict__ yh) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np * KJ) return;
    const int64_t n = i / KJ;
    const int j = (int)(i % KJ);
    const float v = (n < N && j < Dx) ? X[n * Dx + j] : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    Xb[i] = b;
    Xlo[i] = __float2bfloat16_rn(v - __bfloat162float(b));
    if (j == 0) {
        yp[n] = n < N ? y[n] : 0.f;
        yh[n] = n < N ? y[n] - 0.5f : 0.f;   // y - 1/2; 0 on padded rows so their residual is exactly 0
    }
}
// theta [C, D] fp32 -> beta_b [Cp, 128] bf16 (regressor columns only, zero padded)
__global__ void k_hlr_prep_beta(const float* __restrict__ theta, int64_t C, int64_t Cp, int Dx, int D,
                                float scale, __nv_bfloat16* __restrict__ out,
                                __nv_bfloat16* __restrict__ out_lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cp * KJ) return;
    const int64_t c = i / KJ;
    const int j = (int)(i % KJ);
    const float v = (c < C && j < Dx) ? scale * theta[c * D + j] : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    out[i] = b;
    if (out_lo) out_lo[i] = __float2bfloat16_rn(v - __bfloat162float(b));
}

}  // namespace hlrtc

using namespace hlrtc;

static int64_t pad128(int64_t x) { return (x + 127) / 128 * 128; }

size_t hlr_tc_model_ws_bytes(const bk_model_desc& d) {
    if (d.kind != BK_MODEL_HIER_LOGREG || d.dtype != BK_F32 || d.dims - 2 > KJ) return 0;
    const size_t Np = (size_t)pad128(d.n_obs);
    return 2 * align_up(Np * KJ * 2, 256) + 2 * align_up(Np * 4, 256) + 1024;
}

int hlr_tc_prepare(Model& m, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (hlr_tc_model_ws_bytes(m.d) == 0) return BK_OK;
    const int64_t Np = pad128(m.d.n_obs);
    Arena ar(ws, ws_bytes);
    m.Xb = ar.take<__nv_bfloat16>((size_t)Np * KJ);
    m.Xlo = ar.take<__nv_bfloat16>((size_t)Np * KJ);
    m.yp = ar.take<float>((size_t)Np);
    m.yh = ar.take<float>((size_t)Np);
    if (!ar.ok()) {
        m.Xb = m.Xlo = nullptr; m.yp = m.yh = nullptr;
        set_error("bk_model_create: workspace too small for the tensor-core operands");
        return BK_E_WORKSPACE;
    }
    const int64_t n = Np * KJ;
    k_hlr_prep_x<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)m.d.X, (const float*)m.d.y, m.d.n_obs, Np,
                                                              (int)m.d.dims - 2, m.Xb, m.Xlo, m.yp, m.yh);
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

constexpr int HLR_MAX_FUSED_STEPS = 32;    // leapfrog steps per persistent launch (counters reserved in the workspace)
static size_t hlr_steps_cnt_words(int64_t C) {
    const int64_t ctiles = (C + CT - 1) / CT;
    return ctiles <= 148 ? (size_t)HLR_MAX_FUSED_STEPS * ctiles * 2 + 64 : 0;
}

size_t hlr_tc_eval_ws_bytes(const Model& m, int64_t C) {
    const int Dx = (int)m.d.dims - 2;
    const int ns = hlr_tc_splits(C, m.d.n_obs);
    return 2 * align_up((size_t)pad128(C) * KJ * 2, 256) + align_up((size_t)ns * C * Dx * 4, 256) +
           align_up((size_t)ns * C * 4, 256) + align_up(hlr_steps_cnt_words(C) * 4, 256) + 1024;
}

// BK_HLR_PDL=0 (diagnostic): plain stream-ordered launches for the interior leapfrog steps
static bool hlr_pdl() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("BK_HLR_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

// one layout of the evaluation workspace for every entry point (the bf16 operand written by one call is read by the next)
struct HlrWs {
    __nv_bfloat16 *bb, *bl;   // operand beta (hi, lo) [Cp, 128]
    float *pg, *pl;           // per-slice partial gradients [ns, C, Dx] / log-likelihoods [ns, C]
    uint32_t* cnt;            // counters of the multi-step launch + error word
    int ns;
    bool ok;
    size_t need;
};
static HlrWs hlr_carve(const Model& m, int64_t C, void* ws, size_t ws_bytes) {
    const int Dx = (int)m.d.dims - 2;
    const int64_t Cp = pad128(C);
    HlrWs w;
    w.ns = hlr_tc_splits(C, m.d.n_obs);
    Arena ar(ws, ws_bytes);
    w.bb = ar.take<__nv_bfloat16>((size_t)Cp * KJ);
    w.bl = ar.take<__nv_bfloat16>((size_t)Cp * KJ);
    w.pg = ar.take<float>((size_t)w.ns * C * Dx);
    w.pl = ar.take<float>((size_t)w.ns * C);
    w.cnt = ar.take<uint32_t>(hlr_steps_cnt_words(C));
    w.ok = ar.ok();
    w.need = ar.off;
    return w;
}

// partial gradients / log-likelihoods of every observation slice -> part_g, part_ll
// mode: HLR_TC_GRAD (gradient partials only), HLR_TC_GRAD_LL (+ bf16-grade log-likelihood),
//       HLR_TC_LP (split-precision log-likelihood only, part_g untouched)
int hlr_tc_partial(const Model& m, const float* theta, int64_t C, void* ws, size_t ws_bytes, float** part_g,
                   float** part_ll, int* n_split, cudaStream_t st, int mode, bool operand_ready,
                   __nv_bfloat16** operand) {
    const int D = (int)m.d.dims, Dx = D - 2;
    const int64_t N = m.d.n_obs, Np = pad128(N), Cp = pad128(C);
    const HlrWs w = hlr_carve(m, C, ws, ws_bytes);
    if (!w.ok) { set_error("model eval workspace too small (%zu < %zu)", ws_bytes, w.need); return BK_E_WORKSPACE; }
    const int ns = w.ns;
    __nv_bfloat16 *bb = w.bb, *bl = w.bl;
    float *pg = w.pg, *pl = w.pl;
    if (operand) *operand = bb;
    if (!operand_ready) {
        // gradient-only mode folds the 1/2 of sigmoid(z) = 1/2 + tanh(z/2)/2 into the operand (exact in bf16)
        k_hlr_prep_beta<<<(unsigned)((Cp * KJ + 255) / 256), 256, 0, st>>>(theta, C, Cp, Dx, D,
                                                                           mode == HLR_TC_GRAD ? 0.5f : 1.0f, bb,
                                                                           mode == HLR_TC_LP ? bl : nullptr);
        BK_LAUNCH_CHECK();
    }
    CUtensorMap mB, mBl, mX, mXl;
    int rc;
    if ((rc = make_map_bf16(&mB, bb, Cp, KJ, KJ, CT))) return rc;
    if ((rc = make_map_bf16(&mBl, bl, Cp, KJ, KJ, CT))) return rc;
    if ((rc = make_map_bf16(&mX, m.Xb, Np, KJ, KJ, NT))) return rc;
    if ((rc = make_map_bf16(&mXl, m.Xlo, Np, KJ, KJ, NT))) return rc;
    static bool attr = false;
    if (!attr) {
        BK_CUDA(cudaFuncSetAttribute(k_hlr_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        BK_CUDA(cudaFuncSetAttribute(k_hlr_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        BK_CUDA(cudaFuncSetAttribute(k_hlr_lp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, LP_SMEM_BYTES));
        attr = true;
    }
    Args a;
    a.C = C; a.N = N; a.Dx = Dx; a.part_g = pg; a.part_ll = pl;
    a.y = mode == HLR_TC_GRAD ? m.yh : m.yp;
    a.need_ll = mode == HLR_TC_GRAD_LL ? 1 : 0;
    a.g_scale = 1.0f;
    { const char* e = getenv("BK_HLR_DEBUG"); a.debug = e ? atoi(e) : 0; }
    int64_t rows = (N + ns - 1) / ns;
    a.rows_per_split = (rows + NT - 1) / NT * NT;
    dim3 grid((unsigned)(Cp / CT), (unsigned)ns);
    prof_begin(BK_PROF_GRAD, st);
    if (mode == HLR_TC_LP) k_hlr_lp_tc<<<grid, THREADS, LP_SMEM_BYTES, st>>>(mB, mBl, mX, mXl, a);
    else if (a.need_ll) k_hlr_tc<true><<<grid, GTHREADS, SMEM_BYTES, st>>>(mB, mX, a);
    else if (mode == HLR_TC_GRAD && hlr_pdl()) {
        // interior-step gradient: programmatic dependent launch behind the finish kernel (see the producer warp)
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = grid; cfg.blockDim = dim3(GTHREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        BK_CUDA(cudaLaunchKernelEx(&cfg, k_hlr_tc<false>, mB, mX, a));
    }
    else k_hlr_tc<false><<<grid, GTHREADS, SMEM_BYTES, st>>>(mB, mX, a);
    prof_end(BK_PROF_GRAD, st);
    BK_LAUNCH_CHECK();
    *part_g = pg; *part_ll = pl; *n_split = ns;
    return BK_OK;
}

// ---- interior leapfrog step, fused around the tensor-core gradient ---------------------------
// One warp per chain: fixed-order sum of the observation slices + hierarchical prior terms (the
// gradient), the leapfrog kick and drift  r += eps m g ; q += eps r  (hmc.py:48-49), and the bf16
// operand (beta / 2, zero padded) of the NEXT gradient launch -- instead of three kernels
// (finish, step, operand preparation) and two round trips of the gradient through HBM.
// The partial sums of all four dimensions of a lane (up to 64 loads) are requested before any is consumed: the finish
// sits on the critical path of every leapfrog step.  Slices are summed in a fixed
// order (four interleaved accumulators over the full groups of 4, the remainder into the first), independent of
// which kernel calls it.
__device__ __forceinline__ void hlr_finish_chain(float* __restrict__ th, float* __restrict__ rr,
                                                 const float* __restrict__ pg_chain, int64_t step, int Dx,
                                                 int n_split, float eps, const float* __restrict__ metric,
                                                 __nv_bfloat16* __restrict__ oprow, int lane, bool poison) {
    const float mu = th[Dx], lam = th[Dx + 1];
    const float e2 = expf(-2.f * lam), ep2 = expf(2.f * lam);
    float sr = 0.f, ss = 0.f;
    const int ns4 = n_split & ~3;
    constexpr int NJ = KJ / 32;                       // dimensions per lane (4)
    float g[NJ][4];
#pragma unroll
    for (int h = 0; h < NJ; ++h) { g[h][0] = g[h][1] = g[h][2] = g[h][3] = 0.f; }
    // all four dimensions of a lane travel together: 64 loads in flight per lane and chunk of 16 slices
#pragma unroll 1
    for (int s0 = 0; s0 < n_split; s0 += 16) {
        float av[NJ][16];
#pragma unroll
        for (int h = 0; h < NJ; ++h) {
            const int j = lane + 32 * h;
#pragma unroll
            for (int u = 0; u < 16; ++u)
                av[h][u] = (j < Dx && s0 + u < n_split) ? __ldcg(pg_chain + j + (int64_t)(s0 + u) * step) : 0.f;
        }
#pragma unroll
        for (int h = 0; h < NJ; ++h) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                if (s0 + u < ns4) g[h][u & 3] += av[h][u];
                else if (s0 + u < n_split) g[h][0] += av[h][u];
            }
        }
    }
#pragma unroll
    for (int h = 0; h < NJ; ++h) {
        const int j = lane + 32 * h;
        float bq = 0.f;
        if (j < Dx) {
            const float d = th[j] - mu;
            sr += d;
            ss = fmaf(d, d, ss);
            const float gs = ((g[h][0] + g[h][1]) + (g[h][2] + g[h][3])) - e2 * d;
            const float rn = fmaf(eps * (metric ? metric[j] : 1.f), gs, rr[j]);
            rr[j] = rn;
            bq = fmaf(eps, rn, th[j]);
            if (poison) bq = __int_as_float(0x7fc00000);
            th[j] = bq;
        }
        oprow[j] = __float2bfloat16_rn(0.5f * bq);
    }
    sr = warp_sum(sr);
    ss = warp_sum(ss);
    if (lane == 0) {
        const float gmu = e2 * sr - mu, glam = -(float)Dx + e2 * ss - ep2 + 1.f;
        const float r0 = fmaf(eps * (metric ? metric[Dx] : 1.f), gmu, rr[Dx]);
        const float r1 = fmaf(eps * (metric ? metric[Dx + 1] : 1.f), glam, rr[Dx + 1]);
        rr[Dx] = r0; rr[Dx + 1] = r1;
        th[Dx] = fmaf(eps, r0, mu);
        th[Dx + 1] = fmaf(eps, r1, lam);
    }
}

__global__ void k_hlr_finish_step(float* __restrict__ q, float* __restrict__ r, const float* __restrict__ part_g,
                                  int64_t C, int Dx, int D, int n_split, float eps, const float* __restrict__ metric,
                                  __nv_bfloat16* __restrict__ operand) {
    cudaTriggerProgrammaticLaunchCompletion();   // the next gradient launch may start its prologue / X prefetch
    cudaGridDependencySynchronize();             // the partial gradients of the launch before us are complete
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    hlr_finish_chain(q + c * D, r + c * D, part_g + c * Dx, C * (int64_t)Dx, Dx, n_split, eps, metric,
                     operand + c * KJ, lane, false);
}

// ---- all interior leapfrog steps of a trajectory in ONE persistent launch ----------------------------------------
// k_hlr_tc<false> + k_hlr_finish_step per leapfrog step are two launches with an ~18 us fixed cost (prologue,
// pipeline fill, drain, launch gap) around ~45 us of tensor work at c3.  Here the grid (chain tiles x observation
// slices, one CTA per SM, all co-resident: cooperative launch) stays up for n_steps leapfrog steps.  Per step a CTA
//   1. streams its observation slice through the same GEMM1 -> residual -> GEMM2 pipeline and writes its partial
//      gradient [128 chains, Dx] (the X tiles of step s+1 are already prefetched into the ring while step s drains);
//   2. counts itself in on cnt[s][chain tile][0] and waits until all n_split slices of its chain tile are there;
//   3. finishes 128 / n_split of the tile's chains (warp per chain, hlr_finish_chain: fixed-order slice sum, prior
//      terms, kick, drift, next bf16 operand) and counts itself in on cnt[s][chain tile][1];
//   4. its TMA producer waits for that counter to reach n_split, then loads the new operand (generic-proxy writes of
//      other CTAs -> async-proxy read: fence.proxy.async on both sides of the release / acquire pair).
// A chain is always finished by the same warp of the same CTA, so q / r need no cross-CTA ordering; the partials are
// read with ld.global.cg (they are rewritten every step by other SMs).  Results are bit-identical to the per-step
// launches (same device function, same summation order).  Waits are bounded (~4 s, then the error word poisons the
// positions with NaN).
struct StepsArgs {
    Args g;
    float* q;                  // [C, D]
    float* r;                  // [C, D]
    int D;
    float eps;
    const float* metric;       // [D] or NULL
    __nv_bfloat16* operand;    // [Cp, 128] beta / 2 of the current q on entry
    int n_steps;
    uint32_t* cnt;             // [n_steps][gridDim.x][2], zeroed by the host; then the error word
    uint32_t* err;
};

__device__ __forceinline__ void hlr_red_release_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t hlr_ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until *flag >= want; false (and the error word set) after ~4 s or when another waiter already gave up
__device__ __forceinline__ bool hlr_wait_count(const uint32_t* flag, uint32_t want, uint32_t* err) {
    if (hlr_ld_acquire(flag) >= want) return true;
    uint64_t t_start;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    unsigned spins = 0;
    while (hlr_ld_acquire(flag) < want) {
        __nanosleep(32);
        if ((++spins & 1023u) == 0) {
            uint64_t t_now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_now));
            if (t_now - t_start > 4000000000ull || *(volatile uint32_t*)err) {
                atomicExch(err, 1u);
                return false;
            }
        }
    }
    return true;
}

__global__ void __launch_bounds__(GTHREADS, 1)
k_hlr_tc_steps(const __grid_constant__ CUtensorMap mapBeta, const __grid_constant__ CUtensorMap mapX,
               const StepsArgs sa) {
    extern __shared__ uint8_t smem_raw[];
    const Args& a = sa.g;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA1 = base + OFF_A1, sA2 = base + OFF_A2, bars = base + OFF_BARS;
    auto stage = [&](int s) { return base + OFF_STAGE + (uint32_t)s * STAGE_BYTES; };
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (NSTAGE + s); };
    auto d1_full = [&](int b) { return bars + 8u * (2 * NSTAGE + b); };
    auto a2_full = [&](int b) { return bars + 8u * (2 * NSTAGE + 2 + b); };
    auto a2_free = [&](int b) { return bars + 8u * (2 * NSTAGE + 4 + b); };
    const uint32_t d2_full = bars + 8u * (2 * NSTAGE + 6), beta_full = bars + 8u * (2 * NSTAGE + 7),
                   tmem_slot = bars + 8u * (2 * NSTAGE + 8), d2_free = bars + 8u * (2 * NSTAGE + 9);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(gbase + OFF_BARS + 8 * (2 * NSTAGE + 8));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c0 = (int64_t)blockIdx.x * CT;
    const int split = blockIdx.y, n_split = (int)gridDim.y;
    const int64_t n_begin = split * a.rows_per_split;
    const int64_t n_end = n_begin + a.rows_per_split < a.N ? n_begin + a.rows_per_split : a.N;
    const int T = n_end > n_begin ? (int)((n_end - n_begin + NT - 1) / NT) : 0;
    auto cnt_of = [&](int s) { return sa.cnt + ((int64_t)s * gridDim.x + blockIdx.x) * 2; };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(d1_full(s), 1); mbar_init(a2_full(s), 16 * EWARPS); mbar_init(a2_free(s), 1); }
        mbar_init(d2_full, 1); mbar_init(beta_full, 1); mbar_init(d2_free, EWARPS);
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
            auto issue_x = [&](int s, int i) {
                const int gi = s * T + i, sl = gi % NSTAGE;
                const int n0 = (int)(n_begin + (int64_t)i * NT);
                mbar_wait(empty(sl), ((uint32_t)(gi / NSTAGE) & 1u) ^ 1u);
                mbar_expect_tx(full(sl), 2 * ATOM + NT * 4);
                const uint32_t st = stage(sl);
                tma_load_2d(st, &mapX, full(sl), 0, n0);
                tma_load_2d(st + ATOM, &mapX, full(sl), 64, n0);
                bulk_load(st + 2 * ATOM, a.y + n0, NT * 4, full(sl));
            };
            for (int s = 0; s < sa.n_steps; ++s) {
                // the X tiles do not depend on the step: fill the ring while the previous step drains / finishes
                const int pre = T < NSTAGE ? T : NSTAGE;
                for (int i = 0; i < pre; ++i) issue_x(s, i);
                if (s > 0) {
                    // every slice of this chain tile has written its rows of the new operand; all MMAs of the
                    // previous step (the readers of sA1) completed before this CTA counted itself in
                    hlr_wait_count(cnt_of(s - 1) + 1, (uint32_t)n_split, sa.err);
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                if (T > 0) {
                    mbar_expect_tx(beta_full, 2 * ATOM);
                    tma_load_2d(sA1, &mapBeta, beta_full, 0, (int)c0);
                    tma_load_2d(sA1 + ATOM, &mapBeta, beta_full, 64, (int)c0);
                }
                for (int i = pre; i < T; ++i) issue_x(s, i);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && T > 0) {
            auto kdesc = [](uint32_t tile, int kk) { return umma_desc<128>(tile + (kk >> 2) * ATOM + (kk & 3) * 32); };
            for (int s = 0; s < sa.n_steps; ++s) {
                auto mma2 = [&](int t) {   // G += R_t . X_t   (A = R in smem, B = the X tile read MN-major)
                    const int gi = s * T + t, sl = gi % NSTAGE, b = gi & 1;
                    mbar_wait(a2_full(b), (uint32_t)(gi >> 1) & 1u);
                    tc_fence_after();
                    const uint32_t xt = stage(sl), ra = sA2 + (uint32_t)b * 2 * ATOM;
#pragma unroll
                    for (int kk = 0; kk < NT / 16; ++kk)
                        umma(tD2, kdesc(ra, kk), umma_desc_mn128(xt + kk * 2048, ATOM, 1024), IDESC_BMN, (t | kk) != 0);
                    umma_commit(empty(sl));
                    umma_commit(a2_free(b));
                };
                mbar_wait(beta_full, (uint32_t)s & 1u);
                if (s > 0) mbar_wait(d2_free, (uint32_t)(s - 1) & 1u);   // the epilogue has read the previous D2
                tc_fence_after();
                for (int i = 0; i < T; ++i) {
                    const int gi = s * T + i, sl = gi % NSTAGE;
                    mbar_wait(full(sl), (uint32_t)(gi / NSTAGE) & 1u);
                    tc_fence_after();
                    const uint32_t d1 = tD1 + (uint32_t)(gi & 1) * 128;
#pragma unroll
                    for (int kk = 0; kk < KJ / 16; ++kk) umma(d1, kdesc(sA1, kk), kdesc(stage(sl), kk), IDESC, kk != 0);
                    umma_commit(d1_full(gi & 1));
                    if (i >= 1) mma2(i - 1);
                }
                mma2(T - 1);
                umma_commit(d2_full);
            }
        }
    } else {
        const int ew = warp - 2;                               // epilogue warp 0..15
        const int quarter = warp & 3, grp = ew >> 3, part = (ew >> 2) & 1;
        const int part4 = grp * 2 + part;
        constexpr int PW = NT / 2;
        const int cl = quarter * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        const uint32_t rrow0 = sA2 + (uint32_t)((part * PW) >> 6) * ATOM + (uint32_t)(cl >> 3) * 1024 +
                               (uint32_t)(cl & 7) * 128;
        const int kc0 = ((part * PW) & 63) >> 3;
        const int64_t c = c0 + cl;
        const bool elected = threadIdx.x == 64;                // first epilogue thread: the CTA's voice on the counters
        for (int s = 0; s < sa.n_steps; ++s) {
            for (int i = 0; i < T; ++i) {
                const int gi = s * T + i;
                if ((gi & 1) != grp) continue;                 // the two groups of 8 warps take alternate tiles
                const int sl = gi % NSTAGE;
                mbar_wait(full(sl), (uint32_t)(gi / NSTAGE) & 1u);
                mbar_wait(d1_full(gi & 1), (uint32_t)(gi >> 1) & 1u);
                tc_fence_after();
                const float* ys = reinterpret_cast<const float*>(gbase + OFF_STAGE + sl * STAGE_BYTES + 2 * ATOM) + part * PW;
                const uint32_t d1 = tD1 + (uint32_t)(gi & 1) * 128 + lane_sel + (uint32_t)(part * PW);
                uint32_t packed[PW / 2];
#pragma unroll
                for (int ch = 0; ch < PW / 16; ++ch) {
                    uint32_t zv[16];
                    tmem_ld16(d1 + ch * 16, zv);
#pragma unroll
                    for (int j4 = 0; j4 < 16; j4 += 4) {
                        const float4 y4 = *reinterpret_cast<const float4*>(ys + ch * 16 + j4);
                        const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
                        float rr[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) rr[u] = fmaf(-0.5f, tanh_approx(__uint_as_float(zv[j4 + u])), yv[u]);
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(rr[0], rr[1]), p1 = __floats2bfloat162_rn(rr[2], rr[3]);
                        packed[ch * 8 + (j4 >> 1)] = *reinterpret_cast<uint32_t*>(&p0);
                        packed[ch * 8 + (j4 >> 1) + 1] = *reinterpret_cast<uint32_t*>(&p1);
                    }
                }
                mbar_wait(a2_free(gi & 1), ((uint32_t)(gi >> 1) + 1u) & 1u);
                const uint32_t rrow = rrow0 + (uint32_t)(gi & 1) * 2 * ATOM;
#pragma unroll
                for (int kc = 0; kc < PW / 8; ++kc) {
                    const uint32_t addr = rrow + (((uint32_t)(kc0 + kc) ^ (uint32_t)(cl & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[4 * kc]),
                                 "r"(packed[4 * kc + 1]), "r"(packed[4 * kc + 2]), "r"(packed[4 * kc + 3])
                                 : "memory");
                }
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(a2_full(gi & 1));
            }
            // ---- partial gradient of this slice: D2[c, j] -> part_g[split] ----
            if (T > 0) {
                mbar_wait(d2_full, (uint32_t)s & 1u);
                tc_fence_after();
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
                    uint32_t gv[16];
                    tmem_ld16(tD2 + lane_sel + (uint32_t)(part4 * 32 + ch * 16), gv);
                    if (c < a.C) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int jj = part4 * 32 + ch * 16 + j;
                            if (jj < a.Dx) a.part_g[((int64_t)split * a.C + c) * a.Dx + jj] = __uint_as_float(gv[j]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d2_free);           // D2 may be overwritten by the next step's GEMM2
            } else if (c < a.C && s == 0) {
                for (int jj = part4 * 32; jj < part4 * 32 + 32; ++jj)
                    if (jj < a.Dx) a.part_g[((int64_t)split * a.C + c) * a.Dx + jj] = 0.f;
            }
            // ---- all slices of this chain tile in? ----
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EWARPS) : "memory");
            if (elected) {
                __threadfence();
                hlr_red_release_add(cnt_of(s), 1u);
                hlr_wait_count(cnt_of(s), (uint32_t)n_split, sa.err);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EWARPS) : "memory");
            // ---- finish + kick + drift + next operand for this CTA's share of the tile's chains ----
            const bool poison = *(volatile uint32_t*)sa.err != 0u;
            for (int k = split + n_split * ew; k < CT; k += n_split * EWARPS) {
                const int64_t cc = c0 + k;
                if (cc < a.C)
                    hlr_finish_chain(sa.q + cc * sa.D, sa.r + cc * sa.D, a.part_g + cc * a.Dx, a.C * (int64_t)a.Dx, a.Dx,
                                     n_split, sa.eps, sa.metric, sa.operand + cc * KJ, lane, poison);
            }
            asm volatile("fence.proxy.async;" ::: "memory");   // operand rows: generic-proxy writes -> TMA reads elsewhere
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EWARPS) : "memory");
            if (elected) hlr_red_release_add(cnt_of(s) + 1, 1u);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int hlr_tc_interior_step(const Model& m, float* q, float* r, int64_t C, float eps, const float* metric,
                         bool operand_ready, void* ws, size_t ws_bytes, cudaStream_t st) {
    float *pg, *pl;
    int ns;
    __nv_bfloat16* bb;
    int rc = hlr_tc_partial(m, q, C, ws, ws_bytes, &pg, &pl, &ns, st, HLR_TC_GRAD, operand_ready, &bb);
    if (rc) return rc;
    const int D = (int)m.d.dims;
    if (hlr_pdl()) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)((C * 32 + 63) / 64)); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        BK_CUDA(cudaLaunchKernelEx(&cfg, k_hlr_finish_step, q, r, (const float*)pg, C, D - 2, D, ns, eps, metric, bb));
    } else {
        k_hlr_finish_step<<<(unsigned)((C * 32 + 63) / 64), 64, 0, st>>>(q, r, pg, C, D - 2, D, ns, eps, metric, bb);
    }
    BK_LAUNCH_CHECK();
    return BK_OK;
}

// n_steps interior leapfrog steps (gradient at q, kick, drift) in persistent launches of up to HLR_MAX_FUSED_STEPS
// steps each -- only with BK_HLR_FUSE=1: MEASURED SLOWER than the launch pairs (c3: 1.09 vs 0.91 ms per draw,
// profiles/r2_hlr_fuse_check.log; draws bit-identical).  A leapfrog step is a true dependency chain per chain tile --
// drain the MMA pipeline, all slices in, finish, operand published, operand reloaded, pipeline refilled -- and with one
// chain tile per CTA nothing overlaps it; two kernel boundaries cost less than two counter hand-overs plus the finish
// on the critical path.  Default: one gradient launch + one finish launch per step.
int hlr_tc_interior_steps(const Model& m, float* q, float* r, int64_t C, float eps, const float* metric, int n_steps,
                          void* ws, size_t ws_bytes, cudaStream_t st) {
    if (n_steps <= 0) return BK_OK;
    static int fuse_env = -1;
    if (fuse_env < 0) { const char* e = getenv("BK_HLR_FUSE"); fuse_env = (e && e[0] == '1') ? 1 : 0; }
    const int D = (int)m.d.dims, Dx = D - 2;
    const int64_t N = m.d.n_obs, Np = pad128(N), Cp = pad128(C);
    const HlrWs w = hlr_carve(m, C, ws, ws_bytes);
    if (!w.ok) { set_error("model eval workspace too small (%zu < %zu)", ws_bytes, w.need); return BK_E_WORKSPACE; }
    // co-residency: one CTA per SM (200 KB of shared memory), the grid must fit the device in one wave
    static int sms_dev[64] = {0}, occ_dev[64] = {0};
    int dev = 0;
    BK_CUDA(cudaGetDevice(&dev));
    bool fused = fuse_env != 0 && hlr_steps_cnt_words(C) > 0 && dev >= 0 && dev < 64;
    if (fused && !sms_dev[dev]) {
        int sms = 0, occ = 0;
        BK_CUDA(cudaFuncSetAttribute(k_hlr_tc_steps, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_hlr_tc_steps, GTHREADS, SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError();
            sms = 1; occ = 0;
        }
        sms_dev[dev] = sms; occ_dev[dev] = occ;
    }
    if (fused && (Cp / CT) * (int64_t)w.ns > (int64_t)sms_dev[dev] * occ_dev[dev]) fused = false;
    if (!fused) {
        for (int s = 0; s < n_steps; ++s) {
            const int rc = hlr_tc_interior_step(m, q, r, C, eps, metric, s > 0, ws, ws_bytes, st);
            if (rc) return rc;
        }
        return BK_OK;
    }
    k_hlr_prep_beta<<<(unsigned)((Cp * KJ + 255) / 256), 256, 0, st>>>(q, C, Cp, Dx, D, 0.5f, w.bb, nullptr);
    BK_LAUNCH_CHECK();
    CUtensorMap mB, mX;
    int rc;
    if ((rc = make_map_bf16(&mB, w.bb, Cp, KJ, KJ, CT))) return rc;
    if ((rc = make_map_bf16(&mX, m.Xb, Np, KJ, KJ, NT))) return rc;
    StepsArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.g.C = C; sa.g.N = N; sa.g.Dx = Dx; sa.g.part_g = w.pg; sa.g.part_ll = w.pl;
    sa.g.y = m.yh; sa.g.need_ll = 0; sa.g.g_scale = 1.0f; sa.g.debug = 0;
    const int64_t rows = (N + w.ns - 1) / w.ns;
    sa.g.rows_per_split = (rows + NT - 1) / NT * NT;
    sa.q = q; sa.r = r; sa.D = D; sa.eps = eps; sa.metric = metric; sa.operand = w.bb;
    const size_t cnt_words = hlr_steps_cnt_words(C);
    sa.cnt = w.cnt; sa.err = w.cnt + (cnt_words - 1);
    dim3 grid((unsigned)(Cp / CT), (unsigned)w.ns);
    for (int done = 0; done < n_steps;) {
        const int ns_now = n_steps - done < HLR_MAX_FUSED_STEPS ? n_steps - done : HLR_MAX_FUSED_STEPS;
        sa.n_steps = ns_now;
        BK_CUDA(cudaMemsetAsync(w.cnt, 0, cnt_words * sizeof(uint32_t), st));
        void* args[3] = {(void*)&mB, (void*)&mX, (void*)&sa};
        prof_begin(BK_PROF_GRAD, st);
        // cooperative: every CTA spins on counters the others advance, so the whole grid must be resident
        BK_CUDA(cudaLaunchCooperativeKernel((const void*)k_hlr_tc_steps, grid, dim3(GTHREADS), args, SMEM_BYTES, st));
        prof_end(BK_PROF_GRAD, st);
        count_launch();
        done += ns_now;
    }
    return BK_OK;
}

}  // namespace bk
