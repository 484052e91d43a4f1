// Dense-precision Gaussian on the 5th-gen tensor cores (fp32 timed mode).
//
// One leapfrog step of C chains is   G^T[d, c] = sum_k P[d, k] * Theta[c, k]
// (M = dims, N = chains, K = dims): a TMA-fed tcgen05 GEMM with fp32
// accumulators in TMEM.  A = bf16(P) tile [128 dims x 64 k], B = bf16(theta)
// tile [256 chains x 64 k], both K-major with the 128-byte swizzle, staged by a
// 4-deep TMA/mbarrier ring.  TMEM lanes are DIMS and TMEM columns are CHAINS,
// so in the epilogue a warp's 32 lanes touch 32 consecutive dims of one chain:
// every global access of the fused leapfrog update is a coalesced 128-byte row
// segment.  The update is the leapfrog recursion in POSITION form (Stormer-
// Verlet): kick r += eps*m*g and drift q += eps*r combine to
//        q_{n+1} = q_n + ((q_n - q_{n-1}) + eps^2 * m * g(q_n))
// so the momentum never exists in memory: a step reads q_n, q_{n-1} and writes
// q_{n+1} over q_{n-1} (12 B fp32 per element instead of 16) plus the bf16
// operand for the next GEMM; eps*r_{n+1/2} = q_{n+1} - q_n is recovered at the
// trajectory end, where the Hamiltonian needs it.  The accumulator is double-buffered (2 x 256 TMEM columns) so the
// epilogue of tile i overlaps the MMAs of tile i+1.
//
// Precision: interior leapfrog gradients use bf16 operands (the leapfrog map
// stays a volume-preserving, reversible shear for ANY deterministic gradient
// function); the endpoint gradient that enters the Hamiltonian is computed
// with a 3-pass bf16 split (P_hi*q_hi + P_hi*q_lo + P_lo*q_hi, ~2^-16
// relative) so the Metropolis test sees fp32-accurate energies.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread MMA issuer, warps 2..9 = epilogue (TMEM lane quarter = warp%4,
// column half = (warp-2)/4), two 16-chain chunks of r/q loads in flight per warp.
// Work arrays are tile-/box-blocked so each CTA streams contiguous runs.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "dense_tc.h"
#include "sep_common.cuh"

namespace bk {

namespace tc {

constexpr int BM = 128;      // dims per tile  (UMMA M, TMEM lanes)
constexpr int BN = 256;      // chains per tile (UMMA N, TMEM columns)
constexpr int BK = 64;       // k per stage (128 bytes of bf16 = one swizzle atom row)
constexpr int UK = 16;       // UMMA K for bf16
// GRAD mode is tensor-bound: 4 operand stages.  STEP mode is HBM-bound on its epilogue:
// 2 operand stages, and the freed 128 KB is a per-warp cp.async ring that keeps
// EDEPTH-1 chunks (16 chains x 32 dims of r and q = 4 KB) per epilogue warp in flight.
constexpr int CWID = 16;                               // chains per epilogue chunk
constexpr int EDEPTH_MAX = 4;                          // ring depth per epilogue warp (Cfg::EDEPTH)
constexpr int ECHUNK_BYTES = 2 * CWID * 32 * 4;        // r + q of one chunk of one warp
template <int MODE, bool PAIR = false> struct Cfg {
    // STEP stages are HALF k-blocks (32 k = 64-byte rows, SWIZZLE_64B): 4 x 24 KB keeps the same
    // 96 KB of operands in flight as 2 x 48 KB but with twice the pipeline granularity.
    static constexpr int KS = MODE == TC_MODE_STEP ? 32 : 64;          // k per stage
    // pair + STEP: the GEMM side is bound by operand bytes in flight (TMA latency), so it trades one ring
    // slot per epilogue warp (32 KB) for four more 16 KB operand stages
#ifndef BK_STEP_STAGES
#define BK_STEP_STAGES 8
#define BK_STEP_EDEPTH 3
#endif
    static constexpr int STAGES = (PAIR && MODE == TC_MODE_STEP) ? BK_STEP_STAGES : 4;
    static constexpr int EDEPTH = (PAIR && MODE == TC_MODE_STEP) ? BK_STEP_EDEPTH : EDEPTH_MAX;
    static constexpr int A_BYTES = BM * KS * 2;
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * KS * 2;   // pair: each CTA stages half the chains
    static constexpr int SMEM_TILES = STAGES * (A_BYTES + B_BYTES);
    static constexpr int ERING = MODE == TC_MODE_STEP ? 8 * EDEPTH * ECHUNK_BYTES : 0;
    static constexpr int SMEM_BYTES = SMEM_TILES + ERING + 256 + 1024;  // + barriers + alignment slack
};
constexpr int EPI_WARPS = 8;            // 2 per TMEM lane quarter (column halves); ring sized for 8
// + one signalling warp (fused STEP launches: publishes finished tiles to the other clusters, so the
// gpu-scope release -- which waits for the SM's outstanding stores -- never stalls a streaming warp)
constexpr int SIG_WARP = 2 + EPI_WARPS;
constexpr int THREADS = 64 + 32 * EPI_WARPS + 32;
constexpr int SIG_SLOTS = 4;
constexpr uint32_t TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
        : "memory");
}
// ---- CTA pair (cta_group::2) helpers: the two CTAs of a cluster run ONE 256 x 256 MMA; each CTA
// stages its 128 rows of A and HALF of B (128 of the 256 chains), the leader (rank 0) issues
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {   // same smem offset in CTA `rank`
    uint32_t d;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(saddr), "r"(rank));
    return d;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (.release.cta): the cluster-scope release form costs a MEMBAR.ALL + ERRBAR that waits for
    // every outstanding global store of the warp; the consumer (MMA issuer) only needs the TMEM reads to be
    // complete, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync guarantee
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into OWN shared memory whose bytes are credited to the barrier `bar_cluster` (leader's)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on `bar`'s offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major swizzled operand tile: rows of ROWB bytes (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B),
// 8-row groups 8*ROWB bytes apart
template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);            // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8u * ROWB) >> 4) << 32;              // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
    d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;           // layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4
    return d;
}
// kind::f16, A = B = bf16, D = f32, both K-major, M = 128, N = 256
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
// CTA pair: M = 256 (two CTAs x 128 dims), N = 256
constexpr uint32_t IDESC_PAIR = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                ((uint32_t)((2 * BM) >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Internal fp32 work arrays (r, q, endpoint gradient) are TILE-BLOCKED to match the
// epilogue's access pattern: element (chain c, dim d) lives at
//   ((c / BN) * m_tiles + d / BM) * (BN * BM) + (c % BN) * BM + (d % BM)
// so the 256 x 128 patch a CTA updates is one contiguous 128 KB run (DRAM-page
// friendly streaming instead of 128-byte pieces at a 4000-byte stride).
__host__ __device__ __forceinline__ int64_t blk_index(int64_t c, int d, int m_tiles) {
    return (((c / BN) * m_tiles + (d / BM)) * (int64_t)(BN * BM)) + (c % BN) * BM + (d % BM);
}

// bf16 MMA operands are stored BOX-BLOCKED: one TMA box ([rows_per_box x 64 k], 128 B per
// row) is one contiguous run, boxes ordered [row_tile][k_block]:
//   element (row, k) at (((row / RB) * kblocks + k / 64) * RB + row % RB) * 64 + k % 64
// -> every TMA load is a contiguous 16/32 KB read and the epilogue fills whole boxes.
__host__ __device__ __forceinline__ int64_t box_index(int64_t row, int k, int kblocks, int RB) {
    return (((row / RB) * kblocks + (k / BK)) * RB + (row % RB)) * (int64_t)BK + (k % BK);
}

// asynchronous variant: the registers are valid only after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait + register dependency: nothing that reads v may be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),
                   "+r"(v[15])
                 :
                 : "memory");
}

struct StepArgs {
    int mode;          // TC_MODE_STEP | TC_MODE_GRAD
    int n_pass;        // 1 (bf16) or 3 (bf16x3 split)
    int64_t C;
    int D, Dp;
    int m_tiles;
    int64_t n_tiles;
    int kblocks;       // Dp / BK
    float eps;
    const float* metric;  // [D] or NULL
    const float* cvec;    // [D] P*mu or NULL
    float* q_prev;        // tile-blocked (STEP: q_{n-1} in, q_{n+1} out)
    float* q_cur;         // tile-blocked (STEP: q_n, read only within a step; the two swap roles every fused step)
    float* q_out0;        // NULL, or a third tile-blocked array that receives q_{n+1} of the launch's FIRST step, so that
                          // q_prev (the trajectory's start = the chain state) survives; later steps ping-pong between
                          // q_cur and q_out0
    __nv_bfloat16* q_hi_next;  // [C, Dp] (STEP): operand written by even fused steps (0, 2, ...)
    __nv_bfloat16* q_hi_cur;   // operand READ by step 0 (mapB0) = written by odd fused steps; multi-step launches only
    __nv_bfloat16* q_lo_next;  // [C, Dp] or NULL: low half of the LAST fused step's operand
    // Fused leapfrog steps (STEP + pair only): one persistent launch runs n_steps steps.  Step s of chain-tile j
    // needs the bf16 operand rows all m_tiles dim-tiles of (s-1, j) wrote: every epilogue warp adds 1 to
    // sync[(s-1) * n_tiles + j] (release, gpu scope) when its part of the tile is written; the TMA producers
    // spin (acquire) until it reaches m_tiles * EPI_WARPS.  Zeroed by the host before the launch.
    int n_steps;
    uint32_t* sync;
    uint32_t* err;        // set to 1 when a producer gave up waiting (fused grid not co-resident): the end kernel
                          // then reports NaN log densities instead of the launch hanging
    float* g_out;         // [C, D] row-major, or tile-blocked when g_blocked
    int g_blocked;
    int debug;            // BK_TC_DEBUG bits: 1 = skip TMA+MMA, 2 = skip epilogue global traffic, 4 = load B on even stages only,
                          // 16 = accumulator never read
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
k_dense_tc(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
           const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
           const StepArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms (dynamic smem starts at the same offset in both
    // CTAs of a pair, so the carve-up below is identical in both -- the pair MMA relies on that)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    using C_ = Cfg<MODE, PAIR>;
    constexpr int STAGES = C_::STAGES;
    constexpr int SMEM_TILES = C_::SMEM_TILES;
    constexpr int KS = C_::KS, A_BYTES = C_::A_BYTES, B_BYTES = C_::B_BYTES;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs)
    const int n_cta = PAIR ? 2 : 1;
    constexpr int SPK = BK / KS;                               // stages per 64-wide k-block
    const uint32_t sA = base, sB = base + STAGES * A_BYTES;
    const uint32_t ering = base + SMEM_TILES;                 // STEP: per-warp r/q rings
    const uint32_t bars = ering + C_::ERING;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tempty = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    auto twritten = [&](uint32_t k) { return bars + 8u * (2 * STAGES + 5 + (k % SIG_SLOTS)); };
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        // pair: the leader's tempty collects one arrival per epilogue WARP of both CTAs
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), PAIR ? 2 * EPI_WARPS : 32 * EPI_WARPS); }
        for (int s = 0; s < SIG_SLOTS; ++s) mbar_init(twritten(s), EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (PAIR) cluster_sync_all();          // both CTAs' barriers exist before anything remote touches them
    if (warp == 1) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // work units: single CTA -> tiles (n_tile, m_tile); pair -> (n_tile, pair of adjacent m_tiles), CTA `rank`
    // of the cluster owning m_tile = 2 * m_pair + rank.  `unit` below enumerates them; first / stride per CTA.
    const int m_units = PAIR ? a.m_tiles / 2 : a.m_tiles;
    const int64_t total_tiles = a.n_tiles * m_units;
    const int64_t unit0 = PAIR ? blockIdx.x / 2 : blockIdx.x, unit_stride = PAIR ? gridDim.x / 2 : gridDim.x;
    auto m_tile_of = [&](int64_t unit) { return PAIR ? 2 * (int)(unit % m_units) + (int)rank : (int)(unit % m_units); };
    const int iters = a.n_pass * a.kblocks * SPK;
    const int n_steps = (MODE == TC_MODE_STEP && PAIR && a.n_steps > 1) ? a.n_steps : 1;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0 && !(a.debug & 1)) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t sync_full = (uint32_t)a.m_tiles;      // one count per CTA that writes the chain-tile
            for (int step = 0; step < n_steps; ++step)
            for (int64_t tile = unit0; tile < total_tiles; tile += unit_stride) {
                const int m_tile = m_tile_of(tile);
                const int64_t n_tile = tile / m_units;
                if (step > 0) {      // every dim-tile of this chain-tile has written the previous step's operand
                    const uint32_t* flag = a.sync + (int64_t)(step - 1) * a.n_tiles + n_tile;
                    if (ld_acquire_gpu(flag) < sync_full) {
                        // bounded: the fused grid's co-residency rests on an occupancy query, not on a cooperative
                        // launch; if another long-lived kernel holds SMs the wait gives up after ~4 s
                        uint64_t t_start;
                        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
                        unsigned spins = 0;
                        while (ld_acquire_gpu(flag) < sync_full) {
                            __nanosleep(64);
                            if ((++spins & 4095u) == 0) {
                                uint64_t t_now;
                                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_now));
                                if (t_now - t_start > 4000000000ull || (a.err && *(volatile uint32_t*)a.err)) {
                                    if (a.err) atomicExch(a.err, 1u);
                                    break;
                                }
                            }
                        }
                    }
                    fence_proxy_async_all();     // generic-proxy writes (other CTAs) -> this thread's TMA reads
                }
                for (int it = 0; it < iters; ++it) {
                    const int pass = it / (a.kblocks * SPK), ks = it % (a.kblocks * SPK);
                    const int kb = ks / SPK, kx = (ks % SPK) * KS;   // 64-wide box, k offset inside it
                    // pass 0: (A_hi, B_hi)  pass 1: (A_hi, B_lo)  pass 2: (A_lo, B_hi);
                    // fused steps (n_pass = 1): the operand buffers alternate, mapB0 on even steps, mapB1 on odd
                    const CUtensorMap* ma = pass == 2 ? &mapA1 : &mapA0;
                    const CUtensorMap* mb = (pass == 1 || (step & 1)) ? &mapB1 : &mapB0;
                    mbar_wait(empty(stage), phase ^ 1u);
                    if constexpr (PAIR) {
                        // both CTAs load into their own shared memory; every byte is credited to the LEADER's
                        // full barrier (the leader's MMA thread is the only consumer of the stage)
                        const uint32_t fb = mapa_rank(full(stage), 0);
                        if (rank == 0) mbar_expect_tx(full(stage), 2 * (A_BYTES + B_BYTES));
                        tma_load_2d_pair(sA + stage * A_BYTES, ma, fb, kx, (m_tile * a.kblocks + kb) * BM);
                        tma_load_2d_pair(sB + stage * B_BYTES, mb, fb, kx,
                                         (int)((n_tile * a.kblocks + kb) * BN + rank * (BN / 2)));
                    } else {
                        const bool skip_b = (a.debug & 4) && (it & 1);   // timing experiment: half the B traffic
                        mbar_expect_tx(full(stage), A_BYTES + (skip_b ? 0 : B_BYTES));
                        tma_load_2d(sA + stage * A_BYTES, ma, full(stage), kx, (m_tile * a.kblocks + kb) * BM);
                        if (!skip_b)
                            tma_load_2d(sB + stage * B_BYTES, mb, full(stage), kx,
                                        (int)((n_tile * a.kblocks + kb) * BN));
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (PAIR && rank != 0) {
            // the peer CTA's MMA warp only took part in the TMEM allocation
        } else if (lane == 0 && (a.debug & 1)) {   // experiment: no GEMM, just hand the accumulators over
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int step = 0; step < n_steps; ++step)
            for (int64_t tile = unit0; tile < total_tiles; tile += unit_stride) {
                mbar_wait(tempty(acc), acc_phase ^ 1u);
                mbar_arrive(tfull(acc));
                if constexpr (PAIR) mbar_arrive_cluster(mapa_rank(tfull(acc), 1));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        } else if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int step = 0; step < n_steps; ++step)
            for (int64_t tile = unit0; tile < total_tiles; tile += unit_stride) {
                mbar_wait(tempty(acc), acc_phase ^ 1u);   // epilogue(s) drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(full(stage), phase);
                    tc_fence_after();
                    const uint32_t a0 = sA + stage * A_BYTES, b0 = sB + stage * B_BYTES;
#pragma unroll
                    for (int k = 0; k < KS / UK; ++k) {
                        if constexpr (PAIR)
                            umma_pair(d_tmem, umma_desc<KS * 2>(a0 + k * UK * 2), umma_desc<KS * 2>(b0 + k * UK * 2),
                                      IDESC_PAIR, (it | k) != 0 ? 1u : 0u);
                        else
                            umma(d_tmem, umma_desc<KS * 2>(a0 + k * UK * 2), umma_desc<KS * 2>(b0 + k * UK * 2),
                                 (it | k) != 0 ? 1u : 0u);
                    }
                    // smem slot reusable once these MMAs retire (pair: in both CTAs)
                    if constexpr (PAIR) umma_commit_pair(empty(stage)); else umma_commit(empty(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                // accumulator ready for the epilogue (pair: each CTA's own 128 dims x 256 chains)
                if constexpr (PAIR) umma_commit_pair(tfull(acc)); else umma_commit(tfull(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp == SIG_WARP) {
        // ================= signaller (fused STEP launches) =================
        // Epilogue warps arrive on a CTA-local barrier when their part of a tile is written; this thread then
        // releases the tile at gpu scope (cumulative over the writes it observed through the barrier).
        if constexpr (MODE == TC_MODE_STEP && PAIR) {
            if (lane == 0 && n_steps > 1) {
                uint32_t k = 0;
                for (int step = 0; step < n_steps - 1; ++step)
                    for (int64_t tile = unit0; tile < total_tiles; tile += unit_stride, ++k) {
                        mbar_wait(twritten(k), (k / SIG_SLOTS) & 1u);
                        red_release_gpu_add(a.sync + (int64_t)step * a.n_tiles + tile / m_units, 1u);
                    }
            }
        }
    } else {
        // ================= epilogue: 8 warps, lane = dim, registers = chains =================
        // warp -> TMEM lane quarter (warp % 4) and column half (128 chains).
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int NCH = BN / 32 / (EPI_WARPS / 4);     // 32-chain groups per warp
        int acc = 0;
        uint32_t acc_phase = 0;
        if constexpr (MODE == TC_MODE_GRAD) {
            for (int64_t tile = unit0; tile < total_tiles; tile += unit_stride) {
                const int m_tile = m_tile_of(tile);
                const int64_t n_tile = tile / m_units;
                const int d = m_tile * BM + quarter * 32 + lane;
                const bool d_ok = d < a.D;
                const float cv = (a.cvec && d_ok) ? a.cvec[d] : 0.0f;
                const uint32_t t0 = tmem_base + (uint32_t)acc * BN + ((uint32_t)(quarter * 32) << 16) +
                                    (uint32_t)(half * NCH * 32);
                const int64_t cbase = n_tile * BN + half * NCH * 32;
                mbar_wait(tfull(acc), acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch) {
                    uint32_t v[32];
                    tmem_ld32(t0 + ch * 32, v);
                    const int64_t c0 = cbase + ch * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int64_t c = c0 + j;
                        if (d_ok && c < a.C)
                            a.g_out[a.g_blocked ? blk_index(c, d, a.m_tiles) : c * a.D + d] =
                                cv - __uint_as_float(v[j]);
                    }
                }
                tc_fence_before();
                if constexpr (PAIR) {          // one arrival per warp on the LEADER's barrier
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_rank(tempty(acc), 0));
                } else {
                    mbar_arrive(tempty(acc));
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        } else {
            // STEP: fused leapfrog update.  Each warp streams ITS 16-chain x 32-dim pieces of
            // r and q through a private cp.async ring in shared memory (EDEPTH-1 chunks =
            // 12 KB per warp, 96 KB per SM in flight, no registers held, prefetch runs
            // across tile boundaries and ahead of the accumulator), then updates from
            // registers:  q_next = q + ((q - q_prev) + eps^2*m*g) ; q_hi(next operand) = bf16(q_next).
            constexpr int NC = NCH * 32 / CWID;                 // chunks per warp per tile
            const bool no_mem = (a.debug & 2) != 0;
            constexpr int EDEPTH = C_::EDEPTH;
            const uint32_t ring = ering + (uint32_t)(warp - 2) * (EDEPTH * ECHUNK_BYTES);
            const int row_in = lane >> 3, seg = lane & 7;       // cp.async: 4 rows x 8 x 16 B per instruction
            // A unit's 256 x 128 patch of the tile-blocked arrays starts at (n_tile * m_tiles + m_tile) * BN * BM
            // = (pair ? 2 * unit + rank : unit) * BN * BM: no divisions on the streaming side.  Inside the patch
            // this warp's chunks are CWID * BM elements apart.
            const int64_t warp_off = (int64_t)(half * NCH * 32) * BM + quarter * 32;
            auto patch0 = [&](int64_t unit) -> int64_t { return (PAIR ? 2 * unit + rank : unit) * (int64_t)(BN * BM) + warp_off; };
            // prefetch cursor: next chunk to request (tile, chunk in tile, ring slot, source pointers of this lane)
            int64_t pf_tile = unit0;
            int pf_ch = 0, pf_step = 0;
            const float* pf_prev = a.q_prev;     // roles swap every fused step
            const float* pf_cur = a.q_cur;
            float* const q_alt = a.q_out0 ? a.q_out0 : a.q_prev;   // where the first step's q_{n+1} goes
            uint32_t pf_dst = ring + (uint32_t)(row_in * 128 + seg * 16);
            int pf_slot = 0;
            int64_t pf_e = patch0(pf_tile) + (int64_t)row_in * BM + seg * 4;
            if (unit0 >= total_tiles) pf_step = n_steps;      // no work for this CTA
            auto issue = [&]() {
                if (pf_step < n_steps && !no_mem) {
                    const float* sp = pf_prev + pf_e;
                    const float* sc = pf_cur + pf_e;
#pragma unroll
                    for (int it = 0; it < CWID / 4; ++it) {
                        cp_async16(pf_dst + it * 512, sp + it * 4 * BM);
                        cp_async16(pf_dst + CWID * 128 + it * 512, sc + it * 4 * BM);
                    }
                    pf_e += CWID * BM;
                    if (++pf_ch == NC) {
                        pf_ch = 0;
                        pf_tile += unit_stride;
                        if (pf_tile >= total_tiles) {     // next fused step: same units, q_{n-1} / q_n swapped
                            pf_tile = unit0;
                            ++pf_step;
                            if (pf_step == 1) { pf_prev = a.q_cur; pf_cur = q_alt; }
                            else { const float* t = pf_prev; pf_prev = pf_cur; pf_cur = t; }
                        }
                        pf_e = patch0(pf_tile) + (int64_t)row_in * BM + seg * 4;
                    }
                }
                cp_async_commit();
                pf_dst += ECHUNK_BYTES;
                if (++pf_slot == EDEPTH) { pf_slot = 0; pf_dst -= EDEPTH * ECHUNK_BYTES; }
            };
#pragma unroll
            for (int p = 0; p < EDEPTH - 1; ++p) issue();
            uint32_t src = ring + (uint32_t)lane * 4;           // consume cursor (this lane's dim column of the chunk)
            int slot = 0;
            uint32_t sig_k = 0;                                 // tiles handed to the signaller so far
            for (int step = 0; step < n_steps; ++step)
            for (int64_t tile = unit0; tile < total_tiles; tile += unit_stride) {
                float* const q_out = (step & 1) ? a.q_cur : q_alt;               // q_{n+1} replaces q_{n-1} (first step: q_out0)
                __nv_bfloat16* const hi_out = (step & 1) ? a.q_hi_cur : a.q_hi_next;
                __nv_bfloat16* const lo_out = step == n_steps - 1 ? a.q_lo_next : nullptr;
                const int m_tile = m_tile_of(tile);
                const int64_t n_tile = tile / m_units;
                const int d = m_tile * BM + quarter * 32 + lane;
                const bool d_ok = d < a.D;
                const float cv = (a.cvec && d_ok) ? a.cvec[d] : 0.0f;
                const float em2 = a.eps * a.eps * ((a.metric && d_ok) ? a.metric[d] : 1.0f);
                const uint32_t t0 = tmem_base + (uint32_t)acc * BN + ((uint32_t)(quarter * 32) << 16) +
                                    (uint32_t)(half * NCH * 32);
                const int64_t cbase = n_tile * BN + half * NCH * 32;
                // warp-uniform: every lane's dim is real and all of this warp's chains exist -> no predicates
                const bool fast = __all_sync(0xffffffffu, d_ok && cbase + NCH * 32 <= a.C) && !no_mem;
                float* qw = q_out + patch0(tile) + lane;
                const int64_t b0i = box_index(cbase, d, a.kblocks, BN);
                __nv_bfloat16* hw = hi_out + b0i;
                __nv_bfloat16* lw = lo_out ? lo_out + b0i : nullptr;
                mbar_wait(tfull(acc), acc_phase);
                tc_fence_after();
                // one chunk: 16 chains x this lane's dim.  eps*r_{n+1/2} = eps*r_{n-1/2} + eps^2*m*g ;
                // q_{n+1} = q_n + eps*r_{n+1/2}
                auto process = [&](const uint32_t (&v)[CWID], int ch) {
                    if (fast) {
                        if (lw == nullptr) {
#pragma unroll
                            for (int j = 0; j < CWID; ++j) {
                                float pj, qj;
                                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pj) : "r"(src + j * 128));
                                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qj) : "r"(src + CWID * 128 + j * 128));
                                const float gj = cv - __uint_as_float(v[j]);
                                const float qn = qj + fmaf(em2, gj, qj - pj);
                                __stcs(qw + j * BM, qn);
                                hw[j * BK] = __float2bfloat16_rn(qn);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CWID; ++j) {
                                float pj, qj;
                                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pj) : "r"(src + j * 128));
                                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qj) : "r"(src + CWID * 128 + j * 128));
                                const float gj = cv - __uint_as_float(v[j]);
                                const float qn = qj + fmaf(em2, gj, qj - pj);
                                __stcs(qw + j * BM, qn);
                                const __nv_bfloat16 hi = __float2bfloat16_rn(qn);
                                hw[j * BK] = hi;
                                lw[j * BK] = __float2bfloat16_rn(qn - __bfloat162float(hi));
                            }
                        }
                    } else if (!no_mem) {
#pragma unroll
                        for (int j = 0; j < CWID; ++j) {
                            float pj, qj;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pj) : "r"(src + j * 128));
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qj) : "r"(src + CWID * 128 + j * 128));
                            if (d_ok && cbase + ch * CWID + j < a.C) {
                                const float gj = cv - __uint_as_float(v[j]);
                                const float qn = qj + fmaf(em2, gj, qj - pj);
                                __stcs(qw + j * BM, qn);
                                const __nv_bfloat16 hi = __float2bfloat16_rn(qn);
                                hw[j * BK] = hi;
                                if (lw) lw[j * BK] = __float2bfloat16_rn(qn - __bfloat162float(hi));
                            }
                        }
                    }
                    qw += CWID * BM;
                    hw += CWID * BK;
                    if (lw) lw += CWID * BK;
                    src += ECHUNK_BYTES;
                    if (++slot == EDEPTH) { slot = 0; src -= EDEPTH * ECHUNK_BYTES; }
                    __syncwarp();                     // stage may be refilled by the next issue
                };
                // the accumulator chunk of the NEXT iteration is requested before the current one is processed
                static_assert(NC % 2 == 0, "two register sets ping-pong");
                uint32_t va[CWID], vb[CWID];
                const bool no_tmem = (a.debug & 16) != 0;     // debug 16: the accumulator is never read
                if (!no_tmem) tmem_ld16_async(t0, va);
#pragma unroll 1
                for (int ch = 0; ch < NC; ch += 2) {
                    issue();
                    cp_async_wait<EDEPTH - 1>();      // this chunk has landed
                    __syncwarp();
                    tmem_wait_ld16(va);
                    if (!no_tmem) tmem_ld16_async(t0 + (ch + 1) * CWID, vb);
                    process(va, ch);
                    issue();
                    cp_async_wait<EDEPTH - 1>();
                    __syncwarp();
                    tmem_wait_ld16(vb);
                    if (ch + 2 < NC && !no_tmem) tmem_ld16_async(t0 + (ch + 2) * CWID, va);
                    process(vb, ch + 1);
                }
                if (!d_ok && !no_mem) {   // pad dim (last dim-tile only): keep the next operand's padding zero
                    hw -= NC * CWID * BK;
                    if (lw) lw -= NC * CWID * BK;
                    for (int j = 0; j < NCH * 32; ++j)
                        if (cbase + j < a.C) {
                            hw[(int64_t)j * BK] = __float2bfloat16_rn(0.f);
                            if (lw) lw[(int64_t)j * BK] = __float2bfloat16_rn(0.f);
                        }
                }
                tc_fence_before();
                if constexpr (PAIR) {          // one arrival per warp on the LEADER's barrier
                    if (step < n_steps - 1) fence_proxy_async_global();   // this lane's operand stores -> (other CTAs') TMA reads
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive_cluster(mapa_rank(tempty(acc), 0));
                        if (step < n_steps - 1) mbar_arrive(twritten(sig_k));
                    }
                    ++sig_k;
                } else {
                    mbar_arrive(tempty(acc));
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            cp_async_wait<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();      // both CTAs are done with the pair's TMEM and barriers
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                         : "memory");
    }
}

// ---- operand preparation ---------------------------------------------------------
// P [D,D] fp32 -> zero-padded bf16 hi/lo [Dp,Dp]
__global__ void k_split_matrix(const float* __restrict__ P, int D, int Dp, __nv_bfloat16* __restrict__ hi,
                               __nv_bfloat16* __restrict__ lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Dp * Dp) return;
    int r = (int)(i / Dp), c = (int)(i % Dp);
    float v = (r < D && c < D) ? P[(int64_t)r * D + c] : 0.0f;
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    const int64_t o = box_index(r, c, Dp / BK, BM);
    hi[o] = h;
    lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// cvec = P * mu (fp64 accumulate), one warp per row
__global__ void k_pmu(const float* __restrict__ P, const float* __restrict__ mu, int D, float* __restrict__ out) {
    int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (r >= D) return;
    double s = 0;
    for (int k = lane; k < D; k += 32) s += (double)P[(int64_t)r * D + k] * (double)mu[k];
    s = warp_sum(s);
    if (lane == 0) out[r] = (float)s;
}

// theta [C,D] fp32 -> bf16 hi (/lo) [C,Dp]; pad columns written as zeros
__global__ void k_split_rows(const float* __restrict__ x, int64_t C, int D, int Dp,
                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * Dp) return;
    int64_t c = i / Dp;
    int d = (int)(i % Dp);
    float v = d < D ? x[c * D + d] : 0.f;
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    const int64_t o = box_index(c, d, Dp / BK, BN);
    hi[o] = h;
    if (lo) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ---- HMC begin / end around the tensor-core steps (fp32) ----------------------------
struct HmcTcArgs {
    float *theta, *lp, *grad;   // state
    float *q, *qm, *gq, *h0;    // workspace: begin writes q_1 -> q, q_0 -> qm; end reads q_L, q_{L-1}
    __nv_bfloat16 *q_hi, *q_lo; // operand for the first GEMM
    const float *metric, *mu;
    int64_t C;
    int D, Dp, L, m_tiles;
    float eps, half_eps, inv_eps;
    bk_rng rng;
    float *draws, *logp;
    int32_t* accept;
    int out_lp;                 // 1: report log p(theta) (MALA, mala.py:66) instead of the joint (hmc.py:63)
    const uint32_t* err;        // fused STEP launch gave up waiting (see StepArgs::err): report NaN
    // draws 2.. of a sample_n call: the chain state lives in TILE-BLOCKED arrays (qs = position, gs = its gradient)
    // left there by k_hmc_turn_tc; theta / grad (row-major) are only written by the call's last draw
    const float *qs, *gs;
    int old_blocked;            // 1: the state before this draw is (qs, gs); 0: (theta, grad)
    __nv_bfloat16 *nq_hi, *nq_lo;   // turn: operand of the NEXT trajectory's first GEMM
};

// One warp per chain, lane-strided blocks of 4 elements (the Philox block
// mapping); VEC = rows are 16-byte aligned and D % 4 == 0 -> 128-bit accesses.
template <bool VEC>
__device__ __forceinline__ void ld4(const float* p, int e, int D, float (&v)[4]) {
    if (VEC) {
        const float4 t = *reinterpret_cast<const float4*>(p + e);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (e + i < D) ? p[e + i] : 0.f;
    }
}
template <bool VEC>
__device__ __forceinline__ void st4(float* p, int e, int D, const float (&v)[4]) {
    if (VEC) {
        *reinterpret_cast<float4*>(p + e) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (e + i < D) p[e + i] = v[i];
    }
}
template <bool VEC>
__device__ __forceinline__ void st4_bf16(__nv_bfloat16* p, int e, int D, const __nv_bfloat16 (&v)[4]) {
    if (VEC) {
        uint2 u;
        u.x = (uint32_t)__bfloat16_as_ushort(v[0]) | ((uint32_t)__bfloat16_as_ushort(v[1]) << 16);
        u.y = (uint32_t)__bfloat16_as_ushort(v[2]) | ((uint32_t)__bfloat16_as_ushort(v[3]) << 16);
        *reinterpret_cast<uint2*>(p + e) = u;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (e + i < D) p[e + i] = v[i];
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_hmc_begin_tc(HmcTcArgs p, int64_t t, int write_lo) {
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= p.C) return;
    const int D = p.D;
    const int64_t off = c * (int64_t)D;
    float kin = 0.f;
#pragma unroll 2
    for (int b = lane; 4 * b < D; b += 32) {
        const int e = 4 * b;
        float z[4], g[4], th[4], m[4] = {1.f, 1.f, 1.f, 1.f};
        ld4<VEC>(p.grad + off, e, D, g);
        ld4<VEC>(p.theta + off, e, D, th);
        if (p.metric) ld4<VEC>(p.metric, e, D, m);
        if (p.rng.mode == BK_RNG_INJECTED)
            ld4<VEC>(reinterpret_cast<const float*>(p.rng.normals) + (t * p.C + c) * (int64_t)D, e, D, z);
        else
            philox_normal4<float>(p.rng.seed, (uint32_t)b, (uint32_t)(p.rng.chain_offset + (uint64_t)c),
                                  (uint32_t)(p.rng.draw_offset + (uint64_t)t), z);
        float r[4], q[4];
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!VEC && e + i >= D) z[i] = 0.f;
            const float mg = m[i] * g[i];
            kin = fmaf(z[i], m[i] * z[i], kin);
            r[i] = z[i] - p.half_eps * mg;    // backward half kick (hmc.py:46)
            r[i] = r[i] + p.eps * mg;         // first full kick     (hmc.py:48)
            q[i] = th[i] + p.eps * r[i];
            hi[i] = __float2bfloat16_rn(q[i]);
            lo[i] = __float2bfloat16_rn(q[i] - __bfloat162float(hi[i]));
        }
        const int64_t bo = blk_index(c, e, p.m_tiles);   // q_0, q_1 are tile-blocked
        st4<true>(p.qm + bo, 0, 4, th);
        st4<true>(p.q + bo, 0, 4, q);
        const int64_t xo = box_index(c, e, p.Dp / BK, BN);   // e % 4 == 0: stays inside one box row
        st4_bf16<true>(p.q_hi + xo, 0, 4, hi);
        if (write_lo) st4_bf16<true>(p.q_lo + xo, 0, 4, lo);
    }
    // pad dims [D, Dp) of this chain's MMA operand rows must be zero (they meet P's zero padding)
    for (int b = (D + 3) / 4 + lane; 4 * b < p.Dp; b += 32) {
        const __nv_bfloat16 zero4[4] = {__float2bfloat16_rn(0.f), __float2bfloat16_rn(0.f), __float2bfloat16_rn(0.f),
                                        __float2bfloat16_rn(0.f)};
        const int64_t xo = box_index(c, 4 * b, p.Dp / BK, BN);
        st4_bf16<true>(p.q_hi + xo, 0, 4, zero4);
        if (write_lo) st4_bf16<true>(p.q_lo + xo, 0, 4, zero4);
    }
    kin = warp_sum(kin);
    if (lane == 0) p.h0[c] = p.lp[c] - 0.5f * kin;
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_hmc_end_tc(HmcTcArgs p, int64_t t) {
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= p.C) return;
    const int D = p.D;
    const int64_t off = c * (int64_t)D;
    float kin = 0.f, dot = 0.f;
#pragma unroll 2
    for (int b = lane; 4 * b < D; b += 32) {
        const int e = 4 * b;
        float g[4], qm[4], q[4], m[4] = {1.f, 1.f, 1.f, 1.f}, mu[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t bo = blk_index(c, e, p.m_tiles);   // gq, q_{L-1}, q_L are tile-blocked
        ld4<true>(p.gq + bo, 0, 4, g);
        ld4<true>(p.qm + bo, 0, 4, qm);
        ld4<true>(p.q + bo, 0, 4, q);
        if (p.metric) ld4<VEC>(p.metric, e, D, m);
        if (p.mu) ld4<VEC>(p.mu, e, D, mu);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (e + i >= D) break;                                // blocked arrays carry pad dims
            // r_{L-1/2} = (q_L - q_{L-1}) / eps, then the forward half kick (hmc.py:52)
            const float rf = fmaf(q[i] - qm[i], p.inv_eps, p.half_eps * (m[i] * g[i]));
            kin = fmaf(rf, m[i] * rf, kin);
            dot = fmaf(q[i] - mu[i], g[i], dot);                  // log p(q) = 0.5 (q-mu).g
        }
    }
    kin = warp_sum(kin);
    dot = warp_sum(dot);
    const bool lost = p.err && *p.err != 0u;
    const float lpq = lost ? __int_as_float(0x7fc00000) : 0.5f * dot;
    const float h1 = lpq - 0.5f * kin, h0 = lost ? lpq : p.h0[c];
    float u;
    if (p.rng.mode == BK_RNG_INJECTED)
        u = reinterpret_cast<const float*>(p.rng.uniforms)[(t * p.C + c) * p.rng.n_uniform];
    else
        u = philox_accept_uniform<float>(p.rng.seed, (uint32_t)(p.rng.chain_offset + (uint64_t)c),
                                         (uint32_t)(p.rng.draw_offset + (uint64_t)t), D);
    const bool acc = log_u(u) < h1 - h0;
    float* dr = p.draws ? p.draws + (t * p.C + c) * (int64_t)D : nullptr;
#pragma unroll 2
    for (int b = lane; 4 * b < D; b += 32) {
        const int e = 4 * b;
        float v[4];
        if (acc) {
            float g[4];
            const int64_t bo = blk_index(c, e, p.m_tiles);
            ld4<true>(p.q + bo, 0, 4, v);
            ld4<true>(p.gq + bo, 0, 4, g);
            st4<VEC>(p.theta + off, e, D, v);
            st4<VEC>(p.grad + off, e, D, g);
        } else if (p.old_blocked) {
            // the state moved into the blocked arrays after the call's first draw: theta / grad are stale
            float g[4];
            const int64_t bo = blk_index(c, e, p.m_tiles);
            ld4<true>(p.qs + bo, 0, 4, v);
            ld4<true>(p.gs + bo, 0, 4, g);
            st4<VEC>(p.theta + off, e, D, v);
            st4<VEC>(p.grad + off, e, D, g);
        } else if (dr) {
            ld4<VEC>(p.theta + off, e, D, v);
        }
        if (dr) st4<VEC>(dr, e, D, v);
    }
    if (lane == 0) {
        const float lp_old = p.lp[c];
        if (acc) p.lp[c] = lpq;
        if (p.logp) p.logp[t * p.C + c] = p.out_lp ? (acc ? lpq : lp_old) : (acc ? h1 : h0);
        if (p.accept) p.accept[t * p.C + c] = acc ? 1 : 0;
    }
}

// End of draw t AND begin of draw t + 1 in one pass (draws 1 .. n-1 of a sample_n call): the Metropolis decision of
// k_hmc_end_tc, then -- instead of copying the accepted state out to theta / grad and reading it back -- the next
// trajectory starts from where the state already is.  Accepted chains keep q_L / g(q_L) in place (the arrays p.q /
// p.gq become the state arrays of the next draw); rejected chains get their old state copied into those rows (rare).
// Then the momentum of draw t + 1, H_0, the first kick and drift: q_1 -> p.qm (q_{L-1} is dead by then; same warp,
// same rows) and the bf16 operand.  DRAM traffic: 12 B (q_L, q_{L-1}, g) + 10 B (draw, q_1, operand) per element
// against 24 + 18 for end + begin; 0.56-0.60 ms against 0.21 + 0.53 (c2: 2.97 -> 2.78 ms per draw).  Latency-bound like
// the end kernel (ncu profiles/r2_ncu_hmc_turn.md: long-scoreboard stalls at the first use of each phase's loads, IPC
// 1.1): a variant that staged a chain's three rows through shared memory with cp.async (12 KB per warp, next chain
// requested under the second phase or after it) and 3 instead of 4 CTAs per SM all measured within 0.3 % of this one.
template <bool VEC>
__global__ void __launch_bounds__(256, 4) k_hmc_turn_tc(HmcTcArgs p, int64_t t, int write_lo) {
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= p.C) return;
    const int D = p.D;
    const int64_t off = c * (int64_t)D;
    float kin = 0.f, dot = 0.f;
#pragma unroll 2
    for (int b = lane; 4 * b < D; b += 32) {
        const int e = 4 * b;
        float g[4], qm[4], q[4], m[4] = {1.f, 1.f, 1.f, 1.f}, mu[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t bo = blk_index(c, e, p.m_tiles);
        ld4<true>(p.gq + bo, 0, 4, g);
        ld4<true>(p.qm + bo, 0, 4, qm);
        ld4<true>(p.q + bo, 0, 4, q);
        if (p.metric) ld4<VEC>(p.metric, e, D, m);
        if (p.mu) ld4<VEC>(p.mu, e, D, mu);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (e + i >= D) break;
            const float rf = fmaf(q[i] - qm[i], p.inv_eps, p.half_eps * (m[i] * g[i]));
            kin = fmaf(rf, m[i] * rf, kin);
            dot = fmaf(q[i] - mu[i], g[i], dot);
        }
    }
    kin = warp_sum(kin);
    dot = warp_sum(dot);
    const bool lost = p.err && *p.err != 0u;
    const float lpq = lost ? __int_as_float(0x7fc00000) : 0.5f * dot;
    const float h1 = lpq - 0.5f * kin, h0 = lost ? lpq : p.h0[c];
    float u;
    if (p.rng.mode == BK_RNG_INJECTED)
        u = reinterpret_cast<const float*>(p.rng.uniforms)[(t * p.C + c) * p.rng.n_uniform];
    else
        u = philox_accept_uniform<float>(p.rng.seed, (uint32_t)(p.rng.chain_offset + (uint64_t)c),
                                         (uint32_t)(p.rng.draw_offset + (uint64_t)t), D);
    const bool acc = log_u(u) < h1 - h0;
    float* dr = p.draws ? p.draws + (t * p.C + c) * (int64_t)D : nullptr;
    float kin0 = 0.f;
#pragma unroll 2
    for (int b = lane; 4 * b < D; b += 32) {
        const int e = 4 * b;
        const int64_t bo = blk_index(c, e, p.m_tiles);
        float v[4], g[4], z[4], m[4] = {1.f, 1.f, 1.f, 1.f};
        if (acc) {
            ld4<true>(p.q + bo, 0, 4, v);
            ld4<true>(p.gq + bo, 0, 4, g);
        } else {
            if (p.old_blocked) {
                ld4<true>(p.qs + bo, 0, 4, v);
                ld4<true>(p.gs + bo, 0, 4, g);
            } else {
                ld4<VEC>(p.theta + off, e, D, v);
                ld4<VEC>(p.grad + off, e, D, g);
            }
            st4<true>(p.q + bo, 0, 4, v);      // the state arrays of the next draw hold every chain
            st4<true>(p.gq + bo, 0, 4, g);
        }
        if (dr) st4<VEC>(dr, e, D, v);
        // ---- begin of draw t + 1 (k_hmc_begin_tc) from the state in registers ----
        if (p.metric) ld4<VEC>(p.metric, e, D, m);
        if (p.rng.mode == BK_RNG_INJECTED)
            ld4<VEC>(reinterpret_cast<const float*>(p.rng.normals) + ((t + 1) * p.C + c) * (int64_t)D, e, D, z);
        else
            philox_normal4<float>(p.rng.seed, (uint32_t)b, (uint32_t)(p.rng.chain_offset + (uint64_t)c),
                                  (uint32_t)(p.rng.draw_offset + (uint64_t)(t + 1)), z);
        float q1[4];
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (e + i >= D) { z[i] = 0.f; v[i] = 0.f; g[i] = 0.f; }
            const float mg = m[i] * g[i];
            kin0 = fmaf(z[i], m[i] * z[i], kin0);
            float r = z[i] - p.half_eps * mg;    // backward half kick (hmc.py:46)
            r = r + p.eps * mg;                  // first full kick     (hmc.py:48)
            q1[i] = v[i] + p.eps * r;
            hi[i] = __float2bfloat16_rn(q1[i]);
            lo[i] = __float2bfloat16_rn(q1[i] - __bfloat162float(hi[i]));
        }
        st4<true>(p.qm + bo, 0, 4, q1);
        const int64_t xo = box_index(c, e, p.Dp / BK, BN);
        st4_bf16<true>(p.nq_hi + xo, 0, 4, hi);
        if (write_lo) st4_bf16<true>(p.nq_lo + xo, 0, 4, lo);
    }
    for (int b = (D + 3) / 4 + lane; 4 * b < p.Dp; b += 32) {   // pad dims of the operand rows stay zero
        const __nv_bfloat16 zero4[4] = {__float2bfloat16_rn(0.f), __float2bfloat16_rn(0.f), __float2bfloat16_rn(0.f),
                                        __float2bfloat16_rn(0.f)};
        const int64_t xo = box_index(c, 4 * b, p.Dp / BK, BN);
        st4_bf16<true>(p.nq_hi + xo, 0, 4, zero4);
        if (write_lo) st4_bf16<true>(p.nq_lo + xo, 0, 4, zero4);
    }
    kin0 = warp_sum(kin0);
    if (lane == 0) {
        const float lp_old = p.lp[c];
        const float lp_new = acc ? lpq : lp_old;
        if (acc) p.lp[c] = lpq;
        p.h0[c] = lp_new - 0.5f * kin0;
        if (p.logp) p.logp[t * p.C + c] = p.out_lp ? lp_new : (acc ? h1 : h0);
        if (p.accept) p.accept[t * p.C + c] = acc ? 1 : 0;
    }
}

// ---- host side --------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    }
    return fn;
}

// bf16 row-major [rows, Dp] (pitch Dp), box = [box_rows x 64 k], 128B swizzle
// box-blocked bf16 operand: a [n_boxes * box_rows, 64] matrix with 128-byte rows
// `blk_rows` = rows per stored box (the array's blocking); a load may fetch `box_rows` <= blk_rows of one
static int make_map(CUtensorMap* m, const void* ptr, int64_t rows, int Dp, int box_rows, int box_k, int blk_rows = 0) {
    EncodeFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return BK_E_CUDA; }
    if (!blk_rows) blk_rows = box_rows;
    const int64_t row_tiles = (rows + blk_rows - 1) / blk_rows;
    cuuint64_t dims[2] = {(cuuint64_t)BK, (cuuint64_t)(row_tiles * (Dp / BK) * blk_rows)};
    cuuint64_t strides[1] = {(cuuint64_t)BK * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t el[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, el,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return BK_E_CUDA; }
    return BK_OK;
}

constexpr int MAX_FUSED_STEPS = 64;     // sync counters reserved in the workspace: MAX_FUSED_STEPS x n_tiles

// How many leapfrog steps one persistent STEP launch may fuse: needs the pair kernel with EVERY cluster of the
// grid co-resident (the producers spin on counters other clusters advance).  1 = launch per step.
// BK_TC_FUSE=0 (diagnostic) forces 1.
static int max_fused_steps(int m_tiles, int debug) {
    static int fuse_env = -1, pair_env = -1;
    static int clusters_dev[64], sms_dev[64];
    static bool dev_init = false;
    static std::mutex dev_mu;
    std::lock_guard<std::mutex> dev_lock(dev_mu);
    if (!dev_init) { for (int i = 0; i < 64; ++i) { clusters_dev[i] = -1; sms_dev[i] = 0; } dev_init = true; }
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= 64) return 1;
    int& clusters = clusters_dev[cur_dev];      // occupancy and SM count are per DEVICE, not per process
    int& sms = sms_dev[cur_dev];
    if (fuse_env < 0) { const char* e = getenv("BK_TC_FUSE"); fuse_env = (e && e[0] == '0') ? 0 : 1; }
    if (pair_env < 0) { const char* e = getenv("BK_TC_PAIR"); pair_env = (e && e[0] == '0') ? 0 : 1; }
    if (!fuse_env || !pair_env || m_tiles % 2 != 0 || (debug & 4)) return 1;
    if (clusters < 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
        cudaFuncSetAttribute(k_dense_tc<TC_MODE_STEP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg<TC_MODE_STEP, true>::SMEM_BYTES);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2u * (unsigned)(sms / 2));
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = Cfg<TC_MODE_STEP, true>::SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, k_dense_tc<TC_MODE_STEP, true>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
        clusters = n;
    }
    return clusters >= sms / 2 ? MAX_FUSED_STEPS : 1;
}

static int launch_tc(const Model& m, const StepArgs& a, const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo,
                     cudaStream_t st) {
    static bool attr_dev[64] = {false};
    static int sms_by_dev[64] = {0};
    int launch_dev = 0;
    BK_CUDA(cudaGetDevice(&launch_dev));
    if (launch_dev < 0 || launch_dev >= 64) { set_error("device index %d out of range", launch_dev); return BK_E_CUDA; }
    bool& attr = attr_dev[launch_dev];          // function attributes are per device
    if (!attr) {
        BK_CUDA(cudaFuncSetAttribute(k_dense_tc<TC_MODE_STEP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg<TC_MODE_STEP, false>::SMEM_BYTES));
        BK_CUDA(cudaFuncSetAttribute(k_dense_tc<TC_MODE_GRAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg<TC_MODE_GRAD, false>::SMEM_BYTES));
        BK_CUDA(cudaFuncSetAttribute(k_dense_tc<TC_MODE_STEP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg<TC_MODE_STEP, true>::SMEM_BYTES));
        BK_CUDA(cudaFuncSetAttribute(k_dense_tc<TC_MODE_GRAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg<TC_MODE_GRAD, true>::SMEM_BYTES));
        attr = true;
    }
    // CTA pairs (cta_group::2): two adjacent dim-tiles share one 256 x 256 MMA and each CTA stages only
    // half of the chain tile -- needs an even number of dim-tiles; BK_TC_PAIR=0 selects the single-CTA kernel
    static int pair_env = -1;
    if (pair_env < 0) { const char* e = getenv("BK_TC_PAIR"); pair_env = (e && e[0] == '0') ? 0 : 1; }
    const bool pair = pair_env && a.m_tiles % 2 == 0 && !(a.debug & 4);
    int& sms = sms_by_dev[launch_dev];
    if (!sms) BK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, launch_dev));
    CUtensorMap mA0, mA1, mB0, mB1;
    int rc;
    const int bk = a.mode == TC_MODE_STEP ? Cfg<TC_MODE_STEP>::KS : Cfg<TC_MODE_GRAD>::KS;
    if ((rc = make_map(&mA0, m.P_hi, m.Dp, (int)m.Dp, BM, bk))) return rc;
    if ((rc = make_map(&mA1, m.P_lo, m.Dp, (int)m.Dp, BM, bk))) return rc;
    const int brows = pair ? BN / 2 : BN;      // pair: each CTA loads half of a stored 256-row box
    if ((rc = make_map(&mB0, b_hi, a.C, (int)m.Dp, brows, bk, BN))) return rc;
    // fused steps: the operand buffers alternate -- mapB0 = read by even steps, mapB1 = read by odd steps
    const __nv_bfloat16* b1 = a.n_steps > 1 ? a.q_hi_next : (b_lo ? b_lo : b_hi);
    if ((rc = make_map(&mB1, b1, a.C, (int)m.Dp, brows, bk, BN))) return rc;
    if (a.n_steps > 1 && !(pair && a.mode == TC_MODE_STEP)) { set_error("fused steps need the pair STEP kernel"); return BK_E_INVALID; }
    const int tag = a.mode == TC_MODE_STEP ? BK_PROF_STEP : BK_PROF_GRAD;
    if (pair) {
        const int64_t units = a.n_tiles * (a.m_tiles / 2);
        const int64_t max_pairs = sms / 2;
        const unsigned grid = 2u * (unsigned)(units < max_pairs ? units : max_pairs);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(THREADS);
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        // (A cooperative launch would make the driver guarantee the fused grid's co-residency next to kernels of
        // other streams too, but Nsight Compute cannot replay a cooperative cluster launch -- "LaunchFailed" --
        // so co-residency rests on the occupancy query in max_fused_steps().)  Two fused grids that were
        // only partly resident next to each other would wait on each other forever, so fused launches of
        // this process never overlap: each one first waits (on its own stream, asynchronously) for the event
        // recorded behind the previous fused launch of the device, whatever stream that ran on.  Other
        // kernels may overlap freely -- they finish on their own and free their SMs.
        static std::mutex fused_mu;
        static cudaEvent_t fused_ev[64] = {nullptr};
        std::unique_lock<std::mutex> fused_lock(fused_mu, std::defer_lock);
        int fused_dev = -1;
        if (a.n_steps > 1) {
            fused_lock.lock();
            BK_CUDA(cudaGetDevice(&fused_dev));
            if (fused_dev < 0 || fused_dev >= 64) { set_error("device index %d out of range", fused_dev); return BK_E_CUDA; }
            if (!fused_ev[fused_dev]) BK_CUDA(cudaEventCreateWithFlags(&fused_ev[fused_dev], cudaEventDisableTiming));
            else BK_CUDA(cudaStreamWaitEvent(st, fused_ev[fused_dev], 0));
            BK_CUDA(cudaMemsetAsync(a.sync, 0, (size_t)(a.n_steps - 1) * a.n_tiles * sizeof(uint32_t), st));
        }
        prof_begin(tag, st);
        cudaError_t e;
        if (a.mode == TC_MODE_STEP) {
            cfg.dynamicSmemBytes = Cfg<TC_MODE_STEP, true>::SMEM_BYTES;
            e = cudaLaunchKernelEx(&cfg, k_dense_tc<TC_MODE_STEP, true>, mA0, mA1, mB0, mB1, a);
            if (fused_dev >= 0 && e == cudaSuccess) e = cudaEventRecord(fused_ev[fused_dev], st);

        } else {
            cfg.dynamicSmemBytes = Cfg<TC_MODE_GRAD, true>::SMEM_BYTES;
            e = cudaLaunchKernelEx(&cfg, k_dense_tc<TC_MODE_GRAD, true>, mA0, mA1, mB0, mB1, a);
        }
        prof_end(tag, st);
        if (e != cudaSuccess) { set_error("pair launch failed: %s", cudaGetErrorString(e)); return BK_E_CUDA; }
        count_launch();
        return BK_OK;
    }
    const int64_t tiles = a.n_tiles * a.m_tiles;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    prof_begin(tag, st);
    if (a.mode == TC_MODE_STEP)
        k_dense_tc<TC_MODE_STEP, false><<<grid, THREADS, Cfg<TC_MODE_STEP, false>::SMEM_BYTES, st>>>(mA0, mA1, mB0, mB1, a);
    else
        k_dense_tc<TC_MODE_GRAD, false><<<grid, THREADS, Cfg<TC_MODE_GRAD, false>::SMEM_BYTES, st>>>(mA0, mA1, mB0, mB1, a);
    prof_end(tag, st);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------
bool dense_tc_enabled(const Model& m) {
    const char* e = getenv("BK_DISABLE_TC");
    if (e && e[0] == '1') return false;
    return m.d.kind == BK_MODEL_DENSE_PREC_GAUSS && m.d.dtype == BK_F32 && m.P_hi != nullptr;
}

size_t dense_tc_model_ws_bytes(const bk_model_desc& d) {
    if (d.kind != BK_MODEL_DENSE_PREC_GAUSS || d.dtype != BK_F32) return 0;
    const size_t Dp = align_up((size_t)d.dims, 128);
    return 2 * align_up(Dp * Dp * 2, 256) + align_up((size_t)d.dims * 4, 256) + 1024;
}

int dense_tc_prepare(Model& m, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (m.d.kind != BK_MODEL_DENSE_PREC_GAUSS || m.d.dtype != BK_F32) return BK_OK;
    const int D = (int)m.d.dims;
    m.Dp = (int64_t)align_up((size_t)D, 128);
    Arena ar(ws, ws_bytes);
    m.P_hi = ar.take<__nv_bfloat16>((size_t)m.Dp * m.Dp);
    m.P_lo = ar.take<__nv_bfloat16>((size_t)m.Dp * m.Dp);
    m.Pmu = m.d.mu ? (void*)ar.take<float>(D) : nullptr;
    if (!ar.ok()) {
        m.P_hi = m.P_lo = nullptr;
        m.Pmu = nullptr;
        set_error("bk_model_create: workspace too small for the tensor-core operands (%zu < %zu)", ws_bytes,
                  ar.off);
        return BK_E_WORKSPACE;
    }
    const int64_t n = m.Dp * m.Dp;
    tc::k_split_matrix<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)m.d.P, D, (int)m.Dp, m.P_hi,
                                                                     m.P_lo);
    BK_LAUNCH_CHECK();
    if (m.Pmu) {
        tc::k_pmu<<<(unsigned)(((int64_t)D * 32 + 255) / 256), 256, 0, st>>>((const float*)m.d.P,
                                                                             (const float*)m.d.mu, D,
                                                                             (float*)m.Pmu);
        BK_LAUNCH_CHECK();
    }
    return BK_OK;
}

size_t dense_tc_hmc_ws_bytes(const Model& m, int64_t C) {
    const size_t n = (size_t)align_up((size_t)C, tc::BN) * m.Dp;   // tile-blocked, padded
    const size_t nb = n;                                           // box-blocked bf16, padded
    // three positions (state / q_{n-1} / q_n rotate) and two gradients (state / endpoint) fp32; h0; q_hi x2, q_lo
    // bf16; eval scratch for the cache refresh
    const size_t n_tiles = align_up((size_t)C, tc::BN) / tc::BN;
    return 5 * align_up(n * 4, 256) + align_up((size_t)C * 4, 256) + 3 * align_up(nb * 2, 256) +
           align_up(tc::MAX_FUSED_STEPS * n_tiles * 4 + 256, 256) + model_eval_ws_bytes(m, C) + 2048;
}

// gradient of the dense plugin on tensor cores (3-pass split): theta [C,D] -> grad [C,D]
int dense_tc_grad(const Model& m, const float* theta, int64_t C, float* grad, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
    Arena ar(ws, ws_bytes);
    const size_t nb = (size_t)align_up((size_t)C, tc::BN) * m.Dp;   // box-blocked, padded chains
    __nv_bfloat16* hi = ar.take<__nv_bfloat16>(nb);
    __nv_bfloat16* lo = ar.take<__nv_bfloat16>(nb);
    if (!ar.ok()) { set_error("dense_tc_grad: workspace too small"); return BK_E_WORKSPACE; }
    const int D = (int)m.d.dims;
    tc::k_split_rows<<<(unsigned)((C * m.Dp + 255) / 256), 256, 0, st>>>(theta, C, D, (int)m.Dp, hi, lo);
    BK_LAUNCH_CHECK();
    tc::StepArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = TC_MODE_GRAD; a.n_pass = 3; a.C = C; a.D = D; a.Dp = (int)m.Dp;
    a.m_tiles = (int)(m.Dp / tc::BM); a.n_tiles = (C + tc::BN - 1) / tc::BN; a.kblocks = (int)(m.Dp / tc::BK);
    a.cvec = (const float*)m.Pmu; a.g_out = grad;
    return tc::launch_tc(m, a, hi, lo, st);
}

int dense_tc_hmc(const Model& m, float* theta, float* lp, float* grad, int32_t* cache_valid, int64_t C,
                 double eps, int L, const float* metric, int64_t n_draws, const bk_rng* rng,
                 const bk_draw_out& out, void* ws, size_t ws_bytes, cudaStream_t st, bool report_lp) {
    const int D = (int)m.d.dims;
    const size_t n = (size_t)align_up((size_t)C, tc::BN) * m.Dp;   // tile-blocked, padded
    const size_t nb = n;                                           // box-blocked bf16, padded
    Arena ar(ws, ws_bytes);
    // Q[x] = the chain state (q_0 of the running trajectory), Q[y] / Q[z] = q_{n-1} / q_n (roles swap every step);
    // G[ga] = gradient at the state, G[gb] = endpoint gradient.  After every draw but the last the arrays holding
    // q_L / g(q_L) BECOME the state arrays (k_hmc_turn_tc), so the indices rotate instead of the data moving.
    float* Q[3] = {ar.take<float>(n), ar.take<float>(n), ar.take<float>(n)};
    float* G[2] = {ar.take<float>(n), ar.take<float>(n)};
    float* h0 = ar.take<float>(C);
    __nv_bfloat16* qhi[2] = {ar.take<__nv_bfloat16>(nb), ar.take<__nv_bfloat16>(nb)};
    __nv_bfloat16* qlo = ar.take<__nv_bfloat16>(nb);
    uint32_t* sync = ar.take<uint32_t>(tc::MAX_FUSED_STEPS * (align_up((size_t)C, tc::BN) / tc::BN) + 64);
    uint32_t* err_word = sync + tc::MAX_FUSED_STEPS * (align_up((size_t)C, tc::BN) / tc::BN);
    const size_t ebytes = model_eval_ws_bytes(m, C);
    void* ews = ar.take<char>(ebytes);
    if (!ar.ok()) {
        set_error("bk_hmc_diag_sample: workspace too small (need %zu bytes, got %zu)", ar.off, ws_bytes);
        return BK_E_WORKSPACE;
    }
    int rc;
    if (!cache_valid || !*cache_valid) {
        rc = model_eval(m, theta, C, lp, grad, ews, ebytes, st);   // same functions as the trajectory ends: 3-pass split tensor-core gradient, lp = 0.5 (q - mu).g (model.cu eval_t)
        if (rc) return rc;
        if (cache_valid) *cache_valid = 1;
    }
    // pad dims [D, Dp) of the bf16 operands are zeroed by the kernels that write them (begin, STEP
    // epilogue); rows of pad chains only feed output columns nobody reads
    tc::HmcTcArgs h;
    memset(&h, 0, sizeof(h));
    h.theta = theta; h.lp = lp; h.grad = grad; h.h0 = h0;
    h.metric = metric; h.mu = (const float*)m.d.mu; h.C = C; h.D = D; h.Dp = (int)m.Dp; h.L = L;
    h.m_tiles = (int)(m.Dp / tc::BM);
    h.eps = (float)eps; h.half_eps = (float)(0.5 * eps); h.inv_eps = (float)(1.0 / eps); h.rng = *rng;
    h.draws = (float*)out.draws; h.logp = (float*)out.logp; h.accept = out.accept;
    h.out_lp = report_lp ? 1 : 0;
    h.err = err_word;
    BK_CUDA(cudaMemsetAsync(err_word, 0, sizeof(uint32_t), st));
    tc::StepArgs a;
    memset(&a, 0, sizeof(a));
    a.err = err_word;
    a.C = C; a.D = D; a.Dp = (int)m.Dp;
    a.m_tiles = (int)(m.Dp / tc::BM); a.n_tiles = (C + tc::BN - 1) / tc::BN; a.kblocks = (int)(m.Dp / tc::BK);
    a.eps = (float)eps; a.metric = metric; a.cvec = (const float*)m.Pmu;
    a.g_blocked = 1;
    { const char* e = getenv("BK_TC_DEBUG"); a.debug = e ? atoi(e) : 0; }
    static int turn_env = -1;   // BK_TC_TURN=0 (diagnostic): end + begin as two kernels through theta / grad
    if (turn_env < 0) { const char* e = getenv("BK_TC_TURN"); turn_env = (e && e[0] == '0') ? 0 : 1; }
    const unsigned wblocks = (unsigned)((C * 32 + 255) / 256);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = D % 4 == 0 && al16(theta) && al16(grad) && al16(metric) && al16(m.d.mu) &&
                     al16(out.draws) && (rng->mode != BK_RNG_INJECTED || al16(rng->normals));
    int x = 0, y = 1, z = 2, ga = 0, gb = 1;
    bool state_blocked = false, begun = false;
    for (int64_t t = 0; t < n_draws; ++t) {
        if (!begun) {
            h.q_hi = qhi[0]; h.q_lo = qlo;
            h.qm = Q[x]; h.q = Q[y];                // begin: q_0 = theta -> Q[x], q_1 -> Q[y]
            if (vec) tc::k_hmc_begin_tc<true><<<wblocks, 256, 0, st>>>(h, t, L == 1 ? 1 : 0);
            else tc::k_hmc_begin_tc<false><<<wblocks, 256, 0, st>>>(h, t, L == 1 ? 1 : 0);
            BK_LAUNCH_CHECK();
        }
        int cur = 0, iprev = x, icur = y;           // q_{s-1}, q_s
        bool first = true;
        const int max_fuse = tc::max_fused_steps(a.m_tiles, a.debug);
        for (int s = 1; s < L;) {   // fused gradient + kick + drift, bf16 operands; up to max_fuse steps per launch
            const int ns = (L - s) < max_fuse ? (L - s) : max_fuse;
            a.mode = TC_MODE_STEP; a.n_pass = 1;
            a.n_steps = ns; a.sync = sync;
            a.q_hi_next = qhi[cur ^ 1]; a.q_hi_cur = qhi[cur];
            a.q_lo_next = (s + ns == L) ? qlo : nullptr;   // the last step's operand also gets its low half
            a.q_prev = Q[iprev]; a.q_cur = Q[icur];
            a.q_out0 = first ? Q[z] : nullptr;      // the trajectory's first step leaves the state Q[x] intact
            a.g_out = nullptr;
            rc = tc::launch_tc(m, a, qhi[cur], nullptr, st);
            if (rc) return rc;
            if (first) {
                // step 0 wrote Q[z]; later steps ping-pong between Q[y] and Q[z]
                if (ns & 1) { iprev = y; icur = z; } else { iprev = z; icur = y; }
                first = false;
            } else if (ns & 1) {
                const int t2 = iprev; iprev = icur; icur = t2;
            }
            if (ns & 1) cur ^= 1;
            s += ns;
        }
        a.n_steps = 1;
        // endpoint gradient with the 3-pass split: enters the Hamiltonian and the cache
        a.mode = TC_MODE_GRAD; a.n_pass = 3; a.q_hi_next = nullptr; a.q_lo_next = nullptr; a.q_out0 = nullptr;
        a.g_out = G[gb];
        rc = tc::launch_tc(m, a, qhi[cur], qlo, st);
        if (rc) return rc;
        h.q = Q[icur]; h.qm = Q[iprev]; h.gq = G[gb];
        h.qs = Q[x]; h.gs = G[ga]; h.old_blocked = state_blocked ? 1 : 0;
        if (t + 1 < n_draws && turn_env) {
            // end of draw t + begin of draw t + 1: the arrays holding q_L / g(q_L) become the state arrays
            h.nq_hi = qhi[0]; h.nq_lo = qlo;
            if (vec) {
                tc::k_hmc_turn_tc<true><<<wblocks, 256, 0, st>>>(h, t, L == 1 ? 1 : 0);
            } else {
                tc::k_hmc_turn_tc<false><<<wblocks, 256, 0, st>>>(h, t, L == 1 ? 1 : 0);
            }
            BK_LAUNCH_CHECK();
            const int nx = icur, ny = iprev, nz = 3 - icur - iprev;
            x = nx; y = ny; z = nz;
            const int tg = ga; ga = gb; gb = tg;
            state_blocked = true;
            begun = true;
        } else {
            if (vec) tc::k_hmc_end_tc<true><<<wblocks, 256, 0, st>>>(h, t);
            else tc::k_hmc_end_tc<false><<<wblocks, 256, 0, st>>>(h, t);
            BK_LAUNCH_CHECK();
            state_blocked = false;   // theta / grad hold the state again
            begun = false;
        }
    }
    return BK_OK;
}

}  // namespace bk
