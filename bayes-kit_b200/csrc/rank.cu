// Rank normalisation on device (rhat.py:27-108): ranks of all draws of one parameter over the
// concatenation of its chains, then  z = Phi^-1((rank - 0.325) / (S - 0.25)).
//
// The reference ranks with argsort().argsort() (rhat.py:52).  Here: a hand-written stable LSD
// radix sort (8-bit digits) of (order-preserving integer image of the value, flat index) pairs --
// per digit a tile histogram, a two-level exclusive scan and a stable scatter in which every warp
// ranks its elements among equal digits with __match_any_sync -- followed by one pass that turns
// sorted position r into rank r + 1 and its normal score (normcdfinv, a few ulp from scipy's ndtri).
// Ties are ranked in flattened (chain-major) order: the reference's default sort is unstable, so
// its tie order is implementation defined.
#include "diag.h"

namespace bk {

constexpr int RS_WARPS = 8, RS_ITEMS = 8;
constexpr int RS_THREADS = RS_WARPS * 32;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;      // 2048 keys per CTA
constexpr int RS_BINS = 256;

template <typename K> struct KeyOf;
template <> struct KeyOf<uint32_t> {
    __device__ static uint32_t make(double v) {
        const uint32_t b = __float_as_uint((float)v);
        return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
};
template <> struct KeyOf<uint64_t> {
    __device__ static uint64_t make(double v) {
        const uint64_t b = (uint64_t)__double_as_longlong(v);
        return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    }
};

// keys[i] = sortable image of x[chain, draw], vals[i] = i, i = chain * N + draw (np.concatenate order)
template <typename K>
__global__ void k_rank_load(SeriesView v, int64_t n, K* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = i / v.N, t = i - c * v.N;
    double x = v.at(c, t);
    if (x == 0.0) x = 0.0;                     // -0.0 and +0.0 compare equal: same key
    keys[i] = KeyOf<K>::make(x);
    vals[i] = (uint32_t)i;
}

// per-tile digit counts, stored digit-major: hist[d * nb + tile]
template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const K* __restrict__ keys, int64_t n, int shift,
                                                           uint32_t* __restrict__ hist, int64_t nb) {
    __shared__ uint32_t h[RS_BINS];
    for (int i = threadIdx.x; i < RS_BINS; i += RS_THREADS) h[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int64_t i = base + j * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RS_BINS; i += RS_THREADS) hist[(int64_t)i * nb + blockIdx.x] = h[i];
}

// one CTA per digit: exclusive scan of that digit's tile counts in place, total -> digit_tot[d]
__global__ void __launch_bounds__(1024) k_radix_scan_tiles(uint32_t* __restrict__ hist, int64_t nb,
                                                           uint32_t* __restrict__ digit_tot) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    uint32_t* row = hist + (int64_t)blockIdx.x * nb;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t b0 = 0; b0 < nb; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const uint32_t x = i < nb ? row[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int k = 0; k < w; ++k) woff += wsum[k];
        const uint32_t carry = carry_s;
        if (i < nb) row[i] = carry + woff + inc - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_tot[blockIdx.x] = carry_s;
}

// exclusive scan of the 256 digit totals (one warp)
__global__ void k_radix_scan_digits(const uint32_t* __restrict__ digit_tot, uint32_t* __restrict__ digit_base) {
    const int lane = threadIdx.x;
    uint32_t carry = 0;
    for (int b0 = 0; b0 < RS_BINS; b0 += 32) {
        const uint32_t x = digit_tot[b0 + lane];
        uint32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        digit_base[b0 + lane] = carry + inc - x;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// stable scatter: element order inside a tile is (warp, item, lane).  The tile is first sorted by digit
// INSIDE shared memory (each element's tile-local position = exclusive digit offset + its stable rank
// among equal digits), then written out in that order: elements with equal digits are neighbours, so
// each digit's run lands as one contiguous burst instead of 2048 isolated 4-byte writes.
template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const K* __restrict__ keys_in,
                                                              const uint32_t* __restrict__ vals_in, int64_t n,
                                                              int shift, const uint32_t* __restrict__ hist,
                                                              int64_t nb, const uint32_t* __restrict__ digit_base,
                                                              K* __restrict__ keys_out,
                                                              uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t cnt[RS_WARPS][RS_BINS + 1];     // per-warp digit counts, then running tile-local offsets
    __shared__ uint32_t lbase[RS_BINS + 1];             // tile-local exclusive digit offsets
    __shared__ uint32_t gbase[RS_BINS];                 // global position of the tile's first element per digit
    __shared__ uint32_t wtot[RS_THREADS / 32];
    __shared__ K skey[RS_TILE];
    __shared__ uint32_t sval[RS_TILE];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * (RS_BINS + 1); i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (32 * RS_ITEMS);
    K key[RS_ITEMS];
    uint32_t val[RS_ITEMS], dig[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int64_t i = base + j * 32 + lane;
        const bool ok = i < n;
        key[j] = ok ? keys_in[i] : (K)0;
        val[j] = ok ? vals_in[i] : 0u;
        dig[j] = ok ? ((uint32_t)(key[j] >> shift) & 0xffu) : (uint32_t)RS_BINS;   // bin 256: padding, sorts last
        atomicAdd(&cnt[w][dig[j]], 1u);
    }
    __syncthreads();
    // thread d: digit d's total over the warps -> exclusive scan over the 256 digits (block scan)
    static_assert(RS_THREADS == RS_BINS, "one thread per digit");
    uint32_t tot = 0;
#pragma unroll
    for (int k = 0; k < RS_WARPS; ++k) tot += cnt[k][threadIdx.x];
    uint32_t inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wtot[w] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int k = 0; k < w; ++k) woff += wtot[k];
    const uint32_t excl = woff + inc - tot;
    {
        const int d = threadIdx.x;
        lbase[d] = excl;
        if (d == RS_BINS - 1) lbase[RS_BINS] = excl + tot;          // padding elements go behind every digit
        gbase[d] = digit_base[d] + hist[(int64_t)d * nb + blockIdx.x];
        uint32_t run = excl;                                        // per-warp tile-local starting offsets
#pragma unroll
        for (int k = 0; k < RS_WARPS; ++k) {
            const uint32_t c = cnt[k][d];
            cnt[k][d] = run;
            run += c;
        }
        if (d == 0) {
            uint32_t pr = 0;
            for (int k = 0; k < RS_WARPS; ++k) { const uint32_t c = cnt[k][RS_BINS]; cnt[k][RS_BINS] = pr; pr += c; }
        }
    }
    __syncthreads();
    const uint32_t pad0 = lbase[RS_BINS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint32_t peers = __match_any_sync(0xffffffffu, dig[j]);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = cnt[w][dig[j]] + rank;
        __syncwarp();
        if (rank == 0) cnt[w][dig[j]] += __popc(peers);
        __syncwarp();
        if (dig[j] == RS_BINS) pos += pad0;
        skey[pos] = key[j];
        sval[pos] = val[j];
    }
    __syncthreads();
    // write out in tile-sorted order: thread i handles local positions i, i + 256, ...
    const int64_t valid = n - (int64_t)blockIdx.x * RS_TILE < RS_TILE ? n - (int64_t)blockIdx.x * RS_TILE : RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = j * RS_THREADS + threadIdx.x;
        if (i < valid) {
            const K k = skey[i];
            const uint32_t d = (uint32_t)(k >> shift) & 0xffu;
            const uint32_t pos = gbase[d] + ((uint32_t)i - lbase[d]);
            keys_out[pos] = k;
            vals_out[pos] = sval[i];
        }
    }
}

// sorted position r holds flat index vals[r]: rank r + 1 (rhat.py:52), z = ndtri((rank - 0.325) / (S - 0.25)) (:106)
__global__ void k_rank_finish(const uint32_t* __restrict__ vals, int64_t n, double* __restrict__ ranks,
                              double* __restrict__ z) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t idx = vals[r];
    const double rank = (double)(r + 1);
    if (ranks) ranks[idx] = rank;
    if (z) z[idx] = normcdfinv((rank - 0.325) / ((double)n - 0.25));
}

template <typename K>
static int rank_normalize_t(const SeriesView& v, int64_t n, double* ranks, double* z, void* ws, size_t ws_bytes,
                            cudaStream_t st) {
    const int64_t nb = (n + RS_TILE - 1) / RS_TILE;
    Arena ar(ws, ws_bytes);
    K* k0 = ar.take<K>((size_t)n);
    K* k1 = ar.take<K>((size_t)n);
    uint32_t* v0 = ar.take<uint32_t>((size_t)n);
    uint32_t* v1 = ar.take<uint32_t>((size_t)n);
    uint32_t* hist = ar.take<uint32_t>((size_t)RS_BINS * nb);
    uint32_t* dtot = ar.take<uint32_t>(RS_BINS);
    uint32_t* dbase = ar.take<uint32_t>(RS_BINS);
    if (!ar.ok()) {
        set_error("bk_rank_normalize: workspace too small (need %zu bytes, got %zu)", ar.off, ws_bytes);
        return BK_E_WORKSPACE;
    }
    k_rank_load<K><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n, k0, v0);
    BK_LAUNCH_CHECK();
    for (int shift = 0; shift < (int)sizeof(K) * 8; shift += 8) {
        k_radix_hist<K><<<(unsigned)nb, RS_THREADS, 0, st>>>(k0, n, shift, hist, nb);
        BK_LAUNCH_CHECK();
        k_radix_scan_tiles<<<RS_BINS, 1024, 0, st>>>(hist, nb, dtot);
        BK_LAUNCH_CHECK();
        k_radix_scan_digits<<<1, 32, 0, st>>>(dtot, dbase);
        BK_LAUNCH_CHECK();
        k_radix_scatter<K><<<(unsigned)nb, RS_THREADS, 0, st>>>(k0, v0, n, shift, hist, nb, dbase, k1, v1);
        BK_LAUNCH_CHECK();
        K* tk = k0; k0 = k1; k1 = tk;
        uint32_t* tv = v0; v0 = v1; v1 = tv;
    }
    k_rank_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v0, n, ranks, z);
    BK_LAUNCH_CHECK();
    return BK_OK;
}

}  // namespace bk

using namespace bk;

extern "C" {

size_t bk_rank_normalize_workspace_bytes(int64_t n_total, int32_t dtype) {
    if (n_total <= 0) return 256;
    const size_t ks = dtype == BK_F64 ? 8 : 4;
    const size_t nb = (size_t)((n_total + RS_TILE - 1) / RS_TILE);
    return 2 * align_up((size_t)n_total * ks, 256) + 2 * align_up((size_t)n_total * 4, 256) +
           align_up(RS_BINS * nb * 4, 256) + 2 * align_up(RS_BINS * 4, 256) + 1024;
}

int bk_rank_normalize(const void* x, int32_t dtype, const bk_series_layout* layout, double* ranks_out,
                      double* z_out, void* ws, size_t ws_bytes, void* stream) {
    BK_CHECK_ARG(x && layout, "bk_rank_normalize: null argument");
    BK_CHECK_ARG(dtype == BK_F32 || dtype == BK_F64, "bk_rank_normalize: bad dtype %d", dtype);
    BK_CHECK_ARG(layout->n_series >= 0 && layout->n_draws >= 0 && layout->n_inner >= 1, "bk_rank_normalize: bad layout");
    const int64_t n = layout->n_series * layout->n_draws;
    BK_CHECK_ARG(n < ((int64_t)1 << 32), "bk_rank_normalize: at most 2^32 - 1 draws per parameter (got %lld)",
                 (long long)n);
    if (n == 0) return BK_OK;
    const SeriesView v{x, dtype, layout->n_series, layout->n_draws, layout->n_inner, layout->outer_stride,
                       layout->inner_stride, layout->draw_stride};
    if (dtype == BK_F64) return rank_normalize_t<uint64_t>(v, n, ranks_out, z_out, ws, ws_bytes, (cudaStream_t)stream);
    return rank_normalize_t<uint32_t>(v, n, ranks_out, z_out, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
