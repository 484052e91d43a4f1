// DrGhmcDiag (drghmc.py:37-446) for ANY model plugin: dense-precision Gaussian, hierarchical
// logistic regression (gradients on tcgen05 through model_eval), binomial, and separable plugins
// beyond the fused kernel's limits (D > 256).
//
// All C chains advance in lockstep; the data-dependent control flow of the reference -- the
// probabilistic retry (drghmc.py:368-371), acceptance (:378-381), the recursive ghost proposals of
// accept() with their early exit (:424-436) -- becomes per-chain PREDICATION: every trajectory a
// chain could need is integrated for all chains (2^k leapfrog trajectories for proposal k, batched
// gradient evaluations that keep the tensor cores busy), and per-chain status words decide whose
// results count.  The recursion of accept() runs on the host over "levels" of device buffers
// (level 0 = the proposal made from the current state, level l + 1 = ghosts made from level l);
// nothing is read back: no host synchronisation per draw.
//
// The reference's (logp, grad) cache stack (drghmc.py:82, 243-247, 276, 288) only avoids
// recomputation; here every level keeps the (logp, grad) of its own point.
//
// Precision (fp32): every gradient comes from the plugin's fast (tensor-core) path when it has one --
// leapfrog followed by a momentum flip is an exact involution for ANY deterministic gradient function,
// which is what delayed rejection needs -- while every log density that enters a Hastings ratio is
// the precise evaluation.
#include "drghmc_generic.h"

namespace bk {

namespace {

template <typename T>
struct DrgLevel {
    T *q, *r, *g;        // [C, D] point, momentum, gradient at the point
    T *lp, *joint;       // [C] log density at q; joint log density lp - kinetic(r)
    T *hast, *a;         // [C] prop_hastings accumulated over this point's ghosts; log accept prob. of this point
    int32_t* dead;       // [C] the ghost recursion of this point hit accept_logp == 0 (drghmc.py:430-432)
};

template <typename T>
struct DrgArgs {
    T *theta, *rho;      // [C, D] chain state (caller)
    T *g_cur, *lp_cur;   // gradient / log density at theta
    T *cur_logp, *cur_hast, *reject_logp;   // [C]
    int32_t *status, *ui, *moved;           // [C] 0 = still proposing, 1 = stopped, 2 = accepted; uniforms used
    int64_t C;
    int D;
    const T* metric;
    T s_keep, s_new;
    int prob_retry;
    bk_rng rng;
    T* draws;
    T* logp;
    int32_t* accept;
    int32_t* n_used;
};

#define DRG_ROW()                                                                  \
    using A = Ar<T>;                                                                \
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;       \
    const int lane = threadIdx.x & 31;                                              \
    if (c >= p.C) return;                                                           \
    const int D = p.D;                                                              \
    const int64_t off = c * (int64_t)D;                                             \
    (void)lane; (void)off;

template <typename T>
__device__ __forceinline__ T drg_uniform(const bk_rng& rng, int64_t C, int64_t c, int64_t t, int k) {
    if (rng.mode == BK_RNG_INJECTED)
        return reinterpret_cast<const T*>(rng.uniforms)[(t * C + c) * rng.n_uniform + k];
    return philox_uniform<T>(rng.seed, (uint32_t)k, (uint32_t)(rng.chain_offset + (uint64_t)c),
                             (uint32_t)(rng.draw_offset + (uint64_t)t));
}
template <typename T>
__device__ __forceinline__ T drg_log_u(T u) { return u > T(0) ? Ar<T>::log_(u) : neg_inf<T>(); }
template <typename T>
__device__ __forceinline__ T drg_log1m_exp(T a) { return Ar<T>::log1p_(-Ar<T>::exp_(a)); }   // np.log1p(-np.exp(a))
// drghmc.py:317: bool * float (False * -inf = nan on purpose)
template <typename T>
__device__ __forceinline__ T drg_retry(int prob_retry, T reject_logp) {
    return prob_retry ? reject_logp : Ar<T>::mul(T(0), reject_logp);
}

// partial momentum refresh (drghmc.py:360-364) and the joint log density of the current draw (:365)
template <typename T>
__global__ void k_drg_refresh(DrgArgs<T> p, int64_t t) {
    DRG_ROW();
    T kin = T(0);
    for (int b = lane; 4 * b < D; b += 32) {
        T z[4];
        if (p.rng.mode == BK_RNG_INJECTED) {
            const T* zp = reinterpret_cast<const T*>(p.rng.normals) + (t * p.C + c) * (int64_t)D;
#pragma unroll
            for (int i = 0; i < 4; ++i) z[i] = (4 * b + i < D) ? zp[4 * b + i] : T(0);
        } else {
            philox_normal4<T>(p.rng.seed, (uint32_t)b, (uint32_t)(p.rng.chain_offset + (uint64_t)c),
                              (uint32_t)(p.rng.draw_offset + (uint64_t)t), z);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = 4 * b + i;
            if (e >= D) break;
            const T r = A::add(A::mul(p.rho[off + e], p.s_keep), A::mul(p.s_new, z[i]));
            p.rho[off + e] = r;
            kin = A::add(kin, A::mul(r, p.metric ? A::mul(p.metric[e], r) : r));
        }
    }
    kin = warp_sum(kin);
    if (lane == 0) {
        p.cur_logp[c] = A::sub(p.lp_cur[c], A::mul(T(0.5), kin));
        p.cur_hast[c] = T(0);
        p.reject_logp[c] = T(0);
        p.status[c] = 0;
        p.ui[c] = 0;
        p.moved[c] = 0;
    }
}

// retry test of proposal k (drghmc.py:368-371): consumes one uniform of every chain still proposing
template <typename T>
__global__ void k_drg_retry(DrgArgs<T> p, int64_t t) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C || p.status[c] != 0) return;
    const T u = drg_uniform<T>(p.rng, p.C, c, t, p.ui[c]++);
    if (!(drg_log_u(u) < drg_retry<T>(p.prob_retry, p.reject_logp[c]))) p.status[c] = 1;
}

// first half kick + drift of a trajectory (drghmc.py:276-278): from (q0, r0, g0) into (q, r)
template <typename T>
__global__ void k_drg_first(DrgArgs<T> p, const T* __restrict__ q0, const T* __restrict__ r0,
                            const T* __restrict__ g0, T* __restrict__ q, T* __restrict__ r, T eps, T half) {
    DRG_ROW();
    for (int e = lane; e < D; e += 32) {
        const T g = g0[off + e];
        const T mg = p.metric ? A::mul(p.metric[e], g) : g;
        const T rm = A::add(r0[off + e], A::mul(half, mg));
        r[off + e] = rm;
        q[off + e] = A::add(q0[off + e], A::mul(eps, rm));
    }
}
// interior step (drghmc.py:280-283): r += eps m g(q); q += eps r
template <typename T>
__global__ void k_drg_step(DrgArgs<T> p, T* __restrict__ q, T* __restrict__ r, const T* __restrict__ g, T eps) {
    DRG_ROW();
    for (int e = lane; e < D; e += 32) {
        const T gg = g[off + e];
        const T mg = p.metric ? A::mul(p.metric[e], gg) : gg;
        const T rm = A::add(r[off + e], A::mul(eps, mg));
        r[off + e] = rm;
        q[off + e] = A::add(q[off + e], A::mul(eps, rm));
    }
}
// last half kick (drghmc.py:285-286), momentum flip (:345), joint log density of the proposal (accept():
// prop_logp, :418) and reset of this point's ghost bookkeeping
template <typename T>
__global__ void k_drg_last(DrgArgs<T> p, DrgLevel<T> lv, T half) {
    DRG_ROW();
    T kin = T(0);
    for (int e = lane; e < D; e += 32) {
        const T gg = lv.g[off + e];
        const T m = p.metric ? p.metric[e] : T(1);
        const T mg = p.metric ? A::mul(m, gg) : gg;
        const T r = -A::add(lv.r[off + e], A::mul(half, mg));
        lv.r[off + e] = r;
        kin = A::add(kin, A::mul(r, p.metric ? A::mul(m, r) : r));
    }
    kin = warp_sum(kin);
    if (lane == 0) {
        lv.joint[c] = A::sub(lv.lp[c], A::mul(T(0.5), kin));
        lv.hast[c] = T(0);
        lv.dead[c] = 0;
    }
}
// after ghost i of the point at `lv` was evaluated (its log accept probability in gh.a), drghmc.py:428-436
template <typename T>
__global__ void k_drg_ghost_update(DrgArgs<T> p, DrgLevel<T> lv, const T* __restrict__ a_ghost) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C || lv.dead[c]) return;
    const T ai = a_ghost[c];
    if (ai == T(0)) lv.dead[c] = 1;
    else lv.hast[c] = Ar<T>::add(lv.hast[c], drg_log1m_exp(ai));
}
// log acceptance probability of the point at `lv` against (cur_hastings, cur_logp) (drghmc.py:438-446)
template <typename T>
__global__ void k_drg_accept_frac(DrgArgs<T> p, DrgLevel<T> lv, const T* __restrict__ ch, const T* __restrict__ cl) {
    using A = Ar<T>;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C) return;
    if (lv.dead[c]) { lv.a[c] = neg_inf<T>(); return; }
    const T ph = lv.hast[c], h = ch[c];
    const T frac = A::add(A::add(A::sub(lv.joint[c], cl[c]), A::sub(ph, h)),
                          A::sub(drg_retry<T>(p.prob_retry, ph), drg_retry<T>(p.prob_retry, h)));
    lv.a[c] = frac < T(0) ? frac : T(0);   // python min(0, frac): nan -> 0
}
// accept test of proposal k for the chains still proposing (drghmc.py:378-385)
template <typename T>
__global__ void k_drg_decide(DrgArgs<T> p, DrgLevel<T> lv, int64_t t) {
    DRG_ROW();
    if (p.status[c] != 0) return;        // warp-uniform: one chain per warp
    const T a = lv.a[c];
    T u = T(0);
    if (lane == 0) u = drg_uniform<T>(p.rng, p.C, c, t, p.ui[c]);
    u = __shfl_sync(0xffffffffu, u, 0);
    const bool acc = drg_log_u(u) < a;
    if (acc) {
        for (int e = lane; e < D; e += 32) {
            p.theta[off + e] = lv.q[off + e];
            p.rho[off + e] = lv.r[off + e];
            p.g_cur[off + e] = lv.g[off + e];
        }
    }
    __syncwarp();
    if (lane == 0) {
        p.ui[c] += 1;
        if (acc) {
            p.lp_cur[c] = lv.lp[c];
            p.cur_logp[c] = lv.joint[c];
            p.status[c] = 2;
            p.moved[c] = 1;
        } else {
            const T rej = drg_log1m_exp(a);
            p.reject_logp[c] = rej;
            p.cur_hast[c] = A::add(p.cur_hast[c], rej);
        }
    }
}
// unconditional momentum flip (drghmc.py:388) and the draw's outputs
template <typename T>
__global__ void k_drg_finish(DrgArgs<T> p, int64_t t) {
    DRG_ROW();
    T* dr = p.draws ? p.draws + (t * p.C + c) * (int64_t)D : nullptr;
    for (int e = lane; e < D; e += 32) {
        p.rho[off + e] = -p.rho[off + e];
        if (dr) dr[e] = p.theta[off + e];
    }
    if (lane == 0) {
        if (p.logp) p.logp[t * p.C + c] = p.cur_logp[c];
        if (p.accept) p.accept[t * p.C + c] = p.moved[c];
        if (p.n_used) p.n_used[t * p.C + c] = p.ui[c];
    }
}

template <typename T>
struct Engine {
    const Model& m;
    DrgArgs<T> p;
    DrgLevel<T> lv[DRG_KMAX];
    T eps[DRG_KMAX], half[DRG_KMAX];
    int cnt[DRG_KMAX];
    void* ews;
    size_t ebytes;
    cudaStream_t st;
    bool fast;
    unsigned rb, sb;   // row-kernel / scalar-kernel grid sizes

    // (lp, grad) at q: gradients from the fast path where the plugin has one, density always precise
    int eval(const T* q, T* lp, T* g, bool want_lp) {
        if (fast) {
            int rc = model_eval(m, q, p.C, nullptr, g, ews, ebytes, st, false);
            if (rc || !want_lp) return rc;
            return model_eval(m, q, p.C, lp, nullptr, ews, ebytes, st, true);
        }
        return model_eval(m, q, p.C, lp, g, ews, ebytes, st, true);
    }
    // proposal_map (drghmc.py:319-346) with the parameters of proposal i: from (q0, r0, g0) into level `to`
    int propose(int i, const T* q0, const T* r0, const T* g0, int to) {
        DrgLevel<T>& L = lv[to];
        k_drg_first<T><<<rb, 256, 0, st>>>(p, q0, r0, g0, L.q, L.r, eps[i], half[i]);
        BK_LAUNCH_CHECK();
        for (int s = 1; s < cnt[i]; ++s) {
            if (int rc = eval(L.q, L.lp, L.g, false)) return rc;
            k_drg_step<T><<<rb, 256, 0, st>>>(p, L.q, L.r, L.g, eps[i]);
            BK_LAUNCH_CHECK();
        }
        if (int rc = eval(L.q, L.lp, L.g, true)) return rc;
        k_drg_last<T><<<rb, 256, 0, st>>>(p, L, half[i]);
        BK_LAUNCH_CHECK();
        return BK_OK;
    }
    // accept() (drghmc.py:391-446) for the point at level `at`, proposal number k
    int accept(int k, int at, const T* ch, const T* cl) {
        DrgLevel<T>& L = lv[at];
        for (int i = 0; i < k; ++i) {
            if (int rc = propose(i, L.q, L.r, L.g, at + 1)) return rc;
            if (int rc = accept(i, at + 1, L.hast, L.joint)) return rc;
            k_drg_ghost_update<T><<<sb, 256, 0, st>>>(p, L, lv[at + 1].a);
            BK_LAUNCH_CHECK();
        }
        k_drg_accept_frac<T><<<sb, 256, 0, st>>>(p, L, ch, cl);
        BK_LAUNCH_CHECK();
        return BK_OK;
    }
    int run(int K, int64_t n_draws) {
        // (logp, grad) at the current positions: the reference's cache invariant between draws
        if (int rc = eval(p.theta, p.lp_cur, p.g_cur, true)) return rc;
        for (int64_t t = 0; t < n_draws; ++t) {
            k_drg_refresh<T><<<rb, 256, 0, st>>>(p, t);
            BK_LAUNCH_CHECK();
            for (int k = 0; k < K; ++k) {
                k_drg_retry<T><<<sb, 256, 0, st>>>(p, t);
                BK_LAUNCH_CHECK();
                if (int rc = propose(k, p.theta, p.rho, p.g_cur, 0)) return rc;
                if (int rc = accept(k, 0, p.cur_hast, p.cur_logp)) return rc;
                k_drg_decide<T><<<rb, 256, 0, st>>>(p, lv[0], t);
                BK_LAUNCH_CHECK();
            }
            k_drg_finish<T><<<rb, 256, 0, st>>>(p, t);
            BK_LAUNCH_CHECK();
        }
        return BK_OK;
    }
};

template <typename T>
size_t ws_layout(const Model& m, int64_t C, int K, void* ws, size_t ws_bytes, Engine<T>* e) {
    Arena ar(ws, ws_bytes);
    const size_t n = (size_t)C * m.d.dims;
    T* g_cur = ar.take<T>(n);
    T* sc[4];
    for (auto& s : sc) s = ar.take<T>(C);
    int32_t* ic[3];
    for (auto& s : ic) s = ar.take<int32_t>(C);
    if (e) {
        e->p.g_cur = g_cur;
        e->p.lp_cur = sc[0]; e->p.cur_logp = sc[1]; e->p.cur_hast = sc[2]; e->p.reject_logp = sc[3];
        e->p.status = ic[0]; e->p.ui = ic[1]; e->p.moved = ic[2];
    }
    for (int l = 0; l < K; ++l) {
        DrgLevel<T> L;
        L.q = ar.take<T>(n); L.r = ar.take<T>(n); L.g = ar.take<T>(n);
        L.lp = ar.take<T>(C); L.joint = ar.take<T>(C); L.hast = ar.take<T>(C); L.a = ar.take<T>(C);
        L.dead = ar.take<int32_t>(C);
        if (e) e->lv[l] = L;
    }
    const size_t eb = model_eval_ws_bytes(m, C);
    char* ews = ar.take<char>(eb);
    if (e) { e->ews = ews; e->ebytes = eb; }
    return ar.off + 256;
}

template <typename T>
int run_t(const Model& m, void* theta, void* rho, int64_t C, int K, const double* sizes, const int32_t* counts,
          double damping, int prob_retry, const void* metric, int64_t n, const bk_rng* rng, const bk_draw_out& out,
          int32_t* n_used, void* ws, size_t ws_bytes, cudaStream_t st) {
    Engine<T> e{m};
    memset(&e.p, 0, sizeof(e.p));
    if (!ws || ws_layout<T>(m, C, K, ws, ws_bytes, &e) > ws_bytes) {
        set_error("bk_drghmc_sample: workspace too small (need %zu bytes, got %zu)", ws_layout<T>(m, C, K, nullptr, 0, nullptr),
                  ws_bytes);
        return BK_E_WORKSPACE;
    }
    e.p.theta = (T*)theta;
    e.p.rho = (T*)rho;
    e.p.C = C;
    e.p.D = (int)m.d.dims;
    e.p.metric = (const T*)metric;
    e.p.s_keep = (T)sqrt(1 - damping);
    e.p.s_new = (T)sqrt(damping);
    e.p.prob_retry = prob_retry;
    e.p.rng = *rng;
    e.p.draws = (T*)out.draws;
    e.p.logp = (T*)out.logp;
    e.p.accept = out.accept;
    e.p.n_used = n_used;
    for (int k = 0; k < K; ++k) {
        e.eps[k] = (T)sizes[k];
        e.half[k] = (T)(0.5 * sizes[k]);
        e.cnt[k] = counts[k];
    }
    e.st = st;
    e.fast = sizeof(T) == 4 && model_has_fast_path(m);
    e.rb = (unsigned)((C * 32 + 255) / 256);
    e.sb = (unsigned)((C + 255) / 256);
    return e.run(K, n);
}

}  // namespace

size_t drghmc_generic_ws_bytes(const Model& m, int64_t C, int K) {
    if (K > DRG_KMAX) K = DRG_KMAX;
    if (K < 1) K = 1;
    return m.d.dtype == BK_F64 ? ws_layout<double>(m, C, K, nullptr, 0, nullptr)
                               : ws_layout<float>(m, C, K, nullptr, 0, nullptr);
}

int drghmc_generic(const Model& m, void* theta, void* rho, int64_t C, int K, const double* sizes, const int32_t* counts,
                   double damping, int prob_retry, const void* metric, int64_t n, const bk_rng* rng,
                   const bk_draw_out& out, int32_t* n_used, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (m.d.dtype == BK_F64)
        return run_t<double>(m, theta, rho, C, K, sizes, counts, damping, prob_retry, metric, n, rng, out, n_used, ws,
                             ws_bytes, st);
    return run_t<float>(m, theta, rho, C, K, sizes, counts, damping, prob_retry, metric, n, rng, out, n_used, ws, ws_bytes,
                        st);
}

}  // namespace bk
