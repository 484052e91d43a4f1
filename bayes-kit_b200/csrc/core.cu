// Error plumbing + ABI bookkeeping.
#include <stdarg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace bk {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- event profiler -----------------------------------------------------------
struct ProfRec { int tag; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;       // recorded since last read
static std::vector<ProfRec> g_free;       // recycled event pairs
static ProfRec* g_open[BK_PROF_NTAGS] = {nullptr};

void prof_begin(int tag, cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    if (!g_free.empty()) { r = g_free.back(); g_free.pop_back(); }
    else {
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    }
    r.tag = tag;
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
    g_open[tag] = &g_recs.back();
}
void prof_end(int tag, cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    // the open record of this tag is the last one pushed with it
    for (auto it = g_recs.rbegin(); it != g_recs.rend(); ++it)
        if (it->tag == tag) { cudaEventRecord(it->b, st); break; }
}
}  // namespace bk

extern "C" {
int bk_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(bk::g_prof_mu);
    bk::g_prof_on = on != 0;
    return BK_OK;
}
int bk_profile_read(int32_t tag, double* total_ms_out, uint64_t* launches_out) {
    std::lock_guard<std::mutex> lk(bk::g_prof_mu);
    double tot = 0;
    uint64_t n = 0;
    std::vector<bk::ProfRec> keep;
    for (auto& r : bk::g_recs) {
        if (r.tag != tag) { keep.push_back(r); continue; }
        float ms = 0;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            tot += ms;
            ++n;
        }
        bk::g_free.push_back(r);
    }
    bk::g_recs.swap(keep);
    if (total_ms_out) *total_ms_out = tot;
    if (launches_out) *launches_out = n;
    return BK_OK;
}
const char* bk_last_error(void) { return bk::g_err; }
int bk_abi_version(void) { return BK_ABI_VERSION; }
uint64_t bk_launch_count(void) { return bk::g_launches.load(); }
}
