// Model plugin registry (host side) and batched device evaluation.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace bk {

struct Model {
    bk_model_desc d;
    // derived operands living in the caller-provided model workspace
    void* Pmu = nullptr;            // [D]  P * mu (DENSE with mu), dtype
    __nv_bfloat16* P_hi = nullptr;  // [Dp, Dp] bf16(P), zero padded   (fp32 DENSE: tcgen05 operand)
    __nv_bfloat16* P_lo = nullptr;  // [Dp, Dp] bf16(P - P_hi)
    int64_t Dp = 0;                 // D rounded up to 128
    __nv_bfloat16* Xb = nullptr;    // [Np, 128] bf16(X), zero padded      (fp32 HIER_LOGREG: tcgen05 operands)
    __nv_bfloat16* Xlo = nullptr;   // [Np, 128] bf16(X - Xb): split-precision logits
    float* yp = nullptr;            // [Np] responses, zero padded
    float* yh = nullptr;            // [Np] responses - 1/2 (0 on padded rows)
    bool separable() const {
        return d.kind == BK_MODEL_ISO_GAUSS || d.kind == BK_MODEL_DIAG_GAUSS;
    }
};

// returns nullptr (and sets the error) for an unknown handle
const Model* get_model(uint64_t handle);

size_t model_eval_ws_bytes(const Model& m, int64_t C);
// true when model_eval(..., precise=false) is a genuinely cheaper (tensor-core) evaluation
bool model_has_fast_path(const Model& m);
// theta [C,D] -> lp [C], grad [C,D] (nullable).  All on `st`.  precise = false lets a plugin
// answer with its reduced-precision tensor-core path: allowed for INTERIOR leapfrog
// gradients only (any deterministic gradient keeps the leapfrog map reversible and
// volume preserving); everything that enters a Metropolis test must be precise.
int model_eval(const Model& m, const void* theta, int64_t C, void* lp, void* grad, void* ws,
               size_t ws_bytes, cudaStream_t st, bool precise = true);

// hierarchical logistic regression evaluator (logreg.cu)
size_t hlr_eval_ws_bytes(const Model& m, int64_t C);
int hlr_eval(const Model& m, const void* theta, int64_t C, void* lp, void* grad, void* ws, size_t ws_bytes,
             cudaStream_t st, bool precise);
// tcgen05 path (logreg_tc.cu)
size_t hlr_tc_model_ws_bytes(const bk_model_desc& d);
int hlr_tc_prepare(Model& m, void* ws, size_t ws_bytes, cudaStream_t st);
bool hlr_tc_enabled(const Model& m);
size_t hlr_tc_eval_ws_bytes(const Model& m, int64_t C);
enum { HLR_TC_GRAD = 0, HLR_TC_GRAD_LL = 1, HLR_TC_LP = 2 };
int hlr_tc_partial(const Model& m, const float* theta, int64_t C, void* ws, size_t ws_bytes, float** part_g,
                   float** part_ll, int* n_split, cudaStream_t st, int mode, bool operand_ready = false,
                   __nv_bfloat16** operand = nullptr);
// interior leapfrog step of HMC on this plugin: tensor-core gradient of q, then ONE kernel that finishes
// the gradient, applies kick + drift to (r, q) in place and writes the next launch's bf16 operand.
// operand_ready: the operand of q was written by the previous call (same ws, same C).
int hlr_tc_interior_step(const Model& m, float* q, float* r, int64_t C, float eps, const float* metric,
                         bool operand_ready, void* ws, size_t ws_bytes, cudaStream_t st);
// n_steps of them; one persistent launch when the grid fits the device in one wave (logreg_tc.cu)
int hlr_tc_interior_steps(const Model& m, float* q, float* r, int64_t C, float eps, const float* metric, int n_steps,
                          void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace bk
