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
    bool separable() const {
        return d.kind == BK_MODEL_ISO_GAUSS || d.kind == BK_MODEL_DIAG_GAUSS;
    }
};

// returns nullptr (and sets the error) for an unknown handle
const Model* get_model(uint64_t handle);

size_t model_eval_ws_bytes(const Model& m, int64_t C);
// theta [C,D] -> lp [C], grad [C,D] (nullable).  All on `st`.
int model_eval(const Model& m, const void* theta, int64_t C, void* lp, void* grad, void* ws,
               size_t ws_bytes, cudaStream_t st);

// hierarchical logistic regression evaluator (logreg.cu)
size_t hlr_eval_ws_bytes(const Model& m, int64_t C);
int hlr_eval(const Model& m, const void* theta, int64_t C, void* lp, void* grad, void* ws, size_t ws_bytes,
             cudaStream_t st);

}  // namespace bk
