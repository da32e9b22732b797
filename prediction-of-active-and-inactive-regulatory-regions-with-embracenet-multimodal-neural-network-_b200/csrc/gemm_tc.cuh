// tcgen05 / TMEM / TMA GEMM back end (bf16 operands, fp32 accumulation in tensor memory).
#pragma once
#include "common.cuh"

namespace emb {

inline int tc_init() { return 0; }

inline int tc_linear_f32(const float*, const float*, const float*, int, int, int, int, float*, cudaStream_t) {
    return set_error(-5, "tensor-core GEMM not built yet");
}

}  // namespace emb
