// Constants shared by the network kernels.
#pragma once
#include "common.cuh"

namespace scsfm {
constexpr int PADMODE_ZERO = SCSFM_PADMODE_ZERO;
constexpr int PADMODE_REFLECT = SCSFM_PADMODE_REFLECT;
constexpr int ACT_NONE = SCSFM_ACT_NONE;
constexpr int ACT_RELU = SCSFM_ACT_RELU;
constexpr int ACT_ELU = SCSFM_ACT_ELU;
constexpr int ACT_DISP = SCSFM_ACT_DISP;
constexpr int ROUND_TF32 = SCSFM_ROUND_TF32;

__device__ __forceinline__ float tf32_round(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ float maybe_round(float x, bool on) { return on ? tf32_round(x) : x; }
}  // namespace scsfm
