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
}  // namespace scsfm
