// Shared declarations of the tensor-core convolution kernels (conv_tc.cu: cp.async gather producers; conv_tma.cu:
// TMA halo-patch producers).
#pragma once
#include <cuda.h>

#include "nn_common.cuh"
#include "tc_common.cuh"

namespace scsfm {

constexpr int TBM = 128;            // tile rows (UMMA M)
constexpr int TBK = 32;             // floats per k-block = one 128-byte swizzle row
constexpr int TC_THREADS = 160;       // wgrad kernel: 4 producer/epilogue warps + 1 MMA warp
constexpr int FW_PWARPS = 8;          // forward/dgrad kernel: 8 producer/epilogue warps + 1 MMA warp
constexpr int FW_THREADS = (FW_PWARPS + 1) * 32;
constexpr int A_STAGE_BYTES = TBM * 128;

// The tensor core adds into a TMEM accumulator with truncation: a chain of n tcgen05.mma carries a systematic relative error
// of ~3e-8 * n (measured; 1e-4 after a few thousand).  The kernels whose epilogue warps are busy producing (cp.async gather
// kernels) therefore spread consecutive k-blocks round-robin over NACC accumulators (chains NACC times shorter) and add them
// up in registers (round-to-nearest) in the epilogue.
template <int BN>
struct TcCfg {
    static constexpr int STAGES = 3;
    static constexpr int B_STAGE_BYTES = BN * 128;
    static constexpr int NACC = BN >= 128 ? 2 : 4;
    static constexpr int ACC_COLS = BN < 32 ? 32 : BN;      // column stride between accumulators
    static constexpr int TMEM_COLS = NACC * ACC_COLS;       // 64 .. 256: two CTAs per SM still fit in the 512 columns
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256;
};

// Operand precision: kind::tf32 TRUNCATES the low 13 mantissa bits of whatever fp32 pattern sits in shared memory
// (measured: a systematic -7e-4 relative bias per dot product).  Converting with cvt.rna inside the loaders costs
// ~50% of the loader-bound kernel time, so the operands are rounded ONCE where they are produced instead: every
// kernel that writes a tensor later consumed by a convolution takes the SCSFM_ROUND_TF32 flag, and the weights are
// rounded per optimizer step (scsfm_round_tf32).  The loaders below therefore copy bits unchanged.

__device__ __forceinline__ float tc_act(float v, int act) {
    switch (act & 0xff) {
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_ELU: return v > 0.f ? v : expm1f(v);
        case ACT_DISP: return 10.0f * (1.0f / (1.0f + expf(-v))) + 0.01f;
        default: return v;
    }
}

// Geometry of one (sub-)convolution as the kernel sees it.  A plain convolution uses the identity output map; a
// stride-2 data gradient is run as four parity-class stride-1 sub-convolutions (output pixels 2h+py, 2w+px) whose
// taps are the kernel rows/columns of matching parity -- no multiplications by inserted zeros.
struct TcView {
    int kh, kw;            // tap grid
    int oy0, ox0;          // input row = ho * in_stride + oy0 + dy
    int in_stride;
    int out_sy, out_oy, out_sx, out_ox, out_H, out_W;   // output pixel (ho, wo) -> (ho*out_sy + out_oy, wo*out_sx + out_ox)
    int border;            // 1: the GEMM rows are only the image-border pixels (2*(Ho+Wo)-4 per image), see border_pixel()
};

// Row j of the border-only view -> pixel: top row, bottom row, then the left/right pixels of the rows in between.
__host__ __device__ __forceinline__ void border_pixel(int j, int Ho, int Wo, int& ho, int& wo) {
    if (j < Wo) { ho = 0; wo = j; }
    else if (j < 2 * Wo) { ho = Ho - 1; wo = j - Wo; }
    else { const int k = j - 2 * Wo; ho = 1 + (k >> 1); wo = (k & 1) ? Wo - 1 : 0; }
}
__host__ __device__ __forceinline__ int border_count(int Ho, int Wo) { return 2 * (Ho + Wo) - 4; }


// cuTensorMapEncodeTiled resolved through the runtime (cudaGetDriverEntryPoint): libscsfm.so does not link libcuda, so it
// loads (and its host-side argument checks run) on machines without a driver.  Returns CUDA_ERROR_NOT_FOUND if absent.
CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, cuuint32_t rank, void* gaddr, const cuuint64_t* gdim,
                      const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                      CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob);

// conv_tma.cu: stride-1 zero-padded (sub-)convolutions with kh, kw <= 3 through the TMA halo-patch kernel
bool conv_tma_eligible(const ScsfmConv& p, const TcView& v);
bool conv_tma_forced(const ScsfmConv& p);      // a tile configuration is being forced through ScsfmConv.tune (tests / experiments)
int launch_conv_tma(const ScsfmConv& p, const TcView& v, cudaStream_t st);

// conv_wgrad_tma.cu: stride-1 zero-padded weight gradient with TMA-delivered operands
bool conv_wgrad_tma_eligible(const ScsfmConv& p);
int launch_conv_wgrad_tma(const ScsfmConv& p, cudaStream_t st);

// conv_wgrad_thin.cu: 3x3 stride-1 pad-1 layers with Cout = 16 and Cin in {16, 32} on the fp32 FMA pipes (exact products: no low parts)
bool conv_wgrad_thin_eligible(const ScsfmConv& p);
int launch_conv_wgrad_thin(const ScsfmConv& p, cudaStream_t st);

}  // namespace scsfm
