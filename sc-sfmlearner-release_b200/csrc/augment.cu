// Training-time image transforms on the device.  Reference: custom_transforms.py applied per
// sample by the dataset classes (datasets/sequence_folders.py:59-62):
//     RandomHorizontalFlip (:46-60)  ->  RandomScaleCrop (:63-89)  ->  ArrayToTensor (:32-43)  ->  Normalize (:21-29)
// The host decodes the JPEGs and draws the random numbers (scsfm/augment.py, the reference's draw order); the batch crosses
// PCIe as uint8 (a quarter of the float32 bytes the reference's loader ships) and one kernel writes the normalised NCHW float
// tensors the networks take -- bit for bit what the reference chain produces:
//
//  * RandomScaleCrop zooms with PIL's Image.resize on the uint8 image: Pillow's BICUBIC (src/libImaging/Resample.c), two
//    separable passes with 22-bit fixed-point taps and an 8-bit intermediate image.  Only the H x W crop window of the zoomed
//    image is ever needed, so an output pixel evaluates its <= 5 vertical taps, each from <= 5 horizontal taps of the source
//    (the flip is an index mirror on the source columns), rounding and clipping to 8 bits after each pass exactly like the
//    two-pass original;
//  * the taps depend on (sample, output coordinate) only: a first tiny kernel computes them in double precision with explicitly
//    rounded operations (__dmul_rn, __dadd_rn, ...: no FMA contraction), i.e. the IEEE sequence of Resample.c's
//    precompute_coeffs / normalize_coeffs_8bpc built for a baseline x86-64;
//  * ArrayToTensor / Normalize: float(u8) / 255, - mean, / std as three correctly rounded float32 operations (true divisions:
//    the reference runs them on the CPU, where torch divides).
// HBM-bound and tiny: 3 B/pixel in (re-read through L1/L2 by the 25-tap window), 12 B/pixel out.
#include "common.cuh"

namespace scsfm {

constexpr int AUG_PRECISION_BITS = 32 - 8 - 2;
constexpr int AUG_TAPS = 5;               // 2 * ceil(support) + 1 with support = 2 (zooming in: the filter is not stretched)
constexpr int AUG_COEF_INTS = 8;          // per (sample, coordinate): first source coordinate, tap count, 5 taps, pad

// Resample.c bicubic_filter (a = -0.5), every operation rounded on its own
__device__ __forceinline__ double aug_bicubic(double x) {
    if (x < 0.0) x = -x;
    if (x < 1.0) {
        // ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        double t = __dsub_rn(__dmul_rn(1.5, x), 2.5);
        t = __dmul_rn(__dmul_rn(t, x), x);
        return __dadd_rn(t, 1.0);
    }
    if (x < 2.0) {
        // (((x - 5) * x + 8) * x - 4) * a
        double t = __dsub_rn(x, 5.0);
        t = __dadd_rn(__dmul_rn(t, x), 8.0);
        t = __dsub_rn(__dmul_rn(t, x), 4.0);
        return __dmul_rn(t, -0.5);
    }
    return 0.0;
}

// params[b] = {flip, scaled_w, scaled_h, offset_x, offset_y}; coef[b][0][x] for the W crop columns, coef[b][1][y] for the H rows
__global__ void augment_coeffs_kernel(const int* __restrict__ params, int B, int H, int W, int* __restrict__ coef) {
    const int per = W + H;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * per) return;
    const int b = i / per, j = i - b * per;
    const int* pr = params + 5 * b;
    const bool is_x = j < W;
    const int in_size = is_x ? W : H;
    const int out_size = is_x ? pr[1] : pr[2];
    const int xx = (is_x ? pr[3] + j : pr[4] + (j - W));            // coordinate in the zoomed image
    // precompute_coeffs with in0 = 0, in1 = in_size; zooming in: filterscale = 1, support = 2, ss = 1
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double center = __dmul_rn((double)xx + 0.5, scale);
    int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, 2.0), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, 2.0), 0.5));
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    if (xmax > AUG_TAPS) xmax = AUG_TAPS;                              // (cannot happen for out_size >= in_size)
    double w[AUG_TAPS];
    double ww = 0.0;
#pragma unroll
    for (int x = 0; x < AUG_TAPS; ++x) {
        w[x] = 0.0;
        if (x < xmax) {
            w[x] = aug_bicubic(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5));
            ww = __dadd_rn(ww, w[x]);
        }
    }
    int* o = coef + (size_t)i * AUG_COEF_INTS;
    o[0] = xmin;
    o[1] = xmax;
#pragma unroll
    for (int x = 0; x < AUG_TAPS; ++x) {
        int k = 0;
        if (x < xmax) {
            const double v = ww != 0.0 ? __ddiv_rn(w[x], ww) : w[x];
            const double s = __dmul_rn(v, (double)(1 << AUG_PRECISION_BITS));
            k = v < 0 ? __double2int_rz(__dadd_rn(-0.5, s)) : __double2int_rz(__dadd_rn(0.5, s));
        }
        o[2 + x] = k;
    }
    o[7] = 0;
}

__device__ __forceinline__ int aug_clip8(int v) {
    v >>= AUG_PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

struct AugNorm {
    float mean[3], std[3];
};

// images [n_img][B][H][W][3] uint8 -> out [n_img][B][3][H][W] float32; one thread per output pixel (three channels)
__global__ void __launch_bounds__(256)
augment_kernel(const unsigned char* __restrict__ images, const int* __restrict__ params, const int* __restrict__ coef, int n_img, int B, int H,
               int W, AugNorm nm, float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int ib = blockIdx.z;                     // image slot * B + sample
    if (x >= W) return;
    const int b = ib % B;
    const bool flip = params[5 * b] != 0;
    const int* cx = coef + ((size_t)b * (W + H) + x) * AUG_COEF_INTS;
    const int* cy = coef + ((size_t)b * (W + H) + W + y) * AUG_COEF_INTS;
    const int4 cx0 = *reinterpret_cast<const int4*>(cx), cx1 = *reinterpret_cast<const int4*>(cx + 4);
    const int4 cy0 = *reinterpret_cast<const int4*>(cy), cy1 = *reinterpret_cast<const int4*>(cy + 4);
    const int kx[AUG_TAPS] = {cx0.z, cx0.w, cx1.x, cx1.y, cx1.z}, ky[AUG_TAPS] = {cy0.z, cy0.w, cy1.x, cy1.y, cy1.z};
    const int xmin = cx0.x, nx = cx0.y, ymin = cy0.x, ny = cy0.y;
    const unsigned char* src = images + (size_t)ib * H * W * 3;
    int v[3] = {1 << (AUG_PRECISION_BITS - 1), 1 << (AUG_PRECISION_BITS - 1), 1 << (AUG_PRECISION_BITS - 1)};
#pragma unroll
    for (int j = 0; j < AUG_TAPS; ++j) {
        if (j < ny) {
            const unsigned char* row = src + (size_t)(ymin + j) * W * 3;
            int h[3] = {1 << (AUG_PRECISION_BITS - 1), 1 << (AUG_PRECISION_BITS - 1), 1 << (AUG_PRECISION_BITS - 1)};
#pragma unroll
            for (int i = 0; i < AUG_TAPS; ++i) {
                if (i < nx) {
                    const int col = flip ? W - 1 - (xmin + i) : xmin + i;       // RandomHorizontalFlip runs first: mirror the source
                    const unsigned char* px = row + 3 * col;
                    h[0] += (int)__ldg(px) * kx[i];
                    h[1] += (int)__ldg(px + 1) * kx[i];
                    h[2] += (int)__ldg(px + 2) * kx[i];
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] += aug_clip8(h[c]) * ky[j];         // 8-bit intermediate of the horizontal pass
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float f = __fdiv_rn((float)aug_clip8(v[c]), 255.0f);                     // ArrayToTensor
        f = __fdiv_rn(__fsub_rn(f, nm.mean[c]), nm.std[c]);                      // Normalize
        out[(((size_t)ib * 3 + c) * H + y) * W + x] = f;
    }
}

}  // namespace scsfm

using namespace scsfm;

extern "C" long long scsfm_augment_workspace_ints(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return (long long)B * (W + H) * AUG_COEF_INTS;
}

extern "C" int scsfm_augment_batch(const unsigned char* images, const int* params, int n_img, int B, int H, int W, const float* mean3,
                                   const float* std3, float* out, int* workspace, long long workspace_ints, void* stream) {
    SCSFM_CHECK_ARG(images && params && out && workspace && mean3 && std3, "augment_batch: null pointer");
    SCSFM_CHECK_ARG(n_img > 0 && B > 0 && H > 0 && W > 0, "augment_batch: bad geometry");
    SCSFM_CHECK_ARG((long long)n_img * B <= 65535 && H <= 65535, "augment_batch: too many images / rows for one launch");
    SCSFM_CHECK_ARG(workspace_ints >= scsfm_augment_workspace_ints(B, H, W), "augment_batch: workspace too small (scsfm_augment_workspace_ints)");
    SCSFM_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "augment_batch: workspace must be 16-byte aligned");
    AugNorm nm;
    for (int c = 0; c < 3; ++c) {
        nm.mean[c] = mean3[c];
        nm.std[c] = std3[c];
        SCSFM_CHECK_ARG(std3[c] != 0.f, "augment_batch: zero std");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = B * (W + H);
    augment_coeffs_kernel<<<(n + 127) / 128, 128, 0, st>>>(params, B, H, W, workspace);
    SCSFM_CHECK_LAUNCH();
    augment_kernel<<<dim3((W + 255) / 256, H, n_img * B), 256, 0, st>>>(images, params, workspace, n_img, B, H, W, nm, out);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}
