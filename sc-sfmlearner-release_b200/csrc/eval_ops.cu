// Validation metrics of the depth network (reference loss_functions.py:163-205, `compute_errors`): per image the Garg /
// NYU crop + depth-range mask, median scaling of the prediction (torch.median = the LOWER median, element (n-1)/2 of the
// sorted masked values), then abs_diff, abs_rel, sq_rel and the three threshold accuracies.
//
// The reference does this with boolean-mask gathers, two sorts and ~20 elementwise kernels per image; here one CTA per
// (image, tensor) finds the exact median with a 4-pass radix select on the (positive) float bit patterns and a second
// kernel accumulates the six sums -- no host synchronisation, no temporaries.  HBM-bound: 2 x 4 passes + 1 pass over
// 8 B/pixel.
#include "nn_common.cuh"

namespace scsfm {

constexpr int EV_THREADS = 1024;

struct EvalGeom {
    int B, H, W, y1, y2, x1, x2;
    float max_depth;
};

__device__ __forceinline__ bool eval_keep(const EvalGeom& g, int pix, float gt) {
    const int y = pix / g.W, x = pix - y * g.W;
    return y >= g.y1 && y < g.y2 && x >= g.x1 && x < g.x2 && gt > 0.1f && gt < g.max_depth;
}

// grid (B, 2): blockIdx.y = 0 -> median of the masked ground truth, 1 -> of the masked, clamped prediction.
// med[b][which] = value of rank (n - 1) / 2 (NaN when the mask is empty), cnt[b] = n.
__global__ void __launch_bounds__(EV_THREADS)
eval_median_kernel(const float* __restrict__ gt, const float* __restrict__ pred, EvalGeom g, float* __restrict__ med, int* __restrict__ cnt) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_rank, s_count;
    const int b = blockIdx.x, which = blockIdx.y, tid = threadIdx.x;
    const int HW = g.H * g.W;
    const float* gi = gt + (size_t)b * HW;
    const float* pi = pred + (size_t)b * HW;
    // count the masked pixels
    if (tid == 0) s_count = 0;
    __syncthreads();
    unsigned local = 0;
    for (int i = tid; i < HW; i += EV_THREADS) local += eval_keep(g, i, __ldg(gi + i)) ? 1u : 0u;
    local = __reduce_add_sync(0xffffffffu, local);
    if ((tid & 31) == 0 && local) atomicAdd(&s_count, local);
    __syncthreads();
    const unsigned n = s_count;
    if (n == 0) {
        if (tid == 0) {
            med[b * 2 + which] = __int_as_float(0x7fc00000);
            if (which == 0) cnt[b] = 0;
        }
        return;
    }
    if (tid == 0) {
        s_prefix = 0;
        s_rank = (n - 1) / 2;              // torch.median: lower median
        if (which == 0) cnt[b] = (int)n;
    }
    // radix select, most significant byte first: all candidates are positive floats, so their bit patterns order like the values
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 256; i += EV_THREADS) hist[i] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < HW; i += EV_THREADS) {
            const float gv = __ldg(gi + i);
            if (!eval_keep(g, i, gv)) continue;
            const float v = which == 0 ? gv : fminf(fmaxf(__ldg(pi + i), 1e-3f), g.max_depth);
            const unsigned u = __float_as_uint(v);
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned r = s_rank, acc = 0;
            int d = 0;
            for (; d < 256; ++d) {
                if (acc + hist[d] > r) break;
                acc += hist[d];
            }
            s_rank = r - acc;
            s_prefix = prefix | ((unsigned)d << shift);
        }
        __syncthreads();
    }
    if (tid == 0) med[b * 2 + which] = __uint_as_float(s_prefix);
}

// grid B: out[b][0..5] = abs_diff, abs_rel, sq_rel, a1, a2, a3 (means over the image's masked pixels), out[b][6..7] = medians
__global__ void __launch_bounds__(EV_THREADS)
eval_metrics_kernel(const float* __restrict__ gt, const float* __restrict__ pred, EvalGeom g, const float* __restrict__ med,
                    const int* __restrict__ cnt, float* __restrict__ out) {
    __shared__ double red[32][6];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int HW = g.H * g.W;
    const float mg = med[b * 2], mp = med[b * 2 + 1];
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int i = tid; i < HW; i += EV_THREADS) {
        const float vg = __ldg(gt + (size_t)b * HW + i);
        if (!eval_keep(g, i, vg)) continue;
        float vp = fminf(fmaxf(__ldg(pred + (size_t)b * HW + i), 1e-3f), g.max_depth);
        vp = __fdiv_rn(__fmul_rn(vp, mg), mp);                         // valid_pred * median(gt) / median(pred), in that order
        const float th = fmaxf(__fdiv_rn(vg, vp), __fdiv_rn(vp, vg));
        const float e = fabsf(vg - vp);
        s[0] += e;
        s[1] += __fdiv_rn(e, vg);
        s[2] += __fdiv_rn(__fmul_rn(vg - vp, vg - vp), vg);
        s[3] += th < 1.25f ? 1.0 : 0.0;
        s[4] += th < 1.25f * 1.25f ? 1.0 : 0.0;
        s[5] += th < 1.25f * 1.25f * 1.25f ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if ((tid & 31) == 0) red[tid >> 5][k] = s[k];
    }
    __syncthreads();
    if (tid < 6) {
        double t = 0;
        for (int w = 0; w < EV_THREADS / 32; ++w) t += red[w][tid];
        const int n = cnt[b];
        out[b * 8 + tid] = n > 0 ? (float)(t / n) : __int_as_float(0x7fc00000);
    }
    if (tid == 6) out[b * 8 + 6] = mg;
    if (tid == 7) out[b * 8 + 7] = mp;
}

}  // namespace scsfm

using namespace scsfm;

// gt, pred [B,H,W]; crop rows [y1,y2) x columns [x1,x2); work: 2*B floats + B ints; out [B][8]
extern "C" int scsfm_compute_errors(const float* gt, const float* pred, int B, int H, int W, int y1, int y2, int x1, int x2,
                                    float max_depth, void* work, float* out, void* stream) {
    SCSFM_CHECK_ARG(gt && pred && work && out && B > 0 && H > 0 && W > 0, "compute_errors: bad arguments");
    SCSFM_CHECK_ARG(0 <= y1 && y1 <= y2 && y2 <= H && 0 <= x1 && x1 <= x2 && x2 <= W && max_depth > 0.1f, "compute_errors: bad crop / depth range");
    EvalGeom g{B, H, W, y1, y2, x1, x2, max_depth};
    float* med = reinterpret_cast<float*>(work);
    int* cnt = reinterpret_cast<int*>(med + 2 * B);
    cudaStream_t st = (cudaStream_t)stream;
    eval_median_kernel<<<dim3(B, 2), EV_THREADS, 0, st>>>(gt, pred, g, med, cnt);
    SCSFM_CHECK_LAUNCH();
    eval_metrics_kernel<<<B, EV_THREADS, 0, st>>>(gt, pred, g, med, cnt, out);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}
