// K10: edge-aware smoothness of the mean-normalised depth (reference loss_functions.py:132-159,
// get_smooth_loss :133-152), forward and backward, for njobs images per launch.
//
// HBM-bound stencil: 16 B/pixel forward (depth + 3 image planes), 20 B/pixel backward.  Neighbour
// reads are served by L1 (each row segment is re-used by the thread rows above/below).
//
// Backward uses Euler's theorem for the 1-homogeneous loss L(d_hat): sum_q dL/dd_hat(q) d_hat(q) = L,
// so the gradient through the per-image mean needs only the per-sample loss saved by the forward:
//   dL/dd(p) = dL/dd_hat(p) / (mu + eps)  -  g * L_b / (H W (mu + eps)).
#include "common.cuh"

namespace scsfm {

constexpr int SM_THREADS = 256;
constexpr int SM_ROWS = 4;  // pixels per thread (a CTA covers SM_ROWS*SM_THREADS consecutive pixels)

struct SmoothJobs {
    ScsfmSmoothJob j[SCSFM_MAX_JOBS];
};

// stats layout per (job, b): double[4] = { sum(depth), Sx, Sy, - }
constexpr int SMOOTH_STATS = 4;

__global__ void __launch_bounds__(SM_THREADS)
smooth_mean_kernel(SmoothJobs jobs, int B, int HW, double* __restrict__ stats) {
    __shared__ float red[32];
    const int job = blockIdx.y / B, b = blockIdx.y % B;
    const float* d = jobs.j[job].depth + (size_t)b * HW;
    float acc = 0.f;
    for (int i = blockIdx.x * SM_THREADS + threadIdx.x; i < HW; i += gridDim.x * SM_THREADS) acc += __ldg(d + i);
    acc = block_sum<SM_THREADS / 32>(acc, red);
    if (threadIdx.x == 0) atomicAdd(stats + (size_t)blockIdx.y * SMOOTH_STATS, (double)acc);
}

__device__ __forceinline__ float edge_weight(const float* __restrict__ img, int HW, int p, int q) {
    const float a = fabsf(__ldg(img + p) - __ldg(img + q));
    const float b = fabsf(__ldg(img + HW + p) - __ldg(img + HW + q));
    const float c = fabsf(__ldg(img + 2 * HW + p) - __ldg(img + 2 * HW + q));
    return expf(-((a + b + c) / 3.0f));
}

__global__ void __launch_bounds__(SM_THREADS)
smooth_fwd_kernel(SmoothJobs jobs, int B, int H, int W, double* __restrict__ stats) {
    __shared__ float red[32];
    const int job = blockIdx.z / B, b = blockIdx.z % B;
    const int HW = H * W;
    const float* d = jobs.j[job].depth + (size_t)b * HW;
    const float* img = jobs.j[job].img + (size_t)b * 3 * HW;
    double* st = stats + (size_t)blockIdx.z * SMOOTH_STATS;
    const float denom = (float)(st[0] / (double)HW) + 1e-7f;   // mean + 1e-7 (loss_functions.py:139-140)
    float ax = 0.f, ay = 0.f;
#pragma unroll
    for (int r = 0; r < SM_ROWS; ++r) {
        const int p = (blockIdx.x * SM_ROWS + r) * SM_THREADS + threadIdx.x;   // flat pixel index: coalesced rows
        if (p < HW) {
            const int y = p / W, x = p - y * W;
            const float dc = __ldg(d + p) / denom;
            if (x + 1 < W) ax += fabsf(dc - __ldg(d + p + 1) / denom) * edge_weight(img, HW, p, p + 1);
            if (y + 1 < H) ay += fabsf(dc - __ldg(d + p + W) / denom) * edge_weight(img, HW, p, p + W);
        }
    }
    ax = block_sum<SM_THREADS / 32>(ax, red);
    ay = block_sum<SM_THREADS / 32>(ay, red);
    if (threadIdx.x == 0) {
        atomicAdd(st + 1, (double)ax);
        atomicAdd(st + 2, (double)ay);
    }
}

__global__ void smooth_finalize_kernel(const double* __restrict__ stats, int njobs, int B, int H, int W, float* __restrict__ loss_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double nx = (double)B * H * (W - 1), ny = (double)B * (H - 1) * W;
    double total = 0.0;
    for (int j = 0; j < njobs; ++j) {
        double sx = 0.0, sy = 0.0;
        for (int b = 0; b < B; ++b) {
            sx += stats[((size_t)j * B + b) * SMOOTH_STATS + 1];
            sy += stats[((size_t)j * B + b) * SMOOTH_STATS + 2];
        }
        total += (double)((float)(sx / nx) + (float)(sy / ny));
    }
    loss_out[0] = (float)total;
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(SM_THREADS)
smooth_bwd_kernel(SmoothJobs jobs, int B, int H, int W, const double* __restrict__ stats, const float* __restrict__ grad_out) {
    const int job = blockIdx.z / B, b = blockIdx.z % B;
    const int HW = H * W;
    float* gd = jobs.j[job].grad_depth;
    if (gd == nullptr) return;
    gd += (size_t)b * HW;
    const float* d = jobs.j[job].depth + (size_t)b * HW;
    const float* img = jobs.j[job].img + (size_t)b * 3 * HW;
    const double* st = stats + (size_t)blockIdx.z * SMOOTH_STATS;
    const float denom = (float)(st[0] / (double)HW) + 1e-7f;
    const float g = grad_out[0];
    const float inx = 1.0f / ((float)B * (float)H * (float)(W - 1)), iny = 1.0f / ((float)B * (float)(H - 1) * (float)W);
    const float Lb = (float)(st[1] * (double)inx + st[2] * (double)iny);      // this sample's share of the loss
    const float mean_term = g * Lb / ((float)HW * denom);
#pragma unroll
    for (int r = 0; r < SM_ROWS; ++r) {
        const int p = (blockIdx.x * SM_ROWS + r) * SM_THREADS + threadIdx.x;
        if (p >= HW) break;
        const int y = p / W, x = p - y * W;
        const float dc = __ldg(d + p) / denom;
        float acc = 0.f;
        if (x + 1 < W) acc += sgn(dc - __ldg(d + p + 1) / denom) * edge_weight(img, HW, p, p + 1) * inx;
        if (x > 0) acc -= sgn(__ldg(d + p - 1) / denom - dc) * edge_weight(img, HW, p - 1, p) * inx;
        if (y + 1 < H) acc += sgn(dc - __ldg(d + p + W) / denom) * edge_weight(img, HW, p, p + W) * iny;
        if (y > 0) acc -= sgn(__ldg(d + p - W) / denom - dc) * edge_weight(img, HW, p - W, p) * iny;
        red_add(gd + p, g * acc / denom - mean_term);
    }
}

}  // namespace scsfm

using namespace scsfm;

static int check_smooth(const ScsfmSmoothJob* jobs, int njobs, int B, int H, int W) {
    SCSFM_CHECK_ARG(jobs != nullptr && njobs >= 1 && njobs <= SCSFM_MAX_JOBS, "smooth: njobs must be in [1,%d], got %d", SCSFM_MAX_JOBS, njobs);
    SCSFM_CHECK_ARG(B >= 1 && H >= 2 && W >= 2 && (long long)njobs * B <= 65535, "smooth: wrong size B=%d H=%d W=%d", B, H, W);
    for (int i = 0; i < njobs; ++i) SCSFM_CHECK_ARG(jobs[i].depth && jobs[i].img, "smooth: job %d has a null input", i);
    return SCSFM_OK;
}

extern "C" size_t scsfm_smooth_stats_bytes(int njobs, int B) { return (size_t)njobs * B * SMOOTH_STATS * sizeof(double); }

extern "C" int scsfm_smooth_fwd(const ScsfmSmoothJob* jobs_host, int njobs, int B, int H, int W, void* stats, float* loss_out,
                                void* stream) {
    if (int rc = check_smooth(jobs_host, njobs, B, H, W)) return rc;
    SCSFM_CHECK_ARG(stats && loss_out, "smooth_fwd: null stats/loss_out");
    cudaStream_t st = (cudaStream_t)stream;
    SmoothJobs sj;
    memset(&sj, 0, sizeof(sj));
    for (int i = 0; i < njobs; ++i) sj.j[i] = jobs_host[i];
    SCSFM_CHECK_CUDA(cudaMemsetAsync(stats, 0, scsfm_smooth_stats_bytes(njobs, B), st));
    const int HW = H * W;
    int chunks = (HW + SM_THREADS * 8 - 1) / (SM_THREADS * 8);
    if (chunks > 64) chunks = 64;
    smooth_mean_kernel<<<dim3(chunks, njobs * B), SM_THREADS, 0, st>>>(sj, B, HW, (double*)stats);
    SCSFM_CHECK_LAUNCH();
    dim3 grid((H * W + SM_THREADS * SM_ROWS - 1) / (SM_THREADS * SM_ROWS), 1, njobs * B);
    smooth_fwd_kernel<<<grid, SM_THREADS, 0, st>>>(sj, B, H, W, (double*)stats);
    SCSFM_CHECK_LAUNCH();
    smooth_finalize_kernel<<<1, 32, 0, st>>>((const double*)stats, njobs, B, H, W, loss_out);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}

extern "C" int scsfm_smooth_bwd(const ScsfmSmoothJob* jobs_host, int njobs, int B, int H, int W, void* stats,
                                const float* grad_out, void* stream) {
    if (int rc = check_smooth(jobs_host, njobs, B, H, W)) return rc;
    SCSFM_CHECK_ARG(stats && grad_out, "smooth_bwd: null stats/grad_out");
    SmoothJobs sj;
    memset(&sj, 0, sizeof(sj));
    for (int i = 0; i < njobs; ++i) sj.j[i] = jobs_host[i];
    dim3 grid((H * W + SM_THREADS * SM_ROWS - 1) / (SM_THREADS * SM_ROWS), 1, njobs * B);
    smooth_bwd_kernel<<<grid, SM_THREADS, 0, (cudaStream_t)stream>>>(sj, B, H, W, (const double*)stats, grad_out);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}
