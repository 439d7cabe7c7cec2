/*
 * scsfm.h -- C ABI of libscsfm.so: hand-written sm_100a kernels for the SC-SfMLearner
 * training hot path.
 *
 * The reference (JiawangBian/SC-SfMLearner-Release) is pure Python and has no FFI of its
 * own; every entry point below names the reference function (file:line under
 * /root/reference) whose device work it replaces.  Conventions:
 *   - extern "C", plain pointers and sizes; no torch types cross this boundary.
 *   - every pointer is DEVICE memory owned by the caller unless the name ends in _host;
 *     the library keeps no pointer past return and allocates no persistent memory.
 *   - tensors are contiguous fp32; images/depths are NCHW exactly as the reference's
 *     Python API passes them (train.py:254-266).
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising; return 0 on success, negative on error;
 *     scsfm_last_error() gives a thread-local message.
 */
#ifndef SCSFM_H_
#define SCSFM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCSFM_OK 0
#define SCSFM_ERR_ARG (-1)
#define SCSFM_ERR_CUDA (-2)

#define SCSFM_MAX_JOBS 16

/* flags of the pairwise loss (reference loss_functions.py:95, int flags compared with == True) */
#define SCSFM_WITH_SSIM 1
#define SCSFM_WITH_MASK 2
#define SCSFM_WITH_AUTO_MASK 4
/* padding_mode of F.grid_sample (reference inverse_warp.py:262,267) */
#define SCSFM_PAD_ZEROS 0
#define SCSFM_PAD_BORDER 1

const char* scsfm_last_error(void);
int scsfm_version(void);

/* One pair-direction of compute_photo_and_geometry_loss (reference loss_functions.py:84-87):
 * warp `ref_*` into the view of `tgt_*`.  Depth maps may be coarser than the image by a power
 * of two (the reference nearest-upsamples them first, loss_functions.py:77-82): element (y,x)
 * of the full-resolution map is depth[(y >> shift) * (W >> shift) + (x >> shift)]. */
typedef struct ScsfmPairJob {
    const float* tgt_img;    /* [B,3,H,W] */
    const float* ref_img;    /* [B,3,H,W] */
    const float* tgt_depth;  /* [B,1,H>>tgt_shift,W>>tgt_shift] */
    const float* ref_depth;  /* [B,1,H>>ref_shift,W>>ref_shift] */
    const float* pose;       /* [B,6] tx,ty,tz,rx,ry,rz (euler) */
    float* grad_tgt_depth;   /* backward only: accumulated into (atomic +=), same shape as tgt_depth */
    float* grad_ref_depth;   /* backward only: accumulated into (atomic +=) */
    float* grad_pose;        /* backward only: [B,6], accumulated into (+=) */
    int tgt_shift;
    int ref_shift;
} ScsfmPairJob;

/* Optional per-pixel outputs of job 0 (the four returns of inverse_warp2, reference
 * inverse_warp.py:230-269, plus the final mask/diff maps of loss_functions.py:99-113).
 * Any pointer may be NULL. */
typedef struct ScsfmPairMaps {
    float* warped;       /* [B,3,H,W] projected_img */
    float* valid;        /* [B,1,H,W] valid_mask of inverse_warp2 (before auto-mask) */
    float* proj_depth;   /* [B,1,H,W] projected_depth */
    float* comp_depth;   /* [B,1,H,W] computed_depth */
    float* mask;         /* [B,1,H,W] valid mask after the auto-mask */
    float* diff_img;     /* [B,3,H,W] final photometric map */
    float* diff_depth;   /* [B,1,H,W] */
} ScsfmPairMaps;

/* Bytes of the `stats` buffer needed by scsfm_pairwise_fwd/bwd for njobs jobs and batch B. */
size_t scsfm_pairwise_stats_bytes(int njobs, int B);

/* Fused pixel2cam -> pose -> cam2pixel2 -> bilinear sample -> L1 + SSIM -> depth consistency ->
 * auto-mask -> masked sums, for njobs pair-directions in one launch.
 * Replaces compute_pairwise_loss + inverse_warp2 (+ SSIM.forward, mean_on_mask):
 * reference loss_functions.py:95-129, :11-42; inverse_warp.py:29-44,77-112,139-154,194-269.
 *   intrinsics [B,3,3]; flags = SCSFM_WITH_*; padding_mode = SCSFM_PAD_*.
 *   stats: scratch of scsfm_pairwise_stats_bytes() bytes, kept by the caller for the backward.
 *   loss_out[2]: (photo_loss, geometry_loss) summed over the jobs (loss_functions.py:89-90).
 *   maps: optional per-pixel outputs for job 0 (NULL for none). */
int scsfm_pairwise_fwd(const ScsfmPairJob* jobs_host, int njobs, const float* intrinsics, int B, int H, int W,
                       int flags, int padding_mode, void* stats, float* loss_out, const ScsfmPairMaps* maps_host,
                       void* stream);

/* Backward of scsfm_pairwise_fwd: given d(loss)/d(photo) and d(loss)/d(geometry) (device scalars
 * grad_out[2]) accumulates gradients into jobs[i].grad_tgt_depth / grad_ref_depth / grad_pose.
 * Hand-written replacement for autograd through reference loss_functions.py:95-129 and
 * inverse_warp.py:230-269 (grid_sampler_2d_backward scatter, bmm/inverse backward). */
int scsfm_pairwise_bwd(const ScsfmPairJob* jobs_host, int njobs, const float* intrinsics, int B, int H, int W,
                       int flags, int padding_mode, void* stats, const float* grad_out, void* stream);

/* Forward of inverse_warp2 alone (reference inverse_warp.py:230-269): the four maps, no loss. */
int scsfm_inverse_warp2_fwd(const float* img, const float* depth, const float* ref_depth, const float* pose,
                            const float* intrinsics, int B, int H, int W, int padding_mode, float* warped,
                            float* valid, float* proj_depth, float* comp_depth, void* stream);

/* Backward of inverse_warp2 alone: grads of the three differentiable maps -> depth, ref_depth, pose
 * (all accumulated into). grad_* inputs may be NULL (treated as zero). */
int scsfm_inverse_warp2_bwd(const float* img, const float* depth, const float* ref_depth, const float* pose,
                            const float* intrinsics, int B, int H, int W, int padding_mode,
                            const float* grad_warped, const float* grad_proj_depth, const float* grad_comp_depth,
                            float* grad_depth, float* grad_ref_depth, float* grad_pose, void* scratch_12B_doubles,
                            void* stream);

/* pose_vec2mat (reference inverse_warp.py:139-154): [B,6] -> [B,3,4]; rotation_mode 0 = euler, 1 = quat. */
int scsfm_pose_vec2mat(const float* vec, int B, int rotation_mode, float* out, void* stream);

/* One image of compute_smooth_loss (reference loss_functions.py:132-159). */
typedef struct ScsfmSmoothJob {
    const float* depth;   /* [B,1,H,W] (the reference passes scale-0 depth, called "disp" there) */
    const float* img;     /* [B,3,H,W] */
    float* grad_depth;    /* backward only, accumulated into (atomic +=) */
} ScsfmSmoothJob;

size_t scsfm_smooth_stats_bytes(int njobs, int B);

/* Edge-aware smoothness of the mean-normalised map, summed over njobs images
 * (get_smooth_loss, reference loss_functions.py:133-152). loss_out[1]. */
int scsfm_smooth_fwd(const ScsfmSmoothJob* jobs_host, int njobs, int B, int H, int W, void* stats, float* loss_out,
                     void* stream);
int scsfm_smooth_bwd(const ScsfmSmoothJob* jobs_host, int njobs, int B, int H, int W, void* stats,
                     const float* grad_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCSFM_H_ */
