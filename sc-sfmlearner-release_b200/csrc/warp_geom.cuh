// Per-pixel geometry of the inverse warp, shared by the fused loss kernels and the
// stand-alone inverse_warp2 kernels.
//
// Follows the arithmetic of reference inverse_warp.py (pixel2cam :29-44, pose_vec2mat :139-154,
// euler2mat :77-112, cam2pixel2 :194-227) and of F.grid_sample(bilinear, align_corners=False)
// as called at inverse_warp.py:262,267 -- same operation order in fp32, no fast-math.
#pragma once
#include "common.cuh"

namespace scsfm {

struct WarpCtx {
    float kinv[9];  // K^-1, row major
    float m[12];    // K * [R|t], row major 3x4  ("proj_cam_to_src_pixel", inverse_warp.py:258)
};

// Euler rotation R = Rx Ry Rz (inverse_warp.py:77-112) written out.
__host__ __device__ inline void euler_to_matrix(float rx, float ry, float rz, float* R) {
    float sx = sinf(rx), cx = cosf(rx), sy = sinf(ry), cy = cosf(ry), sz = sinf(rz), cz = cosf(rz);
    R[0] = cy * cz;                 R[1] = -cy * sz;                R[2] = sy;
    R[3] = cx * sz + sx * sy * cz;  R[4] = cx * cz - sx * sy * sz;  R[5] = -sx * cy;
    R[6] = sx * sz - cx * sy * cz;  R[7] = sx * cz + cx * sy * sz;  R[8] = cx * cy;
}

__device__ inline void make_warp_ctx(const float* __restrict__ K, const float* __restrict__ pose, WarpCtx& c) {
    const float a = K[0], b = K[1], cc = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i = K[8];
    const float A = e * i - f * h, Bc = f * g - d * i, C = d * h - e * g;
    const float inv = 1.0f / (a * A + b * Bc + cc * C);
    c.kinv[0] = A * inv;   c.kinv[1] = (cc * h - b * i) * inv;  c.kinv[2] = (b * f - cc * e) * inv;
    c.kinv[3] = Bc * inv;  c.kinv[4] = (a * i - cc * g) * inv;  c.kinv[5] = (cc * d - a * f) * inv;
    c.kinv[6] = C * inv;   c.kinv[7] = (b * g - a * h) * inv;   c.kinv[8] = (a * e - b * d) * inv;
    float T[12];
    float R[9];
    euler_to_matrix(pose[3], pose[4], pose[5], R);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T[r * 4 + 0] = R[r * 3 + 0];
        T[r * 4 + 1] = R[r * 3 + 1];
        T[r * 4 + 2] = R[r * 3 + 2];
        T[r * 4 + 3] = pose[r];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int col = 0; col < 4; ++col)
            c.m[r * 4 + col] = K[r * 3 + 0] * T[col] + K[r * 3 + 1] * T[4 + col] + K[r * 3 + 2] * T[8 + col];
}

struct Geom {
    float rayx, rayy, rayz;  // K^-1 (x, y, 1)
    float camx, camy, camz;  // ray * depth
    float X, Y, Zr, Z;       // projected point, raw and clamped depth
    bool gradx, grady;       // d(ix)/d(xn) is alive (not overwritten by 2 / not clipped)
    bool valid;              // max(|xn|,|yn|) <= 1
    float fx, fy;            // bilinear fractions
    int x0, y0;              // top-left tap (clamped into [-2, size])
    bool in_x0, in_x1, in_y0, in_y1;
};

// Geometry of target pixel (x, y) with depth D.  PADDING: SCSFM_PAD_ZEROS / SCSFM_PAD_BORDER.
__device__ __forceinline__ Geom project_pixel(const WarpCtx& c, float D, int x, int y, int H, int W, int padding) {
    Geom g;
    const float xf = (float)x, yf = (float)y;
    g.rayx = c.kinv[0] * xf + c.kinv[1] * yf + c.kinv[2];
    g.rayy = c.kinv[3] * xf + c.kinv[4] * yf + c.kinv[5];
    g.rayz = c.kinv[6] * xf + c.kinv[7] * yf + c.kinv[8];
    g.camx = g.rayx * D;
    g.camy = g.rayy * D;
    g.camz = g.rayz * D;
    g.X = c.m[0] * g.camx + c.m[1] * g.camy + c.m[2] * g.camz + c.m[3];
    g.Y = c.m[4] * g.camx + c.m[5] * g.camy + c.m[6] * g.camz + c.m[7];
    g.Zr = c.m[8] * g.camx + c.m[9] * g.camy + c.m[10] * g.camz + c.m[11];
    g.Z = fmaxf(g.Zr, 1e-3f);
    float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fdiv_rn(g.X, g.Z)), (float)(W - 1)), 1.0f);
    float yn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fdiv_rn(g.Y, g.Z)), (float)(H - 1)), 1.0f);
    g.gradx = true;
    g.grady = true;
    if (padding == SCSFM_PAD_ZEROS) {
        // inverse_warp.py:219-224: out-of-range coordinates are overwritten with 2 (gradient cut)
        if (xn > 1.0f || xn < -1.0f) { xn = 2.0f; g.gradx = false; }
        if (yn > 1.0f || yn < -1.0f) { yn = 2.0f; g.grady = false; }
    }
    g.valid = fmaxf(fabsf(xn), fabsf(yn)) <= 1.0f;  // NaN -> false, like (abs().max() <= 1)
    // grid_sampler unnormalize, align_corners=False
    float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(xn, 1.0f), (float)W), 1.0f), 2.0f);
    float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(yn, 1.0f), (float)H), 1.0f), 2.0f);
    if (padding == SCSFM_PAD_BORDER) {
        // clip_coordinates_set_grad: zero gradient at and beyond the limits
        if (!(ix > 0.0f)) { ix = 0.0f; g.gradx = false; }
        else if (ix >= (float)(W - 1)) { ix = (float)(W - 1); g.gradx = false; }
        if (!(iy > 0.0f)) { iy = 0.0f; g.grady = false; }
        else if (iy >= (float)(H - 1)) { iy = (float)(H - 1); g.grady = false; }
    }
    const float x0f = floorf(ix), y0f = floorf(iy);
    g.fx = ix - x0f;
    g.fy = iy - y0f;
    g.in_x0 = (x0f >= 0.0f) && (x0f <= (float)(W - 1));
    g.in_x1 = (x0f >= -1.0f) && (x0f <= (float)(W - 2));
    g.in_y0 = (y0f >= 0.0f) && (y0f <= (float)(H - 1));
    g.in_y1 = (y0f >= -1.0f) && (y0f <= (float)(H - 2));
    g.x0 = (int)fminf(fmaxf(x0f, -2.0f), (float)W);   // NaN -> -2 via fmaxf semantics; flags are false anyway
    g.y0 = (int)fminf(fmaxf(y0f, -2.0f), (float)H);
    return g;
}

// The four bilinear taps of a plane (values of out-of-image taps are 0).
struct Taps {
    float v00, v01, v10, v11;  // v[dy][dx]
};

__device__ __forceinline__ Taps gather_taps(const Geom& g, const float* __restrict__ plane, int W, int shift) {
    Taps t;
    const int ws = W >> shift;
    const int xa = g.x0 >> shift, xb = (g.x0 + 1) >> shift, ya = g.y0 >> shift, yb = (g.y0 + 1) >> shift;
    t.v00 = (g.in_y0 && g.in_x0) ? __ldg(plane + ya * ws + xa) : 0.0f;
    t.v01 = (g.in_y0 && g.in_x1) ? __ldg(plane + ya * ws + xb) : 0.0f;
    t.v10 = (g.in_y1 && g.in_x0) ? __ldg(plane + yb * ws + xa) : 0.0f;
    t.v11 = (g.in_y1 && g.in_x1) ? __ldg(plane + yb * ws + xb) : 0.0f;
    return t;
}

__device__ __forceinline__ float blend(const Geom& g, const Taps& t) {
    // nw*(1-fx)(1-fy) + ne*fx(1-fy) + sw*(1-fx)fy + se*fx*fy  (ATen grid_sampler_2d order)
    const float gx = 1.0f - g.fx, gy = 1.0f - g.fy;
    return t.v00 * (gx * gy) + t.v01 * (g.fx * gy) + t.v10 * (gx * g.fy) + t.v11 * (g.fx * g.fy);
}

// d(blend)/d(ix), d(blend)/d(iy) (taps outside the image contribute nothing)
__device__ __forceinline__ void blend_grad(const Geom& g, const Taps& t, float& dix, float& diy) {
    const float gx = 1.0f - g.fx, gy = 1.0f - g.fy;
    dix = (t.v01 - t.v00) * gy + (t.v11 - t.v10) * g.fy;
    diy = (t.v10 - t.v00) * gx + (t.v11 - t.v01) * g.fx;
}

}  // namespace scsfm
