// Stride-1 convolution on the tensor cores with TMA halo-patch producers (tcgen05 kind::tf32, accumulators in TMEM).
//
// The cp.async kernel of conv_tc.cu gathers one 128 x 32-float im2col slice per (tap, channel chunk): every input
// pixel crosses L2 -> shared memory kh*kw times and 256 threads spend their issue slots on 16-byte copies.  Here the
// CTA's output tile is a 2-D patch of ONE image (MT*TH rows x TW columns, TH*TW = 128, TW in {8, 16}), and per
// (channel chunk, dx) ONE 4-D tiled TMA load brings the (MT*TH + kh - 1) x TW x 32-channel input patch, already in
// the 128B-swizzled K-major layout the UMMA descriptors expect.  Because TW is a multiple of 8, the im2col operand of
// tap row dy is the SAME patch shifted by dy*TW rows = dy*TW*128 bytes (a multiple of the 1024-byte swizzle atom), so
// the kh vertical taps reuse one load: the input crosses L2 -> smem kw*(1 + (kh-1)/(MT*TH)) times instead of kh*kw.
// Zero padding is the TMA's out-of-bounds fill (negative / past-the-end coordinates); channels beyond Cin in the last
// 32-wide chunk are zero-filled the same way (the matching weight columns then multiply zeros).
//
//   warps 0-7   epilogue only: tcgen05.ld -> bias / residual addend / activation / TF32 rounding / BatchNorm sums
//   warp 8      lane 0: TMA producer (1 activation box + kh weight boxes per stage, mbarrier expect_tx)
//   warp 9      TMEM allocation; lane 0: MMA issuer (MT * kh * 4 tcgen05.mma M128 x BN x K8 per stage)
//
// Used for: forward of every stride-1 layer with kh, kw <= 3 (reflection-padded layers run it with zero padding and
// the cp.async kernel then recomputes the 2*(H+W)-4 border pixels per image, see scsfm_conv2d_fwd_tc), stride-1 data
// gradients, and the four parity-class sub-convolutions of stride-2 data gradients.
#include <stdlib.h>

#include "conv_tc.cuh"

namespace scsfm {

constexpr int TMA_EWARPS = 8;
constexpr int TMA_THREADS = (TMA_EWARPS + 2) * 32;
constexpr int TMA_MAX_KH = 3;

template <int BN, int MT, int STAGES>
struct TmaCfg {
    static constexpr int A_BYTES = (MT * 16 + TMA_MAX_KH - 1) * 8 * 128 > (MT * 8 + TMA_MAX_KH - 1) * 16 * 128
                                       ? (MT * 16 + TMA_MAX_KH - 1) * 8 * 128
                                       : (MT * 8 + TMA_MAX_KH - 1) * 16 * 128;       // worst case of TW = 8 / TW = 16
    static constexpr int A_STAGE = (A_BYTES + 1023) / 1024 * 1024;
    static constexpr int B_TILE = BN * 128;                       // one tap: BN rows x 32 floats
    static constexpr int STAGE = A_STAGE + TMA_MAX_KH * B_TILE;
    static constexpr int TMEM_COLS = MT * BN < 32 ? 32 : MT * BN;
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE + 256;
};

struct TmaGeom {
    int tw_log2;           // TW = 1 << tw_log2 (3 or 4), TH = 128 >> tw_log2
    int tiles_x, tiles_y;  // tiles per image
};

template <int BN, int MT, int STAGES>
__global__ void __launch_bounds__(TMA_THREADS)
conv_tma_kernel(ScsfmConv p, TcView v, TmaGeom g, const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap) {
    using Cfg = TmaCfg<BN, MT, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * Cfg::STAGE);
    uint64_t* bar_empty = bar_full + STAGES;
    uint64_t* bar_acc = bar_empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TW = 1 << g.tw_log2, TH = TBM >> g.tw_log2;
    const int N = p.Cout, n0 = blockIdx.y * BN;
    // tile -> (image, y0, x0)
    int t = blockIdx.x;
    const int tx = t % g.tiles_x; t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int b = t / g.tiles_y;
    const int y0 = ty * (MT * TH), x0 = tx * TW;
    const int chunks = (p.Cin + TBK - 1) / TBK;
    const int NIT = chunks * v.kw;                      // pipeline iterations: (channel chunk, dx)
    const int patch_rows = MT * TH + v.kh - 1;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(bar_full + s, 1);              // the producer's expect_tx arrival; TMA completes the bytes
            tc::mbar_init(bar_empty + s, 1);
        }
        tc::mbar_init(bar_acc, 1);
        tc::fence_barrier_init();
    }
    if (warp == TMA_EWARPS + 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = tc::smem_u32(smem);

    if (warp == TMA_EWARPS) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            tc::tma_prefetch_desc(&amap);
            tc::tma_prefetch_desc(&wmap);
            const uint32_t tx_bytes = (uint32_t)(patch_rows * TW * 128 + v.kh * Cfg::B_TILE);
            int it = 0;
            for (int ck = 0; ck < chunks; ++ck) {
                for (int dx = 0; dx < v.kw; ++dx, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    tc::mbar_wait(bar_empty + s, ph ^ 1);
                    const uint32_t st = smem_base + (uint32_t)(s * Cfg::STAGE);
                    tc::mbar_arrive_expect_tx(bar_full + s, tx_bytes);
                    tc::tma_load_4d(st, &amap, ck * TBK, x0 + v.ox0 + dx, y0 + v.oy0, b, bar_full + s);
                    for (int dy = 0; dy < v.kh; ++dy)
                        tc::tma_load_2d(st + (uint32_t)(Cfg::A_STAGE + dy * Cfg::B_TILE), &wmap, (dy * v.kw + dx) * p.Cin + ck * TBK, n0,
                                        bar_full + s);
                }
            }
        }
        __syncwarp();
    } else if (warp == TMA_EWARPS + 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = tc::make_idesc_tf32(TBM, BN, 0, 0);
        if (lane == 0) {
            const uint32_t dy_bytes = (uint32_t)(TW * 128);
            for (int it = 0; it < NIT; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                tc::mbar_wait(bar_full + s, ph);
                tc::fence_after_thread_sync();
                const uint32_t a_addr = smem_base + (uint32_t)(s * Cfg::STAGE);
                const uint32_t b_addr = a_addr + Cfg::A_STAGE;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    for (int dy = 0; dy < v.kh; ++dy) {
                        const uint32_t a_tap = a_addr + (uint32_t)(mt * TH + dy) * dy_bytes;
                        const uint32_t b_tap = b_addr + (uint32_t)(dy * Cfg::B_TILE);
#pragma unroll
                        for (int j = 0; j < TBK / 8; ++j) {
                            const uint64_t da = tc::make_smem_desc(a_tap + j * 32, 16, 1024, tc::LAYOUT_SW128);
                            const uint64_t db = tc::make_smem_desc(b_tap + j * 32, 16, 1024, tc::LAYOUT_SW128);
                            tc::mma_tf32(tmem_base + (uint32_t)(mt * BN), da, db, idesc, (it | dy | j) != 0 ? 1u : 0u);
                        }
                    }
                }
                tc::mma_commit(bar_empty + s);
            }
            tc::mma_commit(bar_acc);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0-7)
        tc::mbar_wait(bar_acc, 0);
        tc::fence_after_thread_sync();
        const int quarter = warp & 3, half = warp >> 2;
        const int l = quarter * 32 + lane;                  // TMEM lane = pixel index inside the 128-pixel sub-tile
        const int lr = l >> g.tw_log2, lc = l & (TW - 1);
        float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 33);      // all MMAs retired: operand smem is free
        constexpr int CW = BN < 32 ? BN : 32;
        constexpr int NCH = BN / CW;
        const int groups = p.bn_groups > 0 ? p.bn_groups : 1;
        const int grp = b / (p.B / groups);
#pragma unroll 1
        for (int idx = half; idx < MT * NCH; idx += TMA_EWARPS / 4) {
            const int mt = idx / NCH, cc = idx - mt * NCH;
            if (n0 + cc * CW >= N) continue;                // columns beyond Cout (N tile wider than the layer)
            const int ho = y0 + mt * TH + lr, wo = x0 + lc;
            const bool row_ok = ho < p.Ho && wo < p.Wo;
            const size_t out_row = ((size_t)b * v.out_H + (ho * v.out_sy + v.out_oy)) * v.out_W + (wo * v.out_sx + v.out_ox);
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * BN + cc * CW);
            if (CW == 32) tc::tmem_ld32(taddr, r);
            else tc::tmem_ld16(taddr, r);
            tc::tmem_ld_wait();
            float o[32];
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const int n = n0 + cc * CW + j;
                float x = __uint_as_float(r[j]);
                if (row_ok && n < N) {
                    if (p.bias) x += __ldg(p.bias + n);
                    if (p.addend) x += __ldg(p.addend + out_row * N + n);
                    x = tc_act(x, p.act);
                    if (p.act & ROUND_TF32) x = tf32_round(x);
                } else {
                    x = 0.f;
                }
                o[j] = x;
            }
            if (row_ok) {
                float* dst = p.out + out_row * N + n0 + cc * CW;
                if ((N & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < CW; j += 4)
                        if (n0 + cc * CW + j < N) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < CW; ++j)
                        if (n0 + cc * CW + j < N) dst[j] = o[j];
                }
            }
            if (p.bn_sums != nullptr) {
                // the whole tile lies in one image, hence in one BatchNorm group: column sums over the warp's 32 rows
                // (masked rows hold zeros) in fp64, one atomic pair per column into one of SCSFM_BN_SLOTS replicas
#pragma unroll
                for (int j = 0; j < CW; ++j) stage[lane * 33 + j] = o[j];
                __syncwarp();
                if (lane < CW) {
                    const int n = n0 + cc * CW + lane;
                    double s1 = 0.0, s2 = 0.0;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) {
                        const double tv = (double)stage[rr * 33 + lane];
                        s1 += tv;
                        s2 += tv * tv;
                    }
                    if (n < N && (s1 != 0.0 || s2 != 0.0)) {
                        double* d = p.bn_sums + (((size_t)(blockIdx.x % SCSFM_BN_SLOTS) * groups + grp) * N + n) * 2;
                        atomicAdd(d, s1);
                        atomicAdd(d + 1, s2);
                    }
                }
                __syncwarp();
            }
        }
    }

    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == TMA_EWARPS + 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int env_enable() {                     // SCSFM_CONV_TMA=0 routes every layer through the gather kernel
    const char* e = getenv("SCSFM_CONV_TMA");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
}
static int g_tma_enable = env_enable(), g_force_mt = 0, g_force_bn = 0, g_force_tw = 0;

bool conv_tma_eligible(const ScsfmConv& p, const TcView& v) {
    if (!g_tma_enable || v.border) return false;
    if (v.in_stride != 1 || v.kh > TMA_MAX_KH || v.kw > TMA_MAX_KH || v.kh < 1 || v.kw < 1) return false;
    if ((p.Cin & 3) != 0) return false;
    // measured (tools/check_conv_tma.py, round 1): this non-persistent version only wins where the K loop is long
    // enough to amortise the per-CTA prologue / epilogue (deep layers); forced configurations bypass the rule
    if (!g_force_mt && !g_force_bn && !g_force_tw && ((p.Cin + TBK - 1) / TBK) * v.kw < 24) return false;
    if (p.bn_sums && p.B % (p.bn_groups > 0 ? p.bn_groups : 1) != 0) return false;
    return true;
}

template <int BN, int MT, int STAGES>
static int launch_tma_cfg(const ScsfmConv& p, const TcView& v, int tw_log2, cudaStream_t st) {
    using Cfg = TmaCfg<BN, MT, STAGES>;
    static bool configured = false;
    if (!configured) {
        SCSFM_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<BN, MT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured = true;
    }
    const int TW = 1 << tw_log2, TH = TBM >> tw_log2;
    TmaGeom g{tw_log2, (p.Wo + TW - 1) / TW, (p.Ho + MT * TH - 1) / (MT * TH)};
    CUtensorMap amap, wmap;
    {
        // activations [B][Hi][Wi][Cin] (Cin contiguous); box = 32 channels x TW x (MT*TH + kh - 1) x 1, 128B swizzle
        const cuuint64_t gdim[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.B};
        const cuuint64_t gstride[3] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Wi * p.Cin * 4, (cuuint64_t)p.Hi * p.Wi * p.Cin * 4};
        const cuuint32_t box[4] = {(cuuint32_t)TBK, (cuuint32_t)TW, (cuuint32_t)(MT * TH + v.kh - 1), 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = encode_tiled(&amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.in), gdim, gstride, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(activations %d x %d x %d x %d) failed with CUresult %d", p.B, p.Hi, p.Wi, p.Cin, (int)r);
            return SCSFM_ERR_CUDA;
        }
    }
    {
        const int K = v.kh * v.kw * p.Cin;
        const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)p.Cout};
        const cuuint64_t gstride[1] = {(cuuint64_t)K * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)BN};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode_tiled(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.w), gdim, gstride, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(weights %d x %d) failed with CUresult %d", p.Cout, K, (int)r);
            return SCSFM_ERR_CUDA;
        }
    }
    dim3 grid(g.tiles_x * g.tiles_y * p.B, (p.Cout + BN - 1) / BN);
    conv_tma_kernel<BN, MT, STAGES><<<grid, TMA_THREADS, Cfg::SMEM, st>>>(p, v, g, amap, wmap);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}

static int tile_waste(int H, int W, int th, int tw) {      // padded area of the tiling
    return ((H + th - 1) / th * th) * ((W + tw - 1) / tw * tw);
}

int launch_conv_tma(const ScsfmConv& p, const TcView& v, cudaStream_t st) {
    const int N = p.Cout;
    // N tile
    // N tile: the smallest tile that covers Cout in one pass (rows beyond Cout are zero-filled by the TMA and masked in
    // the epilogue) -- re-reading the activation patch per N tile costs more than the idle MMA columns
    int bn;
    if (N <= 16) bn = 16;
    else if (N <= 32) bn = 32;
    else if (N <= 64) bn = 64;
    else bn = 128;
    // M tile: 1 or 2 stacked 128-pixel sub-tiles, TW = 8 or 16 -- least padded area first, then the larger tile
    int best_mt = 1, best_tw = 4;
    long best_cost = -1;
    for (int mt = 1; mt <= 2; ++mt)
        for (int twl = 3; twl <= 4; ++twl) {
            const int tw = 1 << twl, th = mt * (TBM >> twl);
            const long area = tile_waste(p.Ho, p.Wo, th, tw);
            const long ctas = (long)(area / (th * tw)) * p.B * ((N + bn - 1) / bn);
            if (mt == 2 && ctas < 2 * 148) continue;          // keep every SM busy before growing the tile
            // cost: padded area, inflated by the halo rows re-read per tile (kh - 1 extra rows per th)
            const long cost = area * (th + v.kh - 1) / th;
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_mt = mt; best_tw = twl; }
        }
    if (bn == 128 && N % 128 == 0 &&
        (long)(tile_waste(p.Ho, p.Wo, best_mt * (TBM >> best_tw), 1 << best_tw) / (best_mt * TBM)) * p.B * (N / 128) < 148)
        bn = 64;                                            // deep layers at 8x26 / 16x52: more CTAs
    if (g_force_mt) best_mt = g_force_mt;
    if (g_force_tw) best_tw = g_force_tw;
    if (g_force_bn) bn = g_force_bn;
    if (best_mt == 1) {
        switch (bn) {
            case 16: return launch_tma_cfg<16, 1, 4>(p, v, best_tw, st);
            case 32: return launch_tma_cfg<32, 1, 3>(p, v, best_tw, st);
            case 64: return launch_tma_cfg<64, 1, 2>(p, v, best_tw, st);
            default: return launch_tma_cfg<128, 1, 3>(p, v, best_tw, st);
        }
    }
    switch (bn) {
        case 16: return launch_tma_cfg<16, 2, 3>(p, v, best_tw, st);
        case 32: return launch_tma_cfg<32, 2, 3>(p, v, best_tw, st);
        case 64: return launch_tma_cfg<64, 2, 3>(p, v, best_tw, st);
        default: return launch_tma_cfg<128, 2, 2>(p, v, best_tw, st);
    }
}

}  // namespace scsfm

// Experiment / test hook: enable = 0 routes everything through the cp.async kernel; force_* = 0 keeps the heuristic
// (force_mt in {1,2}, force_bn in {16,32,64,128}, force_tw_log2 in {3,4}).
extern "C" int scsfm_conv_tma_config(int enable, int force_mt, int force_bn, int force_tw_log2) {
    SCSFM_CHECK_ARG((force_mt >= 0 && force_mt <= 2) && (force_bn == 0 || force_bn == 16 || force_bn == 32 || force_bn == 64 || force_bn == 128) &&
                        (force_tw_log2 == 0 || force_tw_log2 == 3 || force_tw_log2 == 4),
                    "conv_tma_config: bad arguments");
    scsfm::g_tma_enable = enable;
    scsfm::g_force_mt = force_mt;
    scsfm::g_force_bn = force_bn;
    scsfm::g_force_tw = force_tw_log2;
    return SCSFM_OK;
}
