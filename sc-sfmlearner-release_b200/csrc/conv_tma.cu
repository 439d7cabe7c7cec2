// Stride-1 convolution on the tensor cores: persistent CTAs, TMA halo-patch producer, double-buffered TMEM accumulators
// (tcgen05 kind::tf32, fp32 accumulation).
//
// The cp.async kernel of conv_tc.cu gathers one 128 x 32-float im2col slice per (tap, channel chunk): every input
// pixel crosses L2 -> shared memory kh*kw times, 256 threads spend their issue slots on 16-byte copies, and every
// 128-pixel tile pays a CTA launch, a TMEM allocation and a non-overlapped epilogue.  Here:
//
//  * the output tile is a 2-D patch of ONE image (MT*TH rows x TW columns, TH*TW = 128, TW in {8, 16}); per (channel
//    chunk, dx) ONE 4-D tiled TMA load brings the (MT*TH + kh - 1) x TW x 32-channel input patch, already in the
//    128B-swizzled K-major layout the UMMA descriptors expect.  TW is a multiple of 8, so the im2col operand of tap row
//    dy is the SAME patch shifted by dy*TW rows = dy*TW*128 bytes (a multiple of the 1024-byte swizzle atom): the kh
//    vertical taps reuse one load and the input crosses L2 -> smem kw*(1 + (kh-1)/(MT*TH)) times instead of kh*kw.
//    Zero padding is the TMA's out-of-bounds fill; channels beyond Cin in the last 32-wide chunk are zero-filled the
//    same way (the matching weight columns then multiply zeros) and whole all-zero K8 slices are not issued at all.
//  * one CTA per SM walks the (tile, N tile) work list; the shared-memory stage ring and the mbarrier phases run across
//    tiles, so the producer prefetches the next tile's patches while the current one is still being multiplied;
//  * two TMEM accumulator buffers: the epilogue warps drain tile j (tcgen05.ld -> bias / residual addend / activation /
//    TF32 rounding / BatchNorm sums -> global) while the MMA warp already accumulates tile j+1.
//
//   warps 0-7   epilogue
//   warp 8      lane 0: TMA producer (1 activation box + kh weight boxes per stage, mbarrier expect_tx)
//   warp 9      TMEM allocation; lane 0: MMA issuer (MT * kh * <=4 tcgen05.mma M128 x BN x K8 per stage)
//
// Used for: forward of every stride-1 layer with kh, kw <= 3 (reflection-padded layers run it with zero padding and
// the cp.async kernel then recomputes the 2*(H+W)-4 border pixels per image, see tc_dispatch in conv_tc.cu), stride-1
// data gradients, and the four parity-class sub-convolutions of stride-2 data gradients.
#include <stdlib.h>

#include "conv_tc.cuh"

namespace scsfm {

constexpr int TMA_EWARPS = 8;
constexpr int TMA_THREADS = (TMA_EWARPS + 2) * 32;
constexpr int TMA_MAX_KH = 3;
constexpr int TMA_MAX_STAGES = 8;
constexpr int TMA_EPI_BYTES = TMA_EWARPS * 32 * 33 * 4;      // BatchNorm column-sum staging, one 32 x 33 pad per warp
constexpr int TMA_SMEM_MAX = 232448;                         // 227 KB: the most one CTA may opt into on sm_100

struct TmaGeom {
    int tw_log2;             // TW = 1 << tw_log2 (3 or 4), TH = 128 >> tw_log2
    int tiles_x, tiles_y;    // M tiles per image
    int n_tiles, num_work;   // N tiles; work items = B * tiles_y * tiles_x * n_tiles (N tile fastest)
    int stages;              // shared-memory ring depth
    int a_bytes, stage_bytes;
    unsigned long long* dbg; // optional per-CTA cycle counters (scsfm_conv_tma_debug), 8 per CTA; NULL = off
};

__device__ __forceinline__ long long tma_clock() { return clock64(); }

template <int BN, int MT>
__global__ void __launch_bounds__(TMA_THREADS, 1)
conv_tma_kernel(ScsfmConv p, TcView v, TmaGeom g, const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap) {
    constexpr int ACC_COLS = MT * BN;                        // TMEM columns of one accumulator buffer
    constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
    constexpr int B_TILE = BN * 128;                          // one tap: BN rows x 32 floats
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + (size_t)g.stages * g.stage_bytes;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(epi_smem + (p.bn_sums != nullptr ? TMA_EPI_BYTES : 0));
    uint64_t* bar_empty = bar_full + TMA_MAX_STAGES;
    uint64_t* acc_full = bar_empty + TMA_MAX_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TW = 1 << g.tw_log2, TH = TBM >> g.tw_log2;
    const int N = p.Cout;
    const int chunks = (p.Cin + TBK - 1) / TBK;
    const int patch_rows = MT * TH + v.kh - 1;

    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            tc::mbar_init(bar_full + s, 1);              // the producer's expect_tx arrival; TMA completes the bytes
            tc::mbar_init(bar_empty + s, 1);             // tcgen05.commit of the MMAs that read the stage
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(acc_full + i, 1);              // tcgen05.commit after the tile's last MMA
            tc::mbar_init(acc_empty + i, TMA_EWARPS);    // one arrival per epilogue warp
        }
        tc::fence_barrier_init();
    }
    if (warp == TMA_EWARPS + 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
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
            const uint32_t tx_bytes = (uint32_t)(patch_rows * TW * 128 + v.kh * B_TILE);
            int s = 0;
            uint32_t ph = 0;
            long long t_wait = 0;
            const long long t_begin = tma_clock();
            for (int w = blockIdx.x; w < g.num_work; w += gridDim.x) {
                int t = w / g.n_tiles;
                const int n0 = (w - t * g.n_tiles) * BN;
                const int tx = t % g.tiles_x; t /= g.tiles_x;
                const int ty = t % g.tiles_y;
                const int b = t / g.tiles_y;
                const int y0 = ty * (MT * TH) + v.oy0, x0 = tx * TW + v.ox0;
                for (int ck = 0; ck < chunks; ++ck) {
                    for (int dx = 0; dx < v.kw; ++dx) {
                        const long long t0 = g.dbg ? tma_clock() : 0;
                        tc::mbar_wait(bar_empty + s, ph ^ 1);
                        if (g.dbg) t_wait += tma_clock() - t0;
                        const uint32_t st = smem_base + (uint32_t)(s * g.stage_bytes);
                        tc::mbar_arrive_expect_tx(bar_full + s, tx_bytes);
                        tc::tma_load_4d(st, &amap, ck * TBK, x0 + dx, y0, b, bar_full + s);
                        for (int dy = 0; dy < v.kh; ++dy)
                            tc::tma_load_2d(st + (uint32_t)(g.a_bytes + dy * B_TILE), &wmap, (dy * v.kw + dx) * p.Cin + ck * TBK, n0,
                                            bar_full + s);
                        if (++s == g.stages) { s = 0; ph ^= 1; }
                    }
                }
            }
            if (g.dbg) {
                g.dbg[blockIdx.x * 8 + 0] = (unsigned long long)t_wait;
                g.dbg[blockIdx.x * 8 + 1] = (unsigned long long)(tma_clock() - t_begin);
            }
        }
        __syncwarp();
    } else if (warp == TMA_EWARPS + 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = tc::make_idesc_tf32(TBM, BN, 0, 0);
        if (lane == 0) {
            const uint32_t dy_bytes = (uint32_t)(TW * 128);
            // descriptor = constant fields (LBO 16, SBO 1024, 128B swizzle) + (shared address >> 4) in the low 14 bits
            const uint64_t desc0 = tc::make_smem_desc(0, 16, 1024, tc::LAYOUT_SW128);
            int s = 0;
            uint32_t ph = 0;
            int j = 0;
            long long t_full = 0, t_acc = 0;
            const long long t_begin = tma_clock();
            for (int w = blockIdx.x; w < g.num_work; w += gridDim.x, ++j) {
                const int buf = j & 1;
                const long long ta = g.dbg ? tma_clock() : 0;
                tc::mbar_wait(acc_empty + buf, ((j >> 1) & 1) ^ 1);      // epilogue of tile j-2 has drained this buffer
                if (g.dbg) t_acc += tma_clock() - ta;
                tc::fence_after_thread_sync();
                const uint32_t acc = tmem_base + (uint32_t)(buf * ACC_COLS);
                for (int ck = 0; ck < chunks; ++ck) {
                    const int rem = p.Cin - ck * TBK;
                    const int k8 = rem >= TBK ? TBK / 8 : (rem + 7) / 8;           // K8 slices holding real channels
                    for (int dx = 0; dx < v.kw; ++dx) {
                        const long long t0 = g.dbg ? tma_clock() : 0;
                        tc::mbar_wait(bar_full + s, ph);
                        if (g.dbg) t_full += tma_clock() - t0;
                        tc::fence_after_thread_sync();
                        const uint32_t a_addr = smem_base + (uint32_t)(s * g.stage_bytes);
                        const uint32_t b_addr = a_addr + (uint32_t)g.a_bytes;
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            for (int dy = 0; dy < v.kh; ++dy) {
                                const uint32_t a_tap = a_addr + (uint32_t)(mt * TH + dy) * dy_bytes;
                                const uint32_t b_tap = b_addr + (uint32_t)(dy * B_TILE);
                                uint64_t da = desc0 + (uint64_t)(a_tap >> 4), db = desc0 + (uint64_t)(b_tap >> 4);
                                for (int q = 0; q < k8; ++q) {
                                    tc::mma_tf32(acc + (uint32_t)(mt * BN), da, db, idesc, (ck | dx | dy | q) != 0 ? 1u : 0u);
                                    da += 2;                 // next K8 slice: +32 bytes inside the 128-byte swizzle row
                                    db += 2;
                                }
                            }
                        }
                        tc::mma_commit(bar_empty + s);
                        if (++s == g.stages) { s = 0; ph ^= 1; }
                    }
                }
                tc::mma_commit(acc_full + buf);
            }
            if (g.dbg) {
                g.dbg[blockIdx.x * 8 + 2] = (unsigned long long)t_full;
                g.dbg[blockIdx.x * 8 + 3] = (unsigned long long)t_acc;
                g.dbg[blockIdx.x * 8 + 4] = (unsigned long long)(tma_clock() - t_begin);
                g.dbg[blockIdx.x * 8 + 7] = (unsigned long long)j;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0-7)
        const int quarter = warp & 3, half = warp >> 2;
        const int l = quarter * 32 + lane;                  // TMEM lane = pixel index inside the 128-pixel sub-tile
        const int lr = l >> g.tw_log2, lc = l & (TW - 1);
        float* stage = reinterpret_cast<float*>(epi_smem) + warp * (32 * 33);
        constexpr int CW = BN < 32 ? BN : 32;
        constexpr int NCH = BN / CW;
        const int groups = p.bn_groups > 0 ? p.bn_groups : 1;
        int j = 0;
        long long t_wait = 0;
        const long long t_begin = tma_clock();
        for (int w = blockIdx.x; w < g.num_work; w += gridDim.x, ++j) {
            int t = w / g.n_tiles;
            const int n0 = (w - t * g.n_tiles) * BN;
            const int tx = t % g.tiles_x; t /= g.tiles_x;
            const int ty = t % g.tiles_y;
            const int b = t / g.tiles_y;
            const int y0 = ty * (MT * TH), x0 = tx * TW;
            const int grp = b / (p.B / groups);
            const int buf = j & 1;
            const long long t0 = g.dbg ? tma_clock() : 0;
            tc::mbar_wait(acc_full + buf, (j >> 1) & 1);
            if (g.dbg) t_wait += tma_clock() - t0;
            tc::fence_after_thread_sync();
            const uint32_t acc = tmem_base + (uint32_t)(buf * ACC_COLS) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int idx = half; idx < MT * NCH; idx += TMA_EWARPS / 4) {
                const int mt = idx / NCH, cc = idx - mt * NCH;
                if (n0 + cc * CW >= N) continue;                // columns beyond Cout (N tile wider than the layer)
                const int ho = y0 + mt * TH + lr, wo = x0 + lc;
                const bool row_ok = ho < p.Ho && wo < p.Wo;
                const size_t out_row = ((size_t)b * v.out_H + (ho * v.out_sy + v.out_oy)) * v.out_W + (wo * v.out_sx + v.out_ox);
                uint32_t r[32];
                if (CW == 32) tc::tmem_ld32(acc + (uint32_t)(mt * BN + cc * CW), r);
                else tc::tmem_ld16(acc + (uint32_t)(mt * BN + cc * CW), r);
                tc::tmem_ld_wait();
                float o[32];
#pragma unroll
                for (int q = 0; q < CW; ++q) {
                    const int n = n0 + cc * CW + q;
                    float x = __uint_as_float(r[q]);
                    if (row_ok && n < N) {
                        if (p.bias) x += __ldg(p.bias + n);
                        if (p.addend) x += __ldg(p.addend + out_row * N + n);
                        x = tc_act(x, p.act);
                        if (p.act & ROUND_TF32) x = tf32_round(x);
                    } else {
                        x = 0.f;
                    }
                    o[q] = x;
                }
                if (row_ok) {
                    float* dst = p.out + out_row * N + n0 + cc * CW;
                    if ((N & 3) == 0) {
#pragma unroll
                        for (int q = 0; q < CW; q += 4)
                            if (n0 + cc * CW + q < N) *reinterpret_cast<float4*>(dst + q) = make_float4(o[q], o[q + 1], o[q + 2], o[q + 3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < CW; ++q)
                            if (n0 + cc * CW + q < N) dst[q] = o[q];
                    }
                }
                if (p.bn_sums != nullptr) {
                    // the whole tile lies in one image, hence in one BatchNorm group: column sums over the warp's 32 rows
                    // (masked rows hold zeros) in fp64, one atomic pair per column into one of SCSFM_BN_SLOTS replicas
#pragma unroll
                    for (int q = 0; q < CW; ++q) stage[lane * 33 + q] = o[q];
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
                            double* d = p.bn_sums + (((size_t)(w % SCSFM_BN_SLOTS) * groups + grp) * N + n) * 2;
                            atomicAdd(d, s1);
                            atomicAdd(d + 1, s2);
                        }
                    }
                    __syncwarp();
                }
            }
            // all of this warp's tcgen05.ld of the buffer have completed (tmem_ld_wait above): hand it back to the MMA warp
            tc::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_empty + buf);
        }
        if (g.dbg && tid == 0) {
            g.dbg[blockIdx.x * 8 + 5] = (unsigned long long)t_wait;
            g.dbg[blockIdx.x * 8 + 6] = (unsigned long long)(tma_clock() - t_begin);
        }
    }

    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == TMA_EWARPS + 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int env_enable() {                     // SCSFM_CONV_TMA=0 routes every layer through the gather kernel
    const char* e = getenv("SCSFM_CONV_TMA");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
}
static int g_tma_enable = env_enable(), g_force_mt = 0, g_force_bn = 0, g_force_tw = 0;
static unsigned long long* g_dbg = nullptr;

bool conv_tma_eligible(const ScsfmConv& p, const TcView& v) {
    if (!g_tma_enable || v.border) return false;
    if (v.in_stride != 1 || v.kh > TMA_MAX_KH || v.kw > TMA_MAX_KH || v.kh < 1 || v.kw < 1) return false;
    if ((p.Cin & 3) != 0) return false;
    if (p.bn_sums && p.B % (p.bn_groups > 0 ? p.bn_groups : 1) != 0) return false;
    return true;
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BN, int MT>
static int launch_tma_cfg(const ScsfmConv& p, const TcView& v, int tw_log2, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        SCSFM_CHECK_CUDA(cudaFuncSetAttribute(conv_tma_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_MAX));
        configured = true;
    }
    const int TW = 1 << tw_log2, TH = TBM >> tw_log2;
    TmaGeom g;
    g.tw_log2 = tw_log2;
    g.tiles_x = (p.Wo + TW - 1) / TW;
    g.tiles_y = (p.Ho + MT * TH - 1) / (MT * TH);
    g.n_tiles = (p.Cout + BN - 1) / BN;
    g.num_work = g.tiles_x * g.tiles_y * p.B * g.n_tiles;
    g.a_bytes = ((MT * TH + v.kh - 1) * TW * 128 + 1023) / 1024 * 1024;
    g.stage_bytes = g.a_bytes + v.kh * BN * 128;
    const int fixed = 1024 /* alignment slack */ + (p.bn_sums ? TMA_EPI_BYTES : 0) + 256 /* barriers */;
    g.stages = (TMA_SMEM_MAX - fixed) / g.stage_bytes;
    if (g.stages > TMA_MAX_STAGES) g.stages = TMA_MAX_STAGES;
    if (g.stages < 2) {
        set_error("conv_tma: stage of %d bytes does not fit twice in shared memory", g.stage_bytes);
        return SCSFM_ERR_ARG;
    }
    g.dbg = g_dbg;
    const size_t smem = (size_t)fixed + (size_t)g.stages * g.stage_bytes;
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
    int ctas = sm_count();
    if (ctas > g.num_work) ctas = g.num_work;
    conv_tma_kernel<BN, MT><<<ctas, TMA_THREADS, smem, st>>>(p, v, g, amap, wmap);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}

int launch_conv_tma(const ScsfmConv& p, const TcView& v, cudaStream_t st) {
    const int N = p.Cout;
    // Tile choice.  Candidates: N tile = the smallest of 16/32/64/128 covering Cout (rows beyond Cout are zero-filled by
    // the TMA and masked in the epilogue; for Cout > 64 also 64), 1 or 2 stacked 128-pixel sub-tiles, TW = 8 or 16.
    // Cost = L2 -> smem traffic per (channel chunk, dx) in units of 128 B: padded area x N tiles x (activation patch incl.
    // halo rows + kh weight rows shared by MT*128 pixels), divided by the fraction of the persistent CTAs' waves that has work.
    int bn_lo;
    if (N <= 16) bn_lo = 16;
    else if (N <= 32) bn_lo = 32;
    else if (N <= 64) bn_lo = 64;
    else bn_lo = 128;
    const int nsm = sm_count();
    int bn = bn_lo, best_mt = 1, best_tw = 4;
    double best_cost = -1.0;
    const int cands[2] = {bn_lo, 64};
    for (int ci = 0; ci < (bn_lo > 64 ? 2 : 1); ++ci) {
        const int cand = cands[ci];
        for (int mt = 1; mt <= 2; ++mt)
            for (int twl = 3; twl <= 4; ++twl) {
                const int tw = 1 << twl, th = mt * (TBM >> twl);
                const long ty = (p.Ho + th - 1) / th, tx = (p.Wo + tw - 1) / tw, nt = (N + cand - 1) / cand;
                const long work = ty * tx * p.B * nt;
                const long waves = (work + nsm - 1) / nsm;
                const double eff = (double)work / (double)(waves * nsm);
                const double per_px = (double)(th + v.kh - 1) / th + (double)(v.kh * cand) / (double)(mt * TBM);
                const double cost = (double)(ty * th) * (double)(tx * tw) * (double)nt * per_px / eff;
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; bn = cand; best_mt = mt; best_tw = twl; }
            }
    }
    if (g_force_mt) best_mt = g_force_mt;
    if (g_force_tw) best_tw = g_force_tw;
    if (g_force_bn) bn = g_force_bn;
    if (best_mt == 1) {
        switch (bn) {
            case 16: return launch_tma_cfg<16, 1>(p, v, best_tw, st);
            case 32: return launch_tma_cfg<32, 1>(p, v, best_tw, st);
            case 64: return launch_tma_cfg<64, 1>(p, v, best_tw, st);
            default: return launch_tma_cfg<128, 1>(p, v, best_tw, st);
        }
    }
    switch (bn) {
        case 16: return launch_tma_cfg<16, 2>(p, v, best_tw, st);
        case 32: return launch_tma_cfg<32, 2>(p, v, best_tw, st);
        case 64: return launch_tma_cfg<64, 2>(p, v, best_tw, st);
        default: return launch_tma_cfg<128, 2>(p, v, best_tw, st);
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

// Profiling hook: buf = device array of 8 x (number of SMs) 64-bit counters, or NULL to switch it off.  Per CTA (clock64
// cycles): [0] producer waiting for a free stage, [1] producer total, [2] MMA thread waiting for operands, [3] MMA thread
// waiting for a drained accumulator, [4] MMA thread total, [5] epilogue warp 0 waiting for an accumulator, [6] epilogue
// total, [7] tiles processed.
extern "C" int scsfm_conv_tma_debug(unsigned long long* buf) {
    scsfm::g_dbg = buf;
    return SCSFM_OK;
}
