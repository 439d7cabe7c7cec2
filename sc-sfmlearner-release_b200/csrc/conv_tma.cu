// Stride-1 convolution on the tensor cores: persistent CTAs, TMA halo-patch producer, double-buffered TMEM accumulators
// (tcgen05 kind::tf32, fp32 accumulation), weights on the M side of the MMA.
//
//   D^T[Cout tile (M = 128 TMEM lanes), pixels (N = 128 or 256 TMEM columns)] = W[Cout, K] x im2col(x)[pixels, K]^T
//
// Why this shape.  Measured on B200 (tools/tma_profile.py, profiles/r01_tma_roles.txt): with both operands in shared
// memory a tcgen05.mma with M = 128 takes ~130-140 cycles whatever N is (the M-side operand streams from shared memory
// one 32-byte row per cycle), so an M128 x N64 x K8 instruction runs the tensor pipe at a quarter of its rate.  The
// layers here have 16..128 output channels (only the deepest have 256/512) but >= 10^4 pixels, so the PIXELS go on the
// N side (up to 256 per instruction) and the output channels on the M side (rows beyond Cout compute garbage lanes that
// nobody reads and cost nothing extra).  The accumulator comes out channel-major: TMEM lane = output channel, column =
// pixel, which also makes the epilogue stores coalesced along NHWC channels and the BatchNorm sums a per-lane loop.
//
// Versus the cp.async kernel of conv_tc.cu (one 128 x 32-float im2col slice gathered per (tap, channel chunk) by 256
// threads, one CTA per 128-pixel tile):
//  * the output tile is a 2-D patch of ONE image (MT*TH rows x TW columns, TH*TW = 128, TW in {8, 16}); per (channel
//    chunk, dx) ONE 4-D tiled TMA load brings the (MT*TH + kh - 1) x TW x 32-channel input patch in the 128B-swizzled
//    K-major layout the UMMA descriptors expect.  TW is a multiple of 8, so the operand of tap row dy is the SAME patch
//    shifted by dy*TW rows = dy*TW*128 bytes (a multiple of the 1024-byte swizzle atom): the kh vertical taps reuse one
//    load.  Zero padding is the TMA's out-of-bounds fill; channels beyond Cin in the last 32-wide chunk are zero-filled
//    the same way (the matching weight columns then multiply zeros) and all-zero K8 slices are not issued at all;
//  * one CTA per SM walks the (tile, Cout tile) work list; the shared-memory stage ring and the mbarrier phases run
//    across tiles, so the producer prefetches the next tile's patches while the current one is still being multiplied;
//  * two TMEM accumulator buffers: the epilogue warps drain tile j (tcgen05.ld -> bias / residual addend / activation /
//    TF32 rounding / BatchNorm sums -> global) while the MMA warp already accumulates tile j+1.
//
//   warps 0-7   epilogue (warp w reads TMEM lanes 32*(w%4).., i.e. channels; w/4 picks every other 32-pixel chunk)
//   warp 8      lane 0: TMA producer (1 activation box + kh weight boxes per stage, mbarrier expect_tx)
//   warp 9      TMEM allocation; lane 0: MMA issuer (kh * <=4 tcgen05.mma M128 x N(128|256) x K8 per stage)
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
constexpr int TMA_SMEM_MAX = 232448;                         // 227 KB: the most one CTA may opt into on sm_100

struct TmaGeom {
    int tw_log2;             // TW = 1 << tw_log2 (3 or 4), TH = 128 >> tw_log2
    int tiles_x, tiles_y;    // pixel tiles per image
    int n_tiles, num_work;   // Cout tiles; work items = B * tiles_y * tiles_x * n_tiles (Cout tile fastest)
    int stages;              // shared-memory ring depth
    int a_bytes, stage_bytes;
    // split-accumulate passes per (channel chunk, dx): pass i multiplies (activations: lo if a_lo bit i else raw) by
    // (weights: lo if w_lo bit i else raw); plain TF32 = one pass with both masks 0
    int npass, a_lo, w_lo;
    // split mode with <= 64 output channels: the weight tile stacks W (rows 0..) and lo(W) (rows 64..) on the M side of ONE
    // tcgen05.mma, so the two passes lo(x) and x produce all four products (TMEM lanes c and 64 + c are added in the epilogue):
    // two MMAs per K8 slice instead of three
    int stack;
    // accumulation chunks: the tensor core adds into the TMEM accumulator with truncation (measured: relative error ~ 3e-8 per
    // tcgen05.mma of the chain, a systematic bias), so a chain is cut after `cpg` channel chunks (~100 MMAs) and the epilogue
    // warps add the partial accumulators in registers (round-to-nearest).  cpg >= chunks: one chain per tile (plain TF32 mode).
    int cpg;
    unsigned long long* dbg; // optional per-CTA cycle counters (ScsfmConv.debug), 8 per CTA; NULL = off
};

__device__ __forceinline__ long long tma_clock() { return clock64(); }

// BNW: rows of the weight tile kept in shared memory (16/32/64/128 >= the Cout tile); the MMA always reads 128 rows from
// the tile start (M = 128), the rows past BNW are whatever follows in shared memory and only feed unread lanes.
// MT: 128-pixel sub-tiles stacked vertically, N = MT * 128 pixels per instruction.
template <int BNW, int MT>
__global__ void __launch_bounds__(TMA_THREADS, 1)
conv_tma_kernel(ScsfmConv p, TcView v, TmaGeom g, const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap,
                const __grid_constant__ CUtensorMap amap_lo, const __grid_constant__ CUtensorMap wmap_lo) {
    constexpr int NPIX = MT * TBM;                           // TMEM columns of one accumulator buffer
    constexpr int TMEM_COLS = 2 * NPIX;                      // 256 or 512
    const int W_TILE = g.stack ? TBM * 128 : BNW * 128;      // one tap: BNW rows x 32 floats (stacked: W at row 0, lo(W) at row 64)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // barriers live in FRONT of the stage ring: the MMA's 128-row read of a BNW-row weight tile may run past the last stage
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar_empty = bar_full + TMA_MAX_STAGES;
    uint64_t* acc_full = bar_empty + TMA_MAX_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint8_t* ring = smem + 1024;
    // epilogue transpose buffer T[64 pixels][BNW + 4] behind the ring and the over-read pad
    float* T = reinterpret_cast<float*>(ring + (size_t)g.stages * g.stage_bytes + (g.stack ? 0 : (TBM - BNW) * 128));

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
    const uint32_t ring_base = tc::smem_u32(ring);

    if (warp == TMA_EWARPS) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            tc::tma_prefetch_desc(&amap);
            tc::tma_prefetch_desc(&wmap);
            if (g.a_lo) tc::tma_prefetch_desc(&amap_lo);
            if (g.w_lo) tc::tma_prefetch_desc(&wmap_lo);
            const uint32_t tx_bytes = (uint32_t)(patch_rows * TW * 128 + v.kh * (g.stack ? 2 : 1) * BNW * 128);
            int s = 0;
            uint32_t ph = 0;
            long long t_wait = 0;
            const long long t_begin = tma_clock();
            for (int w = blockIdx.x; w < g.num_work; w += gridDim.x) {
                int t = w / g.n_tiles;
                const int n0 = (w - t * g.n_tiles) * TBM;
                const int tx = t % g.tiles_x; t /= g.tiles_x;
                const int ty = t % g.tiles_y;
                const int b = t / g.tiles_y;
                const int y0 = ty * (MT * TH) + v.oy0, x0 = tx * TW + v.ox0;
                // one accumulation chain = cpg channel chunks; inside a chain the low-part passes run FIRST (their sums are
                // 2^-11 of the main term: added while the accumulator is still small they lose nothing to its truncation),
                // the raw x raw pass last
                for (int ck0 = 0; ck0 < chunks; ck0 += g.cpg)
                for (int q = 0; q < g.npass; ++q) {
                    const int ps = (q + 1) % g.npass;
                    const CUtensorMap* am = ((g.a_lo >> ps) & 1) ? &amap_lo : &amap;
                    const CUtensorMap* wm = ((g.w_lo >> ps) & 1) ? &wmap_lo : &wmap;
                    for (int ck = ck0; ck < min(chunks, ck0 + g.cpg); ++ck) {
                        for (int dx = 0; dx < v.kw; ++dx) {
                            const long long t0 = g.dbg ? tma_clock() : 0;
                            tc::mbar_wait(bar_empty + s, ph ^ 1);
                            if (g.dbg) t_wait += tma_clock() - t0;
                            const uint32_t st = ring_base + (uint32_t)(s * g.stage_bytes);
                            tc::mbar_arrive_expect_tx(bar_full + s, tx_bytes);
                            tc::tma_load_4d(st, am, ck * TBK, x0 + dx, y0, b, bar_full + s);
                            for (int dy = 0; dy < v.kh; ++dy) {
                                const uint32_t wdst = st + (uint32_t)(g.a_bytes + dy * W_TILE);
                                const int kcol = (dy * v.kw + dx) * p.Cin + ck * TBK;
                                if (g.stack) {
                                    tc::tma_load_2d(wdst, &wmap, kcol, n0, bar_full + s);
                                    tc::tma_load_2d(wdst + 64 * 128, &wmap_lo, kcol, n0, bar_full + s);
                                } else {
                                    tc::tma_load_2d(wdst, wm, kcol, n0, bar_full + s);
                                }
                            }
                            if (++s == g.stages) { s = 0; ph ^= 1; }
                        }
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
        constexpr uint32_t idesc = tc::make_idesc_tf32(TBM, NPIX, 0, 0);      // M = 128 weight rows, N = NPIX pixels
        if (lane == 0) {
            const uint32_t dy_bytes = (uint32_t)(TW * 128);
            // descriptor = constant fields (LBO 16, SBO 1024, 128B swizzle) + (shared address >> 4) in the low 14 bits
            const uint64_t desc0 = tc::make_smem_desc(0, 16, 1024, tc::LAYOUT_SW128);
            int s = 0;
            uint32_t ph = 0;
            int j = 0;
            long long t_full = 0, t_acc = 0;
            const long long t_begin = tma_clock();
            int jc = 0;                                    // accumulation chunks issued so far: TMEM buffer jc & 1
            for (int w = blockIdx.x; w < g.num_work; w += gridDim.x, ++j) {
                for (int ck0 = 0; ck0 < chunks; ck0 += g.cpg) {          // one accumulation chain (same loop nest as the producer)
                    const int buf = jc & 1;
                    const long long ta = g.dbg ? tma_clock() : 0;
                    tc::mbar_wait(acc_empty + buf, ((jc >> 1) & 1) ^ 1);      // the epilogue has drained this buffer
                    if (g.dbg) t_acc += tma_clock() - ta;
                    tc::fence_after_thread_sync();
                    const uint32_t acc = tmem_base + (uint32_t)(buf * NPIX);
                for (int q = 0; q < g.npass; ++q) {
                    for (int ck = ck0; ck < min(chunks, ck0 + g.cpg); ++ck) {
                    const bool first_of_chain = q == 0 && ck == ck0;
                    const int rem = p.Cin - ck * TBK;
                    const int k8 = rem >= TBK ? TBK / 8 : (rem + 7) / 8;           // K8 slices holding real channels
                        for (int dx = 0; dx < v.kw; ++dx) {
                            const long long t0 = g.dbg ? tma_clock() : 0;
                            tc::mbar_wait(bar_full + s, ph);
                            if (g.dbg) t_full += tma_clock() - t0;
                            tc::fence_after_thread_sync();
                            const uint32_t x_addr = ring_base + (uint32_t)(s * g.stage_bytes);
                            const uint32_t w_addr = x_addr + (uint32_t)g.a_bytes;
                            for (int dy = 0; dy < v.kh; ++dy) {
                                uint64_t dw = desc0 + (uint64_t)((w_addr + (uint32_t)(dy * W_TILE)) >> 4);      // M side: weights
                                uint64_t dx_ = desc0 + (uint64_t)((x_addr + (uint32_t)dy * dy_bytes) >> 4);    // N side: pixels
                                for (int q8 = 0; q8 < k8; ++q8) {
                                    tc::mma_tf32(acc, dw, dx_, idesc, (!first_of_chain || (dx | dy | q8) != 0) ? 1u : 0u);
                                    dw += 2;                     // next K8 slice: +32 bytes inside the 128-byte swizzle row
                                    dx_ += 2;
                                }
                            }
                            tc::mma_commit(bar_empty + s);
                            if (++s == g.stages) { s = 0; ph ^= 1; }
                        }
                    }
                }
                    tc::mma_commit(acc_full + buf);          // end of the chain: hand the buffer to the epilogue warps
                    ++jc;
                }
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
        // ------------------------------------------------------------------ epilogue (warps 0-7, 256 threads)
        // The accumulator is channel-major (TMEM lane = output channel, column = pixel) but the tensor is NHWC, and a thin
        // layer occupies only 16 or 32 of the 128 lanes.  So the tile is transposed through shared memory in rounds of 64
        // pixels: the warps whose lane quarter holds real channels copy tcgen05.ld results into T[pixel][channel], then
        // ALL 256 threads walk T as float4 channel groups: bias / residual addend / activation / TF32 rounding, 16-byte
        // coalesced stores, BatchNorm partial sums per thread (a thread always owns the same 4 channels).
        constexpr int CT = BNW;                      // channels per T row
        constexpr int TS = CT + 4;                   // row stride in floats (keeps rows 16-byte aligned)
        constexpr int GPR = CT / 4;                  // float4 groups per pixel
        constexpr int ITER = (64 * GPR) / 256;       // groups per thread per round (CT / 16)
        constexpr int PXSTEP = 256 / GPR;            // pixel distance between a thread's groups
        const int quarter = warp & 3, half = warp >> 2;
        const int c4 = tid % GPR, px0 = tid / GPR;   // this thread's channel group and first pixel of a round
        const int groups = p.bn_groups > 0 ? p.bn_groups : 1;
        const int act = p.act & 0xff;
        const bool round = (p.act & ROUND_TF32) != 0;
        const int Ho = p.Ho, Wo = p.Wo;
        const long long img_step = (long long)v.out_H * v.out_W * N;       // elements between images
        const int row_step = v.out_sy * v.out_W * N;                        // ... between tile rows
        const int px_step = v.out_sx * N;                                   // ... between tile columns
        constexpr int RPT = NPIX / 64;               // 32-column groups of the accumulator this warp owns: (2 * rd + half) * 32
        const int nchains = (chunks + g.cpg - 1) / g.cpg;
        float accr[RPT][32];                         // the tile's sums (this thread's TMEM lane = channel, RPT * 32 pixels)
        int j = 0, jc = 0;
        long long t_wait = 0, t_ld = 0;
        const long long t_begin = tma_clock();
        for (int w = blockIdx.x; w < g.num_work; w += gridDim.x, ++j) {
            int t = w / g.n_tiles;
            const int n0 = (w - t * g.n_tiles) * TBM;
            const int tx = t % g.tiles_x; t /= g.tiles_x;
            const int ty = t % g.tiles_y;
            const int b = t / g.tiles_y;
            const int y0 = ty * (MT * TH), x0 = tx * TW;
            const int cv = min(TBM, N - n0);                     // real channels of this Cout tile (multiple of 4)
            // warp-uniform: this lane quarter holds real channels (stacked: quarters 2, 3 hold the lo(W) products of channels 0..63)
            const bool loader = (g.stack ? (quarter & 1) : quarter) * 32 < cv;
            // ---- drain the tile's accumulation chains into registers (TMEM -> registers, fp32 round-to-nearest adds)
            for (int c = 0; c < nchains; ++c, ++jc) {
                const int buf = jc & 1;
                const long long t0 = g.dbg ? tma_clock() : 0;
                tc::mbar_wait(acc_full + buf, (jc >> 1) & 1);
                if (g.dbg) t_wait += tma_clock() - t0;
                tc::fence_after_thread_sync();
                if (loader) {
                    const uint32_t src = tmem_base + (uint32_t)(buf * NPIX) + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 32);
                    const long long tl0 = g.dbg ? tma_clock() : 0;
                    if (c == 0) {
#pragma unroll
                        for (int rd = 0; rd < RPT; ++rd) {
                            uint32_t r[32];
                            tc::tmem_ld32(src + (uint32_t)(rd * 64), r);
                            tc::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) accr[rd][i] = __uint_as_float(r[i]);
                        }
                    } else {
#pragma unroll
                        for (int rd = 0; rd < RPT; ++rd) {
                            uint32_t r[32];
                            tc::tmem_ld32(src + (uint32_t)(rd * 64), r);
                            tc::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) accr[rd][i] += __uint_as_float(r[i]);
                        }
                    }
                    if (g.dbg) t_ld += tma_clock() - tl0;
                }
                tc::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(acc_empty + buf);      // this warp's tcgen05.ld of the buffer are done
            }
            const bool ch_ok = 4 * c4 < cv;
            const int n = n0 + 4 * c4;                            // first of this thread's 4 channels
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr && ch_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            const long long tile_off = (long long)b * img_step + (long long)(y0 * v.out_sy + v.out_oy) * (v.out_W * N) +
                                       (long long)(x0 * v.out_sx + v.out_ox) * N + n;
            float bs1[4] = {0.f, 0.f, 0.f, 0.f}, bs2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int rd = 0; rd < RPT; ++rd) {
                asm volatile("bar.sync 1, 256;" ::: "memory");      // A: everybody has finished reading the previous round
                if (loader) {
                    const int tq = g.stack ? (quarter & 1) : quarter, plane = g.stack ? (quarter >> 1) : 0;
                    float* dst = T + plane * (64 * TS) + (32 * half) * TS + tq * 32 + lane;
                    if (tq * 32 + lane < CT) {            // (a 16-channel tile only has 16 real lanes)
#pragma unroll
                        for (int i = 0; i < 32; ++i) dst[i * TS] = accr[rd][i];     // consecutive lanes = consecutive words
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");      // B: the round's 64 x cv values are in T
#pragma unroll
                for (int k = 0; k < ITER; ++k) {
                    const int px = px0 + k * PXSTEP;
                    const int l = rd * 64 + px;                 // pixel index inside the tile
                    const int row = l >> g.tw_log2, col = l & (TW - 1);
                    if (ch_ok && y0 + row < Ho && x0 + col < Wo) {
                        float4 x = *reinterpret_cast<const float4*>(T + px * TS + 4 * c4);
                        if (g.stack) {                          // + the lo(W) products from TMEM lanes 64 + c
                            const float4 x2 = *reinterpret_cast<const float4*>(T + 64 * TS + px * TS + 4 * c4);
                            x.x += x2.x; x.y += x2.y; x.z += x2.z; x.w += x2.w;
                        }
                        x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
                        const long long off = tile_off + (long long)row * row_step + (long long)col * px_step;
                        if (p.addend != nullptr) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(p.addend + off));
                            x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
                        }
                        if (act != ACT_NONE) { x.x = tc_act(x.x, act); x.y = tc_act(x.y, act); x.z = tc_act(x.z, act); x.w = tc_act(x.w, act); }
                        if (round) { x.x = tf32_round(x.x); x.y = tf32_round(x.y); x.z = tf32_round(x.z); x.w = tf32_round(x.w); }
                        *reinterpret_cast<float4*>(p.out + off) = x;
                        bs1[0] += x.x; bs1[1] += x.y; bs1[2] += x.z; bs1[3] += x.w;
                        bs2[0] += x.x * x.x; bs2[1] += x.y * x.y; bs2[2] += x.z * x.z; bs2[3] += x.w * x.w;
                    }
                }
            }
            if (p.bn_sums != nullptr) {
                // lanes l, l + GPR, l + 2 GPR ... of a warp own the same 4 channels: butterfly them together, then one fp64
                // atomic pair per (warp, channel).  The tile lies in one image, hence in one BatchNorm group.
#pragma unroll
                for (int o = GPR; o < 32; o <<= 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        bs1[q] += __shfl_xor_sync(0xffffffffu, bs1[q], o);
                        bs2[q] += __shfl_xor_sync(0xffffffffu, bs2[q], o);
                    }
                }
                if (lane < GPR && ch_ok) {
                    const int grp = b / (p.B / groups);
                    double* d = p.bn_sums + (((size_t)(w % SCSFM_BN_SLOTS) * groups + grp) * N + n) * 2;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        atomicAdd(d + 2 * q, (double)bs1[q]);
                        atomicAdd(d + 2 * q + 1, (double)bs2[q]);
                    }
                }
            }
        }
        if (g.dbg && tid == 0) {
            g.dbg[blockIdx.x * 8 + 5] = (unsigned long long)t_wait;
            g.dbg[blockIdx.x * 8 + 6] = (unsigned long long)(tma_clock() - t_begin);
            g.dbg[blockIdx.x * 8 + 0] = (unsigned long long)t_ld;        // (overwrites the producer's wait counter)
        }
    }

    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == TMA_EWARPS + 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// per-call knobs (ScsfmConv.tune, include/scsfm.h): no process-global state
static int tune_mt(const ScsfmConv& p) { return (int)((p.tune >> 4) & 3u); }
static int tune_tw(const ScsfmConv& p) { const int t = (int)((p.tune >> 6) & 3u); return t ? t + 2 : 0; }
static int tune_bn(const ScsfmConv& p) { const int t = (int)((p.tune >> 8) & 15u); return t ? 8 << t : 0; }

bool conv_tma_eligible(const ScsfmConv& p, const TcView& v) {
    if ((p.tune & SCSFM_TUNE_NO_TMA) || v.border) return false;
    if (v.in_stride != 1 || v.kh > TMA_MAX_KH || v.kw > TMA_MAX_KH || v.kh < 1 || v.kw < 1) return false;
    if ((p.Cin & 3) != 0 || (p.Cout & 3) != 0) return false;      // 16-byte TMA rows / float4 epilogue
    if (p.bn_sums && p.B % (p.bn_groups > 0 ? p.bn_groups : 1) != 0) return false;
    return true;
}

bool conv_tma_forced(const ScsfmConv& p) { return tune_mt(p) != 0 || tune_bn(p) != 0 || tune_tw(p) != 0; }

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BNW, int MT>
static int launch_tma_cfg(const ScsfmConv& p, const TcView& v, int tw_log2, cudaStream_t st) {
    // one-time opt-in to 227 KB of dynamic shared memory (C++11 thread-safe static initialisation)
    static const cudaError_t attr_rc = cudaFuncSetAttribute(conv_tma_kernel<BNW, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_MAX);
    SCSFM_CHECK_CUDA(attr_rc);
    const int TW = 1 << tw_log2, TH = TBM >> tw_log2;
    TmaGeom g;
    g.tw_log2 = tw_log2;
    g.tiles_x = (p.Wo + TW - 1) / TW;
    g.tiles_y = (p.Ho + MT * TH - 1) / (MT * TH);
    g.n_tiles = (p.Cout + TBM - 1) / TBM;
    g.num_work = g.tiles_x * g.tiles_y * p.B * g.n_tiles;
    g.a_bytes = ((MT * TH + v.kh - 1) * TW * 128 + 1023) / 1024 * 1024;
    // split mode with a single <= 64-channel Cout tile: W and lo(W) stacked on the M side (see TmaGeom.stack)
    g.stack = (BNW <= 64 && p.Cout <= 64 && p.in_lo != nullptr && p.w_lo != nullptr) ? 1 : 0;
    g.stage_bytes = g.a_bytes + v.kh * (g.stack ? TBM : BNW) * 128;
    // 1024 alignment slack + 1024 barrier block + ring + the MMA's over-read past a BNW-row weight tile (M = 128 rows) +
    // the epilogue's transpose buffer T[64][BNW + 4] (two planes when stacked)
    const int fixed = 1024 + 1024 + (g.stack ? 0 : (TBM - BNW) * 128) + (g.stack ? 2 : 1) * 64 * (BNW + 4) * 4;
    g.stages = (TMA_SMEM_MAX - fixed) / g.stage_bytes;
    if (g.stages > TMA_MAX_STAGES) g.stages = TMA_MAX_STAGES;
    if (g.stages < 2) {
        set_error("conv_tma: stage of %d bytes does not fit twice in shared memory", g.stage_bytes);
        return SCSFM_ERR_ARG;
    }
    g.dbg = p.debug;
    // split-accumulate passes: raw x raw, then lo(activations) x raw(weights), then raw(activations) x lo(weights)
    g.npass = 1; g.a_lo = 0; g.w_lo = 0;
    if (p.in_lo != nullptr) { g.a_lo |= 1 << g.npass; ++g.npass; }
    if (p.w_lo != nullptr && !g.stack) { g.w_lo |= 1 << g.npass; ++g.npass; }
    {
        const int chunks = (p.Cin + TBK - 1) / TBK;
        const int mma_per_chunk = v.kw * g.npass * v.kh * (TBK / 8);
        // split mode: chains of ~100 MMAs (bias ~3e-6); plain TF32 (operand rounding 3e-4 dominates): one chain per tile
        g.cpg = g.npass > 1 ? (96 + mma_per_chunk - 1) / mma_per_chunk : chunks;
        if (g.cpg < 1) g.cpg = 1;
    }
    const size_t smem = (size_t)fixed + (size_t)g.stages * g.stage_bytes;
    CUtensorMap amap, wmap, amap_lo, wmap_lo;
    for (int lo = 0; lo < 2; ++lo) {
        const float* base = lo ? p.in_lo : p.in;
        CUtensorMap& amap_ = lo ? amap_lo : amap;
        if (base == nullptr) { amap_lo = amap; continue; }
        // activations [B][Hi][Wi][Cin] (Cin contiguous); box = 32 channels x TW x (MT*TH + kh - 1) x 1, 128B swizzle
        const cuuint64_t gdim[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.B};
        const cuuint64_t gstride[3] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Wi * p.Cin * 4, (cuuint64_t)p.Hi * p.Wi * p.Cin * 4};
        const cuuint32_t box[4] = {(cuuint32_t)TBK, (cuuint32_t)TW, (cuuint32_t)(MT * TH + v.kh - 1), 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = encode_tiled(&amap_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(activations %d x %d x %d x %d) failed with CUresult %d", p.B, p.Hi, p.Wi, p.Cin, (int)r);
            return SCSFM_ERR_CUDA;
        }
    }
    for (int lo = 0; lo < 2; ++lo) {
        const float* base = lo ? p.w_lo : p.w;
        CUtensorMap& wmap_ = lo ? wmap_lo : wmap;
        if (base == nullptr) { wmap_lo = wmap; continue; }
        // weights [Cout][K] (K contiguous); box = 32 columns x BNW rows (rows past Cout are zero-filled)
        const int K = v.kh * v.kw * p.Cin;
        const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)p.Cout};
        const cuuint64_t gstride[1] = {(cuuint64_t)K * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)BNW};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode_tiled(&wmap_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(weights %d x %d) failed with CUresult %d", p.Cout, K, (int)r);
            return SCSFM_ERR_CUDA;
        }
    }
    int ctas = sm_count();
    if (ctas > g.num_work) ctas = g.num_work;
    conv_tma_kernel<BNW, MT><<<ctas, TMA_THREADS, smem, st>>>(p, v, g, amap, wmap, amap_lo, wmap_lo);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}

int launch_conv_tma(const ScsfmConv& p, const TcView& v, cudaStream_t st) {
    const int N = p.Cout;
    // weight rows kept in shared memory: the smallest of 16/32/64/128 covering Cout (Cout > 128: 128-row tiles)
    int bnw;
    if (N <= 16) bnw = 16;
    else if (N <= 32) bnw = 32;
    else if (N <= 64) bnw = 64;
    else bnw = 128;
    // Pixel tile: 1 or 2 stacked 128-pixel sub-tiles (N = 128 / 256 of the MMA), TW = 8 or 16.  An MMA costs the same
    // for N = 128 and N = 256 (measured), so a tile costs the same either way: minimise the number of waves of the
    // persistent CTAs, then the padded area (halo rows included), then prefer the smaller tile.
    const int nsm = sm_count();
    const long nt = (N + TBM - 1) / TBM;
    int best_mt = 1, best_tw = 4;
    double best_cost = -1.0;
    for (int mt = 1; mt <= 2; ++mt)
        for (int twl = 4; twl >= 3; --twl) {
            const int tw = 1 << twl, th = mt * (TBM >> twl);
            const long ty = (p.Ho + th - 1) / th, tx = (p.Wo + tw - 1) / tw;
            const long work = ty * tx * p.B * nt;
            const long waves = (work + nsm - 1) / nsm;
            const double area = (double)(ty * (th + v.kh - 1)) * (double)(tx * tw) / ((double)p.Ho * p.Wo);     // >= 1
            const double cost = (double)waves + 0.05 * area + 0.01 * mt;
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_mt = mt; best_tw = twl; }
        }
    if (tune_mt(p)) best_mt = tune_mt(p);
    if (tune_tw(p)) best_tw = tune_tw(p);
    if (tune_bn(p)) bnw = tune_bn(p) < bnw ? bnw : tune_bn(p);       // never fewer rows than the Cout tile needs
    if (best_mt == 1) {
        switch (bnw) {
            case 16: return launch_tma_cfg<16, 1>(p, v, best_tw, st);
            case 32: return launch_tma_cfg<32, 1>(p, v, best_tw, st);
            case 64: return launch_tma_cfg<64, 1>(p, v, best_tw, st);
            default: return launch_tma_cfg<128, 1>(p, v, best_tw, st);
        }
    }
    switch (bnw) {
        case 16: return launch_tma_cfg<16, 2>(p, v, best_tw, st);
        case 32: return launch_tma_cfg<32, 2>(p, v, best_tw, st);
        case 64: return launch_tma_cfg<64, 2>(p, v, best_tw, st);
        default: return launch_tma_cfg<128, 2>(p, v, best_tw, st);
    }
}

}  // namespace scsfm

