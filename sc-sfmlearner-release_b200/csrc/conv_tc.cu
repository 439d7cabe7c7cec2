// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05, kind::tf32, fp32 accumulate in TMEM).
//
//   C[M = B*Ho*Wo, N = Cout] = A[M, K = kh*kw*Cin] (im2col gather of the NHWC input) x W[N, K]^T
//
// CTA = one 128 x BN output tile, 5 warps:
//   warps 0-3  producers: gather 128 x 32-float A slices (any stride / zero or reflection padding) and BN x 32 W
//              slices with 16-byte loads, write them into the 128B-swizzled K-major layout the UMMA descriptors
//              expect, fence.proxy.async, arrive on the stage's "full" mbarrier; afterwards they are the epilogue
//              (tcgen05.ld 32x32b -> bias / activation / residual addend / BatchNorm partial sums -> 16B stores)
//   warp 4     allocates TMEM, one elected lane issues 4 x tcgen05.mma (M128 x BN x K8) per stage and
//              tcgen05.commit's to the stage's "empty" mbarrier / the accumulator-ready mbarrier.
// A software gather is used instead of TMA-im2col because the same loader folds in reflection padding and the
// transposed-convolution view used for the data gradient; the tile still never touches registers twice.
//
// dgrad (stride 1) is the same kernel run on dout with flipped/transposed weights (weight_flip_kernel below);
// for reflection-padded layers it returns the gradient of the padded tensor (pad' = 2), folded afterwards.
//
// Replaces cuDNN implicit-GEMM fwd/dgrad for every layer with Cin % 4 == 0.
#include <stdlib.h>

#include "conv_tc.cuh"

namespace scsfm {

// STACK = true (split mode, Cout <= BN / 2): the weight tile holds W in rows [0, BN/2) and lo(W) in rows [BN/2, BN), so the two
// passes lo(in) and in give all four products (accumulator columns n and BN/2 + n are added in the epilogue): 2 passes, not 3.
template <int BN, bool STACK = false>
__global__ void __launch_bounds__(FW_THREADS)
conv_fwd_tc_kernel(ScsfmConv p, TcView v, const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap wmap_lo) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_STAGE_BYTES);
    uint64_t* bar_empty = bar_full + STAGES;
    uint64_t* bar_acc = bar_empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows_per_img = v.border ? border_count(p.Ho, p.Wo) : p.Ho * p.Wo;
    const int M = p.B * rows_per_img, N = p.Cout, K = v.kh * v.kw * p.Cin;
    const int m0 = blockIdx.x * TBM, n0 = blockIdx.y * BN;
    const int KB = (K + TBK - 1) / TBK;
    // split-accumulate passes over the whole K range (ScsfmConv.in_lo / w_lo): raw x raw, lo(in) x raw(w), raw(in) x lo(w)
    const int npass = STACK ? 2 : 1 + (p.in_lo != nullptr ? 1 : 0) + (p.w_lo != nullptr ? 1 : 0);
    // round-robin accumulators only in split mode (plain TF32 is bounded by its operand rounding: one accumulator, fewer TMEM
    // columns, more resident CTAs)
    const int nacc_rt = npass > 1 ? Cfg::NACC : 1;
    const uint32_t tmem_cols = (uint32_t)(nacc_rt * Cfg::ACC_COLS);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(bar_full + s, FW_PWARPS * 32 + 1);      // cp.async arrivals (activations) + 1 expect_tx arrival (weights, TMA)
            tc::mbar_init(bar_empty + s, 1);
        }
        tc::mbar_init(bar_acc, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&wmap);
        if (p.w_lo != nullptr) tc::tma_prefetch_desc(&wmap_lo);
    }
    if (warp == FW_PWARPS) tc::tmem_alloc(tmem_slot, tmem_cols);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < FW_PWARPS) {
        // ------------------------------------------------------------------ producers
        // Per k-block every thread fetches 4 A chunks (16 B each); the BN x 32 weight slice is one TMA box.  The loop is software-pipelined:
        // the loads of k-block kb+1 are in flight while k-block kb is written to shared memory, all addressing is
        // 32-bit offset arithmetic and out-of-image taps are handled without branches (clamped address + select).
        constexpr int ROWS = TBM / (FW_PWARPS * 4);    // A rows per thread (4): rows r0 + 32*i
        const int c = tid & 7;               // 16-byte chunk column inside the 128-byte row
        const int r0 = tid >> 3;             // 0..31
        const int cs = c ^ (r0 & 7);         // 128B swizzle: chunk ^= row % 8 (rows r0+32i share row % 8)
        const bool reflect = p.pad_mode == PADMODE_REFLECT;
        int hi0[ROWS], wi0[ROWS], rbase[ROWS];   // first-tap input coordinates and the image offset of each row
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const int m = m0 + r0 + 32 * i;
            if (m < M) {
                const int b = m / rows_per_img, rem = m - b * rows_per_img;
                int ho, wo;
                if (v.border) border_pixel(rem, p.Ho, p.Wo, ho, wo);
                else { ho = rem / p.Wo; wo = rem - ho * p.Wo; }
                hi0[i] = ho * v.in_stride + v.oy0;
                wi0[i] = wo * v.in_stride + v.ox0;
                rbase[i] = b * p.Hi * p.Wi * p.Cin;
            } else {
                hi0[i] = -(1 << 28);          // always out of range -> zero rows
                wi0[i] = -(1 << 28);
                rbase[i] = 0;
            }
        }
        const uint32_t a_smem = tc::smem_u32(sA) + (uint32_t)(r0 * 128 + cs * 16);
        const uint32_t b_smem = tc::smem_u32(sB);
        int kc, dy, dx, ch;
        // Asynchronous copies (cp.async / LDGSTS, 16 B, zero-fill for out-of-image taps): no register staging, so the
        // loads of all pipeline stages are in flight at once; the stage's "full" mbarrier gets this thread's arrival
        // when its copies have landed (cp.async.mbarrier.arrive.noinc).  The MMA thread issues fence.proxy.async after
        // waiting on the barrier to order these generic-proxy writes before the tensor core's async-proxy reads.
        // Row offsets / validity only change with the tap, i.e. every Cin/32 k-blocks: they are cached in between.
        int aoff[ROWS];
        uint32_t okm = 0;
        int it = 0;                                  // k-block counter across the passes: stage = it % STAGES
        for (int q = 0; q < npass; ++q) {
        const int ps = (q + 1) % npass;              // low-part passes first (added while the accumulators are small), raw x raw last
        const float* a_base = (ps == 1 && p.in_lo != nullptr) ? p.in_lo : p.in;
        const CUtensorMap* wm = (ps == npass - 1 && ps > 0 && p.w_lo != nullptr) ? &wmap_lo : &wmap;
        kc = 4 * c;
        {
            const int tap = kc / p.Cin;
            ch = kc - tap * p.Cin;
            dy = tap / v.kw;
            dx = tap - dy * v.kw;
        }
        int cur_dy = -1, cur_dx = -1;
        for (int kb = 0; kb < KB; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            if (lane == 0) tc::mbar_wait(bar_empty + s, ph ^ 1);
            __syncwarp();
            const uint32_t a_st = a_smem + (uint32_t)(s * A_STAGE_BYTES);
            if (tid == 0) {
                // weights: TMA box (32 K-columns x BN rows) straight into the 128B-swizzled stage; rows / columns beyond
                // Cout / K are zero-filled by the hardware and still count towards the expected bytes
                tc::mbar_arrive_expect_tx(bar_full + s, (uint32_t)Cfg::B_STAGE_BYTES);
                if (STACK) {
                    tc::tma_load_2d(b_smem + (uint32_t)(s * Cfg::B_STAGE_BYTES), &wmap, kb * TBK, 0, bar_full + s);
                    tc::tma_load_2d(b_smem + (uint32_t)(s * Cfg::B_STAGE_BYTES + (BN / 2) * 128), &wmap_lo, kb * TBK, 0, bar_full + s);
                } else {
                    tc::tma_load_2d(b_smem + (uint32_t)(s * Cfg::B_STAGE_BYTES), wm, kb * TBK, n0, bar_full + s);
                }
            }
            if (dy != cur_dy || dx != cur_dx) {
                cur_dy = dy; cur_dx = dx;
                okm = 0;
#pragma unroll
                for (int i = 0; i < ROWS; ++i) {
                    int hi = hi0[i] + dy, wi = wi0[i] + dx;
                    bool ok;
                    if (reflect) {
                        ok = hi0[i] > -(1 << 27);
                        hi = reflect_index(hi, p.Hi);
                        wi = reflect_index(wi, p.Wi);
                    } else {
                        ok = (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi;
                    }
                    aoff[i] = ok ? rbase[i] + (hi * p.Wi + wi) * p.Cin : 0;
                    okm |= (ok ? 1u : 0u) << i;
                }
            }
            const bool kok = kc < K;
#pragma unroll
            for (int i = 0; i < ROWS; ++i) {
                const bool ok = kok && ((okm >> i) & 1u);
                tc::cp_async_16(a_st + i * 4096, a_base + (ok ? aoff[i] + ch : 0), ok ? 16u : 0u);
            }
            tc::cp_async_arrive_noinc(bar_full + s);
            // advance this thread's K index by one k-block
            kc += TBK;
            ch += TBK;
            while (ch >= p.Cin) {
                ch -= p.Cin;
                if (++dx == v.kw) { dx = 0; ++dy; }
            }
        }
        }

        // ------------------------------------------------------------------ epilogue (same 4 warps)
        tc::mbar_wait(bar_acc, 0);
        tc::fence_after_thread_sync();
        const int quarter = warp & 3, half = warp >> 2;      // TMEM lane quarter / which half of the column chunks
        const int m = m0 + quarter * 32 + lane;
        const bool row_ok = m < M;
        size_t out_row = (size_t)m;          // row of the output / addend tensors
        if (row_ok && (v.out_sy != 1 || v.out_sx != 1 || v.border)) {
            const int b = m / rows_per_img, rem = m - b * rows_per_img;
            int ho, wo;
            if (v.border) border_pixel(rem, p.Ho, p.Wo, ho, wo);
            else { ho = rem / p.Wo; wo = rem - ho * p.Wo; }
            out_row = ((size_t)b * v.out_H + (ho * v.out_sy + v.out_oy)) * v.out_W + (wo * v.out_sx + v.out_ox);
        }
        float* stage = reinterpret_cast<float*>(sA) + warp * (32 * 33);    // all MMAs retired: operand smem is free
        constexpr int CREAL = STACK ? BN / 2 : BN;                // accumulator columns holding output channels
        constexpr int CW = CREAL < 32 ? CREAL : 32;
        const int nacc = min(nacc_rt, KB * npass);              // accumulators that received at least one k-block
        constexpr int NCH = CREAL / CW;                           // column chunks of real output channels
#pragma unroll 1
        for (int cc = half; cc < NCH; cc += FW_PWARPS / 4) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cc * CW);
            if (CW == 32) tc::tmem_ld32(taddr, r);
            else tc::tmem_ld16(taddr, r);
            tc::tmem_ld_wait();
            // the round-robin partial accumulators (and, stacked, the lo(W) columns BN/2 further on), added in fp32 (RN)
            for (int a = 0; a < nacc; ++a) {
                for (int h = 0; h < (STACK ? 2 : 1); ++h) {
                    if (a == 0 && h == 0) continue;
                    uint32_t q[32];
                    const uint32_t src = taddr + (uint32_t)(a * Cfg::ACC_COLS + h * (BN / 2));
                    if (CW == 32) tc::tmem_ld32(src, q);
                    else tc::tmem_ld16(src, q);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < CW; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(q[j]));
                }
            }
            float v[32];
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
                v[j] = x;
            }
            if (row_ok) {
                float* o = p.out + out_row * N + n0 + cc * CW;
                if ((N & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < CW; j += 4)
                        if (n0 + cc * CW + j < N) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < CW; ++j)
                        if (n0 + cc * CW + j < N) o[j] = v[j];
                }
            }
            if (p.bn_sums != nullptr) {
                // column sums over the warp's 32 rows through a padded smem transpose, then fp64 atomics into one of
                // SCSFM_BN_SLOTS replicas (spreads the per-address serialisation at L2)
#pragma unroll
                for (int j = 0; j < CW; ++j) stage[lane * 33 + j] = v[j];
                __syncwarp();
                if (lane < CW) {
                    // a tile may straddle BatchNorm groups (network calls batched into one launch): walk the warp's
                    // 32 rows and flush the column sums whenever the group changes
                    const int groups = p.bn_groups > 0 ? p.bn_groups : 1;
                    const int rows_per_group = (p.B / groups) * rows_per_img;
                    const int row_base = m0 + quarter * 32;
                    const int n = n0 + cc * CW + lane;
                    int g_cur = row_base / rows_per_group;
                    int next_edge = (g_cur + 1) * rows_per_group - row_base;      // first row index of the next group
                    // fp64 accumulation: var = E[x^2] - mean^2 cancels catastrophically when |mean| >> std, so the
                    // partial sums must not carry fp32 rounding (B200 issues DFMA at half the FFMA rate: negligible here)
                    double s1 = 0.0, s2 = 0.0;
                    for (int rr = 0; rr <= 32; ++rr) {
                        if (rr == 32 || rr == next_edge) {
                            if (n < N && g_cur < groups && (s1 != 0.0 || s2 != 0.0)) {
                                double* d = p.bn_sums + (((size_t)(blockIdx.x % SCSFM_BN_SLOTS) * groups + g_cur) * N + n) * 2;
                                atomicAdd(d, s1);
                                atomicAdd(d + 1, s2);
                            }
                            if (rr == 32) break;
                            s1 = 0.0; s2 = 0.0;
                            ++g_cur;
                            next_edge += rows_per_group;
                        }
                        const double t = (double)stage[rr * 33 + lane];
                        s1 += t;
                        s2 += t * t;
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer (warp 4)
        constexpr uint32_t idesc = tc::make_idesc_tf32(TBM, BN, 0, 0);
        if (lane == 0) {                                 // one thread waits, issues and commits
            for (int kb = 0; kb < KB * npass; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                tc::mbar_wait(bar_full + s, ph);
                tc::fence_proxy_async();                 // cp.async (generic proxy) writes -> UMMA (async proxy) reads
                tc::fence_after_thread_sync();
                const uint32_t a_addr = tc::smem_u32(sA + s * A_STAGE_BYTES);
                const uint32_t b_addr = tc::smem_u32(sB + s * Cfg::B_STAGE_BYTES);
#pragma unroll
                for (int j = 0; j < TBK / 8; ++j) {
                    const uint64_t da = tc::make_smem_desc(a_addr + j * 32, 16, 1024, tc::LAYOUT_SW128);
                    const uint64_t db = tc::make_smem_desc(b_addr + j * 32, 16, 1024, tc::LAYOUT_SW128);
                    tc::mma_tf32(tmem_base + (uint32_t)((kb % nacc_rt) * Cfg::ACC_COLS), da, db, idesc, (kb >= nacc_rt || j != 0) ? 1u : 0u);
                }
                tc::mma_commit(bar_empty + s);           // frees the stage once these MMAs have read it
            }
            tc::mma_commit(bar_acc);                     // accumulator complete
        }
        __syncwarp();
    }

    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == FW_PWARPS) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// w [Co][kh][kw][Ci] -> wt [Ci][jh][jw][Co] with wt[c][jy][jx][o] = w[o][dy_max - step*jy][dx_max - step*jx][c]:
// the (TF32-rounded) weights of the transposed conv.  step 1, d*_max = k-1: the full flipped kernel (stride-1 dgrad);
// step 2: the taps of one output-parity class of a stride-2 dgrad.
__global__ void weight_flip_kernel(const float* __restrict__ w, int Co, int kh, int kw, int Ci, int jh, int jw, int dy_max,
                                   int dx_max, int step, float* __restrict__ wt, int operand) {
    const long long total = (long long)Co * jh * jw * Ci;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int o = (int)(i % Co);
        long long t2 = i / Co;
        const int jx = (int)(t2 % jw); t2 /= jw;
        const int jy = (int)(t2 % jh);
        const int c = (int)(t2 / jh);
        const int dy = dy_max - step * jy, dx = dx_max - step * jx;
        wt[i] = tc_operand(__ldg(w + (((size_t)o * kh + dy) * kw + dx) * Ci + c), operand);
    }
}


// ---------------------------------------------------------------------------------------------------------
// weight gradient on the tensor cores:  dW^T[(tap,c), o] += sum_pix  in[pix (+) tap, c] * dout[pix, o]
//   GEMM  M = kh*kw*Cin (gathered input, "MN-major": channels contiguous),  N = Cout (dout, "MN-major"),
//   K = B*Ho*Wo pixels, split over gridDim.z; partial tiles are added into dw with fp32 atomics.
// Both operands are MN-major (channels contiguous).  For 32-bit MN-major operands the only shared-memory layout the
// UMMA unit accepts is SWIZZLE_128B_BASE32B (layout type 1): atom = 4 pixels (K) x 32 channels (128-byte rows),
// 32-byte chunks XOR-swizzled with (pixel % 4); atoms along M/N at LBO = 512 B, 4-pixel groups at SBO.  One
// tcgen05.mma (K = 8 for tf32) therefore consumes two pixel groups.
// ---------------------------------------------------------------------------------------------------------
template <int BN>
struct WgCfg {
    static constexpr int STAGES = 3;
    static constexpr int A_BYTES = 32 * TBM * 4;        // 32 pixels x 128 (tap,c)
    static constexpr int B_BYTES = 32 * BN * 4;         // 32 pixels x BN output channels
    static constexpr int NACC = BN >= 128 ? 1 : (BN >= 64 ? 2 : 4);      // round-robin accumulators (see TcCfg); 3 CTAs per SM fit
    static constexpr int ACC_COLS = BN < 32 ? 32 : BN;
    static constexpr int TMEM_COLS = NACC * ACC_COLS;   // 128
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * (A_BYTES + B_BYTES) + 256;
};

// BORDER = true (used after the zero-padding TMA weight-gradient kernel on a reflection-padded layer): the K dimension only
// runs over the 2*(Ho+Wo)-4 border pixels of every image and only the taps that fall OUTSIDE the image contribute, read
// at their reflected positions -- exactly the part of the gradient the zero-padded pass left out.
// STACK = true (split mode, Cout <= BN / 2): the N-side tile holds dout in columns [0, BN/2) and lo(dout) in [BN/2, BN), so the two
// passes lo(in) and in give all four products (columns o and BN/2 + o are both added into dw[o]): 2 passes instead of 3.
template <int BN, bool BORDER = false, bool STACK = false>
__global__ void __launch_bounds__(FW_THREADS)
conv_wgrad_tc_kernel(ScsfmConv p, int pix_per_split) {
    using Cfg = WgCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
    uint64_t* bar_empty = bar_full + STAGES;
    uint64_t* bar_acc = bar_empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = BORDER ? border_count(p.Ho, p.Wo) : p.Ho * p.Wo;        // K pixels per image
    const int Mtot = p.kh * p.kw * p.Cin, N = p.Cout, npix = p.B * nb;
    const int m0 = blockIdx.x * TBM, n0 = blockIdx.y * BN;
    const int pix_begin = blockIdx.z * pix_per_split, pix_end = min(npix, pix_begin + pix_per_split);
    const int KB = (pix_end - pix_begin + 31) / 32;
    if (KB <= 0) return;
    // split-accumulate passes over the CTA's pixel range (ScsfmConv.in_lo / dout_lo): raw x raw, lo(in) x raw(dout), raw(in) x lo(dout)
    const int npass = STACK ? 2 : 1 + (p.in_lo != nullptr ? 1 : 0) + (p.dout_lo != nullptr ? 1 : 0);
    const int nacc_rt = npass > 1 ? Cfg::NACC : 1;            // round-robin accumulators only in split mode
    const uint32_t tmem_cols = (uint32_t)(nacc_rt * Cfg::ACC_COLS);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(bar_full + s, FW_PWARPS * 32);
            tc::mbar_init(bar_empty + s, 1);
        }
        tc::mbar_init(bar_acc, 1);
        tc::fence_barrier_init();
    }
    if (warp == FW_PWARPS) tc::tmem_alloc(tmem_slot, tmem_cols);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < FW_PWARPS) {
        // ------------------------------------------------------------------ producers
        // A: chunk c4 = tid % 32 (4 consecutive (tap,c) entries, fixed for the whole kernel) x PPT CONSECUTIVE pixels
        // starting at PPT * (tid / 32): a warp reads 512 contiguous bytes per pixel, and stepping to the next pixel is
        // one add (+ a rare row wrap).  cp.async with zero-fill, completion on the stage's mbarrier.
        constexpr int PPT = 32 / FW_PWARPS;                   // pixels per thread per k-block (4)
        const int c4 = tid & 31, kq = tid >> 5;              // kq: which group of PPT pixels of the 32-pixel k-block
        const int mm = m0 + 4 * c4;
        const bool a_ok = mm < Mtot;
        int a_dy = 0, a_dx = 0, a_ch = 0;
        if (a_ok) {
            const int tap = mm / p.Cin;
            a_ch = mm - tap * p.Cin;
            a_dy = tap / p.kw - p.pad;
            a_dx = tap - (tap / p.kw) * p.kw - p.pad;
        }
        const bool reflect = p.pad_mode == PADMODE_REFLECT;
        // smem byte offsets of this thread's 8 A chunks inside a stage: pixel k = 8*kq + i -> group k/4, row k%4
        const uint32_t a_smem = tc::smem_u32(sA) + (uint32_t)((c4 >> 3) * 512);
        const int a_chunk = c4 & 7;
        // B: BN/4 chunks per pixel row
        constexpr int BCH = BN / 4;                          // chunks per row: 8, 16 or 32
        constexpr int B_IT = (32 * BCH) / (FW_PWARPS * 32);  // per-thread chunk loads: 1, 2, 4
        constexpr int B_STEP = (FW_PWARPS * 32) / BCH;       // pixel-row step between a thread's B chunks
        const int b_c4 = tid % BCH, b_kr0 = tid / BCH;
        // stacked: chunk columns >= BN/2 read lo(dout) at channel (column - BN/2)
        const bool b_lo_half = STACK && 4 * b_c4 >= BN / 2;
        const int nn = STACK ? (4 * b_c4 - (b_lo_half ? BN / 2 : 0)) : n0 + 4 * b_c4;
        const bool b_ok = nn < N;
        const uint32_t b_smem = tc::smem_u32(sB) + (uint32_t)((b_c4 >> 3) * 512);
        // pixel (b, ho, wo) of this thread's first A row in the current k-block, advanced by 32 per block
        int pb, pho, pwo;
        int it = 0;                                  // k-block counter across the passes: stage = it % STAGES
        for (int q = 0; q < npass; ++q) {
        const int ps = (q + 1) % npass;              // low-part passes first (added while the accumulators are small), raw x raw last
        const float* a_base = (ps == 1 && p.in_lo != nullptr) ? p.in_lo : p.in;
        const float* b_base = STACK ? (b_lo_half ? p.dout_lo : p.dout)
                                    : ((ps == npass - 1 && ps > 0 && p.dout_lo != nullptr) ? p.dout_lo : p.dout);
        {
            const int px = pix_begin + PPT * kq;
            pb = px / (p.Ho * p.Wo);
            const int rem = px - pb * p.Ho * p.Wo;
            pho = rem / p.Wo;
            pwo = rem - pho * p.Wo;
        }
        int pix0 = pix_begin;
        for (int kb = 0; kb < KB; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            if (lane == 0) tc::mbar_wait(bar_empty + s, ph ^ 1);
            __syncwarp();
            const uint32_t a_st = a_smem + (uint32_t)(s * Cfg::A_BYTES), b_st = b_smem + (uint32_t)(s * Cfg::B_BYTES);
            int b = pb, ho = pho, wo = pwo;
#pragma unroll
            for (int i = 0; i < PPT; ++i) {
                const int k = PPT * kq + i;                 // pixel row inside the block: group k/4, row k%4
                const int px = pix0 + k;
                bool ok = a_ok && px < pix_end;
                if (BORDER) {
                    // k-th border pixel of the block; only taps leaving the image count, at their mirrored position
                    b = px / nb;
                    border_pixel(px - b * nb, p.Ho, p.Wo, ho, wo);
                }
                int hi = ho * p.stride + a_dy, wi = wo * p.stride + a_dx;
                if (BORDER) {
                    ok = ok && !((unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi);
                    hi = reflect_index(hi, p.Hi); wi = reflect_index(wi, p.Wi);
                } else if (reflect) { hi = reflect_index(hi, p.Hi); wi = reflect_index(wi, p.Wi); }
                else ok = ok && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi;
                const int off = ok ? ((b * p.Hi + hi) * p.Wi + wi) * p.Cin + a_ch : 0;
                tc::cp_async_16(a_st + (k >> 2) * (4 * 512) + (k & 3) * 128 + (((((a_chunk >> 1) ^ (k & 3)) << 1) | (a_chunk & 1)) * 16),
                                a_base + off, ok ? 16u : 0u);
                if (!BORDER) { if (++wo == p.Wo) { wo = 0; if (++ho == p.Ho) { ho = 0; ++b; } } }
            }
#pragma unroll
            for (int i = 0; i < B_IT; ++i) {
                const int k = b_kr0 + B_STEP * i;
                const int px = pix0 + k;
                const bool ok = b_ok && px < pix_end;
                size_t row = (size_t)px;                       // row of dout
                if (BORDER && ok) {
                    const int bb = px / nb;
                    int bho, bwo;
                    border_pixel(px - bb * nb, p.Ho, p.Wo, bho, bwo);
                    row = ((size_t)bb * p.Ho + bho) * p.Wo + bwo;
                }
                tc::cp_async_16(b_st + (k >> 2) * ((BN / 32) * 512) + (k & 3) * 128 + ((((((b_c4 & 7) >> 1) ^ (k & 3)) << 1) | (b_c4 & 1)) * 16),
                                b_base + (ok ? row * N + nn : 0), ok ? 16u : 0u);
            }
            tc::cp_async_arrive_noinc(bar_full + s);
            pix0 += 32;
            pwo += 32;
            while (pwo >= p.Wo) { pwo -= p.Wo; if (++pho == p.Ho) { pho = 0; ++pb; } }
        }
        }

        // ------------------------------------------------------------------ epilogue: dw[o][mm] += D[mm][o]
        tc::mbar_wait(bar_acc, 0);
        tc::fence_after_thread_sync();
        const int quarter = warp & 3, half = warp >> 2;
        const int row = m0 + quarter * 32 + lane;
        const int nacc = min(nacc_rt, KB * npass);
#pragma unroll 1
        for (int cc = half; cc < BN / 32; cc += FW_PWARPS / 4) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cc * 32);
            tc::tmem_ld32(taddr, r);
            tc::tmem_ld_wait();
            for (int a = 1; a < nacc; ++a) {                   // round-robin partial accumulators, added in fp32 (RN)
                uint32_t q[32];
                tc::tmem_ld32(taddr + (uint32_t)(a * Cfg::ACC_COLS), q);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(q[j]));
            }
            if (row < Mtot) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = cc * 32 + j;
                    const int o = STACK ? (col >= BN / 2 ? col - BN / 2 : col) : n0 + col;
                    if (o < N) red_add(p.dw + (size_t)o * Mtot + row, __uint_as_float(r[j]));
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer (warp 4)
        constexpr uint32_t idesc = tc::make_idesc_tf32(TBM, BN, 1, 1);       // both operands MN-major
        if (lane == 0) {
            for (int kb = 0; kb < KB * npass; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                tc::mbar_wait(bar_full + s, ph);
                tc::fence_proxy_async();
                tc::fence_after_thread_sync();
                const uint32_t a_addr = tc::smem_u32(sA + s * Cfg::A_BYTES);
                const uint32_t b_addr = tc::smem_u32(sB + s * Cfg::B_BYTES);
#pragma unroll
                for (int j = 0; j < 4; ++j) {               // 4 x (2 pixel groups of 4): UMMA K = 8 for tf32
                    // LBO = 512 B between 32-channel atoms, SBO = distance between 4-pixel groups; 2 groups per MMA
                    const uint64_t da = tc::make_smem_desc(a_addr + j * (2 * 4 * 512), 512, 4 * 512, tc::LAYOUT_SW128_BASE32B);
                    const uint64_t db = tc::make_smem_desc(b_addr + j * (2 * (BN / 32) * 512), 512, (BN / 32) * 512, tc::LAYOUT_SW128_BASE32B);
                    tc::mma_tf32(tmem_base + (uint32_t)((kb % nacc_rt) * Cfg::ACC_COLS), da, db, idesc, (kb >= nacc_rt || j != 0) ? 1u : 0u);
                }
                tc::mma_commit(bar_empty + s);
            }
            tc::mma_commit(bar_acc);
        }
        __syncwarp();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == FW_PWARPS) tc::tmem_dealloc(tmem_base, tmem_cols);
}

CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, cuuint32_t rank, void* gaddr, const cuuint64_t* gdim,
                      const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                      CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
    using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Fn fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || sym == nullptr ||
            qres != cudaDriverEntryPointSuccess)
            return CUDA_ERROR_NOT_FOUND;
        fn = reinterpret_cast<Fn>(sym);
    }
    return fn(map, dtype, rank, gaddr, gdim, gstride, box, estr, il, sw, l2, oob);
}

int launch_bias_grad(const float* dout, int rows, int C, float* dbias, cudaStream_t st);   // conv_simt.cu

template <int BN, bool BORDER = false, bool STACK = false>
static int launch_wgrad_tc(const ScsfmConv& p, cudaStream_t st) {
    using Cfg = WgCfg<BN>;
    static const cudaError_t attr_rc = cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN, BORDER, STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    SCSFM_CHECK_CUDA(attr_rc);
    const int Mtot = p.kh * p.kw * p.Cin, npix = p.B * (BORDER ? border_count(p.Ho, p.Wo) : p.Ho * p.Wo);
    const int mt = (Mtot + TBM - 1) / TBM, nt = STACK ? 1 : (p.Cout + BN - 1) / BN;
    int splits = (148 * 3 + mt * nt - 1) / (mt * nt);          // 3 CTAs per SM fit
    const int max_splits = (npix + 1023) / 1024;             // at least 32 k-blocks per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int npass = STACK ? 2 : 1 + (p.in_lo != nullptr ? 1 : 0) + (p.dout_lo != nullptr ? 1 : 0);
    if (npass > 1) {
        // split-accumulate (parity) mode: bound every accumulation chain to ~160 tcgen05.mma (truncation bias ~5e-6):
        // chain = k-blocks * 4 MMAs * passes / NACC.  More, shorter CTAs; their partial tiles are added with fp32 atomics (RN)
        const int kb_max = 160 * Cfg::NACC / (4 * npass);
        const int need = (npix + 32 * kb_max - 1) / (32 * kb_max);
        if (splits < need) splits = need;
    }
    const int pps = ((npix + splits - 1) / splits + 31) / 32 * 32;
    dim3 grid(mt, nt, (npix + pps - 1) / pps);
    conv_wgrad_tc_kernel<BN, BORDER, STACK><<<grid, FW_THREADS, Cfg::SMEM, st>>>(p, pps);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}

static TcView plain_view(const ScsfmConv& p) {
    return TcView{p.kh, p.kw, -p.pad, -p.pad, p.stride, 1, 0, 1, 0, p.Ho, p.Wo};
}

template <int BN, bool STACK = false>
static int launch_fwd_tc(const ScsfmConv& p, const TcView& v, cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static const cudaError_t attr_rc = cudaFuncSetAttribute(conv_fwd_tc_kernel<BN, STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    SCSFM_CHECK_CUDA(attr_rc);
    const int M = p.B * (v.border ? border_count(p.Ho, p.Wo) : p.Ho * p.Wo);
    // TMA descriptor of the weight matrix [Cout rows][K columns] (K contiguous), box = 32 columns x BN rows, 128B swizzle
    const int K = v.kh * v.kw * p.Cin;
    CUtensorMap wmap, wmap_lo;
    for (int lo = 0; lo < 2; ++lo) {
        const float* base = lo ? p.w_lo : p.w;
        CUtensorMap& wmap_ = lo ? wmap_lo : wmap;
        if (base == nullptr) { wmap_lo = wmap; continue; }
        const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)p.Cout};
        const cuuint64_t gstride[1] = {(cuuint64_t)K * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)(STACK ? BN / 2 : BN)};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode_tiled(&wmap_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(weights %d x %d) failed with CUresult %d", p.Cout, K, (int)r);
            return SCSFM_ERR_CUDA;
        }
    }
    dim3 grid((M + TBM - 1) / TBM, STACK ? 1 : (p.Cout + BN - 1) / BN);
    conv_fwd_tc_kernel<BN, STACK><<<grid, FW_THREADS, Cfg::SMEM, st>>>(p, v, wmap, wmap_lo);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}

}  // namespace scsfm

using namespace scsfm;

// cp.async gather kernel: any stride / padding mode / border-only rows
static int tc_dispatch_gather(const ScsfmConv& p, const TcView& v, cudaStream_t st) {
    const int N = p.Cout;
    if (p.in_lo != nullptr && p.w_lo != nullptr && N <= 64) {       // split mode, thin: W / lo(W) stacked on the N side
        if (N <= 16) return launch_fwd_tc<32, true>(p, v, st);
        if (N <= 32) return launch_fwd_tc<64, true>(p, v, st);
        return launch_fwd_tc<128, true>(p, v, st);
    }
    if (N <= 16) return launch_fwd_tc<16>(p, v, st);
    if (N <= 32 || N % 64 != 0) return launch_fwd_tc<32>(p, v, st);
    if (N <= 64 || N % 128 != 0) return launch_fwd_tc<64>(p, v, st);
    // prefer more CTAs when the M extent is small (deep layers at 8x26 / 16x52)
    const int M = p.B * (v.border ? border_count(p.Ho, p.Wo) : p.Ho * p.Wo);
    if (((M + TBM - 1) / TBM) * (N / 128) < 148) return launch_fwd_tc<64>(p, v, st);
    return launch_fwd_tc<128>(p, v, st);
}

static int tc_dispatch(const ScsfmConv& p, const TcView& v, cudaStream_t st) {
    if (p.pad_mode == PADMODE_ZERO && conv_tma_eligible(p, v)) return launch_conv_tma(p, v, st);
    if (p.pad_mode == PADMODE_REFLECT && p.bn_sums == nullptr && p.Ho >= 3 && p.Wo >= 3 &&
        (p.Ho * p.Wo >= 64 * 208 || conv_tma_forced(p)) && conv_tma_eligible(p, v)) {
        // reflection padding only changes the outermost ring of output pixels: run the TMA kernel with zero padding
        // (interior exact), then recompute the 2*(Ho+Wo)-4 border pixels per image with the reflecting gather kernel.
        // Measured (tools/check_conv_tma.py): pays off from 64x208 upwards; below that the ring is too large a share
        // of the image and the gather kernel alone is faster.
        ScsfmConv q = p;
        q.pad_mode = PADMODE_ZERO;
        if (int rc = launch_conv_tma(q, v, st)) return rc;
        TcView bv = v;
        bv.border = 1;
        return tc_dispatch_gather(p, bv, st);
    }
    return tc_dispatch_gather(p, v, st);
}

static int check_tc(const ScsfmConv* p, const char* who) {
    SCSFM_CHECK_ARG(p != nullptr && p->in && p->w && p->out, "%s: null tensor", who);
    SCSFM_CHECK_ARG(p->B > 0 && p->Hi > 0 && p->Wi > 0 && p->Cin > 0 && p->Cout > 0 && p->kh > 0 && p->kw > 0 && p->stride > 0 && p->pad >= 0,
                    "%s: bad geometry", who);
    SCSFM_CHECK_ARG((p->Cin & 3) == 0, "%s: the tensor-core kernel needs Cin %% 4 == 0 (got %d); use the CUDA-core kernel", who, p->Cin);
    SCSFM_CHECK_ARG(p->Ho == (p->Hi + 2 * p->pad - p->kh) / p->stride + 1 && p->Wo == (p->Wi + 2 * p->pad - p->kw) / p->stride + 1,
                    "%s: output size does not match geometry", who);
    SCSFM_CHECK_ARG(p->pad_mode != PADMODE_REFLECT || p->pad == 1, "%s: reflect pad must be 1", who);
    if (p->bn_sums) {
        const int g = p->bn_groups > 0 ? p->bn_groups : 1;
        SCSFM_CHECK_ARG(p->B % g == 0, "%s: batch not divisible by the number of BatchNorm groups", who);
    }
    return SCSFM_OK;
}

extern "C" int scsfm_conv2d_fwd_tc(const ScsfmConv* p, void* stream) {
    if (int rc = check_tc(p, "conv2d_fwd_tc")) return rc;
    return tc_dispatch(*p, plain_view(*p), (cudaStream_t)stream);
}

extern "C" int scsfm_weight_flip(const float* w, int Cout, int kh, int kw, int Cin, float* wt, int operand, void* stream) {
    SCSFM_CHECK_ARG(w && wt && Cout > 0 && kh > 0 && kw > 0 && Cin > 0 && operand >= 0 && operand <= 2, "weight_flip: bad arguments");
    const long long total = (long long)Cout * kh * kw * Cin;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    weight_flip_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, Cout, kh, kw, Cin, kh, kw, kh - 1, kw - 1, 1, wt, operand);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}

// All flips of a network in ONE launch.  table = n rows of 12 int64: {src pointer, dst pointer, Co, kh, kw, Ci, jh, jw,
// dy_max, dx_max, step, first block}; a row is one weight_flip_kernel job (a stride-2 layer contributes one row per parity
// class with taps) and owns blocks [first block, next row's first block); row n is a sentinel holding the total.
// Each block transposes one 32 (Cout) x 32 (Cin) tile of one tap through shared memory: reads coalesced along Cin, writes
// coalesced along Cout.  A row owns ceil(Co/32) * ceil(Ci/32) * jh * jw blocks.
__global__ void __launch_bounds__(256)
weight_flip_batched_kernel(const long long* __restrict__ table, int n) {
    __shared__ int s_row;
    __shared__ float tile[32][33];
    if (threadIdx.x == 0) {
        int lo = 0, hi = n;                       // last row with first block <= blockIdx.x
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (table[mid * 12 + 11] <= (long long)blockIdx.x) lo = mid; else hi = mid;
        }
        s_row = lo;
    }
    __syncthreads();
    const long long* e = table + s_row * 12;
    const float* w = reinterpret_cast<const float*>(e[0]);
    float* wt = reinterpret_cast<float*>(e[1]);
    const int Co = (int)e[2], kh = (int)e[3], kw = (int)e[4], Ci = (int)e[5], jh = (int)e[6], jw = (int)e[7];
    const int dy_max = (int)e[8], dx_max = (int)e[9], step = (int)(e[10] & 0xff), operand = (int)(e[10] >> 8);
    const int tco = (Co + 31) >> 5, tci = (Ci + 31) >> 5;
    int t = (int)((long long)blockIdx.x - e[11]);
    const int to = t % tco; t /= tco;
    const int tc = t % tci; t /= tci;
    const int jy = t / jw, jx = t - jy * jw;
    const int dy = dy_max - step * jy, dx = dx_max - step * jx;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int o = to * 32 + ty + 8 * k, c = tc * 32 + tx;
        if (o < Co && c < Ci) tile[ty + 8 * k][tx] = tc_operand(__ldg(w + (((size_t)o * kh + dy) * kw + dx) * Ci + c), operand);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = tc * 32 + ty + 8 * k, o = to * 32 + tx;
        if (c < Ci && o < Co) wt[(((size_t)c * jh + jy) * jw + jx) * Co + o] = tile[tx][ty + 8 * k];
    }
}

extern "C" int scsfm_weight_flip_batched(const long long* table, int n_rows, int total_blocks, void* stream) {
    SCSFM_CHECK_ARG(table != nullptr && n_rows > 0 && total_blocks > 0, "weight_flip_batched: bad arguments");
    weight_flip_batched_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>(table, n_rows);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}

// Stride-2 data gradient: wt4 receives the four parity-class weight sets back to back (class order (py,px) =
// (0,0),(0,1),(1,0),(1,1)); total size = Cin*kh*kw*Cout floats, the same as the full flipped kernel.
extern "C" int scsfm_weight_flip_s2(const float* w, int Cout, int kh, int kw, int Cin, int pad, float* wt4, int operand, void* stream) {
    SCSFM_CHECK_ARG(w && wt4 && Cout > 0 && kh > 0 && kw > 0 && Cin > 0 && pad >= 0 && operand >= 0 && operand <= 2, "weight_flip_s2: bad arguments");
    size_t off = 0;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            int dy_max = kh - 1, dx_max = kw - 1;
            while (dy_max >= 0 && ((py + pad - dy_max) & 1)) --dy_max;
            while (dx_max >= 0 && ((px + pad - dx_max) & 1)) --dx_max;
            const int jh = dy_max < 0 ? 0 : dy_max / 2 + 1, jw = dx_max < 0 ? 0 : dx_max / 2 + 1;
            const long long total = (long long)Cout * jh * jw * Cin;
            if (total > 0) {
                int grid = (int)((total + 255) / 256);
                if (grid > 148 * 16) grid = 148 * 16;
                weight_flip_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, Cout, kh, kw, Cin, jh, jw, dy_max, dx_max, 2, wt4 + off, operand);
                SCSFM_CHECK_LAUNCH();
            }
            off += (size_t)total;
        }
    return SCSFM_OK;
}

// stride-1 data gradient = forward kernel on dout with the flipped weights `wt` ([Cin][kh][kw][Cout], from
// scsfm_weight_flip) passed in p->w; p->din [B,Hi,Wi,Cin] (+ p->addend).
extern "C" int scsfm_conv2d_dgrad_tc(const ScsfmConv* p, void* stream) {
    SCSFM_CHECK_ARG(p != nullptr && p->dout && p->w && p->din, "conv2d_dgrad_tc: null tensor");
    SCSFM_CHECK_ARG(p->stride == 1 || p->stride == 2, "conv2d_dgrad_tc: stride must be 1 or 2");
    SCSFM_CHECK_ARG(p->kh == p->kw && p->kh - 1 - p->pad >= 0, "conv2d_dgrad_tc: square kernels only");
    cudaStream_t st = (cudaStream_t)stream;
    ScsfmConv q = *p;
    q.in = p->dout; q.in_lo = p->dout_lo; q.out = p->din; q.bias = nullptr; q.bn_sums = nullptr; q.act = SCSFM_ACT_NONE;
    q.dout_lo = nullptr;                       // (w_lo: the flipped low-part weights, laid out like w)
    q.Hi = p->Ho; q.Wi = p->Wo; q.Cin = p->Cout;
    q.Cout = p->Cin; q.pad_mode = SCSFM_PADMODE_ZERO;
    if (p->stride == 1) {
        q.Ho = p->Hi; q.Wo = p->Wi;
        q.pad = p->kh - 1 - p->pad;
        if (int rc = check_tc(&q, "conv2d_dgrad_tc")) return rc;
        return tc_dispatch(q, plain_view(q), st);
    }
    // stride 2: p->w holds the four parity-class weight sets of scsfm_weight_flip_s2
    SCSFM_CHECK_ARG((q.Cin & 3) == 0, "conv2d_dgrad_tc: needs Cout %% 4 == 0");
    q.stride = 1;
    size_t woff = 0;
    bool covered_all = true;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            int dy_max = p->kh - 1, dx_max = p->kw - 1;
            while (dy_max >= 0 && ((py + p->pad - dy_max) & 1)) --dy_max;
            while (dx_max >= 0 && ((px + p->pad - dx_max) & 1)) --dx_max;
            const int jh = dy_max < 0 ? 0 : dy_max / 2 + 1, jw = dx_max < 0 ? 0 : dx_max / 2 + 1;
            const int Hs = (p->Hi - py + 1) / 2, Ws = (p->Wi - px + 1) / 2;
            if (jh == 0 || jw == 0) { if (Hs > 0 && Ws > 0) covered_all = false; }
            woff += (size_t)p->Cout * jh * jw * p->Cin;
        }
    if (!covered_all) {
        // parity classes without taps (1x1 stride 2): their gradient is the addend alone (or zero)
        const size_t bytes = (size_t)p->B * p->Hi * p->Wi * p->Cin * sizeof(float);
        if (p->addend) SCSFM_CHECK_CUDA(cudaMemcpyAsync(p->din, p->addend, bytes, cudaMemcpyDeviceToDevice, st));
        else SCSFM_CHECK_CUDA(cudaMemsetAsync(p->din, 0, bytes, st));
    }
    woff = 0;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            int dy_max = p->kh - 1, dx_max = p->kw - 1;
            while (dy_max >= 0 && ((py + p->pad - dy_max) & 1)) --dy_max;
            while (dx_max >= 0 && ((px + p->pad - dx_max) & 1)) --dx_max;
            const int jh = dy_max < 0 ? 0 : dy_max / 2 + 1, jw = dx_max < 0 ? 0 : dx_max / 2 + 1;
            const int Hs = (p->Hi - py + 1) / 2, Ws = (p->Wi - px + 1) / 2;
            const size_t wcount = (size_t)p->Cout * jh * jw * p->Cin;
            if (jh > 0 && jw > 0 && Hs > 0 && Ws > 0) {
                ScsfmConv r = q;
                r.w = p->w + woff;
                r.w_lo = p->w_lo ? p->w_lo + woff : nullptr;
                r.Ho = Hs; r.Wo = Ws; r.kh = jh; r.kw = jw;
                // input (dout) row of output hy and tap jy: hy + (py + pad - dy_max)/2 + jy
                TcView v{jh, jw, (py + p->pad - dy_max) / 2, (px + p->pad - dx_max) / 2, 1, 2, py, 2, px, p->Hi, p->Wi};
                if (int rc = tc_dispatch(r, v, st)) return rc;
            }
            woff += wcount;
        }
    return SCSFM_OK;
}

// dw [Cout,kh,kw,Cin] += dout^T x gather(in); dbias += column sums of dout.  Needs Cin % 4 == 0 and Cout % 4 == 0.
extern "C" int scsfm_conv2d_wgrad_tc(const ScsfmConv* p, void* stream) {
    SCSFM_CHECK_ARG(p != nullptr && p->in && p->dout && p->dw, "conv2d_wgrad_tc: null tensor");
    SCSFM_CHECK_ARG(p->B > 0 && p->Hi > 0 && p->Wi > 0 && p->Cin > 0 && p->Cout > 0 && p->kh > 0 && p->kw > 0 && p->stride > 0 && p->pad >= 0,
                    "conv2d_wgrad_tc: bad geometry");
    SCSFM_CHECK_ARG((p->Cin & 3) == 0 && (p->Cout & 3) == 0, "conv2d_wgrad_tc: needs Cin %% 4 == 0 and Cout %% 4 == 0");
    SCSFM_CHECK_ARG(p->pad_mode != PADMODE_REFLECT || p->pad == 1, "conv2d_wgrad_tc: reflect pad must be 1");
    SCSFM_CHECK_ARG((long long)p->B * p->Ho * p->Wo < (1LL << 31), "conv2d_wgrad_tc: too many pixels");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    ScsfmConv zp = *p;                         // the same layer with zero padding (what the TMA kernel computes)
    zp.pad_mode = SCSFM_PADMODE_ZERO;
    int kernel = (int)((p->tune >> 12) & 3u);       // SCSFM_TUNE_WGRAD: 0 auto, 1 cp.async kernel, 2 TMA kernel, 3 thin-layer fp32 kernel
    if (kernel == 3 && !conv_wgrad_thin_eligible(*p)) kernel = 0;
    if (kernel == 0) {
        // split-accumulate (parity) mode: the TMA kernel drains its accumulation chains into registers, so it needs no extra
        // split-K to bound the truncation bias; the cp.async kernel does (cheap only when dW is small, i.e. thin layers).
        // Plain TF32: the cp.async kernel is as fast or faster everywhere (profiles/r02_wgrad_tma_check.txt).
        const bool split = p->in_lo != nullptr || p->dout_lo != nullptr;
        // (measured per layer, profiles/r02_layers_tf32x3_wgrad.txt: from 64 output channels up the TMA kernel wins, below
        // that the cp.async kernel with dout / lo(dout) stacked on its N side)
        kernel = (split && p->Cout >= 64) ? 2 : 1;
        // the 16-output-channel decoder layers are bound by the MMA instruction count (K = 8 pixels per tcgen05.mma, 16 of 128 rows
        // used) and in split mode read four tensors: the fp32 FMA kernel reads two and is exact per product (conv_wgrad_thin.cu)
        if (split && conv_wgrad_thin_eligible(*p)) kernel = 3;
    }
    if (kernel == 3) {
        rc = launch_conv_wgrad_thin(*p, st);
    } else if (kernel == 2 && p->pad_mode == PADMODE_ZERO && conv_wgrad_tma_eligible(*p)) {
        rc = launch_conv_wgrad_tma(*p, st);
    } else if (kernel == 2 && p->pad_mode == PADMODE_REFLECT && p->pad == 1 && p->Ho >= 3 && p->Wo >= 3 &&
               p->Ho * p->Wo >= 64 * 208 && conv_wgrad_tma_eligible(zp)) {
        // reflection padding: zero-padded pass + the contributions of the taps that leave the image (border pixels only)
        rc = launch_conv_wgrad_tma(zp, st);
        if (rc == SCSFM_OK) {
            if (p->Cout <= 32) rc = launch_wgrad_tc<32, true>(*p, st);
            else if (p->Cout <= 64) rc = launch_wgrad_tc<64, true>(*p, st);
            else rc = launch_wgrad_tc<128, true>(*p, st);
        }
    } else if (p->in_lo != nullptr && p->dout_lo != nullptr && p->Cout <= 16) rc = launch_wgrad_tc<32, false, true>(*p, st);
    else if (p->in_lo != nullptr && p->dout_lo != nullptr && p->Cout <= 32) rc = launch_wgrad_tc<64, false, true>(*p, st);
    else if (p->in_lo != nullptr && p->dout_lo != nullptr && p->Cout <= 64) rc = launch_wgrad_tc<128, false, true>(*p, st);
    else if (p->Cout <= 32) rc = launch_wgrad_tc<32>(*p, st);
    else if (p->Cout <= 64) rc = launch_wgrad_tc<64>(*p, st);
    else rc = launch_wgrad_tc<128>(*p, st);
    if (rc) return rc;
    if (p->dbias) return launch_bias_grad(p->dout, p->B * p->Ho * p->Wo, p->Cout, p->dbias, st);
    return SCSFM_OK;
}
