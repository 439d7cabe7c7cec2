// Weight gradient of a stride-1, zero-padded convolution with both operands delivered by TMA (validated on B200 against
// the cp.async kernel and fp64: profiles/r02_wgrad_tma_check.txt; default for Cout >= 64 in split mode).
//
//   D[o (M = 128 TMEM lanes: output channels), (dy, chunk c, ch) (N = kh * G * 32 columns)] +=
//        sum over the pixels p of a tile   dout[p, o] * in[p + (dy, dx) - pad, 32 * (G * cg + c) + ch]
//
//  * pixels are the K dimension and both operands are "MN-major" (channels contiguous), which for 32-bit types means the
//    UMMA layout SWIZZLE_128B_BASE32B; CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B is its TMA twin (rows of 128 B = 32 channels of
//    one pixel, 32-byte chunks XOR-swizzled with the pixel index mod 4);
//  * a CTA owns one (Cout tile, group of G <= 2 channel chunks, dx) slice of dW and a range of pixel tiles (TH x TW = 64
//    pixels of one image, TW in {8, 16}); a stage holds an M part (the dout tile, one TMA box per 32 output channels) and an
//    N part (the input rows y0-pad .. y0-pad+TH+kh-2 shifted by dx, laid out [row][chunk][x][32 ch], so the (dy, chunk)
//    operand atoms of the tile row r start at ((r + dy) * G + c) * TW * 128 bytes: one uniform stride TW*128 = the
//    descriptor's LBO; the kh vertical taps share one load, as in conv_tma.cu);
//  * one tcgen05.mma (K = 8 pixels) per 8-pixel slice of a tile row: 8 per tile and pass, M = 128, N = kh*G*32 <= 192;
//  * split mode (ScsfmConv.in_lo / dout_lo): a tile takes TWO stages and its passes share operands ACROSS them, so every
//    operand crosses L2 -> shared memory once per tile instead of once per pass (the round-2 measurement: with one full stage
//    per pass the kernel asked for 46 B/cycle/SM, above the chip-wide L2 cap of ~42):
//        stage A = (dout,    lo(x))        pass 1: M = A, N = A      lo(x) . dout
//        stage B = (lo(dout), x   )        pass 2: M = B, N = B      x . lo(dout)
//                                          pass 3: M = A, N = B      x . dout         (the big term last)
//    With Cout <= 64 dout sits in atoms 0-1 and lo(dout) in atoms 2-3 of stage A's M part (both red.add into dw), stage B
//    carries only x:  pass 1: M = A, N = A;  pass 2: M = A, N = B  -- two MMAs per K8 slice instead of three;
//  * the tensor core adds into TMEM with truncation (bias ~3e-8 per MMA of a chain): in split mode an accumulation chain is
//    2-3 tiles (48 MMAs), drained by the epilogue warps into registers (round-to-nearest adds) while the MMAs of the next
//    chain run in the second TMEM buffer (one tile per chain starves the MMAs: a drain takes about as long as 16 MMAs);
//    plain TF32: one chain per CTA;
//  * split-K over pixel-tile ranges (gridDim.z, wave-aware count), fp32 vector red.add of the partial dW tiles.
// Reflection-padded layers (beyond the zero-padded pass), stride 2 and kernels larger than 3x3 stay on the cp.async kernel.
#include "conv_tc.cuh"

namespace scsfm {

constexpr int WT_EWARPS = 8;
constexpr int WT_THREADS = (WT_EWARPS + 2) * 32;
constexpr int WT_MAX_STAGES = 6;
constexpr int WT_PIX = 64;                           // pixels per tile
constexpr int WT_ATOM_BYTES = WT_PIX * 128;          // one 32-channel atom of the dout tile
constexpr int WT_M_BYTES = 4 * WT_ATOM_BYTES;        // M part of a stage: four atoms (atoms beyond the loaded ones stay stale)
constexpr int WT_SMEM_MAX = 232448;

struct WtGeom {
    int tw_log2;                 // TW = 1 << tw_log2 (3 or 4), TH = 64 >> tw_log2
    int tiles_x, tiles_y;        // pixel tiles per image
    int tiles_total;             // B * tiles_y * tiles_x
    int tiles_per_split;
    int groups, G;               // channel-chunk groups of G chunks (the last group may hold fewer real chunks)
    int atoms_m;                 // 32-channel atoms of the Cout tile that are loaded (1..4; stacked: 1..2 for each of dout / lo(dout))
    int stages, stage_bytes;
    int split;                   // 1: split-accumulate (two stages per tile, see above); 0: plain TF32 (one stage, one pass)
    int stack;                   // split mode with Cout <= 64: dout / lo(dout) stacked in the M part
    int tpc;                     // tiles per accumulation chain (split mode: 2-3 tiles = 48 MMAs; plain TF32: the whole CTA)
};

__global__ void __launch_bounds__(WT_THREADS, 1)
conv_wgrad_tma_kernel(ScsfmConv p, WtGeom g, const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap dmap,
                      const __grid_constant__ CUtensorMap xmap_lo, const __grid_constant__ CUtensorMap dmap_lo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar_empty = bar_full + WT_MAX_STAGES;
    uint64_t* acc_full = bar_empty + WT_MAX_STAGES;      // [2] tcgen05.commit at the end of a chain
    uint64_t* acc_empty = acc_full + 2;                  // [2] one arrival per epilogue warp once the buffer is drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint8_t* ring = smem + 1024;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TW = 1 << g.tw_log2, TH = WT_PIX >> g.tw_log2;
    const int cg = blockIdx.x / p.kw, dx = blockIdx.x - cg * p.kw;        // channel-chunk group, horizontal tap
    const int n0 = blockIdx.y * TBM;                                     // first output channel of this CTA
    const int t_begin = blockIdx.z * g.tiles_per_split, t_end = min(g.tiles_total, t_begin + g.tiles_per_split);
    const int ntiles = t_end - t_begin;
    if (ntiles <= 0) return;
    const int chunk0 = cg * g.G;                                         // first 32-channel chunk of the group
    const int natoms = p.kh * g.G;                                       // N atoms of the accumulator
    const int NCOLS = natoms * 32;
    const int spt = g.split ? 2 : 1;                                     // stages per tile
    const int tpc = g.split ? g.tpc : ntiles;                            // tiles per accumulation chain
    const int nchains = (ntiles + tpc - 1) / tpc;
    // optional per-CTA cycle counters (ScsfmConv.debug, 8 per CTA): [0] producer waiting for a free stage, [1] producer total,
    // [2] MMA thread waiting for operands, [3] MMA thread waiting for a drained accumulator, [4] MMA thread total,
    // [5] epilogue warp 0 waiting for a chain, [6] epilogue total, [7] tiles
    unsigned long long* dbg = p.debug ? p.debug + 8 * (size_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;

    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            tc::mbar_init(bar_full + s, 1);
            tc::mbar_init(bar_empty + s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(acc_full + i, 1);
            tc::mbar_init(acc_empty + i, WT_EWARPS);
        }
        tc::fence_barrier_init();
    }
    if (warp == WT_EWARPS + 1) tc::tmem_alloc(tmem_slot, 512);       // two accumulator buffers of 256 columns
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t ring_base = tc::smem_u32(ring);

    if (warp == WT_EWARPS) {
        // ------------------------------------------------------------------ TMA producer (the whole warp)
        // Measured with the per-role counters (profiles/r02_wgrad_roles.txt): ONE thread issuing the 14-28 TMA boxes of a stage
        // spends ~190 cycles per cp.async.bulk.tensor and starves the MMA thread (41 % of its time waiting for operands).  So
        // lane 0 waits for the free stage and posts the expected byte count, then the 32 lanes issue the boxes in parallel.
        if (lane == 0) {
            tc::tma_prefetch_desc(&xmap);
            tc::tma_prefetch_desc(&dmap);
            if (g.split) {
                tc::tma_prefetch_desc(&xmap_lo);
                tc::tma_prefetch_desc(&dmap_lo);
            }
        }
        const int prows = TH + p.kh - 1;
        const uint32_t row_bytes = (uint32_t)(TW * 128);
        const uint32_t n_bytes = (uint32_t)(prows * g.G) * row_bytes;
        const uint32_t m_bytes = (uint32_t)(g.atoms_m * WT_ATOM_BYTES);
        int s = 0;
        uint32_t ph = 0;
        long long tw = 0;
        const long long tb = clock64();
        for (int t = t_begin; t < t_end; ++t) {
            int q = t;
            const int tx = q % g.tiles_x; q /= g.tiles_x;
            const int ty = q % g.tiles_y;
            const int b = q / g.tiles_y;
            const int y0 = ty * TH, x0 = tx * TW;
            for (int j = 0; j < spt; ++j) {
                // stage j of the tile: which tensors feed its M part (dout side) and its N part (input side)
                const CUtensorMap* xm = (g.split && j == 0) ? &xmap_lo : &xmap;       // A: lo(x);  B (or plain): x
                const bool load_m = !(g.stack && j == 1);                             // stacked: stage B has no M part
                const int m_ops = load_m ? (g.stack ? 2 * g.atoms_m : g.atoms_m) : 0;
                const uint32_t st = ring_base + (uint32_t)(s * g.stage_bytes);
                if (lane == 0) {
                    const long long t0 = dbg ? clock64() : 0;
                    tc::mbar_wait(bar_empty + s, ph ^ 1);
                    if (dbg) tw += clock64() - t0;
                    tc::mbar_arrive_expect_tx(bar_full + s, (uint32_t)m_ops * WT_ATOM_BYTES + n_bytes);
                }
                __syncwarp();
                const int n_ops = prows * g.G;
                for (int k = lane; k < m_ops + n_ops; k += 32) {
                    if (k < m_ops) {
                        // M part: dout tile, one box (32 channels x TW x TH) per atom -> [atom][pixel][32]
                        int a, slot;
                        const CUtensorMap* dm;
                        if (g.stack) { a = k >> 1; slot = (k & 1) ? 2 + a : a; dm = (k & 1) ? &dmap_lo : &dmap; }
                        else { a = k; slot = k; dm = (g.split && j == 1) ? &dmap_lo : &dmap; }      // A (or plain): dout;  B: lo(dout)
                        tc::tma_load_4d(st + (uint32_t)(slot * WT_ATOM_BYTES), dm, n0 + 32 * a, x0, y0, b, bar_full + s);
                    } else {
                        // N part: input rows shifted by dx, [row][chunk][x][32]; channels / pixels outside the tensor are zero-filled
                        const int rc = k - m_ops, r = rc / g.G, c = rc - r * g.G;
                        tc::tma_load_4d(st + WT_M_BYTES + (uint32_t)rc * row_bytes, xm, 32 * (chunk0 + c), x0 - p.pad + dx, y0 - p.pad + r, b,
                                        bar_full + s);
                    }
                }
                if (++s == g.stages) { s = 0; ph ^= 1; }
            }
        }
        (void)m_bytes;
        if (dbg && lane == 0) { dbg[0] = (unsigned long long)tw; dbg[1] = (unsigned long long)(clock64() - tb); }
        __syncwarp();
    } else if (warp == WT_EWARPS + 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_tf32(TBM, NCOLS, 1, 1);               // both operands MN-major
            // constant descriptor fields; the start address (>> 4) is added per MMA
            const uint64_t dm0 = tc::make_smem_desc(0, WT_ATOM_BYTES, 512, tc::LAYOUT_SW128_BASE32B);         // M atoms one dout atom apart
            const uint64_t dn0 = tc::make_smem_desc(0, (uint32_t)(TW * 128), 512, tc::LAYOUT_SW128_BASE32B);   // N atoms one (row, chunk) block apart
            const int slices = TW >> 3;                       // 8-pixel slices per tile row
            int s = 0;
            uint32_t ph = 0;
            int jc = 0;                                       // chains issued: TMEM buffer jc & 1
            long long tfull = 0, tacc = 0;
            const long long tb = clock64();
            for (int t = 0; t < ntiles; ++t) {
                const bool first_of_chain = t % tpc == 0;
                const uint32_t acc = tmem_base + (uint32_t)((jc & 1) * 256);
                if (first_of_chain) {
                    const long long t0 = dbg ? clock64() : 0;
                    tc::mbar_wait(acc_empty + (jc & 1), ((jc >> 1) & 1) ^ 1);      // drained by the epilogue warps
                    if (dbg) tacc += clock64() - t0;
                    tc::fence_after_thread_sync();
                }
                const int sA = s, sB = s + 1;                 // (sB only in split mode: stages is even, so it never wraps)
                const uint32_t aM = ring_base + (uint32_t)(sA * g.stage_bytes), aN = aM + WT_M_BYTES;
                const uint32_t bM = ring_base + (uint32_t)(sB * g.stage_bytes), bN = bM + WT_M_BYTES;
                // passes of the tile as (M part address, N part address); the big x . dout term comes last
                const int npass = g.split ? (g.stack ? 2 : 3) : 1;
                for (int ps = 0; ps < npass; ++ps) {
                    uint32_t m_addr, n_addr;
                    const long long t0 = dbg ? clock64() : 0;
                    if (ps == 0) {                            // A x A
                        tc::mbar_wait(bar_full + sA, ph);
                        if (dbg) tfull += clock64() - t0;
                        tc::fence_after_thread_sync();
                        m_addr = aM; n_addr = aN;
                    } else if (ps == 1) {
                        tc::mbar_wait(bar_full + sB, ph);
                        if (dbg) tfull += clock64() - t0;
                        tc::fence_after_thread_sync();
                        m_addr = g.stack ? aM : bM; n_addr = bN;      // stacked: [dout; lo(dout)] . x   |   lo(dout) . x
                    } else {                                  // A's dout . B's x
                        m_addr = aM; n_addr = bN;
                    }
                    for (int r = 0; r < TH; ++r) {
                        for (int kq = 0; kq < slices; ++kq) {
                            const uint64_t dm = dm0 + (uint64_t)((m_addr + (uint32_t)((r * TW + kq * 8) * 128)) >> 4);
                            const uint64_t dn = dn0 + (uint64_t)((n_addr + (uint32_t)((r * g.G * TW + kq * 8) * 128)) >> 4);
                            tc::mma_tf32(acc, dm, dn, idesc, (!first_of_chain || ps != 0 || r != 0 || kq != 0) ? 1u : 0u);
                        }
                    }
                }
                tc::mma_commit(bar_empty + sA);               // both stages are free once every MMA of the tile has read them
                if (g.split) tc::mma_commit(bar_empty + sB);
                s += spt;
                if (s >= g.stages) { s = 0; ph ^= 1; }
                if ((t + 1) % tpc == 0 || t == ntiles - 1) {
                    tc::mma_commit(acc_full + (jc & 1));
                    ++jc;
                }
            }
            if (dbg) {
                dbg[2] = (unsigned long long)tfull; dbg[3] = (unsigned long long)tacc; dbg[4] = (unsigned long long)(clock64() - tb);
                dbg[7] = (unsigned long long)ntiles;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: dw[o][(dy, dx, channel)] += D[o][...]
        const int quarter = warp & 3, half = warp >> 2;
        const int oq = g.stack ? (quarter & 1) : quarter; // stacked: lane quarters 2, 3 hold the lo(dout) products of channels 0..63
        const int o = n0 + oq * 32 + lane;
        const int Mtot = p.kh * p.kw * p.Cin;
        const bool real = n0 + oq * 32 < p.Cout;          // warp-uniform: this lane quarter holds real output channels
        constexpr int APW = 3;                            // accumulator atoms (32 columns) per warp: a = half + 2 * ai < kh * G <= 6
        float accr[APW][32];
        long long tw = 0;
        const long long tb = clock64();
        for (int c = 0; c < nchains; ++c) {
            const int buf = c & 1;
            const long long t0 = dbg ? clock64() : 0;
            tc::mbar_wait(acc_full + buf, (c >> 1) & 1);
            if (dbg) tw += clock64() - t0;
            tc::fence_after_thread_sync();
            if (real) {
                const uint32_t src = tmem_base + (uint32_t)(buf * 256) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
                for (int ai = 0; ai < APW; ++ai) {
                    const int a = half + 2 * ai;
                    if (a < natoms) {
                        uint32_t r[32];
                        tc::tmem_ld32(src + (uint32_t)(a * 32), r);
                        tc::tmem_ld_wait();
                        if (c == 0) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) accr[ai][i] = __uint_as_float(r[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) accr[ai][i] += __uint_as_float(r[i]);
                        }
                    }
                }
            }
            tc::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_empty + buf);
        }
        if (dbg && tid == 0) { dbg[5] = (unsigned long long)tw; dbg[6] = (unsigned long long)(clock64() - tb); }
        if (real) {
#pragma unroll
            for (int ai = 0; ai < APW; ++ai) {
                const int a = half + 2 * ai;
                if (a >= natoms) continue;
                const int dy = a / g.G, c = a - dy * g.G;
                const int ch0 = 32 * (chunk0 + c);            // first input channel of this atom
                if (ch0 >= p.Cin) continue;                   // padding chunk of the last group
                if (o < p.Cout) {
                    float* dst = p.dw + (size_t)o * Mtot + (size_t)(dy * p.kw + dx) * p.Cin + ch0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (ch0 + j < p.Cin)                  // Cin % 4 == 0: a vector never straddles the end
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(accr[ai][j]), "f"(accr[ai][j + 1]),
                                         "f"(accr[ai][j + 2]), "f"(accr[ai][j + 3])
                                         : "memory");
                }
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == WT_EWARPS + 1) tc::tmem_dealloc(tmem_base, 512);
}

bool conv_wgrad_tma_eligible(const ScsfmConv& p) {
    return p.stride == 1 && p.pad_mode == PADMODE_ZERO && p.kh >= 1 && p.kh <= 3 && p.kw >= 1 && p.kw <= 3 && (p.Cin & 3) == 0 &&
           (p.Cout & 3) == 0 && p.Ho == p.Hi + 2 * p.pad - p.kh + 1 && p.Wo == p.Wi + 2 * p.pad - p.kw + 1 &&
           ((p.in_lo != nullptr) == (p.dout_lo != nullptr));      // split mode needs both low parts
}

int launch_conv_wgrad_tma(const ScsfmConv& p, cudaStream_t st) {
    static const cudaError_t attr_rc = cudaFuncSetAttribute(conv_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_MAX);
    SCSFM_CHECK_CUDA(attr_rc);
    WtGeom g;
    // tile shape (64 pixels): least padded area
    const long a16 = (long)((p.Ho + 3) / 4 * 4) * ((p.Wo + 15) / 16 * 16), a8 = (long)((p.Ho + 7) / 8 * 8) * ((p.Wo + 7) / 8 * 8);
    g.tw_log2 = a8 < a16 ? 3 : 4;
    const int TW = 1 << g.tw_log2, TH = WT_PIX >> g.tw_log2;
    g.tiles_x = (p.Wo + TW - 1) / TW;
    g.tiles_y = (p.Ho + TH - 1) / TH;
    g.tiles_total = g.tiles_x * g.tiles_y * p.B;
    const int chunks = (p.Cin + 31) / 32;
    g.G = chunks >= 2 ? 2 : 1;
    g.groups = (chunks + g.G - 1) / g.G;
    const int nt = (p.Cout + TBM - 1) / TBM;
    const int cout_tile = p.Cout < TBM ? p.Cout : TBM;             // (the last Cout tile may need fewer atoms; extra rows are zero-filled)
    g.split = (p.in_lo != nullptr && p.dout_lo != nullptr) ? 1 : 0;
    g.stack = (g.split && p.Cout <= 64) ? 1 : 0;
    g.tpc = g.stack ? 3 : 2;
    g.atoms_m = (cout_tile + 31) / 32;
    const int patch_bytes = (TH + p.kh - 1) * g.G * TW * 128;
    g.stage_bytes = WT_M_BYTES + (patch_bytes + 1023) / 1024 * 1024;
    g.stages = (WT_SMEM_MAX - 2048) / g.stage_bytes;
    if (g.stages > WT_MAX_STAGES) g.stages = WT_MAX_STAGES;
    g.stages &= ~1;                                                // even: a split-mode tile takes two consecutive stages
    if (g.stages < 2) {
        set_error("conv_wgrad_tma: stage of %d bytes does not fit twice in shared memory", g.stage_bytes);
        return SCSFM_ERR_ARG;
    }
    // Split the pixel tiles over gridDim.z.  One CTA per SM is resident (shared memory), so the kernel takes
    //   waves * (tiles per CTA + fixed cost per CTA)   with waves = ceil(slices * splits / SMs):
    // pick the split count that minimises it (an "about two waves" rule gave e.g. 297 CTAs = 2.007 waves on 148 SMs,
    // i.e. three waves of 13 tiles where one wave of 26 would do).  The fixed cost (prologue + red.add epilogue of a
    // 128 x 192 partial tile) is worth about 6 of the 64-pixel tiles; at least 4 tiles per CTA.
    const int slices = g.groups * p.kw * nt;
    int nsm = 148;
    {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) nsm = v;
    }
    int splits = 1;
    {
        const int max_splits = g.tiles_total >= 4 ? g.tiles_total / 4 : 1;
        long best = -1;
        for (int sp = 1; sp <= max_splits && sp <= 4096; ++sp) {
            const int tps = (g.tiles_total + sp - 1) / sp;
            const int real = (g.tiles_total + tps - 1) / tps;
            const long waves = ((long)slices * real + nsm - 1) / nsm;
            const long cost = waves * (tps + 6);
            if (best < 0 || cost < best) { best = cost; splits = real; }
        }
    }
    g.tiles_per_split = (g.tiles_total + splits - 1) / splits;
    splits = (g.tiles_total + g.tiles_per_split - 1) / g.tiles_per_split;
    CUtensorMap xmap, dmap, xmap_lo, dmap_lo;
    for (int lo = 0; lo < 2; ++lo) {
        const float* base = lo ? p.in_lo : p.in;
        CUtensorMap& xmap_ = lo ? xmap_lo : xmap;
        if (base == nullptr) { xmap_lo = xmap; continue; }
        const cuuint64_t gdim[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.B};
        const cuuint64_t gstride[3] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Wi * p.Cin * 4, (cuuint64_t)p.Hi * p.Wi * p.Cin * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)TW, 1, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = encode_tiled(&xmap_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(wgrad input %d x %d x %d x %d) failed with CUresult %d", p.B, p.Hi, p.Wi, p.Cin, (int)r);
            return SCSFM_ERR_CUDA;
        }
    }
    for (int lo = 0; lo < 2; ++lo) {
        const float* base = lo ? p.dout_lo : p.dout;
        CUtensorMap& dmap_ = lo ? dmap_lo : dmap;
        if (base == nullptr) { dmap_lo = dmap; continue; }
        const cuuint64_t gdim[4] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.B};
        const cuuint64_t gstride[3] = {(cuuint64_t)p.Cout * 4, (cuuint64_t)p.Wo * p.Cout * 4, (cuuint64_t)p.Ho * p.Wo * p.Cout * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)TW, (cuuint32_t)TH, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = encode_tiled(&dmap_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(wgrad dout %d x %d x %d x %d) failed with CUresult %d", p.B, p.Ho, p.Wo, p.Cout, (int)r);
            return SCSFM_ERR_CUDA;
        }
    }
    const size_t smem = 2048 + (size_t)g.stages * g.stage_bytes;
    dim3 grid(g.groups * p.kw, nt, splits);
    conv_wgrad_tma_kernel<<<grid, WT_THREADS, smem, st>>>(p, g, xmap, dmap, xmap_lo, dmap_lo);
    SCSFM_CHECK_LAUNCH();
    return SCSFM_OK;
}

}  // namespace scsfm
