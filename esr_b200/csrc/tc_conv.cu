// tc_conv.cu -- 3x3 / 1x1 convolution as an implicit GEMM on the 5th-gen tensor cores (sm_100a).
//
//   D[128 pixels, Cout] += A[128 pixels, 64 ch of one tap] * B[Cout, 64]^T      per K-block (tap x 64-channel chunk)
//
// * Activations live in HBM as "split bf16" NHWC tensors (value = hi + lo).  A K-block's A tile is ONE TMA box
//   (64 ch, TW, TH, 1 image, 1 plane) taken at the tap's (dy, dx) shift; out-of-image coordinates are zero-filled
//   by TMA, which is exactly the conv's zero padding.  The box lands in shared memory in the 128B-swizzled K-major
//   layout that tcgen05.mma consumes, so no thread touches the operands.
// * Channel concatenation (torch.cat in the reference) is a K-split over up to 3 source tensors.
// * fp32 parity: three bf16 MMAs per K-step (lo*hi + hi*lo + hi*hi) accumulate in fp32 in TMEM => ~2^-17 relative
//   operand error instead of bf16's 2^-9 (the reference network is fp32-only).
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld -> bias / residual / activation / GRU gating -> split-bf16 or fp32 NHWC stores).
// * mbarrier ring: full[s] (TMA -> MMA), empty[s] (tcgen05.commit -> TMA), accum_full (tcgen05.commit -> epilogue).
//
// Reference layers served: every Conv2d of models/model.py at feature resolution (Cin multiple of 64), the ConvGRU
// gates (models/submodules.py:496-514) and the DCNv2 contraction (models/DCNv2/src/cuda/dcn_v2_cuda.cu:90-92).
#include "tc_common.cuh"
#include <cstdlib>

namespace esr {


// ------------------------------------------------------------------------------------------------
// the kernel: one CTA = one tile of 128 output pixels (TH x TW) of one image, all output channels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 2) k_conv_tc(const __grid_constant__ ConvTCArgs a)
{
    PDL_LAUNCH_DEPENDENTS();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B needs 1024-B alignment
    const uint32_t b_bytes = (uint32_t)a.npad * 128u;
    const uint32_t stage_bytes = 2u * TC_A_BYTES + 2u * b_bytes;
    const uint32_t bar_base = smem_base + (uint32_t)a.stages * stage_bytes;
    // barriers: full[stages], empty[stages], accum_full, then the TMEM base address slot
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * a.stages, bar_accum = bar_base + 16u * a.stages;
    const uint32_t tmem_slot = bar_accum + 8u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < (a.stack ? 2 * a.npad : a.npad)) tmem_cols <<= 1;

    // tile -> (image, y0, x0)
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    const int img = blockIdx.x / tiles_per_img;
    const int trem = blockIdx.x - img * tiles_per_img;
    const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    PDL_WAIT();                      // everything above is CTA-local set-up; global memory only from here on

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one_sync()) {
            uint32_t s = 0, ph = 0;
            int src = 0, chunk_base = 0;
            for (int kb = 0; kb < a.nkb; ++kb) {
                const int gchunk = kb / a.ntaps, tap = kb - gchunk * a.ntaps;
                while (gchunk >= a.chunk_end[src]) { chunk_base = a.chunk_end[src]; ++src; }
                const int dy = a.ntaps == 9 ? tap / 3 - 1 : 0, dx = a.ntaps == 9 ? tap % 3 - 1 : 0;
                const int simg = a.src_img[src] ? a.src_img[src][img] : img;
                mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                mbar_expect_tx(bar_full + 8u * s, stage_bytes);
                const uint32_t st = smem_base + s * stage_bytes;
                const int c0 = (gchunk - chunk_base) * 64;
                tma_load_5d(&a.amap[src], bar_full + 8u * s, st, c0, x0 + dx, y0 + dy, simg, 0);
                tma_load_5d(&a.amap[src], bar_full + 8u * s, st + TC_A_BYTES, c0, x0 + dx, y0 + dy, simg, 1);
                tma_load_3d(&a.bmap, bar_full + 8u * s, st + 2u * TC_A_BYTES, 0, 0, kb);
                tma_load_3d(&a.bmap, bar_full + 8u * s, st + 2u * TC_A_BYTES + b_bytes, 0, 0, a.nkb + kb);
                if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc(TC_BLOCK_M, a.npad), idesc2 = umma_idesc(TC_BLOCK_M, 2 * a.npad);
            uint32_t s = 0, ph = 0;
            for (int kb = 0; kb < a.nkb; ++kb) {
                mbar_wait(bar_full + 8u * s, ph);
                tc_fence_after();
                const uint32_t st = smem_base + s * stage_bytes;
                const uint32_t a_hi = st, a_lo = st + TC_A_BYTES, b_hi = st + 2u * TC_A_BYTES, b_lo = b_hi + b_bytes;
                if (a.stack) {
                    // B_hi and B_lo are adjacent in the stage: ONE descriptor over 2 npad rows = [B_hi; B_lo]
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t dah = umma_desc(umma_desc_lo(a_hi) + 2u * k, UMMA_HI_1024), dal = umma_desc(umma_desc_lo(a_lo) + 2u * k, UMMA_HI_1024);
                        const uint64_t dbh = umma_desc(umma_desc_lo(b_hi) + 2u * k, UMMA_HI_1024);
                        umma_bf16(tmem_base, dah, dbh, idesc2, (kb | k) != 0 ? 1u : 0u);   // cols [0,npad) += A_hi B_hi, [npad,2npad) += A_hi B_lo
                        umma_bf16(tmem_base, dal, dbh, idesc, 1u);                          // cols [0,npad) += A_lo B_hi
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                       // 4 x (K = 16 bf16 = 32 bytes) per 128-byte row
                        const uint64_t dah = umma_desc(umma_desc_lo(a_hi) + 2u * k, UMMA_HI_1024), dal = umma_desc(umma_desc_lo(a_lo) + 2u * k, UMMA_HI_1024);
                        const uint64_t dbh = umma_desc(umma_desc_lo(b_hi) + 2u * k, UMMA_HI_1024), dbl = umma_desc(umma_desc_lo(b_lo) + 2u * k, UMMA_HI_1024);
                        umma_bf16(tmem_base, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);   // small terms first
                        umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                        umma_bf16(tmem_base, dah, dbh, idesc, 1u);
                    }
                }
                umma_commit(bar_empty + 8u * s);                    // frees the stage once these MMAs retire
                if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(bar_accum);                                 // accumulator complete
        }
    } else {
        // ===================== epilogue (4 warps, one TMEM lane quadrant each) =====================
        const int quad = warp & 3;                                  // tcgen05.ld: warp w may touch lanes 32*(w%4)..+31
        const int m = quad * 32 + lane;                             // accumulator row = pixel within the tile
        const int y = y0 + m / a.TW, x = x0 + m % a.TW;
        const bool valid = (y < a.H) && (x < a.W);
        const size_t pix = ((size_t)img * a.H + (valid ? y : 0)) * a.W + (valid ? x : 0);
        mbar_wait_backoff(bar_accum, 0);   // four warps idle for the whole main loop: poll with back-off
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int n0 = 0; n0 < a.npad; n0 += 32) {
            uint32_t raw[32];
            if (a.stack) tmem_ld_chunk_stacked(taddr, n0, a.npad, raw);
            else tmem_ld_chunk(taddr, n0, a.npad, raw);
            if (valid) epilogue_chunk(a, raw, n0, pix, img, y, x);
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant for the wide layers (N > 128: one stage ring fills most of the shared memory, so only ONE CTA fits per
// SM and k_conv_tc's epilogue -- 6-7 chunks of tcgen05.ld + bias/residual/activation + stores -- runs with the tensor core
// idle).  Here a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... with TWO accumulators in TMEM (2 x 256 columns):
// the MMA thread starts tile i+1 in the other accumulator while the epilogue warps drain tile i; the TMA producer simply
// streams K-blocks across tile boundaries.  Same operand order, same epilogue code: results are bit-identical to k_conv_tc.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t TCP_STG_PLANE = 32 * 64, TCP_STG_BUF = 2 * TCP_STG_PLANE;   // per-warp staging of the TMA-store epilogue (split outputs)
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tc_persist(const __grid_constant__ ConvTCArgs a)
{
    PDL_LAUNCH_DEPENDENTS();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)a.npad * 128u;
    const uint32_t stage_bytes = 2u * TC_A_BYTES + 2u * b_bytes;
    const uint32_t stg_ring = smem_base + (uint32_t)a.stages * stage_bytes;      // out_tma: [4 warps][2 buffers][2 planes][32 px x 64 B] (tc_conv_halo.cu)
    const uint32_t bar_base = stg_ring + (a.out_tma ? 4u * 2u * TCP_STG_BUF : 0u);
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * a.stages;
    const uint32_t bar_afull = bar_base + 16u * a.stages, bar_aempty = bar_afull + 16u;      // per accumulator
    const uint32_t tmem_slot = bar_aempty + 16u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = a.tiles_x * a.tiles_y, n_tiles = a.n_img * tiles_per_img;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_afull + 8u * i, 1); mbar_init(bar_aempty + 8u * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    PDL_WAIT();                      // everything above is CTA-local set-up; global memory only from here on

    if (warp == 0) {
        if (elect_one_sync()) {
            uint32_t s = 0, ph = 0;
            long long *tr = (a.trace && blockIdx.x == 0) ? a.trace : nullptr;
            int tn = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
                const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
                int src = 0, chunk_base = 0;
                for (int kb = 0; kb < a.nkb; ++kb) {
                    const int gchunk = kb / a.ntaps, tap = kb - gchunk * a.ntaps;
                    while (gchunk >= a.chunk_end[src]) { chunk_base = a.chunk_end[src]; ++src; }
                    const int dy = a.ntaps == 9 ? tap / 3 - 1 : 0, dx = a.ntaps == 9 ? tap % 3 - 1 : 0;
                    const int simg = a.src_img[src] ? a.src_img[src][img] : img;
                    mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                    if (tr && tn < TRACE_N) tr[tn * 8 + 0] = clock64();
                    mbar_expect_tx(bar_full + 8u * s, stage_bytes - ((a.diag & 1) ? b_bytes : 0u) - ((a.diag & 2) ? (uint32_t)TC_A_BYTES : 0u));
                    const uint32_t st = smem_base + s * stage_bytes;
                    const int c0 = (gchunk - chunk_base) * 64;
                    tma_load_5d(&a.amap[src], bar_full + 8u * s, st, c0, x0 + dx, y0 + dy, simg, 0);
                    if (!(a.diag & 2)) tma_load_5d(&a.amap[src], bar_full + 8u * s, st + TC_A_BYTES, c0, x0 + dx, y0 + dy, simg, 1);
                    tma_load_3d(&a.bmap, bar_full + 8u * s, st + 2u * TC_A_BYTES, 0, 0, kb);
                    if (!(a.diag & 1)) tma_load_3d(&a.bmap, bar_full + 8u * s, st + 2u * TC_A_BYTES + b_bytes, 0, 0, a.nkb + kb);
                    if (tr && tn < TRACE_N) { tr[tn * 8 + 1] = clock64(); ++tn; }
                    if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc(TC_BLOCK_M, a.npad), idesc2 = umma_idesc(TC_BLOCK_M, 2 * a.npad);
            uint32_t s = 0, ph = 0;
            int it = 0;
            long long *tr = (a.trace && blockIdx.x == 0) ? a.trace : nullptr;
            int tn = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t ai = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
                mbar_wait(bar_aempty + 8u * ai, aph ^ 1u);          // the epilogue of tile it-2 has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + ai * 256u;
                for (int kb = 0; kb < a.nkb; ++kb) {
                    if (tr && tn < TRACE_N) tr[tn * 8 + 2] = clock64();
                    mbar_wait(bar_full + 8u * s, ph);
                    tc_fence_after();
                    if (tr && tn < TRACE_N) tr[tn * 8 + 3] = clock64();
                    const uint32_t st = smem_base + s * stage_bytes;
                    const uint32_t a_hi = st, a_lo = st + TC_A_BYTES, b_hi = st + 2u * TC_A_BYTES, b_lo = b_hi + b_bytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t dah = umma_desc(umma_desc_lo(a_hi) + 2u * k, UMMA_HI_1024), dal = umma_desc(umma_desc_lo(a_lo) + 2u * k, UMMA_HI_1024);
                        const uint64_t dbh = umma_desc(umma_desc_lo(b_hi) + 2u * k, UMMA_HI_1024), dbl = umma_desc(umma_desc_lo(b_lo) + 2u * k, UMMA_HI_1024);
                        if (a.stack) {
                            umma_bf16(acc, dah, dbh, idesc2, (kb | k) != 0 ? 1u : 0u);   // [B_hi; B_lo] as one 2 npad-row operand
                            umma_bf16(acc, dal, dbh, idesc, 1u);
                        } else if (!(a.diag & 4)) {
                            umma_bf16(acc, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);
                            umma_bf16(acc, dah, dbl, idesc, 1u);
                            umma_bf16(acc, dah, dbh, idesc, 1u);
                        } else {
                            umma_bf16(acc, dah, dbh, idesc, (kb | k) != 0 ? 1u : 0u);   // measurement aid: one pass only
                        }
                    }
                    umma_commit(bar_empty + 8u * s);
                    if (tr && tn < TRACE_N) { tr[tn * 8 + 4] = clock64(); ++tn; }
                    if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(bar_afull + 8u * ai);
            }
        }
    } else {
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        int it = 0;
        uint32_t stg_n = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t ai = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
            const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
            const int y = y0 + m / a.TW, x = x0 + m % a.TW;
            const bool valid = (y < a.H) && (x < a.W);
            const size_t pix = ((size_t)img * a.H + (valid ? y : 0)) * a.W + (valid ? x : 0);
            mbar_wait_backoff(bar_afull + 8u * ai, aph);
            tc_fence_after();
            if (a.trace && blockIdx.x == 0 && threadIdx.x == 64 && it < TRACE_N) a.trace[it * 8 + 5] = clock64();
            const uint32_t taddr = tmem_base + ai * 256u + ((uint32_t)(quad * 32) << 16);
            for (int n0 = 0; n0 < a.npad; n0 += 32) {
                uint32_t raw[32];
                if (a.diag & 16) continue;                          // measurement aid: no TMEM reads, no stores
                if (a.stack) tmem_ld_chunk_stacked(taddr, n0, a.npad, raw);
                else tmem_ld_chunk(taddr, n0, a.npad, raw);
                if (a.out_tma == 1 && !(a.diag & 8)) {
                    // staged epilogue as in k_conv_tc_halo: this warp's 32 pixels x 32 channels -> swizzled shared memory -> TMA store
                    const uint32_t buf = stg_ring + ((uint32_t)quad * 2u + (stg_n & 1u)) * TCP_STG_BUF;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                    float v[32];
                    epilogue_values(a, raw, n0, img, y, x, valid, v);
                    uint32_t hw[16], lw[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) split_pack2(v[2 * e], v[2 * e + 1], hw[e], lw[e]);
                    const uint32_t row = buf + (uint32_t)lane * 64u, sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t o = row + (((uint32_t)q ^ sw) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o), "r"(hw[4 * q]), "r"(hw[4 * q + 1]), "r"(hw[4 * q + 2]), "r"(hw[4 * q + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + TCP_STG_PLANE), "r"(lw[4 * q]), "r"(lw[4 * q + 1]), "r"(lw[4 * q + 2]), "r"(lw[4 * q + 3]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const int rows = 32 / a.TW, yq = y0 + quad * rows;
                        tma_store_5d(&a.omap, buf, a.out_coff + n0, x0, yq, img, 0);
                        tma_store_5d(&a.omap, buf + TCP_STG_PLANE, a.out_coff + n0, x0, yq, img, 1);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++stg_n;
                } else
                if (valid && !(a.diag & 8)) epilogue_chunk(a, raw, n0, pix, img, y, x);
                else if (a.diag & 8) { if (raw[0] == 0x7fc12345u && raw[31] == 0x7fc54321u) a.out_f32[0] = 1.0f; }   // keep the loads alive
                __syncwarp();
            }
            tc_fence_before();                                       // this thread's TMEM reads are done: release the accumulator
            if (a.trace && blockIdx.x == 0 && threadIdx.x == 64 && it < TRACE_N) a.trace[it * 8 + 6] = clock64();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_aempty + 8u * ai) : "memory");
        }
        if (a.out_tma == 1 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the persistent kernel: clusters of two CTAs (the two SMs of a TPC) run ONE tcgen05.mma.cta_group::2 over
// two adjacent pixel tiles (M = 2 x 128).  Each CTA stages only its own A tile and HALF of the weight tile (rows
// [rank * npad/2, +npad/2)); the tensor cores of both SMs read both halves.  Per CTA a stage shrinks from 32 KB + 2 npad 128 B
// to 32 KB + npad 128 B, which is what buys pipeline depth: measured on B200 (profiles/r2_notes.md) the single-CTA kernel is
// bound by the TMA round trip (~2200 cycles) divided by the stages in flight -- 2 stages at N = 192, 4 at N = 64 -- not by
// bytes or MMA issue.  Here N = 192 runs 4 stages deep, N = 64 five.
// Protocol (S stages, two accumulators in TMEM as in k_conv_tc_persist):
//   * every CTA: TMA producer -> own full[s]; own empty[s] is signalled by the leader's tcgen05.commit (multicast to both CTAs);
//   * rank 1: one thread forwards "my stage s has landed" to the leader's pfull[s] (remote mbarrier arrive);
//   * rank 0 (leader): one thread issues the MMAs once full[s] and pfull[s] are complete; commits multicast to both CTAs'
//     empty[s] / afull[acc];
//   * epilogue warps of both CTAs drain their own 128 TMEM lanes and arrive (one lane per warp) on the LEADER's aempty[acc].
// Same operand order per output element as k_conv_tc / k_conv_tc_persist => bit-identical results.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) k_conv_tc_pair(const __grid_constant__ ConvTCArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bh_bytes = (uint32_t)a.npad * 64u;                 // one plane of this CTA's HALF weight tile (npad/2 rows)
    const uint32_t stage_bytes = 2u * TC_A_BYTES + 2u * bh_bytes;
    const uint32_t bar_base = smem_base + (uint32_t)a.stages * stage_bytes;
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * a.stages, bar_pfull = bar_base + 16u * a.stages;
    const uint32_t bar_afull = bar_base + 24u * a.stages, bar_aempty = bar_afull + 16u;
    const uint32_t tmem_slot = bar_aempty + 16u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = pair_rank();
    const int tiles_per_img = a.tiles_x * a.tiles_y, n_tiles = a.n_img * tiles_per_img;
    const int n_pairs = (n_tiles + 1) >> 1, pair0 = (int)(blockIdx.x >> 1), pair_step = (int)(gridDim.x >> 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); mbar_init(bar_pfull + 8u * s, 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_afull + 8u * i, 1); mbar_init(bar_aempty + 8u * i, 8); }   // 4 + 4 epilogue warps
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    pair_sync();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer (both CTAs): own A tile + own half of the weights =====================
        if (elect_one_sync()) {
            uint32_t s = 0, ph = 0;
            const int brow = (int)rank * (a.npad >> 1);
            long long *tr = (a.trace && blockIdx.x == 0) ? a.trace : nullptr;
            int tn = 0;
            for (int pr = pair0; pr < n_pairs; pr += pair_step) {
                const int tile = min(2 * pr + (int)rank, n_tiles - 1);
                const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
                const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
                int src = 0, chunk_base = 0;
                for (int kb = 0; kb < a.nkb; ++kb) {
                    const int gchunk = kb / a.ntaps, tap = kb - gchunk * a.ntaps;
                    while (gchunk >= a.chunk_end[src]) { chunk_base = a.chunk_end[src]; ++src; }
                    const int dy = a.ntaps == 9 ? tap / 3 - 1 : 0, dx = a.ntaps == 9 ? tap % 3 - 1 : 0;
                    const int simg = a.src_img[src] ? a.src_img[src][img] : img;
                    mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                    if (tr && tn < TRACE_N) tr[tn * 8 + 0] = clock64();
                    mbar_expect_tx(bar_full + 8u * s, stage_bytes);
                    const uint32_t st = smem_base + s * stage_bytes;
                    const int c0 = (gchunk - chunk_base) * 64;
                    tma_load_5d(&a.amap[src], bar_full + 8u * s, st, c0, x0 + dx, y0 + dy, simg, 0);
                    tma_load_5d(&a.amap[src], bar_full + 8u * s, st + TC_A_BYTES, c0, x0 + dx, y0 + dy, simg, 1);
                    tma_load_3d(&a.bmap_half, bar_full + 8u * s, st + 2u * TC_A_BYTES, 0, brow, kb);
                    tma_load_3d(&a.bmap_half, bar_full + 8u * s, st + 2u * TC_A_BYTES + bh_bytes, 0, brow, a.nkb + kb);
                    if (tr && tn < TRACE_N) { tr[tn * 8 + 1] = clock64(); ++tn; }
                    if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        const bool one = elect_one_sync();
        if (one && rank == 1) {
            // ===================== rank 1: forward "stage landed" to the leader =====================
            uint32_t s = 0, ph = 0;
            long long *tr = (a.trace && blockIdx.x == 1) ? a.trace : nullptr;
            int tn = 0;
            for (int pr = pair0; pr < n_pairs; pr += pair_step)
                for (int kb = 0; kb < a.nkb; ++kb) {
                    mbar_wait(bar_full + 8u * s, ph);
                    if (tr && tn < TRACE_N) { tr[tn * 8 + 6] = clock64(); ++tn; }
                    mbar_arrive_remote(bar_pfull + 8u * s, 0u);
                    if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
                }
        } else if (one) {
            // ===================== leader: MMA issuer for the pair =====================
            const uint32_t idesc = umma_idesc(2 * TC_BLOCK_M, a.npad);
            uint32_t s = 0, ph = 0;
            int it = 0;
            long long *tr = (a.trace && blockIdx.x == 0) ? a.trace : nullptr;
            int tn = 0;
            for (int pr = pair0; pr < n_pairs; pr += pair_step, ++it) {
                const uint32_t ai = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
                mbar_wait(bar_aempty + 8u * ai, aph ^ 1u);  // both CTAs' epilogues of pair it-2 have drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + ai * 256u;
                for (int kb = 0; kb < a.nkb; ++kb) {
                    if (tr && tn < TRACE_N) tr[tn * 8 + 2] = clock64();
                    mbar_wait(bar_full + 8u * s, ph);
                    if (tr && tn < TRACE_N) tr[tn * 8 + 3] = clock64();
                    mbar_wait(bar_pfull + 8u * s, ph);
                    tc_fence_after();
                    if (tr && tn < TRACE_N) tr[tn * 8 + 5] = clock64();
                    const uint32_t st = smem_base + s * stage_bytes;
                    const uint32_t a_hi = st, a_lo = st + TC_A_BYTES, b_hi = st + 2u * TC_A_BYTES, b_lo = b_hi + bh_bytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t dah = umma_desc(umma_desc_lo(a_hi) + 2u * k, UMMA_HI_1024), dal = umma_desc(umma_desc_lo(a_lo) + 2u * k, UMMA_HI_1024);
                        const uint64_t dbh = umma_desc(umma_desc_lo(b_hi) + 2u * k, UMMA_HI_1024), dbl = umma_desc(umma_desc_lo(b_lo) + 2u * k, UMMA_HI_1024);
                        umma_bf16_pair(acc, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_bf16_pair(acc, dah, dbl, idesc, 1u);
                        umma_bf16_pair(acc, dah, dbh, idesc, 1u);
                    }
                    umma_commit_pair(bar_empty + 8u * s);
                    if (tr && tn < TRACE_N) { tr[tn * 8 + 4] = clock64(); ++tn; }
                    if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
                }
                umma_commit_pair(bar_afull + 8u * ai);
            }
        }
    } else {
        // ===================== epilogue (both CTAs): own 128 accumulator rows = own pixel tile =====================
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        int it = 0;
        for (int pr = pair0; pr < n_pairs; pr += pair_step, ++it) {
            const uint32_t ai = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            const int tile_raw = 2 * pr + (int)rank;
            const int tile = min(tile_raw, n_tiles - 1);
            const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
            const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
            const int y = y0 + m / a.TW, x = x0 + m % a.TW;
            const bool valid = (tile_raw < n_tiles) && (y < a.H) && (x < a.W);
            const size_t pix = ((size_t)img * a.H + (valid ? y : 0)) * a.W + (valid ? x : 0);
            mbar_wait_backoff(bar_afull + 8u * ai, aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ai * 256u + ((uint32_t)(quad * 32) << 16);
            for (int n0 = 0; n0 < a.npad; n0 += 32) {
                uint32_t raw[32];
                if (a.npad - n0 >= 32) {
                    tmem_ld32(taddr + (uint32_t)n0, raw);
                } else {
                    uint32_t r16[16];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32"
                                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(r16[0]), "=r"(r16[1]), "=r"(r16[2]), "=r"(r16[3]), "=r"(r16[4]), "=r"(r16[5]),
                                   "=r"(r16[6]), "=r"(r16[7]), "=r"(r16[8]), "=r"(r16[9]), "=r"(r16[10]), "=r"(r16[11]),
                                   "=r"(r16[12]), "=r"(r16[13]), "=r"(r16[14]), "=r"(r16[15])
                                 : "r"(taddr + (uint32_t)n0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) { raw[j] = r16[j]; raw[16 + j] = 0u; }
                }
                if (valid) epilogue_chunk(a, raw, n0, pix, img, y, x);
                __syncwarp();
            }
            tc_fence_before();                                       // this warp's TMEM reads are done: release the accumulator
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(bar_aempty + 8u * ai, 0u);   // leader's barrier (for the leader: its own)
        }
    }

    tc_fence_before();
    pair_sync();                     // nobody leaves while the peer may still read this CTA's operands or arrive on its barriers
    if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps and launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode()
{
    static PFN_tmapEncodeTiled fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (PFN_tmapEncodeTiled)p;
    }();
    return fn;
}

static int make_amap(const SplitTensor &t, int BW, int BH, CUtensorMap *out)
{
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ESR_ECUDA; }
    const cuuint64_t C = t.C, W = t.W, H = t.H, N = t.n_img;
    cuuint64_t gdim[5] = {C, W, H, N, 2};
    cuuint64_t gstr[4] = {C * 2, W * C * 2, H * W * C * 2, N * H * W * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)BW, (cuuint32_t)BH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, t.base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A) failed: %d (C=%d W=%d H=%d N=%d)", (int)r, t.C, t.W, t.H, t.n_img); return ESR_ECUDA; }
    return ESR_OK;
}

static int make_bmap(const void *w, int npad, int nkb, int box_rows, CUtensorMap *out)
{
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ESR_ECUDA; }
    cuuint64_t gdim[3] = {64, (cuuint64_t)npad, (cuuint64_t)2 * nkb};
    cuuint64_t gstr[2] = {128, (cuuint64_t)npad * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(w), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d (npad=%d nkb=%d)", (int)r, npad, nkb); return ESR_ECUDA; }
    return ESR_OK;
}

// Output side: 32 channels x (BW x BH = 32 pixels) per box, 64-byte rows with the 64-byte swizzle (conflict-free 16-byte
// shared-memory stores by 32 lanes that own one pixel each).
int tc_make_omap(const SplitTensor &t, int BW, int BH, CUtensorMap *out)
{
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ESR_ECUDA; }
    const cuuint64_t C = t.C, W = t.W, H = t.H, N = t.n_img;
    cuuint64_t gdim[5] = {C, W, H, N, 2};
    cuuint64_t gstr[4] = {C * 2, W * C * 2, H * W * C * 2, (cuuint64_t)t.plane() * 2};
    cuuint32_t box[5] = {32, (cuuint32_t)BW, (cuuint32_t)BH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, t.base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out) failed: %d (C=%d W=%d H=%d N=%d)", (int)r, t.C, t.W, t.H, t.n_img); return ESR_ECUDA; }
    return ESR_OK;
}
int tc_make_omap_f32(float *base, int n_img, int H_, int W_, int C_, int BW, int BH, CUtensorMap *out)
{
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ESR_ECUDA; }
    const cuuint64_t C = C_, W = W_, H = H_, N = n_img;
    cuuint64_t gdim[5] = {C, W, H, N, 1};
    cuuint64_t gstr[4] = {C * 4, W * C * 4, H * W * C * 4, N * H * W * C * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)BW, (cuuint32_t)BH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out f32) failed: %d (C=%d W=%d H=%d N=%d)", (int)r, C_, W_, H_, n_img); return ESR_ECUDA; }
    return ESR_OK;
}
int tc_make_amap(const SplitTensor &t, int BW, int BH, CUtensorMap *out) { return make_amap(t, BW, BH, out); }
int tc_make_bmap(const void *w, int npad, int nkb, int box_rows, CUtensorMap *out) { return make_bmap(w, npad, nkb, box_rows, out); }

static size_t tc_smem_bytes(int npad, int stages)
{
    return 1024 + (size_t)stages * (2 * TC_A_BYTES + 2 * (size_t)npad * 128) + 16 * (size_t)stages + 64;
}

static size_t tc_pair_smem_bytes(int npad, int stages)
{
    return 1024 + (size_t)stages * (2 * TC_A_BYTES + (size_t)npad * 128) + 24 * (size_t)stages + 96;
}

int conv_tc_prepare(const ConvTCDesc &d, ConvTCArgs *args)
{
    ConvTCArgs &a = *args;
    memset(&a, 0, sizeof(a));
    ESR_REQUIRE(d.n_src >= 1 && d.n_src <= TC_MAX_SRC, "conv_tc: n_src=%d", d.n_src);
    ESR_REQUIRE(d.ntaps == 9 || d.ntaps == 1, "conv_tc: ntaps=%d", d.ntaps);
    const int H = d.src[0].H, W = d.src[0].W;
    int chunks = 0;
    // tile shape: 128 pixels; prefer wide tiles, but do not waste more than half a tile on narrow images
    int TW = W >= 24 ? 32 : (W >= 12 ? 16 : 8);
    int TH = TC_BLOCK_M / TW;
    // The halo-reuse + weight-multicast kernel (tc_conv3.cu: 8 x 16 tiles, 16 x 18 halo boxes) is opt-in (ESR_TC_V3=1):
    // measured on B200 it is 5-50 % SLOWER than this kernel with two co-resident CTAs per SM (profiles/r1_notes.md) --
    // the main loop turned out to be latency / shared-memory-port bound, not L2-bandwidth bound.
    static const bool use_v3 = getenv("ESR_TC_V3") != nullptr;
    const int npad_ = tc_npad(d.cout);
    int a_st = 0, b_st = 0;
    static const int v3_min_n = getenv("ESR_TC_V3_MIN_N") ? atoi(getenv("ESR_TC_V3_MIN_N")) : 0;   // restrict the experiment to wide layers
    const bool v3 = d.ntaps == 9 && use_v3 && npad_ >= v3_min_n && conv_tc3_plan(npad_, &a_st, &b_st);
    int BW = TW, BH = TH;
    if (v3) { TW = 8; TH = 16; BW = 16; BH = 18; }
    // multi-wave 3x3 layers: persistent halo-reuse kernel (tc_conv_halo.cu; ESR_TC_NO_HALO=1 keeps k_conv_tc_persist)
    static const bool no_halo = getenv("ESR_TC_NO_HALO") != nullptr;
    int h_as = 0, h_bs = 0;
    // epilogue through shared memory + TMA stores (ESR_TC_NO_OUT_TMA=1: direct 16-byte stores): plain split outputs only
    static const bool no_out_tma = getenv("ESR_TC_NO_OUT_TMA") != nullptr;
    int out_tma = 0, stg_bufs = 0;
    if (!no_out_tma && d.epi_mode == EPI_STD) {
        if (d.out.base && !d.out_f32 && d.cout % 32 == 0 && d.out.C % 8 == 0 && d.out_coff % 8 == 0) out_tma = 1;
        else if (!d.out.base && d.out_f32 && !d.out_f32_nchw && d.out_f32_C % 4 == 0 && ((uintptr_t)d.out_f32 & 15) == 0) out_tma = 2;
    }
    bool halo = !v3 && !no_halo && d.ntaps == 9 && getenv("ESR_TC_PAIR") == nullptr &&
                d.n_img * ((W + 7) / 8) * ((H + 15) / 16) > dev_info().sm_count;
    if (halo && out_tma) {
        if (conv_tc_halo_plan(npad_, 2, &h_as, &h_bs)) stg_bufs = 2;
        else if (conv_tc_halo_plan(npad_, 1, &h_as, &h_bs)) stg_bufs = 1;
        else out_tma = 0;
    }
    if (halo && !out_tma) halo = conv_tc_halo_plan(npad_, 0, &h_as, &h_bs);
    const int out_tma_req = out_tma;                              // eligibility of the output, whichever kernel takes the layer
    if (!halo) out_tma = 0;                                       // (decided below for the persistent kernel; the one-tile kernel keeps direct stores)
    if (halo) { TW = 8; TH = 16; BW = 10; BH = 18; }
    for (int s = 0; s < d.n_src; ++s) {
        const SplitTensor &t = d.src[s];
        ESR_REQUIRE(t.base && t.C % 64 == 0 && t.H == H && t.W == W, "conv_tc: source %d has C=%d H=%d W=%d", s, t.C, t.H, t.W);
        chunks += t.C / 64;
        a.chunk_end[s] = chunks;
        a.src_img[s] = d.src_img[s];
        int rc = make_amap(t, BW, BH, &a.amap[s]);
        if (rc) return rc;
    }
    for (int s = d.n_src; s < TC_MAX_SRC; ++s) a.chunk_end[s] = 1 << 30;
    a.n_src = d.n_src; a.ntaps = d.ntaps; a.nkb = chunks * d.ntaps;
    a.npad = tc_npad(d.cout); a.cout = d.cout;
    ESR_REQUIRE(a.npad <= 256, "conv_tc: cout=%d too large", d.cout);
    int rc = make_bmap(d.wpacked, a.npad, a.nkb, a.npad, &a.bmap);
    if (rc) return rc;
    a.kernel_ver = v3 ? 3 : (halo ? 4 : 1);
    // stacked weights: for N <= 128 the two products that share A_hi run as ONE MMA over [B_hi; B_lo] (N' = 2 N <= 256): 8 instead of 12
    // MMAs per K-block and A_hi is read from shared memory once instead of twice (the kernels are shared-memory-port bound)
    { static const bool no_stack = getenv("ESR_TC_NO_STACK") != nullptr; a.stack = (!v3 && !no_stack && a.npad <= 128) ? 1 : 0; }
    { static const int diag = getenv("ESR_TC_DIAG") ? atoi(getenv("ESR_TC_DIAG")) : 0; a.diag = diag; }
    a.cluster = 1; a.a_stages = a_st;
    a.H = H; a.W = W; a.TW = TW; a.TH = TH;
    a.tiles_x = (W + TW - 1) / TW; a.tiles_y = (H + TH - 1) / TH; a.n_img = d.n_img;
    // Pipeline depth.  Grids of more than one wave keep two CTAs resident per SM (<= half the shared memory each) so that
    // one CTA's prologue / epilogue overlaps the other's main loop; single-wave grids take all the stages that fit.
    const size_t smem_cap = (size_t)dev_info().max_smem_optin;
    const int n_tiles = d.n_img * a.tiles_x * a.tiles_y;
    int stages = 6;
    while (stages > 2 && tc_smem_bytes(a.npad, stages) > smem_cap) --stages;
    // multi-wave grids of layers with N >= 64: persistent CTAs (one per SM, all the stages that fit) with two TMEM accumulators
    static const int persist_min_n = getenv("ESR_TC_NO_PERSIST") ? 1 << 30 : (getenv("ESR_TC_PERSIST_MIN_N") ? atoi(getenv("ESR_TC_PERSIST_MIN_N")) : 64);   // measured: 129 -> 3.196 ms, 64 -> 3.159 ms, 16 -> 3.165 ms per cfg2 step
    a.persist = (!v3 && !halo && a.npad >= persist_min_n && n_tiles > dev_info().sm_count) ? 1 : 0;
    if (a.persist && out_tma_req == 1 && 32 % TW == 0 && getenv("ESR_TC_PAIR") == nullptr) {
        int st2 = stages;                                         // staged TMA-store epilogue also here: 32 KB of staging beside the ring
        while (st2 > 2 && tc_smem_bytes(a.npad, st2) + 64 + 4 * 2 * 4096 > smem_cap) --st2;
        if (tc_smem_bytes(a.npad, st2) + 64 + 4 * 2 * 4096 <= smem_cap) { stages = st2; out_tma = 1; stg_bufs = 2; }
    }
    if (n_tiles > dev_info().sm_count && !a.persist) {
        int s2 = stages;
        while (s2 > 2 && 2 * (tc_smem_bytes(a.npad, s2) + 1024) > smem_cap) --s2;
        if (2 * (tc_smem_bytes(a.npad, s2) + 1024) <= smem_cap) stages = s2;
    }
    { static const int cap = getenv("ESR_TC_STAGES") ? atoi(getenv("ESR_TC_STAGES")) : 0; if (cap >= 2 && stages > cap) stages = cap; }   // measurement aid
    if (stages > a.nkb) stages = a.nkb < 2 ? 2 : a.nkb;
    a.stages = stages;
    // CTA pairs (tcgen05 cta_group::2) for the persistent layers: half the weight bytes per CTA -> deeper pipeline (ESR_TC_NO_PAIR=1: off)
    static const bool no_pair = getenv("ESR_TC_PAIR") == nullptr;       // opt-in: measured slower than the single-CTA kernel (profiles/r2_notes.md)
    a.pair = 0;
    if (a.persist && !no_pair && n_tiles >= 2 && dev_info().sm_count >= 2) {
        int sp = 8;
        while (sp > 2 && tc_pair_smem_bytes(a.npad, sp) > smem_cap) --sp;
        if (tc_pair_smem_bytes(a.npad, sp) <= smem_cap) {
            if (sp > a.nkb) sp = a.nkb < 2 ? 2 : a.nkb;
            a.pair = 1; a.stages = sp;
            if ((rc = make_bmap(d.wpacked, a.npad, a.nkb, a.npad / 2, &a.bmap_half))) return rc;
        }
    }
    if (halo) { a.a_stages = h_as; a.stages = h_bs; }
    if (v3) {
        static const bool no_mc = getenv("ESR_TC_NO_MULTICAST") != nullptr;
        a.stages = b_st;
        a.cluster = (!no_mc && n_tiles >= 2) ? 2 : 1;
        if (a.cluster == 2 && (rc = make_bmap(d.wpacked, a.npad, a.nkb, a.npad / 2, &a.bmap_half))) return rc;
    }
    a.bias = d.bias;
    a.act = d.act; a.act_from = d.act_from; a.res_mode = d.res_mode; a.epi_mode = d.epi_mode;
    if (d.res_mode != RES_NONE) {
        ESR_REQUIRE(d.res.base && d.res.H == H && d.res.W == W && d.cout % 32 == 0 && d.res.C >= d.cout, "conv_tc: bad residual");
        a.res = d.res.base; a.res_plane = d.res.plane(); a.res_C = d.res.C; a.res_img = d.res_img;
    }
    if (d.out.base) {
        ESR_REQUIRE(d.out.H == H && d.out.W == W && d.out.n_img >= d.n_img && d.out.C % 8 == 0 && d.out_coff % 8 == 0,
                    "conv_tc: bad split output");
        a.out = d.out.base; a.out_plane = d.out.plane(); a.out_C = d.out.C; a.out_coff = d.out_coff;
        if (out_tma == 1) {
            a.out_tma = 1; a.stg_bufs = stg_bufs;
            if ((rc = tc_make_omap(d.out, TW, 32 / TW, &a.omap))) return rc;
        }
    }
    if (out_tma == 2) {
        a.out_tma = 2; a.stg_bufs = stg_bufs;
        if ((rc = tc_make_omap_f32(d.out_f32, d.n_img, H, W, d.out_f32_C, TW, 32 / TW, &a.omap))) return rc;
    }
    a.out_f32 = d.out_f32; a.out_f32_C = d.out_f32_C; a.out_f32_nchw = d.out_f32_nchw;
    if (d.epi_mode != EPI_STD) {
        ESR_REQUIRE(d.h_prev.base && d.h_prev.C == 64 && d.z_buf && d.out.base, "conv_tc: GRU epilogue needs h_prev, z_buf, out");
        ESR_REQUIRE((d.epi_mode == EPI_GRU_ZR && a.npad == 128) || (d.epi_mode == EPI_GRU_OUT && a.npad == 64), "conv_tc: GRU epilogue width");
        a.h_prev = d.h_prev.base; a.h_plane = d.h_prev.plane(); a.z_buf = d.z_buf;
    } else {
        ESR_REQUIRE(!d.out.base || d.cout % 32 == 0, "conv_tc: split output needs cout %% 32 == 0");
    }
    return ESR_OK;
}

int conv_tc_launch(const ConvTCArgs &a, cudaStream_t st)
{
    if (a.kernel_ver == 3) return conv_tc3_launch(a, st);
    if (a.kernel_ver == 4) return conv_tc_halo_launch(a, st);
    static int max_set = 0;
    const size_t smem = tc_smem_bytes(a.npad, a.stages);
    if (!(a.persist && a.pair) && (int)smem > max_set) {
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_set = (int)smem;
    }
    const unsigned grid = (unsigned)(a.n_img * a.tiles_x * a.tiles_y);
    // wide layers (one CTA per SM) on multi-wave grids: persistent CTAs with two TMEM accumulators (ESR_TC_NO_PERSIST=1: off)
    if (a.persist && a.pair) {
        static int max_set_q = 0;
        const size_t smem_q = tc_pair_smem_bytes(a.npad, a.stages);
        if ((int)smem_q > max_set_q) {
            ESR_CUDA_CHECK(cudaFuncSetAttribute(k_conv_tc_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
            max_set_q = (int)smem_q;
        }
        const int n_pairs = ((int)grid + 1) / 2;
        const unsigned g2 = 2u * (unsigned)min(n_pairs, dev_info().sm_count / 2);
        static const char *trace_path = getenv("ESR_TC_TRACE");          // measurement aid, see the persistent kernel below
        if (trace_path) {
            static long long *dbuf = nullptr;
            int want_n = 0, want_k = 0; char path[512] = {0};
            if (sscanf(trace_path, "%511[^:]:%d:%d", path, &want_n, &want_k) == 3 && want_n == a.npad && want_k == a.nkb) {
                if (!dbuf) cudaMalloc(&dbuf, sizeof(long long) * 8 * TRACE_N);
                cudaMemsetAsync(dbuf, 0, sizeof(long long) * 8 * TRACE_N, st);
                ConvTCArgs b = a; b.trace = dbuf;
                k_conv_tc_pair<<<g2, TC_THREADS, smem_q, st>>>(b);
                cudaStreamSynchronize(st);
                static long long host[8 * TRACE_N];
                cudaMemcpy(host, dbuf, sizeof(host), cudaMemcpyDeviceToHost);
                FILE *f = fopen(path, "w");
                if (f) {
                    fprintf(f, "# PAIR npad=%d nkb=%d stages=%d tiles=%d; per K-block: prod_after_empty_wait, prod_after_issue, mma_before_full_wait, mma_after_full_wait, mma_after_commit, mma_after_pfull_wait, relay_after_full_wait(rank1 clock)\n", a.npad, a.nkb, a.stages, (int)grid);
                    for (int i = 0; i < TRACE_N; ++i) { for (int c = 0; c < 7; ++c) fprintf(f, "%lld%c", host[i * 8 + c], c == 6 ? '\n' : ','); }
                    fclose(f);
                }
                ESR_LAUNCH_CHECK();
                return ESR_OK;
            }
        }
        k_conv_tc_pair<<<g2, TC_THREADS, smem_q, st>>>(a);
        ESR_LAUNCH_CHECK();
        return ESR_OK;
    }
    if (a.persist) {
        static const char *trace_path = getenv("ESR_TC_TRACE");          // measurement aid: "<file>:<npad>:<nkb>" traces launches of that shape
        if (trace_path) {
            static long long *dbuf = nullptr;
            int want_n = 0, want_k = 0; char path[512] = {0};
            if (sscanf(trace_path, "%511[^:]:%d:%d", path, &want_n, &want_k) == 3 && want_n == a.npad && want_k == a.nkb) {
                if (!dbuf) cudaMalloc(&dbuf, sizeof(long long) * 8 * TRACE_N);
                cudaMemsetAsync(dbuf, 0, sizeof(long long) * 8 * TRACE_N, st);
                ConvTCArgs b = a; b.trace = dbuf;
                static int max_set_t = 0;
                const size_t smem_t = smem + 64 + (a.out_tma ? 4 * 2 * 4096 : 0);
                if ((int)smem_t > max_set_t) { cudaFuncSetAttribute(k_conv_tc_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t); max_set_t = (int)smem_t; }
                k_conv_tc_persist<<<(unsigned)dev_info().sm_count, TC_THREADS, smem_t, st>>>(b);
                cudaStreamSynchronize(st);
                static long long host[8 * TRACE_N];
                cudaMemcpy(host, dbuf, sizeof(host), cudaMemcpyDeviceToHost);
                FILE *f = fopen(path, "w");
                if (f) {
                    fprintf(f, "# npad=%d nkb=%d stages=%d tiles=%d; per K-block: prod_after_empty_wait, prod_after_issue, mma_before_full_wait, mma_after_full_wait, mma_after_commit; per tile (same rows, by tile index): epi_start, epi_end\n", a.npad, a.nkb, a.stages, a.n_img * a.tiles_x * a.tiles_y);
                    for (int i = 0; i < TRACE_N; ++i) { for (int c = 0; c < 7; ++c) fprintf(f, "%lld%c", host[i * 8 + c], c == 6 ? '\n' : ','); }
                    fclose(f);
                }
                ESR_LAUNCH_CHECK();
                return ESR_OK;
            }
        }
        static int max_set_p = 0;
        const size_t smem_p = smem + 64 + (a.out_tma ? 4 * 2 * 4096 : 0);
        if ((int)smem_p > max_set_p) {
            ESR_CUDA_CHECK(cudaFuncSetAttribute(k_conv_tc_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
            max_set_p = (int)smem_p;
        }
        ESR_CUDA_CHECK(launch_pdl(k_conv_tc_persist, dim3((unsigned)dev_info().sm_count), dim3(TC_THREADS), smem_p, st, a));
        esr::count_launch();
        return ESR_OK;
    }
    ESR_CUDA_CHECK(launch_pdl(k_conv_tc, dim3(grid), dim3(TC_THREADS), smem, st, a));
    esr::count_launch();
    return ESR_OK;
}

// ------------------------------------------------------------------------------------------------
// weight packing and layout conversion kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_weight(const float *__restrict__ w0, const float *__restrict__ w1, int cout_each, int cin, int ksz,
                              int npad, int nkb, __nv_bfloat16 *__restrict__ dst)
{
    const int ntaps = ksz * ksz;
    const size_t total = (size_t)nkb * npad * 64;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % 64);
        const int n = (int)((i / 64) % npad);
        const int kb = (int)(i / ((size_t)64 * npad));
        const int chunk = kb / ntaps, tap = kb % ntaps;
        const int ci = chunk * 64 + c;
        float v = 0.0f;
        const int cout_tot = w1 ? 2 * cout_each : cout_each;
        if (n < cout_tot) {
            const float *w = (n < cout_each) ? w0 : w1;
            const int nn = n < cout_each ? n : n - cout_each;
            v = w[((size_t)nn * cin + ci) * ntaps + tap];
        }
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        dst[i] = hi;
        dst[total + i] = lo;
    }
}

int pack_conv_weight2(const float *w0, const float *w1, int cout_each, int cin, int ksz, void *dst, cudaStream_t st)
{
    ESR_REQUIRE(cin % 64 == 0 && (ksz == 1 || ksz == 3), "pack_conv_weight: cin=%d ksz=%d", cin, ksz);
    const int cout = w1 ? 2 * cout_each : cout_each;
    const int npad = tc_npad(cout), nkb = tc_nkb(cin, ksz * ksz);
    const size_t total = (size_t)nkb * npad * 64;
    k_pack_weight<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(w0, w1, cout_each, cin, ksz, npad, nkb,
                                                                             (__nv_bfloat16 *)dst);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}
int pack_conv_weight(const float *w, int cout, int cin, int ksz, void *dst, cudaStream_t st)
{
    return pack_conv_weight2(w, nullptr, cout, cin, ksz, dst, st);
}

__global__ void k_split_from_nchw(const float *__restrict__ src, int n_img, int C, int H, int W, __nv_bfloat16 *__restrict__ dst,
                                  size_t plane)
{
    const size_t count = (size_t)n_img * H * W * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t p = i / C;
        const int x = (int)(p % W), y = (int)((p / W) % H), n = (int)(p / ((size_t)W * H));
        __nv_bfloat16 hi, lo;
        split_bf16(src[(((size_t)n * C + c) * H + y) * W + x], hi, lo);
        dst[i] = hi;
        dst[plane + i] = lo;
    }
}
__global__ void k_split_to_nchw(const __nv_bfloat16 *__restrict__ src, size_t plane, int n_img, int C, int H, int W,
                                float *__restrict__ dst)
{
    const size_t count = (size_t)n_img * H * W * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const int c = (int)((i / ((size_t)W * H)) % C), n = (int)(i / ((size_t)W * H * C));
        const size_t s = (((size_t)n * H + y) * W + x) * C + c;
        dst[i] = join_bf16(src[s], src[plane + s]);
    }
}
int split_from_nchw_planes(const float *src, int n_img, int C, int H, int W, __nv_bfloat16 *dst, size_t plane, cudaStream_t st)
{
    const size_t count = (size_t)n_img * H * W * C;
    k_split_from_nchw<<<(unsigned)ceil_div64((int64_t)count, 256), 256, 0, st>>>(src, n_img, C, H, W, dst, plane);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}
int split_to_nchw_planes(const __nv_bfloat16 *src, size_t plane, int n_img, int C, int H, int W, float *dst, cudaStream_t st)
{
    const size_t count = (size_t)n_img * H * W * C;
    k_split_to_nchw<<<(unsigned)ceil_div64((int64_t)count, 256), 256, 0, st>>>(src, plane, n_img, C, H, W, dst);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}
int split_from_nchw(const float *src, int n_img, int C, int H, int W, __nv_bfloat16 *dst, cudaStream_t st)
{
    return split_from_nchw_planes(src, n_img, C, H, W, dst, (size_t)n_img * H * W * C, st);
}
int split_to_nchw(const __nv_bfloat16 *src, int n_img, int C, int H, int W, float *dst, cudaStream_t st)
{
    return split_to_nchw_planes(src, (size_t)n_img * H * W * C, n_img, C, H, W, dst, st);
}

} // namespace esr

// ------------------------------------------------------------------------------------------------
// C ABI (layer level)
// ------------------------------------------------------------------------------------------------
using namespace esr;

static SplitTensor mk_split(const void *p, int n_img, int H, int W, int C)
{
    SplitTensor t;
    t.base = (__nv_bfloat16 *)const_cast<void *>(p);
    t.n_img = n_img; t.H = H; t.W = W; t.C = C;
    return t;
}

extern "C" int esr_conv_tc(const esr_conv_desc *c, esr_stream_t stream)
{
    ESR_REQUIRE(c, "esr_conv_tc: null descriptor");
    ConvTCDesc d;
    d.n_src = c->n_src;
    for (int s = 0; s < c->n_src && s < TC_MAX_SRC; ++s) {
        d.src[s] = mk_split(c->src[s], c->src_n_img[s], c->H, c->W, c->src_C[s]);
        d.src_img[s] = c->src_img[s];
    }
    d.ntaps = c->ntaps; d.cout = c->cout; d.wpacked = c->wpacked; d.bias = c->bias; d.n_img = c->n_img;
    d.act = c->act; d.act_from = c->act_from; d.res_mode = c->res_mode; d.epi_mode = c->epi_mode;
    if (c->res) { d.res = mk_split(c->res, c->res_n_img, c->H, c->W, c->res_C); d.res_img = c->res_img; }
    if (c->out) { d.out = mk_split(c->out, c->out_n_img, c->H, c->W, c->out_C); d.out_coff = c->out_coff; }
    d.out_f32 = c->out_f32; d.out_f32_C = c->out_f32_C;
    if (c->h_prev) d.h_prev = mk_split(c->h_prev, c->h_n_img, c->H, c->W, 64);
    d.z_buf = c->z_buf;
    ConvTCArgs args;
    int rc = conv_tc_prepare(d, &args);
    if (rc) return rc;
    return conv_tc_launch(args, (cudaStream_t)stream);
}
extern "C" size_t esr_conv_weight_bytes(int cout, int cin, int ksz) { return tc_packed_weight_bytes(cout, cin, ksz * ksz); }
extern "C" int esr_pack_conv_weight(const float *w0, const float *w1, int cout_each, int cin, int ksz, void *dst, esr_stream_t st)
{
    return pack_conv_weight2(w0, w1, cout_each, cin, ksz, dst, (cudaStream_t)st);
}
extern "C" int esr_split_from_nchw(const float *src, int n_img, int C, int H, int W, void *dst, esr_stream_t st)
{
    return split_from_nchw(src, n_img, C, H, W, (__nv_bfloat16 *)dst, (cudaStream_t)st);
}
extern "C" int esr_split_to_nchw(const void *src, int n_img, int C, int H, int W, float *dst, esr_stream_t st)
{
    return split_to_nchw((const __nv_bfloat16 *)src, n_img, C, H, W, dst, (cudaStream_t)st);
}
