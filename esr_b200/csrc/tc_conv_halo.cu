// tc_conv_halo.cu -- persistent tcgen05 3x3 convolution with halo reuse: the kernel for the multi-wave 3x3 layers.
//
// Why (profiles/r2_notes.md, in-kernel clock traces): the tcgen05 conv kernels are bound by the 128 B/clk shared-memory port,
// through which BOTH the SS-mode MMA operand reads and the TMA fill of the next stages go.  k_conv_tc_persist moves, per
// K-block (one tap of one 64-channel chunk) and N = 64: 72 KB of MMA reads + 48 KB of TMA writes.  This kernel cuts both:
//   * A by halo reuse: per 64-channel chunk ONE TMA box of the (8+2) x (16+2) pixel halo tile per plane (2 x 23 KB) serves all
//     nine taps -- 5 KB of fill per tap instead of 32 KB.  The output tile is 8 pixels wide and 16 tall, halo rows are 10 pixels
//     (1280 B) apart, so the 128 A rows of tap (ty, tx) are a canonical 128B-swizzled K-major operand that starts at halo pixel
//     (ty, tx): 8-row groups SBO = 1280 B apart, start address advanced by (ty * 10 + tx) * 128 B.  tcgen05 applies the swizzle to
//     the absolute shared-memory address like TMA does when it writes the box (measured in round 1, tools/diag_halo.py), so the
//     shifted views need no correction.
//   * MMA reads by stacking [B_hi; B_lo] along N (ConvTCArgs::stack, N <= 128): A_hi x [B_hi; B_lo] + A_lo x B_hi = two MMAs per
//     K-step instead of three, A_hi read once.
// Weights stream through their own ring (one stage per tap: 2 x npad x 128 B), deep enough to cover the TMA round trip.
// Persistent CTAs (one per SM) with two TMEM accumulators as in k_conv_tc_persist; identical epilogue.
// Accumulation order per output element equals k_conv_tc / k_conv_tc_persist (same taps, k-steps and pass order), so the three
// kernels are bit-identical and the plan may pick per layer.
#include "tc_common.cuh"
#include <cstdlib>

namespace esr {

constexpr int TH_TW = 8, TH_TH = 16;                                   // output tile
constexpr int TH_HW = TH_TW + 2, TH_HH = TH_TH + 2;                    // halo box 10 x 18 pixels
constexpr uint32_t TH_A_PLANE = ((TH_HW * TH_HH * 128 + 1023) / 1024) * 1024;   // 23040 -> 23552: planes stay 1024-B aligned
constexpr uint32_t TH_A_STAGE = 2 * TH_A_PLANE;
constexpr uint32_t TH_STG_PLANE = 32 * 64;                             // one warp's 32 pixels x 32 channels of one split plane
constexpr uint32_t TH_STG_BUF = 2 * TH_STG_PLANE;                      // one staging buffer: two split planes, or 32 pixels x 32 fp32
static inline uint32_t th_stg_bytes(int bufs) { return 4u * (uint32_t)bufs * TH_STG_BUF; }

__device__ __forceinline__ uint64_t umma_smem_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// TRACE: the clock-stamp instrumentation (ESR_TC_TRACE) is a separate instantiation -- even with a null trace pointer the per-tap checks
// in the single-thread MMA / producer loops cost 2.5 % of the step (measured: 2.578 -> 2.514 ms without them).
template <bool TRACE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tc_halo(const __grid_constant__ ConvTCArgs a)
{
    PDL_LAUNCH_DEPENDENTS();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)a.npad * 128u, b_stage = 2u * b_bytes;
    const uint32_t a_ring = smem_base, b_ring = smem_base + (uint32_t)a.a_stages * TH_A_STAGE;
    const uint32_t stg_ring = b_ring + (uint32_t)a.stages * b_stage;            // out_tma: [4 warps][2 buffers][2 planes][32 px x 64 B]
    const uint32_t bar_base = stg_ring + (a.out_tma ? 4u * (uint32_t)a.stg_bufs * TH_STG_BUF : 0u);
    const uint32_t bar_afull = bar_base, bar_aempty = bar_afull + 8u * a.a_stages;
    const uint32_t bar_bfull = bar_aempty + 8u * a.a_stages, bar_bempty = bar_bfull + 8u * a.stages;
    const uint32_t bar_cfull = bar_bempty + 8u * a.stages, bar_cempty = bar_cfull + 16u;        // per accumulator
    const uint32_t tmem_slot = bar_cempty + 16u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = a.tiles_x * a.tiles_y, n_tiles = a.n_img * tiles_per_img;
    const int n_chunks = a.nkb / 9;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.a_stages; ++s) { mbar_init(bar_afull + 8u * s, 1); mbar_init(bar_aempty + 8u * s, 1); }
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_bfull + 8u * s, 1); mbar_init(bar_bempty + 8u * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_cfull + 8u * i, 1); mbar_init(bar_cempty + 8u * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    PDL_WAIT();                      // everything above is CTA-local set-up; global memory only from here on
    if (TRACE && a.trace && (int)blockIdx.x == a.trace_cta && threadIdx.x == 0) a.trace[7] = clock64();

    if (warp == 0) {
        // ===================== TMA producer: per chunk one halo box per plane, then nine weight tiles =====================
        if (elect_one_sync()) {
            uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
            long long *const tr = (TRACE && a.trace && (int)blockIdx.x == a.trace_cta) ? a.trace : nullptr;      // ESR_TC_TRACE: clock stamps of one CTA
            int pit = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
                const int y0 = (trem / a.tiles_x) * TH_TH, x0 = (trem % a.tiles_x) * TH_TW;
                int src = 0, chunk_base = 0;
                for (int c = 0; c < n_chunks; ++c) {
                    while (c >= a.chunk_end[src]) { chunk_base = a.chunk_end[src]; ++src; }
                    const int simg = a.src_img[src] ? a.src_img[src][img] : img;
                    mbar_wait(bar_aempty + 8u * sa, pha ^ 1u);
                    if (tr && pit < TRACE_N) tr[pit * 8 + 1] = clock64();                   // halo box of the chunk starting at this K-block
                    mbar_expect_tx(bar_afull + 8u * sa, 2u * (uint32_t)(TH_HW * TH_HH * 128));
                    const uint32_t sta = a_ring + sa * TH_A_STAGE;
                    const int c0 = (c - chunk_base) * 64;
                    tma_load_5d(&a.amap[src], bar_afull + 8u * sa, sta, c0, x0 - 1, y0 - 1, simg, 0);
                    tma_load_5d(&a.amap[src], bar_afull + 8u * sa, sta + TH_A_PLANE, c0, x0 - 1, y0 - 1, simg, 1);
                    if (++sa == (uint32_t)a.a_stages) { sa = 0; pha ^= 1u; }
                    for (int t = 0; t < 9; ++t) {
                        const int kb = c * 9 + t;
                        mbar_wait(bar_bempty + 8u * sb, phb ^ 1u);
                        if (tr && pit < TRACE_N) tr[pit * 8 + 0] = clock64();
                        ++pit;
                        mbar_expect_tx(bar_bfull + 8u * sb, b_stage);
                        const uint32_t stb = b_ring + sb * b_stage;
                        tma_load_3d(&a.bmap, bar_bfull + 8u * sb, stb, 0, 0, kb);
                        tma_load_3d(&a.bmap, bar_bfull + 8u * sb, stb + b_bytes, 0, 0, a.nkb + kb);
                        if (++sb == (uint32_t)a.stages) { sb = 0; phb ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc(TC_BLOCK_M, a.npad), idesc2 = umma_idesc(TC_BLOCK_M, 2 * a.npad);
            constexpr uint32_t SBO = (uint32_t)TH_HW * 128u;
            constexpr uint32_t A_DESC_HI = (SBO >> 4) | (1u << 14) | (2u << 29), B_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
            uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
            int it = 0, mit = 0;
            long long *const tr = (TRACE && a.trace && (int)blockIdx.x == a.trace_cta) ? a.trace : nullptr;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t ai = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
                if (tr && mit < TRACE_N) tr[mit * 8 + 5] = clock64();
                mbar_wait(bar_cempty + 8u * ai, aph ^ 1u);          // the epilogue of tile it-2 has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + ai * 256u;
                for (int c = 0; c < n_chunks; ++c) {
                    if (tr && mit < TRACE_N) tr[mit * 8 + 6] = clock64();
                    mbar_wait(bar_afull + 8u * sa, pha);
                    const uint32_t a_hi = a_ring + sa * TH_A_STAGE, a_lo = a_hi + TH_A_PLANE;
                    // This thread's instruction stream between two tcgen05.mma is on the critical path (the tensor pipe's queue is short).
                    // A descriptor's upper word is a constant and its lower word is the address field plus a constant, so the operands of
                    // a tap are base + small immediates instead of a shift / mask / or chain per MMA.
                    const uint32_t ah0 = umma_desc_lo(a_hi), al0 = umma_desc_lo(a_lo);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        if (tr && mit < TRACE_N) tr[mit * 8 + 2] = clock64();
                        mbar_wait(bar_bfull + 8u * sb, phb);
                        tc_fence_after();
                        if (tr && mit < TRACE_N) tr[mit * 8 + 3] = clock64();
                        const uint32_t tap16 = (uint32_t)((t / 3) * TH_HW + (t % 3)) * 8u;         // halo pixel (ty, tx), in 16-byte units
                        const uint32_t b_hi = b_ring + sb * b_stage;
                        const uint32_t bh0 = umma_desc_lo(b_hi), bl0 = bh0 + (b_bytes >> 4);
                        if (a.stack) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t dah = umma_desc(ah0 + tap16 + 2u * k, A_DESC_HI), dal = umma_desc(al0 + tap16 + 2u * k, A_DESC_HI);
                                const uint64_t dbh = umma_desc(bh0 + 2u * k, B_DESC_HI);
                                umma_bf16(acc, dah, dbh, idesc2, (c | t | k) != 0 ? 1u : 0u);      // [B_hi; B_lo] as one 2 npad-row operand
                                umma_bf16(acc, dal, dbh, idesc, 1u);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t dah = umma_desc(ah0 + tap16 + 2u * k, A_DESC_HI), dal = umma_desc(al0 + tap16 + 2u * k, A_DESC_HI);
                                const uint64_t dbh = umma_desc(bh0 + 2u * k, B_DESC_HI), dbl = umma_desc(bl0 + 2u * k, B_DESC_HI);
                                umma_bf16(acc, dal, dbh, idesc, (c | t | k) != 0 ? 1u : 0u);
                                umma_bf16(acc, dah, dbl, idesc, 1u);
                                umma_bf16(acc, dah, dbh, idesc, 1u);
                            }
                        }
                        umma_commit(bar_bempty + 8u * sb);
                        if (tr && mit < TRACE_N) tr[mit * 8 + 4] = clock64();
                        ++mit;
                        if (++sb == (uint32_t)a.stages) { sb = 0; phb ^= 1u; }
                    }
                    umma_commit(bar_aempty + 8u * sa);
                    if (++sa == (uint32_t)a.a_stages) { sa = 0; pha ^= 1u; }
                }
                umma_commit(bar_cfull + 8u * ai);
            }
        }
    } else {
        // ===================== epilogue (4 warps, one TMEM lane quadrant each) =====================
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        int it = 0;
        uint32_t stg_n = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t ai = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
            const int y0 = (trem / a.tiles_x) * TH_TH, x0 = (trem % a.tiles_x) * TH_TW;
            const int y = y0 + m / TH_TW, x = x0 + m % TH_TW;
            const bool valid = (y < a.H) && (x < a.W);
            const size_t pix = ((size_t)img * a.H + (valid ? y : 0)) * a.W + (valid ? x : 0);
            mbar_wait_backoff(bar_cfull + 8u * ai, aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ai * 256u + ((uint32_t)(quad * 32) << 16);
            for (int n0 = 0; n0 < a.npad; n0 += 32) {
                if (a.diag & 16) continue;                          // measurement aid (ESR_TC_DIAG): no TMEM reads, no stores
                uint32_t raw[32];
                if (a.stack) tmem_ld_chunk_stacked(taddr, n0, a.npad, raw);
                else tmem_ld_chunk(taddr, n0, a.npad, raw);
                if (a.out_tma && !(a.diag & 8)) {
                    // The direct form (one pixel per lane, 16-byte stores 128+ bytes apart) costs 32 sector writes per store
                    // instruction and those go through the same L1 / shared-memory pipeline the MMA operands are read from: the
                    // clock trace shows ~1100 cycles of MMA back-pressure per 32-channel chunk.  Here the warp's 32 pixels x 32
                    // channels go to shared memory (swizzled: 4 wavefronts per instruction) and leave with one TMA store per plane.
                    const bool two = a.stg_bufs == 2;
                    const uint32_t buf = stg_ring + ((uint32_t)quad * (uint32_t)a.stg_bufs + (two ? (stg_n & 1u) : 0u)) * TH_STG_BUF;
                    if (lane == 0) {                                    // the store that last read this buffer
                        if (two) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                    float v[32];
                    epilogue_values(a, raw, n0, img, y, x, valid, v);
                    if (a.out_tma == 1) {
                        uint32_t hw[16], lw[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) split_pack2(v[2 * e], v[2 * e + 1], hw[e], lw[e]);
                        const uint32_t row = buf + (uint32_t)lane * 64u, sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t o = row + (((uint32_t)q ^ sw) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o), "r"(hw[4 * q]), "r"(hw[4 * q + 1]), "r"(hw[4 * q + 2]), "r"(hw[4 * q + 3]) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + TH_STG_PLANE), "r"(lw[4 * q]), "r"(lw[4 * q + 1]), "r"(lw[4 * q + 2]), "r"(lw[4 * q + 3]) : "memory");
                        }
                    } else {                                             // fp32 rows of 128 B, SWIZZLE_128B pattern
                        const uint32_t row = buf + (uint32_t)lane * 128u, sw = (uint32_t)lane & 7u;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint32_t o = row + (((uint32_t)q ^ sw) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o), "r"(__float_as_uint(v[4 * q])), "r"(__float_as_uint(v[4 * q + 1])),
                                         "r"(__float_as_uint(v[4 * q + 2])), "r"(__float_as_uint(v[4 * q + 3])) : "memory");
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const int yq = y0 + 4 * quad;
                        if (a.out_tma == 1) {
                            tma_store_5d(&a.omap, buf, a.out_coff + n0, x0, yq, img, 0);
                            tma_store_5d(&a.omap, buf + TH_STG_PLANE, a.out_coff + n0, x0, yq, img, 1);
                        } else {
                            tma_store_5d(&a.omap, buf, n0, x0, yq, img, 0);      // channels >= out_f32_C are clipped by the tensor map
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++stg_n;
                } else if (valid && !(a.diag & 8)) epilogue_chunk(a, raw, n0, pix, img, y, x);
                else if (a.diag & 8) { if (raw[0] == 0x7fc12345u && raw[31] == 0x7fc54321u) a.out_f32[0] = 1.0f; }   // keep the loads alive
                __syncwarp();
            }
            tc_fence_before();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_cempty + 8u * ai) : "memory");
        }
        if (a.out_tma && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging buffers are read until then
    }

    tc_fence_before();
    __syncthreads();
    if (TRACE && a.trace && (int)blockIdx.x == a.trace_cta && threadIdx.x == 0) a.trace[15] = clock64();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static size_t th_smem_bytes(int npad, int a_stages, int b_stages, int staging_bufs)
{
    return 1024 + (size_t)a_stages * TH_A_STAGE + (size_t)b_stages * 2 * npad * 128 + th_stg_bytes(staging_bufs) +
           16 * (size_t)(a_stages + b_stages) + 96;
}

// pipeline depths for a layer of padded width npad (staging: with the epilogue's TMA-store buffers); false if nothing useful fits
bool conv_tc_halo_plan(int npad, int staging, int *a_stages, int *b_stages)
{
    const size_t cap = (size_t)dev_info().max_smem_optin;
    for (int as = 2; as >= 2; --as) {
        int bs = 9;
        while (bs >= 2 && th_smem_bytes(npad, as, bs, staging) > cap) --bs;
        if (bs >= 2) { *a_stages = as; *b_stages = bs; return true; }
    }
    return false;
}

int conv_tc_halo_launch(const ConvTCArgs &a, cudaStream_t st)
{
    static int max_set = 0;
    const size_t smem = th_smem_bytes(a.npad, a.a_stages, a.stages, a.out_tma ? a.stg_bufs : 0);
    if ((int)smem > max_set) {
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_conv_tc_halo<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_conv_tc_halo<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_set = (int)smem;
    }
    const int n_tiles = a.n_img * a.tiles_x * a.tiles_y;
    const unsigned grid = (unsigned)(n_tiles < dev_info().sm_count ? n_tiles : dev_info().sm_count);
    static const char *trace_path = getenv("ESR_TC_TRACE");          // measurement aid: "<file>:<npad>:<nkb>" traces launches of that shape
    if (trace_path) {
        static long long *dbuf = nullptr;
        int want_n = 0, want_k = 0; char path[512] = {0};
        if (sscanf(trace_path, "%511[^:]:%d:%d", path, &want_n, &want_k) == 3 && want_n == a.npad && want_k == a.nkb) {
            if (!dbuf) cudaMalloc(&dbuf, sizeof(long long) * 8 * TRACE_N);
            cudaMemsetAsync(dbuf, 0, sizeof(long long) * 8 * TRACE_N, st);
            ConvTCArgs b = a; b.trace = dbuf; b.trace_cta = getenv("ESR_TC_TRACE_CTA") ? atoi(getenv("ESR_TC_TRACE_CTA")) : 0;
            k_conv_tc_halo<true><<<grid, TC_THREADS, smem, st>>>(b);
            cudaStreamSynchronize(st);
            static long long host[8 * TRACE_N];
            cudaMemcpy(host, dbuf, sizeof(host), cudaMemcpyDeviceToHost);
            FILE *f = fopen(path, "w");
            if (f) {
                fprintf(f, "# HALO npad=%d nkb=%d a_stages=%d b_stages=%d tiles=%d stack=%d; per K-block (tap): prod_after_bempty_wait, prod_after_aempty_wait(first tap of a chunk), mma_before_bfull_wait, mma_after_bfull_wait, mma_after_commit, mma_before_cempty_wait(first tap of a tile), mma_before_afull_wait(first tap of a chunk)\n",
                        a.npad, a.nkb, a.a_stages, a.stages, n_tiles, a.stack);
                for (int i = 0; i < TRACE_N; ++i) { for (int c = 0; c < 8; ++c) fprintf(f, "%lld%c", host[i * 8 + c], c == 7 ? '\n' : ','); }
                fclose(f);
            }
            esr::count_launch();
            return ESR_OK;
        }
    }
    ESR_CUDA_CHECK(launch_pdl(k_conv_tc_halo<false>, dim3(grid), dim3(TC_THREADS), smem, st, a));
    esr::count_launch();
    return ESR_OK;
}

} // namespace esr
