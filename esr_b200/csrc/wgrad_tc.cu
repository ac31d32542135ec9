// wgrad_tc.cu -- weight gradient of a stride-1 3x3 (pad 1) / 1x1 convolution on the tensor cores.
//
//   dw[co, ci, ky, kx] = sum over (image, y, x) of g[image, y, x, co] * x[image, y + ky - 1, x + kx - 1, ci]
//
// is, per tap, a GEMM whose reduction dimension is the PIXEL index: D[a-channel, b-channel] += A^T B with A and B both
// stored pixel-major (NHWC rows of 64 channels = one 128-byte shared-memory row per pixel).  That is exactly the
// MN-major SWIZZLE_128B operand layout of tcgen05 (canonical ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units: 64
// channels contiguous in a row, 8 pixel rows per 1024-byte swizzle atom, SBO = 1024 between K atoms, LBO = distance
// between 64-channel tiles), so the same TMA boxes the forward kernel loads (64 ch x TW x TH pixels, tap shift + zero
// fill = padding) feed the MMA directly -- no transposition anywhere.  fp32 parity: operands are split bf16 (hi, lo),
// three MMAs per K step (lo*hi + hi*lo + hi*hi) into fp32 TMEM accumulators, like the forward.
//
// One CTA owns (128 M-side channels, 64 N-side channels, a group of <= 5 taps, a slice of the pixel tiles): it keeps one
// 128 x 64 fp32 accumulator per tap in TMEM (5 x 64 = 320 of the 512 columns; all 9 taps would need 576), streams its
// pixel tiles through a TMA/mbarrier pipeline (the unshifted tensor once per tile, the shifted one once per tap) and adds
// its partial sums to dw with fp32 atomics at the end.  Which tensor sits on the 128-row M side is chosen per layer:
// g (Cout >= 128) or x (Cout == 64 and Cin >= 128); a 64-channel tensor on the M side is loaded twice (rows 64..127
// ignored).  Warps: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue.
#include "tc_common.cuh"
#include "net.cuh"

namespace esr {

constexpr int WG_THREADS = 192;
constexpr uint32_t WG_TILE = TC_BLOCK_M * 128u;       // one 64-channel x 128-pixel plane: 16 KB

struct WgradArgs {
    CUtensorMap gmap, xmap;           // 5-D (C, W, H, img, plane), box (64, TW, TH, 1, 1)
    float *dw;                        // [Cout][Cin][KK]
    int Cout, CoutPad, Cin, KK;       // g has CoutPad (64-multiple) channels, the first Cout are real
    int a_is_x;                       // 1: M side = x channels (shifted per tap), N side = g; 0: M side = g, N side = x
    CUtensorMap hmap;                 // halo mode: x as (64 ch, TW + 2, TH + 2, 1, 1) boxes
    int halo;                         // 1: ONE halo tile of x per pixel tile feeds all taps (shifted MN-major descriptors)
    int m_blocks, n_chunks, groups;   // grid decomposition
    int n_img, H, W, TW, TH, tiles_x, tiles_y, slices;
};

// MN-major, 128B-swizzled operand: start | LBO (between 64-channel tiles) | SBO = 1024 (between 8-pixel K atoms)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// lower word of the same descriptor (the upper word is UMMA_HI_1024); advancing the operand by n bytes is lo + n / 16
__device__ __forceinline__ uint32_t umma_desc_mn_lo(uint32_t smem_addr, uint32_t lbo_bytes)
{
    return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}

__global__ void __launch_bounds__(WG_THREADS, 1) k_wgrad_tc(const __grid_constant__ WgradArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // shared memory: [once buffers x2][per-tap ring x2][barriers]; sizes depend on which tensor is on the M side
    const uint32_t m_bytes = 4u * WG_TILE, n_bytes = 2u * WG_TILE;      // M side: 2 tiles x 2 planes; N side: 1 tile x 2 planes
    const uint32_t once_bytes = a.a_is_x ? n_bytes : m_bytes;           // the unshifted tensor (g)
    const uint32_t halo_plane = ((uint32_t)((a.TW + 2) * (a.TH + 2)) * 128u + 1023u) & ~1023u;
    const uint32_t tap_bytes = a.halo ? 2u * halo_plane : (a.a_is_x ? m_bytes : n_bytes);   // the shifted tensor (x), or its halo box
    const uint32_t once_base = smem_base, ring_base = smem_base + 2u * once_bytes;
    const uint32_t bar_base = ring_base + 2u * tap_bytes;
    const uint32_t bar_gfull = bar_base, bar_gempty = bar_base + 16u, bar_xfull = bar_base + 32u, bar_xempty = bar_base + 48u;
    const uint32_t bar_accum = bar_base + 64u, tmem_slot = bar_base + 72u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // block -> (M block, N chunk, tap group, slice)
    int bid = blockIdx.x;
    const int slice = bid % a.slices; bid /= a.slices;
    const int grp = bid % a.groups; bid /= a.groups;
    const int nch = bid % a.n_chunks; const int mblk = bid / a.n_chunks;
    const int tap0 = grp == 0 ? 0 : 5;
    const int ntap = a.KK == 1 ? 1 : (grp == 0 ? 5 : 4);
 const int m_ch = a.a_is_x ? a.Cin : a.CoutPad;               // channels of the M-side tensor as stored
    const int m_real = a.a_is_x ? a.Cin : a.Cout;
    const int m0 = mblk * 128;
    const bool m_dup = m0 + 64 >= m_ch;                                  // only 64 channels left on the M side
    const int n0 = nch * 64;
    const int tiles_per_img = a.tiles_x * a.tiles_y, n_tiles = a.n_img * tiles_per_img;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_gfull + 8u * s, 1); mbar_init(bar_gempty + 8u * s, 1);
            mbar_init(bar_xfull + 8u * s, 1); mbar_init(bar_xempty + 8u * s, 1);
        }
        mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (elect_one_sync()) {
            uint32_t gs = 0, gph = 0, xs = 0, xph = 0;
            for (int t = slice; t < n_tiles; t += a.slices) {
                const int img = t / tiles_per_img, tr = t - img * tiles_per_img;
                const int y0 = (tr / a.tiles_x) * a.TH, x0 = (tr % a.tiles_x) * a.TW;
                // ---- g tile(s), unshifted
                mbar_wait(bar_gempty + 8u * gs, gph ^ 1u);
                mbar_expect_tx(bar_gfull + 8u * gs, once_bytes);
                const uint32_t gb = once_base + gs * once_bytes;
                if (!a.a_is_x) {                                         // M side: [hi: tile0, tile1][lo: tile0, tile1]
                    const int c1 = m_dup ? m0 : m0 + 64;
                    tma_load_5d(&a.gmap, bar_gfull + 8u * gs, gb, m0, x0, y0, img, 0);
                    tma_load_5d(&a.gmap, bar_gfull + 8u * gs, gb + WG_TILE, c1, x0, y0, img, 0);
                    tma_load_5d(&a.gmap, bar_gfull + 8u * gs, gb + 2u * WG_TILE, m0, x0, y0, img, 1);
                    tma_load_5d(&a.gmap, bar_gfull + 8u * gs, gb + 3u * WG_TILE, c1, x0, y0, img, 1);
                } else {
                    tma_load_5d(&a.gmap, bar_gfull + 8u * gs, gb, n0, x0, y0, img, 0);
                    tma_load_5d(&a.gmap, bar_gfull + 8u * gs, gb + WG_TILE, n0, x0, y0, img, 1);
                }
                if (++gs == 2) { gs = 0; gph ^= 1u; }
                if (a.halo) {
                    // ---- ONE halo box of x for all taps of this pixel tile (hi and lo planes); zero fill outside = padding
                    mbar_wait(bar_xempty + 8u * xs, xph ^ 1u);
                    mbar_expect_tx(bar_xfull + 8u * xs, 2u * (uint32_t)((a.TW + 2) * (a.TH + 2)) * 128u);
                    const uint32_t xb = ring_base + xs * tap_bytes;
                    tma_load_5d(&a.hmap, bar_xfull + 8u * xs, xb, n0, x0 - 1, y0 - 1, img, 0);
                    tma_load_5d(&a.hmap, bar_xfull + 8u * xs, xb + halo_plane, n0, x0 - 1, y0 - 1, img, 1);
                    if (++xs == 2) { xs = 0; xph ^= 1u; }
                    continue;
                }
                // ---- x tile(s), one per tap, shifted; out-of-image pixels are zero-filled = the conv padding
                for (int j = 0; j < ntap; ++j) {
                    const int tap = tap0 + j;
                    const int dy = a.KK == 9 ? tap / 3 - 1 : 0, dx = a.KK == 9 ? tap % 3 - 1 : 0;
                    mbar_wait(bar_xempty + 8u * xs, xph ^ 1u);
                    mbar_expect_tx(bar_xfull + 8u * xs, tap_bytes);
                    const uint32_t xb = ring_base + xs * tap_bytes;
                    if (a.a_is_x) {
                        const int c1 = m_dup ? m0 : m0 + 64;
                        tma_load_5d(&a.xmap, bar_xfull + 8u * xs, xb, m0, x0 + dx, y0 + dy, img, 0);
                        tma_load_5d(&a.xmap, bar_xfull + 8u * xs, xb + WG_TILE, c1, x0 + dx, y0 + dy, img, 0);
                        tma_load_5d(&a.xmap, bar_xfull + 8u * xs, xb + 2u * WG_TILE, m0, x0 + dx, y0 + dy, img, 1);
                        tma_load_5d(&a.xmap, bar_xfull + 8u * xs, xb + 3u * WG_TILE, c1, x0 + dx, y0 + dy, img, 1);
                    } else {
                        tma_load_5d(&a.xmap, bar_xfull + 8u * xs, xb, n0, x0 + dx, y0 + dy, img, 0);
                        tma_load_5d(&a.xmap, bar_xfull + 8u * xs, xb + WG_TILE, n0, x0 + dx, y0 + dy, img, 1);
                    }
                    if (++xs == 2) { xs = 0; xph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            // MN-major A and B (bits 15, 16), fp32 accumulate, bf16 operands, M = 128, N = 64
            const uint32_t idesc = umma_idesc(TC_BLOCK_M, 64) | (1u << 15) | (1u << 16);
            uint32_t gs = 0, gph = 0, xs = 0, xph = 0;
            bool first = true;
            for (int t = slice; t < n_tiles; t += a.slices) {
                mbar_wait(bar_gfull + 8u * gs, gph);
                const uint32_t gb = once_base + gs * once_bytes;
                if (a.halo) {
                    mbar_wait(bar_xfull + 8u * xs, xph);
                    tc_fence_after();
                    const uint32_t xb = ring_base + xs * tap_bytes;
                    const uint32_t m_hi = gb, m_lo = gb + 2u * WG_TILE;
                    const uint32_t hw = (uint32_t)(a.TW + 2);
                    for (int j = 0; j < ntap; ++j) {
                        const int tap = tap0 + j, ky = tap / 3, kx = tap % 3;
                        const uint32_t d = tmem_base + (uint32_t)j * 64u;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {                    // K step k = tile row k: 16 pixels = 16 rows of the halo box
                            const uint32_t sh = (((uint32_t)(k + ky)) * hw + (uint32_t)kx) * 128u;
                            const uint64_t dah = umma_desc(umma_desc_mn_lo(m_hi, WG_TILE) + 128u * k, UMMA_HI_1024), dal = umma_desc(umma_desc_mn_lo(m_lo, WG_TILE) + 128u * k, UMMA_HI_1024);
                            const uint64_t dbh = umma_desc(umma_desc_mn_lo(xb, WG_TILE) + (sh >> 4), UMMA_HI_1024), dbl = umma_desc(umma_desc_mn_lo(xb + halo_plane, WG_TILE) + (sh >> 4), UMMA_HI_1024);
                            umma_bf16(d, dal, dbh, idesc, (first && k == 0) ? 0u : 1u);
                            umma_bf16(d, dah, dbl, idesc, 1u);
                            umma_bf16(d, dah, dbh, idesc, 1u);
                        }
                    }
                    umma_commit(bar_xempty + 8u * xs);
                    if (++xs == 2) { xs = 0; xph ^= 1u; }
                    umma_commit(bar_gempty + 8u * gs);
                    if (++gs == 2) { gs = 0; gph ^= 1u; }
                    first = false;
                    continue;
                }
                for (int j = 0; j < ntap; ++j) {
                    mbar_wait(bar_xfull + 8u * xs, xph);
                    tc_fence_after();
                    const uint32_t xb = ring_base + xs * tap_bytes;
                    const uint32_t m_hi = a.a_is_x ? xb : gb, m_lo = m_hi + 2u * WG_TILE;       // M side planes (2 tiles each)
                    const uint32_t n_hi = a.a_is_x ? gb : xb, n_lo = n_hi + WG_TILE;            // N side planes
                    const uint32_t d = tmem_base + (uint32_t)j * 64u;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {                        // 8 x (K = 16 pixels = 16 rows = 2048 bytes)
                        const uint64_t dah = umma_desc(umma_desc_mn_lo(m_hi, WG_TILE) + 128u * k, UMMA_HI_1024), dal = umma_desc(umma_desc_mn_lo(m_lo, WG_TILE) + 128u * k, UMMA_HI_1024);
                        const uint64_t dbh = umma_desc(umma_desc_mn_lo(n_hi, WG_TILE) + 128u * k, UMMA_HI_1024), dbl = umma_desc(umma_desc_mn_lo(n_lo, WG_TILE) + 128u * k, UMMA_HI_1024);
                        umma_bf16(d, dal, dbh, idesc, (first && k == 0) ? 0u : 1u);
                        umma_bf16(d, dah, dbl, idesc, 1u);
                        umma_bf16(d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(bar_xempty + 8u * xs);
                    if (++xs == 2) { xs = 0; xph ^= 1u; }
                }
                umma_commit(bar_gempty + 8u * gs);
                if (++gs == 2) { gs = 0; gph ^= 1u; }
                first = false;
            }
            umma_commit(bar_accum);
        }
    } else if (slice < n_tiles) {
        // ===================== epilogue: TMEM -> fp32 atomics into dw =====================
        const int quad = warp & 3;
        const int m = quad * 32 + lane;                                  // accumulator row = M-side channel
        const bool row_ok = (m < 64 || !m_dup) && (m0 + m < m_real);
        mbar_wait_backoff(bar_accum, 0);
        tc_fence_after();
        for (int j = 0; j < ntap; ++j) {
            const int tap = tap0 + j;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                uint32_t raw[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * 64 + half * 32), raw);
                if (row_ok) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int nc = n0 + half * 32 + c;
                        const int co = a.a_is_x ? nc : m0 + m, ci = a.a_is_x ? m0 + m : nc;
                        if (co < a.Cout) atomicAdd(a.dw + ((size_t)co * a.Cin + ci) * a.KK + tap, __uint_as_float(raw[c]));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// x_split: split NHWC [2][B][H][W][Cin]; g_split: split NHWC [2][B][H][W][CoutPad] (channels >= Cout zero); dw pre-zeroed.
int wgrad_tc(const __nv_bfloat16 *x_split, const __nv_bfloat16 *g_split, int B, int Cin, int H, int W, int Cout, int CoutPad, int ksz,
             float *dw, cudaStream_t st)
{
    if (Cin % 64 != 0 || CoutPad % 64 != 0 || CoutPad < Cout || (ksz != 3 && ksz != 1)) return ESR_EINVAL;
    int rc;
    SplitTensor xs; xs.base = const_cast<__nv_bfloat16 *>(x_split); xs.n_img = B; xs.H = H; xs.W = W; xs.C = Cin;
    SplitTensor gs; gs.base = const_cast<__nv_bfloat16 *>(g_split); gs.n_img = B; gs.H = H; gs.W = W; gs.C = CoutPad;
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.dw = dw; a.Cout = Cout; a.Cin = Cin; a.KK = ksz * ksz;
    a.a_is_x = (Cout == 64 && Cin >= 128) ? 1 : 0;
    // halo mode: 16 x 8 tiles (a K step = one 16-pixel tile row, contiguous in the halo box); ESR_WGRAD_NO_HALO=1 turns it off
    a.halo = (ksz == 3 && !a.a_is_x && W >= 12 && getenv("ESR_WGRAD_NO_HALO") == nullptr) ? 1 : 0;
    a.TW = a.halo ? 16 : (W >= 24 ? 32 : (W >= 12 ? 16 : 8)); a.TH = TC_BLOCK_M / a.TW;
    if ((rc = tc_make_amap(gs, a.TW, a.TH, &a.gmap))) return rc;
    if ((rc = tc_make_amap(xs, a.TW, a.TH, &a.xmap))) return rc;
    if (a.halo && (rc = tc_make_amap(xs, a.TW + 2, a.TH + 2, &a.hmap))) return rc;
    const int m_ch = a.a_is_x ? Cin : CoutPad, n_ch = a.a_is_x ? CoutPad : Cin;
    a.CoutPad = CoutPad;
    a.m_blocks = (m_ch + 127) / 128; a.n_chunks = n_ch / 64; a.groups = ksz == 3 ? 2 : 1;
    a.n_img = B; a.H = H; a.W = W;
    a.tiles_x = (W + a.TW - 1) / a.TW; a.tiles_y = (H + a.TH - 1) / a.TH;
    const int n_tiles = B * a.tiles_x * a.tiles_y, base = a.m_blocks * a.n_chunks * a.groups;
    // one CTA per SM.  (Every CTA ends with 128 x 64 x taps fp32 atomics; giving small problems fewer, longer CTAs was measured
    // slower on B200: min 6 tiles per CTA +0.3 ms, min 16 +2 ms per cfg2 training iteration -- ESR_WGRAD_MIN_TILES to retest.)
    int slices = (dev_info().sm_count + base - 1) / base;
    static const int min_tiles = getenv("ESR_WGRAD_MIN_TILES") ? atoi(getenv("ESR_WGRAD_MIN_TILES")) : 1;
    if (slices > n_tiles / min_tiles) slices = n_tiles / min_tiles;
    if (slices < 1) slices = 1;
    a.slices = slices;
    const size_t halo_plane = align_up((size_t)(a.TW + 2) * (a.TH + 2) * 128, 1024);
    const size_t smem = 1024 + 2 * (size_t)(4 * WG_TILE) + 2 * (a.halo ? 2 * halo_plane : (size_t)(2 * WG_TILE)) + 128;
    static size_t attr_done = 0;
    if (smem > attr_done) {
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = smem;
    }
    k_wgrad_tc<<<(unsigned)(base * slices), WG_THREADS, smem, st>>>(a);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

} // namespace esr
