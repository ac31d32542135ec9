// tc_conv3.cu -- 3x3 convolution, second-generation tcgen05 kernel: halo reuse + weight multicast.
//
// The first kernel (tc_conv.cu) fetches one shifted 128-pixel A tile per tap: 9 x 32 KB per 64-channel chunk, and every
// CTA pulls every weight tile from L2.  On B200 that made the main loop L2->SM bandwidth bound (ncu: ~4.7 TB/s, tensor
// pipe ~10 %).  This kernel cuts the operand traffic:
//   * A: per 64-channel chunk ONE TMA box of the (16+2) x 16 pixel halo tile (hi and lo planes, 2 x 36 KB).  The tile is
//     8 pixels wide and 16 tall, the halo rows are 16 pixels (2048 B) apart, so the 128 rows of tap (dy,dx) are the
//     canonical 128B-swizzled K-major layout starting at halo pixel (dy,dx): 8-row groups 2048 B apart (SBO), start
//     address advanced by (dy*16+dx)*128 B (the swizzle is a function of the absolute shared-memory address, so the
//     shifted view stays consistent with what TMA wrote).  All 9 taps x 4 K-steps x 3 split products run from the
//     same shared-memory tile.
//   * B: thread-block clusters of 2 CTAs (adjacent tiles); each CTA loads half of every weight tile and TMA-multicasts it
//     into both CTAs, so the per-CTA weight traffic halves.  Stage release is cluster-wide (tcgen05.commit multicast).
// Epilogue, operand precision (3-pass split bf16, fp32 accumulate in TMEM) and tensor formats are those of tc_conv.cu.
// Warp roles: 0 = A (halo) producer, 1 = MMA issuer + TMEM owner, 2 = B (weights) producer, 3..6 = epilogue.
#include "tc_common.cuh"
#include <cstdlib>

namespace esr {

constexpr int T3_THREADS = 224;
constexpr int T3_TW = 8, T3_TH = 16;                       // output tile (pixels)
constexpr int T3_HALO_W = 16, T3_HALO_H = T3_TH + 2;       // halo box: 16 x 18 pixels (10 columns used)
constexpr uint32_t T3_A_PLANE = T3_HALO_W * T3_HALO_H * 128;   // 36864 B per plane
constexpr uint32_t T3_A_STAGE = 2 * T3_A_PLANE;                // hi + lo

__device__ __forceinline__ void tma_load_3d_mc(const CUtensorMap *map, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                               uint16_t mask)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
                 " [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B descriptor whose start is NOT 1024-byte aligned: SBO = 2048 B (16-pixel halo rows).
// Measured on B200 (tools/diag_halo.py): the tensor core applies the 128B swizzle to the absolute shared-memory address
// (start + row offset), exactly like TMA does when it writes the box, so a start advanced by whole 128-byte rows needs
// NO base-offset correction (setting bits 49-51 to the start's row phase breaks every tap with dx != 0).
__device__ __forceinline__ uint64_t umma_smem_desc_halo(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(2048 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(T3_THREADS, 1) k_conv_tc3(const __grid_constant__ ConvTCArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)a.npad * 128u;           // one plane of one weight tile
    const uint32_t b_stage = 2u * b_bytes;
    const uint32_t a_ring = smem_base, b_ring = smem_base + (uint32_t)a.a_stages * T3_A_STAGE;
    const uint32_t bar_base = b_ring + (uint32_t)a.stages * b_stage;
    const uint32_t bar_afull = bar_base, bar_aempty = bar_afull + 8u * a.a_stages;
    const uint32_t bar_bfull = bar_aempty + 8u * a.a_stages, bar_bempty = bar_bfull + 8u * a.stages;
    const uint32_t bar_accum = bar_bempty + 8u * a.stages, tmem_slot = bar_accum + 8u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t csize = (uint32_t)a.cluster, crank = csize > 1 ? cluster_rank() : 0u;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < a.npad) tmem_cols <<= 1;

    // tile -> (image, y0, x0); the grid is padded to a multiple of the cluster size, padded CTAs recompute the last tile
    const int tiles_per_img = a.tiles_x * a.tiles_y, n_tiles = a.n_img * tiles_per_img;
    const bool tile_valid = (int)blockIdx.x < n_tiles;
    const int tile = tile_valid ? (int)blockIdx.x : n_tiles - 1;
    const int img = tile / tiles_per_img;
    const int trem = tile - img * tiles_per_img;
    const int y0 = (trem / a.tiles_x) * T3_TH, x0 = (trem % a.tiles_x) * T3_TW;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.a_stages; ++s) { mbar_init(bar_afull + 8u * s, 1); mbar_init(bar_aempty + 8u * s, 1); }
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_bfull + 8u * s, 1); mbar_init(bar_bempty + 8u * s, csize); }
        mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    if (csize > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int n_chunks = a.nkb / 9;
    if (warp == 0) {
        // ===================== A producer: one halo box per 64-channel chunk and plane =====================
        if (elect_one_sync()) {
            uint32_t s = 0, ph = 0;
            int src = 0, chunk_base = 0;
            for (int c = 0; c < n_chunks; ++c) {
                while (c >= a.chunk_end[src]) { chunk_base = a.chunk_end[src]; ++src; }
                const int simg = a.src_img[src] ? a.src_img[src][img] : img;
                mbar_wait(bar_aempty + 8u * s, ph ^ 1u);
                mbar_expect_tx(bar_afull + 8u * s, T3_A_STAGE);
                const uint32_t st = a_ring + s * T3_A_STAGE;
                const int c0 = (c - chunk_base) * 64;
                tma_load_5d(&a.amap[src], bar_afull + 8u * s, st, c0, x0 - 1, y0 - 1, simg, 0);
                tma_load_5d(&a.amap[src], bar_afull + 8u * s, st + T3_A_PLANE, c0, x0 - 1, y0 - 1, simg, 1);
                if (++s == (uint32_t)a.a_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ===================== B producer: weight tiles, multicast across the cluster =====================
        if (elect_one_sync()) {
            uint32_t s = 0, ph = 0;
            const uint32_t half_rows = (uint32_t)a.npad / csize;
            const uint16_t mask = (uint16_t)((1u << csize) - 1u);
            for (int kb = 0; kb < a.nkb; ++kb) {
                mbar_wait(bar_bempty + 8u * s, ph ^ 1u);                  // every CTA of the cluster released the stage
                mbar_expect_tx(bar_bfull + 8u * s, b_stage);              // own rows + the peer's rows
                const uint32_t st = b_ring + s * b_stage + crank * half_rows * 128u;
                if (csize > 1) {
                    tma_load_3d_mc(&a.bmap_half, bar_bfull + 8u * s, st, 0, (int)(crank * half_rows), kb, mask);
                    tma_load_3d_mc(&a.bmap_half, bar_bfull + 8u * s, st + b_bytes, 0, (int)(crank * half_rows), a.nkb + kb, mask);
                } else {
                    tma_load_3d(&a.bmap, bar_bfull + 8u * s, st, 0, 0, kb);
                    tma_load_3d(&a.bmap, bar_bfull + 8u * s, st + b_bytes, 0, 0, a.nkb + kb);
                }
                if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc(TC_BLOCK_M, a.npad);
            const uint16_t mask = (uint16_t)((1u << csize) - 1u);
            uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(bar_afull + 8u * sa, pha);
                const uint32_t a_hi = a_ring + sa * T3_A_STAGE, a_lo = a_hi + T3_A_PLANE;
                for (int t = 0; t < 9; ++t) {
                    mbar_wait(bar_bfull + 8u * sb, phb);
                    tc_fence_after();
                    const uint32_t tap_off = (uint32_t)((t / 3) * T3_HALO_W + (t % 3)) * 128u;   // halo pixel (dy, dx)
                    const uint32_t b_hi = b_ring + sb * b_stage, b_lo = b_hi + b_bytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t dah = umma_smem_desc_halo(a_hi + tap_off + 32u * k);
                        const uint64_t dal = umma_smem_desc_halo(a_lo + tap_off + 32u * k);
                        const uint64_t dbh = umma_desc(umma_desc_lo(b_hi) + 2u * k, UMMA_HI_1024), dbl = umma_desc(umma_desc_lo(b_lo) + 2u * k, UMMA_HI_1024);
                        umma_bf16(tmem_base, dal, dbh, idesc, (c | t | k) != 0 ? 1u : 0u);
                        umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                        umma_bf16(tmem_base, dah, dbh, idesc, 1u);
                    }
                    if (csize > 1) umma_commit_mc(bar_bempty + 8u * sb, mask); else umma_commit(bar_bempty + 8u * sb);
                    if (++sb == (uint32_t)a.stages) { sb = 0; phb ^= 1u; }
                }
                umma_commit(bar_aempty + 8u * sa);
                if (++sa == (uint32_t)a.a_stages) { sa = 0; pha ^= 1u; }
            }
            umma_commit(bar_accum);
        }
    } else {
        // ===================== epilogue (warps 3..6 -> TMEM lane quadrants 3,0,1,2) =====================
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        const int y = y0 + m / T3_TW, x = x0 + m % T3_TW;
        const bool valid = tile_valid && (y < a.H) && (x < a.W);
        const size_t pix = ((size_t)img * a.H + (valid ? y : 0)) * a.W + (valid ? x : 0);
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
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
    }

    tc_fence_before();
    // nobody may leave while the peer can still multicast into this CTA's shared memory or arrive on its barriers
    if (csize > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

static size_t t3_smem_bytes(int npad, int a_stages, int b_stages)
{
    return 1024 + (size_t)a_stages * T3_A_STAGE + (size_t)b_stages * 2 * npad * 128 + 16 * (size_t)(a_stages + b_stages) + 64;
}

// chooses the pipeline depths; returns false if the configuration does not fit
bool conv_tc3_plan(int npad, int *a_stages, int *b_stages)
{
    const size_t cap = (size_t)dev_info().max_smem_optin;
    for (int as = 2; as >= 1; --as) {
        int bs = 6;
        while (bs >= 2 && t3_smem_bytes(npad, as, bs) > cap) --bs;
        if (bs >= 2) { *a_stages = as; *b_stages = bs; return true; }
    }
    return false;
}

int conv_tc3_launch(const ConvTCArgs &a, cudaStream_t st)
{
    static int max_set = 0;

    const size_t smem = t3_smem_bytes(a.npad, a.a_stages, a.stages);
    if ((int)smem > max_set) {
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_conv_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_set = (int)smem;
    }
    const int n_tiles = a.n_img * a.tiles_x * a.tiles_y;
    const unsigned grid = (unsigned)((n_tiles + a.cluster - 1) / a.cluster * a.cluster);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(T3_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ESR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_conv_tc3, a));
    esr::count_launch();
    return ESR_OK;
}

} // namespace esr
