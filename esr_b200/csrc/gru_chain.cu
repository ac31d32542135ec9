// gru_chain.cu -- the whole bidirectional ConvGRU recurrence of a sequence batch in ONE cooperative kernel.
//
// Reference: TimePropagation.global_time_corre (models/model.py:91-124) calling RecurrentConvLayer / ConvGRU
// (models/submodules.py:340-344, 496-514) once per frame and direction, window after window with the state carried.
// Per step:  z, r = sigmoid(conv3x3(cat(x, h)))   (update | reset gates, one N = 128 GEMM)
//            o    = tanh(conv3x3(cat(x, h * r)))  (N = 64 GEMM),   h' = h (1 - z) + o z
// The two convolutions of a step and consecutive steps are separated by true grid-wide dependencies (3x3 halos), so
// the per-launch version (tc_conv.cu, 2 launches per step) spends most of each ~25 us launch on launch latency,
// TMEM allocation, barrier set-up, pipeline fill and drain.  Here the CTAs stay resident: TMEM, mbarriers and tensor
// maps are set up once, every phase (2 per step) is a tcgen05 implicit GEMM over the CTA's tile(s) followed by the
// fused gate epilogue, and phases are separated by a grid barrier (co-residency guaranteed by the cooperative launch).
//
// Memory ordering across a phase boundary: the epilogue writes h*r / h' with generic-proxy stores; the next phase reads
// them through TMA (async proxy) from OTHER CTAs.  Writers: stores -> fence.proxy.async -> __threadfence -> barrier
// arrive (release); readers: barrier wait (acquire) -> fence.proxy.async -> TMA.  z and h are re-read only by the thread
// that wrote them (same tile -> same CTA -> same TMEM lane in every phase).
#include "tc_common.cuh"
#include "net.cuh"
#include <cstdlib>

namespace esr {

constexpr int GC_THREADS = 320;    // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quadrant)

struct GruChainArgs {
    CUtensorMap amap_xc, amap_hs, amap_rh;      // 5-D maps (64 ch, W, H, img, plane), box (64, TW, TH, 1, 1)
    CUtensorMap bmap_zr, bmap_go;               // packed weights: update|reset (N=128), out gate (N=64); 18 K-blocks each
    const float *bias_zr, *bias_go;             // [128], [64]
    __nv_bfloat16 *hs; size_t hs_plane;         // state slots: image = slot * 2B + j
    __nv_bfloat16 *rh; size_t rh_plane;         // h * r, [2B, H, W, 64]
    float *zbuf;                                // update gate, fp32 [2B, H, W, 64]
    unsigned int *barrier;                      // grid barrier counter (zeroed before the launch)
    int B, N, nsteps, H, W, TW, TH, tiles_x, tiles_y, stages;
    int diag;                                   // measurement aid (ESR_GRU_DIAG): 1 = the epilogue skips its global stores, 2 = also its global loads (wrong results)
    int wpre;                                   // 1: weight tiles of the first state-side K-blocks are issued before the phase barrier wait (measured slower)
    int cluster;                                // > 1: one thread-block cluster per image (tiles_per_img CTAs), phase barrier through DSMEM mbarriers
};

__device__ __forceinline__ void gc_grid_barrier(unsigned int *counter, unsigned int target)
{
    asm volatile("fence.proxy.async;" ::: "memory");      // this thread's generic stores before later async-proxy reads
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
    asm volatile("fence.proxy.async;" ::: "memory");
}

// Fused gate epilogue of one tile: TMEM -> registers -> sigmoid / tanh / blend -> z (fp32), h*r or h' (split bf16).
// Eight warps: two per TMEM lane quadrant, each taking half of the columns.
__device__ __forceinline__ void gc_epilogue(const GruChainArgs &a, int which, int g, int img, int y0, int x0, int warp, int lane,
                                            uint32_t tmem_acc)
{
    const int B2 = 2 * a.B;
    const int npad = which == 0 ? 128 : 64;
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int y = y0 + m / a.TW, x = x0 + m % a.TW;
    const bool valid = (y < a.H) && (x < a.W);
    const size_t pix = ((size_t)img * a.H + (valid ? y : 0)) * a.W + (valid ? x : 0);   // within a 2B-image tensor
    const __nv_bfloat16 *h_prev = a.hs + ((size_t)g * B2 * a.H * a.W + pix) * 64;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16);
    for (int n0 = half * (npad / 2); n0 < (half + 1) * (npad / 2); n0 += 32) {
        uint32_t raw[32];
        tmem_ld32(taddr + (uint32_t)n0, raw);
        if (valid) {
            float v[32];
            const float4 *bp = reinterpret_cast<const float4 *>((which == 0 ? a.bias_zr : a.bias_go) + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b = bp[q];
                v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b.x;
                v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
                v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z;
                v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
            }
            if (which == 0) {
                act32(v, ACT_SIGMOID);
                if (n0 < 64) {                                      // update gate z (fp32)
                    float4 *zp = reinterpret_cast<float4 *>(a.zbuf + pix * 64 + n0);
                    if (!(a.diag & 1)) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) zp[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                } else {                                            // reset gate r -> h * r
                    float h[32];
                    if (!(a.diag & 2)) load_split32(h_prev + (n0 - 64), a.hs_plane, h);
                    else { for (int j = 0; j < 32; ++j) h[j] = 1.0f; }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= h[j];
                    if (!(a.diag & 1)) store_split32(a.rh + pix * 64 + (n0 - 64), a.rh_plane, v);
                }
            } else {                                                // h' = h (1 - z) + tanh(.) z
                float h[32];
                if (!(a.diag & 2)) load_split32(h_prev + n0, a.hs_plane, h);
                else { for (int j = 0; j < 32; ++j) h[j] = 1.0f; }
                const float4 *zp = reinterpret_cast<const float4 *>(a.zbuf + pix * 64 + n0);
                float4 zq[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) zq[q] = (a.diag & 2) ? make_float4(0.5f, 0.5f, 0.5f, 0.5f) : zp[q];
                act32(v, ACT_TANH);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float zz[4] = {zq[q].x, zq[q].y, zq[q].z, zq[q].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = 4 * q + e;
                        v[j] = h[j] * (1.0f - zz[e]) + v[j] * zz[e];
                    }
                }
                __nv_bfloat16 *h_new = a.hs + ((size_t)(g + 1) * B2 * a.H * a.W + pix) * 64;
                if (!(a.diag & 1)) store_split32(h_new + n0, a.hs_plane, v);
            }
        }
        __syncwarp();
    }
}

// The same epilogue for the pipelined kernel, split in two: the h / z operands of this warp's columns do not depend on the
// phase's MMAs (h is the previous step's state, z was written by this very thread one phase earlier), so they are loaded into
// registers BEFORE the accumulator barrier is awaited -- their L2 latency (queued behind the next phase's operand prefetch)
// overlaps the main loop instead of extending the serial epilogue.
struct GcPre { uint4 hh[2][4], hl[2][4]; float4 zq[8]; };

__device__ __forceinline__ void gc_prefetch(const GruChainArgs &a, int which, int g, int img, int y0, int x0, int warp, int lane, GcPre &pre)
{
    const int B2 = 2 * a.B;
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int y = y0 + m / a.TW, x = x0 + m % a.TW;
    if (!(y < a.H && x < a.W)) return;
    const size_t pix = ((size_t)img * a.H + y) * a.W + x;
    const __nv_bfloat16 *h_prev = a.hs + ((size_t)g * B2 * a.H * a.W + pix) * 64;
    if (which == 0) {
        if (half == 1) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint4 *ph = reinterpret_cast<const uint4 *>(h_prev + c * 32), *pl = reinterpret_cast<const uint4 *>(h_prev + c * 32 + a.hs_plane);
#pragma unroll
                for (int q = 0; q < 4; ++q) { pre.hh[c][q] = ph[q]; pre.hl[c][q] = pl[q]; }
            }
        }
    } else {
        const uint4 *ph = reinterpret_cast<const uint4 *>(h_prev + half * 32), *pl = reinterpret_cast<const uint4 *>(h_prev + half * 32 + a.hs_plane);
        const float4 *zp = reinterpret_cast<const float4 *>(a.zbuf + pix * 64 + half * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) { pre.hh[0][q] = ph[q]; pre.hl[0][q] = pl[q]; }
#pragma unroll
        for (int q = 0; q < 8; ++q) pre.zq[q] = zp[q];
    }
}

__device__ __forceinline__ void gc_unpack32(const uint4 (&hh)[4], const uint4 (&hl)[4], float (&o)[32])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t hw[4] = {hh[q].x, hh[q].y, hh[q].z, hh[q].w}, lw[4] = {hl[q].x, hl[q].y, hl[q].z, hl[q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            o[q * 8 + e * 2 + 0] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
            o[q * 8 + e * 2 + 1] = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
        }
    }
}

__device__ __forceinline__ void gc_epilogue_pre(const GruChainArgs &a, int which, int g, int img, int y0, int x0, int warp, int lane,
                                                uint32_t tmem_acc, const GcPre &pre)
{
    const int B2 = 2 * a.B;
    const int npad = which == 0 ? 128 : 64;
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int y = y0 + m / a.TW, x = x0 + m % a.TW;
    const bool valid = (y < a.H) && (x < a.W);
    const size_t pix = ((size_t)img * a.H + (valid ? y : 0)) * a.W + (valid ? x : 0);
    const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16);
    int c = 0;
    for (int n0 = half * (npad / 2); n0 < (half + 1) * (npad / 2); n0 += 32, ++c) {
        uint32_t raw[32];
        tmem_ld32(taddr + (uint32_t)n0, raw);
        if (valid) {
            float v[32];
            const float4 *bp = reinterpret_cast<const float4 *>((which == 0 ? a.bias_zr : a.bias_go) + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b = bp[q];
                v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b.x;
                v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
                v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z;
                v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
            }
            if (which == 0) {
                act32(v, ACT_SIGMOID);
                if (n0 < 64) {
                    float4 *zp = reinterpret_cast<float4 *>(a.zbuf + pix * 64 + n0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) zp[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                } else {
                    float h[32];
                    if (c == 0) gc_unpack32(pre.hh[0], pre.hl[0], h); else gc_unpack32(pre.hh[1], pre.hl[1], h);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= h[j];
                    store_split32(a.rh + pix * 64 + (n0 - 64), a.rh_plane, v);
                }
            } else {
                float h[32];
                gc_unpack32(pre.hh[0], pre.hl[0], h);
                act32(v, ACT_TANH);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float zz[4] = {pre.zq[q].x, pre.zq[q].y, pre.zq[q].z, pre.zq[q].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = 4 * q + e;
                        v[j] = h[j] * (1.0f - zz[e]) + v[j] * zz[e];
                    }
                }
                __nv_bfloat16 *h_new = a.hs + ((size_t)(g + 1) * B2 * a.H * a.W + pix) * 64;
                if (!(a.diag & 1)) store_split32(h_new + n0, a.hs_plane, v);
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(GC_THREADS, 1) k_gru_chain(const __grid_constant__ GruChainArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t B_MAX = 128u * 128u;                               // one plane of the N = 128 weight tile
    constexpr uint32_t STAGE = 2u * TC_A_BYTES + 2u * B_MAX;              // 64 KB
    const uint32_t bar_base = smem_base + (uint32_t)a.stages * STAGE;
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * a.stages, bar_accum = bar_base + 16u * a.stages;
    const uint32_t tmem_slot = bar_accum + 8u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B2 = 2 * a.B;
    const int tiles_per_img = a.tiles_x * a.tiles_y, n_tiles = B2 * tiles_per_img;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // pipeline state, private to each role, identical sequences on every role
    uint32_t ps = 0, pph = 0;          // producer stage / phase bit
    uint32_t ms = 0, mph = 0;          // MMA stage / phase bit
    uint32_t acc_ph = 0;               // accumulator barrier parity (flips once per tile)

    for (int p = 0; p < 2 * a.nsteps; ++p) {
        const int g = p >> 1, which = p & 1;              // which: 0 = update|reset gates, 1 = candidate + blend
        const int w_idx = g / a.N, s_idx = g - w_idx * a.N;
        const int npad = which == 0 ? 128 : 64;
        const uint32_t b_bytes = (uint32_t)npad * 128u;
        const uint32_t stage_bytes = 2u * TC_A_BYTES + 2u * b_bytes;
        const CUtensorMap *bmap = which == 0 ? &a.bmap_zr : &a.bmap_go;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int img = tile / tiles_per_img;         // j in [0, 2B): j < B forward direction, else time-reversed
            const int trem = tile - img * tiles_per_img;
            const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
            // frame of this image at this step: forward reads window slot s, reverse reads slot N-1-s (model.py:95-96,107-108)
            const int bb = img < a.B ? img : img - a.B;
            const int xc_img = (w_idx * a.B + bb) * a.N + (img < a.B ? s_idx : a.N - 1 - s_idx);
            const int hs_img = g * B2 + img;
            if (warp == 0) {
                if (elect_one_sync()) {
                    for (int kb = 0; kb < 18; ++kb) {
                        const int src = kb / 9, tap = kb - src * 9;
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        mbar_wait(bar_empty + 8u * ps, pph ^ 1u);
                        mbar_expect_tx(bar_full + 8u * ps, stage_bytes);
                        const uint32_t st = smem_base + ps * STAGE;
                        const CUtensorMap *am = src == 0 ? &a.amap_xc : (which == 0 ? &a.amap_hs : &a.amap_rh);
                        const int simg = src == 0 ? xc_img : (which == 0 ? hs_img : img);
                        tma_load_5d(am, bar_full + 8u * ps, st, 0, x0 + dx, y0 + dy, simg, 0);
                        tma_load_5d(am, bar_full + 8u * ps, st + TC_A_BYTES, 0, x0 + dx, y0 + dy, simg, 1);
                        tma_load_3d(bmap, bar_full + 8u * ps, st + 2u * TC_A_BYTES, 0, 0, kb);
                        tma_load_3d(bmap, bar_full + 8u * ps, st + 2u * TC_A_BYTES + b_bytes, 0, 0, 18 + kb);
                        if (++ps == (uint32_t)a.stages) { ps = 0; pph ^= 1u; }
                    }
                }
            } else if (warp == 1) {
                if (elect_one_sync()) {
                    const uint32_t idesc = umma_idesc(TC_BLOCK_M, npad);
                    for (int kb = 0; kb < 18; ++kb) {
                        mbar_wait(bar_full + 8u * ms, mph);
                        tc_fence_after();
                        const uint32_t st = smem_base + ms * STAGE;
                        const uint32_t a_hi = st, a_lo = st + TC_A_BYTES, b_hi = st + 2u * TC_A_BYTES, b_lo = b_hi + b_bytes;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t dah = umma_desc(umma_desc_lo(a_hi) + 2u * k, UMMA_HI_1024), dal = umma_desc(umma_desc_lo(a_lo) + 2u * k, UMMA_HI_1024);
                            const uint64_t dbh = umma_desc(umma_desc_lo(b_hi) + 2u * k, UMMA_HI_1024), dbl = umma_desc(umma_desc_lo(b_lo) + 2u * k, UMMA_HI_1024);
                            umma_bf16(tmem_base, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);
                            umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                            umma_bf16(tmem_base, dah, dbh, idesc, 1u);
                        }
                        umma_commit(bar_empty + 8u * ms);
                        if (++ms == (uint32_t)a.stages) { ms = 0; mph ^= 1u; }
                    }
                    umma_commit(bar_accum);
                }
            } else {
                mbar_wait(bar_accum, acc_ph);
                tc_fence_after();
                gc_epilogue(a, which, g, img, y0, x0, warp, lane, tmem_base);
                tc_fence_before();
            }
            acc_ph ^= 1u;
            // the next tile's MMAs overwrite the accumulator: wait until this tile's epilogue has read it
            __syncthreads();
            tc_fence_after();
        }
        gc_grid_barrier(a.barrier, (unsigned int)(p + 1) * gridDim.x);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}


// ------------------------------------------------------------------------------------------------
// Pipelined variant (one tile per CTA, the usual case): the x-side half of a phase's K loop (9 of 18 K-blocks: taps over
// the step's input features) does not depend on the previous phase, so the producer and the MMA thread run it BEFORE the
// grid barrier, into the second of two TMEM accumulators, while the epilogue warps are still finishing the previous
// phase.  Only the 9 state-side K-blocks (h, or h*r) wait for the barrier.  The barrier is split: the epilogue warps
// arrive (after their stores and fences), the producer thread alone waits.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GC_THREADS, 1) k_gru_chain_pipe(const __grid_constant__ GruChainArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t B_MAX = 128u * 128u;
    constexpr uint32_t STAGE = 2u * TC_A_BYTES + 2u * B_MAX;
    const uint32_t bar_base = smem_base + (uint32_t)a.stages * STAGE;
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * a.stages, bar_accum = bar_base + 16u * a.stages;   // 2 accum barriers
    const uint32_t bar_phase = bar_accum + 16u;                    // cluster mode: one arrival per CTA of the image and phase
    const uint32_t tmem_slot = bar_phase + 8u;
    const bool cl = a.cluster > 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B2 = 2 * a.B;
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    const int tile = blockIdx.x;                                  // grid == number of tiles
    const int img = tile / tiles_per_img;
    const int trem = tile - img * tiles_per_img;
    const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
    const int bb = img < a.B ? img : img - a.B;
    const int n_phases = 2 * a.nsteps;
    // A tile's 3x3 halo only reaches tiles of the SAME image, so the phase barrier is per image (tiles_per_img CTAs, own
    // counter 32 bytes apart) instead of grid-wide: 16 x fewer arrivals per counter and no waiting for other images' stragglers.
    const bool per_img = B2 <= 64;
    unsigned int *bar_ctr = a.barrier + (per_img ? img * 8 : 0);
    const unsigned int bar_n = per_img ? (unsigned int)tiles_per_img : gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        mbar_init(bar_accum, 1); mbar_init(bar_accum + 8u, 1);
        mbar_init(bar_phase, cl ? (uint32_t)a.cluster : 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);                    // two 128-column accumulators
    tc_fence_before();
    if (cl) pair_sync(); else __syncthreads();                    // cluster mode: every CTA's phase barrier exists before anyone arrives on it
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (elect_one_sync()) {
            uint32_t ps = 0, pph = 0;
            for (int p = 0; p < n_phases; ++p) {
                const int g = p >> 1, which = p & 1;
                const int w_idx = g / a.N, s_idx = g - w_idx * a.N;
                const uint32_t b_bytes = (which == 0 ? 128u : 64u) * 128u;
                const uint32_t stage_bytes = 2u * TC_A_BYTES + 2u * b_bytes;
                const CUtensorMap *bmap = which == 0 ? &a.bmap_zr : &a.bmap_go;
                const int xc_img = (w_idx * a.B + bb) * a.N + (img < a.B ? s_idx : a.N - 1 - s_idx);
                // weight tiles never depend on the recurrent state: for the first `pre` state-side K-blocks they are issued BEFORE the
                // phase barrier is awaited (as soon as their stages are free), only the h / h*r tiles wait for it
                const int pre = (p > 0 && a.wpre) ? (a.stages < 9 ? a.stages : 9) : 0;
                for (int kb = 0; kb < 18; ++kb) {
                    const int src = kb / 9, tap = kb - src * 9;
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    if (kb == 9 && p > 0) {
                        uint32_t qs = ps, qph = pph;
                        for (int j = 0; j < pre; ++j) {
                            mbar_wait(bar_empty + 8u * qs, qph ^ 1u);
                            mbar_expect_tx(bar_full + 8u * qs, stage_bytes);
                            const uint32_t stq = smem_base + qs * STAGE;
                            tma_load_3d(bmap, bar_full + 8u * qs, stq + 2u * TC_A_BYTES, 0, 0, 9 + j);
                            tma_load_3d(bmap, bar_full + 8u * qs, stq + 2u * TC_A_BYTES + b_bytes, 0, 0, 18 + 9 + j);
                            if (++qs == (uint32_t)a.stages) { qs = 0; qph ^= 1u; }
                        }
                        // state-side operands are written by every CTA's epilogue of phase p-1: wait for all of them
                        if (cl) {
                            mbar_wait_cluster(bar_phase, (uint32_t)((p - 1) & 1));
                        } else {
                            const unsigned int target = (unsigned int)p * bar_n;
                            unsigned int v;
                            do {
                                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar_ctr) : "memory");
                                                        } while (v < target);
                        }
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    const bool b_done = kb >= 9 && kb < 9 + pre;          // this stage's weights (and its expect_tx) are already on the way
                    const uint32_t st = smem_base + ps * STAGE;
                    const CUtensorMap *am = src == 0 ? &a.amap_xc : (which == 0 ? &a.amap_hs : &a.amap_rh);
                    const int simg = src == 0 ? xc_img : (which == 0 ? g * B2 + img : img);
                    if (!b_done) {
                        mbar_wait(bar_empty + 8u * ps, pph ^ 1u);
                        mbar_expect_tx(bar_full + 8u * ps, stage_bytes);
                    }
                    tma_load_5d(am, bar_full + 8u * ps, st, 0, x0 + dx, y0 + dy, simg, 0);
                    tma_load_5d(am, bar_full + 8u * ps, st + TC_A_BYTES, 0, x0 + dx, y0 + dy, simg, 1);
                    if (!b_done) {
                        tma_load_3d(bmap, bar_full + 8u * ps, st + 2u * TC_A_BYTES, 0, 0, kb);
                        tma_load_3d(bmap, bar_full + 8u * ps, st + 2u * TC_A_BYTES + b_bytes, 0, 0, 18 + kb);
                    }
                    if (++ps == (uint32_t)a.stages) { ps = 0; pph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            uint32_t ms = 0, mph = 0;
            for (int p = 0; p < n_phases; ++p) {
                const int npad = (p & 1) == 0 ? 128 : 64;
                const uint32_t b_bytes = (uint32_t)npad * 128u;
                const uint32_t idesc = umma_idesc(TC_BLOCK_M, npad);
                const uint32_t acc = tmem_base + (uint32_t)(p & 1) * 128u;   // phase p's epilogue reads this one while p+1 fills the other
                for (int kb = 0; kb < 18; ++kb) {
                    mbar_wait(bar_full + 8u * ms, mph);
                    tc_fence_after();
                    const uint32_t st = smem_base + ms * STAGE;
                    const uint32_t a_hi = st, a_lo = st + TC_A_BYTES, b_hi = st + 2u * TC_A_BYTES, b_lo = b_hi + b_bytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t dah = umma_desc(umma_desc_lo(a_hi) + 2u * k, UMMA_HI_1024), dal = umma_desc(umma_desc_lo(a_lo) + 2u * k, UMMA_HI_1024);
                        const uint64_t dbh = umma_desc(umma_desc_lo(b_hi) + 2u * k, UMMA_HI_1024), dbl = umma_desc(umma_desc_lo(b_lo) + 2u * k, UMMA_HI_1024);
                        umma_bf16(acc, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_bf16(acc, dah, dbl, idesc, 1u);
                        umma_bf16(acc, dah, dbh, idesc, 1u);
                    }
                    umma_commit(bar_empty + 8u * ms);
                    if (++ms == (uint32_t)a.stages) { ms = 0; mph ^= 1u; }
                }
                umma_commit(bar_accum + 8u * (uint32_t)(p & 1));
            }
        }
    } else {
        // accumulator p&1 is overwritten by phase p+2, whose MMAs only start after phase p+1's state-side loads, i.e.
        // after the grid barrier that this CTA's epilogue of phase p has already arrived on: no extra hand-shake needed
        for (int p = 0; p < n_phases; ++p) {
            GcPre pre;
            gc_prefetch(a, p & 1, p >> 1, img, y0, x0, warp, lane, pre);     // h / z operands: in flight during the main loop
            mbar_wait_backoff(bar_accum + 8u * (uint32_t)(p & 1), (uint32_t)((p >> 1) & 1));
            tc_fence_after();
            gc_epilogue_pre(a, p & 1, p >> 1, img, y0, x0, warp, lane, tmem_base + (uint32_t)(p & 1) * 128u, pre);
            tc_fence_before();
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");            // the eight epilogue warps
            // one release for the CTA: the named barrier orders the other threads' stores before this thread's gpu-scope fence
            // (cumulativity -- the pattern of cooperative-groups grid.sync), so 255 threads skip their own membar.gl
            if (warp == 2 && lane == 0) {
                __threadfence();
                if (cl) { for (int r = 0; r < a.cluster; ++r) mbar_arrive_remote(bar_phase, (uint32_t)r); }
                else atomicAdd(bar_ctr, 1u);
            }
        }
    }

    tc_fence_before();
    if (cl) pair_sync(); else __syncthreads();                    // nobody leaves while a peer may still arrive on its phase barrier
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ------------------------------------------------------------------------------------------------
// Re-scheduled pipelined variant (default when one tile per CTA fits): the update gate z is only needed for the final blend,
// so the step is split differently from the reference's (update|reset), (candidate) order:
//   part A   r = sigmoid(conv_r(x, h))                       N = 64: 9 state-side K-blocks on the critical path (was N = 128)
//            epilogue: h * r (64 columns instead of 128)      -> barrier
//   part B   z = sigmoid(conv_z(x, h)),  o = tanh(conv_o(x, h * r))      two N = 64 GEMMs into adjacent TMEM columns
//            z's 18 K-blocks and o's 9 x-side K-blocks depend only on data that was complete before part A's barrier: the producer
//            / MMA threads run them while the epilogue warps of part A work and the barrier is pending; only o's 9 (h * r)-side
//            K-blocks wait.  epilogue: h' = h (1 - z) + o z straight from the two accumulators -- z never leaves the SM (the fp32
//            z buffer round trip is gone).
// Per output element the sums are those of k_gru_chain_pipe (a GEMM's columns are independent), so the results are
// bit-identical.  Stages are 48 KB (A 32 KB + two 8 KB weight planes): four fit.  ESR_GRU_RZO=0 selects the older kernel.
// ------------------------------------------------------------------------------------------------
// out_tma epilogue helpers: a warp's 32 pixels x 32 channels -> swizzled shared-memory rows -> one TMA store per plane, and the
// wait for their COMPLETION (the values are read by other CTAs' TMA loads after the phase barrier)
__device__ __forceinline__ void gc_stage_split32(uint32_t stg, int lane, const float (&x)[32])
{
    uint32_t hw[16], lw[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) split_pack2(x[2 * e], x[2 * e + 1], hw[e], lw[e]);
    const uint32_t row = stg + (uint32_t)lane * 64u, sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t o = row + (((uint32_t)q ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o), "r"(hw[4 * q]), "r"(hw[4 * q + 1]), "r"(hw[4 * q + 2]), "r"(hw[4 * q + 3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + 32u * 64u), "r"(lw[4 * q]), "r"(lw[4 * q + 1]), "r"(lw[4 * q + 2]), "r"(lw[4 * q + 3]) : "memory");
    }
}
__device__ __forceinline__ void gc_tma_store_wait(const CUtensorMap *map, uint32_t stg, int lane, int c0, int x0, int yq, int img)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        tma_store_5d(map, stg, c0, x0, yq, img, 0);
        tma_store_5d(map, stg + 32u * 64u, c0, x0, yq, img, 1);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
}

struct GruRzoArgs {
    GruChainArgs g;
    CUtensorMap bmap_zr64;                      // the update|reset pack with a 64-row box: rows [0, 64) = update (z), [64, 128) = reset (r)
    CUtensorMap omap_rh, omap_hs;               // out_tma: store side of rh / the state slots, box (32 ch, TW, 32 / TW, 1, 1), SWIZZLE_64B
    int out_tma;                                // 1: h * r and h' leave through shared memory + TMA stores (see tc_conv_halo.cu)
    int tiles_per_cta;                          // k_gru_chain_mt: consecutive tiles owned by one CTA (<= 4)
    int warp_arrive;                            // k_gru_chain_rzo with out_tma: per-warp barrier arrivals (ESR_GRU_WARP_ARRIVE=0: one per CTA behind a CTA barrier + membar)
    int sched;                                  // k_gru_chain_rzo: 1 = candidate's x-side issued with the NEXT step's reset x-side (see the kernel)
    long long *trace;                           // ESR_GRU_TRACE: [nsteps][8] clock stamps of CTA 0 (k_gru_chain_rzo<true>)
};
constexpr uint32_t GC_STG_PLANE = 32 * 64;      // one epilogue warp's 32 pixels x 32 channels of one split plane
constexpr uint32_t GC_STG_BYTES = 8 * 2 * GC_STG_PLANE;

// TRACE: clock stamps of CTA 0 per step (ESR_GRU_TRACE=<file>), a separate instantiation (see tc_conv_halo.cu).  Columns per step:
// 0 MMA thread: reset-side GEMM issued, 1 before / 2 after its wait for the first candidate-side (h * r) K-block, 3 candidate-side issued,
// 4 epilogue (warp 2): reset accumulator ready, 5 h * r stored and complete, 6 barrier arrival done, 7 producer: barrier 2g seen
template <bool TRACE>
__global__ void __launch_bounds__(GC_THREADS, 1) k_gru_chain_rzo(const __grid_constant__ GruRzoArgs aa)
{
    const GruChainArgs &a = aa.g;
    long long *const tr = (TRACE && blockIdx.x == 0) ? aa.trace : nullptr;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t B_BYTES = 64u * 128u;                              // one plane of a 64-row weight tile
    constexpr uint32_t STAGE = 2u * TC_A_BYTES + 2u * B_BYTES;            // 48 KB
    const uint32_t stg_base = smem_base + (uint32_t)a.stages * STAGE;         // out_tma: [8 epilogue warps][2 planes][32 px x 64 B]
    const uint32_t bar_base = stg_base + (aa.out_tma ? GC_STG_BYTES : 0u);
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * a.stages, bar_accum = bar_base + 16u * a.stages;   // [0] = r, [1] = z|o
    const uint32_t tmem_slot = bar_accum + 16u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B2 = 2 * a.B;
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    const int tile = blockIdx.x;                                  // grid == number of tiles
    const int img = tile / tiles_per_img;
    const int trem = tile - img * tiles_per_img;
    const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
    const int bb = img < a.B ? img : img - a.B;
    const bool per_img = B2 <= 64;
    unsigned int *bar_ctr = a.barrier + (per_img ? img * 8 : 0);
    const bool warp_arrive = aa.out_tma && aa.warp_arrive;               // each epilogue warp releases the phase barrier itself (8 arrivals per CTA)
    const unsigned int bar_n = (per_img ? (unsigned int)tiles_per_img : gridDim.x) * (warp_arrive ? 8u : 1u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        mbar_init(bar_accum, 1); mbar_init(bar_accum + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);                    // r: columns [0, 64); z: [64, 128); o: [128, 192)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // the six K-segments of a step, in issue order: {r: x, h}, {z: x, h}, {o: x, h*r}
    //   seg     weights            A operand      waits for barrier
    //   0 r.x   zr64 row 64        xc             -
    //   1 r.h   zr64 row 64        hs[g]          2g - 1   (state of the previous step; none at g = 0)
    //   2 z.x   zr64 row 0         xc             -
    //   3 z.h   zr64 row 0         hs[g]          -        (same data as segment 1)
    //   4 o.x   go                 xc             -
    //   5 o.rh  go                 rh             2g       (this step's h * r)
    if (warp == 0) {
        if (elect_one_sync()) {
            uint32_t ps = 0, pph = 0;
            auto do_seg = [&](int seg, int g) {
                {
                    const int w_idx = g / a.N, s_idx = g - w_idx * a.N;
                    const int xc_img = (w_idx * a.B + bb) * a.N + (img < a.B ? s_idx : a.N - 1 - s_idx);
                    const CUtensorMap *bmap = seg < 4 ? &aa.bmap_zr64 : &a.bmap_go;
                    const int brow = seg < 2 ? 64 : 0;
                    const int kb0 = (seg & 1) ? 9 : 0;
                    const CUtensorMap *am = (seg & 1) == 0 ? &a.amap_xc : (seg == 5 ? &a.amap_rh : &a.amap_hs);
                    const int simg = (seg & 1) == 0 ? xc_img : (seg == 5 ? img : g * B2 + img);
                    const int wait_bar = seg == 1 ? 2 * g - 1 : (seg == 5 ? 2 * g : -1);
                    int pre = 0;
                    if (wait_bar >= 0) {
                        // the weight tiles do not depend on the state: those of the first stages go out before the barrier is awaited
                        pre = a.wpre ? (a.stages < 9 ? a.stages : 9) : 0;
                        uint32_t qs = ps, qph = pph;
                        for (int j = 0; j < pre; ++j) {
                            mbar_wait(bar_empty + 8u * qs, qph ^ 1u);
                            mbar_expect_tx(bar_full + 8u * qs, STAGE);
                            const uint32_t stq = smem_base + qs * STAGE;
                            tma_load_3d(bmap, bar_full + 8u * qs, stq + 2u * TC_A_BYTES, 0, brow, kb0 + j);
                            tma_load_3d(bmap, bar_full + 8u * qs, stq + 2u * TC_A_BYTES + B_BYTES, 0, brow, 18 + kb0 + j);
                            if (++qs == (uint32_t)a.stages) { qs = 0; qph ^= 1u; }
                        }
                        const unsigned int target = (unsigned int)(wait_bar + 1) * bar_n;
                        unsigned int v;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar_ctr) : "memory");
                                                } while (v < target);
                        asm volatile("fence.proxy.async;" ::: "memory");
                        if (tr && seg == 5) tr[g * 8 + 7] = clock64();
                    }
                    for (int t = 0; t < 9; ++t) {
                        const int dy = t / 3 - 1, dx = t % 3 - 1;
                        const uint32_t st = smem_base + ps * STAGE;
                        const bool b_done = t < pre;
                        if (!b_done) {
                            mbar_wait(bar_empty + 8u * ps, pph ^ 1u);
                            mbar_expect_tx(bar_full + 8u * ps, STAGE);
                        }
                        tma_load_5d(am, bar_full + 8u * ps, st, 0, x0 + dx, y0 + dy, simg, 0);
                        tma_load_5d(am, bar_full + 8u * ps, st + TC_A_BYTES, 0, x0 + dx, y0 + dy, simg, 1);
                        if (!b_done) {
                            tma_load_3d(bmap, bar_full + 8u * ps, st + 2u * TC_A_BYTES, 0, brow, kb0 + t);
                            tma_load_3d(bmap, bar_full + 8u * ps, st + 2u * TC_A_BYTES + B_BYTES, 0, brow, 18 + kb0 + t);
                        }
                        if (++ps == (uint32_t)a.stages) { ps = 0; pph ^= 1u; }
                    }
                }
            };
            // sched 1: between the reset-side GEMM r.h and the candidate-side GEMM o.rh only z.x and z.h are issued (what fits into the
            // epilogue + barrier turnaround between them, profiles/r2_trace_gru.csv); the candidate's x-side moves next to the NEXT
            // step's reset x-side, into the turnaround after o.rh, and writes the other of two candidate accumulators.
            if (aa.sched) {
                do_seg(0, 0); do_seg(4, 0);
                for (int g = 0; g < a.nsteps; ++g) {
                    do_seg(1, g); do_seg(2, g); do_seg(3, g); do_seg(5, g);
                    if (g + 1 < a.nsteps) { do_seg(0, g + 1); do_seg(4, g + 1); }
                }
            } else {
                for (int g = 0; g < a.nsteps; ++g)
                    for (int seg = 0; seg < 6; ++seg) do_seg(seg, g);
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc(TC_BLOCK_M, 64);
            uint32_t ms = 0, mph = 0;
            auto do_seg = [&](int seg, int g) {
                {
                    const uint32_t acc = tmem_base + (uint32_t)(seg >> 1) * 64u + ((aa.sched && seg >= 4) ? (uint32_t)(g & 1) * 64u : 0u);   // r | z | o (| o')
                    for (int t = 0; t < 9; ++t) {
                        if (tr && seg == 5 && t == 0) tr[g * 8 + 1] = clock64();
                        mbar_wait(bar_full + 8u * ms, mph);
                        tc_fence_after();
                        if (tr && seg == 5 && t == 0) tr[g * 8 + 2] = clock64();
                        const uint32_t st = smem_base + ms * STAGE;
                        // descriptors as base + immediates (see tc_conv_halo.cu): this thread's instruction stream is on the critical path
                        constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
                        const uint32_t ah0 = umma_desc_lo(st), al0 = ah0 + (TC_A_BYTES >> 4), bh0 = ah0 + (2u * TC_A_BYTES >> 4), bl0 = bh0 + (B_BYTES >> 4);
                        const uint32_t first = ((seg & 1) == 0 && t == 0) ? 0u : 1u;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t dah = umma_desc(ah0 + 2u * k, DESC_HI), dal = umma_desc(al0 + 2u * k, DESC_HI);
                            const uint64_t dbh = umma_desc(bh0 + 2u * k, DESC_HI), dbl = umma_desc(bl0 + 2u * k, DESC_HI);
                            umma_bf16(acc, dal, dbh, idesc, k == 0 ? first : 1u);
                            umma_bf16(acc, dah, dbl, idesc, 1u);
                            umma_bf16(acc, dah, dbh, idesc, 1u);
                        }
                        umma_commit(bar_empty + 8u * ms);
                        if (++ms == (uint32_t)a.stages) { ms = 0; mph ^= 1u; }
                    }
                    if (seg == 1) { umma_commit(bar_accum); if (tr) tr[g * 8 + 0] = clock64(); }                 // r complete
                    if (seg == 5) { umma_commit(bar_accum + 8u); if (tr) tr[g * 8 + 3] = clock64(); }            // z and o complete
                }
            };
            if (aa.sched) {
                do_seg(0, 0); do_seg(4, 0);
                for (int g = 0; g < a.nsteps; ++g) {
                    do_seg(1, g); do_seg(2, g); do_seg(3, g); do_seg(5, g);
                    if (g + 1 < a.nsteps) { do_seg(0, g + 1); do_seg(4, g + 1); }
                }
            } else {
                for (int g = 0; g < a.nsteps; ++g)
                    for (int seg = 0; seg < 6; ++seg) do_seg(seg, g);
            }
        }
    } else {
        // ===================== epilogue: 8 warps, two per TMEM lane quadrant, 32 channels each =====================
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int m = quad * 32 + lane;
        const int y = y0 + m / a.TW, x = x0 + m % a.TW;
        const bool valid = (y < a.H) && (x < a.W);
        const size_t pix = ((size_t)img * a.H + (valid ? y : 0)) * a.W + (valid ? x : 0);
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int c0 = half * 32;                                            // this warp's channels
        const uint32_t stg_w = stg_base + (uint32_t)(warp - 2) * (2u * GC_STG_PLANE);
        for (int g = 0; g < a.nsteps; ++g) {
            const uint32_t par = (uint32_t)(g & 1);
            const __nv_bfloat16 *h_prev = a.hs + ((size_t)g * B2 * a.H * a.W + pix) * 64 + c0;
            // h of this pixel (previous step's state: final before this step's first barrier) -- in flight during the main loops.
            // L2 loads (ld.cg): with per-warp arrivals no CTA barrier separates this read from the OTHER half-warp's store of the same
            // 128-byte line, and an L1 copy taken now could serve that warp's own read a stale half later
            uint4 hh[4], hl[4];
            if (valid) {
                const uint4 *ph = reinterpret_cast<const uint4 *>(h_prev), *pl = reinterpret_cast<const uint4 *>(h_prev + a.hs_plane);
#pragma unroll
                for (int q = 0; q < 4; ++q) { hh[q] = __ldcg(ph + q); hl[q] = __ldcg(pl + q); }
            }
            float h[32];
            // ---- part A: r -> h * r
            mbar_wait_backoff(bar_accum, par);
            tc_fence_after();
            if (tr && warp == 2 && lane == 0) tr[g * 8 + 4] = clock64();
            {
                uint32_t raw[32];
                tmem_ld32(taddr + (uint32_t)c0, raw);
                if (valid) {
                    gc_unpack32(hh, hl, h);
                    float v[32];
                    const float4 *bp = reinterpret_cast<const float4 *>(a.bias_zr + 64 + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = bp[q];
                        v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b.x;
                        v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
                        v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z;
                        v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
                    }
                    act32(v, ACT_SIGMOID);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= h[j];
                    if (aa.out_tma) gc_stage_split32(stg_w, lane, v);
                    else store_split32(a.rh + pix * 64 + c0, a.rh_plane, v);
                }
                __syncwarp();
                if (aa.out_tma) gc_tma_store_wait(&aa.omap_rh, stg_w, lane, c0, x0, y0 + quad * (32 / a.TW), img);
            }
            if (tr && warp == 2 && lane == 0) tr[g * 8 + 5] = clock64();
            tc_fence_before();
            if (warp_arrive) {
                // this warp's h * r went out as bulk stores that gc_tma_store_wait() has seen COMPLETE, and its TMEM reads are done:
                // nothing of the other warps is needed for its arrival (no CTA barrier, no membar in the turnaround)
                if (lane == 0) { atomicAdd(bar_ctr, 1u); if (tr && warp == 2) tr[g * 8 + 6] = clock64(); }
                __syncwarp();
            } else {
                asm volatile("fence.proxy.async;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (warp == 2 && lane == 0) { __threadfence(); atomicAdd(bar_ctr, 1u); if (tr) tr[g * 8 + 6] = clock64(); }
            }
            // ---- part B: z, o -> h' = h (1 - z) + o z
            mbar_wait_backoff(bar_accum + 8u, par);
            tc_fence_after();
            {
                uint32_t rz[32], ro[32];
                tmem_ld32(taddr + 64u + (uint32_t)c0, rz);
                tmem_ld32(taddr + 128u + (aa.sched ? (uint32_t)(g & 1) * 64u : 0u) + (uint32_t)c0, ro);
                if (valid) {
                    float z[32], o[32];
                    const float4 *bz = reinterpret_cast<const float4 *>(a.bias_zr + c0), *bo = reinterpret_cast<const float4 *>(a.bias_go + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b1 = bz[q], b2 = bo[q];
                        z[4 * q + 0] = __uint_as_float(rz[4 * q + 0]) + b1.x; o[4 * q + 0] = __uint_as_float(ro[4 * q + 0]) + b2.x;
                        z[4 * q + 1] = __uint_as_float(rz[4 * q + 1]) + b1.y; o[4 * q + 1] = __uint_as_float(ro[4 * q + 1]) + b2.y;
                        z[4 * q + 2] = __uint_as_float(rz[4 * q + 2]) + b1.z; o[4 * q + 2] = __uint_as_float(ro[4 * q + 2]) + b2.z;
                        z[4 * q + 3] = __uint_as_float(rz[4 * q + 3]) + b1.w; o[4 * q + 3] = __uint_as_float(ro[4 * q + 3]) + b2.w;
                    }
                    act32(z, ACT_SIGMOID);
                    act32(o, ACT_TANH);
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] = h[j] * (1.0f - z[j]) + o[j] * z[j];
                    __nv_bfloat16 *h_new = a.hs + ((size_t)(g + 1) * B2 * a.H * a.W + pix) * 64 + c0;
                    if (aa.out_tma) gc_stage_split32(stg_w, lane, o);
                    else store_split32(h_new, a.hs_plane, o);
                }
                __syncwarp();
                if (aa.out_tma) gc_tma_store_wait(&aa.omap_hs, stg_w, lane, c0, x0, y0 + quad * (32 / a.TW), (g + 1) * B2 + img);
            }
            tc_fence_before();
            if (warp_arrive) {
                if (lane == 0) atomicAdd(bar_ctr, 1u);
                __syncwarp();
            } else {
                asm volatile("fence.proxy.async;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (warp == 2 && lane == 0) { __threadfence(); atomicAdd(bar_ctr, 1u); }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ------------------------------------------------------------------------------------------------
// k_gru_chain_x3: the state-independent x-side of all three gates as ONE N = 192 GEMM per step.
//
// Every kernel above loads and reads the x tile once per gate (k_gru_chain_rzo: three times, 27 of its 54 K-blocks per step).  The
// tensor-core convs and this chain are bound by bytes through the shared-memory pipeline (TMA fill + SS-mode operand reads), so:
//   X(g)   9 K-blocks, A = x,      B = [z; r; o] rows (zr pack + go pack, adjacent in the stage),  N = 192 -> accumulators z | r | o
//   H(g)   9 K-blocks, A = h,      B = [z; r] (zr pack, state-side K-blocks),                        N = 128 -> += z | r
//   RH(g)  9 K-blocks, A = h * r,  B = o (go pack, state-side K-blocks),                            N = 64  -> += o
// 27 K-blocks and 4.3 MB through the pipeline per step and CTA instead of 54 and 6.5 MB; per output element the products are still
// accumulated x taps 0..8 then state taps 0..8, so the results are bit-identical to the other kernels.
// X(g + 1) does not depend on the state: it is issued between H(g) and RH(g) -- i.e. while this step's reset epilogue and phase
// barrier are pending -- into the OTHER of two accumulator sets (2 x 192 TMEM columns).  That set was last read by the part-B
// epilogue of step g - 1, which precedes phase barrier 2g - 1, which H(g)'s operand loads wait for: no extra synchronisation.
// Stages are 80 KB (A 32 KB + B 2 x 24 KB): two fit beside the 32 KB store staging.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GC_THREADS, 1) k_gru_chain_x3(const __grid_constant__ GruRzoArgs aa)
{
    const GruChainArgs &a = aa.g;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t B_PLANE = 192u * 128u;                             // [z; r; o] rows of one plane: 24 KB
    constexpr uint32_t STAGE = 2u * TC_A_BYTES + 2u * B_PLANE;            // 80 KB
    constexpr uint32_t NST = 2;
    const uint32_t stg_base = smem_base + NST * STAGE;                    // out_tma: [8 epilogue warps][2 planes][32 px x 64 B]
    const uint32_t bar_base = stg_base + (aa.out_tma ? GC_STG_BYTES : 0u);
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * NST, bar_accum = bar_base + 16u * NST;   // [0] = z|r, [1] = o
    const uint32_t tmem_slot = bar_accum + 16u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B2 = 2 * a.B;
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    const int tile = blockIdx.x;                                  // grid == number of tiles
    const int img = tile / tiles_per_img;
    const int trem = tile - img * tiles_per_img;
    const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
    const int bb = img < a.B ? img : img - a.B;
    const bool per_img = B2 <= 64;
    unsigned int *bar_ctr = a.barrier + (per_img ? img * 8 : 0);
    const unsigned int bar_n = per_img ? (unsigned int)tiles_per_img : gridDim.x;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < NST; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        mbar_init(bar_accum, 1); mbar_init(bar_accum + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);                    // two sets of z [0, 64) | r [64, 128) | o [128, 192)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // issue order (producer and MMA thread alike): X(0), then per step g: H(g), X(g + 1), RH(g)
    if (warp == 0) {
        if (elect_one_sync()) {
            uint32_t ps = 0, pph = 0;
            auto xc_of = [&](int g) {
                const int w_idx = g / a.N, s_idx = g - w_idx * a.N;
                return (w_idx * a.B + bb) * a.N + (img < a.B ? s_idx : a.N - 1 - s_idx);
            };
            auto wait_phase = [&](int idx) {
                const unsigned int target = (unsigned int)(idx + 1) * bar_n;
                unsigned int v;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar_ctr) : "memory");
                                } while (v < target);
                asm volatile("fence.proxy.async;" ::: "memory");
            };
            // kind 0: X, 1: H, 2: RH
            auto segment = [&](int kind, const CUtensorMap *am, int simg) {
                const uint32_t bytes = 2u * TC_A_BYTES + (kind == 0 ? 2u * B_PLANE : kind == 1 ? 2u * 128u * 128u : 2u * 64u * 128u);
                for (int t = 0; t < 9; ++t) {
                    const int dy = t / 3 - 1, dx = t % 3 - 1;
                    const uint32_t st = smem_base + ps * STAGE, bf = bar_full + 8u * ps;
                    mbar_wait(bar_empty + 8u * ps, pph ^ 1u);
                    mbar_expect_tx(bf, bytes);
                    tma_load_5d(am, bf, st, 0, x0 + dx, y0 + dy, simg, 0);
                    tma_load_5d(am, bf, st + TC_A_BYTES, 0, x0 + dx, y0 + dy, simg, 1);
                    const uint32_t bh = st + 2u * TC_A_BYTES, bl = bh + B_PLANE;
                    if (kind == 0) {                  // [z; r] from the zr pack, o from the go pack, x-side K-block t
                        tma_load_3d(&a.bmap_zr, bf, bh, 0, 0, t);
                        tma_load_3d(&a.bmap_zr, bf, bl, 0, 0, 18 + t);
                        tma_load_3d(&a.bmap_go, bf, bh + 128u * 128u, 0, 0, t);
                        tma_load_3d(&a.bmap_go, bf, bl + 128u * 128u, 0, 0, 18 + t);
                    } else if (kind == 1) {           // [z; r], state-side K-block 9 + t
                        tma_load_3d(&a.bmap_zr, bf, bh, 0, 0, 9 + t);
                        tma_load_3d(&a.bmap_zr, bf, bl, 0, 0, 18 + 9 + t);
                    } else {                          // o, state-side K-block 9 + t
                        tma_load_3d(&a.bmap_go, bf, bh, 0, 0, 9 + t);
                        tma_load_3d(&a.bmap_go, bf, bl, 0, 0, 18 + 9 + t);
                    }
                    if (++ps == NST) { ps = 0; pph ^= 1u; }
                }
            };
            segment(0, &a.amap_xc, xc_of(0));
            for (int g = 0; g < a.nsteps; ++g) {
                if (g > 0) wait_phase(2 * g - 1);                             // h of the previous step
                segment(1, &a.amap_hs, g * B2 + img);
                if (g + 1 < a.nsteps) segment(0, &a.amap_xc, xc_of(g + 1));
                wait_phase(2 * g);                                            // this step's h * r
                segment(2, &a.amap_rh, img);
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            const uint32_t id192 = umma_idesc(TC_BLOCK_M, 192), id128 = umma_idesc(TC_BLOCK_M, 128), id64 = umma_idesc(TC_BLOCK_M, 64);
            uint32_t ms = 0, mph = 0;
            auto segment = [&](uint32_t acc, uint32_t idesc, bool fresh) {
#pragma unroll 1
                for (int t = 0; t < 9; ++t) {
                    mbar_wait(bar_full + 8u * ms, mph);
                    tc_fence_after();
                    const uint32_t ah0 = umma_desc_lo(smem_base + ms * STAGE), al0 = ah0 + (TC_A_BYTES >> 4);
                    const uint32_t bh0 = ah0 + (2u * TC_A_BYTES >> 4), bl0 = bh0 + (B_PLANE >> 4);
                    const uint32_t first = (fresh && t == 0) ? 0u : 1u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t dah = umma_desc(ah0 + 2u * k, UMMA_HI_1024), dal = umma_desc(al0 + 2u * k, UMMA_HI_1024);
                        const uint64_t dbh = umma_desc(bh0 + 2u * k, UMMA_HI_1024), dbl = umma_desc(bl0 + 2u * k, UMMA_HI_1024);
                        umma_bf16(acc, dal, dbh, idesc, k == 0 ? first : 1u);
                        umma_bf16(acc, dah, dbl, idesc, 1u);
                        umma_bf16(acc, dah, dbh, idesc, 1u);
                    }
                    umma_commit(bar_empty + 8u * ms);
                    if (++ms == NST) { ms = 0; mph ^= 1u; }
                }
            };
            segment(tmem_base, id192, true);
            for (int g = 0; g < a.nsteps; ++g) {
                const uint32_t set = tmem_base + (uint32_t)(g & 1) * 192u, other = tmem_base + (uint32_t)((g + 1) & 1) * 192u;
                segment(set, id128, false);
                umma_commit(bar_accum);                                       // z and r of step g complete
                if (g + 1 < a.nsteps) segment(other, id192, true);
                segment(set + 128u, id64, false);
                umma_commit(bar_accum + 8u);                                  // o complete
            }
        }
    } else {
        // ===================== epilogue: 8 warps, two per TMEM lane quadrant, 32 channels each =====================
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int m = quad * 32 + lane;
        const int y = y0 + m / a.TW, x = x0 + m % a.TW;
        const bool valid = (y < a.H) && (x < a.W);
        const size_t pix = ((size_t)img * a.H + (valid ? y : 0)) * a.W + (valid ? x : 0);
        const int c0 = half * 32;                                            // this warp's channels
        const uint32_t stg_w = stg_base + (uint32_t)(warp - 2) * (2u * GC_STG_PLANE);
        for (int g = 0; g < a.nsteps; ++g) {
            const uint32_t par = (uint32_t)(g & 1);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + par * 192u;
            const __nv_bfloat16 *h_prev = a.hs + ((size_t)g * B2 * a.H * a.W + pix) * 64 + c0;
            // h of this pixel (previous step's state: final before this step's first barrier) -- in flight during the main loops
            uint4 hh[4], hl[4];
            if (valid) {
                const uint4 *ph = reinterpret_cast<const uint4 *>(h_prev), *pl = reinterpret_cast<const uint4 *>(h_prev + a.hs_plane);
#pragma unroll
                for (int q = 0; q < 4; ++q) { hh[q] = ph[q]; hl[q] = pl[q]; }
            }
            float h[32];
            // ---- part A: r -> h * r
            mbar_wait_backoff(bar_accum, par);
            tc_fence_after();
            {
                uint32_t raw[32];
                tmem_ld32(taddr + 64u + (uint32_t)c0, raw);
                if (valid) {
                    gc_unpack32(hh, hl, h);
                    float v[32];
                    const float4 *bp = reinterpret_cast<const float4 *>(a.bias_zr + 64 + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = bp[q];
                        v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b.x;
                        v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
                        v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z;
                        v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
                    }
                    act32(v, ACT_SIGMOID);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= h[j];
                    if (aa.out_tma) gc_stage_split32(stg_w, lane, v);
                    else store_split32(a.rh + pix * 64 + c0, a.rh_plane, v);
                }
                __syncwarp();
                if (aa.out_tma) gc_tma_store_wait(&aa.omap_rh, stg_w, lane, c0, x0, y0 + quad * (32 / a.TW), img);
            }
            tc_fence_before();
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (warp == 2 && lane == 0) { __threadfence(); atomicAdd(bar_ctr, 1u); }
            // ---- part B: z, o -> h' = h (1 - z) + o z
            mbar_wait_backoff(bar_accum + 8u, par);
            tc_fence_after();
            {
                uint32_t rz[32], ro[32];
                tmem_ld32(taddr + (uint32_t)c0, rz);
                tmem_ld32(taddr + 128u + (uint32_t)c0, ro);
                if (valid) {
                    float z[32], o[32];
                    const float4 *bz = reinterpret_cast<const float4 *>(a.bias_zr + c0), *bo = reinterpret_cast<const float4 *>(a.bias_go + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b1 = bz[q], b2 = bo[q];
                        z[4 * q + 0] = __uint_as_float(rz[4 * q + 0]) + b1.x; o[4 * q + 0] = __uint_as_float(ro[4 * q + 0]) + b2.x;
                        z[4 * q + 1] = __uint_as_float(rz[4 * q + 1]) + b1.y; o[4 * q + 1] = __uint_as_float(ro[4 * q + 1]) + b2.y;
                        z[4 * q + 2] = __uint_as_float(rz[4 * q + 2]) + b1.z; o[4 * q + 2] = __uint_as_float(ro[4 * q + 2]) + b2.z;
                        z[4 * q + 3] = __uint_as_float(rz[4 * q + 3]) + b1.w; o[4 * q + 3] = __uint_as_float(ro[4 * q + 3]) + b2.w;
                    }
                    act32(z, ACT_SIGMOID);
                    act32(o, ACT_TANH);
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] = h[j] * (1.0f - z[j]) + o[j] * z[j];
                    __nv_bfloat16 *h_new = a.hs + ((size_t)(g + 1) * B2 * a.H * a.W + pix) * 64 + c0;
                    if (aa.out_tma) gc_stage_split32(stg_w, lane, o);
                    else store_split32(h_new, a.hs_plane, o);
                }
                __syncwarp();
                if (aa.out_tma) gc_tma_store_wait(&aa.omap_hs, stg_w, lane, c0, x0, y0 + quad * (32 / a.TW), (g + 1) * B2 + img);
            }
            tc_fence_before();
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (warp == 2 && lane == 0) { __threadfence(); atomicAdd(bar_ctr, 1u); }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// k_gru_chain_mt: the chain when there are more tiles than SMs (4x SR configurations: 256 / 512 tiles).
//
// k_gru_chain above serialises main loop and epilogue per tile (one 128-column accumulator, a block barrier per tile), stores z in
// fp32 and reads z and h back with one-pixel-per-lane accesses: at cfg4 the epilogue's strided loads and stores alone are 23 % of its
// 4.97 ms (ESR_GRU_DIAG).  Here a CTA owns up to FOUR tiles and gives each its own 128 TMEM columns for the whole step:
//     columns [128 t, 128 t + 64)      z of tile t  -- written in phase 1, read in phase 2, never leaves the SM
//     columns [128 t + 64, 128 t + 128) r in phase 1, then the candidate o in phase 2 (r is dead once h * r is stored)
// so the MMA thread runs ahead into the next tile while the eight epilogue warps drain the previous one, and no barrier separates the
// tiles of a phase.  h arrives through TMA into a per-warp 4 KB buffer (issued before the accumulator is awaited), h * r and h' leave
// from the same buffer by TMA stores (tc_conv_halo.cu explains why strided per-lane accesses are to be avoided in these kernels).
// Same K-block order and per-element accumulation order as every other chain kernel: bit-identical results.
// ------------------------------------------------------------------------------------------------
constexpr int GC_MT_MAX = 4;

__global__ void __launch_bounds__(GC_THREADS, 1) k_gru_chain_mt(const __grid_constant__ GruRzoArgs aa)
{
    const GruChainArgs &a = aa.g;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t B_MAX = 128u * 128u;                               // one plane of the N = 128 weight tile
    constexpr uint32_t STAGE = 2u * TC_A_BYTES + 2u * B_MAX;              // 64 KB
    const uint32_t stg_base = smem_base + (uint32_t)a.stages * STAGE;     // [8 epilogue warps][2 planes][32 px x 64 B]
    const uint32_t bar_base = stg_base + GC_STG_BYTES;
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * a.stages, bar_accum = bar_base + 16u * a.stages;   // [GC_MT_MAX]
    const uint32_t bar_hload = bar_accum + 8u * GC_MT_MAX;                // [8]: one per epilogue warp
    const uint32_t tmem_slot = bar_hload + 64u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B2 = 2 * a.B;
    const int tiles_per_img = a.tiles_x * a.tiles_y, n_tiles = B2 * tiles_per_img;
    const int per_cta = aa.tiles_per_cta;
    const int tile0 = blockIdx.x * per_cta;
    const int my_tiles = min(per_cta, n_tiles - tile0);                   // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        for (int t = 0; t < GC_MT_MAX; ++t) mbar_init(bar_accum + 8u * t, 1);
        for (int w = 0; w < 8; ++w) mbar_init(bar_hload + 8u * w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    uint32_t ps = 0, pph = 0, ms = 0, mph = 0;
    uint32_t hl_ph = 0;                                                    // parity of this warp's h-load barrier
    for (int p = 0; p < 2 * a.nsteps; ++p) {
        const int g = p >> 1, which = p & 1;              // which: 0 = update|reset gates, 1 = candidate + blend
        const int w_idx = g / a.N, s_idx = g - w_idx * a.N;
        const int npad = which == 0 ? 128 : 64;
        const uint32_t b_bytes = (uint32_t)npad * 128u;
        const uint32_t stage_bytes = 2u * TC_A_BYTES + 2u * b_bytes;
        const CUtensorMap *bmap = which == 0 ? &a.bmap_zr : &a.bmap_go;
        if (warp == 0) {
            if (elect_one_sync()) {
                for (int t = 0; t < my_tiles; ++t) {
                    const int tile = tile0 + t;
                    const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
                    const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
                    const int bb = img < a.B ? img : img - a.B;
                    const int xc_img = (w_idx * a.B + bb) * a.N + (img < a.B ? s_idx : a.N - 1 - s_idx);
                    for (int kb = 0; kb < 18; ++kb) {
                        const int src = kb / 9, tap = kb - src * 9;
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        mbar_wait(bar_empty + 8u * ps, pph ^ 1u);
                        mbar_expect_tx(bar_full + 8u * ps, stage_bytes);
                        const uint32_t st = smem_base + ps * STAGE;
                        const CUtensorMap *am = src == 0 ? &a.amap_xc : (which == 0 ? &a.amap_hs : &a.amap_rh);
                        const int simg = src == 0 ? xc_img : (which == 0 ? g * B2 + img : img);
                        tma_load_5d(am, bar_full + 8u * ps, st, 0, x0 + dx, y0 + dy, simg, 0);
                        tma_load_5d(am, bar_full + 8u * ps, st + TC_A_BYTES, 0, x0 + dx, y0 + dy, simg, 1);
                        tma_load_3d(bmap, bar_full + 8u * ps, st + 2u * TC_A_BYTES, 0, 0, kb);
                        tma_load_3d(bmap, bar_full + 8u * ps, st + 2u * TC_A_BYTES + b_bytes, 0, 0, 18 + kb);
                        if (++ps == (uint32_t)a.stages) { ps = 0; pph ^= 1u; }
                    }
                }
            }
        } else if (warp == 1) {
            if (elect_one_sync()) {
                const uint32_t idesc = umma_idesc(TC_BLOCK_M, npad);
                for (int t = 0; t < my_tiles; ++t) {
                    const uint32_t acc = tmem_base + (uint32_t)t * 128u + (which ? 64u : 0u);
#pragma unroll 1
                    for (int kb = 0; kb < 18; ++kb) {
                        mbar_wait(bar_full + 8u * ms, mph);
                        tc_fence_after();
                        const uint32_t ah0 = umma_desc_lo(smem_base + ms * STAGE), al0 = ah0 + (TC_A_BYTES >> 4);
                        const uint32_t bh0 = ah0 + (2u * TC_A_BYTES >> 4), bl0 = bh0 + (b_bytes >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t dah = umma_desc(ah0 + 2u * k, UMMA_HI_1024), dal = umma_desc(al0 + 2u * k, UMMA_HI_1024);
                            const uint64_t dbh = umma_desc(bh0 + 2u * k, UMMA_HI_1024), dbl = umma_desc(bl0 + 2u * k, UMMA_HI_1024);
                            umma_bf16(acc, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);
                            umma_bf16(acc, dah, dbl, idesc, 1u);
                            umma_bf16(acc, dah, dbh, idesc, 1u);
                        }
                        umma_commit(bar_empty + 8u * ms);
                        if (++ms == (uint32_t)a.stages) { ms = 0; mph ^= 1u; }
                    }
                    umma_commit(bar_accum + 8u * t);
                }
            }
        } else {
            // ===================== epilogue: 8 warps, two per TMEM lane quadrant, 32 channels each =====================
            const int quad = warp & 3, half = (warp - 2) >> 2;
            const int c0 = half * 32;
            const uint32_t stg_w = stg_base + (uint32_t)(warp - 2) * (2u * GC_STG_PLANE);
            const uint32_t hbar = bar_hload + 8u * (uint32_t)(warp - 2);
            const uint32_t row = stg_w + (uint32_t)lane * 64u, sw = ((uint32_t)lane >> 1) & 3u;
            for (int t = 0; t < my_tiles; ++t) {
                const int tile = tile0 + t;
                const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
                const int y0 = (trem / a.tiles_x) * a.TH, x0 = (trem % a.tiles_x) * a.TW;
                const int yq = y0 + quad * (32 / a.TW);
                // h of this warp's 32 pixels x 32 channels (the previous step's state) -> staging buffer, while the tile's MMAs run
                if (lane == 0) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the store that last read the buffer
                    mbar_expect_tx(hbar, 2u * GC_STG_PLANE);
                    tma_load_5d(&aa.omap_hs, hbar, stg_w, c0, x0, yq, g * B2 + img, 0);
                    tma_load_5d(&aa.omap_hs, hbar, stg_w + GC_STG_PLANE, c0, x0, yq, g * B2 + img, 1);
                }
                mbar_wait_backoff(bar_accum + 8u * t, (uint32_t)(p & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)t * 128u;
                uint32_t raw[32];
                tmem_ld32(taddr + 64u + (uint32_t)c0, raw);              // r (phase 1) or o (phase 2)
                mbar_wait(hbar, hl_ph);
                hl_ph ^= 1u;
                float h[32];
                {
                    uint4 hh[4], hl[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t o = row + (((uint32_t)q ^ sw) << 4);
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hh[q].x), "=r"(hh[q].y), "=r"(hh[q].z), "=r"(hh[q].w) : "r"(o));
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hl[q].x), "=r"(hl[q].y), "=r"(hl[q].z), "=r"(hl[q].w) : "r"(o + GC_STG_PLANE));
                    }
                    gc_unpack32(hh, hl, h);
                }
                float v[32];
                if (which == 0) {
                    const float4 *bp = reinterpret_cast<const float4 *>(a.bias_zr + 64 + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = bp[q];
                        v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b.x;
                        v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
                        v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z;
                        v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
                    }
                    act32(v, ACT_SIGMOID);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= h[j];
                } else {
                    uint32_t rz[32];
                    tmem_ld32(taddr + (uint32_t)c0, rz);                 // z, still where phase 1 left it
                    float z[32];
                    const float4 *bz = reinterpret_cast<const float4 *>(a.bias_zr + c0), *bo = reinterpret_cast<const float4 *>(a.bias_go + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b1 = bz[q], b2 = bo[q];
                        z[4 * q + 0] = __uint_as_float(rz[4 * q + 0]) + b1.x; v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b2.x;
                        z[4 * q + 1] = __uint_as_float(rz[4 * q + 1]) + b1.y; v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b2.y;
                        z[4 * q + 2] = __uint_as_float(rz[4 * q + 2]) + b1.z; v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b2.z;
                        z[4 * q + 3] = __uint_as_float(rz[4 * q + 3]) + b1.w; v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b2.w;
                    }
                    act32(z, ACT_SIGMOID);
                    act32(v, ACT_TANH);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = h[j] * (1.0f - z[j]) + v[j] * z[j];
                }
                __syncwarp();                                             // every lane has read its h row
                gc_stage_split32(stg_w, lane, v);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const CUtensorMap *om = which == 0 ? &aa.omap_rh : &aa.omap_hs;
                    const int oimg = which == 0 ? img : (g + 1) * B2 + img;
                    tma_store_5d(om, stg_w, c0, x0, yq, oimg, 0);
                    tma_store_5d(om, stg_w + GC_STG_PLANE, c0, x0, yq, oimg, 1);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                tc_fence_before();
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // complete before the phase barrier releases
        }
        gc_grid_barrier(a.barrier, (unsigned int)(p + 1) * gridDim.x);
        tc_fence_after();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
struct GruChainPlan {
    GruChainArgs args;
    int grid;
    size_t smem;
    bool pipelined;
    bool rzo = false;               // k_gru_chain_rzo (reset gate first; update gate and candidate together, off the critical path)
    GruRzoArgs rzo_args;
    size_t rzo_smem = 0;
    bool mt = false;                // k_gru_chain_mt (more tiles than SMs); ESR_GRU_MT=0: k_gru_chain
    int mt_grid = 0;
    size_t mt_smem = 0;
    bool x3 = false;                // k_gru_chain_x3 (x-side of the three gates as one N = 192 GEMM); ESR_GRU_X3=1
    size_t x3_smem = 0;
};

int gru_chain_prepare(const SplitTensor &xc, const SplitTensor &hs, const SplitTensor &rh, float *zbuf, const void *w_zr,
                      const float *b_zr, const void *w_go, const float *b_go, unsigned int *barrier, int B, int N,
                      int nsteps, void **plan_out)
{
    GruChainPlan *p = new GruChainPlan();
    GruChainArgs &a = p->args;
    memset(&a, 0, sizeof(a));
    const int H = xc.H, W = xc.W;
    int TW = W >= 24 ? 32 : (W >= 12 ? 16 : 8);
    int TH = TC_BLOCK_M / TW;
    int rc;
    if ((rc = tc_make_amap(xc, TW, TH, &a.amap_xc)) || (rc = tc_make_amap(hs, TW, TH, &a.amap_hs)) ||
        (rc = tc_make_amap(rh, TW, TH, &a.amap_rh)) || (rc = tc_make_bmap(w_zr, 128, 18, 128, &a.bmap_zr)) ||
        (rc = tc_make_bmap(w_go, 64, 18, 64, &a.bmap_go))) { delete p; return rc; }
    a.bias_zr = b_zr; a.bias_go = b_go;
    a.hs = hs.base; a.hs_plane = hs.plane(); a.rh = rh.base; a.rh_plane = rh.plane(); a.zbuf = zbuf; a.barrier = barrier;
    a.B = B; a.N = N; a.nsteps = nsteps; a.H = H; a.W = W; a.TW = TW; a.TH = TH;
    a.tiles_x = (W + TW - 1) / TW; a.tiles_y = (H + TH - 1) / TH;
    a.stages = 3;
    a.diag = getenv("ESR_GRU_DIAG") ? atoi(getenv("ESR_GRU_DIAG")) : 0;
    a.wpre = getenv("ESR_GRU_WPRE") != nullptr;     // 476 -> 492 us with it (profiles/r2_notes.md): off
    p->smem = 1024 + (size_t)a.stages * (2 * TC_A_BYTES + 2 * 128 * 128) + 16 * a.stages + 96;
    ESR_CUDA_CHECK(cudaFuncSetAttribute(k_gru_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    int per_sm = 0;
    ESR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gru_chain, GC_THREADS, p->smem));
    if (per_sm < 1) { set_error("gru_chain: kernel does not fit on an SM"); delete p; return ESR_EUNSUPPORTED; }
    ESR_CUDA_CHECK(cudaFuncSetAttribute(k_gru_chain_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    const int n_tiles = 2 * B * a.tiles_x * a.tiles_y;
    const int max_grid = dev_info().sm_count * per_sm;
    p->grid = n_tiles < max_grid ? n_tiles : max_grid;
    static const bool no_pipe = getenv("ESR_GRU_NO_PIPE") != nullptr;
    p->pipelined = !no_pipe && n_tiles <= max_grid;            // one tile per CTA: overlap the x-side K-blocks with the epilogue
    // a tile's halo only reaches tiles of its own image: with <= 8 tiles per image each image is ONE thread-block cluster whose
    // CTAs synchronise phases through mbarriers in each other's shared memory (remote arrive + local try_wait: a few hundred
    // cycles) instead of a global-memory counter polled with ld.acquire.gpu (~5 k cycles per phase, profiles/r1_notes.md); the
    // clusters are independent, so no cooperative launch is needed.  ESR_GRU_NO_CLUSTER=1: counter barrier.
    static const bool no_cluster = getenv("ESR_GRU_NO_CLUSTER") != nullptr;
    const int tpi = a.tiles_x * a.tiles_y;
    a.cluster = (p->pipelined && !no_cluster && tpi >= 2 && tpi <= 8) ? tpi : 1;
    if (a.cluster > 1) {
        // every image's cluster must be resident at once, or the images run in waves and the serial chain takes twice as long
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(n_tiles); cfg.blockDim = dim3(GC_THREADS); cfg.dynamicSmemBytes = p->smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)a.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, k_gru_chain_pipe, &cfg) != cudaSuccess) { cudaGetLastError(); max_clusters = 0; }
        if (getenv("ESR_DEBUG")) fprintf(stderr, "[esr] gru_chain: %d clusters of %d CTAs wanted, %d can be resident\n", n_tiles / a.cluster, a.cluster, max_clusters);
        if (max_clusters * a.cluster < n_tiles) a.cluster = 1;
    }
    static const bool mt_off = getenv("ESR_GRU_MT") && atoi(getenv("ESR_GRU_MT")) == 0;
    if (!p->pipelined && !mt_off && 32 % TW == 0) {
        const int per_cta = (n_tiles + dev_info().sm_count - 1) / dev_info().sm_count;
        const size_t smem_mt = 1024 + (size_t)3 * (2 * TC_A_BYTES + 2 * 128 * 128) + GC_STG_BYTES + 16 * 3 + 8 * GC_MT_MAX + 64 + 96;
        if (per_cta <= GC_MT_MAX && smem_mt <= (size_t)dev_info().max_smem_optin &&
            tc_make_omap(rh, TW, 32 / TW, &p->rzo_args.omap_rh) == ESR_OK && tc_make_omap(hs, TW, 32 / TW, &p->rzo_args.omap_hs) == ESR_OK) {
            p->mt = true;
            p->rzo_args.g = a;
            p->rzo_args.g.stages = 3;
            p->rzo_args.tiles_per_cta = per_cta;
            p->mt_grid = (n_tiles + per_cta - 1) / per_cta;
            p->mt_smem = smem_mt;
            ESR_CUDA_CHECK(cudaFuncSetAttribute(k_gru_chain_mt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mt));
        }
    }
    static const bool rzo_off = getenv("ESR_GRU_RZO") && atoi(getenv("ESR_GRU_RZO")) == 0;
    if (p->pipelined && a.cluster <= 1 && !rzo_off) {
        p->rzo = true;
        p->rzo_args.g = a;
        { static const bool sched_off = getenv("ESR_GRU_SCHED") && atoi(getenv("ESR_GRU_SCHED")) == 0; p->rzo_args.sched = sched_off ? 0 : 1; }
        { static const bool wa_off = getenv("ESR_GRU_WARP_ARRIVE") && atoi(getenv("ESR_GRU_WARP_ARRIVE")) == 0; p->rzo_args.warp_arrive = wa_off ? 0 : 1; }
        p->rzo_args.g.stages = 4;
        if ((rc = tc_make_bmap(w_zr, 128, 18, 64, &p->rzo_args.bmap_zr64))) { delete p; return rc; }
        p->rzo_smem = 1024 + (size_t)4 * (2 * TC_A_BYTES + 2 * 64 * 128) + 16 * 4 + 96;
        static const bool no_out_tma = getenv("ESR_TC_NO_OUT_TMA") != nullptr || getenv("ESR_GRU_NO_OUT_TMA") != nullptr;
        if (!no_out_tma && p->rzo_smem + GC_STG_BYTES <= (size_t)dev_info().max_smem_optin && 32 % TW == 0 &&
            tc_make_omap(rh, TW, 32 / TW, &p->rzo_args.omap_rh) == ESR_OK && tc_make_omap(hs, TW, 32 / TW, &p->rzo_args.omap_hs) == ESR_OK) {
            p->rzo_args.out_tma = 1;
            p->rzo_smem += GC_STG_BYTES;
        }
        static const bool x3_on = getenv("ESR_GRU_X3") && atoi(getenv("ESR_GRU_X3")) != 0;
        p->x3_smem = 1024 + (size_t)2 * (2 * TC_A_BYTES + 2 * 192 * 128) + (p->rzo_args.out_tma ? GC_STG_BYTES : 0) + 16 * 2 + 96;
        if (x3_on && p->x3_smem <= (size_t)dev_info().max_smem_optin) {
            p->x3 = true;
            ESR_CUDA_CHECK(cudaFuncSetAttribute(k_gru_chain_x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->x3_smem));
        }
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_gru_chain_rzo<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->rzo_smem));
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_gru_chain_rzo<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->rzo_smem));
    }
    *plan_out = p;
    return ESR_OK;
}

int gru_chain_launch(void *plan, cudaStream_t st)
{
    GruChainPlan *p = (GruChainPlan *)plan;
    ESR_CUDA_CHECK(cudaMemsetAsync(p->args.barrier, 0, 64 * 8 * sizeof(unsigned int), st));   // per-image counters, 32 bytes apart
    if (p->mt) {
        void *kargs[] = {(void *)&p->rzo_args};
        ESR_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_gru_chain_mt, dim3(p->mt_grid), dim3(GC_THREADS), kargs, p->mt_smem, st));
        esr::count_launch();
        return ESR_OK;
    }
    if (p->rzo && p->x3) {
        void *kargs[] = {(void *)&p->rzo_args};
        ESR_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_gru_chain_x3, dim3(p->grid), dim3(GC_THREADS), kargs, p->x3_smem, st));
        esr::count_launch();
        return ESR_OK;
    }
    if (p->rzo) {
        static const char *trace_path = getenv("ESR_GRU_TRACE");          // measurement aid: clock stamps of CTA 0, one row per step
        if (trace_path) {
            static long long *dbuf = nullptr;
            const int ns = p->rzo_args.g.nsteps;
            if (!dbuf) cudaMalloc(&dbuf, sizeof(long long) * 8 * 256);
            cudaMemsetAsync(dbuf, 0, sizeof(long long) * 8 * 256, st);
            GruRzoArgs b = p->rzo_args; b.trace = dbuf;
            void *targs[] = {(void *)&b};
            ESR_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_gru_chain_rzo<true>, dim3(p->grid), dim3(GC_THREADS), targs, p->rzo_smem, st));
            cudaStreamSynchronize(st);
            static long long host[8 * 256];
            cudaMemcpy(host, dbuf, sizeof(host), cudaMemcpyDeviceToHost);
            FILE *f = fopen(trace_path, "w");
            if (f) {
                fprintf(f, "# per step: mma_reset_side_issued, mma_before_wait_rh, mma_after_wait_rh, mma_candidate_issued, epi_reset_acc_ready, epi_rh_stored, epi_barrier_arrived, prod_barrier_seen\n");
                for (int i = 0; i < ns && i < 256; ++i) { for (int c = 0; c < 8; ++c) fprintf(f, "%lld%c", host[i * 8 + c], c == 7 ? '\n' : ','); }
                fclose(f);
            }
            esr::count_launch();
            return ESR_OK;
        }
        void *kargs[] = {(void *)&p->rzo_args};
        ESR_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_gru_chain_rzo<false>, dim3(p->grid), dim3(GC_THREADS), kargs, p->rzo_smem, st));
        esr::count_launch();
        return ESR_OK;
    }
    if (p->pipelined && p->args.cluster > 1) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(p->grid); cfg.blockDim = dim3(GC_THREADS); cfg.dynamicSmemBytes = p->smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)p->args.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        ESR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gru_chain_pipe, p->args));
        esr::count_launch();
        return ESR_OK;
    }
    void *kargs[] = {(void *)&p->args};
    ESR_CUDA_CHECK(cudaLaunchCooperativeKernel(p->pipelined ? (void *)k_gru_chain_pipe : (void *)k_gru_chain, dim3(p->grid),
                                               dim3(GC_THREADS), kargs, p->smem, st));
    esr::count_launch();
    return ESR_OK;
}

void gru_chain_destroy(void *plan) { delete (GruChainPlan *)plan; }

} // namespace esr
