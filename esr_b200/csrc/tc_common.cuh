// tc_common.cuh -- device-side building blocks shared by the tcgen05 convolution kernels (tc_conv.cu, tc_conv3.cu):
// PTX wrappers (mbarrier, TMA, tcgen05 alloc/mma/commit/ld), UMMA descriptors, split-bf16 vector I/O and the fused epilogue.
#pragma once
#include "tc_conv.cuh"

namespace esr {

constexpr int TC_THREADS = 192;
constexpr int TC_BLOCK_M = 128;
constexpr int TC_A_BYTES = TC_BLOCK_M * 128;          // one plane of one A tile: 128 rows x 64 bf16

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp.  Unlike `lane == 0`, elect.sync tells ptxas that exactly one thread runs the guarded code, so the
// uniform-datapath instructions inside (UTCHMMA, UTMALDG, UTCBAR) are emitted directly instead of each being wrapped in an
// ELECT / BRA.U.ANY loop over the "possibly many" active lanes -- measured on B200: 69-74 cycles per tcgen05.mma issued from an
// `if (lane == 0)` block, which made the N <= 64 layers issue-bound (profiles/r2_notes.md).
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}

// ---- thread-block cluster helpers (CTA pairs of tc_conv.cu, per-image clusters of gru_chain.cu)
__device__ __forceinline__ uint32_t pair_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void pair_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta)
{
    asm volatile(
        "{\n\t"
        ".reg .b32 raddr;\n\t"
        "mapa.shared::cluster.u32 raddr, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [raddr];\n\t"
        "}" ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// same, for waits that last microseconds (an MMA / producer thread idling while other warps of the CTA do the real work): back
// off between polls so the spin does not eat the issue slots of the warps it is waiting for
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done) __nanosleep(40);
    } while (!done);
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap *map, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                            int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint32_t bar, uint32_t dst, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
                 " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// one epilogue chunk of accumulator columns [n0, n0 + 32) (16 valid columns when fewer remain: npad is a multiple of 16)
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, int n0, int npad, uint32_t (&raw)[32])
{
    if (npad - n0 >= 32) {
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
}
// "stacked" accumulators (ConvTCArgs::stack): columns [0, npad) hold A_hi.B_hi + A_lo.B_hi, columns [npad, 2 npad) hold A_hi.B_lo
__device__ __forceinline__ void tmem_ld_chunk_stacked(uint32_t taddr, int n0, int npad, uint32_t (&raw)[32])
{
    uint32_t lo[32];
    tmem_ld_chunk(taddr, n0, npad, raw);
    tmem_ld_chunk(taddr + (uint32_t)npad, n0, npad, lo);
#pragma unroll
    for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(lo[j]));
}

// K-major, 128B-swizzled operand tile (rows of 128 bytes, 8-row groups 1024 bytes apart).
// Bits: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=SWIZZLE_128B(2) [61,64)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO = 1024 B
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// The same descriptor in two words: lo = address field | LBO, hi = SBO | version | SWIZZLE_128B (a constant per operand kind).  Advancing
// the operand by n bytes (n a multiple of 16, the result below 256 KB) is lo + n / 16.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
constexpr uint32_t UMMA_HI_1024 = (1024u >> 4) | (1u << 14) | (2u << 29);      // upper word of umma_smem_desc(): SBO = 1024 B, version 1, SWIZZLE_128B
// c=F32 [4,6)=1 | a=BF16 [7,10)=1 | b=BF16 [10,13)=1 | K-major A,B | N>>3 [17,23) | M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float apply_act(float x, int act)
{
    if (act == ACT_RELU) return fmaxf(x, 0.0f);
    if (act == ACT_SIGMOID) return fast_sigmoid(x);
    if (act == ACT_TANH) return fast_tanh(x);
    return x;
}

// 32 consecutive channels of one pixel of a split tensor -> fp32
__device__ __forceinline__ void load_split32(const __nv_bfloat16 *hi_ptr, size_t plane, float (&o)[32])
{
    const uint4 *ph = reinterpret_cast<const uint4 *>(hi_ptr);
    const uint4 *pl = reinterpret_cast<const uint4 *>(hi_ptr + plane);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 h = ph[q], l = pl[q];
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            o[q * 8 + e * 2 + 0] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
            o[q * 8 + e * 2 + 1] = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
        }
    }
}
__device__ __forceinline__ void store_split32(__nv_bfloat16 *hi_ptr, size_t plane, const float (&x)[32])
{
    uint32_t hw[16], lw[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        split_pack2(x[2 * e], x[2 * e + 1], hw[e], lw[e]);
    }
    uint4 *ph = reinterpret_cast<uint4 *>(hi_ptr);
    uint4 *pl = reinterpret_cast<uint4 *>(hi_ptr + plane);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        ph[q] = make_uint4(hw[q * 4], hw[q * 4 + 1], hw[q * 4 + 2], hw[q * 4 + 3]);
        pl[q] = make_uint4(lw[q * 4], lw[q * 4 + 1], lw[q * 4 + 2], lw[q * 4 + 3]);
    }
}

// activation over 32 values with a warp-uniform selector (no per-element branching)
__device__ __forceinline__ void act32(float (&v)[32], int act)
{
    if (act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    } else if (act == ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fast_sigmoid(v[j]);
    } else if (act == ACT_TANH) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fast_tanh(v[j]);
    }
}

// One thread's 32 accumulator columns [n0, n0+32) of one valid output pixel -> global memory.
// residual / activation part of the standard epilogue on 32 biased accumulator values (in place)
__device__ __forceinline__ void epilogue_std_math(const ConvTCArgs &a, float (&v)[32], int n0, int img, int y, int x, bool valid)
{
    // activation selector for this chunk: uniform unless act_from falls inside it (conv_offset_mask: 144 = 4.5 chunks)
    const bool mixed = (a.act_from > n0) && (a.act_from < n0 + 32);
    const int act = (n0 >= a.act_from) ? a.act : ACT_NONE;
    float r[32];
    const bool has_res = valid && a.res_mode != RES_NONE && n0 < a.cout;
    if (has_res) {
        const size_t rpix = ((size_t)(a.res_img ? a.res_img[img] : img) * a.H + y) * a.W + x;
        load_split32(a.res + rpix * a.res_C + n0, a.res_plane, r);
        if (a.res_mode == RES_PRE_ACT) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += r[j];
        }
    }
    if (!mixed) act32(v, act);
    else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (n0 + j >= a.act_from) v[j] = apply_act(v[j], a.act);
    }
    if (has_res && a.res_mode == RES_POST_ACT) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += r[j];
    }
}
// bias + standard epilogue math of one chunk, values only (the caller stores them): the out_tma path
__device__ __forceinline__ void epilogue_values(const ConvTCArgs &a, const uint32_t (&raw)[32], int n0, int img, int y, int x, bool valid,
                                                float (&v)[32])
{
    const float4 *bp = reinterpret_cast<const float4 *>(a.bias + n0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 b = bp[q];
        v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b.x;
        v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
        v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z;
        v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
    }
    epilogue_std_math(a, v, n0, img, y, x, valid);
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3, int c4)
{
    // L2 evict_last: the tensor is the next layer's input and must stay in the 126 MB L2 like the lines of a plain st.global do
    // (without the hint every consumer kernel got 3-5 us slower: its input came from HBM)
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5, %6}], [%1], %7;"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "l"(0x14F0000000000000ull) : "memory");
}

__device__ __forceinline__ void epilogue_chunk(const ConvTCArgs &a, const uint32_t (&raw)[32], int n0, size_t pix, int img,
                                               int y, int x)
{
    float v[32];
    {
        // bias: 16-byte broadcast loads (npad is a multiple of 16, the blob is 256-byte aligned)
        const float4 *bp = reinterpret_cast<const float4 *>(a.bias + n0);
        const int nq = (a.npad - n0 >= 32) ? 8 : 4;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < nq) b = bp[q];
            v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b.x;
            v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
            v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z;
            v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
        }
    }

    if (a.epi_mode == EPI_GRU_ZR) {
        // channels [0,64): update gate z -> fp32; [64,128): reset gate r -> rh = h * r (split)
        act32(v, ACT_SIGMOID);
        if (n0 < 64) {
            float4 *zp = reinterpret_cast<float4 *>(a.z_buf + pix * 64 + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) zp[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
            float h[32];
            load_split32(a.h_prev + pix * 64 + (n0 - 64), a.h_plane, h);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= h[j];
            store_split32(a.out + pix * a.out_C + a.out_coff + (n0 - 64), a.out_plane, v);
        }
        return;
    }
    if (a.epi_mode == EPI_GRU_OUT) {
        // h' = h (1 - z) + tanh(.) z        (models/submodules.py:511-512)
        float h[32];
        load_split32(a.h_prev + pix * 64 + n0, a.h_plane, h);
        const float4 *zp = reinterpret_cast<const float4 *>(a.z_buf + pix * 64 + n0);
        float4 zq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) zq[q] = zp[q];
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
        store_split32(a.out + pix * a.out_C + a.out_coff + n0, a.out_plane, v);
        return;
    }
    // ---- standard epilogue: bias (+ residual before or after the activation)
    epilogue_std_math(a, v, n0, img, y, x, true);
    if (a.out && n0 + 32 <= a.cout) store_split32(a.out + pix * a.out_C + a.out_coff + n0, a.out_plane, v);
    if (a.out_f32 && a.out_f32_nchw) {
        // the autograd layout of the training operators: consecutive lanes = consecutive pixels of a tile row -> each of the
        // 32 per-channel stores of a warp is one or two contiguous segments
        float *op = a.out_f32 + (((size_t)img * a.out_f32_C + n0) * a.H + y) * a.W + x;
        const size_t cs = (size_t)a.H * a.W;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (n0 + j < a.cout) op[j * cs] = v[j];
    } else if (a.out_f32) {
        float *op = a.out_f32 + pix * a.out_f32_C + n0;
        if ((a.out_f32_C & 3) == 0) {                 // 16-byte stores (conv_offset_mask: 216 channels)
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (n0 + 4 * q + 4 <= a.cout)
                    reinterpret_cast<float4 *>(op)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < a.cout) op[j] = v[j];
        }
    }
}

} // namespace esr
