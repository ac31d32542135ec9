// tc_conv.cuh -- interface of the tcgen05 implicit-GEMM convolution (see tc_conv.cu).
#pragma once
#include "common.cuh"
#include <cuda.h>   // CUtensorMap (types only; the driver entry point is resolved at run time)

namespace esr {
constexpr int TRACE_N = 512;                 // K-blocks / tiles recorded by the ESR_TC_TRACE measurement aid


enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_TANH = 3 };
enum EpiMode : int { EPI_STD = 0, EPI_GRU_ZR = 1, EPI_GRU_OUT = 2 };
enum ResMode : int { RES_NONE = 0, RES_PRE_ACT = 1, RES_POST_ACT = 2 };

// A "split" activation tensor: [2 planes][n_img][H][W][C] bf16, value = hi + lo.
struct SplitTensor {
    __nv_bfloat16 *base = nullptr;
    int n_img = 0, H = 0, W = 0, C = 0;
    size_t plane_override = 0;      // set on views into a larger tensor (distance between the hi and lo planes)
    size_t plane() const { return plane_override ? plane_override : (size_t)n_img * H * W * C; }
    size_t bytes() const { return plane() * 2 * sizeof(__nv_bfloat16); }
};

constexpr int TC_MAX_SRC = 3;

// Everything one launch needs.  Passed to the kernel by value (__grid_constant__).
struct ConvTCArgs {
    CUtensorMap amap[TC_MAX_SRC];   // 5-D maps over the source split tensors (C, W, H, img, plane), box (64, TW, TH, 1, 1)
    CUtensorMap bmap;               // 3-D map over packed weights (64, npad, 2*nkb), box (64, npad, 1)
    // The argument block must stay <= 1024 bytes: above that every launch of these kernels measured ~4 us slower (the store-side map
    // added as its own member made it 1088 bytes and cost 19 launches x 4 us per step).  The two maps below are never used together.
    union {
        CUtensorMap bmap_half;      // same tensor, box (64, npad/2, 1): the half a CTA multicasts in a 2-CTA cluster (pair / v3 kernels)
        CUtensorMap omap;           // out_tma (halo kernel): 5-D map over the OUTPUT (C, W, H, img, plane), box (32, TW, 32 / TW, 1, 1)
    };
    int out_tma;                    // 1: the epilogue stages each warp's 32 pixels x 32 channels in shared memory and stores them with TMA (split output);
                                    // 2: same for the fp32 NHWC output (omap: (C, W, H, img, 1) fp32, SWIZZLE_128B)
    int stg_bufs;                   // staging buffers per epilogue warp (2, or 1 when shared memory is short)
    int kernel_ver;                 // 1: tc_conv.cu (tap-shifted tiles), 3: tc_conv3.cu (halo reuse + weight multicast), 4: tc_conv_halo.cu
    int stack;                      // 1: [B_hi; B_lo] stacked along N -- two MMAs per K-step (A_hi x [B_hi;B_lo], A_lo x B_hi) instead of three
    int pair;                       // 1 (with persist): k_conv_tc_pair, two-CTA clusters issuing tcgen05.mma.cta_group::2
    int persist;                    // 1: k_conv_tc_persist (one CTA per SM walks the tiles, two TMEM accumulators)
    int a_stages, cluster;          // v3 only: halo ring depth, cluster size (1 or 2)
    const int *src_img[TC_MAX_SRC]; // output image -> source image (nullptr = identity)
    int chunk_end[TC_MAX_SRC];      // cumulative number of 64-channel chunks after source s
    int n_src, ntaps, nkb, npad, cout;
    int H, W, TW, TH, tiles_x, tiles_y, n_img;
    int stages;
    int trace_cta;                  // the CTA whose stamps are recorded (ESR_TC_TRACE_CTA, default 0)
    long long *trace;               // measurement aid (ESR_TC_TRACE): clock64 stamps of CTA 0's producer / MMA / epilogue threads
    int diag;                       // measurement aid (ESR_TC_DIAG): bit 0 = do not load the lo plane of B, bit 1 = not the lo plane of A (wrong results)
    // epilogue
    const float *bias;              // [npad]
    int act, act_from, res_mode, epi_mode;
    const __nv_bfloat16 *res; size_t res_plane; int res_C; const int *res_img;
    __nv_bfloat16 *out; size_t out_plane; int out_C, out_coff;   // split output (may be null)
    float *out_f32; int out_f32_C;                                // fp32 NHWC output (may be null)
    int out_f32_nchw;                                             // 1: out_f32 is NCHW [n_img][out_f32_C][H][W] instead
    // GRU extras
    const __nv_bfloat16 *h_prev; size_t h_plane;                  // [*,H,W,64] split
    float *z_buf;                                                 // [n_img,H,W,64] fp32 (ZR writes, OUT reads)
};

static_assert(sizeof(ConvTCArgs) <= 1024, "ConvTCArgs: keep the kernel argument block within 1 KB");

// Host-side description used to build ConvTCArgs.
struct ConvTCDesc {
    SplitTensor src[TC_MAX_SRC];
    const int *src_img[TC_MAX_SRC] = {nullptr, nullptr, nullptr};
    int n_src = 1;
    int ntaps = 9;                  // 9 (3x3, pad 1) or 1 (1x1)
    int cout = 64;
    const void *wpacked = nullptr;  // packed by pack_conv_weight: [2][nkb][npad][64] bf16
    const float *bias = nullptr;    // [npad] fp32, zero padded
    int n_img = 0;                  // number of output images
    int act = ACT_NONE, act_from = 0, res_mode = RES_NONE, epi_mode = EPI_STD;
    SplitTensor res; const int *res_img = nullptr;
    SplitTensor out; int out_coff = 0;
    float *out_f32 = nullptr; int out_f32_C = 0; int out_f32_nchw = 0;
    SplitTensor h_prev; float *z_buf = nullptr;
};

static inline int tc_npad(int cout) { return (cout + 15) / 16 * 16; }
static inline int tc_nkb(int cin_total, int ntaps) { return cin_total / 64 * ntaps; }
static inline size_t tc_packed_weight_bytes(int cout, int cin_total, int ntaps)
{
    return (size_t)2 * tc_nkb(cin_total, ntaps) * tc_npad(cout) * 64 * sizeof(__nv_bfloat16);
}

// Builds tensor maps + launch geometry.  H, W taken from src[0].
int conv_tc_prepare(const ConvTCDesc &d, ConvTCArgs *args);
int conv_tc_launch(const ConvTCArgs &args, cudaStream_t st);
// tensor-map builders (shared with gru_chain.cu)
int tc_make_amap(const SplitTensor &t, int box_w, int box_h, CUtensorMap *out);
int tc_make_bmap(const void *w, int npad, int nkb, int box_rows, CUtensorMap *out);
// tc_conv_halo.cu: persistent halo-reuse kernel for multi-wave 3x3 layers (kernel_ver 4)
bool conv_tc_halo_plan(int npad, int staging_bufs, int *a_stages, int *b_stages);
int tc_make_omap(const SplitTensor &t, int box_w, int box_h, CUtensorMap *out);
int tc_make_omap_f32(float *base, int n_img, int H, int W, int C, int box_w, int box_h, CUtensorMap *out);
int conv_tc_halo_launch(const ConvTCArgs &args, cudaStream_t st);
// tc_conv3.cu
bool conv_tc3_plan(int npad, int *a_stages, int *b_stages);
int conv_tc3_launch(const ConvTCArgs &args, cudaStream_t st);

// w: fp32 [cout, cin, k, k] (device) -> packed split bf16 [2][nkb][npad][64]; kb = chunk*ntaps + tap
int pack_conv_weight(const float *w, int cout, int cin, int ksz, void *dst, cudaStream_t st);
// concatenates two [cout_i, cin, k, k] weights along cout before packing (GRU update|reset gates)
int pack_conv_weight2(const float *w0, const float *w1, int cout_each, int cin, int ksz, void *dst, cudaStream_t st);

// NCHW fp32 <-> split NHWC
int split_from_nchw(const float *src, int n_img, int C, int H, int W, __nv_bfloat16 *dst, cudaStream_t st);
int split_to_nchw(const __nv_bfloat16 *src, int n_img, int C, int H, int W, float *dst, cudaStream_t st);
// same, with an explicit hi->lo plane distance (views into a larger split tensor)
int split_from_nchw_planes(const float *src, int n_img, int C, int H, int W, __nv_bfloat16 *dst, size_t plane, cudaStream_t st);
int split_to_nchw_planes(const __nv_bfloat16 *src, size_t plane, int n_img, int C, int H, int W, float *dst, cudaStream_t st);

} // namespace esr
