// dcn.cu -- modulated deformable sampling (DCNv2) producing the A operand of the 3x3xC contraction.
// Semantics restated from models/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:25-54 (bilinear with zero outside the image)
// and :125-195 (per deformable group g, tap k = i*3+j: offsets at channels g*18+2k (row) / +1 (col), mask at g*9+k;
// sample at (y-1+i+off_h, x-1+j+off_w) only if it lies in (-1, H) x (-1, W); columns = value * mask).
// Stride 1, pad 1, dilation 1, 3x3, 64 channels, 8 deformable groups of 8 channels (models/model.py:173).
//
// Layout choices for the B200: features are split-bf16 NHWC so the 8 channels of one group at one corner are one
// 16-byte load per plane; one thread = one (pixel, tap, group); its 8 outputs land at column tap*64 + g*8, i.e. eight
// consecutive threads (g = 0..7) write one contiguous 128-byte row segment.  The columns tensor is then consumed by the
// tcgen05 GEMM (tc_conv.cu, 1x1 mode over 9 K-chunks) with the DCN weight packed tap-major, which equals the
// reference's W[Co, Ci*9] . columns contraction (dcn_v2_cuda.cu:90-92).
#include "net.cuh"

namespace esr {

__device__ __forceinline__ void ld8(const __nv_bfloat16 *hi, size_t plane, float (&o)[8])
{
    const uint4 h = *reinterpret_cast<const uint4 *>(hi);
    const uint4 l = *reinterpret_cast<const uint4 *>(hi + plane);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        o[2 * e] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
        o[2 * e + 1] = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
    }
}

__global__ void __launch_bounds__(256)
k_dcn_columns(const __nv_bfloat16 *__restrict__ feat, size_t f_plane, const int *__restrict__ feat_img,
              const float *__restrict__ om, int n_img, int H, int W, __nv_bfloat16 *__restrict__ cols, size_t c_plane)
{
    const size_t total = (size_t)n_img * H * W * 72;          // 9 taps x 8 groups
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % 8), k = (int)((i / 8) % 9);
        const size_t p = i / 72;
        const int x = (int)(p % W), y = (int)((p / W) % H), img = (int)(p / ((size_t)W * H));
        const float *o = om + p * 216;
        const float off_h = o[g * 18 + 2 * k], off_w = o[g * 18 + 2 * k + 1], m = o[144 + g * 9 + k];
        const float h_im = (float)(y - 1 + k / 3) + off_h;
        const float w_im = (float)(x - 1 + k % 3) + off_w;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.0f;
        if (h_im > -1.0f && w_im > -1.0f && h_im < (float)H && w_im < (float)W) {
            const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
            const int h_high = h_low + 1, w_high = w_low + 1;
            const float lh = h_im - (float)h_low, lw = w_im - (float)w_low;
            const float hh = 1.0f - lh, hw = 1.0f - lw;
            const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
            const size_t base = (size_t)(feat_img ? feat_img[img] : img) * H * W;
            float c1[8], c2[8], c3[8], c4[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) c1[e] = c2[e] = c3[e] = c4[e] = 0.0f;
            if (h_low >= 0 && w_low >= 0) ld8(feat + ((base + (size_t)h_low * W + w_low) * 64 + g * 8), f_plane, c1);
            if (h_low >= 0 && w_high <= W - 1) ld8(feat + ((base + (size_t)h_low * W + w_high) * 64 + g * 8), f_plane, c2);
            if (h_high <= H - 1 && w_low >= 0) ld8(feat + ((base + (size_t)h_high * W + w_low) * 64 + g * 8), f_plane, c3);
            if (h_high <= H - 1 && w_high <= W - 1) ld8(feat + ((base + (size_t)h_high * W + w_high) * 64 + g * 8), f_plane, c4);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (w1 * c1[e] + w2 * c2[e] + w3 * c3[e] + w4 * c4[e]) * m;
        }
        uint32_t hw_[4], lw_[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(v[2 * e], h0, l0);
            split_bf16(v[2 * e + 1], h1, l1);
            hw_[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lw_[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        __nv_bfloat16 *dst = cols + p * 576 + k * 64 + g * 8;
        *reinterpret_cast<uint4 *>(dst) = make_uint4(hw_[0], hw_[1], hw_[2], hw_[3]);
        *reinterpret_cast<uint4 *>(dst + c_plane) = make_uint4(lw_[0], lw_[1], lw_[2], lw_[3]);
    }
}

int dcn_columns(const SplitTensor &feat, const int *feat_img, const float *om, int n_img, const SplitTensor &cols, cudaStream_t st)
{
    ESR_REQUIRE(feat.C == 64 && cols.C == 576 && cols.H == feat.H && cols.W == feat.W, "dcn_columns: bad shapes");
    const size_t total = (size_t)n_img * feat.H * feat.W * 72;
    k_dcn_columns<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(feat.base, feat.plane(), feat_img, om, n_img, feat.H,
                                                                             feat.W, cols.base, cols.plane());
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

} // namespace esr
