// direct_conv.cu -- CUDA-core 3x3 convolutions for the small-channel, full-resolution layers (HBM/latency bound,
// Cin <= 64 and Cout <= 64 with too few channels to fill a tensor-core tile):
//   head 2->8 (models/model.py:301,330), encoder 8->16->32->64 stride 2 (models/model.py:20-45),
//   attention maps C->1 sigmoid (models/model.py:195-199,262), decoder bilinear x2 + conv (models/submodules.py:254-299),
//   tail 8->2 (models/model.py:309,337).
// One block = 16x16 output pixels, all output channels; the input patch (+halo) is staged in shared memory as
// [ci][py][px] fp32, weights as [tap][ci][co] fp32 (broadcast reads).  fp32 math throughout.
#include "net.cuh"

namespace esr {

constexpr int DC_T = 16;   // output tile edge

template <int CIN, int COUT, int STRIDE, bool UPS, int INF, int OUTF>
__global__ void __launch_bounds__(256) k_conv_direct(const DirectArgs a)
{
    constexpr int PW = (DC_T - 1) * STRIDE + 3;
    extern __shared__ float dsm[];
    float *patch = dsm;                          // [CIN][PW][PW]
    float *wsm = dsm + CIN * PW * PW;            // [9][CIN][COUT]
    float *bsm = wsm + 9 * CIN * COUT;           // [COUT]

    const int img = blockIdx.z;
    const int oy0 = blockIdx.y * DC_T, ox0 = blockIdx.x * DC_T;
    const int tid = threadIdx.x;

    for (int i = tid; i < 9 * CIN * COUT; i += 256) wsm[i] = a.w[i];
    if (tid < COUT) bsm[tid] = a.bias[tid];

    // ---- stage the input patch.  Conv input coordinates (virtual, before any upsampling): [0,Hc) x [0,Wc)
    const int Hc = UPS ? 2 * a.Hin : a.Hin + a.pad_top + a.pad_bottom;
    const int Wc = UPS ? 2 * a.Win : a.Win + a.pad_left + a.pad_right;
    const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
    const int simg = a.in_img ? a.in_img[img] : img;
    if (INF == FMT_NCHW_F32) {
        for (int i = tid; i < CIN * PW * PW; i += 256) {
            const int px = i % PW, py = (i / PW) % PW, ci = i / (PW * PW);
            const int y = iy0 + py, x = ix0 + px;
            float v = 0.0f;
            if (y >= 0 && y < Hc && x >= 0 && x < Wc) {
                const int sy = y - a.pad_top, sx = x - a.pad_left;      // CropSize zero padding (model_util.py:148-152)
                if (sy >= 0 && sy < a.Hin && sx >= 0 && sx < a.Win)
                    v = a.in_f32[(((size_t)simg * CIN + ci) * a.Hin + sy) * a.Win + sx];
            }
            patch[i] = v;
        }
    } else {
        const __nv_bfloat16 *hi = a.in_split;
        const size_t plane = a.in_plane;
        for (int i = tid; i < CIN * PW * PW; i += 256) {
            const int ci = i % CIN, pp = i / CIN;
            const int px = pp % PW, py = pp / PW;
            const int y = iy0 + py, x = ix0 + px;
            float v = 0.0f;
            if (y >= 0 && y < Hc && x >= 0 && x < Wc) {
                if (!UPS) {
                    const size_t o = (((size_t)simg * a.Hin + y) * a.Win + x) * CIN + ci;
                    v = join_bf16(hi[o], hi[plane + o]);
                } else {
                    // F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) (submodules.py:290):
                    // src = max(0, (dst + 0.5) * 0.5 - 0.5); neighbours clamped to the image
                    const float fy = fmaxf(0.0f, ((float)y + 0.5f) * 0.5f - 0.5f);
                    const float fx = fmaxf(0.0f, ((float)x + 0.5f) * 0.5f - 0.5f);
                    const int y_0 = (int)fy, x_0 = (int)fx;
                    const int y_1 = min(y_0 + 1, a.Hin - 1), x_1 = min(x_0 + 1, a.Win - 1);
                    const float ly = fy - (float)y_0, lx = fx - (float)x_0;
                    const size_t b0 = ((size_t)simg * a.Hin + y_0) * a.Win, b1 = ((size_t)simg * a.Hin + y_1) * a.Win;
                    const size_t o00 = (b0 + x_0) * CIN + ci, o01 = (b0 + x_1) * CIN + ci;
                    const size_t o10 = (b1 + x_0) * CIN + ci, o11 = (b1 + x_1) * CIN + ci;
                    const float v00 = join_bf16(hi[o00], hi[plane + o00]), v01 = join_bf16(hi[o01], hi[plane + o01]);
                    const float v10 = join_bf16(hi[o10], hi[plane + o10]), v11 = join_bf16(hi[o11], hi[plane + o11]);
                    // same association as ATen's upsample_bilinear2d: w00*v00 + w01*v01 + w10*v10 + w11*v11
                    v = (1.0f - ly) * ((1.0f - lx) * v00 + lx * v01) + ly * ((1.0f - lx) * v10 + lx * v11);
                }
            }
            patch[(ci * PW + py) * PW + px] = v;
        }
    }
    __syncthreads();

    const int ty = tid / DC_T, tx = tid % DC_T;
    const int oy = oy0 + ty, ox = ox0 + tx;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = bsm[co];
#pragma unroll 1
    for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float xv = patch[(ci * PW + ty * STRIDE + ky) * PW + tx * STRIDE + kx];
                const float *wp = wsm + ((ky * 3 + kx) * CIN + ci) * COUT;
#pragma unroll
                for (int co = 0; co < COUT; ++co) acc[co] = fmaf(xv, wp[co], acc[co]);
            }
        }
    }
    if (oy >= a.Hout || ox >= a.Wout) return;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        float v = acc[co];
        if (a.act == ACT_RELU) v = fmaxf(v, 0.0f);
        else if (a.act == ACT_SIGMOID) v = 1.0f / (1.0f + expf(-v));
        acc[co] = v;
    }
    if (OUTF == FMT_SPLIT) {
        const size_t o = (((size_t)img * a.Hout + oy) * a.Wout + ox) * COUT;
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            __nv_bfloat16 h, l;
            split_bf16(acc[co], h, l);
            a.out_split[o + co] = h;
            a.out_split[a.out_plane + o + co] = l;
        }
    } else if (OUTF == FMT_NHWC_F32) {
        const size_t o = (((size_t)img * a.Hout + oy) * a.Wout + ox) * COUT;
#pragma unroll
        for (int co = 0; co < COUT; ++co) a.out_f32[o + co] = acc[co];
    } else {
        // NCHW fp32 with the CropSize crop (model_util.py:154-164): only rows/cols inside the crop window are stored
        const int cy = oy - a.crop_top, cx = ox - a.crop_left;
        if (cy >= 0 && cy < a.out_H && cx >= 0 && cx < a.out_W) {
#pragma unroll
            for (int co = 0; co < COUT; ++co) a.out_f32[(((size_t)img * COUT + co) * a.out_H + cy) * a.out_W + cx] = acc[co];
        }
    }
}

template <int CIN, int COUT, int STRIDE, bool UPS, int INF, int OUTF>
static int launch_direct(const DirectArgs &a, cudaStream_t st)
{
    constexpr int PW = (DC_T - 1) * STRIDE + 3;
    constexpr size_t smem = sizeof(float) * (size_t)(CIN * PW * PW + 9 * CIN * COUT + COUT);
    static bool attr_set = false;
    if (!attr_set) {
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_conv_direct<CIN, COUT, STRIDE, UPS, INF, OUTF>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid((a.Wout + DC_T - 1) / DC_T, (a.Hout + DC_T - 1) / DC_T, a.n_img);
    k_conv_direct<CIN, COUT, STRIDE, UPS, INF, OUTF><<<grid, 256, smem, st>>>(a);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

int conv_direct(DirectKind kind, const DirectArgs &a, cudaStream_t st)
{
    switch (kind) {
    case DK_HEAD:    return launch_direct<2, 8, 1, false, FMT_NCHW_F32, FMT_SPLIT>(a, st);
    case DK_ENC0:    return launch_direct<8, 16, 2, false, FMT_SPLIT, FMT_SPLIT>(a, st);
    case DK_ENC1:    return launch_direct<16, 32, 2, false, FMT_SPLIT, FMT_SPLIT>(a, st);
    case DK_ENC2:    return launch_direct<32, 64, 2, false, FMT_SPLIT, FMT_SPLIT>(a, st);
    case DK_ATT32:   return launch_direct<32, 1, 1, false, FMT_SPLIT, FMT_NHWC_F32>(a, st);
    case DK_ATT16:   return launch_direct<16, 1, 1, false, FMT_SPLIT, FMT_NHWC_F32>(a, st);
    case DK_RECON0:  return launch_direct<64, 32, 1, true, FMT_SPLIT, FMT_SPLIT>(a, st);
    case DK_RECON1:  return launch_direct<32, 16, 1, true, FMT_SPLIT, FMT_SPLIT>(a, st);
    case DK_RECON2:  return launch_direct<16, 8, 1, true, FMT_SPLIT, FMT_SPLIT>(a, st);
    case DK_TAIL:    return launch_direct<8, 2, 1, false, FMT_SPLIT, FMT_NCHW_F32>(a, st);
    }
    set_error("conv_direct: unknown kind %d", (int)kind);
    return ESR_EINVAL;
}

// fp32 [Cout, Cin, 3, 3] -> [tap][ci][co]
__global__ void k_pack_direct_weight(const float *__restrict__ w, int cout, int cin, float *__restrict__ dst)
{
    const int total = 9 * cin * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % cout, ci = (i / cout) % cin, tap = i / (cout * cin);
        dst[i] = w[((size_t)co * cin + ci) * 9 + tap];
    }
}
int pack_direct_weight(const float *w, int cout, int cin, float *dst, cudaStream_t st)
{
    k_pack_direct_weight<<<(9 * cin * cout + 255) / 256, 256, 0, st>>>(w, cout, cin, dst);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

} // namespace esr
