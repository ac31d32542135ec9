// resize.cu -- the two plane resizes of the dataset's per-frame tensor factory:
//   F.interpolate(x, size=(Hout, Wout), mode='bicubic', align_corners=False)   dataloader/h5dataset.py:341-342, infer_ours_cnt.py:76-78
//   F.interpolate(x, size=(Hout, Wout), mode='nearest')                        dataloader/h5dataset.py:343-344
// following ATen's CPU kernels (the reference's dependency that carries the arithmetic): source index
// scale * (dst + 0.5) - 0.5 with scale = in / out in fp32, Keys cubic coefficients with A = -0.75, taps clamped to the image,
// x taps accumulated left to right per row, rows top to bottom; legacy nearest: min(floor(dst * scale), in - 1).
// One thread per output element; the kernels are HBM-bound (4 B written per element, the 16 taps come from L1/L2).
#include "common.cuh"

namespace esr {

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4])
{
    const float A = -0.75f;
    c[0] = cubic2(t + 1.0f, A);
    c[1] = cubic1(t, A);
    c[2] = cubic1(1.0f - t, A);
    c[3] = cubic2(2.0f - t, A);
}

__global__ void __launch_bounds__(256) k_resize_bicubic(const float *__restrict__ x, int Hin, int Win, int Hout, int Wout,
                                                        float sh, float sw, float *__restrict__ out)
{
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
    if (ox >= Wout) return;
    const float *src = x + (size_t)blockIdx.z * Hin * Win;
    const float ry = sh * ((float)oy + 0.5f) - 0.5f, rx = sw * ((float)ox + 0.5f) - 0.5f;
    const float fy = floorf(ry), fx = floorf(rx);
    float cy[4], cx[4];
    cubic_coeffs(fminf(fmaxf(ry - fy, 0.0f), 1.0f), cy);
    cubic_coeffs(fminf(fmaxf(rx - fx, 0.0f), 1.0f), cx);
    const int iy = (int)fy, ix = (int)fx;
    int xs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = min(max(ix - 1 + j, 0), Win - 1);
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float *row = src + (size_t)min(max(iy - 1 + i, 0), Hin - 1) * Win;
        float r = __fmul_rn(__ldg(row + xs[0]), cx[0]);                     // no FMA contraction: ATen's scalar order
        r = __fadd_rn(r, __fmul_rn(__ldg(row + xs[1]), cx[1]));
        r = __fadd_rn(r, __fmul_rn(__ldg(row + xs[2]), cx[2]));
        r = __fadd_rn(r, __fmul_rn(__ldg(row + xs[3]), cx[3]));
        acc = i == 0 ? __fmul_rn(r, cy[0]) : __fadd_rn(acc, __fmul_rn(r, cy[i]));
    }
    out[((size_t)blockIdx.z * Hout + oy) * Wout + ox] = acc;
}

__global__ void __launch_bounds__(256) k_resize_nearest(const float *__restrict__ x, int Hin, int Win, int Hout, int Wout,
                                                        float sh, float sw, float *__restrict__ out)
{
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
    if (ox >= Wout) return;
    const int iy = Hout == Hin ? oy : (Hout == 2 * Hin ? oy >> 1 : min((int)floorf((float)oy * sh), Hin - 1));
    const int ix = Wout == Win ? ox : (Wout == 2 * Win ? ox >> 1 : min((int)floorf((float)ox * sw), Win - 1));
    out[((size_t)blockIdx.z * Hout + oy) * Wout + ox] = __ldg(x + ((size_t)blockIdx.z * Hin + iy) * Win + ix);
}

} // namespace esr

using namespace esr;

extern "C" int esr_resize_planes(const float *x, int planes, int Hin, int Win, int Hout, int Wout, int mode, float *out,
                                 esr_stream_t stream)
{
    ESR_REQUIRE(x && out, "esr_resize_planes: null pointer");
    ESR_REQUIRE(planes > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "esr_resize_planes: bad dims");
    ESR_REQUIRE(mode == 0 || mode == 1, "esr_resize_planes: mode must be 0 (nearest) or 1 (bicubic)");
    ESR_REQUIRE(planes <= 65535 && Hout <= 65535, "esr_resize_planes: at most 65535 planes / output rows per call");
    const float sh = (float)Hin / (float)Hout, sw = (float)Win / (float)Wout;     // area_pixel_compute_scale, align_corners=False
    dim3 grid((unsigned)((Wout + 255) / 256), (unsigned)Hout, (unsigned)planes);
    if (mode == 1) k_resize_bicubic<<<grid, 256, 0, (cudaStream_t)stream>>>(x, Hin, Win, Hout, Wout, sh, sw, out);
    else k_resize_nearest<<<grid, 256, 0, (cudaStream_t)stream>>>(x, Hin, Win, Hout, Wout, sh, sw, out);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}
