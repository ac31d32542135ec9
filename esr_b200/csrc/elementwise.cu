// elementwise.cu -- the fused element-wise / reduction glue of DeepRecurrNet on split-bf16 NHWC tensors.
// All HBM-bound: 16-byte vector accesses, 8 channels per thread.
#include "net.cuh"

namespace esr {

__device__ __forceinline__ void load8(const __nv_bfloat16 *hi, size_t plane, float (&o)[8])
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
__device__ __forceinline__ void store8(__nv_bfloat16 *hi, size_t plane, const float (&x)[8])
{
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        split_pack2(x[2 * e], x[2 * e + 1], hw[e], lw[e]);
    }
    *reinterpret_cast<uint4 *>(hi) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4 *>(hi + plane) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ltc_cat(const __nv_bfloat16 *__restrict__ f, size_t f_plane, const float *__restrict__ maps, const int *__restrict__ idx,
          int n_img, int HW, __nv_bfloat16 *__restrict__ out, size_t out_plane)
{
    const size_t total = (size_t)n_img * HW * 24;            // 24 groups of 8 channels = 192
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % 24);
        const size_t p = i / 24;
        const int pix = (int)(p % HW), img = (int)(p / HW);
        const int part = g / 8, c0 = (g % 8) * 8;            // part 0: f0*m0, 1: f1, 2: f2*m1
        const int *ix = idx + img * 5;
        const size_t src = ((size_t)ix[part] * HW + pix) * 64 + c0;
        float v[8];
        load8(f + src, f_plane, v);
        if (part != 1) {
            const float m = maps[(size_t)ix[part == 0 ? 3 : 4] * HW + pix];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] *= m;
        }
        store8(out + p * 192 + part * 64 + c0, out_plane, v);
    }
}
int ltc_cat(const SplitTensor &f, const float *maps, const int *idx, int n_img, const SplitTensor &out, cudaStream_t st)
{
    const int HW = f.H * f.W;
    const size_t total = (size_t)n_img * HW * 24;
    k_ltc_cat<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(f.base, f.plane(), maps, idx, n_img, HW, out.base,
                                                                         out.plane());
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
// per-image channel max: grid (slices, n_img); partial maxima merged with integer atomicMax on the ordered
// bit pattern (monotone map of fp32), out pre-filled with -inf by the launcher.
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256)
k_chan_max(const __nv_bfloat16 *__restrict__ t, size_t plane, int HW, int *__restrict__ out_ord)
{
    const int img = blockIdx.y;
    const int c8 = threadIdx.x & 7;                      // 8 groups of 8 channels
    const int lane_pix = threadIdx.x >> 3;               // 32 pixels per pass
    float mx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) mx[e] = -INFINITY;
    for (int pix = blockIdx.x * 32 + lane_pix; pix < HW; pix += gridDim.x * 32) {
        float v[8];
        load8(t + ((size_t)img * HW + pix) * 64 + c8 * 8, plane, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) mx[e] = fmaxf(mx[e], v[e]);
    }
    __shared__ float sm[32][64];
#pragma unroll
    for (int e = 0; e < 8; ++e) sm[lane_pix][c8 * 8 + e] = mx[e];
    __syncthreads();
    if (threadIdx.x < 64) {
        float m = -INFINITY;
        for (int r = 0; r < 32; ++r) m = fmaxf(m, sm[r][threadIdx.x]);
        atomicMax(out_ord + img * 64 + threadIdx.x, f2ord(m));
    }
}
static int f2ord_host_neg_inf() { return (int)(0xFF800000u ^ 0x7FFFFFFFu); }
__global__ void k_fill_int(int *p, int n, int v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
int chan_max(const SplitTensor &t, int n_img, float *out, cudaStream_t st)
{
    const int HW = t.H * t.W;
    k_fill_int<<<(n_img * 64 + 255) / 256, 256, 0, st>>>((int *)out, n_img * 64, f2ord_host_neg_inf());
    ESR_LAUNCH_CHECK();
    int slices = (HW + 32 * 8 - 1) / (32 * 8);
    if (slices > 64) slices = 64;
    k_chan_max<<<dim3(slices, n_img), 256, 0, st>>>(t.base, t.plane(), HW, (int *)out);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

// channel attention: ck = sigmoid(W1 relu(W0 mx + b0) + b1); mx arrives as ordered ints from chan_max
__global__ void __launch_bounds__(128)
k_attn_mlp(const int *__restrict__ mx_ord, const float *__restrict__ w0, const float *__restrict__ b0,
           const float *__restrict__ w1, const float *__restrict__ b1, float *__restrict__ ck)
{
    __shared__ float m[64], hdn[32];
    const int img = blockIdx.x, t = threadIdx.x;
    if (t < 64) m[t] = ord2f(mx_ord[img * 64 + t]);
    __syncthreads();
    if (t < 32) {
        float s = b0[t];
        for (int k = 0; k < 64; ++k) s = fmaf(w0[t * 64 + k], m[k], s);
        hdn[t] = fmaxf(s, 0.0f);
    }
    __syncthreads();
    float s = b1[t];
    for (int k = 0; k < 32; ++k) s = fmaf(w1[t * 32 + k], hdn[k], s);
    ck[img * 128 + t] = 1.0f / (1.0f + expf(-s));
}
int attn_mlp(const float *mx, int n_img, const float *w0, const float *b0, const float *w1, const float *b1, float *ck,
             cudaStream_t st)
{
    k_attn_mlp<<<n_img, 128, 0, st>>>((const int *)mx, w0, b0, w1, b1, ck);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_attn_apply(const __nv_bfloat16 *__restrict__ al, size_t al_plane, const __nv_bfloat16 *__restrict__ mid, size_t mid_plane,
             const int *__restrict__ mid_img, const float *__restrict__ sk, const float *__restrict__ ck, int n_img, int HW,
             __nv_bfloat16 *__restrict__ out, size_t out_plane)
{
    const size_t total = (size_t)n_img * HW * 16;            // 16 groups of 8 channels = 128
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % 16);
        const size_t p = i / 16;
        const int pix = (int)(p % HW), img = (int)(p / HW);
        const int part = g / 8, c0 = (g % 8) * 8;
        float v[8];
        if (part == 0) load8(al + p * 64 + c0, al_plane, v);
        else load8(mid + ((size_t)(mid_img ? mid_img[img] : img) * HW + pix) * 64 + c0, mid_plane, v);
        const float s = sk[p * 2 + part];
        const float *c = ck + img * 128 + part * 64 + c0;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (v[e] * s) * c[e];       // (feat * spatial) * channel, as model.py:224-227
        store8(out + p * 128 + part * 64 + c0, out_plane, v);
    }
}
int attn_apply(const SplitTensor &aligned, const SplitTensor &mid_src, const int *mid_img, const float *sk, const float *ck,
               int n_img, const SplitTensor &out, cudaStream_t st)
{
    const int HW = aligned.H * aligned.W;
    const size_t total = (size_t)n_img * HW * 16;
    k_attn_apply<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(aligned.base, aligned.plane(), mid_src.base,
                                                                            mid_src.plane(), mid_img, sk, ck, n_img, HW,
                                                                            out.base, out.plane());
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_scale_aggregate(const __nv_bfloat16 *__restrict__ x, size_t x_plane, const __nv_bfloat16 *__restrict__ feats, size_t f_plane,
                  const float *__restrict__ att, const int *__restrict__ fidx, int B, int N, int HW, int C,
                  __nv_bfloat16 *__restrict__ out, size_t out_plane)
{
    const int G = C / 8;
    const size_t total = (size_t)B * HW * G;
    const float inv = 1.0f / (float)N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const size_t p = i / G;
        const int pix = (int)(p % HW), b = (int)(p / HW);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
        for (int n = 0; n < N; ++n) {
            const size_t fp = (size_t)(fidx ? fidx[b * N + n] : b * N + n) * HW + pix;
            float v[8];
            load8(feats + fp * C + g * 8, f_plane, v);
            const float a = att[fp];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v[e] * a;         // sum over frames in order, then / N (mean)
        }
        float xv[8];
        load8(x + p * C + g * 8, x_plane, xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) xv[e] += acc[e] * inv;
        store8(out + p * C + g * 8, out_plane, xv);
    }
}
int scale_aggregate(const SplitTensor &x, const SplitTensor &feats, const float *att, const int *fidx, int B, int N,
                    const SplitTensor &out, cudaStream_t st)
{
    const int HW = x.H * x.W, C = x.C;
    const size_t total = (size_t)B * HW * (C / 8);
    k_scale_aggregate<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(x.base, x.plane(), feats.base, feats.plane(),
                                                                                 att, fidx, B, N, HW, C, out.base, out.plane());
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
// F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) (models/submodules.py:290) on a split tensor:
// src index = max(0, (dst + 0.5) / 2 - 0.5), neighbours clamped; same association as ATen's upsample_bilinear2d.
__global__ void __launch_bounds__(256)
k_upsample2x(const __nv_bfloat16 *__restrict__ src, size_t s_plane, int n_img, int H, int W, int C,
             __nv_bfloat16 *__restrict__ dst, size_t d_plane)
{
    const int G = C / 8, H2 = 2 * H, W2 = 2 * W;
    const size_t total = (size_t)n_img * H2 * W2 * G;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const size_t p = i / G;
        const int x = (int)(p % W2), y = (int)((p / W2) % H2), img = (int)(p / ((size_t)W2 * H2));
        const float fy = fmaxf(0.0f, ((float)y + 0.5f) * 0.5f - 0.5f), fx = fmaxf(0.0f, ((float)x + 0.5f) * 0.5f - 0.5f);
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const size_t b0 = ((size_t)img * H + y0) * W, b1 = ((size_t)img * H + y1) * W;
        float v00[8], v01[8], v10[8], v11[8], o[8];
        load8(src + (b0 + x0) * C + g * 8, s_plane, v00);
        load8(src + (b0 + x1) * C + g * 8, s_plane, v01);
        load8(src + (b1 + x0) * C + g * 8, s_plane, v10);
        load8(src + (b1 + x1) * C + g * 8, s_plane, v11);
#pragma unroll
        for (int e = 0; e < 8; ++e)
            o[e] = (1.0f - ly) * ((1.0f - lx) * v00[e] + lx * v01[e]) + ly * ((1.0f - lx) * v10[e] + lx * v11[e]);
        store8(dst + p * C + g * 8, d_plane, o);
    }
}
int upsample2x(const SplitTensor &src, int n_img, const SplitTensor &dst, cudaStream_t st)
{
    ESR_REQUIRE(dst.H == 2 * src.H && dst.W == 2 * src.W && dst.C == src.C && src.C % 8 == 0, "upsample2x: bad shapes");
    const size_t total = (size_t)n_img * dst.H * dst.W * (src.C / 8);
    k_upsample2x<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(src.base, src.plane(), n_img, src.H, src.W, src.C,
                                                                            dst.base, dst.plane());
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_copy_split(const __nv_bfloat16 *__restrict__ src, size_t s_plane, const int *__restrict__ src_img, int n_img, size_t per_img8,
             __nv_bfloat16 *__restrict__ dst, size_t d_plane)
{
    const size_t total = (size_t)n_img * per_img8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t img = i / per_img8, r = i % per_img8;
        const size_t s = ((size_t)(src_img ? src_img[img] : (int)img) * per_img8 + r) * 8;
        const size_t d = i * 8;
        *reinterpret_cast<uint4 *>(dst + d) = *reinterpret_cast<const uint4 *>(src + s);
        *reinterpret_cast<uint4 *>(dst + d_plane + d) = *reinterpret_cast<const uint4 *>(src + s_plane + s);
    }
}
int copy_split(const SplitTensor &src, const int *src_img, int n_img, const SplitTensor &dst, cudaStream_t st)
{
    const size_t per_img8 = (size_t)src.H * src.W * src.C / 8;
    const size_t total = (size_t)n_img * per_img8;
    k_copy_split<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(src.base, src.plane(), src_img, n_img, per_img8,
                                                                            dst.base, dst.plane());
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

} // namespace esr
