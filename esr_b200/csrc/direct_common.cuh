// direct_common.cuh -- device helpers shared by the small-channel convolution kernels (direct_conv.cu, mma_conv.cu).
#pragma once
#include "net.cuh"

namespace esr {

__device__ __forceinline__ void dc_unpack8(const uint4 h, const uint4 l, float (&o)[8])
{
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        o[2 * e] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
        o[2 * e + 1] = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
    }
}
__device__ __forceinline__ void dc_ld8(const __nv_bfloat16 *p, size_t plane, float (&o)[8])
{
    dc_unpack8(*reinterpret_cast<const uint4 *>(p), *reinterpret_cast<const uint4 *>(p + plane), o);
}
// NV values (2, 4 or 8) -> split-bf16, vector stores
template <int NV>
__device__ __forceinline__ void dc_store(__nv_bfloat16 *p, size_t plane, const float *x)
{
    uint32_t hw[NV / 2], lw[NV / 2];
#pragma unroll
    for (int e = 0; e < NV / 2; ++e) {
        split_pack2(x[2 * e], x[2 * e + 1], hw[e], lw[e]);
    }
    if constexpr (NV == 8) {
        *reinterpret_cast<uint4 *>(p) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4 *>(p + plane) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    } else if constexpr (NV == 4) {
        *reinterpret_cast<uint2 *>(p) = make_uint2(hw[0], hw[1]);
        *reinterpret_cast<uint2 *>(p + plane) = make_uint2(lw[0], lw[1]);
    } else {
        *reinterpret_cast<uint32_t *>(p) = hw[0];
        *reinterpret_cast<uint32_t *>(p + plane) = lw[0];
    }
}

__device__ __forceinline__ float dc_act(float v, int act)
{
    if (act == ACT_RELU) return fmaxf(v, 0.0f);
    if (act == ACT_SIGMOID) return fast_sigmoid(v);
    return v;
}

} // namespace esr
