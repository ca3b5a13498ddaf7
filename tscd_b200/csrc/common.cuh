// Shared device helpers for the TSCD aggregation-stage kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/tscd_b200.h"

extern "C" void tscd_set_last_cuda_error(int code, const char* where);
#define TSCD_CUDA_CHECK_LAUNCH()                                   \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) {                                  \
            tscd_set_last_cuda_error((int)e__, __FILE__);          \
            return TSCD_ERR_CUDA;                                  \
        }                                                          \
    } while (0)

namespace tscd {

constexpr int kWarp = 32;

// ---- order-preserving float <-> uint32 (larger float -> larger uint) --------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- element loads (boundary tensors may be fp32 / fp16 / bf16) --------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
}
__device__ __forceinline__ float ldf_reg(float v) { return v; }
__device__ __forceinline__ float ldf_reg(__half v) { return __half2float(v); }
__device__ __forceinline__ float ldf_reg(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T cvt_from_float(float v);
template <> __device__ __forceinline__ float cvt_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt_from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_from_float<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}

// sigmoid matching ATen's  1 / (1 + exp(-x))  expression (used at seam S1 only)
__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// ---- warp / block reductions ------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sumf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one int per thread; returns exclusive prefix, writes total to *total.
// scratch: >= 33 ints of shared memory.  All threads must call.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ int block_excl_scan(int v, int* scratch, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = warp_incl_scan(v, lane);
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? scratch[lane] : 0;
        int winc = warp_incl_scan(w, lane);
        scratch[lane] = winc - w;
        if (lane == 31) scratch[32] = winc;
    }
    __syncthreads();
    int res = inc - v + scratch[wid];
    *total = scratch[32];
    __syncthreads();
    return res;
}

// ---- anchor geometry ----------------------------------------------------------------------------------
struct AnchorPos {
    int level;
    int local;  // index within the level
    float gx, gy, stride;
};
__device__ __forceinline__ AnchorPos anchor_pos(const tscd_anchors& an, int a) {
    AnchorPos p;
    int l = 0;
#pragma unroll
    for (int i = 1; i < TSCD_MAX_LEVELS; ++i)
        if (i < an.num_levels && a >= an.level_start[i]) l = i;
    p.level = l;
    p.local = a - an.level_start[l];
    int w = an.level_w[l];
    int y = p.local / w;
    p.gx = (float)(p.local - y * w);
    p.gy = (float)y;
    p.stride = (float)an.level_stride[l];
    return p;
}

template <typename T>
__device__ __forceinline__ const T* view_ptr(const tscd_view& v, int level, int frame, int local) {
    return reinterpret_cast<const T*>(v.ptr[level]) + (int64_t)frame * v.frame_stride[level] +
           (int64_t)local * v.anchor_stride[level];
}

// Box of one anchor as the reference computes it: decode_outputs (tscd_head.py:768-769) then
// cxcywh->xyxy (tscd_head.py:1561-1566).  Explicit round-to-nearest ops: no FMA contraction.
template <typename T>
__device__ __forceinline__ float4 anchor_box(const tscd_view& reg, const AnchorPos& p, int frame, bool decode) {
    const T* r = view_ptr<T>(reg, p.level, frame, p.local);
    const int64_t cs = reg.chan_stride[p.level];
    float cx = ldf(r), cy = ldf(r + cs), w = ldf(r + 2 * cs), h = ldf(r + 3 * cs);
    if (decode) {
        cx = __fmul_rn(__fadd_rn(cx, p.gx), p.stride);
        cy = __fmul_rn(__fadd_rn(cy, p.gy), p.stride);
        w = __fmul_rn(expf(w), p.stride);
        h = __fmul_rn(expf(h), p.stride);
    }
    float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);  // w/2 is exact either way
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

}  // namespace tscd
