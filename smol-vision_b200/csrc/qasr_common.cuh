// qasr_common.cuh - shared device helpers for libqasr_cuda.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define QASR_WARP 32
#define QASR_FULL 0xffffffffu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(QASR_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(QASR_FULL, v, o));
    return v;
}

// 128-bit streaming load that does not allocate in L1 (weights are read once per token).
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Exact bf16 -> f32 upcast (bits << 16), as the reference does (qwen_asr_kernels.c:232-236).
__device__ __forceinline__ float bf16lo_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi_f32(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

// dot of 8 bf16 weights (one uint4) with 8 f32 activations
__device__ __forceinline__ float dot8(const uint4 w, const float4 a, const float4 b, float acc) {
    acc = fmaf(bf16lo_f32(w.x), a.x, acc);
    acc = fmaf(bf16hi_f32(w.x), a.y, acc);
    acc = fmaf(bf16lo_f32(w.y), a.z, acc);
    acc = fmaf(bf16hi_f32(w.y), a.w, acc);
    acc = fmaf(bf16lo_f32(w.z), b.x, acc);
    acc = fmaf(bf16hi_f32(w.z), b.y, acc);
    acc = fmaf(bf16lo_f32(w.w), b.z, acc);
    acc = fmaf(bf16hi_f32(w.w), b.w, acc);
    return acc;
}

// f32 -> (hi, lo) bf16 split: hi = RN(x), lo = RN(x - hi).  hi + lo carries ~16 mantissa bits,
// so a bf16 tensor-core product against exact-bf16 weights reproduces the f32 reference.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// tanh-approximation GELU, reference qwen_asr_kernels.c:937-944
__device__ __forceinline__ float gelu_tanh(float v) {
    float inner = 0.7978845608028654f * (v + 0.044715f * v * v * v);
    return 0.5f * v * (1.0f + tanhf(inner));
}
// SiLU, reference qwen_asr_kernels.c:930-935
__device__ __forceinline__ float silu(float g) { return g / (1.0f + expf(-g)); }
// MUFU-based form for epilogues that sit on a latency path (a few ulp from silu(); the consumers round the result to bf16 hi / lo planes or add it into f32 sums)
__device__ __forceinline__ float silu_fast(float g) { return __fdividef(g, 1.0f + __expf(-g)); }
