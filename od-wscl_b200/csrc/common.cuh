// common.cuh -- shared helpers for the sm_100a kernels of libodwscl_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "odwscl.h"

#define ODW_API extern "C" __attribute__((visibility("default")))
#define ODW_LAUNCH_CHECK()                                \
  do {                                                    \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) return (int)e__;              \
  } while (0)
#define ODW_CUDA(call)                                    \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return (int)e__;              \
  } while (0)

static inline int odw_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t odw_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
#define ODW_NUM_SMS 148   // B200: 2 dies x 74 SMs
// SMs the persistent kernels (conv3x3 split-K pairs, fc GEMM pairs) leave free -- set with odwscl_set_sm_margin() when a
// concurrent NCCL all-reduce must find SMs without waiting for a persistent kernel to end (capi.cu)
int odw_sm_margin();

__device__ __forceinline__ float odw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float odw_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// IoU of two xyxy boxes; `one` = 1.f for the legacy +1 pixel convention
// (structures/boxlist_ops.py:151-159), 0.f for torchvision's.  Operation order follows the
// reference: inter / ((area_a + area_b) - inter); no FMA contraction (explicit intrinsics).
__device__ __forceinline__ float odw_iou(const float4 a, const float4 b, const float one) {
  const float ltx = fmaxf(a.x, b.x), lty = fmaxf(a.y, b.y);
  const float rbx = fminf(a.z, b.z), rby = fminf(a.w, b.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(rbx, ltx), one), 0.f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(rby, lty), one), 0.f);
  const float inter = __fmul_rn(w, h);
  const float aa = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), one), __fadd_rn(__fsub_rn(a.w, a.y), one));
  const float ab = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), one), __fadd_rn(__fsub_rn(b.w, b.y), one));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
}
