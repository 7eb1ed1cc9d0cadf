// cta_utils.cuh -- block-cooperative primitives shared by boxes.cu and discover.cu:
// exclusive scan, ordered compaction, bitonic sort on (score desc, position asc) and the
// greedy NMS sweep executed ON DEVICE by one CTA (the reference pulls a 64x64 bitmask to the host
// and sweeps there: csrc/cuda/nms.cu:99-123; torchvision does the same) -- no D2H, no sync.
#pragma once
#include "common.cuh"

namespace odw {

constexpr int kCtaThreads = 1024;

// Exclusive prefix sum of one int per thread across the CTA; *total gets the block sum.
// `s_warp` must hold 33 ints.  Contains two __syncthreads().
__device__ __forceinline__ int cta_exclusive_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    int w = lane < nw ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;            // exclusive per-warp offset
    if (lane == 31) s_warp[32] = winc;  // total
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[wid] + inc - v;
}

// Ordered compaction: out[k] = j for the k-th j in [0,n) with flag(j) != 0 (ascending j).
// Returns the count.  Each thread owns a contiguous chunk so order is preserved.
template <typename FlagFn, typename EmitFn>
__device__ __forceinline__ int cta_compact(int n, int* s_warp, FlagFn flag, EmitFn emit) {
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  int cnt = 0;
  for (int j = lo; j < hi; ++j) cnt += flag(j) ? 1 : 0;
  int total;
  int pos = cta_exclusive_scan(cnt, s_warp, &total);
  for (int j = lo; j < hi; ++j)
    if (flag(j)) emit(pos++, j);
  __syncthreads();
  return total;
}

__device__ __forceinline__ bool sort_before(float ka, int ia, float kb, int ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}

// In-place bitonic sort of (key, id) pairs in shared memory into (key desc, id asc) order.
// L is a power of two; padding entries must carry key = -inf, id = INT_MAX.
__device__ __forceinline__ void cta_bitonic_sort(float* s_key, int* s_id, int L) {
  for (int k = 2; k <= L; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < L; t += blockDim.x) {
        const int p = t ^ j;
        if (p > t) {
          const float ka = s_key[t], kb = s_key[p];
          const int ia = s_id[t], ib = s_id[p];
          const bool up = ((t & k) == 0);                 // ascending block in "before" order
          const bool swap = up ? sort_before(kb, ib, ka, ia) : sort_before(ka, ia, kb, ib);
          if (swap) { s_key[t] = kb; s_key[p] = ka; s_id[t] = ib; s_id[p] = ia; }
        }
      }
      __syncthreads();
    }
  }
}

// Greedy sweep over boxes already in sorted order.  s_box[n] sorted boxes, s_sup[n] zeroed.
// Calls keep_fn(k, sorted_pos) from thread 0 for the k-th kept box.  `one` selects the +1
// convention.  Returns the kept count (uniform across the CTA).
template <typename KeepFn>
__device__ __forceinline__ int cta_nms_sweep(const float4* s_box, unsigned char* s_sup, int n, float thr,
                                             float one, KeepFn keep_fn) {
  int nk = 0;
  for (int cur = 0; cur < n; ++cur) {
    if (s_sup[cur]) continue;                     // uniform: everyone reads the same byte
    if (threadIdx.x == 0) keep_fn(nk, cur);
    ++nk;
    const float4 bc = s_box[cur];
    for (int j = cur + 1 + threadIdx.x; j < n; j += blockDim.x)
      if (!s_sup[j] && odw_iou(bc, s_box[j], one) > thr) s_sup[j] = 1;
    __syncthreads();
  }
  return nk;
}

__device__ __forceinline__ int next_pow2(int n) {
  int L = 1;
  while (L < n) L <<= 1;
  return L;
}

}  // namespace odw
