// boxes.cu -- pairwise IoU and single-CTA on-device NMS.
//   odwscl_box_iou_f32      : boxlist_iou (structures/boxlist_ops.py:127-160) / torchvision IoU
//   odwscl_nms_f32          : torchvision.ops.nms semantics (structures/boxlist_ops.py:57)
//   odwscl_nms_legacy_f32   : `_C.nms` (csrc/nms.h:10-28 -> csrc/cuda/nms.cu:23-130)
// Both NMS variants sort, sweep and compact inside ONE CTA with everything in shared memory
// (n <= 8192): no bitmask round trip to the host, no stream synchronisation.
#include "cta_utils.cuh"

namespace {

constexpr int kNmsMax = 8192;

__global__ void box_iou_kernel(const float4* __restrict__ a, int na, const float4* __restrict__ b, int nb,
                               float one, float* __restrict__ out) {
  const long long total = (long long)na * nb;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / nb), j = (int)(t % nb);
    out[t] = odw_iou(__ldg(a + i), __ldg(b + j), one);
  }
}

// dynamic smem: key[L] f32 | id[L] i32 | box[L] float4 | sup[L] u8 | flag[n] u8 (legacy) | scan[33]
template <bool kLegacy>
__global__ void __launch_bounds__(odw::kCtaThreads, 1)
nms_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, int n, int L, float thr,
           int64_t* __restrict__ keep, int32_t* __restrict__ n_keep) {
  extern __shared__ __align__(16) unsigned char smem[];
  float4* s_box = reinterpret_cast<float4*>(smem);
  float* s_key = reinterpret_cast<float*>(s_box + L);
  int* s_id = reinterpret_cast<int*>(s_key + L);
  int* s_scan = s_id + L;
  unsigned char* s_sup = reinterpret_cast<unsigned char*>(s_scan + 64);
  unsigned char* s_flag = s_sup + L;

  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    s_key[t] = t < n ? __ldg(scores + t) : -INFINITY;
    s_id[t] = t < n ? t : INT_MAX;
    s_sup[t] = 0;
    if (kLegacy) s_flag[t] = 0;
  }
  __syncthreads();
  odw::cta_bitonic_sort(s_key, s_id, L);
  for (int t = threadIdx.x; t < n; t += blockDim.x) s_box[t] = __ldg(boxes + s_id[t]);
  __syncthreads();
  if (!kLegacy) {
    const int nk = odw::cta_nms_sweep(s_box, s_sup, n, thr, 0.f,
                                      [&](int k, int pos) { keep[k] = (int64_t)s_id[pos]; });
    if (threadIdx.x == 0) *n_keep = nk;
  } else {
    odw::cta_nms_sweep(s_box, s_sup, n, thr, 1.f, [&](int, int pos) { s_flag[s_id[pos]] = 1; });
    __syncthreads();
    const int nk = odw::cta_compact(
        n, s_scan, [&](int j) { return s_flag[j] != 0; }, [&](int k, int j) { keep[k] = (int64_t)j; });
    if (threadIdx.x == 0) *n_keep = nk;
  }
}

// Test-time filter (box_head/inference.py:216-258, filter_results): one CTA per foreground class j -- candidates with
// scores[i,j] > score_thr (ascending i), sorted by score (desc, index asc on ties), greedy NMS at `thr` without the +1
// (torchvision.ops.nms through boxlist_nms, structures/boxlist_ops.py:13-36).  boxes [N, C*4], scores [N, C];
// keep [C, N] int32 (row j: kept proposal indices in descending-score order), n_keep [C] (row 0 = background = 0).
__global__ void __launch_bounds__(odw::kCtaThreads, 1)
nms_per_class_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int n, int C, int L,
                     float score_thr, float thr, int32_t* __restrict__ keep, int32_t* __restrict__ n_keep) {
  extern __shared__ __align__(16) unsigned char smem[];
  float4* s_box = reinterpret_cast<float4*>(smem);
  float* s_key = reinterpret_cast<float*>(s_box + L);
  int* s_id = reinterpret_cast<int*>(s_key + L);
  int* s_scan = s_id + L;
  unsigned char* s_sup = reinterpret_cast<unsigned char*>(s_scan + 64);
  const int j = blockIdx.x + 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) n_keep[0] = 0;
  const int nc = odw::cta_compact(
      n, s_scan, [&](int i) { return __ldg(scores + (size_t)i * C + j) > score_thr; },
      [&](int k, int i) { s_key[k] = __ldg(scores + (size_t)i * C + j); s_id[k] = i; });
  int nk = 0;
  if (nc > 0) {
    const int L2 = odw::next_pow2(nc < 32 ? 32 : nc);
    for (int t = nc + threadIdx.x; t < L2; t += blockDim.x) { s_key[t] = -INFINITY; s_id[t] = INT_MAX; }
    for (int t = threadIdx.x; t < L2; t += blockDim.x) s_sup[t] = 0;
    __syncthreads();
    odw::cta_bitonic_sort(s_key, s_id, L2);
    for (int t = threadIdx.x; t < nc; t += blockDim.x)
      s_box[t] = __ldg(reinterpret_cast<const float4*>(boxes + ((size_t)s_id[t] * C + j) * 4));
    __syncthreads();
    int32_t* out = keep + (size_t)j * n;
    nk = odw::cta_nms_sweep(s_box, s_sup, nc, thr, 0.f, [&](int k, int pos) { out[k] = s_id[pos]; });
  }
  if (threadIdx.x == 0) n_keep[j] = nk;
}

template <bool kLegacy>
int launch_nms(const float* boxes, const float* scores, int n, float thr, int64_t* keep, int32_t* n_keep,
               odwscl_stream_t stream) {
  if (n < 0 || n > kNmsMax) return ODWSCL_EINVAL;
  if (!n_keep) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    ODW_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int32_t), st));
    return 0;
  }
  if (!boxes || !scores || !keep) return ODWSCL_EINVAL;
  int L = 32;
  while (L < n) L <<= 1;
  const size_t smem = (size_t)L * (16 + 4 + 4 + 1 + 1) + 64 * 4;
  ODW_CUDA(cudaFuncSetAttribute(nms_kernel<kLegacy>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_kernel<kLegacy><<<1, odw::kCtaThreads, smem, st>>>(reinterpret_cast<const float4*>(boxes), scores, n, L,
                                                         thr, keep, n_keep);
  ODW_LAUNCH_CHECK();
  return 0;
}

}  // namespace

ODW_API int odwscl_box_iou_f32(const float* a, int na, const float* b, int nb, int plus_one, float* out,
                               odwscl_stream_t stream) {
  if (na < 0 || nb < 0) return ODWSCL_EINVAL;
  if (na == 0 || nb == 0) return 0;
  if (!a || !b || !out) return ODWSCL_EINVAL;
  const long long total = (long long)na * nb;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 8, (total + 255) / 256);
  box_iou_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(a), na,
                                                           reinterpret_cast<const float4*>(b), nb,
                                                           plus_one ? 1.f : 0.f, out);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_nms_f32(const float* boxes, const float* scores, int n, float thr, int64_t* keep,
                           int32_t* n_keep, odwscl_stream_t stream) {
  return launch_nms<false>(boxes, scores, n, thr, keep, n_keep, stream);
}

ODW_API int odwscl_nms_legacy_f32(const float* boxes, const float* scores, int n, float thr, int64_t* keep,
                                  int32_t* n_keep, odwscl_stream_t stream) {
  return launch_nms<true>(boxes, scores, n, thr, keep, n_keep, stream);
}

ODW_API int odwscl_nms_per_class_f32(const float* boxes, const float* scores, int n, int C, float score_thr, float thr,
                                     int32_t* keep, int32_t* n_keep, odwscl_stream_t stream) {
  if (n < 0 || n > kNmsMax || C < 1) return ODWSCL_EINVAL;
  if (!n_keep) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 || C == 1) {
    ODW_CUDA(cudaMemsetAsync(n_keep, 0, (size_t)C * sizeof(int32_t), st));
    return 0;
  }
  if (!boxes || !scores || !keep) return ODWSCL_EINVAL;
  int L = 32;
  while (L < n) L <<= 1;
  const size_t smem = (size_t)L * (16 + 4 + 4 + 1) + 64 * 4;
  ODW_CUDA(cudaFuncSetAttribute(nms_per_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_per_class_kernel<<<C - 1, odw::kCtaThreads, smem, st>>>(boxes, scores, n, C, L, score_thr, thr, keep, n_keep);
  ODW_LAUNCH_CHECK();
  return 0;
}
