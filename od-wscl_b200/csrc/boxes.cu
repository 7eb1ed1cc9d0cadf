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

// ---------------------------------------------------------------------------------------------
// Large inputs (n > 8192: the single-CTA kernels keep every box in shared memory).  Same semantics, three launches:
//   nms_rank     stable descending rank of every candidate (score > score_thr) by counting -- O(n^2) compares, no sort
//                pass structure to get wrong; order[rank] = original index, *n_valid = candidates
//   nms_mask     64 x 64 blocks of the suppression relation between sorted boxes as bit masks
//                (bit j of mask[i][w]: box 64w+j suppressed by box i, j > i)
//   nms_sweep    one CTA walks the sorted list: a box survives iff its bit in the running `removed` bitmap is clear,
//                then ORs its mask row in -- the order-dependent part, kept on the device (torchvision copies the masks
//                to the host for this)
// boxes are read through a row stride so that one class column of a [N, C*4] regression output needs no copy.
__global__ void __launch_bounds__(256)
nms_rank_kernel(const float* __restrict__ scores, int score_stride, int n, float score_thr, int32_t* __restrict__ order,
                int32_t* __restrict__ n_valid) {
  __shared__ float s_sc[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float si = i < n ? __ldg(scores + (size_t)i * score_stride) : -INFINITY;
  const bool vi = i < n && si > score_thr;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    const int j = j0 + threadIdx.x;
    s_sc[threadIdx.x] = j < n ? __ldg(scores + (size_t)j * score_stride) : -INFINITY;
    __syncthreads();
    if (vi) {
      const int m = min(256, n - j0);
      for (int t = 0; t < m; ++t) {
        const float sj = s_sc[t];
        rank += (sj > score_thr && (sj > si || (sj == si && j0 + t < i))) ? 1 : 0;
      }
    }
    __syncthreads();
  }
  if (vi) order[rank] = i;
  const int cnt = __syncthreads_count(vi ? 1 : 0);
  if (threadIdx.x == 0 && cnt) atomicAdd(n_valid, cnt);
}

__global__ void __launch_bounds__(64)
nms_mask_kernel(const float* __restrict__ boxes, int box_stride, const int32_t* __restrict__ order,
                const int32_t* __restrict__ n_valid, float thr, float one, int ge, unsigned long long* __restrict__ mask,
                int words) {
  const int nv = *n_valid;
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (rb * 64 >= nv || cb * 64 >= nv || cb < rb) return;
  __shared__ float4 s_col[64];
  const int cj = cb * 64 + threadIdx.x;
  if (cj < nv) s_col[threadIdx.x] = *reinterpret_cast<const float4*>(boxes + (size_t)order[cj] * box_stride);
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= nv) return;
  const float4 bi = *reinterpret_cast<const float4*>(boxes + (size_t)order[i] * box_stride);
  unsigned long long bits = 0ull;
  const int jn = min(64, nv - cb * 64);
  for (int t = (rb == cb ? threadIdx.x + 1 : 0); t < jn; ++t) {
    const float v = odw_iou(bi, s_col[t], one);
    if (ge ? v >= thr : v > thr) bits |= 1ull << t;
  }
  mask[(size_t)i * words + cb] = bits;
}

__global__ void __launch_bounds__(1024)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, int words, const int32_t* __restrict__ order,
                 const int32_t* __restrict__ n_valid, int sorted_out, int64_t* __restrict__ keep64,
                 int32_t* __restrict__ keep32, int32_t* __restrict__ n_keep, unsigned char* __restrict__ flag) {
  extern __shared__ unsigned long long s_removed[];          // [words]
  const int nv = *n_valid;
  const int nw = (nv + 63) / 64;
  for (int w = threadIdx.x; w < nw; w += blockDim.x) s_removed[w] = 0ull;
  __syncthreads();
  int nk = 0;
  for (int i = 0; i < nv; ++i) {
    const bool alive = !((s_removed[i >> 6] >> (i & 63)) & 1ull);        // uniform
    if (alive) {
      if (threadIdx.x == 0) {
        const int id = order[i];
        if (sorted_out) { if (keep64) keep64[nk] = id; else keep32[nk] = id; }
        else flag[id] = 1;
      }
      ++nk;
      const unsigned long long* row = mask + (size_t)i * words;
      for (int w = (i >> 6) + threadIdx.x; w < nw; w += blockDim.x) s_removed[w] |= row[w];
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) *n_keep = nk;
}

// ascending original indices of the flagged boxes (the legacy `_C.nms` output order)
__global__ void __launch_bounds__(odw::kCtaThreads, 1)
nms_flag_compact_kernel(const unsigned char* __restrict__ flag, int n, int64_t* __restrict__ keep) {
  __shared__ int s_scan[64];
  odw::cta_compact(n, s_scan, [&](int j) { return flag[j] != 0; }, [&](int k, int j) { keep[k] = (int64_t)j; });
}

size_t nms_large_ws_bytes(int n) {
  const size_t words = (size_t)(n + 63) / 64;
  return odw_align(4 * (size_t)n) + odw_align(8 * (size_t)n * words) + odw_align((size_t)n) + 256;
}

// keep64 / keep32: exactly one is non-null.  legacy: +1 IoU (suppress iff IoU > thr, csrc/cuda/nms.cu), ascending output.
int launch_nms_large(const float* boxes, int box_stride, const float* scores, int score_stride, int n, float score_thr,
                     float thr, bool legacy, int64_t* keep64, int32_t* keep32, int32_t* n_keep, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  if (!ws || ws_bytes < nms_large_ws_bytes(n)) return ODWSCL_ENOWS;
  if ((box_stride & 3) || ((uintptr_t)boxes & 15)) return ODWSCL_EINVAL;      // 16-byte box loads
  const int words = (n + 63) / 64;
  if ((size_t)words * 8 > 200 * 1024) return ODWSCL_EINVAL;                    // removed bitmap in shared memory: n <= 1.6 M
  unsigned char* p = reinterpret_cast<unsigned char*>(ws);
  int32_t* order = reinterpret_cast<int32_t*>(p); p += odw_align(4 * (size_t)n);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(p); p += odw_align(8 * (size_t)n * words);
  unsigned char* flag = p; p += odw_align((size_t)n);
  int32_t* n_valid = reinterpret_cast<int32_t*>(p);
  ODW_CUDA(cudaMemsetAsync(n_valid, 0, sizeof(int32_t), st));
  if (legacy) ODW_CUDA(cudaMemsetAsync(flag, 0, (size_t)n, st));
  nms_rank_kernel<<<odw_cdiv(n, 256), 256, 0, st>>>(scores, score_stride, n, score_thr, order, n_valid);
  ODW_LAUNCH_CHECK();
  nms_mask_kernel<<<dim3(words, words), 64, 0, st>>>(boxes, box_stride, order, n_valid, thr, legacy ? 1.f : 0.f,
                                                      0 /* strict '>' in both conventions (csrc/cuda/nms.cu:41) */, mask, words);
  ODW_LAUNCH_CHECK();
  const size_t smem = (size_t)words * 8;
  if (smem > 48 * 1024)
    ODW_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_sweep_kernel<<<1, 1024, smem, st>>>(mask, words, order, n_valid, legacy ? 0 : 1, keep64, keep32, n_keep, flag);
  ODW_LAUNCH_CHECK();
  if (legacy) {
    nms_flag_compact_kernel<<<1, odw::kCtaThreads, 0, st>>>(flag, n, keep64);
    ODW_LAUNCH_CHECK();
  }
  return 0;
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

ODW_API size_t odwscl_nms_large_ws_bytes(int n) { return n > 0 ? nms_large_ws_bytes(n) : 0; }

ODW_API int odwscl_nms_large_f32(const float* boxes, int box_stride, const float* scores, int score_stride, int n,
                                 float score_thr, float thr, int legacy, int64_t* keep64, int32_t* keep32, int32_t* n_keep,
                                 void* ws, size_t ws_bytes, odwscl_stream_t stream) {
  if (n < 0 || box_stride < 4 || score_stride < 1 || !n_keep || (!keep64) == (!keep32)) return ODWSCL_EINVAL;
  if (legacy && !keep64) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    ODW_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int32_t), st));
    return 0;
  }
  if (!boxes || !scores) return ODWSCL_EINVAL;
  return launch_nms_large(boxes, box_stride, scores, score_stride, n, score_thr, thr, legacy != 0, keep64, keep32, n_keep, ws,
                          ws_bytes, st);
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
