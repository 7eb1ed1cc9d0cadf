// dropblock.cu -- DropBlock2D apply.
// Reference: modeling/dropblock/drop_block.py:29-66 -- CPU-sampled mask, H2D copy, max_pool2d,
// two full elementwise passes over [R,C,7,7] (200 MB each at R=2000) plus a global reduction.
// Here the centre mask is sampled on the device by the caller; the block mask (block x block
// dilation of the centres, window rows y - block/2 .. y - block/2 + block - 1, as max_pool2d with
// padding block/2 and the even-size crop of :61-62) is rebuilt per roi in shared memory and applied
// in ONE pass:  y = x * block_mask * (numel / sum(block_mask)).
// scale_io[0] = sum(block_mask), scale_io[1] = numel / sum.  With reuse_scale != 0 the stored scale
// is used (backward: dx = dy * block_mask * scale).
#include "common.cuh"

namespace {

__device__ __forceinline__ float block_mask_at(const float* __restrict__ cen, int ph, int pw, int block, int y,
                                               int x) {
  const int y0 = y - block / 2, x0 = x - block / 2;
  float mx = 0.f;
  for (int dy = 0; dy < block; ++dy)
    for (int dx = 0; dx < block; ++dx) {
      const int yy = y0 + dy, xx = x0 + dx;
      if (yy >= 0 && yy < ph && xx >= 0 && xx < pw) mx = fmaxf(mx, cen[yy * pw + xx]);
    }
  return 1.f - mx;
}

__global__ void dropblock_sum_kernel(const float* __restrict__ centres, int R, int ph, int pw, int block,
                                     float* __restrict__ scale_io, const int32_t* __restrict__ n_valid) {
  const int cells = ph * pw;
  if (n_valid != nullptr) R = min(R, max(*n_valid, 0));
  float part = 0.f;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)R * cells;
       t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / cells), k = (int)(t % cells);
    part += block_mask_at(centres + (size_t)r * cells, ph, pw, block, k / pw, k % pw);
  }
  part = odw_warp_sum(part);
  if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(scale_io, part);   // integer-valued: order-free
}

__global__ void dropblock_scale_kernel(int R, int ph, int pw, float* __restrict__ scale_io,
                                       const int32_t* __restrict__ n_valid) {
  if (n_valid != nullptr) R = min(R, max(*n_valid, 0));
  scale_io[1] = (float)((long long)R * ph * pw) / scale_io[0];
}

__global__ void __launch_bounds__(256)
dropblock_apply_kernel(const float* __restrict__ x, const float* __restrict__ centres, int C, int ph, int pw,
                       int block, const float* __restrict__ scale_io, float* __restrict__ y,
                       const int32_t* __restrict__ n_valid) {
  extern __shared__ float s_bm[];
  const int r = blockIdx.x, cells = ph * pw;
  if (n_valid != nullptr && r >= *n_valid) {               // padding row: defined output, no contribution
    for (int t = threadIdx.x; t < C * cells; t += blockDim.x) y[(size_t)r * C * cells + t] = 0.f;
    return;
  }
  const float scale = scale_io[1];
  for (int k = threadIdx.x; k < cells; k += blockDim.x)
    s_bm[k] = block_mask_at(centres + (size_t)r * cells, ph, pw, block, k / pw, k % pw) * scale;
  __syncthreads();
  const size_t base = (size_t)r * C * cells;
  const int n = C * cells;
  if ((n & 3) == 0) {                                        // 16-byte streaming loads / stores ((r*n) % 4 == 0)
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    float4* y4 = reinterpret_cast<float4*>(y + base);
    for (int t = threadIdx.x; t < n / 4; t += blockDim.x) {
      float4 v = __ldcs(x4 + t);
      const int k = (4 * t) % cells;                         // 4 consecutive bins, wrapping into the next channel
      v.x *= s_bm[k];
      v.y *= s_bm[k + 1 < cells ? k + 1 : k + 1 - cells];
      v.z *= s_bm[k + 2 < cells ? k + 2 : k + 2 - cells];
      v.w *= s_bm[k + 3 < cells ? k + 3 : k + 3 - cells];
      __stcs(y4 + t, v);
    }
    return;
  }
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const float v = __ldcs(x + base + t) * s_bm[t % cells];
    __stcs(y + base + t, v);
  }
}

__global__ void dropblock_mask_kernel(const float* __restrict__ centres, int R, int ph, int pw, int block,
                                      const float* __restrict__ scale_io, float* __restrict__ mask_out) {
  const int cells = ph * pw;
  const float scale = scale_io[1];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)R * cells;
       t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / cells), k = (int)(t % cells);
    mask_out[t] = block_mask_at(centres + (size_t)r * cells, ph, pw, block, k / pw, k % pw) * scale;
  }
}

// ---- segmented variant: rows [seg_off[p], seg_off[p+1]) form segment p (one (image, class) pair of the contrastive
// branch); every segment is renormalised on its own, exactly as the reference's per-pair drop_pool calls
// (roi_heads/weak_head/loss.py:299 -> vgg16.py:173-175 -> drop_block.py:53).  scale_seg [P,2] = (sum, numel/sum).
__global__ void __launch_bounds__(256)
dropblock_seg_sum_kernel(const float* __restrict__ centres, int R, int ph, int pw, int block,
                         const int32_t* __restrict__ seg_off, float* __restrict__ scale_seg) {
  const int p = blockIdx.x, cells = ph * pw;
  const int r0 = min(max(seg_off[p], 0), R), r1 = min(max(seg_off[p + 1], r0), R);
  float part = 0.f;
  for (long long t = (long long)r0 * cells + threadIdx.x; t < (long long)r1 * cells; t += blockDim.x) {
    const int r = (int)(t / cells), k = (int)(t % cells);
    part += block_mask_at(centres + (size_t)r * cells, ph, pw, block, k / pw, k % pw);
  }
  __shared__ float sm[8];
  part = odw_warp_sum(part);                                   // integer-valued partial sums: exact in any order
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += sm[i];
    scale_seg[2 * p] = tot;
    scale_seg[2 * p + 1] = (float)((long long)(r1 - r0) * cells) / tot;
  }
}

__global__ void __launch_bounds__(256)
dropblock_seg_apply_kernel(const float* __restrict__ x, const float* __restrict__ centres, int C, int ph, int pw,
                           int block, const int32_t* __restrict__ seg_off, int P, const float* __restrict__ scale_seg,
                           float* __restrict__ y) {
  extern __shared__ float s_bm[];
  const int r = blockIdx.x, cells = ph * pw;
  const size_t base = (size_t)r * C * cells;
  const int n = C * cells;
  if (r >= seg_off[P]) {                                       // padding row
    for (int t = threadIdx.x; t < n; t += blockDim.x) y[base + t] = 0.f;
    return;
  }
  int lo = 0, hi = P - 1;                                      // segment of row r: last p with seg_off[p] <= r
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_off[mid] <= r) lo = mid; else hi = mid - 1;
  }
  const float scale = scale_seg[2 * lo + 1];
  for (int k = threadIdx.x; k < cells; k += blockDim.x)
    s_bm[k] = block_mask_at(centres + (size_t)r * cells, ph, pw, block, k / pw, k % pw) * scale;
  __syncthreads();
  if ((n & 3) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    float4* y4 = reinterpret_cast<float4*>(y + base);
    for (int t = threadIdx.x; t < n / 4; t += blockDim.x) {
      float4 v = __ldcs(x4 + t);
      const int k = (4 * t) % cells;
      v.x *= s_bm[k];
      v.y *= s_bm[k + 1 < cells ? k + 1 : k + 1 - cells];
      v.z *= s_bm[k + 2 < cells ? k + 2 : k + 2 - cells];
      v.w *= s_bm[k + 3 < cells ? k + 3 : k + 3 - cells];
      __stcs(y4 + t, v);
    }
    return;
  }
  for (int t = threadIdx.x; t < n; t += blockDim.x) __stcs(y + base + t, __ldcs(x + base + t) * s_bm[t % cells]);
}

}  // namespace

ODW_API int odwscl_dropblock_mask_f32(const float* centres, int R, int ph, int pw, int block, const float* scale_io,
                                      float* mask_out, odwscl_stream_t stream) {
  if (R < 0 || ph <= 0 || pw <= 0 || block <= 0) return ODWSCL_EINVAL;
  if (R == 0) return 0;
  if (!centres || !scale_io || !mask_out) return ODWSCL_EINVAL;
  const long long total = (long long)R * ph * pw;
  dropblock_mask_kernel<<<(int)min((long long)ODW_NUM_SMS * 4, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      centres, R, ph, pw, block, scale_io, mask_out);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_dropblock_rows_f32(const float* x, const float* centres, int R, int C, int ph, int pw, int block,
                                      float* y, float* scale_io, int reuse_scale, const int32_t* n_valid_dev,
                                      odwscl_stream_t stream) {
  if (R < 0 || C < 0 || ph <= 0 || pw <= 0 || block <= 0) return ODWSCL_EINVAL;
  if (R == 0 || C == 0) return 0;
  if (!x || !centres || !y || !scale_io) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (!reuse_scale) {
    ODW_CUDA(cudaMemsetAsync(scale_io, 0, 2 * sizeof(float), st));
    const long long total = (long long)R * ph * pw;
    dropblock_sum_kernel<<<(int)min((long long)ODW_NUM_SMS, (total + 255) / 256), 256, 0, st>>>(centres, R, ph, pw,
                                                                                             block, scale_io, n_valid_dev);
    ODW_LAUNCH_CHECK();
    dropblock_scale_kernel<<<1, 1, 0, st>>>(R, ph, pw, scale_io, n_valid_dev);
    ODW_LAUNCH_CHECK();
  }
  dropblock_apply_kernel<<<R, 256, ph * pw * sizeof(float), st>>>(x, centres, C, ph, pw, block, scale_io, y,
                                                                  n_valid_dev);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_dropblock_f32(const float* x, const float* centres, int R, int C, int ph, int pw, int block,
                                 float* y, float* scale_io, int reuse_scale, odwscl_stream_t stream) {
  return odwscl_dropblock_rows_f32(x, centres, R, C, ph, pw, block, y, scale_io, reuse_scale, nullptr, stream);
}

ODW_API int odwscl_dropblock_seg_f32(const float* x, const float* centres, int R, int C, int ph, int pw, int block,
                                     float* y, const int32_t* seg_off_dev, int P, float* scale_seg, int reuse_scale,
                                     odwscl_stream_t stream) {
  if (R < 0 || C < 0 || P < 0 || ph <= 0 || pw <= 0 || block <= 0) return ODWSCL_EINVAL;
  if (R == 0 || C == 0) return 0;
  if (!x || !centres || !y || !seg_off_dev || !scale_seg || P == 0) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (!reuse_scale) {
    dropblock_seg_sum_kernel<<<P, 256, 0, st>>>(centres, R, ph, pw, block, seg_off_dev, scale_seg);
    ODW_LAUNCH_CHECK();
  }
  dropblock_seg_apply_kernel<<<R, 256, ph * pw * sizeof(float), st>>>(x, centres, C, ph, pw, block, seg_off_dev, P,
                                                                      scale_seg, y);
  ODW_LAUNCH_CHECK();
  return 0;
}

// scale_io and the per-(roi, bin) factor block_mask * numel/sum from the centre mask alone (both are independent of the
// values DropBlock multiplies): lets the ROIPool forward write the augmented copy in its own epilogue.
ODW_API int odwscl_dropblock_prepare_f32(const float* centres, int R, int ph, int pw, int block, float* scale_io,
                                         float* mask_out, odwscl_stream_t stream) {
  if (R < 0 || ph <= 0 || pw <= 0 || block <= 0) return ODWSCL_EINVAL;
  if (R == 0) return 0;
  if (!centres || !scale_io || !mask_out) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  ODW_CUDA(cudaMemsetAsync(scale_io, 0, 2 * sizeof(float), st));
  const long long total = (long long)R * ph * pw;
  dropblock_sum_kernel<<<(int)min((long long)ODW_NUM_SMS, (total + 255) / 256), 256, 0, st>>>(centres, R, ph, pw, block,
                                                                                           scale_io, nullptr);
  ODW_LAUNCH_CHECK();
  dropblock_scale_kernel<<<1, 1, 0, st>>>(R, ph, pw, scale_io, nullptr);
  ODW_LAUNCH_CHECK();
  dropblock_mask_kernel<<<(int)min((long long)ODW_NUM_SMS * 4, (total + 255) / 256), 256, 0, st>>>(centres, R, ph, pw, block,
                                                                                                  scale_io, mask_out);
  ODW_LAUNCH_CHECK();
  return 0;
}
