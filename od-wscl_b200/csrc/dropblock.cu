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

// ---------------------------------------------------------------------------------------------
// The augmented positives of the contrastive branch (roi_heads/weak_head/loss.py:296-305) in one pass over the pooled
// rows: for every Phase-A positive k (proposal rows[k]) the DropBlock(1x1, p = 0.3) view x * mask * numel/sum
// (vgg16.py:173-175, renormalised per (image, class) segment as the reference's per-group calls are) goes to out[k] and
// the multiplicative-noise view eps * x + x, eps ~ N(0,1) (vgg16.py:177-180) to out[Kc + k].  Replaces a gather, a
// DropBlock pass, randn, a multiply, an add and a concatenation (6 passes over [Kc, C*49]) and their backward.  The
// noise is Philox4x32-10 + Box-Muller keyed by (seed, k * D + e) -- regenerated in the backward -- unless a noise tensor
// is handed in (tests replay the CPU checker's draws); rows >= seg_off[P] are padding (zero).
namespace {

__device__ __forceinline__ void philox_round_db(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// four N(0,1) draws for counter `ctr`
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long ctr) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5EED0A06u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round_db(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float u0 = ((c[0] >> 8) + 1u) * (1.f / 16777216.f), u1 = (c[1] >> 8) * (1.f / 16777216.f);   // u0 in (0,1]
  const float u2 = ((c[2] >> 8) + 1u) * (1.f / 16777216.f), u3 = (c[3] >> 8) * (1.f / 16777216.f);
  const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.f * u1, &s0, &c0);
  sincospif(2.f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

template <bool kBackward>
__global__ void __launch_bounds__(256)
aug_positives_kernel(const float* __restrict__ src /* fwd: pooled [R,D]; bwd: g [2Kc,D] */, int D, int cells,
                     const int64_t* __restrict__ rows, int Kc, const int32_t* __restrict__ seg_off, int P,
                     const float* __restrict__ centres, int block, const float* __restrict__ scale_seg,
                     const float* __restrict__ noise, unsigned long long seed, float* __restrict__ dst) {
  extern __shared__ float s_bm[];
  const int k = blockIdx.x;
  const int ph = 7, pw = cells / 7;
  const bool valid = k < seg_off[P];
  float* d_drop = dst + (size_t)k * D;                               // fwd: out[k];  bwd: gx[k]
  float* d_noise = kBackward ? nullptr : dst + (size_t)(Kc + k) * D; // fwd: out[Kc + k]
  if (!valid) {
    for (int t = threadIdx.x; t < D / 4; t += blockDim.x) {
      reinterpret_cast<float4*>(d_drop)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!kBackward) reinterpret_cast<float4*>(d_noise)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  int lo = 0, hi = P - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_off[mid] <= k) lo = mid; else hi = mid - 1;
  }
  const float scale = scale_seg[2 * lo + 1];
  for (int b = threadIdx.x; b < cells; b += blockDim.x)
    s_bm[b] = block_mask_at(centres + (size_t)k * cells, ph, pw, block, b / pw, b % pw) * scale;
  __syncthreads();
  const float4* a4 = reinterpret_cast<const float4*>(kBackward ? src + (size_t)k * D : src + (size_t)rows[k] * D);
  const float4* b4 = kBackward ? reinterpret_cast<const float4*>(src + (size_t)(Kc + k) * D) : nullptr;
  const float4* n4 = noise ? reinterpret_cast<const float4*>(noise + (size_t)k * D) : nullptr;
  for (int t = threadIdx.x; t < D / 4; t += blockDim.x) {
    const float4 x = __ldg(a4 + t);                                  // fwd: pooled row;  bwd: gradient of the drop view
    const float4 e = n4 ? __ldg(n4 + t) : philox_normal4(seed, ((unsigned long long)k * D + 4ull * t) >> 2);
    const int q = (4 * t) % cells;
    const float m0 = s_bm[q], m1 = s_bm[q + 1 < cells ? q + 1 : q + 1 - cells], m2 = s_bm[q + 2 < cells ? q + 2 : q + 2 - cells],
                m3 = s_bm[q + 3 < cells ? q + 3 : q + 3 - cells];
    if (!kBackward) {
      reinterpret_cast<float4*>(d_drop)[t] = make_float4(x.x * m0, x.y * m1, x.z * m2, x.w * m3);
      reinterpret_cast<float4*>(d_noise)[t] = make_float4(__fadd_rn(__fmul_rn(e.x, x.x), x.x), __fadd_rn(__fmul_rn(e.y, x.y), x.y),
                                                         __fadd_rn(__fmul_rn(e.z, x.z), x.z), __fadd_rn(__fmul_rn(e.w, x.w), x.w));
    } else {
      const float4 g2 = __ldg(b4 + t);                               // gradient of the noise view: d/dx = eps + 1
      reinterpret_cast<float4*>(d_drop)[t] =
          make_float4(__fadd_rn(x.x * m0, __fadd_rn(__fmul_rn(g2.x, e.x), g2.x)), __fadd_rn(x.y * m1, __fadd_rn(__fmul_rn(g2.y, e.y), g2.y)),
                      __fadd_rn(x.z * m2, __fadd_rn(__fmul_rn(g2.z, e.z), g2.z)), __fadd_rn(x.w * m3, __fadd_rn(__fmul_rn(g2.w, e.w), g2.w)));
    }
  }
}

}  // namespace

// forward: compute_scale != 0 first fills scale_seg from the centres (per-segment numel / sum)
ODW_API int odwscl_aug_positives_f32(const float* src, int D, int cells, const int64_t* rows, int Kc, const int32_t* seg_off_dev,
                                     int P, const float* centres, int block, float* scale_seg, int compute_scale,
                                     const float* noise, unsigned long long seed, int backward, float* dst,
                                     odwscl_stream_t stream) {
  if (D <= 0 || (D & 3) || cells <= 0 || cells % 7 || D % cells || Kc < 0 || P <= 0 || block <= 0) return ODWSCL_EINVAL;
  if (Kc == 0) return 0;
  if (!src || !seg_off_dev || !centres || !scale_seg || !dst || (!backward && !rows)) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (compute_scale) {
    dropblock_seg_sum_kernel<<<P, 256, 0, st>>>(centres, Kc, 7, cells / 7, block, seg_off_dev, scale_seg);
    ODW_LAUNCH_CHECK();
  }
  if (backward)
    aug_positives_kernel<true><<<Kc, 256, cells * sizeof(float), st>>>(src, D, cells, rows, Kc, seg_off_dev, P, centres, block,
                                                                       scale_seg, noise, seed, dst);
  else
    aug_positives_kernel<false><<<Kc, 256, cells * sizeof(float), st>>>(src, D, cells, rows, Kc, seg_off_dev, P, centres, block,
                                                                        scale_seg, noise, seed, dst);
  ODW_LAUNCH_CHECK();
  return 0;
}
