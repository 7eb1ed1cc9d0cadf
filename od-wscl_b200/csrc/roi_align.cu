// roi_align.cu -- legacy (non-aligned) ROIAlign forward / backward.
// Reference: wetectron/csrc/cuda/ROIAlign_cuda.cu:64-122 (fwd), :177-254 (bwd).  No shipped
// config selects ROIAlign (all use ROIPool); it is on the operator surface, so it is provided
// with the same sampling rule: no coordinate rounding, roi extent clamped to >= 1, adaptive
// ceil(roi/pooled) sampling grid when sampling_ratio <= 0, samples outside [-1, H] x [-1, W]
// contribute 0.  One thread per output scalar; sample weights are computed once per sample and
// shared by forward and backward through `Sample`.
#include "common.cuh"

namespace {

struct Sample {
  int lo_y, hi_y, lo_x, hi_x;
  float w_ll, w_lh, w_hl, w_hh;
  bool valid;
};

__device__ __forceinline__ Sample make_sample(float y, float x, int H, int W) {
  Sample s;
  s.valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
  if (!s.valid) return s;
  y = fmaxf(y, 0.f);
  x = fmaxf(x, 0.f);
  s.lo_y = (int)y;
  s.lo_x = (int)x;
  if (s.lo_y >= H - 1) { s.hi_y = s.lo_y = H - 1; y = (float)s.lo_y; } else s.hi_y = s.lo_y + 1;
  if (s.lo_x >= W - 1) { s.hi_x = s.lo_x = W - 1; x = (float)s.lo_x; } else s.hi_x = s.lo_x + 1;
  const float ly = y - (float)s.lo_y, lx = x - (float)s.lo_x;
  const float hy = 1.f - ly, hx = 1.f - lx;
  s.w_ll = hy * hx; s.w_lh = hy * lx; s.w_hl = ly * hx; s.w_hh = ly * lx;
  return s;
}

struct AlignGeom {
  int b, gh, gw;
  float sh, sw, bh, bw;
};

__device__ __forceinline__ AlignGeom align_geom(const float* __restrict__ roi, float scale, int PH, int PW,
                                                int sampling_ratio) {
  AlignGeom g;
  g.b = (int)roi[0];
  g.sw = roi[1] * scale; g.sh = roi[2] * scale;
  const float ew = roi[3] * scale, eh = roi[4] * scale;
  const float rw = fmaxf(ew - g.sw, 1.f), rh = fmaxf(eh - g.sh, 1.f);
  g.bh = rh / (float)PH; g.bw = rw / (float)PW;
  g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
  g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
  return g;
}

template <bool kBackward>
__global__ void roi_align_kernel(const float* __restrict__ src, const float* __restrict__ rois,
                                 long long total, int C, int H, int W, float scale, int PH, int PW,
                                 int sampling_ratio, float* __restrict__ dst) {
  for (long long index = blockIdx.x * (long long)blockDim.x + threadIdx.x; index < total;
       index += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(index % PW);
    const int ph = (int)((index / PW) % PH);
    const int c = (int)((index / PW / PH) % C);
    const int n = (int)(index / PW / PH / C);
    const AlignGeom g = align_geom(rois + (size_t)n * 5, scale, PH, PW, sampling_ratio);
    const float count = (float)(g.gh * g.gw);
    const size_t plane = ((size_t)g.b * C + c) * H * W;
    float acc = 0.f;
    const float gtop = kBackward ? src[index] : 0.f;
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = g.sh + ph * g.bh + ((float)iy + .5f) * g.bh / (float)g.gh;
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = g.sw + pw * g.bw + ((float)ix + .5f) * g.bw / (float)g.gw;
        const Sample s = make_sample(y, x, H, W);
        if (!s.valid) continue;
        if (kBackward) {
          float* p = dst + plane;
          atomicAdd(p + s.lo_y * W + s.lo_x, gtop * s.w_ll / count);
          atomicAdd(p + s.lo_y * W + s.hi_x, gtop * s.w_lh / count);
          atomicAdd(p + s.hi_y * W + s.lo_x, gtop * s.w_hl / count);
          atomicAdd(p + s.hi_y * W + s.hi_x, gtop * s.w_hh / count);
        } else {
          const float* p = src + plane;
          acc += s.w_ll * __ldg(p + s.lo_y * W + s.lo_x) + s.w_lh * __ldg(p + s.lo_y * W + s.hi_x) +
                 s.w_hl * __ldg(p + s.hi_y * W + s.lo_x) + s.w_hh * __ldg(p + s.hi_y * W + s.hi_x);
        }
      }
    }
    if (!kBackward) dst[index] = acc / count;
  }
}

// ---------------------------------------------------------------------------------------------
// Channels-last (NHWC) path, the layout the conv stack hands over -- same sampling rule, re-mapped for B200:
//   forward   CTA = (roi, 128-channel slab), warp = bin row, lane = 4 consecutive channels: one 16-byte load per bilinear
//             corner serves 4 channels (512 contiguous bytes per warp) and the sample weights -- a function of (roi, bin,
//             sample) only -- are computed once per warp instead of once per channel; results staged in shared memory
//             in [c][49] order and written as contiguous 16-byte stores (the reference walks a stride-W plane per thread).
//   backward  plane-centric like the ROIPool backward: a CTA owns 4 channels of one image for a chunk of rois and
//             accumulates the 4 x samples x 49 scatter-adds per (roi, channel) in SHARED memory (the reference issues them
//             as global atomics); one flush per CTA with red.global.add.v4.f32.
constexpr int kAlSlab = 128, kAlBins = 49;

__global__ void __launch_bounds__(7 * 32, 4)
roi_align_fwd_nhwc7_kernel(const float4* __restrict__ feat4, const float* __restrict__ rois, int C, int H, int W, float scale,
                           int sampling_ratio, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_val[];              // [128][49]
  const int CQ = C >> 2;
  const int n = blockIdx.x, c0 = blockIdx.y * kAlSlab;
  const int nch = min(kAlSlab, C - c0);
  const int ph = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AlignGeom g = align_geom(rois + (size_t)n * 5, scale, 7, 7, sampling_ratio);
  const float count = (float)(g.gh * g.gw);
  if (4 * lane < nch) {
    const float4* base = feat4 + (size_t)g.b * H * W * CQ + (c0 >> 2) + lane;
#pragma unroll 1
    for (int pw = 0; pw < 7; ++pw) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int iy = 0; iy < g.gh; ++iy) {
        const float y = g.sh + ph * g.bh + ((float)iy + .5f) * g.bh / (float)g.gh;
        for (int ix = 0; ix < g.gw; ++ix) {
          const float x = g.sw + pw * g.bw + ((float)ix + .5f) * g.bw / (float)g.gw;
          const Sample s = make_sample(y, x, H, W);
          if (!s.valid) continue;
          const float4 a = __ldg(base + (size_t)(s.lo_y * W + s.lo_x) * CQ), b = __ldg(base + (size_t)(s.lo_y * W + s.hi_x) * CQ);
          const float4 c = __ldg(base + (size_t)(s.hi_y * W + s.lo_x) * CQ), d = __ldg(base + (size_t)(s.hi_y * W + s.hi_x) * CQ);
          acc.x += s.w_ll * a.x + s.w_lh * b.x + s.w_hl * c.x + s.w_hh * d.x;
          acc.y += s.w_ll * a.y + s.w_lh * b.y + s.w_hl * c.y + s.w_hh * d.y;
          acc.z += s.w_ll * a.z + s.w_lh * b.z + s.w_hl * c.z + s.w_hh * d.z;
          acc.w += s.w_ll * a.w + s.w_lh * b.w + s.w_hl * c.w + s.w_hh * d.w;
        }
      }
      const int o = (4 * lane) * kAlBins + ph * 7 + pw;
      s_val[o] = acc.x / count; s_val[o + kAlBins] = acc.y / count;
      s_val[o + 2 * kAlBins] = acc.z / count; s_val[o + 3 * kAlBins] = acc.w / count;
    }
  }
  __syncthreads();
  const size_t obase = ((size_t)n * C + c0) * kAlBins;
  float4* o4 = reinterpret_cast<float4*>(out + obase);
  const float4* sv4 = reinterpret_cast<const float4*>(s_val);
  for (int i = threadIdx.x; i < nch * kAlBins / 4; i += blockDim.x) __stcs(o4 + i, sv4[i]);
}

constexpr int kAlChunk = 256;                                 // rois per backward CTA

__global__ void __launch_bounds__(512, 1)
roi_align_bwd_plane_kernel(const float* __restrict__ grad_out, const float* __restrict__ rois, int R, int C, int H, int W,
                           float scale, int sampling_ratio, float* __restrict__ grad_in_nhwc) {
  extern __shared__ __align__(16) float acc[];                // [H*W][4]
  __shared__ int s_list[kAlChunk];
  __shared__ int s_n;
  const int HW = H * W;
  const int cg = blockIdx.x, b = blockIdx.y;
  const int r0 = blockIdx.z * kAlChunk, r1 = min(R, r0 + kAlChunk);
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x)
    if ((int)__ldg(rois + (size_t)r * 5) == b) s_list[atomicAdd(&s_n, 1)] = r;
  __syncthreads();
  const int nroi = s_n;
  if (nroi == 0) return;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  // thread = (roi of the chunk, bin): the sample weights are shared by the 4 channels
  for (int it = threadIdx.x; it < nroi * kAlBins; it += blockDim.x) {
    const int rl = it / kAlBins, bin = it - rl * kAlBins;
    const int r = s_list[rl], ph = bin / 7, pw = bin - ph * 7;
    const AlignGeom g = align_geom(rois + (size_t)r * 5, scale, 7, 7, sampling_ratio);
    const float count = (float)(g.gh * g.gw);
    const float* go = grad_out + ((size_t)r * C + (size_t)cg * 4) * kAlBins + bin;
    const float g0 = __ldg(go), g1 = __ldg(go + kAlBins), g2 = __ldg(go + 2 * kAlBins), g3 = __ldg(go + 3 * kAlBins);
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = g.sh + ph * g.bh + ((float)iy + .5f) * g.bh / (float)g.gh;
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = g.sw + pw * g.bw + ((float)ix + .5f) * g.bw / (float)g.gw;
        const Sample s = make_sample(y, x, H, W);
        if (!s.valid) continue;
        const float w[4] = {s.w_ll, s.w_lh, s.w_hl, s.w_hh};
        const int cell[4] = {s.lo_y * W + s.lo_x, s.lo_y * W + s.hi_x, s.hi_y * W + s.lo_x, s.hi_y * W + s.hi_x};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float* a = acc + cell[k] * 4;
          atomicAdd(a, g0 * w[k] / count); atomicAdd(a + 1, g1 * w[k] / count);
          atomicAdd(a + 2, g2 * w[k] / count); atomicAdd(a + 3, g3 * w[k] / count);
        }
      }
    }
  }
  __syncthreads();
  float* dst = grad_in_nhwc + (size_t)b * HW * C + (size_t)cg * 4;
  for (int cell = threadIdx.x; cell < HW; cell += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(acc)[cell];
    if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + (size_t)cell * C), "f"(v.x), "f"(v.y), "f"(v.z),
                   "f"(v.w) : "memory");
  }
}

}  // namespace

ODW_API int odwscl_roi_align_fwd_nhwc_f32(const float* feat_nhwc, int B, int C, int H, int W, const float* rois, int R,
                                          float scale, int sampling_ratio, float* out, odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H <= 0 || W <= 0 || R < 0 || (C & 3)) return ODWSCL_EINVAL;
  if (R == 0 || C == 0) return 0;
  if (!feat_nhwc || !rois || !out) return ODWSCL_EINVAL;
  const int smem = kAlSlab * kAlBins * (int)sizeof(float);
  dim3 grid(R, odw_cdiv(C, kAlSlab));
  roi_align_fwd_nhwc7_kernel<<<grid, 7 * 32, smem, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(feat_nhwc), rois, C, H,
                                                                          W, scale, sampling_ratio, out);
  ODW_LAUNCH_CHECK();
  return 0;
}

// grad_in_nhwc [B,H,W,C] is zeroed here.  Returns ODWSCL_ENOWS when one 4-channel plane (16 * H * W bytes) does not fit
// the 227 KB of shared memory (the caller then uses the NCHW entry point).
ODW_API int odwscl_roi_align_bwd_nhwc_f32(const float* grad_out, const float* rois, int R, float scale, int B, int C, int H,
                                          int W, int sampling_ratio, float* grad_in_nhwc, odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || R < 0 || (C & 3)) return ODWSCL_EINVAL;
  const size_t bytes = (size_t)B * C * H * W * sizeof(float);
  if (bytes == 0) return 0;
  if (!grad_in_nhwc) return ODWSCL_EINVAL;
  const size_t smem = (size_t)H * W * 4 * sizeof(float);
  if (smem > 220 * 1024) return ODWSCL_ENOWS;
  cudaStream_t st = (cudaStream_t)stream;
  ODW_CUDA(cudaMemsetAsync(grad_in_nhwc, 0, bytes, st));
  if (R == 0) return 0;
  if (!grad_out || !rois) return ODWSCL_EINVAL;
  ODW_CUDA(cudaFuncSetAttribute(roi_align_bwd_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(C / 4, B, odw_cdiv(R, kAlChunk));
  roi_align_bwd_plane_kernel<<<grid, 512, smem, st>>>(grad_out, rois, R, C, H, W, scale, sampling_ratio, grad_in_nhwc);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_roi_align_fwd_f32(const float* feat, int B, int C, int H, int W, const float* rois, int R,
                                     float scale, int ph, int pw, int sampling_ratio, float* out,
                                     odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || R < 0 || ph <= 0 || pw <= 0) return ODWSCL_EINVAL;
  if (R == 0 || C == 0) return 0;
  if (!feat || !rois || !out) return ODWSCL_EINVAL;
  const long long total = (long long)R * C * ph * pw;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 16, (total + 255) / 256);
  roi_align_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(feat, rois, total, C, H, W, scale, ph, pw,
                                                                    sampling_ratio, out);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_roi_align_bwd_f32(const float* grad_out, const float* rois, int R, float scale, int ph,
                                     int pw, int B, int C, int H, int W, int sampling_ratio, float* grad_in,
                                     odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || R < 0 || ph <= 0 || pw <= 0) return ODWSCL_EINVAL;
  const size_t bytes = (size_t)B * C * H * W * sizeof(float);
  if (bytes == 0) return 0;
  if (!grad_in) return ODWSCL_EINVAL;
  ODW_CUDA(cudaMemsetAsync(grad_in, 0, bytes, (cudaStream_t)stream));
  if (R == 0) return 0;
  if (!grad_out || !rois) return ODWSCL_EINVAL;
  const long long total = (long long)R * C * ph * pw;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 16, (total + 255) / 256);
  roi_align_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(grad_out, rois, total, C, H, W, scale, ph, pw,
                                                                   sampling_ratio, grad_in);
  ODW_LAUNCH_CHECK();
  return 0;
}
