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

}  // namespace

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
