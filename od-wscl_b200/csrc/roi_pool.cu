// roi_pool.cu -- ROIPool forward / backward for sm_100a.
//
// Reference being replaced: wetectron/csrc/cuda/ROIPool_cuda.cu:16-108 (one thread per output
// scalar, stride-W scalar loads).  Arithmetic contract (bit-exact): SURVEY Appendix B.
//
// Forward, fast path (7x7 bins, C % 4 == 0):
//   1. nchw_to_nhwc: the map is transposed once into channels-last scratch ([B][H][W][C]); it is
//      19.9 MB per 608x1024 image and stays L2-resident (126 MB L2) for the pooling kernel.
//   2. roi_pool_fwd_nhwc7: one CTA per (roi, 128-channel slab); warp = bin row ph, lane = 4
//      consecutive channels (one 16-byte load per cell, 512 contiguous bytes per warp).  Every
//      lane runs the reference's scan (row-major, strict '>') for its channels, so the tie-break
//      is the reference's by construction.  Results are staged in shared memory in [c][49] order
//      and written out as contiguous 16-byte streaming stores (out and argmax are each one
//      contiguous 25 KB run per CTA).
// Generic path (other bin shapes / channel counts): one thread per output scalar.
//
// Backward: grad_in is zeroed, then one thread per output scalar issues a fire-and-forget
// red.global.add.f32 at argmax (same arithmetic as the reference, :100-105).
#include "common.cuh"

namespace {

// ----------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                    int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* src = in + (size_t)b * C * HW;
  float* dst = out + (size_t)b * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? __ldg(src + (size_t)c * HW + p) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) dst[(size_t)p * C + c] = tile[threadIdx.x][i];
  }
}

struct RoiGeom {
  int b, x1, y1;
  float bh, bw;
};

__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float scale, int PH, int PW) {
  RoiGeom g;
  g.b = (int)roi[0];
  g.x1 = (int)roundf(__fmul_rn(roi[1], scale));
  g.y1 = (int)roundf(__fmul_rn(roi[2], scale));
  const int x2 = (int)roundf(__fmul_rn(roi[3], scale));
  const int y2 = (int)roundf(__fmul_rn(roi[4], scale));
  const int rw = max(x2 - g.x1 + 1, 1);
  const int rh = max(y2 - g.y1 + 1, 1);
  g.bh = __fdiv_rn((float)rh, (float)PH);
  g.bw = __fdiv_rn((float)rw, (float)PW);
  return g;
}

__device__ __forceinline__ void bin_range(int p, float bsz, int start, int limit, int& lo, int& hi) {
  lo = (int)floorf(__fmul_rn((float)p, bsz));
  hi = (int)ceilf(__fmul_rn((float)(p + 1), bsz));
  lo = min(max(lo + start, 0), limit);
  hi = min(max(hi + start, 0), limit);
}

constexpr int kSlab = 128;            // channels per CTA in the fast path
constexpr int kBins = 49;

__global__ void __launch_bounds__(7 * 32, 4)
roi_pool_fwd_nhwc7_kernel(const float* __restrict__ feat_nhwc, const float* __restrict__ rois, int C,
                          int H, int W, float scale, float* __restrict__ out,
                          int32_t* __restrict__ argmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_val = reinterpret_cast<float*>(smem_raw);
  int* s_idx = reinterpret_cast<int*>(smem_raw + kSlab * kBins * sizeof(float));

  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kSlab;
  const int nch = min(kSlab, C - c0);
  const int ph = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, 7, 7);

  if (4 * lane < nch) {
    int hs, he;
    bin_range(ph, g.bh, g.y1, H, hs, he);
    const float* base = feat_nhwc + (size_t)g.b * H * W * C + c0 + 4 * lane;
#pragma unroll 1
    for (int pw = 0; pw < 7; ++pw) {
      int ws, we;
      bin_range(pw, g.bw, g.x1, W, ws, we);
      const bool empty = (he <= hs) || (we <= ws);
      float m0, m1, m2, m3;
      m0 = m1 = m2 = m3 = empty ? 0.f : -FLT_MAX;
      int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
      for (int h = hs; h < he; ++h) {
        const float* row = base + (size_t)h * W * C;
        int idx = h * W + ws;
#pragma unroll 4
        for (int w = ws; w < we; ++w, ++idx) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(row + (size_t)w * C));
          if (v.x > m0) { m0 = v.x; i0 = idx; }
          if (v.y > m1) { m1 = v.y; i1 = idx; }
          if (v.z > m2) { m2 = v.z; i2 = idx; }
          if (v.w > m3) { m3 = v.w; i3 = idx; }
        }
      }
      const int o = (4 * lane) * kBins + ph * 7 + pw;
      s_val[o] = m0;             s_idx[o] = i0;
      s_val[o + kBins] = m1;     s_idx[o + kBins] = i1;
      s_val[o + 2 * kBins] = m2; s_idx[o + 2 * kBins] = i2;
      s_val[o + 3 * kBins] = m3; s_idx[o + 3 * kBins] = i3;
    }
  }
  __syncthreads();
  // contiguous, 16-byte aligned run of nch*49 scalars ((n*C + c0) % 4 == 0, nch % 4 == 0)
  const size_t obase = ((size_t)n * C + c0) * kBins;
  const int nvec = nch * kBins / 4;
  float4* o4 = reinterpret_cast<float4*>(out + obase);
  int4* a4 = reinterpret_cast<int4*>(argmax + obase);
  const float4* sv4 = reinterpret_cast<const float4*>(s_val);
  const int4* si4 = reinterpret_cast<const int4*>(s_idx);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    __stcs(o4 + i, sv4[i]);
    __stcs(a4 + i, si4[i]);
  }
}

// one thread per output scalar, NCHW direct (any bin shape / channel count)
__global__ void roi_pool_fwd_generic_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                                            long long total, int C, int H, int W, float scale, int PH,
                                            int PW, float* __restrict__ out, int32_t* __restrict__ argmax) {
  for (long long index = blockIdx.x * (long long)blockDim.x + threadIdx.x; index < total;
       index += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(index % PW);
    const int ph = (int)((index / PW) % PH);
    const int c = (int)((index / PW / PH) % C);
    const int n = (int)(index / PW / PH / C);
    const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, PH, PW);
    int hs, he, ws, we;
    bin_range(ph, g.bh, g.y1, H, hs, he);
    bin_range(pw, g.bw, g.x1, W, ws, we);
    const bool empty = (he <= hs) || (we <= ws);
    float m = empty ? 0.f : -FLT_MAX;
    int mi = -1;
    const float* plane = feat + ((size_t)g.b * C + c) * H * W;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const float v = __ldg(plane + h * W + w);
        if (v > m) { m = v; mi = h * W + w; }
      }
    out[index] = m;
    argmax[index] = mi;
  }
}

__global__ void roi_pool_bwd_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ argmax,
                                    const float* __restrict__ rois, long long total, int C, int HW, int bins,
                                    float* __restrict__ grad_in, int nhwc) {
  for (long long index = blockIdx.x * (long long)blockDim.x + threadIdx.x; index < total;
       index += (long long)gridDim.x * blockDim.x) {
    const int a = __ldcs(argmax + index);
    if (a == -1) continue;
    const long long nc = index / bins;
    const int c = (int)(nc % C);
    const int n = (int)(nc / C);
    const int b = (int)__ldg(rois + (size_t)n * 5);
    float* dst = nhwc ? grad_in + ((size_t)b * HW + a) * C + c : grad_in + ((size_t)b * C + c) * HW + a;
    atomicAdd(dst, __ldcs(grad_out + index));
  }
}

}  // namespace

ODW_API size_t odwscl_roi_pool_fwd_ws_bytes(int B, int C, int H, int W, int R, int ph, int pw) {
  (void)R;
  if (ph == 7 && pw == 7 && C % 4 == 0) return odw_align((size_t)B * C * H * W * sizeof(float));
  return 0;
}

ODW_API int odwscl_roi_pool_fwd_f32(const float* feat, int B, int C, int H, int W, const float* rois, int R,
                                    float scale, int ph, int pw, float* out, int32_t* argmax, void* ws,
                                    size_t ws_bytes, odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || R < 0 || ph <= 0 || pw <= 0) return ODWSCL_EINVAL;
  if (R == 0 || C == 0) return 0;
  if (!feat || !rois || !out || !argmax) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const bool fast = (ph == 7 && pw == 7 && C % 4 == 0 && H > 0 && W > 0);
  if (fast) {
    if (!ws || ws_bytes < odwscl_roi_pool_fwd_ws_bytes(B, C, H, W, R, ph, pw)) return ODWSCL_ENOWS;
    float* nhwc = reinterpret_cast<float*>(ws);
    const int HW = H * W;
    dim3 tg(odw_cdiv(HW, 32), odw_cdiv(C, 32), B);
    nchw_to_nhwc_kernel<<<tg, dim3(32, 8), 0, st>>>(feat, nhwc, C, HW);
    ODW_LAUNCH_CHECK();
    const int smem = kSlab * kBins * (int)(sizeof(float) + sizeof(int));
    static bool attr_set = false;   // idempotent; benign if raced
    if (!attr_set) {
      ODW_CUDA(cudaFuncSetAttribute(roi_pool_fwd_nhwc7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    dim3 grid(R, odw_cdiv(C, kSlab));
    roi_pool_fwd_nhwc7_kernel<<<grid, 7 * 32, smem, st>>>(nhwc, rois, C, H, W, scale, out, argmax);
    ODW_LAUNCH_CHECK();
    return 0;
  }
  const long long total = (long long)R * C * ph * pw;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 16, (total + 255) / 256);
  roi_pool_fwd_generic_kernel<<<blocks, 256, 0, st>>>(feat, rois, total, C, H, W, scale, ph, pw, out, argmax);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_roi_pool_bwd_f32(const float* grad_out, const int32_t* argmax, const float* rois, int R,
                                    int B, int C, int H, int W, int ph, int pw, float* grad_in,
                                    odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || R < 0 || ph <= 0 || pw <= 0) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)B * C * H * W * sizeof(float);
  if (bytes == 0) return 0;
  if (!grad_in) return ODWSCL_EINVAL;
  ODW_CUDA(cudaMemsetAsync(grad_in, 0, bytes, st));
  if (R == 0) return 0;
  if (!grad_out || !argmax || !rois) return ODWSCL_EINVAL;
  const long long total = (long long)R * C * ph * pw;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 16, (total + 255) / 256);
  roi_pool_bwd_kernel<<<blocks, 256, 0, st>>>(grad_out, argmax, rois, total, C, H * W, ph * pw, grad_in, 0);
  ODW_LAUNCH_CHECK();
  return 0;
}

// Channels-last variants: the conv stack of this library produces / consumes NHWC maps, so the model path
// skips the layout transpose.  feat_nhwc / grad_in_nhwc are [B,H,W,C]; out / argmax / grad_out keep the
// reference layout [R,C,7,7] (argmax = h*W+w as always).  7x7 bins and C % 4 == 0 only.
ODW_API int odwscl_roi_pool_fwd_nhwc_f32(const float* feat_nhwc, int B, int C, int H, int W, const float* rois, int R,
                                         float scale, float* out, int32_t* argmax, odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H <= 0 || W <= 0 || R < 0 || (C & 3)) return ODWSCL_EINVAL;
  if (R == 0 || C == 0) return 0;
  if (!feat_nhwc || !rois || !out || !argmax) return ODWSCL_EINVAL;
  const int smem = kSlab * kBins * (int)(sizeof(float) + sizeof(int));
  ODW_CUDA(cudaFuncSetAttribute(roi_pool_fwd_nhwc7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(R, odw_cdiv(C, kSlab));
  roi_pool_fwd_nhwc7_kernel<<<grid, 7 * 32, smem, (cudaStream_t)stream>>>(feat_nhwc, rois, C, H, W, scale, out, argmax);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_roi_pool_bwd_nhwc_f32(const float* grad_out, const int32_t* argmax, const float* rois, int R, int B,
                                         int C, int H, int W, float* grad_in_nhwc, odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || R < 0) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)B * C * H * W * sizeof(float);
  if (bytes == 0) return 0;
  if (!grad_in_nhwc) return ODWSCL_EINVAL;
  ODW_CUDA(cudaMemsetAsync(grad_in_nhwc, 0, bytes, st));
  if (R == 0) return 0;
  if (!grad_out || !argmax || !rois) return ODWSCL_EINVAL;
  const long long total = (long long)R * C * 49;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 16, (total + 255) / 256);
  roi_pool_bwd_kernel<<<blocks, 256, 0, st>>>(grad_out, argmax, rois, total, C, H * W, 49, grad_in_nhwc, 1);
  ODW_LAUNCH_CHECK();
  return 0;
}
