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
#include <stdlib.h>
#include <string.h>

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

#define ODW_RP_CMP4(v, id)                  \
  if (v.x > m0) { m0 = v.x; i0 = (id); }    \
  if (v.y > m1) { m1 = v.y; i1 = (id); }    \
  if (v.z > m2) { m2 = v.z; i2 = (id); }    \
  if (v.w > m3) { m3 = v.w; i3 = (id); }

#define ODW_RP_CMP4N(v, id)                 \
  if (v.x > n0) { n0 = v.x; j0 = (id); }    \
  if (v.y > n1) { n1 = v.y; j1 = (id); }    \
  if (v.z > n2) { n2 = v.z; j2 = (id); }    \
  if (v.w > n3) { n3 = v.w; j3 = (id); }

// CQT = C/4 as a compile-time constant (cell stride becomes an immediate load offset), 0 = runtime C.
template <int CQT, bool kPair>
__global__ void __launch_bounds__(7 * 32, 4)
roi_pool_fwd_nhwc7_kernel(const float4* __restrict__ feat4, const float* __restrict__ rois, int C,
                          int H, int W, float scale, float* __restrict__ out,
                          int32_t* __restrict__ argmax, const float* __restrict__ aug_mask,
                          float* __restrict__ out_aug) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_val = reinterpret_cast<float*>(smem_raw);
  int* s_idx = reinterpret_cast<int*>(smem_raw + kSlab * kBins * sizeof(float));

  const int CQ = CQT ? CQT : (C >> 2);
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kSlab;
  const int nch = min(kSlab, C - c0);
  const int ph = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, 7, 7);

  if (4 * lane < nch) {
    int hs, he;
    bin_range(ph, g.bh, g.y1, H, hs, he);
    const float4* base = feat4 + (size_t)g.b * H * W * CQ + (c0 >> 2) + lane;
#pragma unroll 1
    for (int pw = 0; pw < 7; ++pw) {
      int ws, we;
      bin_range(pw, g.bw, g.x1, W, ws, we);
      const bool empty = (he <= hs) || (we <= ws);
      float m0, m1, m2, m3;
      m0 = m1 = m2 = m3 = empty ? 0.f : -FLT_MAX;
      int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
      const int nw = we - ws;
      int h = hs;
      if (kPair) {
        // Two rows in flight: 4 outstanding 16-byte loads per lane instead of 2.  The kernel is bound by L2 latency x loads in
        // flight (a bin row is only ~5 cells).  Row h+1 scans into its own (n, j) and is merged with the same
        // strict '>' afterwards, so the winner is still the first cell in row-major order that attains the maximum.
#pragma unroll 1
        for (; h + 1 < he; h += 2) {
          int idx = h * W + ws;
          const float4* p = base + (size_t)idx * CQ;
          const float4* p2 = p + (size_t)W * CQ;
          float n0, n1, n2, n3;
          n0 = n1 = n2 = n3 = -FLT_MAX;
          int j0 = -1, j1 = -1, j2 = -1, j3 = -1;
          int k = nw;
#pragma unroll 1
          for (; k >= 2; k -= 2) {
            const float4 a = __ldg(p), b = __ldg(p + CQ), c = __ldg(p2), d = __ldg(p2 + CQ);
            p += 2 * CQ; p2 += 2 * CQ;
            ODW_RP_CMP4(a, idx);
            ODW_RP_CMP4(b, idx + 1);
            ODW_RP_CMP4N(c, idx + W);
            ODW_RP_CMP4N(d, idx + W + 1);
            idx += 2;
          }
          if (k > 0) {
            const float4 a = __ldg(p), c = __ldg(p2);
            ODW_RP_CMP4(a, idx);
            ODW_RP_CMP4N(c, idx + W);
          }
          if (n0 > m0) { m0 = n0; i0 = j0; }
          if (n1 > m1) { m1 = n1; i1 = j1; }
          if (n2 > m2) { m2 = n2; i2 = j2; }
          if (n3 > m3) { m3 = n3; i3 = j3; }
        }
      }
#pragma unroll 1
      for (; h < he; ++h) {                    // reference scan order: rows, then columns, strict '>'
        int idx = h * W + ws;
        const float4* p = base + (size_t)idx * CQ;
        int k = nw;
#pragma unroll 1
        for (; k >= 2; k -= 2) {
          const float4 a = __ldg(p), b = __ldg(p + CQ);
          p += 2 * CQ;
          ODW_RP_CMP4(a, idx);
          ODW_RP_CMP4(b, idx + 1);
          idx += 2;
        }
        if (k > 0) {
          const float4 a = __ldg(p);
          ODW_RP_CMP4(a, idx);
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
  // second output: the DropBlock-augmented copy of the pooled features (weak_head.py:111: x * block_mask * scale),
  // written from the same staged values -- the separate apply pass would re-read the 200 MB this kernel just wrote
  if (out_aug != nullptr) {
    const float* mk = aug_mask + (size_t)n * kBins;            // per-(roi, bin) factor, scale included
    float4* g4 = reinterpret_cast<float4*>(out_aug + obase);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      float4 v = sv4[i];
      const int k = (4 * i) % kBins;
      v.x *= __ldg(mk + k);
      v.y *= __ldg(mk + (k + 1 < kBins ? k + 1 : k + 1 - kBins));
      v.z *= __ldg(mk + (k + 2 < kBins ? k + 2 : k + 2 - kBins));
      v.w *= __ldg(mk + (k + 3 < kBins ? k + 3 : k + 3 - kBins));
      __stcs(g4 + i, v);
    }
  }
}


static bool rp_pairs() {                                     // ODWSCL_ROIPOOL=rows: the one-row-at-a-time scan (A/B aid)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ODWSCL_ROIPOOL");
    v = (e && !strcmp(e, "rows")) ? 0 : 1;
  }
  return v == 1;
}

template <typename K>
static int launch_rp(K kern, dim3 grid, int smem, cudaStream_t st, const float4* f4, const float* rois, int C, int H, int W,
                     float scale, float* out, int32_t* argmax, const float* aug_mask, float* out_aug) {
  ODW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, 7 * 32, smem, st>>>(f4, rois, C, H, W, scale, out, argmax, aug_mask, out_aug);
  ODW_LAUNCH_CHECK();
  return 0;
}

static int launch_fwd_nhwc7(const float* nhwc, const float* rois, int C, int H, int W, int R, float scale,
                            float* out, int32_t* argmax, cudaStream_t st, const float* aug_mask = nullptr,
                            float* out_aug = nullptr) {
  const int smem = kSlab * kBins * (int)(sizeof(float) + sizeof(int));
  dim3 grid(R, odw_cdiv(C, kSlab));
  const float4* f4 = reinterpret_cast<const float4*>(nhwc);
#define ODW_RP_GO(CQT_, P_) \
  return launch_rp(roi_pool_fwd_nhwc7_kernel<CQT_, P_>, grid, smem, st, f4, rois, C, H, W, scale, out, argmax, aug_mask, out_aug)
  if (C == 512) {
    if (rp_pairs()) ODW_RP_GO(128, true);
    ODW_RP_GO(128, false);
  }
  if (rp_pairs()) ODW_RP_GO(0, true);
  ODW_RP_GO(0, false);
#undef ODW_RP_GO
}

// Measurement aid: the FLOOR of any "stage the roi's region once per (roi, 128-channel slab)" design -- every
// cell of the roi's clamped extent is loaded exactly once per slab with the same 16-byte-per-lane pattern as the pooling
// kernel and reduced with a single FMNMX (no bins, no argmax, no 50 KB of output).  Its time is what the L2 -> SM fabric
// charges for the bytes alone.
__global__ void __launch_bounds__(7 * 32, 4)
roi_stream_probe_kernel(const float4* __restrict__ feat4, const float* __restrict__ rois, int C, int H, int W, float scale,
                        float* __restrict__ out) {
  const int CQ = C >> 2;
  const int n = blockIdx.x, c0 = blockIdx.y * kSlab;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, 7, 7);
  int hs, he, ws, we, t0, t1;
  bin_range(0, g.bh, g.y1, H, hs, t0);
  bin_range(6, g.bh, g.y1, H, t1, he);
  bin_range(0, g.bw, g.x1, W, ws, t0);
  bin_range(6, g.bw, g.x1, W, t1, we);
  float m = -FLT_MAX;
  if (4 * lane < min(kSlab, C - c0)) {
    const float4* base = feat4 + (size_t)g.b * H * W * CQ + (c0 >> 2) + lane;
    for (int h = hs + wid; h < he; h += 7) {                 // rows round-robin over the 7 warps
      const float4* p = base + (size_t)(h * W + ws) * CQ;
#pragma unroll 4
      for (int w = ws; w < we; ++w, p += CQ) {
        const float4 a = __ldg(p);
        m = fmaxf(m, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
      }
    }
  }
  m = odw_warp_max(m);
  if (lane == 0) out[((size_t)n * gridDim.y + blockIdx.y) * 7 + wid] = m;
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

// ----------------------------------------------------------------------------------------
// Backward, plane-centric: one CTA owns NCH channels of one image for a chunk of rois and accumulates the
// scatter-add in SHARED memory (the whole H x W plane of its channels: 76*128*4ch*4 B = 152 KB), so the 49*R*C
// scalar adds never reach L2 as atomics; grad_out / argmax are read as contiguous NCH*49-float runs per roi
// (16-byte vector loads, streaming).  Each CTA then flushes its plane once with vector reds (NHWC:
// red.global.add.v4.f32, one per cell) into the zeroed grad map.  Sum order differs from the reference's
// atomicAdd order exactly as the reference's own runs differ from each other (:100-105).
constexpr int kBwdThreads = 1024;
constexpr int kBwdChunkRois = 512;

template <int NCH, bool NHWC>
__global__ void __launch_bounds__(kBwdThreads, 1)
roi_pool_bwd_plane_kernel(const float* __restrict__ grad_out, const float* __restrict__ grad_out2,
                          const float* __restrict__ mask2, const int64_t* __restrict__ srows, const float* __restrict__ sgrad, int S,
                          const int32_t* __restrict__ argmax, const float* __restrict__ rois, int R, int C, int HW,
                          float* __restrict__ grad_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* acc = reinterpret_cast<float*>(smem_raw);                         // [HW][NCH]
  __shared__ int s_list[kBwdChunkRois];
  __shared__ int s_n;
  const int cg = blockIdx.x, b = blockIdx.y;
  const int r0 = blockIdx.z * kBwdChunkRois, r1 = min(R, r0 + kBwdChunkRois);
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x)
    if ((int)__ldg(rois + (size_t)r * 5) == b) s_list[atomicAdd(&s_n, 1)] = r;
  __syncthreads();
  const int nroi = s_n;
  if (nroi == 0) return;                                                   // chunk has no roi of this image
  {
    float4* a4 = reinterpret_cast<float4*>(acc);
    const int n4 = HW * NCH / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = n4 * 4 + threadIdx.x; i < HW * NCH; i += blockDim.x) acc[i] = 0.f;
  }
  __syncthreads();
  constexpr int kRun = NCH * 49;                                           // scalars per (roi, channel group)
  if constexpr (NCH == 4) {
    constexpr int kVec = kRun / 4;                                         // 49 float4 per roi
    const int items = nroi * kVec;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int rl = it / kVec, q = it - rl * kVec;
      const size_t off = ((size_t)s_list[rl] * C + (size_t)cg * NCH) * 49 + 4 * q;
      float4 g = __ldcs(reinterpret_cast<const float4*>(grad_out + off));
      if (grad_out2 != nullptr) {                  // second consumer of the pooled features: summed on the fly
        float4 g2 = __ldcs(reinterpret_cast<const float4*>(grad_out2 + off));
        if (mask2 != nullptr) {                    // ... through a per-(roi, bin) multiplier (the DropBlock backward)
          const float* mk = mask2 + (size_t)s_list[rl] * 49;
          const int k0 = (4 * q) % 49;
          g2.x *= __ldg(mk + k0);
          g2.y *= __ldg(mk + (k0 + 1 < 49 ? k0 + 1 : k0 - 48));
          g2.z *= __ldg(mk + (k0 + 2 < 49 ? k0 + 2 : k0 - 47));
          g2.w *= __ldg(mk + (k0 + 3 < 49 ? k0 + 3 : k0 - 46));
        }
        g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
      }
      const int4 a = __ldcs(reinterpret_cast<const int4*>(argmax + off));
      const int e = 4 * q;
      if (a.x >= 0) atomicAdd(acc + a.x * NCH + (e) / 49, g.x);
      if (a.y >= 0) atomicAdd(acc + a.y * NCH + (e + 1) / 49, g.y);
      if (a.z >= 0) atomicAdd(acc + a.z * NCH + (e + 2) / 49, g.z);
      if (a.w >= 0) atomicAdd(acc + a.w * NCH + (e + 3) / 49, g.w);
    }
    // sparse consumer: gradients of a few gathered rows (the contrastive branch's augmented positives)
    const int sitems = S * kVec;
    for (int it = threadIdx.x; it < sitems; it += blockDim.x) {
      const int k = it / kVec, q = it - k * kVec;
      const long long r = srows[k];
      if (r < r0 || r >= r1 || (int)__ldg(rois + (size_t)r * 5) != b) continue;
      const float4 g = __ldg(reinterpret_cast<const float4*>(sgrad + ((size_t)k * C + (size_t)cg * NCH) * 49 + 4 * q));
      const int4 a = __ldg(reinterpret_cast<const int4*>(argmax + ((size_t)r * C + (size_t)cg * NCH) * 49 + 4 * q));
      const int e = 4 * q;
      if (a.x >= 0) atomicAdd(acc + a.x * NCH + (e) / 49, g.x);
      if (a.y >= 0) atomicAdd(acc + a.y * NCH + (e + 1) / 49, g.y);
      if (a.z >= 0) atomicAdd(acc + a.z * NCH + (e + 2) / 49, g.z);
      if (a.w >= 0) atomicAdd(acc + a.w * NCH + (e + 3) / 49, g.w);
    }
  } else {
    const int items = nroi * kRun;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int rl = it / kRun, e = it - rl * kRun;
      const size_t off = ((size_t)s_list[rl] * C + (size_t)cg * NCH) * 49 + e;
      const int a = __ldcs(argmax + off);
      if (a >= 0) {
        float g2 = grad_out2 ? __ldcs(grad_out2 + off) : 0.f;
        if (mask2 != nullptr) g2 *= __ldg(mask2 + (size_t)s_list[rl] * 49 + e % 49);
        atomicAdd(acc + a * NCH + e / 49, __ldcs(grad_out + off) + g2);
      }
    }
    for (int it = threadIdx.x; it < S * kRun; it += blockDim.x) {
      const int k = it / kRun, e = it - k * kRun;
      const long long r = srows[k];
      if (r < r0 || r >= r1 || (int)__ldg(rois + (size_t)r * 5) != b) continue;
      const int a = __ldg(argmax + ((size_t)r * C + (size_t)cg * NCH) * 49 + e);
      if (a >= 0) atomicAdd(acc + a * NCH + e / 49, __ldg(sgrad + ((size_t)k * C + (size_t)cg * NCH) * 49 + e));
    }
  }
  __syncthreads();
  if constexpr (NHWC && NCH == 4) {
    float* dst = grad_in + (size_t)b * HW * C + (size_t)cg * NCH;
    for (int cell = threadIdx.x; cell < HW; cell += blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(acc)[cell];
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + (size_t)cell * C), "f"(v.x), "f"(v.y),
                     "f"(v.z), "f"(v.w)
                     : "memory");
    }
  } else {
    for (int i = threadIdx.x; i < HW * NCH; i += blockDim.x) {
      const int ch = i / HW, cell = i - ch * HW;                           // consecutive threads -> consecutive cells
      const float v = acc[cell * NCH + ch];
      if (v != 0.f) {
        float* dst = NHWC ? grad_in + ((size_t)b * HW + cell) * C + cg * NCH + ch
                          : grad_in + ((size_t)b * C + cg * NCH + ch) * HW + cell;
        atomicAdd(dst, v);
      }
    }
  }
}

struct BwdExtra {                 // optional extra consumers of the pooled features (all device pointers)
  const float* grad_out2;
  const float* mask2;             // [R,49] multiplier applied to grad_out2 (block mask * scale of the DropBlock), or null
  const int64_t* srows;
  const float* sgrad;
  int S;
};

template <int NCH, bool NHWC>
static int launch_bwd_plane(const float* grad_out, const int32_t* argmax, const float* rois, int R, int B, int C,
                            int HW, float* grad_in, cudaStream_t st, BwdExtra ex) {
  const int smem = HW * NCH * (int)sizeof(float);
  ODW_CUDA(cudaFuncSetAttribute(roi_pool_bwd_plane_kernel<NCH, NHWC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(C / NCH, B, odw_cdiv(R, kBwdChunkRois));
  roi_pool_bwd_plane_kernel<NCH, NHWC><<<grid, kBwdThreads, smem, st>>>(grad_out, ex.grad_out2, ex.mask2, ex.srows, ex.sgrad, ex.S,
                                                                         argmax, rois, R, C, HW, grad_in);
  ODW_LAUNCH_CHECK();
  return 0;
}

// picks the widest channel group whose plane fits in shared memory; false = no plane-centric path (7x7 only)
template <bool NHWC>
static int bwd_plane_dispatch(const float* grad_out, const int32_t* argmax, const float* rois, int R, int B, int C,
                              int HW, float* grad_in, cudaStream_t st, bool* done, BwdExtra ex = BwdExtra{nullptr, nullptr, nullptr, nullptr, 0}) {
  const size_t kMaxSmem = 220 * 1024;
  *done = true;
  if (C % 4 == 0 && (size_t)HW * 16 <= kMaxSmem) return launch_bwd_plane<4, NHWC>(grad_out, argmax, rois, R, B, C, HW, grad_in, st, ex);
  if (C % 2 == 0 && (size_t)HW * 8 <= kMaxSmem) return launch_bwd_plane<2, NHWC>(grad_out, argmax, rois, R, B, C, HW, grad_in, st, ex);
  if ((size_t)HW * 4 <= kMaxSmem) return launch_bwd_plane<1, NHWC>(grad_out, argmax, rois, R, B, C, HW, grad_in, st, ex);
  *done = false;
  return 0;
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
    return launch_fwd_nhwc7(nhwc, rois, C, H, W, R, scale, out, argmax, st);
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
  if (ph == 7 && pw == 7) {
    bool done = false;
    const int rc = bwd_plane_dispatch<false>(grad_out, argmax, rois, R, B, C, H * W, grad_in, st, &done);
    if (rc != 0 || done) return rc;
  }
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
  return launch_fwd_nhwc7(feat_nhwc, rois, C, H, W, R, scale, out, argmax, (cudaStream_t)stream);
}

ODW_API int odwscl_roi_pool_fwd_nhwc_aug_f32(const float* feat_nhwc, int B, int C, int H, int W, const float* rois, int R,
                                             float scale, float* out, int32_t* argmax, const float* aug_mask,
                                             float* out_aug, odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H <= 0 || W <= 0 || R < 0 || (C & 3)) return ODWSCL_EINVAL;
  if (R == 0 || C == 0) return 0;
  if (!feat_nhwc || !rois || !out || !argmax || !aug_mask || !out_aug) return ODWSCL_EINVAL;
  return launch_fwd_nhwc7(feat_nhwc, rois, C, H, W, R, scale, out, argmax, (cudaStream_t)stream, aug_mask, out_aug);
}

ODW_API int odwscl_probe_roi_stream_f32(const float* feat_nhwc, int B, int C, int H, int W, const float* rois, int R,
                                        float scale, float* out, odwscl_stream_t stream) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0 || R < 0 || (C & 3)) return ODWSCL_EINVAL;
  if (R == 0) return 0;
  if (!feat_nhwc || !rois || !out) return ODWSCL_EINVAL;
  dim3 grid(R, odw_cdiv(C, kSlab));
  const char* e = getenv("ODWSCL_PROBE_SMEM");               // carve shared memory out of L1 like the pooling kernel does
  const int smem = e ? atoi(e) : 0;
  if (smem > 0) ODW_CUDA(cudaFuncSetAttribute(roi_stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  roi_stream_probe_kernel<<<grid, 7 * 32, smem, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(feat_nhwc), rois, C, H, W,
                                                                    scale, out);
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
  {
    bool done = false;
    const int rc = bwd_plane_dispatch<true>(grad_out, argmax, rois, R, B, C, H * W, grad_in_nhwc, st, &done);
    if (rc != 0 || done) return rc;
  }
  const long long total = (long long)R * C * 49;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 16, (total + 255) / 256);
  roi_pool_bwd_kernel<<<blocks, 256, 0, st>>>(grad_out, argmax, rois, total, C, H * W, 49, grad_in_nhwc, 1);
  ODW_LAUNCH_CHECK();
  return 0;
}

// Backward of a pooled-feature tensor with several consumers (weak_head.py:107-120: fc6 on the clean features, DropBlock
// -> fc6 on the augmented ones, and the contrastive branch's gathered rows): the gradients are summed while they are
// scattered, instead of being materialised as one more [R,C,7,7] tensor per consumer.  grad_out2, and the sparse
// (srows [S] int64, sgrad [S,C,7,7]) source, may be null / empty.  Returns ODWSCL_EINVAL when the map does not fit the
// plane-centric kernel (the caller then sums the gradients itself and calls odwscl_roi_pool_bwd_nhwc_f32).
ODW_API int odwscl_roi_pool_bwd_nhwc_multi_f32(const float* grad_out, const float* grad_out2, const float* mask2,
                                               const int64_t* srows, const float* sgrad, int S, const int32_t* argmax, const float* rois, int R,
                                               int B, int C, int H, int W, float* grad_in_nhwc, odwscl_stream_t stream) {
  if (B < 0 || C < 0 || H < 0 || W < 0 || R < 0 || S < 0) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)B * C * H * W * sizeof(float);
  if (bytes == 0) return 0;
  if (!grad_in_nhwc) return ODWSCL_EINVAL;
  if ((size_t)H * W * 4 > 220 * 1024) return ODWSCL_EINVAL;
  ODW_CUDA(cudaMemsetAsync(grad_in_nhwc, 0, bytes, st));
  if (R == 0) return 0;
  if (!grad_out || !argmax || !rois || (S > 0 && (!srows || !sgrad))) return ODWSCL_EINVAL;
  bool done = false;
  return bwd_plane_dispatch<true>(grad_out, argmax, rois, R, B, C, H * W, grad_in_nhwc, st, &done,
                                  BwdExtra{grad_out2, mask2, srows, sgrad, S});
}
