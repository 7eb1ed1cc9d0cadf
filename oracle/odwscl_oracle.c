/*
 * odwscl_oracle.c -- CPU restatement of the OD-WSCL proposal-feature hot path's
 * integer / index / selection arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  The product path
 * (od-wscl_b200/) never links or calls it.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/wetectron).  Plain scalar C, one thread, fp32 arithmetic with
 * -ffp-contract=off so no step is fused that the reference keeps separate.
 *
 * Pinning (see tests/test_oracle_golden.py, oracle/gen_golden.py):
 *   - ROIPool fwd/bwd: the reference has no CPU ROIPool (csrc/ROIPool.h:23,44), so
 *     the restatement follows csrc/cuda/ROIPool_cuda.cu:16-108 line by line and is
 *     pinned bit-exactly against torch.ops.torchvision.roi_pool /
 *     _roi_pool_backward (same Caffe2 lineage) golden vectors.
 *   - ROIAlign fwd, legacy NMS: pinned against the reference's own CPU extension
 *     sources compiled unmodified (oracle/_ref, recipe oracle/build_ref.py).
 *   - box IoU (+1), torchvision NMS: pinned against the reference's Python
 *     (structures/boxlist_ops.py) imported in the build container and against
 *     torchvision.ops.nms (the reference's third-party NMS, README.md:32).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------- *
 * ROIPool forward: csrc/cuda/ROIPool_cuda.cu:16-77 (RoIPoolFForward<float>)
 * feat [B,C,H,W], rois [R,5] = (batch, x1, y1, x2, y2) in image pixels,
 * out [R,C,PH,PW] fp32, argmax [R,C,PH,PW] int32 plane index (h*W+w) or -1.
 * ------------------------------------------------------------------------- */
ORC_API void orc_roi_pool_fwd_f32(const float* feat, int B, int C, int H, int W,
                                  const float* rois, int R, float scale, int PH, int PW,
                                  float* out, int32_t* argmax) {
  (void)B;
  for (int n = 0; n < R; ++n) {
    const float* roi = rois + (size_t)n * 5;
    int b = (int)roi[0];                               /* :29 */
    int x1 = (int)roundf(roi[1] * scale);              /* :30 */
    int y1 = (int)roundf(roi[2] * scale);              /* :31 */
    int x2 = (int)roundf(roi[3] * scale);              /* :32 */
    int y2 = (int)roundf(roi[4] * scale);              /* :33 */
    int rw = imax(x2 - x1 + 1, 1);                     /* :36 */
    int rh = imax(y2 - y1 + 1, 1);                     /* :37 */
    float bh = (float)rh / (float)PH;                  /* :38-39 */
    float bw = (float)rw / (float)PW;                  /* :40-41 */
    for (int c = 0; c < C; ++c) {
      const float* plane = feat + ((size_t)b * C + c) * H * W;   /* :63-64 */
      for (int ph = 0; ph < PH; ++ph) {
        for (int pw = 0; pw < PW; ++pw) {
          int hs = (int)floorf((float)ph * bh);        /* :43-44 */
          int ws = (int)floorf((float)pw * bw);        /* :45-46 */
          int he = (int)ceilf((float)(ph + 1) * bh);   /* :47-48 */
          int we = (int)ceilf((float)(pw + 1) * bw);   /* :49-50 */
          hs = imin(imax(hs + y1, 0), H);              /* :53 */
          he = imin(imax(he + y1, 0), H);              /* :54 */
          ws = imin(imax(ws + x1, 0), W);              /* :55 */
          we = imin(imax(we + x1, 0), W);              /* :56 */
          int empty = (he <= hs) || (we <= ws);        /* :57 */
          float maxval = empty ? 0.f : -FLT_MAX;       /* :60 */
          int maxidx = -1;                             /* :62 */
          for (int h = hs; h < he; ++h)                /* :65-73 */
            for (int w = ws; w < we; ++w) {
              int idx = h * W + w;
              if (plane[idx] > maxval) { maxval = plane[idx]; maxidx = idx; }
            }
          size_t o = (((size_t)n * C + c) * PH + ph) * PW + pw;
          out[o] = maxval;                             /* :74 */
          argmax[o] = maxidx;                          /* :75 */
        }
      }
    }
  }
}

/* ROIPool backward: csrc/cuda/ROIPool_cuda.cu:79-108.  grad_in is zeroed here
 * (the reference allocates it with at::zeros, :172).  Sum order: ascending
 * output index (the GPU reference's atomic order is unspecified). */
ORC_API void orc_roi_pool_bwd_f32(const float* grad_out, const int32_t* argmax,
                                  const float* rois, int R, int B, int C, int H, int W,
                                  int PH, int PW, float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)B * C * H * W);
  for (int n = 0; n < R; ++n) {
    int b = (int)rois[(size_t)n * 5];                  /* :93 */
    for (int c = 0; c < C; ++c) {
      float* plane = grad_in + ((size_t)b * C + c) * H * W;      /* :94,97 */
      size_t top = ((size_t)n * C + c) * PH * PW;                /* :95 */
      for (int k = 0; k < PH * PW; ++k) {
        int a = argmax[top + k];                       /* :100 */
        if (a != -1) plane[a] += grad_out[top + k];    /* :101-105 */
      }
    }
  }
}

/* ------------------------------------------------------------------------- *
 * ROIAlign forward / backward (legacy, non-aligned):
 * csrc/cuda/ROIAlign_cuda.cu:15-122 (fwd), :125-254 (bwd); the reference's CPU
 * forward (csrc/cpu/ROIAlign_cpu.cpp) computes the same sums with pre-computed
 * weights.
 * ------------------------------------------------------------------------- */
static float bilinear(const float* p, int H, int W, float y, float x) {
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0.f;   /* :22-25 */
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  float v1 = p[yl * W + xl], v2 = p[yl * W + xh], v3 = p[yh * W + xl], v4 = p[yh * W + xh];
  float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;        /* :59 */
}

ORC_API void orc_roi_align_fwd_f32(const float* feat, int B, int C, int H, int W,
                                   const float* rois, int R, float scale, int PH, int PW,
                                   int sampling_ratio, float* out) {
  (void)B;
  for (int n = 0; n < R; ++n) {
    const float* roi = rois + (size_t)n * 5;
    int b = (int)roi[0];
    float sw = roi[1] * scale, sh = roi[2] * scale;    /* :81-84, no rounding */
    float ew = roi[3] * scale, eh = roi[4] * scale;
    float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);    /* :91-92 */
    float bh = rh / (float)PH, bw = rw / (float)PW;
    int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / PH);   /* :99 */
    int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / PW);   /* :100 */
    float count = (float)(gh * gw);
    for (int c = 0; c < C; ++c) {
      const float* plane = feat + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.f;
          for (int iy = 0; iy < gh; ++iy) {
            float y = sh + ph * bh + (iy + .5f) * bh / (float)gh;         /* :108 */
            for (int ix = 0; ix < gw; ++ix) {
              float x = sw + pw * bw + (ix + .5f) * bw / (float)gw;       /* :111 */
              acc += bilinear(plane, H, W, y, x);
            }
          }
          out[(((size_t)n * C + c) * PH + ph) * PW + pw] = acc / count;   /* :117-119 */
        }
    }
  }
}

ORC_API void orc_roi_align_bwd_f32(const float* grad_out, const float* rois, int R, float scale,
                                   int PH, int PW, int B, int C, int H, int W,
                                   int sampling_ratio, float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)B * C * H * W);
  for (int n = 0; n < R; ++n) {
    const float* roi = rois + (size_t)n * 5;
    int b = (int)roi[0];
    float sw = roi[1] * scale, sh = roi[2] * scale;
    float ew = roi[3] * scale, eh = roi[4] * scale;
    float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);
    float bh = rh / (float)PH, bw = rw / (float)PW;
    int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / PH);
    int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / PW);
    float count = (float)(gh * gw);
    for (int c = 0; c < C; ++c) {
      float* plane = grad_in + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float g = grad_out[(((size_t)n * C + c) * PH + ph) * PW + pw];
          for (int iy = 0; iy < gh; ++iy) {
            float y = sh + ph * bh + (iy + .5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ++ix) {
              float x = sw + pw * bw + (ix + .5f) * bw / (float)gw;
              /* bilinear_interpolate_gradient, :125-175 */
              if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
              float yy = y <= 0 ? 0 : y, xx = x <= 0 ? 0 : x;
              int yl = (int)yy, xl = (int)xx, yh, xh;
              if (yl >= H - 1) { yh = yl = H - 1; yy = (float)yl; } else yh = yl + 1;
              if (xl >= W - 1) { xh = xl = W - 1; xx = (float)xl; } else xh = xl + 1;
              float ly = yy - yl, lx = xx - xl, hy = 1.f - ly, hx = 1.f - lx;
              float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              plane[yl * W + xl] += g * w1 / count;    /* :239-249 */
              plane[yl * W + xh] += g * w2 / count;
              plane[yh * W + xl] += g * w3 / count;
              plane[yh * W + xh] += g * w4 / count;
            }
          }
        }
    }
  }
}

/* ------------------------------------------------------------------------- *
 * Box IoU, a [na,4] x b [nb,4] -> out [na,nb].
 * plus_one=1: structures/boxlist_ops.py:127-160 + structures/bounding_box.py:231-241
 *             (TO_REMOVE = 1 in both the intersection and the areas).
 * plus_one=0: torchvision box_iou / the NMS kernel's IoU (no +1).
 * Operation order is the reference's: inter / (area1 + area2 - inter), with
 * (area1 + area2) added first (boxlist_ops.py:159).
 * ------------------------------------------------------------------------- */
static inline float iou_pair(const float* a, const float* b, float one) {
  float ltx = fmaxf(a[0], b[0]), lty = fmaxf(a[1], b[1]);
  float rbx = fminf(a[2], b[2]), rby = fminf(a[3], b[3]);
  float w = rbx - ltx + one, h = rby - lty + one;
  if (w < 0.f) w = 0.f;
  if (h < 0.f) h = 0.f;
  float inter = w * h;
  float aa = (a[2] - a[0] + one) * (a[3] - a[1] + one);
  float ab = (b[2] - b[0] + one) * (b[3] - b[1] + one);
  return inter / (aa + ab - inter);
}

ORC_API void orc_box_iou_f32(const float* a, int na, const float* b, int nb, int plus_one,
                             float* out) {
  float one = plus_one ? 1.f : 0.f;
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = iou_pair(a + 4 * i, b + 4 * j, one);
}

/* ------------------------------------------------------------------------- *
 * NMS with torchvision.ops.nms semantics (the call at
 * structures/boxlist_ops.py:57; torchvision 0.8.2 pinned by README.md:32):
 * stable descending sort by score, IoU without +1, suppress iff IoU > thr,
 * kept indices returned in descending-score order.
 * ------------------------------------------------------------------------- */
typedef struct { float s; int i; } orc_si;
static int cmp_desc_stable(const void* pa, const void* pb) {
  const orc_si* a = (const orc_si*)pa; const orc_si* b = (const orc_si*)pb;
  if (a->s > b->s) return -1;
  if (a->s < b->s) return 1;
  return a->i - b->i;
}

ORC_API int orc_nms_tv_f32(const float* boxes, const float* scores, int n, float thr,
                           int64_t* keep) {
  if (n <= 0) return 0;
  orc_si* ord = (orc_si*)malloc(sizeof(orc_si) * n);
  uint8_t* sup = (uint8_t*)calloc(n, 1);
  for (int i = 0; i < n; ++i) { ord[i].s = scores[i]; ord[i].i = i; }
  qsort(ord, n, sizeof(orc_si), cmp_desc_stable);
  int nk = 0;
  for (int oi = 0; oi < n; ++oi) {
    int i = ord[oi].i;
    if (sup[i]) continue;
    keep[nk++] = i;
    for (int oj = oi + 1; oj < n; ++oj) {
      int j = ord[oj].i;
      if (sup[j]) continue;
      if (iou_pair(boxes + 4 * i, boxes + 4 * j, 0.f) > thr) sup[j] = 1;
    }
  }
  free(ord); free(sup);
  return nk;
}

/* Legacy `_C.nms` (layers/nms.py:6): +1 pixel convention, kept indices returned
 * sorted ASCENDING.  ge=1 reproduces csrc/cpu/nms_cpu.cpp:60 (suppress iff
 * ovr >= thr); ge=0 reproduces csrc/cuda/nms.cu:60 (suppress iff IoU > thr). */
ORC_API int orc_nms_legacy_f32(const float* boxes, const float* scores, int n, float thr,
                               int ge, int64_t* keep) {
  if (n <= 0) return 0;
  orc_si* ord = (orc_si*)malloc(sizeof(orc_si) * n);
  uint8_t* sup = (uint8_t*)calloc(n, 1);
  for (int i = 0; i < n; ++i) { ord[i].s = scores[i]; ord[i].i = i; }
  qsort(ord, n, sizeof(orc_si), cmp_desc_stable);
  for (int oi = 0; oi < n; ++oi) {
    int i = ord[oi].i;
    if (sup[i]) continue;
    for (int oj = oi + 1; oj < n; ++oj) {
      int j = ord[oj].i;
      if (sup[j]) continue;
      float v = iou_pair(boxes + 4 * i, boxes + 4 * j, 1.f);
      if (ge ? (v >= thr) : (v > thr)) sup[j] = 1;
    }
  }
  int nk = 0;
  for (int i = 0; i < n; ++i) if (!sup[i]) keep[nk++] = i;     /* nms_cpu.cpp:64 / nms.cu:127-130 */
  free(ord); free(sup);
  return nk;
}
