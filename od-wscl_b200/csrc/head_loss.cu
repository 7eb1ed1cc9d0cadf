// head_loss.cu -- the MIL + refinement side of RoIRegLossComputation (roi_heads/weak_head/loss.py:233-259,349-406)
// as five launches instead of ~280 eager torch kernels (softmaxes, products, clamps, BCE, 3x weighted CE, 3x
// smooth-L1 gathers, top-k accuracies -- and the whole autograd backward of that graph, which is closed-form here).
//
// The eight predictor heads leave ONE [R, ld] logits buffer (modeling/predictors.py), column blocks
//   [cls C][det C][ref1 C][bbox1 Q][ref2 C][bbox2 Q][ref3 C][bbox3 Q]        Q = 4*C, or 8 when class-agnostic
// Stage 1 (before object discovery, which consumes the scores):
//   head_colstats   (two stages) per (image, class): max / sum-exp of the detection logits over the image's proposals (the
//                   dim-0 softmax of loss.py:237-244) and the column sums of the three refinement logits (accuracy)
//   head_scores     per proposal: softmax_c(cls) * softmax_rois(det) = final_score (loss.py:234-246) and the class
//                   softmaxes of ref1 / ref2 (the supervisors of branches 1 and 2, loss.py:283,313)
//   seg_colsum      img_score[b,c] = sum_j final_score[j,c]  (loss.py:352)
// Stage 2 (after od_layer produced pseudo labels / weights / regression targets):
//   head_loss       per proposal: weighted CE of the three refinement branches (lmda 3,1,1, loss.py:373-377),
//                   smooth-L1(beta=1) on the predicted class's box (loss.py:380-394), and -- closed form -- the
//                   gradient of ALL seven losses w.r.t. every logit, written straight into a [R, ld] buffer
//   head_finalize   loss_img = BCE(clamp(img_score)) (loss.py:353-354), fixed-order sums of the per-block partials,
//                   division by the number of images (loss.py:403-406), and the four top-k accuracies (loss.py:25-34)
// The backward of the whole block is head_grad_scale: the stored gradient times the seven upstream scalars.
#include "common.cuh"

namespace {

constexpr int kMaxC = 96;        // classes incl. background (3 per lane); VOC 21, COCO 81

struct HeadLayout {
  int C, Q, ld;
  __host__ __device__ int cls() const { return 0; }
  __host__ __device__ int det() const { return C; }
  __host__ __device__ int ref(int i) const { return 2 * C + i * (C + Q); }
  __host__ __device__ int bb(int i) const { return 3 * C + i * (C + Q); }
  __host__ __device__ int width() const { return 5 * C + 3 * Q; }
};

__device__ __forceinline__ int image_of_row(const int32_t* __restrict__ img_off, int B, int j) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (img_off[mid] <= j) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Column statistics over the rows of every image, two deterministic stages (a first version with one CTA per
// (image, 32 classes) took 234 us on 2 x 2000 rows: two CTAs walking 2000 rows each are pure latency).
//   stage 1  grid (B, ceil(C/32), kColSplit): CTA z takes every kColSplit-th block of 8 rows; lane = class, warp = row
//            within the block; per CTA: online (max, sum-exp) of the detection logits + sums of the three refinement
//            logits -> partial[b][z][5][C]
//   stage 2  grid (B, ceil(C/32)): the kColSplit partials combined in a fixed order
constexpr int kColSplit = 32;

__global__ void __launch_bounds__(256)
head_colstats_partial_kernel(const float* __restrict__ logits, HeadLayout L, const int32_t* __restrict__ img_off,
                             float* __restrict__ partial) {
  const int b = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5, z = blockIdx.z;
  const int r0 = img_off[b], r1 = img_off[b + 1];
  float m = -INFINITY, s = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (c < L.C)
    for (int j = r0 + z * 8 + w; j < r1; j += 8 * kColSplit) {
      const float* row = logits + (size_t)j * L.ld;
      const float x = row[L.det() + c];
      if (x > m) { s = s * expf(m - x) + 1.f; m = x; } else { s += expf(x - m); }
      a0 += row[L.ref(0) + c]; a1 += row[L.ref(1) + c]; a2 += row[L.ref(2) + c];
    }
  __shared__ float sm[5][8][33];
  const int l = threadIdx.x & 31;
  sm[0][w][l] = m; sm[1][w][l] = s; sm[2][w][l] = a0; sm[3][w][l] = a1; sm[4][w][l] = a2;
  __syncthreads();
  if (w == 0 && c < L.C) {
    float M = -INFINITY;
    for (int k = 0; k < 8; ++k) M = fmaxf(M, sm[0][k][l]);
    float S = 0.f, A0 = 0.f, A1 = 0.f, A2 = 0.f;
    for (int k = 0; k < 8; ++k) {
      if (sm[1][k][l] > 0.f) S += sm[1][k][l] * expf(sm[0][k][l] - M);
      A0 += sm[2][k][l]; A1 += sm[3][k][l]; A2 += sm[4][k][l];
    }
    float* p = partial + ((size_t)(b * kColSplit + z) * 5) * L.C + c;
    p[0] = M; p[L.C] = S; p[2 * L.C] = A0; p[3 * L.C] = A1; p[4 * L.C] = A2;
  }
}

__global__ void __launch_bounds__(32)
head_colstats_final_kernel(const float* __restrict__ partial, int C, int B, float* __restrict__ det_max,
                           float* __restrict__ det_sum, float* __restrict__ ref_colsum) {
  const int b = blockIdx.x, c = blockIdx.y * 32 + threadIdx.x;
  if (c >= C) return;
  const float* p = partial + (size_t)b * kColSplit * 5 * C + c;
  float M = -INFINITY;
  for (int z = 0; z < kColSplit; ++z) M = fmaxf(M, p[(size_t)z * 5 * C]);
  float S = 0.f, A0 = 0.f, A1 = 0.f, A2 = 0.f;
  for (int z = 0; z < kColSplit; ++z) {
    const float* q = p + (size_t)z * 5 * C;
    if (q[C] > 0.f) S += q[C] * expf(q[0] - M);
    A0 += q[2 * C]; A1 += q[3 * C]; A2 += q[4 * C];
  }
  det_max[b * C + c] = M;
  det_sum[b * C + c] = S;
  ref_colsum[(0 * B + b) * C + c] = A0;
  ref_colsum[(1 * B + b) * C + c] = A1;
  ref_colsum[(2 * B + b) * C + c] = A2;
}

// partial[b][z][c] = sum over CTA z's rows of x[j,c]; stage 2 adds the kColSplit partials in order
__global__ void __launch_bounds__(256)
seg_colsum_partial_kernel(const float* __restrict__ x, int ld, int C, const int32_t* __restrict__ img_off,
                          float* __restrict__ partial) {
  const int b = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5, z = blockIdx.z;
  const int r0 = img_off[b], r1 = img_off[b + 1];
  float a = 0.f;
  if (c < C)
    for (int j = r0 + z * 8 + w; j < r1; j += 8 * kColSplit) a += x[(size_t)j * ld + c];
  __shared__ float sm[8][33];
  sm[w][threadIdx.x & 31] = a;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x & 31];
    partial[(size_t)(b * kColSplit + z) * C + c] = t;
  }
}

__global__ void __launch_bounds__(32)
seg_colsum_final_kernel(const float* __restrict__ partial, int C, float* __restrict__ out) {
  const int b = blockIdx.x, c = blockIdx.y * 32 + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int z = 0; z < kColSplit; ++z) t += partial[(size_t)(b * kColSplit + z) * C + c];
  out[b * C + c] = t;
}

// row softmax over C classes held 3 per lane (class = lane + 32 k); returns max and sum, v[] <- exp(x - max)
__device__ __forceinline__ void warp_softmax3(float (&v)[3], int lane, int C, float& m, float& s) {
  m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 3; ++k) if (lane + 32 * k < C) m = fmaxf(m, v[k]);
  m = odw_warp_max(m);
  s = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v[k] = (lane + 32 * k < C) ? expf(v[k] - m) : 0.f;
    s += v[k];
  }
  s = odw_warp_sum(s);
}

// warp per proposal
__global__ void __launch_bounds__(256)
head_scores_kernel(const float* __restrict__ logits, HeadLayout L, const int32_t* __restrict__ img_off, int B, int R,
                   const float* __restrict__ det_max, const float* __restrict__ det_sum, float* __restrict__ final_score,
                   float* __restrict__ sm1, float* __restrict__ sm2) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= R) return;
  const int b = image_of_row(img_off, B, j);
  const float* row = logits + (size_t)j * L.ld;
  float v[3], m, s;
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] = (lane + 32 * k < L.C) ? row[L.cls() + lane + 32 * k] : 0.f;
  warp_softmax3(v, lane, L.C, m, s);
  const float inv = (1.f / s);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int c = lane + 32 * k;
    if (c < L.C) {
      const float d = expf(row[L.det() + c] - det_max[b * L.C + c]) / det_sum[b * L.C + c];
      final_score[(size_t)j * L.C + c] = v[k] * inv * d;
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = (lane + 32 * k < L.C) ? row[L.ref(i) + lane + 32 * k] : 0.f;
    warp_softmax3(v, lane, L.C, m, s);
    const float is = (1.f / s);
    float* o = (i == 0 ? sm1 : sm2) + (size_t)j * L.C;
#pragma unroll
    for (int k = 0; k < 3; ++k) if (lane + 32 * k < L.C) o[lane + 32 * k] = v[k] * is;
  }
}

// warp per proposal: losses of the three refinement branches + the gradient of all seven losses w.r.t. the logits
__global__ void __launch_bounds__(256)
head_loss_kernel(const float* __restrict__ logits, HeadLayout L, int cls_agnostic, const int32_t* __restrict__ img_off,
                 int B, int R, const float* __restrict__ det_max, const float* __restrict__ det_sum,
                 const float* __restrict__ img_score, const float* __restrict__ img_labels,
                 const int64_t* __restrict__ pl, const float* __restrict__ lw, const float* __restrict__ rt, float eps,
                 float* __restrict__ grad, float* __restrict__ partial) {
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + wid;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (j < R) {
    const int b = image_of_row(img_off, B, j);
    const float invN = 1.f / (float)(img_off[b + 1] - img_off[b]);
    const float gs = invN / (float)B;
    const float* row = logits + (size_t)j * L.ld;
    float* grow = grad + (size_t)j * L.ld;
    // ---- MIL: loss_img = mean_c BCE(clamp(sum_j cls_sm * det_sm)) / B     (loss.py:234-246,352-354)
    float v[3], m, s, dimg[3], dsm[3], xr[3];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = (lane + 32 * k < L.C) ? row[L.cls() + lane + 32 * k] : 0.f;
    warp_softmax3(v, lane, L.C, m, s);
    const float inv = (1.f / s);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int c = lane + 32 * k;
      dimg[k] = dsm[k] = xr[k] = 0.f;
      v[k] *= inv;                                             // cls softmax
      if (c < L.C) {
        const float x_raw = img_score[b * L.C + c], y = img_labels[b * L.C + c];
        const float x = fminf(fmaxf(x_raw, eps), 1.f - eps);
        const bool inside = x_raw >= eps && x_raw <= 1.f - eps;   // torch.clamp passes the gradient inside [min, max]
        dimg[k] = inside ? (-y / x + (1.f - y) / (1.f - x)) / ((float)L.C * (float)B) : 0.f;
        dsm[k] = expf(row[L.det() + c] - det_max[b * L.C + c]) / det_sum[b * L.C + c];
        xr[k] = x_raw;
        dot += v[k] * dimg[k] * dsm[k];
      }
    }
    dot = odw_warp_sum(dot);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int c = lane + 32 * k;
      if (c < L.C) {
        grow[L.cls() + c] = v[k] * (dimg[k] * dsm[k] - dot);
        grow[L.det() + c] = dsm[k] * dimg[k] * (v[k] - xr[k]);
      }
    }
    // ---- refinement branches
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
      const float lm = i == 0 ? 3.f : 1.f;                      // loss.py:373
      const int label = (int)pl[(size_t)i * R + j];
      const float w = lw[(size_t)i * R + j];
#pragma unroll
      for (int k = 0; k < 3; ++k) v[k] = (lane + 32 * k < L.C) ? row[L.ref(i) + lane + 32 * k] : 0.f;
      const float zl = row[L.ref(i) + label];
      warp_softmax3(v, lane, L.C, m, s);
      const float ce = (m + logf(s)) - zl;                    // -log softmax[label]
      const float is = (1.f / s);
      const float gsc = lm * w * gs;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int c = lane + 32 * k;
        if (c < L.C) grow[L.ref(i) + c] = gsc * (v[k] * is - (c == label ? 1.f : 0.f));
      }
      if (lane == 0) acc[2 * i] = lm * ce * w * invN;
      // smooth-L1 (beta = 1) on the box of the pseudo label's class, positives only (loss.py:380-394)
      for (int q = lane; q < L.Q; q += 32) grow[L.bb(i) + q] = 0.f;
      __syncwarp();
      if (label > 0 && lane < 4) {
        const int k0 = (cls_agnostic ? 4 : 4 * label) + lane;
        const float d = row[L.bb(i) + k0] - rt[((size_t)i * R + j) * 4 + lane];
        const float n = fabsf(d);
        float sl = n < 1.f ? 0.5f * n * n : n - 0.5f;
        grow[L.bb(i) + k0] = gsc * (n < 1.f ? d : (d > 0.f ? 1.f : -1.f));
        sl += __shfl_xor_sync(0xFu, sl, 1);
        sl += __shfl_xor_sync(0xFu, sl, 2);
        if (lane == 0) acc[2 * i + 1] = lm * w * sl * invN;
      }
    }
  }
  __shared__ float sm[8][6];
  if (lane == 0)
    for (int k = 0; k < 6; ++k) sm[wid][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 6) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
    partial[(size_t)blockIdx.x * 6 + threadIdx.x] = t;
  }
}

// one CTA: fixed-order sums, loss_img, the /B of loss.py:403-406, and compute_avg_img_accuracy (loss.py:25-34)
__global__ void __launch_bounds__(256)
head_finalize_kernel(const float* __restrict__ partial, int nblk, const float* __restrict__ img_score,
                     const float* __restrict__ img_labels, const float* __restrict__ ref_colsum, int B, int C, float eps,
                     float* __restrict__ out) {
  __shared__ float red[256];
  __shared__ float res[11];
  const int t = threadIdx.x;
  for (int k = 0; k < 6; ++k) {                                 // per-block partials, strided then tree (fixed order)
    float a = 0.f;
    for (int i = t; i < nblk; i += 256) a += partial[(size_t)i * 6 + k];
    red[t] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (t < o) red[t] += red[t + o];
      __syncthreads();
    }
    if (t == 0) res[1 + k] = red[0] / (float)B;
    __syncthreads();
  }
  // loss_img: sum_b mean_c BCE(clamp(img_score), label) / B; log clamped at -100 as torch's binary_cross_entropy
  float a = 0.f;
  for (int i = t; i < B * C; i += 256) {
    const float x = fminf(fmaxf(img_score[i], eps), 1.f - eps), y = img_labels[i];
    a += -(y * fmaxf(logf(x), -100.f) + (1.f - y) * fmaxf(log1pf(-x), -100.f));
  }
  red[t] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  if (t == 0) res[0] = red[0] / (float)C / (float)B;
  __syncthreads();
  // accuracies: mean over images of mean(labels[top-k classes of the score]), k = max(#positive classes, 1);
  // 0: img_score over all C columns; 1..3: column sums of the refinement logits over classes 1..C-1
  for (int which = 0; which < 4; ++which) {
    float tot = 0.f;
    for (int b = 0; b < B; ++b) {
      const float* sc = which == 0 ? img_score + b * C : ref_colsum + ((size_t)(which - 1) * B + b) * C;
      const float* lab = img_labels + b * C;
      const int c0 = which == 0 ? 0 : 1;
      float kf = 0.f;
      for (int c = 0; c < C; ++c) kf += lab[c] > 0.5f ? 1.f : 0.f;
      const int k = kf < 1.f ? 1 : (int)kf;
      float hit = 0.f;
      if (t >= c0 && t < C) {
        float x = sc[t];
        if (which == 0) x = fminf(fmaxf(x, eps), 1.f - eps);
        int rank = 0;
        for (int c = c0; c < C; ++c) {
          float xc = sc[c];
          if (which == 0) xc = fminf(fmaxf(xc, eps), 1.f - eps);
          rank += (xc > x || (xc == x && c < t)) ? 1 : 0;
        }
        if (rank < k) hit = lab[t];
      }
      red[t] = hit;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1) {
        if (t < o) red[t] += red[t + o];
        __syncthreads();
      }
      tot += red[0] / (float)k;
      __syncthreads();
    }
    if (t == 0) res[7 + which] = tot / (float)B;
    __syncthreads();
  }
  if (t < 11) out[t] = res[t];
}

// grad[j, block] *= g[loss of that block]   (the backward of the whole head-loss node)
__global__ void __launch_bounds__(256)
head_grad_scale_kernel(float* __restrict__ grad, HeadLayout L, long long R, const float* __restrict__ g) {
  const int W = L.width();
  const long long total = R * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long j = idx / W;
    const int c = (int)(idx - j * W);
    int which;
    if (c < 2 * L.C) which = 0;
    else {
      const int r = c - 2 * L.C, blk = r / (L.C + L.Q);
      which = 1 + 2 * blk + ((r - blk * (L.C + L.Q)) >= L.C ? 1 : 0);
    }
    grad[j * L.ld + c] *= g[which];
  }
}

}  // namespace

ODW_API size_t odwscl_head_scores_ws_bytes(int B, int C) {
  return B > 0 && C > 0 ? (size_t)B * kColSplit * 5 * C * sizeof(float) : 0;
}

ODW_API int odwscl_head_scores_f32(const float* logits, int ld, int R, int C, int Q, const int32_t* img_off, int B,
                                   float* det_max, float* det_sum, float* ref_colsum, float* final_score, float* sm1,
                                   float* sm2, float* img_score, void* ws, size_t ws_bytes, odwscl_stream_t stream) {
  if (R < 0 || B < 0 || C < 2 || C > kMaxC || Q < 0 || ld < 5 * C + 3 * Q) return ODWSCL_EINVAL;
  if (B == 0) return 0;
  if (!img_off || !det_max || !det_sum || !ref_colsum || !img_score) return ODWSCL_EINVAL;
  if (R > 0 && (!logits || !final_score || !sm1 || !sm2)) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const HeadLayout L{C, Q, ld};
  if (!ws || ws_bytes < odwscl_head_scores_ws_bytes(B, C)) return ODWSCL_ENOWS;
  float* partial = reinterpret_cast<float*>(ws);
  dim3 g1(B, odw_cdiv(C, 32), kColSplit), g2(B, odw_cdiv(C, 32));
  head_colstats_partial_kernel<<<g1, 256, 0, st>>>(logits, L, img_off, partial);
  ODW_LAUNCH_CHECK();
  head_colstats_final_kernel<<<g2, 32, 0, st>>>(partial, C, B, det_max, det_sum, ref_colsum);
  ODW_LAUNCH_CHECK();
  if (R > 0) {
    head_scores_kernel<<<odw_cdiv(R, 8), 256, 0, st>>>(logits, L, img_off, B, R, det_max, det_sum, final_score, sm1, sm2);
    ODW_LAUNCH_CHECK();
  }
  seg_colsum_partial_kernel<<<g1, 256, 0, st>>>(final_score, C, C, img_off, partial);
  ODW_LAUNCH_CHECK();
  seg_colsum_final_kernel<<<g2, 32, 0, st>>>(partial, C, img_score);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_head_loss_f32(const float* logits, int ld, int R, int C, int Q, int cls_agnostic, const int32_t* img_off,
                                 int B, const float* det_max, const float* det_sum, const float* img_score,
                                 const float* img_labels, const int64_t* pseudo_labels, const float* label_weights,
                                 const float* reg_targets, const float* ref_colsum, float eps, float* grad_logits,
                                 float* partial, float* out11, odwscl_stream_t stream) {
  if (R < 0 || B <= 0 || C < 2 || C > kMaxC || Q < 0 || ld < 5 * C + 3 * Q) return ODWSCL_EINVAL;
  if (!img_off || !det_max || !det_sum || !img_score || !img_labels || !ref_colsum || !partial || !out11) return ODWSCL_EINVAL;
  if (R > 0 && (!logits || !pseudo_labels || !label_weights || !reg_targets || !grad_logits)) return ODWSCL_EINVAL;
  if (Q > 0 && Q < (cls_agnostic ? 8 : 4 * C)) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const HeadLayout L{C, Q, ld};
  const int nblk = odw_cdiv(R, 8);
  if (R > 0) {
    head_loss_kernel<<<nblk, 256, 0, st>>>(logits, L, cls_agnostic, img_off, B, R, det_max, det_sum, img_score, img_labels,
                                           pseudo_labels, label_weights, reg_targets, eps, grad_logits, partial);
    ODW_LAUNCH_CHECK();
  }
  head_finalize_kernel<<<1, 256, 0, st>>>(partial, nblk, img_score, img_labels, ref_colsum, B, C, eps, out11);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_head_grad_scale_f32(float* grad_logits, int ld, long long R, int C, int Q, const float* upstream7,
                                       odwscl_stream_t stream) {
  if (R < 0 || C < 2 || Q < 0 || ld < 5 * C + 3 * Q) return ODWSCL_EINVAL;
  if (R == 0) return 0;
  if (!grad_logits || !upstream7) return ODWSCL_EINVAL;
  const HeadLayout L{C, Q, ld};
  const long long total = R * L.width();
  head_grad_scale_kernel<<<(int)min((long long)ODW_NUM_SMS * 8, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      grad_logits, L, R, upstream7);
  ODW_LAUNCH_CHECK();
  return 0;
}
