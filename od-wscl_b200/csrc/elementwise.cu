// elementwise.cu -- small fused passes around the GEMM / conv kernels of the path.
//
//   relu_dropout       ReLU + Dropout(p) of the fc6 / fc7 outputs (modeling/backbone/vgg16.py:122-130: nn.ReLU(True),
//                      nn.Dropout()) as ONE in-place pass with a counter-based RNG (Philox4x32-10): no mask tensor.  The
//                      backward needs no mask either: y > 0 exactly where the unit was positive AND kept, so
//                      gx = gy * scale * [y > 0] (one pass over gy and the saved activation).
//   conv_weight_xform  [Cout,Cin,3,3] (the reference's state-dict layout) -> the two operand layouts the tcgen05 conv
//                      kernels read, TF32-rounded, in one pass: [Cout,3,3,Cin] for fprop and the tap-flipped
//                      [Cin,3,3,Cout] for dgrad.
#include "common.cuh"

namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// 4 x 32 random bits for (seed, counter)
__device__ __forceinline__ void philox4x32_10(unsigned long long seed, unsigned long long ctr, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

__global__ void __launch_bounds__(256)
relu_dropout_fwd_kernel(float4* __restrict__ x, long long n4, float p, float scale, unsigned long long seed) {
  // keep iff u >= p with u uniform on [0,1) from 24 random bits (torch's bernoulli(1 - p) convention)
  const uint32_t thr = (uint32_t)(p * 16777216.0f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10(seed, (unsigned long long)i, r);
    float4 v = x[i];
    v.x = (v.x > 0.f && (r[0] >> 8) >= thr) ? v.x * scale : 0.f;
    v.y = (v.y > 0.f && (r[1] >> 8) >= thr) ? v.y * scale : 0.f;
    v.z = (v.z > 0.f && (r[2] >> 8) >= thr) ? v.z * scale : 0.f;
    v.w = (v.w > 0.f && (r[3] >> 8) >= thr) ? v.w * scale : 0.f;
    x[i] = v;
  }
}

__global__ void __launch_bounds__(256)
relu_dropout_bwd_kernel(const float4* __restrict__ y, const float4* __restrict__ gy, float4* __restrict__ gx, long long n4,
                        float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(y + i), g = __ldcs(gy + i);
    float4 o;
    o.x = a.x > 0.f ? g.x * scale : 0.f;
    o.y = a.y > 0.f ? g.y * scale : 0.f;
    o.z = a.z > 0.f ? g.z * scale : 0.f;
    o.w = a.w > 0.f ? g.w * scale : 0.f;
    gx[i] = o;
  }
}

__device__ __forceinline__ float rna_tf32_ew(float v) {
  uint32_t b;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v));
  return __uint_as_float(b);
}

// CTA = 32 output channels x 32 input channels x 9 taps, staged in shared memory: the [Cout,Cin,9] source is read in
// contiguous 1152-byte rows, the fprop layout [co][tap][ci] is written with ci fastest and the dgrad layout
// [ci][8-tap][co] with co fastest -- all three coalesced.
__global__ void __launch_bounds__(256)
conv_weight_xform_kernel(const float* __restrict__ w, int Cout, int Cin, float* __restrict__ w_krsc,
                         float* __restrict__ w_crsk_flip, int round_tf32) {
  __shared__ float s[32][32 * 9 + 1];                       // [co][ci*9 + tap]
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < 32 * 288; i += blockDim.x) {
    const int co = i / 288, k = i % 288;                    // k = ci*9 + tap, contiguous in the source
    float v = 0.f;
    if (co0 + co < Cout && ci0 + k / 9 < Cin) v = __ldg(w + ((size_t)(co0 + co) * Cin + ci0) * 9 + k);
    s[co][k] = round_tf32 ? rna_tf32_ew(v) : v;
  }
  __syncthreads();
  if (w_krsc)
    for (int i = threadIdx.x; i < 32 * 9 * 32; i += blockDim.x) {
      const int ci = i % 32, t = (i / 32) % 9, co = i / 288;
      if (co0 + co < Cout && ci0 + ci < Cin) w_krsc[((size_t)(co0 + co) * 9 + t) * Cin + ci0 + ci] = s[co][ci * 9 + t];
    }
  if (w_crsk_flip)
    for (int i = threadIdx.x; i < 32 * 9 * 32; i += blockDim.x) {
      const int co = i % 32, t = (i / 32) % 9, ci = i / 288;
      if (co0 + co < Cout && ci0 + ci < Cin)
        w_crsk_flip[((size_t)(ci0 + ci) * 9 + (8 - t)) * Cout + co0 + co] = s[co][ci * 9 + t];
    }
}

}  // namespace

ODW_API int odwscl_relu_dropout_fwd_f32(float* x, long long n, float p, unsigned long long seed, odwscl_stream_t stream) {
  if (n < 0 || (n & 3) || p < 0.f || p >= 1.f) return ODWSCL_EINVAL;
  if (n == 0) return 0;
  if (!x) return ODWSCL_EINVAL;
  const long long n4 = n / 4;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 8, (n4 + 255) / 256);
  relu_dropout_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(x), n4, p, 1.f / (1.f - p), seed);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_relu_dropout_bwd_f32(const float* y, const float* gy, float* gx, long long n, float p,
                                        odwscl_stream_t stream) {
  if (n < 0 || (n & 3) || p < 0.f || p >= 1.f) return ODWSCL_EINVAL;
  if (n == 0) return 0;
  if (!y || !gy || !gx) return ODWSCL_EINVAL;
  const long long n4 = n / 4;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 8, (n4 + 255) / 256);
  relu_dropout_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(y),
                                                                    reinterpret_cast<const float4*>(gy),
                                                                    reinterpret_cast<float4*>(gx), n4, 1.f / (1.f - p));
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_conv_weight_xform_f32(const float* w_oihw, int Cout, int Cin, float* w_krsc, float* w_crsk_flip,
                                         int round_tf32, odwscl_stream_t stream) {
  if (Cout <= 0 || Cin <= 0) return ODWSCL_EINVAL;
  if (!w_oihw || (!w_krsc && !w_crsk_flip)) return ODWSCL_EINVAL;
  dim3 grid(odw_cdiv(Cin, 32), odw_cdiv(Cout, 32));
  conv_weight_xform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, w_krsc, w_crsk_flip, round_tf32);
  ODW_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// l2norm_rows: F.normalize(z, dim=1) of Sim_Net (roi_heads/sim_head/sim_net.py:26; eps = 1e-12) forward and backward
// as one launch each (torch: norm + clamp + expand + div, and ~14 kernels in the backward).  Warp per row.
//   y = z / max(||z||, eps);   dz = (g - y (y . g)) / max(||z||, eps)     (||z|| > eps; else dz = g / eps)
namespace {

__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ z, int ldz, int R, int D, float eps, float* __restrict__ y,
                  float* __restrict__ inv_norm) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* zr = z + (size_t)r * ldz;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = zr[c]; ss += v * v; }
  ss = odw_warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  for (int c = lane; c < D; c += 32) y[(size_t)r * D + c] = zr[c] * inv;
  if (lane == 0) inv_norm[r] = inv;
}

__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ g, const float* __restrict__ inv_norm, int R,
                  int D, float eps, float* __restrict__ dz) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* yr = y + (size_t)r * D;
  const float* gr = g + (size_t)r * D;
  float dot = 0.f;
  for (int c = lane; c < D; c += 32) dot += yr[c] * gr[c];
  dot = odw_warp_sum(dot);
  const float inv = inv_norm[r];
  const bool clamped = inv >= 1.f / eps;                   // ||z|| <= eps: y = z / eps is linear in z
  for (int c = lane; c < D; c += 32) dz[(size_t)r * D + c] = clamped ? gr[c] * inv : (gr[c] - yr[c] * dot) * inv;
}

// The index glue of the sync-free contrastive branch (modeling/loss.py, speculative K) in ONE launch instead of ~25 eager
// kernels: rows[k] = rowsA[k] for k < min(K, Kc) else 0 (the padded augmented-positives batch), sel[j] for j < sel_n =
// the row of the padded [2 Kc, 128] embedding matrix that entry j of the [2 K] layout (drop rows, then noise rows) reads,
// padding entries spread over rows; overflow = (K > Kc).
__global__ void __launch_bounds__(256)
spec_index_kernel(const int32_t* __restrict__ k_dev, const int32_t* __restrict__ rowsA, int Kc, long long sel_n,
                  int64_t* __restrict__ rows, int64_t* __restrict__ sel, float* __restrict__ overflow) {
  const long long K = *k_dev, kv = K < Kc ? K : Kc;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i == 0) *overflow = K > Kc ? 1.f : 0.f;
  if (i < Kc) rows[i] = i < kv ? (long long)rowsA[i] : 0;
  if (i < sel_n) {
    const long long pad = i % (2LL * Kc), jj = i - K;
    sel[i] = i < K ? (i < kv ? i : pad) : ((jj < kv && i < 2 * K) ? jj + Kc : pad);
  }
}

}  // namespace

ODW_API int odwscl_l2norm_fwd_f32(const float* z, int ldz, int R, int D, float eps, float* y, float* inv_norm,
                                  odwscl_stream_t stream) {
  if (R < 0 || D <= 0 || ldz < D || eps <= 0.f) return ODWSCL_EINVAL;
  if (R == 0) return 0;
  if (!z || !y || !inv_norm) return ODWSCL_EINVAL;
  l2norm_fwd_kernel<<<odw_cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(z, ldz, R, D, eps, y, inv_norm);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_l2norm_bwd_f32(const float* y, const float* g, const float* inv_norm, int R, int D, float eps, float* dz,
                                  odwscl_stream_t stream) {
  if (R < 0 || D <= 0 || eps <= 0.f) return ODWSCL_EINVAL;
  if (R == 0) return 0;
  if (!y || !g || !inv_norm || !dz) return ODWSCL_EINVAL;
  l2norm_bwd_kernel<<<odw_cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(y, g, inv_norm, R, D, eps, dz);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_spec_index(const int32_t* k_dev, const int32_t* rowsA, int Kc, long long sel_n, int64_t* rows,
                              int64_t* sel, float* overflow, odwscl_stream_t stream) {
  if (Kc <= 0 || sel_n < 0) return ODWSCL_EINVAL;
  if (!k_dev || !rowsA || !rows || !sel || !overflow) return ODWSCL_EINVAL;
  const long long n = sel_n > Kc ? sel_n : Kc;
  spec_index_kernel<<<odw_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(k_dev, rowsA, Kc, sel_n, rows, sel, overflow);
  ODW_LAUNCH_CHECK();
  return 0;
}
