// supcon_tc.cu -- SupConLossV2 forward / backward with the M x M contractions on the tensor cores.
// Reference: roi_heads/sim_head/sim_loss.py:49-80 (one cuBLAS GEMM + ~12 elementwise / reduction kernels over five M x M
// temporaries, doubled by autograd).
//
// supcon.cu keeps the M x M similarity on chip but forms it with FFMA tiles: fine for the ~1.1 k-row bank of a 2-image
// rank, quadratic beyond it (bs 8 per rank: M = 4-6 k).  This path hands both contractions to the persistent tcgen05
// CTA-pair GEMM of fc_gemm.cu at fp32-class accuracy (3xTF32: hi*hi + hi*lo + lo*hi, the dropped lo*lo term is 2^-22):
//   forward   bank rows V = [F ; E][row_src] are gathered ONCE into A3 = [Vh | Vh | Vl], B3 = [Vh | Vl | Vh]
//             (Mp x 384), so S = V V^T is a single K = 384 GEMM; one warp per row folds S/T into (max, pos, all),
//             a one-CTA fixed-order mean finishes the loss (as in supcon.cu);
//   backward  H_rq = (G_rq + G_qr) / T,  G_rq = (g w_r / M) e^{S_rq/T - m_r} (1/all_r - [y_r = y_q]/pos_r), r != q, is
//             written over S as its TF32 hi part plus a separate lo part; dV = Hh Vh + Hh Vl + Hl Vh is two launches of
//             the same GEMM (second operand pair + accumulate epilogue); rows are scattered back through row_src with
//             float reductions (a proposal can appear in the bank more than once).
// M is read from device memory; the GEMMs run on the padded bound Mp (rows >= M are zero).
#include "common.cuh"

extern "C" int odwscl_fc_gemm_tf32(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major, float* C,
                                   int ldc, int M, int N, int K, int flags, const float* bias, const float* mask_src,
                                   int ld_mask, float mask_scale, float dropout_p, unsigned long long seed, int max_pairs,
                                   const float* A2, int lda2, const float* B2, int ldb2, int K2, odwscl_stream_t stream);

namespace {

constexpr int kD = ODWSCL_SIM_DIM;   // 128

__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t b;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v));
  return __uint_as_float(b);
}

// A3[r] = [Vh | Vh | Vl], B3[r] = [Vh | Vl | Vh]; rows >= M are zero.  thread = (row, 4 columns)
__global__ void supcon_gather_split_kernel(const float* __restrict__ F, const float* __restrict__ E, int R,
                                           const int32_t* __restrict__ row_src, const int32_t* __restrict__ M_dev, int Mcap,
                                           int Mp, float* __restrict__ A3, float* __restrict__ B3) {
  const int M = min(*M_dev, Mcap);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Mp * (kD / 4); t += gridDim.x * blockDim.x) {
    const int r = t / (kD / 4), k4 = t % (kD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < M) {
      const int src = __ldg(row_src + r);
      const float* row = src < R ? F + (size_t)src * kD : E + (size_t)(src - R) * kD;
      v = __ldg(reinterpret_cast<const float4*>(row) + k4);
    }
    const float4 h = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
    const float4 l = make_float4(rna_tf32(v.x - h.x), rna_tf32(v.y - h.y), rna_tf32(v.z - h.z), rna_tf32(v.w - h.w));
    float4* a = reinterpret_cast<float4*>(A3 + (size_t)r * 3 * kD) + k4;
    float4* b = reinterpret_cast<float4*>(B3 + (size_t)r * 3 * kD) + k4;
    a[0] = h; a[kD / 4] = h; a[2 * (kD / 4)] = l;
    b[0] = h; b[kD / 4] = l; b[2 * (kD / 4)] = h;
  }
}

__device__ __forceinline__ void merge_stats(float& m, float& p, float& a, float m2, float p2, float a2) {
  const float nm = fmaxf(m, m2);
  const float s1 = (m == -INFINITY) ? 0.f : expf(m - nm);
  const float s2 = (m2 == -INFINITY) ? 0.f : expf(m2 - nm);
  p = p * s1 + p2 * s2;
  a = a * s1 + a2 * s2;
  m = nm;
}

// one warp per bank row: (row max, positive sum, all sum) of S[r][c] / T over c < M, c != r -> parts[r]
__global__ void __launch_bounds__(256)
supcon_rowstats_kernel(const float* __restrict__ S, int ld, const int32_t* __restrict__ row_lab,
                       const int32_t* __restrict__ M_dev, int Mcap, float inv_temp, float4* __restrict__ parts) {
  const int M = min(*M_dev, Mcap);
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= Mcap) return;
  if (r >= M) return;
  const int lab = __ldg(row_lab + r);
  const float* row = S + (size_t)r * ld;
  float m = -INFINITY, p = 0.f, a = 0.f;
  for (int c = lane; c < M; c += 32) {
    if (c == r) continue;
    const float s = __ldg(row + c) * inv_temp;
    const float nm = fmaxf(m, s);
    const float sc = (m == -INFINITY) ? 0.f : expf(m - nm);
    const float e = expf(s - nm);
    p = p * sc + ((__ldg(row_lab + c) == lab) ? e : 0.f);
    a = a * sc + e;
    m = nm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float p2 = __shfl_xor_sync(0xffffffffu, p, o);
    const float a2 = __shfl_xor_sync(0xffffffffu, a, o);
    merge_stats(m, p, a, m2, p2, a2);
  }
  if (lane == 0) parts[r] = make_float4(m, p, a, 0.f);
}

// H over S in place (hi part) + lo part; zero outside [0,M)^2 and on the diagonal.  thread = 4 consecutive columns
__global__ void __launch_bounds__(256)
supcon_hmat_kernel(float* __restrict__ S, float* __restrict__ Hl, int ld, int Mp, const int32_t* __restrict__ row_lab,
                   const float* __restrict__ row_w, const float4* __restrict__ stats, const int32_t* __restrict__ M_dev,
                   int Mcap, float inv_temp, const float* __restrict__ gscale_dev) {
  const int M = min(*M_dev, Mcap);
  const int r = blockIdx.y;
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= Mp) return;
  float4* sp = reinterpret_cast<float4*>(S + (size_t)r * ld + c0);
  float4* lp = reinterpret_cast<float4*>(Hl + (size_t)r * ld + c0);
  if (r >= M || c0 >= M) {
    *sp = make_float4(0.f, 0.f, 0.f, 0.f);
    *lp = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float gM = __ldg(gscale_dev) / (float)M;
  const float4 st = stats[r];
  const float r_m = st.x, r_ip = 1.f / st.y, r_ia = 1.f / st.z, r_cf = gM * __ldg(row_w + r);
  const int r_lab = __ldg(row_lab + r);
  const float4 sv = *sp;
  const float s4[4] = {sv.x, sv.y, sv.z, sv.w};
  float hh[4], hl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + j;
    float h = 0.f;
    if (c < M && c != r) {
      const float4 ct = stats[c];
      const float s = s4[j] * inv_temp;
      const bool same = (__ldg(row_lab + c) == r_lab);
      const float g_rq = r_cf * expf(s - r_m) * (r_ia - (same ? r_ip : 0.f));
      const float g_qr = gM * __ldg(row_w + c) * expf(s - ct.x) * (1.f / ct.z - (same ? 1.f / ct.y : 0.f));
      h = (g_rq + g_qr) * inv_temp;
    }
    hh[j] = rna_tf32(h);
    hl[j] = rna_tf32(h - hh[j]);
  }
  *sp = make_float4(hh[0], hh[1], hh[2], hh[3]);
  *lp = make_float4(hl[0], hl[1], hl[2], hl[3]);
}

__global__ void supcon_scatter_kernel(const float* __restrict__ dV, const int32_t* __restrict__ row_src,
                                      const int32_t* __restrict__ M_dev, int Mcap, int R, float* __restrict__ dF,
                                      float* __restrict__ dE) {
  const int M = min(*M_dev, Mcap);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < M * kD; t += gridDim.x * blockDim.x) {
    const int r = t / kD, d = t % kD;
    const int src = __ldg(row_src + r);
    float* dst = src < R ? dF + (size_t)src * kD : dE + (size_t)(src - R) * kD;
    atomicAdd(dst + d, dV[t]);
  }
}

// the fixed-order merge + mean of supcon.cu, restated for this file's partials (same arithmetic)
__global__ void __launch_bounds__(1024, 1)
supcon_tc_mean_kernel(float4* __restrict__ stats, const float4* __restrict__ parts, const float* __restrict__ row_w,
                      const int32_t* __restrict__ M_dev, int Mcap, float* __restrict__ loss_out) {
  __shared__ float s_v[32];
  const int M = min(*M_dev, Mcap);
  float part = 0.f;
  for (int r = threadIdx.x; r < M; r += blockDim.x) {
    const float4 q = parts[r];
    const float lr = -logf(q.y / q.z) * __ldg(row_w + r);    // sim_loss.py:76-78
    stats[r] = make_float4(q.x, q.y, q.z, lr);
    part += lr;
  }
  part = odw_warp_sum(part);
  if ((threadIdx.x & 31) == 0) s_v[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = s_v[threadIdx.x];
    v = odw_warp_sum(v);
    if (threadIdx.x == 0) *loss_out = M > 0 ? v / (float)M : 0.f;       // sim_loss.py:80 .mean()
  }
}

inline int pad_rows(int Mcap) { return (Mcap + 3) / 4 * 4; }

struct TcWs {
  float *A3, *B3, *S, *Hl, *dV;
};
inline TcWs carve(float* ws, int Mp) {
  TcWs w;
  w.A3 = ws;
  w.B3 = w.A3 + (size_t)Mp * 3 * kD;
  w.S = w.B3 + (size_t)Mp * 3 * kD;
  w.Hl = w.S + (size_t)Mp * Mp;
  w.dV = w.Hl + (size_t)Mp * Mp;
  return w;
}

}  // namespace

ODW_API size_t odwscl_supcon_tc_ws_bytes(int Mcap) {
  if (Mcap <= 0) return 0;
  const size_t Mp = (size_t)pad_rows(Mcap);
  return sizeof(float) * (2 * Mp * 3 * kD + 2 * Mp * Mp + Mp * kD);
}

ODW_API int odwscl_supcon_tc_fwd_f32(const float* F, const float* E, int R, const int32_t* row_src, const int32_t* row_lab,
                                     const float* row_w, const int32_t* M_dev, int Mcap, float inv_temp, float* ws,
                                     size_t ws_bytes, float* stats, float* loss_out, odwscl_stream_t stream) {
  if (R < 0 || Mcap < 0) return ODWSCL_EINVAL;
  if (!loss_out) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (Mcap == 0) { ODW_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st)); return 0; }
  if (!F || !row_src || !row_lab || !row_w || !M_dev || !stats || !ws) return ODWSCL_EINVAL;
  if (ws_bytes < odwscl_supcon_tc_ws_bytes(Mcap)) return ODWSCL_ENOWS;
  if (((uintptr_t)ws & 15) || Mcap > 65532) return ODWSCL_EINVAL;
  const int Mp = pad_rows(Mcap);
  const TcWs w = carve(ws, Mp);
  supcon_gather_split_kernel<<<min(ODW_NUM_SMS * 4, odw_cdiv(Mp * (kD / 4), 256)), 256, 0, st>>>(F, E, R, row_src, M_dev, Mcap,
                                                                                               Mp, w.A3, w.B3);
  ODW_LAUNCH_CHECK();
  int rc = odwscl_fc_gemm_tf32(w.A3, 3 * kD, 0, w.B3, 3 * kD, 0, w.S, Mp, Mp, Mp, 3 * kD, 0, nullptr, nullptr, 0, 1.f, 0.f, 0ull, 0,
                               nullptr, 0, nullptr, 0, 0, stream);
  if (rc) return rc;
  float4* parts = reinterpret_cast<float4*>(stats) + Mcap;              // [kSplit][Mcap] behind the merged rows
  supcon_rowstats_kernel<<<odw_cdiv(Mcap, 8), 256, 0, st>>>(w.S, Mp, row_lab, M_dev, Mcap, inv_temp, parts);
  ODW_LAUNCH_CHECK();
  supcon_tc_mean_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<float4*>(stats), parts, row_w, M_dev, Mcap, loss_out);
  ODW_LAUNCH_CHECK();
  return 0;
}

// `ws` must be the workspace the forward call of the same bank filled (it still holds S and the split bank rows).
ODW_API int odwscl_supcon_tc_bwd_f32(int R, const int32_t* row_src, const int32_t* row_lab, const float* row_w,
                                     const int32_t* M_dev, int Mcap, float inv_temp, float* ws, size_t ws_bytes,
                                     const float* stats, const float* gscale_dev, float* dF, float* dE,
                                     odwscl_stream_t stream) {
  if (R < 0 || Mcap < 0) return ODWSCL_EINVAL;
  if (Mcap == 0) return 0;
  if (!row_src || !row_lab || !row_w || !M_dev || !stats || !gscale_dev || !dF || !ws) return ODWSCL_EINVAL;
  if (ws_bytes < odwscl_supcon_tc_ws_bytes(Mcap)) return ODWSCL_ENOWS;
  cudaStream_t st = (cudaStream_t)stream;
  const int Mp = pad_rows(Mcap);
  const TcWs w = carve(ws, Mp);
  supcon_hmat_kernel<<<dim3(odw_cdiv(Mp / 4, 256), Mp), 256, 0, st>>>(w.S, w.Hl, Mp, Mp, row_lab, row_w,
                                                                      reinterpret_cast<const float4*>(stats), M_dev, Mcap,
                                                                      inv_temp, gscale_dev);
  ODW_LAUNCH_CHECK();
  // dV = Hh Vh + Hh Vl (second operand pair), then += Hl Vh (accumulate epilogue).  Vh / Vl: column blocks of A3, read
  // MN-major ([K = Mp, N = 128], pitch 384)
  int rc = odwscl_fc_gemm_tf32(w.S, Mp, 0, w.A3, 3 * kD, 1, w.dV, kD, Mp, kD, Mp, 0, nullptr, nullptr, 0, 1.f, 0.f, 0ull, 0, w.S,
                               Mp, w.A3 + 2 * kD, 3 * kD, Mp, stream);
  if (rc) return rc;
  rc = odwscl_fc_gemm_tf32(w.Hl, Mp, 0, w.A3, 3 * kD, 1, w.dV, kD, Mp, kD, Mp, ODWSCL_FC_ACCUM, nullptr, nullptr, 0, 1.f, 0.f, 0ull,
                           0, nullptr, 0, nullptr, 0, 0, stream);
  if (rc) return rc;
  supcon_scatter_kernel<<<min(ODW_NUM_SMS * 4, odw_cdiv(Mcap * kD, 256)), 256, 0, st>>>(w.dV, row_src, M_dev, Mcap, R, dF, dE);
  ODW_LAUNCH_CHECK();
  return 0;
}
