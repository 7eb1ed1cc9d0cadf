// supcon.cu -- fused SupConLossV2 forward / backward.
// Reference: roi_heads/sim_head/sim_loss.py:49-80 -- one cuBLAS GEMM plus ~12 elementwise /
// reduction kernels over five M x M fp32 temporaries, doubled again by autograd.
//
// Here the M x M similarity never reaches HBM.  Each CTA owns 64 bank rows and streams all 64-row
// column tiles through shared memory (k-major, so both operands are read as conflict-free
// float4s), forming a 64x64 fp32 tile per step in registers (4x4 per thread) and folding it
// straight into per-row online (max, positive-sum, all-sum) statistics; rows of the bank are
// gathered through `row_src` (the bank is a list of row ids into [F ; E], never materialised).
// The row-max shift cancels in pos/all, so the online max is only a stabiliser (the reference
// detaches it, sim_loss.py:63-64).  Backward recomputes the tiles, builds
//   H_rq = G_rq + G_qr,  G_rq = (g w_r / M) e^{S_rq/T - m_r} (1/all_r - [y_r = y_q]/pos_r),  r != q
// in shared memory and applies dV_r += (1/T) sum_q H_rq V_q as a second register-tiled product,
// scattering into dF / dE with fire-and-forget float reductions (a proposal can appear in the bank
// more than once).  M is read from device memory: no host synchronisation sizes the launch.
#include "common.cuh"

namespace {

constexpr int kD = ODWSCL_SIM_DIM;   // 128
constexpr int kT = 64;               // tile rows / cols
constexpr int kLd = kT + 4;          // k-major leading dimension (floats), keeps float4 alignment
constexpr int kThreads = 256;
constexpr int kSplit = ODWSCL_SUPCON_SPLITS;   // column splits: a bank of ~1100 rows is only 18 row tiles -> 18 x 8 CTAs

__device__ __forceinline__ const float* bank_row(const float* F, const float* E, int R, int src) {
  return src < R ? F + (size_t)src * kD : E + (size_t)(src - R) * kD;
}

// load 64 bank rows starting at r0 into k-major smem tile T[k][row]; rows >= M are zero-filled.
__device__ __forceinline__ void load_tile_kmajor(float* sT, const float* F, const float* E, int R,
                                                 const int32_t* row_src, int r0, int M) {
  for (int t = threadIdx.x; t < kT * (kD / 4); t += kThreads) {
    const int r = t / (kD / 4), k4 = t % (kD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < M) v = __ldg(reinterpret_cast<const float4*>(bank_row(F, E, R, __ldg(row_src + r0 + r))) + k4);
    sT[(4 * k4 + 0) * kLd + r] = v.x;
    sT[(4 * k4 + 1) * kLd + r] = v.y;
    sT[(4 * k4 + 2) * kLd + r] = v.z;
    sT[(4 * k4 + 3) * kLd + r] = v.w;
  }
}

// acc[u][v] = sum_k A[k][ty*4+u] * B[k][tx*4+v]
__device__ __forceinline__ void tile_product(const float* sA, const float* sB, int ty, int tx, float acc[4][4]) {
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
#pragma unroll 8
  for (int k = 0; k < kD; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(sA + k * kLd + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(sB + k * kLd + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
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

__global__ void __launch_bounds__(kThreads, 2)
supcon_fwd_kernel(const float* __restrict__ F, const float* __restrict__ E, int R,
                  const int32_t* __restrict__ row_src, const int32_t* __restrict__ row_lab,
                  const float* __restrict__ row_w, const int32_t* __restrict__ M_dev, int Mcap, float inv_temp,
                  float4* __restrict__ part) {
  extern __shared__ __align__(16) float smem_f[];
  float* sA = smem_f;
  float* sB = smem_f + kD * kLd;
  __shared__ int s_lab[kT];
  const int M = min(*M_dev, Mcap);
  const int r0 = blockIdx.x * kT;
  if (r0 >= M) return;
  // this CTA's share of the column tiles (blockIdx.y of kSplit); partial (max, pos, all) go to part[split][row]
  const int per_split = ((M + kT - 1) / kT + kSplit - 1) / kSplit * kT;
  const int c_begin = blockIdx.y * per_split, c_end = min(M, c_begin + per_split);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  load_tile_kmajor(sA, F, E, R, row_src, r0, M);
  int my_lab[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) my_lab[u] = (r0 + ty * 4 + u < M) ? __ldg(row_lab + r0 + ty * 4 + u) : -1;
  float m[4], ps[4], as[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) { m[u] = -INFINITY; ps[u] = 0.f; as[u] = 0.f; }

  for (int c0 = c_begin; c0 < c_end; c0 += kT) {
    __syncthreads();
    load_tile_kmajor(sB, F, E, R, row_src, c0, M);
    if (threadIdx.x < kT) s_lab[threadIdx.x] = (c0 + threadIdx.x < M) ? __ldg(row_lab + c0 + threadIdx.x) : -2;
    __syncthreads();
    float acc[4][4];
    tile_product(sA, sB, ty, tx, acc);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + ty * 4 + u;
      float s[4];
      float mx = -INFINITY;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int c = c0 + tx * 4 + v;
        s[v] = acc[u][v] * inv_temp;
        if (c < M) mx = fmaxf(mx, s[v]);
      }
      if (mx == -INFINITY) continue;
      const float nm = fmaxf(m[u], mx);
      const float sc = (m[u] == -INFINITY) ? 0.f : expf(m[u] - nm);
      float p = ps[u] * sc, a = as[u] * sc;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int c = c0 + tx * 4 + v;
        if (c < M && c != r) {
          const float e = expf(s[v] - nm);
          a += e;
          if (s_lab[tx * 4 + v] == my_lab[u]) p += e;
        }
      }
      m[u] = nm; ps[u] = p; as[u] = a;
    }
  }
  // merge the 16 column-slices (tx) of each row: lanes of one half-warp
#pragma unroll
  for (int u = 0; u < 4; ++u) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m[u], o);
      const float p2 = __shfl_xor_sync(0xffffffffu, ps[u], o);
      const float a2 = __shfl_xor_sync(0xffffffffu, as[u], o);
      merge_stats(m[u], ps[u], as[u], m2, p2, a2);
    }
    const int r = r0 + ty * 4 + u;
    if (tx == 0 && r < M) part[(size_t)blockIdx.y * Mcap + r] = make_float4(m[u], ps[u], as[u], 0.f);
  }
}

__global__ void __launch_bounds__(1024, 1)
supcon_mean_kernel(float4* __restrict__ stats, const float4* __restrict__ parts, const float* __restrict__ row_w,
                   const int32_t* __restrict__ M_dev, int Mcap, float* __restrict__ loss_out) {
  __shared__ float s_v[32];
  const int M = min(*M_dev, Mcap);
  float part = 0.f;
  for (int r = threadIdx.x; r < M; r += blockDim.x) {
    float m = -INFINITY, p = 0.f, a = 0.f;
    for (int sidx = 0; sidx < kSplit; ++sidx) {              // fixed order: deterministic
      const float4 q = parts[(size_t)sidx * Mcap + r];
      if (q.x != -INFINITY) merge_stats(m, p, a, q.x, q.y, q.z);
    }
    const float lr = -logf(p / a) * __ldg(row_w + r);        // sim_loss.py:76-78
    stats[r] = make_float4(m, p, a, lr);
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

__global__ void __launch_bounds__(kThreads, 2)
supcon_bwd_kernel(const float* __restrict__ F, const float* __restrict__ E, int R,
                  const int32_t* __restrict__ row_src, const int32_t* __restrict__ row_lab,
                  const float* __restrict__ row_w, const int32_t* __restrict__ M_dev, int Mcap, float inv_temp,
                  const float4* __restrict__ stats, const float* __restrict__ gscale_dev, float* __restrict__ dF,
                  float* __restrict__ dE) {
  extern __shared__ __align__(16) float smem_f[];
  float* sA = smem_f;                      // [128][68] k-major rows of this CTA
  float* sB = sA + kD * kLd;               // [128][68] k-major column tile
  float* sH = sB + kD * kLd;               // [64][65]  H tile
  __shared__ float c_m[kT], c_ip[kT], c_ia[kT], c_cf[kT];
  __shared__ int c_lab[kT];
  const int M = min(*M_dev, Mcap);
  const int r0 = blockIdx.x * kT;
  if (r0 >= M) return;
  const int per_split = ((M + kT - 1) / kT + kSplit - 1) / kSplit * kT;
  const int c_begin = blockIdx.y * per_split, c_end = min(M, c_begin + per_split);
  if (c_begin >= c_end) return;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const float gM = __ldg(gscale_dev) / (float)M;
  load_tile_kmajor(sA, F, E, R, row_src, r0, M);
  float r_m[4], r_ip[4], r_ia[4], r_cf[4];
  int r_lab[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + ty * 4 + u;
    if (r < M) {
      const float4 st = stats[r];
      r_m[u] = st.x; r_ip[u] = 1.f / st.y; r_ia[u] = 1.f / st.z;
      r_cf[u] = gM * __ldg(row_w + r); r_lab[u] = __ldg(row_lab + r);
    } else { r_m[u] = 0.f; r_ip[u] = 0.f; r_ia[u] = 0.f; r_cf[u] = 0.f; r_lab[u] = -1; }
  }
  float acc2[4][8];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc2[u][e] = 0.f;

  for (int c0 = c_begin; c0 < c_end; c0 += kT) {
    __syncthreads();
    load_tile_kmajor(sB, F, E, R, row_src, c0, M);
    if (threadIdx.x < kT) {
      const int c = c0 + threadIdx.x;
      if (c < M) {
        const float4 st = stats[c];
        c_m[threadIdx.x] = st.x; c_ip[threadIdx.x] = 1.f / st.y; c_ia[threadIdx.x] = 1.f / st.z;
        c_cf[threadIdx.x] = gM * __ldg(row_w + c); c_lab[threadIdx.x] = __ldg(row_lab + c);
      } else { c_m[threadIdx.x] = 0.f; c_ip[threadIdx.x] = 0.f; c_ia[threadIdx.x] = 0.f; c_cf[threadIdx.x] = 0.f; c_lab[threadIdx.x] = -2; }
    }
    __syncthreads();
    float acc[4][4];
    tile_product(sA, sB, ty, tx, acc);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + ty * 4 + u;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int cl = tx * 4 + v, c = c0 + cl;
        float h = 0.f;
        if (r < M && c < M && r != c) {
          const float s = acc[u][v] * inv_temp;
          const bool same = (c_lab[cl] == r_lab[u]);
          const float g_rq = r_cf[u] * expf(s - r_m[u]) * (r_ia[u] - (same ? r_ip[u] : 0.f));
          const float g_qr = c_cf[cl] * expf(s - c_m[cl]) * (c_ia[cl] - (same ? c_ip[cl] : 0.f));
          h = g_rq + g_qr;
        }
        sH[(ty * 4 + u) * (kT + 1) + cl] = h;
      }
    }
    __syncthreads();
    // acc2[u][e] += sum_q H[row][q] * V_q[d],  d = tx + 16 e
#pragma unroll 4
    for (int q = 0; q < kT; ++q) {
      float hv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) hv[u] = sH[(ty * 4 + u) * (kT + 1) + q];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float b = sB[(tx + 16 * e) * kLd + q];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc2[u][e] = fmaf(hv[u], b, acc2[u][e]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + ty * 4 + u;
    if (r >= M) continue;
    const int src = __ldg(row_src + r);
    float* dst = src < R ? dF + (size_t)src * kD : dE + (size_t)(src - R) * kD;
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dst + tx + 16 * e, acc2[u][e] * inv_temp);
  }
}

}  // namespace

ODW_API int odwscl_supcon_fwd_f32(const float* F, const float* E, int R, const int32_t* row_src,
                                  const int32_t* row_lab, const float* row_w, const int32_t* M_dev, int Mcap,
                                  float inv_temp, float* stats, float* loss_out, odwscl_stream_t stream) {
  if (R < 0 || Mcap < 0) return ODWSCL_EINVAL;
  if (!loss_out) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (Mcap == 0) { ODW_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st)); return 0; }
  if (!F || !row_src || !row_lab || !row_w || !M_dev || !stats) return ODWSCL_EINVAL;
  const int smem = 2 * kD * kLd * (int)sizeof(float);
  ODW_CUDA(cudaFuncSetAttribute(supcon_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float4* parts = reinterpret_cast<float4*>(stats) + Mcap;              // [kSplit][Mcap] behind the merged rows
  supcon_fwd_kernel<<<dim3(odw_cdiv(Mcap, kT), kSplit), kThreads, smem, st>>>(F, E, R, row_src, row_lab, row_w, M_dev,
                                                                             Mcap, inv_temp, parts);
  ODW_LAUNCH_CHECK();
  supcon_mean_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<float4*>(stats), parts, row_w, M_dev, Mcap, loss_out);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_supcon_bwd_f32(const float* F, const float* E, int R, const int32_t* row_src,
                                  const int32_t* row_lab, const float* row_w, const int32_t* M_dev, int Mcap,
                                  float inv_temp, const float* stats, const float* gscale_dev, float* dF,
                                  float* dE, odwscl_stream_t stream) {
  if (R < 0 || Mcap < 0) return ODWSCL_EINVAL;
  if (Mcap == 0) return 0;
  if (!F || !row_src || !row_lab || !row_w || !M_dev || !stats || !gscale_dev || !dF) return ODWSCL_EINVAL;
  const int smem = (2 * kD * kLd + kT * (kT + 1)) * (int)sizeof(float);
  ODW_CUDA(cudaFuncSetAttribute(supcon_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  supcon_bwd_kernel<<<dim3(odw_cdiv(Mcap, kT), kSplit), kThreads, smem, (cudaStream_t)stream>>>(
      F, E, R, row_src, row_lab, row_w, M_dev, Mcap, inv_temp, reinterpret_cast<const float4*>(stats), gscale_dev,
      dF, dE);
  ODW_LAUNCH_CHECK();
  return 0;
}
