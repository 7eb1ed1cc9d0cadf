// sim_nxn.cu -- full N x N proposal-embedding similarity F F^T (drop-in for the torch.mm of
// roi_heads/weak_head/loss.py:319).  The product path never needs the full matrix (discover.cu
// computes only the rows the selection rule reads); this entry point exists for callers that keep
// the reference's dense formulation.  fp32 SIMT register-tiled version (64x64 tile, 4x4 per thread,
// k-major shared-memory operands).
#include "common.cuh"

namespace {
constexpr int kD = ODWSCL_SIM_DIM, kT = 64, kLd = kT + 4;

__device__ __forceinline__ void load_rows_kmajor(float* sT, const float* __restrict__ F, int r0, int N) {
  for (int t = threadIdx.x; t < kT * (kD / 4); t += blockDim.x) {
    const int r = t / (kD / 4), k4 = t % (kD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < N) v = __ldg(reinterpret_cast<const float4*>(F + (size_t)(r0 + r) * kD) + k4);
    sT[(4 * k4 + 0) * kLd + r] = v.x; sT[(4 * k4 + 1) * kLd + r] = v.y;
    sT[(4 * k4 + 2) * kLd + r] = v.z; sT[(4 * k4 + 3) * kLd + r] = v.w;
  }
}

__global__ void __launch_bounds__(256, 2)
sim_nxn_kernel(const float* __restrict__ F, int N, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* sA = sm; float* sB = sm + kD * kLd;
  const int r0 = blockIdx.y * kT, c0 = blockIdx.x * kT;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  load_rows_kmajor(sA, F, r0, N);
  load_rows_kmajor(sB, F, c0, N);
  __syncthreads();
  float acc[4][4] = {};
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
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + ty * 4 + u;
    if (r >= N) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int c = c0 + tx * 4 + v;
      if (c < N) out[(size_t)r * N + c] = acc[u][v];
    }
  }
}
}  // namespace

ODW_API int odwscl_sim_nxn_f32(const float* F, int N, float* out, odwscl_stream_t stream) {
  if (N < 0) return ODWSCL_EINVAL;
  if (N == 0) return 0;
  if (!F || !out) return ODWSCL_EINVAL;
  const int smem = 2 * kD * kLd * (int)sizeof(float);
  ODW_CUDA(cudaFuncSetAttribute(sim_nxn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(odw_cdiv(N, kT), odw_cdiv(N, kT));
  sim_nxn_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(F, N, out);
  ODW_LAUNCH_CHECK();
  return 0;
}
