// gemm_tf32.cu -- C[M,N] = A[M,K] * B[N,K]^T on the 5th-generation tensor cores (tcgen05.mma
// kind::tf32, fp32 accumulators in TMEM), operands streamed by TMA through an mbarrier ring.
//
// Used by odwscl_sim_nxn_f32: the dense N x N proposal-embedding similarity F F^T that
// roi_heads/weak_head/loss.py:319 computes with torch.mm.  To keep the result fp32-accurate (the
// reference's sgemm is fp32; its `Sim >= tau` selections are ulp-sensitive) the
// product is evaluated as a 3xTF32 split: x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi),
// F F^T ~= hi hi^T + hi lo^T + lo hi^T, expressed as ONE K = 384 GEMM over the concatenated
// operands A' = [hi | hi | lo], B' = [hi | lo | hi] (the dropped lo lo^T term is <= 2^-22 relative).
//
// Kernel anatomy (one 128 x BN output tile per CTA, 192 threads, 1 CTA / SM):
//   warp 0     TMA producer: one elected lane waits on empty[s], arms full[s] with the stage's byte
//              count and issues the A and B tile loads (128-byte swizzled rows).
//   warp 1     allocates TMEM (BN fp32 columns x 128 lanes); one elected lane issues 4 tcgen05.mma
//              (K = 8 each) per 32-wide K tile and commits to empty[s]; after the last K tile it
//              commits to tmem_full.
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> 128-byte row segments.
#include "common.cuh"
#include "tc_sm100.cuh"

namespace {

constexpr int kBM = 128;

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 float* __restrict__ C, int M, int N, int K, int ldc) {
  constexpr int A_BYTES = kBM * tc::kTileKBytes, B_BYTES = BN * tc::kTileKBytes, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  const int kiters = (K + tc::kTileK - 1) / tc::kTileK;

  if (warp == 0 && tc::elect_one()) {
    tc::tma_prefetch_desc(&map_a);
    tc::tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tmem_full_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, BN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (tc::elect_one()) {
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* a = tiles + (size_t)s * STAGE_BYTES;
        tc::tma_load_2d(a, &map_a, &full_bar[s], it * tc::kTileK, m0);
        tc::tma_load_2d(a + A_BYTES, &map_b, &full_bar[s], it * tc::kTileK, n0);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::umma_idesc_tf32(kBM, BN);
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(&full_bar[s], (it / STAGES) & 1);
        tc::tc_fence_after();
        const uint32_t a = tc::smem_u32(tiles + (size_t)s * STAGE_BYTES);
        const uint64_t ad = tc::umma_desc_sw128(a), bd = tc::umma_desc_sw128(a + A_BYTES);
#pragma unroll
        for (int k = 0; k < tc::kTileK / tc::kUmmaK; ++k)      // +32 bytes per K step: +2 in the >>4 address field
          tc::umma_tf32(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (it | k) != 0);
        tc::umma_commit(&empty_bar[s]);
      }
      tc::umma_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    tc::mbar_wait(&tmem_full_bar, 0);
    tc::tc_fence_after();
    const int row = m0 + q * 32 + lane;
    float v[32];
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
      tc::tmem_ld_wait();
      const int col = n0 + c * 32;
      if (row < M && col < N) {
        float* dst = C + (size_t)row * ldc + col;
        if (col + 32 <= N && (ldc & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcs(reinterpret_cast<float4*>(dst) + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        } else {
          for (int j = 0; j < 32 && col + j < N; ++j) dst[j] = v[j];
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, BN);
}

// A' = [hi | hi | lo], B' = [hi | lo | hi], rows of 3*D floats
__global__ void split3_kernel(const float* __restrict__ F, int n_elems, int D, float* __restrict__ A3, float* __restrict__ B3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_elems) return;
  const int r = i / D, c = i - r * D;
  const float x = F[i];
  uint32_t hb, lb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
  const float hi = __uint_as_float(hb);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(__fsub_rn(x, hi)));
  const float lo = __uint_as_float(lb);
  const size_t o = (size_t)r * 3 * D + c;
  A3[o] = hi; A3[o + D] = hi; A3[o + 2 * D] = lo;
  B3[o] = hi; B3[o + D] = lo; B3[o + 2 * D] = hi;
}

template <int BN>
int launch_gemm(const float* A, const float* B, float* C, int M, int N, int K, int ldc, cudaStream_t st) {
  constexpr int STAGES = 4;
  CUtensorMap ma, mb;
  const uint64_t da[2] = {(uint64_t)K, (uint64_t)M}, db[2] = {(uint64_t)K, (uint64_t)N};
  const uint64_t sa[1] = {(uint64_t)K * 4};
  const uint32_t ba[2] = {32, kBM}, bb[2] = {32, BN};
  int rc = tc::make_tmap_f32(&ma, A, 2, da, sa, ba);
  if (rc) return rc;
  rc = tc::make_tmap_f32(&mb, B, 2, db, sa, bb);
  if (rc) return rc;
  const int smem = STAGES * (kBM + BN) * tc::kTileKBytes + 1024;
  ODW_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(odw_cdiv(M, kBM), odw_cdiv(N, BN));
  gemm_tf32_kernel<BN, STAGES><<<grid, 192, smem, st>>>(ma, mb, C, M, N, K, ldc);
  ODW_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// C[M,N] = A[M,K] * B[N,K]^T, single-pass TF32 (inputs truncated to 10 mantissa bits by the tensor core).
// K % 4 == 0 (16-byte row pitch for TMA); A, B 16-byte aligned.
ODW_API int odwscl_gemm_nt_tf32(const float* A, const float* B, float* C, int M, int N, int K, int ldc,
                                odwscl_stream_t stream) {
  if (M < 0 || N < 0 || K < 0 || (K & 3) || ldc < N) return ODWSCL_EINVAL;
  if (M == 0 || N == 0) return 0;
  if (!A || !B || !C || K == 0) return ODWSCL_EINVAL;
  return N > 128 ? launch_gemm<256>(A, B, C, M, N, K, ldc, (cudaStream_t)stream)
                 : launch_gemm<128>(A, B, C, M, N, K, ldc, (cudaStream_t)stream);
}

ODW_API size_t odwscl_sim_nxn_ws_bytes(int N) {
  return N > 0 ? 2 * odw_align((size_t)N * 3 * ODWSCL_SIM_DIM * sizeof(float)) : 0;
}

ODW_API int odwscl_sim_nxn_f32(const float* F, int N, float* out, void* ws, size_t ws_bytes, odwscl_stream_t stream) {
  if (N < 0) return ODWSCL_EINVAL;
  if (N == 0) return 0;
  if (!F || !out) return ODWSCL_EINVAL;
  if (!ws || ws_bytes < odwscl_sim_nxn_ws_bytes(N)) return ODWSCL_ENOWS;
  cudaStream_t st = (cudaStream_t)stream;
  float* A3 = reinterpret_cast<float*>(ws);
  float* B3 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + odw_align((size_t)N * 3 * ODWSCL_SIM_DIM * sizeof(float)));
  const int n = N * ODWSCL_SIM_DIM;
  split3_kernel<<<odw_cdiv(n, 256), 256, 0, st>>>(F, n, ODWSCL_SIM_DIM, A3, B3);
  ODW_LAUNCH_CHECK();
  return odwscl_gemm_nt_tf32(A3, B3, out, N, N, 3 * ODWSCL_SIM_DIM, N, stream);
}
