// conv3x3.cu -- the VGG16-OICR conv stack (modeling/backbone/vgg16.py:26-36,58-83)
// as hand-written sm_100a kernels, channels-last (NHWC) end to end.
//
//   conv3x3_tf32_kernel   3x3 / stride 1 / "same" padding / dilation 1 or 2 convolution as an IMPLICIT
//                         GEMM on the tensor cores:  Y[pixel, co] = sum_{tap, ci} X[pixel + tap, ci] W[co, tap, ci]
//                         M = 128 output pixels (a TH x TW patch of one image), N = BN output channels,
//                         K = 9 taps x Cin.  For every (tap, 32-channel slab) the producer issues ONE 4-D
//                         TMA box load of the NHWC input shifted by the tap offset -- out-of-image rows /
//                         columns are zero-filled by the TMA unit, which IS the conv padding -- plus one
//                         2-D box of the K-major weight matrix; tcgen05.mma.kind::tf32 accumulates in TMEM;
//                         the epilogue fuses bias, ReLU, accumulate-into-output (3xTF32 strict mode) and a
//                         ReLU-derivative mask (so the same kernel is the DGRAD: dX = conv(dY, W flipped/transposed),
//                         masked by the previous activation).
//   conv3x3_c3_kernel     conv1_1 (Cin = 3, K = 27: not tensor-core shaped, output-bandwidth bound): fp32 FFMA,
//                         reads the NCHW image, writes NHWC.
//   maxpool2x2_*          2x2 / stride 2 max-pool forward and backward (first-max routing as torch).
//   split_tf32            x -> (hi, lo) for the 3-pass strict-fp32 mode used by the parity tests.
//
// Reference arithmetic: cuDNN through torch.nn.Conv2d; TF32 is what torch 1.7.1 (the reference's pin,
// README.md:23) runs by default on tensor-core GPUs.  Single-pass mode = TF32 inputs / fp32 accumulate.
#include <stdlib.h>

#include "common.cuh"
#include "tc_sm100.cuh"

namespace {

constexpr int kBM = 128;
enum : int { kRelu = 1, kAccum = 2, kMask = 4, kRound = 8 };

__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t b;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v));
  return __uint_as_float(b);
}

// MT = 128-pixel tiles per CTA (1 or 2).  With MT = 2 the CTA keeps two accumulators in TMEM and feeds both from the
// SAME weight tile in shared memory: the operand bytes an SM has to ingest per FLOP drop by a third for BN = 256
// (64 KB per 2x work instead of 48 KB), which is what bounds the Cout >= 256 layers (measured ~58 B/clk/SM).
// CL = thread-block cluster size along the pixel-tile axis (1 or 2): the CL CTAs of a cluster load 1/CL of the
// weight tile each and TMA-multicast it to all of them.  Measured neutral-to-slower on B200 (the limiter is the
// per-SM ingest rate, not aggregate L2 bandwidth), so it is off unless ODWSCL_CONV_CLUSTER=2.
template <int BN, int STAGES, int CL, int MT>
__global__ void __launch_bounds__(192, (STAGES * (MT * kBM + BN) * 128 + 2048 <= 110 * 1024) ? 2 : 1)
conv3x3_tf32_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const float* __restrict__ bias, const float* __restrict__ mask_src, float* __restrict__ y,
                    int H, int W, int Cin, int Cout, int tw_log2, int tiles_w, int tiles_h, int total_tiles, int dil,
                    int flags) {
  constexpr int A_TILE = kBM * tc::kTileKBytes, A_BYTES = MT * A_TILE, B_BYTES = BN * tc::kTileKBytes;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static_assert(MT * BN <= 512, "accumulators exceed the 512 TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int TW = 1 << tw_log2, TH = kBM >> tw_log2;
  int tb[MT], th0[MT], tw0[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    int t = blockIdx.x * MT + m;               // tiles past the end decode to b >= B: TMA zero-fills, epilogue skips
    const int tx = t % tiles_w; t /= tiles_w;
    const int ty = t % tiles_h;
    tb[m] = t / tiles_h;
    tw0[m] = tx * TW;
    th0[m] = ty * TH;
  }
  const int n0 = blockIdx.y * BN;
  const int cchunks = Cin / tc::kTileK;
  const int kiters = 9 * cchunks;

  if (warp == 0 && tc::elect_one()) {
    tc::tma_prefetch_desc(&map_x);
    tc::tma_prefetch_desc(&map_w);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], CL); }
    tc::mbar_init(&tmem_full_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, MT * BN);
  tc::tc_fence_before();
  __syncthreads();
  if (CL > 1) tc::cluster_sync_all();          // peers' barriers are initialised before anything is multicast at them
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t crank = CL > 1 ? tc::cluster_ctarank() : 0;
  constexpr uint16_t kCtaMask = (uint16_t)((1u << CL) - 1);

  if (warp == 0) {
    if (tc::elect_one()) {
      int tap = 0, cc = 0;
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* a = tiles + (size_t)s * STAGE_BYTES;
        const int r = tap / 3, q = tap - 3 * r;
        // input patch shifted by the tap; rows / columns outside the image arrive as zeros (= padding)
#pragma unroll
        for (int m = 0; m < MT; ++m)
          tc::tma_load_4d(a + m * A_TILE, &map_x, &full_bar[s], cc * tc::kTileK, tw0[m] + (q - 1) * dil,
                          th0[m] + (r - 1) * dil, tb[m]);
        if (CL == 1)
          tc::tma_load_2d(a + A_BYTES, &map_w, &full_bar[s], tap * Cin + cc * tc::kTileK, n0);
        else      // my 1/CL of the weight rows, delivered to every CTA of the cluster (each expects the whole tile)
          tc::tma_load_2d_mc(a + A_BYTES + crank * (B_BYTES / CL), &map_w, &full_bar[s], tap * Cin + cc * tc::kTileK,
                             n0 + (int)crank * (BN / CL), kCtaMask);
        if (++cc == cchunks) { cc = 0; ++tap; }
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
        const uint64_t bd = tc::umma_desc_sw128(a + A_BYTES);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint64_t ad = tc::umma_desc_sw128(a + m * A_TILE);
#pragma unroll
          for (int k = 0; k < tc::kTileK / tc::kUmmaK; ++k)
            tc::umma_tf32(tmem_base + m * BN, ad + 2 * k, bd + 2 * k, idesc, (it | k) != 0);
        }
        if (CL == 1) tc::umma_commit(&empty_bar[s]);
        else tc::umma_commit_mc(&empty_bar[s], kCtaMask);  // the stage is refilled by multicast: free it in every CTA
      }
      tc::umma_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    tc::mbar_wait(&tmem_full_bar, 0);
    tc::tc_fence_after();
    float v[32];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int h = th0[m] + (row >> tw_log2), w = tw0[m] + (row & (TW - 1));
      const bool valid = (h < H) && (w < W) && (blockIdx.x * MT + m < total_tiles);
      const size_t pix = ((size_t)tb[m] * H + h) * W + w;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + m * BN + c * 32, v);
        tc::tmem_ld_wait();
        const int co = n0 + c * 32;
        if (!valid || co >= Cout) continue;
        float4* dst = reinterpret_cast<float4*>(y + pix * Cout + co);
        const float4* msk = reinterpret_cast<const float4*>(mask_src + pix * Cout + co);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (bias != nullptr) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + co) + j);
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          if (flags & kAccum) {
            const float4 p = dst[j];
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          if (flags & kRelu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (flags & kMask) {
            const float4 mm = __ldg(msk + j);
            o.x = mm.x > 0.f ? o.x : 0.f; o.y = mm.y > 0.f ? o.y : 0.f; o.z = mm.z > 0.f ? o.z : 0.f; o.w = mm.w > 0.f ? o.w : 0.f;
          }
          if (flags & kRound) { o.x = rna_tf32(o.x); o.y = rna_tf32(o.y); o.z = rna_tf32(o.z); o.w = rna_tf32(o.w); }
          dst[j] = o;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CL > 1) tc::cluster_sync_all();          // no CTA leaves while a peer can still arrive on its barriers
  if (warp == 1) tc::tmem_dealloc(tmem_base, MT * BN);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair version for Cout % 256 == 0 (tcgen05 cta_group::2).  A single-CTA 128 x 256 TF32 tile is bound by the
// SM's shared-memory bandwidth: per K = 8 step the tensor core reads 12 KB of operands and TMA writes the same 12 KB,
// 176 B/clk against the 128 B/clk an SM has (measured: tensor pipe 73 % busy = 128/176).  Two CTAs of a cluster own
// adjacent pixel tiles and ONE 256-channel weight tile: each stages its own A tile and HALF of the B tile, the leader
// CTA issues M = 256 MMAs that read both halves, and each SM's shared-memory traffic drops to 115 B/clk.
template <int STAGES>
__global__ void __launch_bounds__(192, 1)
conv3x3_tf32_2cta_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                         const float* __restrict__ bias, const float* __restrict__ mask_src, float* __restrict__ y,
                         int H, int W, int Cin, int Cout, int tw_log2, int tiles_w, int tiles_h, int dil, int flags) {
  constexpr int BN = 256, A_BYTES = kBM * tc::kTileKBytes, B_BYTES = (BN / 2) * tc::kTileKBytes;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = tc::cluster_ctarank();
  const int TW = 1 << tw_log2, TH = kBM >> tw_log2;
  int t = blockIdx.x;
  const int tx = t % tiles_w; t /= tiles_w;
  const int ty = t % tiles_h;
  const int b = t / tiles_h;
  const int w0 = tx * TW, h0 = ty * TH;
  const int n0 = blockIdx.y * BN;
  const int cchunks = Cin / tc::kTileK;
  const int kiters = 9 * cchunks;

  if (warp == 0 && tc::elect_one()) {
    tc::tma_prefetch_desc(&map_x);
    tc::tma_prefetch_desc(&map_w);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tmem_full_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_2sm(&tmem_base_s, BN);
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();                      // both CTAs' barriers exist before any TMA / commit targets them
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (tc::elect_one()) {
      int tap = 0, cc = 0;
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        // the leader's barrier collects the bytes of BOTH CTAs (each TMA below signals it through the peer-bit mask)
        if (crank == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
        uint8_t* a = tiles + (size_t)s * STAGE_BYTES;
        const int r = tap / 3, q = tap - 3 * r;
        tc::tma_load_4d_2sm(a, &map_x, &full_bar[s], cc * tc::kTileK, w0 + (q - 1) * dil, h0 + (r - 1) * dil, b);
        tc::tma_load_2d_2sm(a + A_BYTES, &map_w, &full_bar[s], tap * Cin + cc * tc::kTileK, n0 + (int)crank * (BN / 2));
        if (++cc == cchunks) { cc = 0; ++tap; }
      }
    }
  } else if (warp == 1) {
    if (crank == 0 && tc::elect_one()) {
      constexpr uint32_t idesc = tc::umma_idesc_tf32(2 * kBM, BN);
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(&full_bar[s], (it / STAGES) & 1);
        tc::tc_fence_after();
        const uint32_t a = tc::smem_u32(tiles + (size_t)s * STAGE_BYTES);
        const uint64_t ad = tc::umma_desc_sw128(a), bd = tc::umma_desc_sw128(a + A_BYTES);
#pragma unroll
        for (int k = 0; k < tc::kTileK / tc::kUmmaK; ++k)
          tc::umma_tf32_2sm(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (it | k) != 0);
        tc::umma_commit_2sm_mc(&empty_bar[s], 3);          // the stage is free in both CTAs
      }
      tc::umma_commit_2sm_mc(&tmem_full_bar, 3);           // both CTAs' epilogues may drain their accumulator rows
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int h = h0 + (row >> tw_log2), w = w0 + (row & (TW - 1));
    const bool valid = (h < H) && (w < W);
    const size_t pix = ((size_t)b * H + h) * W + w;
    tc::mbar_wait(&tmem_full_bar, 0);
    tc::tc_fence_after();
    float v[32];
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
      tc::tmem_ld_wait();
      const int co = n0 + c * 32;
      if (!valid || co >= Cout) continue;
      float4* dst = reinterpret_cast<float4*>(y + pix * Cout + co);
      const float4* msk = reinterpret_cast<const float4*>(mask_src + pix * Cout + co);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        if (bias != nullptr) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + co) + j);
          o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        }
        if (flags & kAccum) {
          const float4 p = dst[j];
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        if (flags & kRelu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (flags & kMask) {
          const float4 mm = __ldg(msk + j);
          o.x = mm.x > 0.f ? o.x : 0.f; o.y = mm.y > 0.f ? o.y : 0.f; o.z = mm.z > 0.f ? o.z : 0.f; o.w = mm.w > 0.f ? o.w : 0.f;
        }
        if (flags & kRound) { o.x = rna_tf32(o.x); o.y = rna_tf32(o.y); o.z = rna_tf32(o.z); o.w = rna_tf32(o.w); }
        dst[j] = o;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();                      // the peer's MMAs read this CTA's shared memory until the very end
  if (warp == 1) tc::tmem_dealloc_2sm(tmem_base, BN);
}

template <int STAGES>
static int launch_conv_2cta(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, const float* msk, float* y,
                            int H, int W, int Cin, int Cout, int best_log2, int tiles_w, int tiles_h, int total_tiles,
                            int dil, int flags, cudaStream_t st) {
  const int smem = STAGES * (kBM + 128) * tc::kTileKBytes + 1024;
  auto kern = conv3x3_tf32_2cta_kernel<STAGES>;
  ODW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(total_tiles, Cout / 256);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ODW_CUDA(cudaLaunchKernelEx(&cfg, kern, mx, mw, bias, msk, y, H, W, Cin, Cout, best_log2, tiles_w, tiles_h, dil, flags));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// PERSISTENT CTA-pair version.  The one-tile-per-pair kernel above loses two things the profile shows: the epilogue /
// pipeline fill of each tile is not overlapped with the next tile's MMAs, and 304 tiles on 148 SMs run as 3 waves
// instead of 2.05.  Here NP resident pairs each take floor(T/NP) WHOLE pair-tiles, round by round and all at the same
// k at the same time (so the 37+ pairs that need the same weight rows ask L2 for them together -- measured: 740 TF/s
// when the tile count divides evenly vs 614 TF/s with skewed k ranges), and the T mod NP leftover tiles are split
// along K over all the pairs (S = NP / leftover slices per tile).  The S partial accumulators of a leftover tile go
// to a global workspace; the S CTAs then reduce 16-column chunks of it in a fixed order and run the fused epilogue
// (all pairs are co-resident, so waiting on the arrival counter cannot deadlock).  Accumulators are double-buffered
// in TMEM (2 x 256 columns): the epilogue warps drain tile i while the MMA thread already accumulates tile i+1.
struct ConvSkParams {
  int n_ptiles;        // T: pair-tiles = (pixel tiles / 2) * (Cout / 256)
  int n_ntiles;        // Cout / 256
  int kiters;          // 9 * Cin / 32
  int np;              // resident pairs
  int rounds;          // T / np whole tiles per pair
  int left;            // T - rounds * np leftover tiles
  int slices;          // K slices per leftover tile (np / left), 0 if none
  float* ws;           // [left * slices][2][128][256] partial accumulators
  int* flags;          // [left][2] arrival counters, zeroed before the launch
};

struct ConvSkItem { int tile, k0, k1, left_idx, slice; };   // left_idx < 0: a whole tile

// Work list of one pair: its K slice of a leftover tile FIRST (the partial is published early and reduced at the very
// end, when every other slice has long arrived -- nobody waits), then the whole tiles round by round.
__device__ __forceinline__ bool conv_sk_item(const ConvSkParams& sk, int pair, int idx, ConvSkItem& it) {
  const bool has_tail = sk.left > 0 && pair < sk.left * sk.slices;
  if (has_tail) {
    if (idx == 0) {
      it.left_idx = pair / sk.slices; it.slice = pair - it.left_idx * sk.slices;
      it.tile = sk.rounds * sk.np + it.left_idx;
      it.k0 = (int)((long long)it.slice * sk.kiters / sk.slices);
      it.k1 = (int)((long long)(it.slice + 1) * sk.kiters / sk.slices);
      return true;
    }
    --idx;
  }
  if (idx < sk.rounds) { it.tile = idx * sk.np + pair; it.k0 = 0; it.k1 = sk.kiters; it.left_idx = -1; it.slice = 0; return true; }
  return false;
}

// HALO: a k-iteration is (tap ROW r, 32-channel slab) and stages ONE halo row of 128 + 2*dil pixels plus the three weight
// tiles of the row's taps; the three column taps read the row through descriptors shifted by q*dil rows of 128 bytes
// (the UMMA swizzle is a function of the absolute shared-memory address, so a window sliding by whole rows over the
// TMA-written image needs no base-offset field -- verified exact).  The A operand is then staged once per 3 taps:
// operand bytes per SM drop from 32 KB to 21.6 KB per tap.  Needs 128 x 1 pixel tiles (tw_log2 == 7).  Measured neutral
// (752 vs 752 TF/s on an evenly dividing shape, 659 vs 679 at 76 rows): operand staging is no longer the limiter of the
// pair kernel, so it stays off unless ODWSCL_CONV_HALO=1.
template <int STAGES, bool HALO, int BN>
__global__ void __launch_bounds__(192, 1)
conv3x3_tf32_2cta_sk_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                            const float* __restrict__ bias, const float* __restrict__ mask_src, float* __restrict__ y,
                            int H, int W, int Cin, int Cout, int tw_log2, int tiles_w, int tiles_h, int dil, int flags,
                            const ConvSkParams sk) {
  constexpr int B_TILE = (BN / 2) * tc::kTileKBytes;
  constexpr int TM_COLS = BN < 32 ? 32 : BN;                  // TMEM columns per accumulator buffer (allocation granule)
  constexpr int A_BYTES = HALO ? 17 * 1024 : kBM * tc::kTileKBytes, B_BYTES = (HALO ? 3 : 1) * B_TILE;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t stage_tx = HALO ? (uint32_t)(128 + 2 * dil) * 128u + B_BYTES : (uint32_t)STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = tc::cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int TW = 1 << tw_log2, TH = kBM >> tw_log2;
  const int cchunks = Cin / tc::kTileK;

  if (warp == 0 && tc::elect_one()) {
    tc::tma_prefetch_desc(&map_x);
    tc::tma_prefetch_desc(&map_w);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tmem_full_bar[i], 1); tc::mbar_init(&tmem_empty_bar[i], 256); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_2sm(&tmem_base_s, 2 * TM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // decode of a pair-tile: n-tile fastest, then the pair of adjacent pixel tiles (this CTA takes 2*m + rank)
  auto tile_coords = [&](int pt, int& b, int& h0, int& w0, int& n0) {
    n0 = (pt % sk.n_ntiles) * BN;
    int t = (pt / sk.n_ntiles) * 2 + (int)crank;
    const int tx = t % tiles_w; t /= tiles_w;
    const int ty = t % tiles_h;
    b = t / tiles_h;
    w0 = tx * TW; h0 = ty * TH;
  };

  if (warp == 0) {
    if (tc::elect_one()) {
      int it = 0;
      ConvSkItem wi;
      for (int item = 0; conv_sk_item(sk, pair, item, wi); ++item) {
        int b, h0, w0, n0;
        tile_coords(wi.tile, b, h0, w0, n0);
        int tap = wi.k0 / cchunks, cc = wi.k0 - tap * cchunks;      // HALO: `tap` counts tap ROWS
        for (int kk = wi.k0; kk < wi.k1; ++kk, ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          if (crank == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * stage_tx);
          uint8_t* a = tiles + (size_t)s * STAGE_BYTES;
          if (HALO) {
            tc::tma_load_4d_2sm(a, &map_x, &full_bar[s], cc * tc::kTileK, w0 - dil, h0 + (tap - 1) * dil, b);
#pragma unroll
            for (int q = 0; q < 3; ++q)
              tc::tma_load_2d_2sm(a + A_BYTES + q * B_TILE, &map_w, &full_bar[s], (tap * 3 + q) * Cin + cc * tc::kTileK,
                                  n0 + (int)crank * (BN / 2));
          } else {
            const int r = tap / 3, q = tap - 3 * r;
            tc::tma_load_4d_2sm(a, &map_x, &full_bar[s], cc * tc::kTileK, w0 + (q - 1) * dil, h0 + (r - 1) * dil, b);
            tc::tma_load_2d_2sm(a + A_BYTES, &map_w, &full_bar[s], tap * Cin + cc * tc::kTileK, n0 + (int)crank * (BN / 2));
          }
          if (++cc == cchunks) { cc = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0 && tc::elect_one()) {
      constexpr uint32_t idesc = tc::umma_idesc_tf32(2 * kBM, BN);
      int it = 0;
      ConvSkItem wi;
      for (int item = 0; conv_sk_item(sk, pair, item, wi); ++item) {
        const int buf = item & 1;
        tc::mbar_wait(&tmem_empty_bar[buf], ((item >> 1) & 1) ^ 1);     // both CTAs drained this accumulator
        tc::tc_fence_after();
        const uint32_t acc = tmem_base + buf * TM_COLS;
        for (int kk = wi.k0; kk < wi.k1; ++kk, ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t a = tc::smem_u32(tiles + (size_t)s * STAGE_BYTES);
          if (HALO) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const uint64_t ad = tc::umma_desc_sw128(a + (uint32_t)(q * dil) * 128u);
              const uint64_t bd = tc::umma_desc_sw128(a + A_BYTES + q * B_TILE);
#pragma unroll
              for (int k = 0; k < tc::kTileK / tc::kUmmaK; ++k)
                tc::umma_tf32_2sm(acc, ad + 2 * k, bd + 2 * k, idesc, (kk != wi.k0) || (q | k) != 0);
            }
          } else {
            const uint64_t ad = tc::umma_desc_sw128(a), bd = tc::umma_desc_sw128(a + A_BYTES);
#pragma unroll
            for (int k = 0; k < tc::kTileK / tc::kUmmaK; ++k)
              tc::umma_tf32_2sm(acc, ad + 2 * k, bd + 2 * k, idesc, (kk != wi.k0) || (k != 0));
          }
          tc::umma_commit_2sm_mc(&empty_bar[s], 3);
        }
        tc::umma_commit_2sm_mc(&tmem_full_bar[buf], 3);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float v[32];
    ConvSkItem wi;
    for (int item = 0; conv_sk_item(sk, pair, item, wi); ++item) {
      const int buf = item & 1;
      int b, h0, w0, n0;
      tile_coords(wi.tile, b, h0, w0, n0);
      const int h = h0 + (row >> tw_log2), w = w0 + (row & (TW - 1));
      const bool valid = (h < H) && (w < W);
      const size_t pix = ((size_t)b * H + h) * W + w;
      tc::mbar_wait(&tmem_full_bar[buf], (item >> 1) & 1);
      tc::tc_fence_after();
      if (wi.left_idx < 0) {
        // ---- whole tile: fused epilogue straight from TMEM
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TM_COLS + c * 32, v);
          tc::tmem_ld_wait();
          const int co = n0 + c * 32;
          if (!valid || co >= Cout) continue;
          float4* dst = reinterpret_cast<float4*>(y + pix * Cout + co);
          const float4* msk = reinterpret_cast<const float4*>(mask_src + pix * Cout + co);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (bias != nullptr) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + co) + j);
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            if (flags & kAccum) {
              const float4 p = dst[j];
              o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            if (flags & kRelu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (flags & kMask) {
              const float4 mm = __ldg(msk + j);
              o.x = mm.x > 0.f ? o.x : 0.f; o.y = mm.y > 0.f ? o.y : 0.f; o.z = mm.z > 0.f ? o.z : 0.f; o.w = mm.w > 0.f ? o.w : 0.f;
            }
            if (flags & kRound) { o.x = rna_tf32(o.x); o.y = rna_tf32(o.y); o.z = rna_tf32(o.z); o.w = rna_tf32(o.w); }
            dst[j] = o;
          }
        }
        tc::tc_fence_before();
        tc::mbar_arrive_leader(&tmem_empty_bar[buf]);
        continue;
      }
      // ---- K slice of a leftover tile: publish the partial now, reduce after the whole tiles
      float* part = sk.ws + ((((size_t)wi.left_idx * sk.slices + wi.slice) * 2 + crank) * kBM + row) * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TM_COLS + c * 32, v);
        tc::tmem_ld_wait();
        float4* dstp = reinterpret_cast<float4*>(part + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) dstp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      tc::tc_fence_before();
      tc::mbar_arrive_leader(&tmem_empty_bar[buf]);
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) atomicAdd(sk.flags + wi.left_idx * 2 + crank, 1);
    }
    // ---- reduction of the leftover tile this pair holds a slice of: 32 chunks of 8 columns, slice s takes chunks
    // s, s + slices, ...; fixed summation order (slice 0, 1, ...), then the fused epilogue
    if (conv_sk_item(sk, pair, 0, wi) && wi.left_idx >= 0) {
      int b, h0, w0, n0;
      tile_coords(wi.tile, b, h0, w0, n0);
      const int h = h0 + (row >> tw_log2), w = w0 + (row & (TW - 1));
      const bool valid = (h < H) && (w < W);
      const size_t pix = ((size_t)b * H + h) * W + w;
      const int* counter = sk.flags + wi.left_idx * 2 + crank;
      while (*reinterpret_cast<const volatile int*>(counter) < sk.slices) __nanosleep(64);
      __threadfence();
      for (int ch = wi.slice; ch < BN / 8; ch += sk.slices) {
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        for (int sl = 0; sl < sk.slices; ++sl) {
          const float4* src = reinterpret_cast<const float4*>(
              sk.ws + ((((size_t)wi.left_idx * sk.slices + sl) * 2 + crank) * kBM + row) * BN + ch * 8);
          const float4 p0 = __ldcg(src), p1 = __ldcg(src + 1);
          s0.x += p0.x; s0.y += p0.y; s0.z += p0.z; s0.w += p0.w;
          s1.x += p1.x; s1.y += p1.y; s1.z += p1.z; s1.w += p1.w;
        }
        const int co = n0 + ch * 8;
        if (!valid || co >= Cout) continue;
        float4* dst = reinterpret_cast<float4*>(y + pix * Cout + co);
        const float4* msk = reinterpret_cast<const float4*>(mask_src + pix * Cout + co);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float4 o = j ? s1 : s0;
          if (bias != nullptr) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + co) + j);
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          if (flags & kAccum) {
            const float4 p = dst[j];
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          if (flags & kRelu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (flags & kMask) {
            const float4 mm = __ldg(msk + j);
            o.x = mm.x > 0.f ? o.x : 0.f; o.y = mm.y > 0.f ? o.y : 0.f; o.z = mm.z > 0.f ? o.z : 0.f; o.w = mm.w > 0.f ? o.w : 0.f;
          }
          if (flags & kRound) { o.x = rna_tf32(o.x); o.y = rna_tf32(o.y); o.z = rna_tf32(o.z); o.w = rna_tf32(o.w); }
          dst[j] = o;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  if (warp == 1) tc::tmem_dealloc_2sm(tmem_base, 2 * TM_COLS);
}

static void* g_sk_ws = nullptr;          // partial accumulators + arrival counters, allocated once per process
static size_t g_sk_ws_bytes = 0;

template <int STAGES, bool HALO, int BN>
static int launch_conv_2cta_sk(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, const float* msk, float* y,
                               int H, int W, int Cin, int Cout, int best_log2, int tiles_w, int tiles_h, int total_tiles,
                               int dil, int flags, cudaStream_t st, bool* used) {
  *used = false;
  const int smem = STAGES * (HALO ? 17 * 1024 + 3 * (BN / 2) * tc::kTileKBytes : (kBM + BN / 2) * tc::kTileKBytes) + 1024;
  auto kern = conv3x3_tf32_2cta_sk_kernel<STAGES, HALO, BN>;
  ODW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int max_pairs = -1;             // pairs that are guaranteed co-resident (leftover slices wait for each other)
  if (max_pairs < 0) {
    cfg.gridDim = dim3(2 * (ODW_NUM_SMS / 2));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    max_pairs = n;
  }
  ConvSkParams sk;
  sk.n_ntiles = Cout / BN;
  sk.n_ptiles = (total_tiles / 2) * sk.n_ntiles;
  sk.kiters = (HALO ? 3 : 9) * (Cin / tc::kTileK);
  sk.np = min(max_pairs, min((ODW_NUM_SMS - odw_sm_margin()) / 2, sk.n_ptiles));
  if (sk.np < 8) return 0;               // not enough resident pairs: caller uses the one-tile-per-pair kernel
  sk.rounds = sk.n_ptiles / sk.np;
  sk.left = sk.n_ptiles - sk.rounds * sk.np;
  sk.slices = sk.left > 0 ? min(sk.np / sk.left, min(32, sk.kiters)) : 0;
  const size_t ws_floats = (size_t)max(sk.left * sk.slices, 1) * 2 * kBM * BN;
  const size_t need = ws_floats * sizeof(float) + (size_t)(sk.left + 1) * 2 * sizeof(int);
  if (g_sk_ws_bytes < need) {
    if (g_sk_ws) cudaFree(g_sk_ws);
    ODW_CUDA(cudaMalloc(&g_sk_ws, need));
    g_sk_ws_bytes = need;
  }
  sk.ws = reinterpret_cast<float*>(g_sk_ws);
  sk.flags = reinterpret_cast<int*>(reinterpret_cast<char*>(g_sk_ws) + ws_floats * sizeof(float));
  if (sk.left > 0) ODW_CUDA(cudaMemsetAsync(sk.flags, 0, (size_t)(sk.left + 1) * 2 * sizeof(int), st));
  cfg.gridDim = dim3(2 * sk.np);
  ODW_CUDA(cudaLaunchKernelEx(&cfg, kern, mx, mw, bias, msk, y, H, W, Cin, Cout, best_log2, tiles_w, tiles_h, dil, flags, sk));
  *used = true;
  return 0;
}

// ODWSCL_CONV_CLUSTER=2 turns the 2-CTA weight multicast on, ODWSCL_CONV_MT=1 turns the two-accumulator tiles off
// (A/B measurements)
static int conv_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0] >= '0' && e[0] <= '9') ? atoi(e) : dflt;
}

template <int BN, int STAGES, int CL, int MT>
static int launch_conv_inst(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, const float* msk, float* y,
                            int H, int W, int Cin, int Cout, int best_log2, int tiles_w, int tiles_h, int total_tiles,
                            int dil, int flags, cudaStream_t st) {
  const int smem = STAGES * (MT * kBM + BN) * tc::kTileKBytes + 1024;
  auto kern = conv3x3_tf32_kernel<BN, STAGES, CL, MT>;
  ODW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(odw_cdiv(total_tiles, MT), odw_cdiv(Cout, BN));
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  ODW_CUDA(cudaLaunchKernelEx(&cfg, kern, mx, mw, bias, msk, y, H, W, Cin, Cout, best_log2, tiles_w, tiles_h, total_tiles,
                              dil, flags));
  return 0;
}

template <int BN, int STAGES>
int launch_conv(const float* x, int B, int H, int W, int Cin, const float* wk, const float* bias, int Cout, int dil,
                int flags, const float* mask_src, float* y, cudaStream_t st) {
  // output patch shape: TH x TW = 128 pixels, minimising the padded area (ties -> wider rows)
  int best_log2 = 3;
  long long best_area = -1;
  for (int l = 3; l <= 7; ++l) {
    const int TW = 1 << l, TH = kBM >> l;
    const long long area = (long long)odw_cdiv(H, TH) * TH * odw_cdiv(W, TW) * TW;
    if (best_area < 0 || area <= best_area) { best_area = area; best_log2 = l; }
  }
  const int TW = 1 << best_log2, TH = kBM >> best_log2;
  const int tiles_w = odw_cdiv(W, TW), tiles_h = odw_cdiv(H, TH);
  const int total_tiles = B * tiles_h * tiles_w;
  const int n_tiles = odw_cdiv(Cout, BN);
  // CTA pairs (cta_group::2) for the 256-channel tiles; ODWSCL_CONV_2CTA=0 falls back to one CTA per tile
  const bool pair = BN >= 64 && Cout % BN == 0 && total_tiles % 2 == 0 && conv_env("ODWSCL_CONV_2CTA", 1) != 0 &&
                    (BN == 256 || conv_env("ODWSCL_CONV_PERSIST", 1) != 0);
  // two accumulators per CTA (ODWSCL_CONV_MT=2; measured slower at the bench shapes, off by default)
  const bool mt2 = !pair && BN == 256 && conv_env("ODWSCL_CONV_MT", 1) >= 2 && odw_cdiv(total_tiles, 2) * n_tiles >= ODW_NUM_SMS;
  const int cl = (pair || (!mt2 && conv_env("ODWSCL_CONV_CLUSTER", 1) >= 2 && total_tiles % 2 == 0)) ? 2 : 1;
  CUtensorMap mx, mw;
  const uint64_t dx[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t sx[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
  const uint32_t bx[4] = {32, (uint32_t)TW, (uint32_t)TH, 1};
  int rc = tc::make_tmap_f32(&mx, x, 4, dx, sx, bx);
  if (rc) return rc;
  const uint64_t dw[2] = {(uint64_t)9 * Cin, (uint64_t)Cout};
  const uint64_t sw[1] = {(uint64_t)9 * Cin * 4};
  const uint32_t bw[2] = {32, (uint32_t)(BN / cl)};
  rc = tc::make_tmap_f32(&mw, wk, 2, dw, sw, bw);
  if (rc) return rc;
  const float* msk = mask_src ? mask_src : y;
  if constexpr (BN >= 64) {
    if (pair) {
      if (conv_env("ODWSCL_CONV_PERSIST", 1) != 0) {
        bool used = false;
        if (best_log2 == 7 && dil <= 4 && conv_env("ODWSCL_CONV_HALO", 0) != 0) {
          // 128 x 1 tiles: one halo row per (tap row, slab), the three column taps slide over it
          CUtensorMap mxh;
          const uint32_t bxh[4] = {32, (uint32_t)(128 + 2 * dil), 1, 1};
          rc = tc::make_tmap_f32(&mxh, x, 4, dx, sx, bxh);
          if (rc) return rc;
          rc = launch_conv_2cta_sk<3, true, BN>(mxh, mw, bias, msk, y, H, W, Cin, Cout, best_log2, tiles_w, tiles_h,
                                                total_tiles, dil, flags, st, &used);
        } else {
          rc = launch_conv_2cta_sk<6, false, BN>(mx, mw, bias, msk, y, H, W, Cin, Cout, best_log2, tiles_w, tiles_h,
                                                 total_tiles, dil, flags, st, &used);
        }
        if (rc != 0 || used) return rc;
      }
      if constexpr (BN == 256)
        return launch_conv_2cta<6>(mx, mw, bias, msk, y, H, W, Cin, Cout, best_log2, tiles_w, tiles_h, total_tiles, dil,
                                   flags, st);
    }
    if (mt2)
      return launch_conv_inst<BN, 3, 1, 2>(mx, mw, bias, msk, y, H, W, Cin, Cout, best_log2, tiles_w, tiles_h, total_tiles,
                                           dil, flags, st);
  }
  if (cl == 2)
    return launch_conv_inst<BN, STAGES, 2, 1>(mx, mw, bias, msk, y, H, W, Cin, Cout, best_log2, tiles_w, tiles_h,
                                              total_tiles, dil, flags, st);
  return launch_conv_inst<BN, STAGES, 1, 1>(mx, mw, bias, msk, y, H, W, Cin, Cout, best_log2, tiles_w, tiles_h, total_tiles,
                                            dil, flags, st);
}

// ---------------------------------------------------------------------------------------------
// WGRAD: dW[co, tap, ci] = sum_pixels dZ[pixel, co] * X[pixel + tap, ci]  (one tap per CTA).
// GEMM with the contraction over PIXELS: M = 128 output channels, N = BN input channels, K = pixels.  Both
// operands are MN-major straight out of the NHWC tensors: a TMA box of {32 channels, 32 pixels of one image
// row} lands as 32 rows (k = pixel) of 128 bytes (32 channels) -- eight 4-row swizzle atoms, exactly the
// canonical MN-major SWIZZLE_128B_BASE32B layout (TMA swizzle mode 128B_ATOM_32B).  Boxes of X are fetched at the tap-shifted coordinate (zero fill =
// padding); boxes past the row end are zero in dZ as well, so partial chunks contribute nothing.  The pixel
// range is split across gridDim.z CTAs; partial tiles are combined with red.global.add.v4.f32 into the zeroed
// [Cout,3,3,Cin] gradient.
constexpr int kWgPix = 32;                                   // pixels (K) per pipeline stage
constexpr int kWgBox = kWgPix * tc::kTileKBytes;             // bytes of one {32 ch x 32 px} box

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_tf32_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x,
                          float* __restrict__ dw, int H, int W, int B, int Cin, int Cout, int dil, int chunks_per_cta) {
  constexpr int A_BYTES = 4 * kWgBox, B_BYTES = (BN / 32) * kWgBox, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cin_tiles = Cin / BN;
  const int co0 = (blockIdx.x / cin_tiles) * kBM, ci0 = (blockIdx.x % cin_tiles) * BN;
  const int tap = blockIdx.y;
  const int r = tap / 3, q3 = tap - 3 * r;
  const int cpr = (W + kWgPix - 1) / kWgPix;                  // chunks per image row
  const int total_chunks = B * H * cpr;
  const int c_begin = blockIdx.z * chunks_per_cta;
  const int c_end = min(total_chunks, c_begin + chunks_per_cta);
  const int kiters = max(c_end - c_begin, 0);

  if (warp == 0 && tc::elect_one()) {
    tc::tma_prefetch_desc(&map_dz);
    tc::tma_prefetch_desc(&map_x);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tmem_full_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, BN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (kiters > 0) {
    if (warp == 0) {
      if (tc::elect_one()) {
        int c = c_begin;
        int wc = c % cpr, hb = c / cpr;
        int h = hb % H, b = hb / H;
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          uint8_t* a = tiles + (size_t)s * STAGE_BYTES;
          const int w0 = wc * kWgPix;
#pragma unroll
          for (int j = 0; j < 4; ++j) tc::tma_load_4d(a + j * kWgBox, &map_dz, &full_bar[s], co0 + 32 * j, w0, h, b);
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            tc::tma_load_4d(a + A_BYTES + j * kWgBox, &map_x, &full_bar[s], ci0 + 32 * j, w0 + (q3 - 1) * dil,
                            h + (r - 1) * dil, b);
          if (++wc == cpr) { wc = 0; if (++h == H) { h = 0; ++b; } }
        }
      }
    } else if (warp == 1) {
      if (tc::elect_one()) {
        constexpr uint32_t idesc = tc::umma_idesc_tf32_mn(kBM, BN);
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t a = tc::smem_u32(tiles + (size_t)s * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < kWgPix / tc::kUmmaK; ++k) {        // 8 pixels per MMA = two 4-row (512-byte) atoms down each box
            const uint64_t ad = tc::umma_desc_mn_sw128_32b(a + k * 1024, kWgBox, 512);
            const uint64_t bd = tc::umma_desc_mn_sw128_32b(a + A_BYTES + k * 1024, kWgBox, 512);
            tc::umma_tf32(tmem_base, ad, bd, idesc, (it | k) != 0);
          }
          tc::umma_commit(&empty_bar[s]);
        }
        tc::umma_commit(&tmem_full_bar);
      }
    } else {
      const int q = warp & 3;
      const int co = co0 + q * 32 + lane;
      tc::mbar_wait(&tmem_full_bar, 0);
      tc::tc_fence_after();
      float v[32];
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
        tc::tmem_ld_wait();
        if (co >= Cout) continue;
        float* dst = dw + ((size_t)co * 9 + tap) * Cin + ci0 + c * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                       "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, BN);
}

template <int BN>
int launch_wgrad(const float* x, const float* dz, int B, int H, int W, int Cin, int Cout, int dil, float* dw, cudaStream_t st) {
  constexpr int STAGES = 4;
  CUtensorMap mz, mx;
  const uint32_t box[4] = {32, (uint32_t)kWgPix, 1, 1};
  const uint64_t dzd[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t dzs[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
  int rc = tc::make_tmap_f32(&mz, dz, 4, dzd, dzs, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  const uint64_t dx[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t sx[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
  rc = tc::make_tmap_f32(&mx, x, 4, dx, sx, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  const int tiles = odw_cdiv(Cout, kBM) * (Cin / BN);
  const int total_chunks = B * H * odw_cdiv(W, kWgPix);
  int splits = max(1, (2 * ODW_NUM_SMS) / (tiles * 9));
  splits = min(splits, total_chunks);
  const int per = odw_cdiv(total_chunks, splits);
  splits = odw_cdiv(total_chunks, per);
  const int smem = STAGES * (4 + BN / 32) * kWgBox + 1024;
  ODW_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_tf32_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ODW_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cout * 9 * Cin * sizeof(float), st));
  dim3 grid(tiles, 9, splits);
  conv3x3_wgrad_tf32_kernel<BN, STAGES><<<grid, 192, smem, st>>>(mz, mx, dw, H, W, B, Cin, Cout, dil, per);
  ODW_LAUNCH_CHECK();
  return 0;
}

// CTA-pair WGRAD (cta_group::2) for Cout % 256 == 0 and Cin % 256 == 0: the pair owns 256 output channels x 256 input
// channels of one tap; each CTA stages its own 128 dZ channels and HALF of the X channels (same shared-memory
// argument as conv3x3_tf32_2cta_kernel), the leader issues M = 256 MMAs, each CTA reduces its 128 accumulator rows.
template <int STAGES>
__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_tf32_2cta_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x,
                               float* __restrict__ dw, int H, int W, int B, int Cin, int Cout, int dil,
                               int chunks_per_cta) {
  constexpr int BN = 256;
  constexpr int A_BYTES = 4 * kWgBox, B_BYTES = (BN / 2 / 32) * kWgBox, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = tc::cluster_ctarank();
  const int co_tiles = Cout / kBM;                             // even; adjacent blockIdx.x = adjacent Cout tiles
  const int co0 = (blockIdx.x % co_tiles) * kBM, ci0 = (blockIdx.x / co_tiles) * BN;
  const int tap = blockIdx.y;
  const int r = tap / 3, q3 = tap - 3 * r;
  const int cpr = (W + kWgPix - 1) / kWgPix;
  const int total_chunks = B * H * cpr;
  const int c_begin = blockIdx.z * chunks_per_cta;
  const int c_end = min(total_chunks, c_begin + chunks_per_cta);
  const int kiters = max(c_end - c_begin, 0);                  // identical in both CTAs of the pair (same blockIdx.z)

  if (warp == 0 && tc::elect_one()) {
    tc::tma_prefetch_desc(&map_dz);
    tc::tma_prefetch_desc(&map_x);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tmem_full_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_2sm(&tmem_base_s, BN);
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (kiters > 0) {
    if (warp == 0) {
      if (tc::elect_one()) {
        int c = c_begin;
        int wc = c % cpr, hb = c / cpr;
        int h = hb % H, b = hb / H;
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          if (crank == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
          uint8_t* a = tiles + (size_t)s * STAGE_BYTES;
          const int w0 = wc * kWgPix;
#pragma unroll
          for (int j = 0; j < 4; ++j) tc::tma_load_4d_2sm(a + j * kWgBox, &map_dz, &full_bar[s], co0 + 32 * j, w0, h, b);
#pragma unroll
          for (int j = 0; j < BN / 2 / 32; ++j)
            tc::tma_load_4d_2sm(a + A_BYTES + j * kWgBox, &map_x, &full_bar[s], ci0 + (int)crank * (BN / 2) + 32 * j,
                                w0 + (q3 - 1) * dil, h + (r - 1) * dil, b);
          if (++wc == cpr) { wc = 0; if (++h == H) { h = 0; ++b; } }
        }
      }
    } else if (warp == 1) {
      if (crank == 0 && tc::elect_one()) {
        constexpr uint32_t idesc = tc::umma_idesc_tf32_mn(2 * kBM, BN);
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t a = tc::smem_u32(tiles + (size_t)s * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < kWgPix / tc::kUmmaK; ++k) {
            const uint64_t ad = tc::umma_desc_mn_sw128_32b(a + k * 1024, kWgBox, 512);
            const uint64_t bd = tc::umma_desc_mn_sw128_32b(a + A_BYTES + k * 1024, kWgBox, 512);
            tc::umma_tf32_2sm(tmem_base, ad, bd, idesc, (it | k) != 0);
          }
          tc::umma_commit_2sm_mc(&empty_bar[s], 3);
        }
        tc::umma_commit_2sm_mc(&tmem_full_bar, 3);
      }
    } else {
      const int q = warp & 3;
      const int co = co0 + q * 32 + lane;
      tc::mbar_wait(&tmem_full_bar, 0);
      tc::tc_fence_after();
      float v[32];
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
        tc::tmem_ld_wait();
        if (co >= Cout) continue;
        float* dst = dw + ((size_t)co * 9 + tap) * Cin + ci0 + c * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                       "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  if (warp == 1) tc::tmem_dealloc_2sm(tmem_base, BN);
}

static int launch_wgrad_2cta(const float* x, const float* dz, int B, int H, int W, int Cin, int Cout, int dil, float* dw,
                             cudaStream_t st) {
  constexpr int STAGES = 6;
  CUtensorMap mz, mx;
  const uint32_t box[4] = {32, (uint32_t)kWgPix, 1, 1};
  const uint64_t dzd[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t dzs[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
  int rc = tc::make_tmap_f32(&mz, dz, 4, dzd, dzs, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  const uint64_t dx[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t sx[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
  rc = tc::make_tmap_f32(&mx, x, 4, dx, sx, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  const int tiles = (Cout / kBM) * (Cin / 256);
  const int total_chunks = B * H * odw_cdiv(W, kWgPix);
  int splits = max(1, (ODW_NUM_SMS) / (tiles * 9));
  splits = min(splits, total_chunks);
  const int per = odw_cdiv(total_chunks, splits);
  splits = odw_cdiv(total_chunks, per);
  const int smem = STAGES * (4 + 4) * kWgBox + 1024;
  auto kern = conv3x3_wgrad_tf32_2cta_kernel<STAGES>;
  ODW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ODW_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cout * 9 * Cin * sizeof(float), st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles, 9, splits);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ODW_CUDA(cudaLaunchKernelEx(&cfg, kern, mz, mx, dw, H, W, B, Cin, Cout, dil, per));
  return 0;
}

// db[co] = sum_pixels dz[pixel, co]
__global__ void bias_grad_kernel(const float* __restrict__ dz, long long P, int C, float* __restrict__ db) {
  // block = 32 channels x 8 pixel lanes; grid.x = channel groups, grid.y = pixel slabs
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int py = threadIdx.x >> 5;
  float acc = 0.f;
  if (c < C)
    for (long long p = (long long)blockIdx.y * 8 + py; p < P; p += (long long)gridDim.y * 8) acc += dz[p * C + c];
  __shared__ float sm[8][33];
  sm[py][threadIdx.x & 31] = acc;
  __syncthreads();
  if (py == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += sm[j][threadIdx.x & 31];
    atomicAdd(db + c, t);
  }
}

// ---------------------------------------------------------------------------------------------
// conv1_1: Cin = 3 (K = 27: not tensor-core shaped).  CTA = 128 consecutive pixels of one image row; the 3 x 3-row
// input patch and the 27 x 64 weights sit in shared memory; thread = (pixel pair, 32-channel half), so one broadcast
// 16-byte weight load feeds 8 FFMA (64 accumulators per thread).  Reads the NCHW image, writes NHWC.
constexpr int kC3Tile = 128;
template <int COUT>
__global__ void __launch_bounds__(128)
conv3x3_c3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int B, int H, int W, int relu) {
  static_assert(COUT == 64, "two 32-channel halves");
  __shared__ __align__(16) float sw[27 * COUT];     // [tap*3+ci][co]
  __shared__ __align__(16) float sb[COUT];
  __shared__ float sx[3][3][kC3Tile + 4];           // [ci][row][col + 1], zero outside the image
  const int tiles_w = (W + kC3Tile - 1) / kC3Tile;
  int t = blockIdx.x;
  const int w0 = (t % tiles_w) * kC3Tile; t /= tiles_w;
  const int hq = t % H;
  const int bq = t / H;
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
    const int co = i % COUT, k = i / COUT;          // k = (r*3+s)*3 + ci ; torch layout [co][ci][r][s]
    const int ci = k % 3, tap = k / 3;
    sw[i] = w[(co * 3 + ci) * 9 + tap];
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  const float* xb = x + (size_t)bq * 3 * H * W;
  for (int i = threadIdx.x; i < 9 * (kC3Tile + 2); i += blockDim.x) {
    const int col = i % (kC3Tile + 2), rc = i / (kC3Tile + 2);
    const int r = rc % 3, ci = rc / 3;
    const int hh = hq + r - 1, ww = w0 + col - 1;
    sx[ci][r][col] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xb + ((size_t)ci * H + hh) * W + ww) : 0.f;
  }
  __syncthreads();
  const int pp = threadIdx.x & 63, half = threadIdx.x >> 6;
  float a0[32], a1[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) a0[j] = a1[j] = sb[half * 32 + j];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float x0 = sx[ci][r][2 * pp + q], x1 = sx[ci][r][2 * pp + 1 + q];
        const float4* wr = reinterpret_cast<const float4*>(sw + ((r * 3 + q) * 3 + ci) * COUT + half * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 wv = wr[j];
          a0[4 * j] = fmaf(x0, wv.x, a0[4 * j]);         a1[4 * j] = fmaf(x1, wv.x, a1[4 * j]);
          a0[4 * j + 1] = fmaf(x0, wv.y, a0[4 * j + 1]); a1[4 * j + 1] = fmaf(x1, wv.y, a1[4 * j + 1]);
          a0[4 * j + 2] = fmaf(x0, wv.z, a0[4 * j + 2]); a1[4 * j + 2] = fmaf(x1, wv.z, a1[4 * j + 2]);
          a0[4 * j + 3] = fmaf(x0, wv.w, a0[4 * j + 3]); a1[4 * j + 3] = fmaf(x1, wv.w, a1[4 * j + 3]);
        }
      }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int wq = w0 + 2 * pp + e;
    if (wq >= W) continue;
    const float* acc = e ? a1 : a0;
    float4* dst = reinterpret_cast<float4*>(y + (((size_t)bq * H + hq) * W + wq) * COUT + half * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 o = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      if (relu & kRelu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (relu & kRound) { o.x = rna_tf32(o.x); o.y = rna_tf32(o.y); o.z = rna_tf32(o.z); o.w = rna_tf32(o.w); }
      dst[j] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void maxpool2x2_nhwc_kernel(const float4* __restrict__ x, float4* __restrict__ y, int B, int H, int W, int C4) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long p = i / C4;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const float4* s = x + (((size_t)b * H + 2 * ho) * W + 2 * wo) * C4 + c;
    const float4 a = __ldg(s), bb = __ldg(s + C4), cc = __ldg(s + (size_t)W * C4), d = __ldg(s + (size_t)W * C4 + C4);
    float4 o;
    o.x = fmaxf(fmaxf(a.x, bb.x), fmaxf(cc.x, d.x)); o.y = fmaxf(fmaxf(a.y, bb.y), fmaxf(cc.y, d.y));
    o.z = fmaxf(fmaxf(a.z, bb.z), fmaxf(cc.z, d.z)); o.w = fmaxf(fmaxf(a.w, bb.w), fmaxf(cc.w, d.w));
    y[i] = o;
  }
}

// gx = grad routed to the FIRST maximal element of each 2x2 window (torch max_pool2d semantics), optionally
// multiplied by the ReLU derivative of x (x is a post-ReLU activation: derivative 0 where x == 0).
__global__ void maxpool2x2_nhwc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                           int B, int H, int W, int C, int relu_mask) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int b = (int)(p / H);
    const int ho = h >> 1, wo = w >> 1;
    float g = 0.f;
    if (ho < Ho && wo < Wo) {
      const float* s = x + (((size_t)b * H + 2 * ho) * W + 2 * wo) * C + c;
      const float v0 = s[0], v1 = s[C], v2 = s[(size_t)W * C], v3 = s[(size_t)W * C + C];
      int am = 0; float m = v0;
      if (v1 > m) { m = v1; am = 1; }
      if (v2 > m) { m = v2; am = 2; }
      if (v3 > m) { m = v3; am = 3; }
      const int me = ((h & 1) << 1) | (w & 1);
      if (me == am && !(relu_mask && !(m > 0.f))) g = gy[(((size_t)b * Ho + ho) * Wo + wo) * C + c];
    }
    gx[i] = g;
  }
}

// same routing, one thread per 2x2 window x 4 channels: every byte of x / gy is read once, gx written once
// (H, W even, C % 4 == 0)
__global__ void __launch_bounds__(256)
maxpool2x2_nhwc_bwd_v4_kernel(const float4* __restrict__ x, const float4* __restrict__ gy, float4* __restrict__ gx,
                              int B, int H, int W, int C4, int relu_mask) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long p = i / C4;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const size_t o00 = (((size_t)b * H + 2 * ho) * W + 2 * wo) * C4 + c;
    const size_t o01 = o00 + C4, o10 = o00 + (size_t)W * C4, o11 = o10 + C4;
    const float4 v0 = __ldcs(x + o00), v1 = __ldcs(x + o01), v2 = __ldcs(x + o10), v3 = __ldcs(x + o11);
    const float4 g = __ldcs(gy + i);
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, r3 = r0;
#define ODW_POOL_ROUTE(f)                                              \
    {                                                                  \
      int am = 0; float m = v0.f;                                      \
      if (v1.f > m) { m = v1.f; am = 1; }                              \
      if (v2.f > m) { m = v2.f; am = 2; }                              \
      if (v3.f > m) { m = v3.f; am = 3; }                              \
      const float gg = (relu_mask && !(m > 0.f)) ? 0.f : g.f;          \
      if (am == 0) r0.f = gg; else if (am == 1) r1.f = gg; else if (am == 2) r2.f = gg; else r3.f = gg; \
    }
    ODW_POOL_ROUTE(x) ODW_POOL_ROUTE(y) ODW_POOL_ROUTE(z) ODW_POOL_ROUTE(w)
#undef ODW_POOL_ROUTE
    gx[o00] = r0; gx[o01] = r1; gx[o10] = r2; gx[o11] = r3;
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ x, long long n, float* __restrict__ hi, float* __restrict__ lo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float h = rna_tf32(v);
    hi[i] = h;
    if (lo != nullptr) lo[i] = rna_tf32(__fsub_rn(v, h));
  }
}

}  // namespace

ODW_API int odwscl_conv3x3_nhwc_tf32(const float* x, int B, int H, int W, int Cin, const float* w_krsc,
                                     const float* bias, int Cout, int dilation, int flags, const float* mask_src,
                                     float* y, odwscl_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || Cin <= 0 || Cout <= 0 || (Cin % 32) || (Cout % 32) || dilation < 1) return ODWSCL_EINVAL;
  if ((long long)B * H * W == 0) return 0;
  if (!x || !w_krsc || !y || ((flags & kMask) && !mask_src)) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cout % 256 == 0) return launch_conv<256, 4>(x, B, H, W, Cin, w_krsc, bias, Cout, dilation, flags, mask_src, y, st);
  if (Cout % 128 == 0) return launch_conv<128, 3>(x, B, H, W, Cin, w_krsc, bias, Cout, dilation, flags, mask_src, y, st);
  if (Cout % 64 == 0) return launch_conv<64, 4>(x, B, H, W, Cin, w_krsc, bias, Cout, dilation, flags, mask_src, y, st);
  return launch_conv<32, 4>(x, B, H, W, Cin, w_krsc, bias, Cout, dilation, flags, mask_src, y, st);
}

ODW_API int odwscl_conv3x3_wgrad_nhwc_tf32(const float* x, const float* dz, int B, int H, int W, int Cin, int Cout,
                                           int dilation, float* dw_krsc, float* db, odwscl_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || Cin <= 0 || Cout <= 0 || (Cin % 32) || (Cout % 32) || dilation < 1) return ODWSCL_EINVAL;
  if (!dw_krsc) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if ((long long)B * H * W == 0) {
    ODW_CUDA(cudaMemsetAsync(dw_krsc, 0, (size_t)Cout * 9 * Cin * sizeof(float), st));
    if (db) ODW_CUDA(cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), st));
    return 0;
  }
  if (!x || !dz) return ODWSCL_EINVAL;
  if (db) {
    ODW_CUDA(cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), st));
    const long long P = (long long)B * H * W;
    dim3 grid(odw_cdiv(Cout, 32), (unsigned)min((long long)ODW_NUM_SMS * 4 / odw_cdiv(Cout, 32) + 1, (P + 7) / 8));
    bias_grad_kernel<<<grid, 256, 0, st>>>(dz, P, Cout, db);
    ODW_LAUNCH_CHECK();
  }
  if (Cin % 256 == 0 && Cout % 256 == 0 && conv_env("ODWSCL_CONV_2CTA", 1) != 0)
    return launch_wgrad_2cta(x, dz, B, H, W, Cin, Cout, dilation, dw_krsc, st);
  if (Cin % 256 == 0) return launch_wgrad<256>(x, dz, B, H, W, Cin, Cout, dilation, dw_krsc, st);
  if (Cin % 128 == 0) return launch_wgrad<128>(x, dz, B, H, W, Cin, Cout, dilation, dw_krsc, st);
  if (Cin % 64 == 0) return launch_wgrad<64>(x, dz, B, H, W, Cin, Cout, dilation, dw_krsc, st);
  return launch_wgrad<32>(x, dz, B, H, W, Cin, Cout, dilation, dw_krsc, st);
}

ODW_API int odwscl_conv3x3_c3_f32(const float* x_nchw, int B, int H, int W, const float* w_oihw, const float* bias,
                                  int Cout, int relu, float* y_nhwc, odwscl_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || Cout != 64) return ODWSCL_EINVAL;
  if ((long long)B * H * W == 0) return 0;
  if (!x_nchw || !w_oihw || !y_nhwc) return ODWSCL_EINVAL;
  const long long blocks = (long long)B * H * odw_cdiv(W, kC3Tile);
  if (blocks > 0x7fffffffLL) return ODWSCL_EINVAL;
  conv3x3_c3_kernel<64><<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(x_nchw, w_oihw, bias, y_nhwc, B, H, W, relu);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_maxpool2x2_nhwc_f32(const float* x, int B, int H, int W, int C, float* y, odwscl_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || C <= 0 || (C & 3)) return ODWSCL_EINVAL;
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 4);
  if (total == 0) return 0;
  if (!x || !y) return ODWSCL_EINVAL;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 16, (total + 255) / 256);
  maxpool2x2_nhwc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x),
                                                                   reinterpret_cast<float4*>(y), B, H, W, C / 4);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_maxpool2x2_nhwc_bwd_f32(const float* x, const float* gy, int B, int H, int W, int C, int relu_mask,
                                           float* gx, odwscl_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || C <= 0) return ODWSCL_EINVAL;
  const long long total = (long long)B * H * W * C;
  if (total == 0) return 0;
  if (!x || !gy || !gx) return ODWSCL_EINVAL;
  if ((C & 3) == 0 && (H & 1) == 0 && (W & 1) == 0) {
    const long long t4 = total / 16;
    const int blocks4 = (int)min((long long)ODW_NUM_SMS * 16, (t4 + 255) / 256);
    maxpool2x2_nhwc_bwd_v4_kernel<<<blocks4, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(gy), reinterpret_cast<float4*>(gx), B, H, W,
        C / 4, relu_mask);
    ODW_LAUNCH_CHECK();
    return 0;
  }
  const int blocks = (int)min((long long)ODW_NUM_SMS * 16, (total + 255) / 256);
  maxpool2x2_nhwc_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, gy, gx, B, H, W, C, relu_mask);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_split_tf32(const float* x, long long n, float* hi, float* lo, odwscl_stream_t stream) {
  if (n < 0) return ODWSCL_EINVAL;
  if (n == 0) return 0;
  if (!x || !hi) return ODWSCL_EINVAL;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 16, (n + 255) / 256);
  split_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, hi, lo);
  ODW_LAUNCH_CHECK();
  return 0;
}
