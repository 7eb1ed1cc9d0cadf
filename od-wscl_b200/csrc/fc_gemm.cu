// fc_gemm.cu -- the fully-connected block of the proposal-feature path:
// fc6 / fc7 (modeling/backbone/vgg16.py:122-130,148-162), Sim_Net (roi_heads/sim_head/sim_net.py:10-26) and the eight
// MIST predictor heads (roi_heads/weak_head/roi_weak_predictors.py:158-165) -- forward, input gradient and weight
// gradient -- as ONE hand-written sm_100a GEMM family with fused epilogues (the reference calls cuBLAS through
// nn.Linear + separate ReLU / Dropout / add kernels).
//
//   C[M,N] (+)= sum_k A(m,k) * B(n,k)            fp32 storage, TF32 tensor-core math, fp32 accumulation in TMEM
//
// Each operand is either K-major (memory [rows, K], the contraction index contiguous) or MN-major (memory [K, rows],
// the row index contiguous), which covers the three GEMMs of a linear layer without any transposed copy:
//   forward   Y[M,N]  = X[M,K]  W[N,K]^T          A = X  K-major,  B = W  K-major
//   dgrad     dX[M,K] = dY[M,N] W[N,K]            A = dY K-major,  B = W  MN-major (contraction over W's ROWS)
//   wgrad     dW[N,K] = dY[M,N]^T X[M,K]          A = dY MN-major, B = X  MN-major (contraction over the batch rows)
//
// Kernel anatomy = the persistent CTA-pair schedule of conv3x3_tf32_2cta_sk_kernel (conv3x3.cu): a cluster of two
// CTAs owns a 256 x 256 output tile (tcgen05 cta_group::2: each CTA stages its own 128 rows of A and HALF of the B
// tile, the leader issues M = 256 MMAs into both CTAs' TMEM), 74 resident pairs take whole tiles round by round in
// lock-step along K (so concurrently running tiles ask L2 for the same operand slabs together), leftover tiles are
// split along K over all pairs and reduced in a fixed order, and the accumulator is double-buffered in TMEM (2 x 256
// columns) so the fused epilogue of tile i overlaps the MMAs of tile i+1.  TF32 operands are 4 bytes for half the
// bf16 MMA rate: at 256 x 256 x 32 per stage every SM ingests 32 KB and the tensor core reads 32 KB per ~520 clk --
// the shared-memory port (128 B/clk) is the bound, exactly as for the conv kernel and for cuBLAS's own 2-SM kernel.
//
// Tiles are rasterised in panels of 8 n-tiles x all m-tiles, so one round of 74 tiles covers ~9 x 8 tiles and reads
// (9 + 8) operand slabs from HBM instead of (1 + 74).
//
// Fused epilogues (flags): + bias[n], accumulate into C (3xTF32 strict mode; folding a second weight gradient),
// ReLU, Dropout(p) with a counter-based Philox stream keyed by the element index (no mask tensor: the backward reads
// y > 0), the ReLU/Dropout DERIVATIVE mask of the layer below (dgrad: dZ = dX * scale * [y_prev > 0]) and rounding
// of the output to TF32 (cvt.rna) when another tensor-core GEMM consumes it (the tensor core truncates otherwise).
#include <stdlib.h>

#include "common.cuh"
#include "tc_sm100.cuh"

namespace {

constexpr int kBM = 128;                 // rows of A per CTA (256 per pair)
constexpr int kBN = 256;                 // columns per pair tile
constexpr int kBox = 32 * 128;           // bytes of one MN-major {32 rows(mn) x 32 k} TMA box
enum : int { kBias = 1, kAccum = 2, kRelu = 4, kDropout = 8, kMask = 16, kRound = 32, kPeerSum = 64, kPeerOwner = 128 };

__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t b;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v));
  return __uint_as_float(b);
}

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(unsigned long long seed, unsigned long long ctr, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x0DDBA11u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

struct FcEpi {
  const float* bias;        // [N] or null
  const float* mask_src;    // [M, ld_mask] (kMask): derivative mask source (the activation the output is a gradient of)
  int ld_mask;
  float mask_scale;         // 1 / (1 - p) of the layer below
  float drop_scale;         // 1 / (1 - p)
  uint32_t drop_thr16;      // keep iff 16 random bits >= thr
  unsigned long long seed;
  int flags;
  long long mc_delta;       // kPeerSum: byte distance from C to its NVSwitch multicast alias
  float out_scale;          // kPeerSum: factor on the stored value (1 / world_size: DDP averages)
  // kPeerSum | kPeerOwner: rows [r * rows_per_owner, (r + 1) * rows_per_owner) are summed on rank r only (reduce-scatter);
  // peer_delta[r] = byte distance from C to rank r's replica of C (peer-mapped symmetric memory)
  long long peer_delta[8];
  int rows_per_owner, n_owners;
};

// Adds 4 floats to the same address of EVERY rank's replica in one NVLink operation: the switch forwards the reduction
// to all members of the multicast object (NVLS), so the cross-rank sum of a weight gradient leaves the GEMM epilogue
// tile by tile while the tensor pipe works on the next tile -- no all-reduce kernel afterwards.
__device__ __forceinline__ void multimem_red_add_v4(float* mc, float a, float b, float c, float d) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void peer_red_add_v4(float* p, float a, float b, float c, float d) {   // one rank's memory, over NVLink
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void peer_red_add(float* p, float a) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void multimem_red_add(float* mc, float a) {
  asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(mc), "f"(a) : "memory");
}

struct FcSk {
  int m_tiles, n_tiles;     // 256-row / 256-column pair tiles
  int n_ptiles, kiters;     // kiters = kiters0 + the 32-wide K steps of the optional second operand pair
  int kiters0;              // K steps served by the first (A, B) pair
  int np, rounds, left, slices;
  int panel;                // n-tiles per raster panel
  float* ws;                // [left * slices][2][128][256] partial accumulators
  int* flags;               // [left][2] arrival counters
};

struct FcItem { int tile, k0, k1, left_idx, slice; };

__device__ __forceinline__ bool fc_item(const FcSk& sk, int pair, int idx, FcItem& it) {
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
  if (idx < sk.rounds) {
    it.tile = idx * sk.np + pair; it.k0 = 0; it.k1 = sk.kiters; it.left_idx = -1; it.slice = 0;
    return it.tile < sk.n_ptiles;         // the last round may be partial
  }
  return false;
}

// pair-tile index -> (first row of the PAIR, first column): panels of `panel` n-tiles, m fastest-but-one
__device__ __forceinline__ void fc_tile_coords(const FcSk& sk, int pt, int& m0, int& n0) {
  const int per_panel = sk.panel * sk.m_tiles;
  const int pn = pt / per_panel;
  const int within = pt - pn * per_panel;
  const int width = min(sk.panel, sk.n_tiles - pn * sk.panel);
  const int mt = within / width;
  const int nt = pn * sk.panel + (within - mt * width);
  m0 = mt * 2 * kBM;
  n0 = nt * kBN;
}

// Fused epilogue of 32 consecutive output columns of ONE row (thread = accumulator row, as tcgen05.ld delivers them):
// 8 x 16-byte accesses per thread.  (A variant that transposed each 32 x 32 chunk through shared memory so that every warp
// access was one full 128-byte row segment was measured 2-3x SLOWER on the epilogue-bound shapes -- 32 scalar stores per
// lane instead of 8 vector stores: the LSU instruction count, not the sector count, is what limits this epilogue.)
__device__ __forceinline__ void fc_epilogue_store(float (&v)[32], float* __restrict__ C, int ldc, int row, int col, int M, int N,
                                                  const FcEpi& ep) {
  if (row >= M || col >= N) return;
  const int flags = ep.flags;
  const bool full = (col + 32 <= N) && ((ldc & 3) == 0);
  float* dst = C + (size_t)row * ldc + col;
  if (flags & kBias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (col + j < N) v[j] += __ldg(ep.bias + col + j);
  }
  if (flags & kAccum) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 p = reinterpret_cast<const float4*>(dst)[j];
        v[4 * j] += p.x; v[4 * j + 1] += p.y; v[4 * j + 2] += p.z; v[4 * j + 3] += p.w;
      }
    } else {
      for (int j = 0; j < 32 && col + j < N; ++j) v[j] += dst[j];
    }
  }
  if (flags & kRelu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (flags & kDropout) {
    // element (row, col + j): 16 random bits of Philox(seed, counter = (row * N + col + j) / 8), slot (.. % 8)
    const unsigned long long e0 = (unsigned long long)row * (unsigned long long)N + (unsigned long long)col;
    if ((e0 & 7ull) == 0) {
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t r[4];
        philox4x32_10(ep.seed, (e0 >> 3) + c4, r);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint32_t bits = (r[t >> 1] >> ((t & 1) * 16)) & 0xFFFFu;
          v[c4 * 8 + t] = bits >= ep.drop_thr16 ? v[c4 * 8 + t] * ep.drop_scale : 0.f;
        }
      }
    } else {                                   // N % 8 != 0: per-element counters (never on the fc6 / fc7 shapes)
      for (int j = 0; j < 32; ++j) {
        uint32_t r[4];
        const unsigned long long e = e0 + j;
        philox4x32_10(ep.seed, e >> 3, r);
        const int t = (int)(e & 7ull);
        const uint32_t bits = (r[t >> 1] >> ((t & 1) * 16)) & 0xFFFFu;
        v[j] = bits >= ep.drop_thr16 ? v[j] * ep.drop_scale : 0.f;
      }
    }
  }
  if (flags & kMask) {
    const float* ms = ep.mask_src + (size_t)row * ep.ld_mask + col;
    if (full && (ep.ld_mask & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 mm = __ldg(reinterpret_cast<const float4*>(ms) + j);
        v[4 * j] = mm.x > 0.f ? v[4 * j] * ep.mask_scale : 0.f;
        v[4 * j + 1] = mm.y > 0.f ? v[4 * j + 1] * ep.mask_scale : 0.f;
        v[4 * j + 2] = mm.z > 0.f ? v[4 * j + 2] * ep.mask_scale : 0.f;
        v[4 * j + 3] = mm.w > 0.f ? v[4 * j + 3] * ep.mask_scale : 0.f;
      }
    } else {
      for (int j = 0; j < 32 && col + j < N; ++j) v[j] = __ldg(ms + j) > 0.f ? v[j] * ep.mask_scale : 0.f;
    }
  }
  if (flags & kRound) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = rna_tf32(v[j]);
  }
  if (full) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    for (int j = 0; j < 32 && col + j < N; ++j) dst[j] = v[j];
  }
}

// kPeerSum epilogue of one warp's 32 rows x 32 columns: every rank's replica of C (zeroed beforehand) += v * out_scale.
// The values go through a padded shared-memory tile so that 8 consecutive lanes add 128 CONTIGUOUS bytes of one row: a
// multimem.red travels over NVLink as its own request, and 16-byte requests at row stride (the row-domain epilogue above)
// ran the whole step 5 ms slower at 2 GPUs -- the fabric is request-bound long before it is byte-bound.
constexpr int kPeerPitch = 36;                 // floats per staged row: 16-byte aligned, conflict-free float4 rows
constexpr int kPeerStageBytes = 4 * 32 * kPeerPitch * (int)sizeof(float);

__device__ __forceinline__ void fc_epilogue_peer(float (&v)[32], float* __restrict__ stage, float* __restrict__ C, int ldc,
                                                 int row0, int col, int M, int N, const FcEpi& ep, int lane) {
  if (col >= N) return;                        // warp-uniform
  const float sc = ep.out_scale;
  float4* srow = reinterpret_cast<float4*>(stage + lane * kPeerPitch);
#pragma unroll
  for (int j = 0; j < 8; ++j) srow[j] = make_float4(v[4 * j] * sc, v[4 * j + 1] * sc, v[4 * j + 2] * sc, v[4 * j + 3] * sc);
  __syncwarp();
  const bool vec = (col + 32 <= N) && ((ldc & 3) == 0);
  const int cg = (lane & 7) * 4;
  if (ep.flags & kPeerOwner) {
    // reduce-scatter form: the warp's 32 rows belong to ONE rank (rows_per_owner % 32 == 0); only that rank's replica
    // receives the tile, so a rank takes in (world - 1) / world of the gradient instead of (world - 1) x all of it
    const int owner = min(row0 / ep.rows_per_owner, ep.n_owners - 1);
    const long long delta = ep.peer_delta[owner];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      const int row = row0 + r;
      const float4 t = *reinterpret_cast<const float4*>(stage + r * kPeerPitch + cg);
      if (row < M) {
        float* dst = reinterpret_cast<float*>(reinterpret_cast<char*>(C + (size_t)row * ldc + col + cg) + delta);
        if (vec) {
          peer_red_add_v4(dst, t.x, t.y, t.z, t.w);
        } else {
          if (col + cg < N) peer_red_add(dst, t.x);
          if (col + cg + 1 < N) peer_red_add(dst + 1, t.y);
          if (col + cg + 2 < N) peer_red_add(dst + 2, t.z);
          if (col + cg + 3 < N) peer_red_add(dst + 3, t.w);
        }
      }
    }
    __syncwarp();
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3);
    const int row = row0 + r;
    const float4 t = *reinterpret_cast<const float4*>(stage + r * kPeerPitch + cg);
    if (row < M) {
      float* mc = reinterpret_cast<float*>(reinterpret_cast<char*>(C + (size_t)row * ldc + col + cg) + ep.mc_delta);
      if (vec) {
        multimem_red_add_v4(mc, t.x, t.y, t.z, t.w);
      } else {
        if (col + cg < N) multimem_red_add(mc, t.x);
        if (col + cg + 1 < N) multimem_red_add(mc + 1, t.y);
        if (col + cg + 2 < N) multimem_red_add(mc + 2, t.z);
        if (col + cg + 3 < N) multimem_red_add(mc + 3, t.w);
      }
    }
  }
  __syncwarp();
}

template <bool AMN, bool BMN, int kStages>
__global__ void __launch_bounds__(192, 1)
fc_gemm_tf32_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
                         float* __restrict__ C, int M, int N, int ldc, const FcEpi ep, const FcSk sk) {
  constexpr int A_BYTES = kBM * tc::kTileKBytes, B_BYTES = (kBN / 2) * tc::kTileKBytes;    // 16 KB + 16 KB per CTA
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = tc::cluster_ctarank();
  const int pair = blockIdx.x >> 1;

  if (warp == 0 && tc::elect_one()) {
    tc::tma_prefetch_desc(&map_a);
    tc::tma_prefetch_desc(&map_b);
    if (sk.kiters0 < sk.kiters) { tc::tma_prefetch_desc(&map_a1); tc::tma_prefetch_desc(&map_b1); }
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tmem_full_bar[i], 1); tc::mbar_init(&tmem_empty_bar[i], 256); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_2sm(&tmem_base_s, 2 * kBN);
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (one lane per CTA)
    if (tc::elect_one()) {
      int it = 0;
      FcItem wi;
      for (int item = 0; fc_item(sk, pair, item, wi); ++item) {
        int m0, n0;
        fc_tile_coords(sk, wi.tile, m0, n0);
        const int am0 = m0 + (int)crank * kBM;                 // this CTA's 128 rows of A
        const int bn0 = n0 + (int)crank * (kBN / 2);           // this CTA's half of the B tile
        for (int kk = wi.k0; kk < wi.k1; ++kk, ++it) {
          const int s = it % kStages;
          tc::mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1);
          if (crank == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
          uint8_t* a = tiles + (size_t)s * STAGE_BYTES;
          // the contraction runs over the K steps of the first operand pair, then over those of the second (two calls
          // of one layer folded into a single weight-gradient GEMM); each pair has its own bounds (zero fill)
          const bool second = kk >= sk.kiters0;
          const CUtensorMap* ma = second ? &map_a1 : &map_a;
          const CUtensorMap* mb = second ? &map_b1 : &map_b;
          const int k0 = (second ? kk - sk.kiters0 : kk) * tc::kTileK;
          if (AMN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) tc::tma_load_2d_2sm(a + j * kBox, ma, &full_bar[s], am0 + 32 * j, k0);
          } else {
            tc::tma_load_2d_2sm(a, ma, &full_bar[s], k0, am0);
          }
          if (BMN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) tc::tma_load_2d_2sm(a + A_BYTES + j * kBox, mb, &full_bar[s], bn0 + 32 * j, k0);
          } else {
            tc::tma_load_2d_2sm(a + A_BYTES, mb, &full_bar[s], k0, bn0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (one lane of the leader CTA)
    if (crank == 0 && tc::elect_one()) {
      constexpr uint32_t idesc = tc::umma_idesc_tf32(2 * kBM, kBN) | (AMN ? (1u << 15) : 0u) | (BMN ? (1u << 16) : 0u);
      int it = 0;
      FcItem wi;
      for (int item = 0; fc_item(sk, pair, item, wi); ++item) {
        const int buf = item & 1;
        tc::mbar_wait(&tmem_empty_bar[buf], ((item >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t acc = tmem_base + buf * kBN;
        for (int kk = wi.k0; kk < wi.k1; ++kk, ++it) {
          const int s = it % kStages;
          tc::mbar_wait(&full_bar[s], (it / kStages) & 1);
          tc::tc_fence_after();
          const uint32_t a = tc::smem_u32(tiles + (size_t)s * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < tc::kTileK / tc::kUmmaK; ++k) {
            // K-major: +32 bytes per K = 8 step inside the 128-byte swizzle row; MN-major: 8 k rows = 1024 bytes down
            const uint64_t ad = AMN ? tc::umma_desc_mn_sw128_32b(a + k * 1024, kBox, 512) : tc::umma_desc_sw128(a) + 2 * k;
            const uint64_t bd = BMN ? tc::umma_desc_mn_sw128_32b(a + A_BYTES + k * 1024, kBox, 512)
                                    : tc::umma_desc_sw128(a + A_BYTES) + 2 * k;
            tc::umma_tf32_2sm(acc, ad, bd, idesc, (kk != wi.k0) || (k != 0));
          }
          tc::umma_commit_2sm_mc(&empty_bar[s], 3);
        }
        tc::umma_commit_2sm_mc(&tmem_full_bar[buf], 3);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 4 warps, thread = accumulator row
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    float v[32];
    FcItem wi;
    const bool peer = (ep.flags & kPeerSum) != 0;      // the launch then carries kPeerStageBytes behind the operand ring
    float* stage = reinterpret_cast<float*>(tiles + (size_t)kStages * (kBM + kBN / 2) * tc::kTileKBytes) + q * 32 * kPeerPitch;
    for (int item = 0; fc_item(sk, pair, item, wi); ++item) {
      const int buf = item & 1;
      int m0, n0;
      fc_tile_coords(sk, wi.tile, m0, n0);
      const int row = m0 + (int)crank * kBM + trow;
      tc::mbar_wait(&tmem_full_bar[buf], (item >> 1) & 1);
      tc::tc_fence_after();
      if (wi.left_idx < 0) {
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN + c * 32, v);
          tc::tmem_ld_wait();
          if (peer) fc_epilogue_peer(v, stage, C, ldc, row - lane, n0 + c * 32, M, N, ep, lane);
          else fc_epilogue_store(v, C, ldc, row, n0 + c * 32, M, N, ep);
        }
        tc::tc_fence_before();
        tc::mbar_arrive_leader(&tmem_empty_bar[buf]);
        continue;
      }
      // K slice of a leftover tile: publish the partial accumulator, reduce after the whole tiles
      float* part = sk.ws + ((((size_t)wi.left_idx * sk.slices + wi.slice) * 2 + crank) * kBM + trow) * kBN;
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN + c * 32, v);
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
    // reduction of the leftover tile this pair holds a slice of: 32-column chunks, slice s takes chunks s, s + slices, ...
    if (fc_item(sk, pair, 0, wi) && wi.left_idx >= 0) {
      int m0, n0;
      fc_tile_coords(sk, wi.tile, m0, n0);
      const int row = m0 + (int)crank * kBM + trow;
      const int* counter = sk.flags + wi.left_idx * 2 + crank;
      while (*reinterpret_cast<const volatile int*>(counter) < sk.slices) __nanosleep(64);
      __threadfence();
      for (int ch = wi.slice; ch < kBN / 32; ch += sk.slices) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
        for (int sl = 0; sl < sk.slices; ++sl) {
          const float4* src = reinterpret_cast<const float4*>(
              sk.ws + ((((size_t)wi.left_idx * sk.slices + sl) * 2 + crank) * kBM + trow) * kBN + ch * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 p = __ldcg(src + j);
            v[4 * j] += p.x; v[4 * j + 1] += p.y; v[4 * j + 2] += p.z; v[4 * j + 3] += p.w;
          }
        }
        if (peer) fc_epilogue_peer(v, stage, C, ldc, row - lane, n0 + ch * 32, M, N, ep, lane);
        else fc_epilogue_store(v, C, ldc, row, n0 + ch * 32, M, N, ep);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  if (warp == 1) tc::tmem_dealloc_2sm(tmem_base, 2 * kBN);
}

void* g_fc_ws = nullptr;                  // split-K partials + arrival counters (allocated once per process)
size_t g_fc_ws_bytes = 0;

int fc_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0] >= '0' && e[0] <= '9') ? atoi(e) : dflt;
}

// K-major operand [rows, K] (pitch ld floats): box {32 k, 128 rows}.  MN-major operand [K, rows]: box {32 rows, 32 k}.
int make_operand_map(CUtensorMap* map, const float* p, int rows, int K, int ld, bool mn) {
  if (mn) {
    const uint64_t d[2] = {(uint64_t)rows, (uint64_t)K};
    const uint64_t s[1] = {(uint64_t)ld * 4};
    const uint32_t b[2] = {32, 32};
    return tc::make_tmap_f32(map, p, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  }
  const uint64_t d[2] = {(uint64_t)K, (uint64_t)rows};
  const uint64_t s[1] = {(uint64_t)ld * 4};
  const uint32_t b[2] = {32, 128};
  return tc::make_tmap_f32(map, p, 2, d, s, b);
}

struct FcOperands { const float* A; int lda; const float* B; int ldb; int K; };

template <bool AMN, bool BMN, int kStages>
int launch_fc_s(const FcOperands& o0, const FcOperands& o1, float* C, int ldc, int M, int N, const FcEpi& ep,
                int max_pairs_cap, cudaStream_t st) {
  CUtensorMap ma, mb, ma1, mb1;
  int rc = make_operand_map(&ma, o0.A, M, o0.K, o0.lda, AMN);
  if (rc) return rc;
  rc = make_operand_map(&mb, o0.B, N, o0.K, o0.ldb, BMN);
  if (rc) return rc;
  ma1 = ma; mb1 = mb;
  if (o1.K > 0) {
    rc = make_operand_map(&ma1, o1.A, M, o1.K, o1.lda, AMN);
    if (rc) return rc;
    rc = make_operand_map(&mb1, o1.B, N, o1.K, o1.ldb, BMN);
    if (rc) return rc;
  }
  const int smem = kStages * (kBM + kBN / 2) * tc::kTileKBytes + 1024 + ((ep.flags & kPeerSum) ? kPeerStageBytes : 0);
  auto kern = fc_gemm_tf32_2cta_kernel<AMN, BMN, kStages>;
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
  static int max_pairs = -1;              // pairs guaranteed co-resident (slices of a leftover tile wait for each other)
  if (max_pairs < 0) {
    cfg.gridDim = dim3(2 * (ODW_NUM_SMS / 2));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    max_pairs = n > 0 ? n : 1;
  }
  FcSk sk;
  sk.m_tiles = odw_cdiv(M, 2 * kBM);
  sk.n_tiles = odw_cdiv(N, kBN);
  sk.n_ptiles = sk.m_tiles * sk.n_tiles;
  sk.kiters0 = odw_cdiv(o0.K, tc::kTileK);
  sk.kiters = sk.kiters0 + (o1.K > 0 ? odw_cdiv(o1.K, tc::kTileK) : 0);
  static const int panel_env = fc_env("ODWSCL_FC_PANEL", 8);
  sk.panel = min(sk.n_tiles, max(panel_env, 1));
  int cap = min(max_pairs, (ODW_NUM_SMS - odw_sm_margin()) / 2);
  if (max_pairs_cap > 0) cap = min(cap, max_pairs_cap);
  if (sk.n_ptiles * 2 <= cap) {
    // few tiles (Sim_Net's 128-wide layer, the predictor heads): every tile is split along K over cap / tiles pairs
    sk.rounds = 0;
    sk.left = sk.n_ptiles;
    sk.slices = max(1, min(cap / sk.left, min(8, sk.kiters)));
    sk.np = sk.left * sk.slices;
  } else {
    sk.np = min(cap, sk.n_ptiles);
    sk.rounds = sk.n_ptiles / sk.np;
    sk.left = sk.n_ptiles - sk.rounds * sk.np;
    sk.slices = sk.left > 0 ? min(sk.np / sk.left, min(8, sk.kiters)) : 0;
    if (sk.slices <= 1) {                 // more than half a round left: run it as one more (partly idle) round
      sk.rounds += sk.left > 0 ? 1 : 0;
      sk.left = 0;
      sk.slices = 0;
    }
  }
  // sized once for the worst case (left * slices <= resident pairs): no cudaMalloc / cudaFree -- both synchronise the
  // device -- when a later call has a different shape
  const size_t ws_floats = (size_t)(ODW_NUM_SMS / 2) * 2 * kBM * kBN;
  const size_t need = ws_floats * sizeof(float) + (size_t)(ODW_NUM_SMS / 2 + 1) * 2 * sizeof(int);
  if (g_fc_ws_bytes < need) {
    if (g_fc_ws) cudaFree(g_fc_ws);
    ODW_CUDA(cudaMalloc(&g_fc_ws, need));
    g_fc_ws_bytes = need;
  }
  sk.ws = reinterpret_cast<float*>(g_fc_ws);
  sk.flags = reinterpret_cast<int*>(reinterpret_cast<char*>(g_fc_ws) + ws_floats * sizeof(float));
  if (sk.left > 0) ODW_CUDA(cudaMemsetAsync(sk.flags, 0, (size_t)(sk.left + 1) * 2 * sizeof(int), st));
  cfg.gridDim = dim3(2 * sk.np);
  ODW_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, ma1, mb1, C, M, N, ldc, ep, sk));
  return 0;
}

// ODWSCL_FC_STAGES = 6 | 7 (shared-memory ring depth, 32 KB per stage per CTA), ODWSCL_FC_PANEL = raster panel width
template <bool AMN, bool BMN>
int launch_fc(const FcOperands& o0, const FcOperands& o1, float* C, int ldc, int M, int N, const FcEpi& ep,
              int max_pairs_cap, cudaStream_t st) {
  static const int stages = fc_env("ODWSCL_FC_STAGES", 6);
  if (stages >= 7 && !(ep.flags & kPeerSum)) return launch_fc_s<AMN, BMN, 7>(o0, o1, C, ldc, M, N, ep, max_pairs_cap, st);
  return launch_fc_s<AMN, BMN, 6>(o0, o1, C, ldc, M, N, ep, max_pairs_cap, st);
}

// out[c] (+)= sum_r x[r, c]  (bias gradients: db = sum over the batch rows of dZ)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long rows, int cols, int ld, float* __restrict__ out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int py = threadIdx.x >> 5;
  float acc = 0.f;
  if (c < cols)
    for (long long r = (long long)blockIdx.y * 8 + py; r < rows; r += (long long)gridDim.y * 8) acc += x[r * ld + c];
  __shared__ float sm[8][33];
  sm[py][threadIdx.x & 31] = acc;
  __syncthreads();
  if (py == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += sm[j][threadIdx.x & 31];
    atomicAdd(out + c, t);
  }
}

}  // namespace

static int fc_gemm_entry(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major, float* C, int ldc,
                         int M, int N, int K, int flags, const float* bias, const float* mask_src, int ld_mask,
                         float mask_scale, float dropout_p, unsigned long long seed, int max_pairs, const float* A2, int lda2,
                         const float* B2, int ldb2, int K2, float* C_multicast, float out_scale, odwscl_stream_t stream,
                         const void* const* peer_C = nullptr, int n_peers = 0, int rows_per_owner = 0) {
  if (M < 0 || N < 0 || K < 0 || K2 < 0 || lda <= 0 || ldb <= 0 || ldc < N) return ODWSCL_EINVAL;
  if ((lda & 3) || (ldb & 3)) return ODWSCL_EINVAL;                      // TMA: 16-byte row pitch
  if (a_mn_major ? lda < M : lda < K) return ODWSCL_EINVAL;
  if (b_mn_major ? ldb < N : ldb < K) return ODWSCL_EINVAL;
  if (dropout_p < 0.f || dropout_p >= 1.f) return ODWSCL_EINVAL;
  if (flags & kPeerOwner) {                                              // reduce-scatter over peer-mapped replicas
    if (!(flags & kPeerSum) || (flags & ~(kPeerSum | kPeerOwner)) || !peer_C || n_peers < 1 || n_peers > 8 ||
        rows_per_owner <= 0 || (rows_per_owner & 31))
      return ODWSCL_EINVAL;
    for (int r = 0; r < n_peers; ++r)
      if (!peer_C[r] || ((uintptr_t)peer_C[r] & 15)) return ODWSCL_EINVAL;
  } else if (flags & kPeerSum) {                                         // a plain sum of products, nothing else fused
    if (!C_multicast || ((uintptr_t)C_multicast & 15) || (flags & ~kPeerSum)) return ODWSCL_EINVAL;
  }
  if (M == 0 || N == 0) return 0;
  if (!A || !B || !C || K == 0) return ODWSCL_EINVAL;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15) || ((uintptr_t)mask_src & 15)) return ODWSCL_EINVAL;
  if (K2 > 0) {
    if (!A2 || !B2 || (lda2 & 3) || (ldb2 & 3) || ((uintptr_t)A2 & 15) || ((uintptr_t)B2 & 15)) return ODWSCL_EINVAL;
    if ((a_mn_major ? lda2 < M : lda2 < K2) || (b_mn_major ? ldb2 < N : ldb2 < K2)) return ODWSCL_EINVAL;
  }
  if ((flags & kBias) && !bias) return ODWSCL_EINVAL;
  if ((flags & kMask) && (!mask_src || ld_mask < N)) return ODWSCL_EINVAL;
  FcEpi ep;
  ep.bias = bias; ep.mask_src = mask_src; ep.ld_mask = ld_mask; ep.mask_scale = mask_scale;
  ep.drop_scale = 1.f / (1.f - dropout_p);
  ep.drop_thr16 = (uint32_t)(dropout_p * 65536.0f);
  ep.seed = seed; ep.flags = flags;
  ep.mc_delta = ((flags & kPeerSum) && C_multicast) ? (long long)(reinterpret_cast<char*>(C_multicast) - reinterpret_cast<char*>(C)) : 0;
  ep.out_scale = out_scale;
  ep.rows_per_owner = rows_per_owner; ep.n_owners = n_peers;
  for (int r = 0; r < 8; ++r)
    ep.peer_delta[r] = (r < n_peers) ? (long long)(reinterpret_cast<const char*>(peer_C[r]) - reinterpret_cast<const char*>(C)) : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const FcOperands o0{A, lda, B, ldb, K}, o1{A2, lda2, B2, ldb2, K2};
  if (!a_mn_major && !b_mn_major) return launch_fc<false, false>(o0, o1, C, ldc, M, N, ep, max_pairs, st);
  if (!a_mn_major && b_mn_major) return launch_fc<false, true>(o0, o1, C, ldc, M, N, ep, max_pairs, st);
  if (a_mn_major && b_mn_major) return launch_fc<true, true>(o0, o1, C, ldc, M, N, ep, max_pairs, st);
  return launch_fc<true, false>(o0, o1, C, ldc, M, N, ep, max_pairs, st);
}

ODW_API int odwscl_fc_gemm_tf32(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major, float* C,
                                int ldc, int M, int N, int K, int flags, const float* bias, const float* mask_src,
                                int ld_mask, float mask_scale, float dropout_p, unsigned long long seed, int max_pairs,
                                const float* A2, int lda2, const float* B2, int ldb2, int K2, odwscl_stream_t stream) {
  if (flags & kPeerSum) return ODWSCL_EINVAL;
  return fc_gemm_entry(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, M, N, K, flags, bias, mask_src, ld_mask, mask_scale,
                       dropout_p, seed, max_pairs, A2, lda2, B2, ldb2, K2, nullptr, 1.f, stream);
}

ODW_API int odwscl_fc_gemm_peer_sum_tf32(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major,
                                         float* C, float* C_multicast, int ldc, int M, int N, int K, float out_scale,
                                         int max_pairs, const float* A2, int lda2, const float* B2, int ldb2, int K2,
                                         odwscl_stream_t stream) {
  return fc_gemm_entry(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, M, N, K, kPeerSum, nullptr, nullptr, 0, 1.f, 0.f, 0ull,
                       max_pairs, A2, lda2, B2, ldb2, K2, C_multicast, out_scale, stream);
}

ODW_API int odwscl_fc_gemm_peer_scatter_tf32(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major,
                                             float* C, const void* const* peer_C, int n_peers, int rows_per_owner, int ldc,
                                             int M, int N, int K, float out_scale, int max_pairs, const float* A2, int lda2,
                                             const float* B2, int ldb2, int K2, odwscl_stream_t stream) {
  return fc_gemm_entry(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, M, N, K, kPeerSum | kPeerOwner, nullptr, nullptr, 0, 1.f,
                       0.f, 0ull, max_pairs, A2, lda2, B2, ldb2, K2, nullptr, out_scale, stream, peer_C, n_peers,
                       rows_per_owner);
}

// every rank's replica[i] = src[i] (one NVLink store, replicated by the switch): the all-gather half after the
// reduce-scatter above -- each rank broadcasts the rows it owns
__global__ void peer_broadcast_kernel(const float4* __restrict__ src, float* __restrict__ mc, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(src + i);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
  }
}

ODW_API int odwscl_peer_broadcast_f32(const float* src, float* dst_multicast, long long n, odwscl_stream_t stream) {
  if (n < 0 || (n & 3)) return ODWSCL_EINVAL;
  if (n == 0) return 0;
  if (!src || !dst_multicast || ((uintptr_t)src & 15) || ((uintptr_t)dst_multicast & 15)) return ODWSCL_EINVAL;
  const int blocks = (int)min((long long)(ODW_NUM_SMS - odw_sm_margin()) * 8, (n / 4 + 255) / 256);
  peer_broadcast_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src), dst_multicast, n / 4);
  ODW_LAUNCH_CHECK();
  return 0;
}

// every rank's replica[i] += scale * src[i] through the multicast alias (the path for a gradient that was not produced by
// the GEMM above: strict-mode chunks, a layer applied more than twice)
__global__ void peer_add_kernel(const float4* __restrict__ src, float* __restrict__ mc, long long n4, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(src + i);
    multimem_red_add_v4(mc + 4 * i, v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
}

ODW_API int odwscl_peer_add_f32(const float* src, float* dst_multicast, long long n, float scale, odwscl_stream_t stream) {
  if (n < 0 || (n & 3)) return ODWSCL_EINVAL;
  if (n == 0) return 0;
  if (!src || !dst_multicast || ((uintptr_t)src & 15) || ((uintptr_t)dst_multicast & 15)) return ODWSCL_EINVAL;
  const int blocks = (int)min((long long)ODW_NUM_SMS * 4, (n / 4 + 255) / 256);
  peer_add_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src), dst_multicast, n / 4, scale);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_colsum_f32(const float* x, long long rows, int cols, int ld, float* out, int accumulate,
                              odwscl_stream_t stream) {
  if (rows < 0 || cols < 0 || ld < cols) return ODWSCL_EINVAL;
  if (cols == 0) return 0;
  if (!out || (rows > 0 && !x)) return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) ODW_CUDA(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), st));
  if (rows == 0) return 0;
  dim3 grid(odw_cdiv(cols, 32), (unsigned)min((long long)ODW_NUM_SMS * 8 / odw_cdiv(cols, 32) + 1, (rows + 7) / 8));
  colsum_kernel<<<grid, 256, 0, st>>>(x, rows, cols, ld, out);
  ODW_LAUNCH_CHECK();
  return 0;
}
