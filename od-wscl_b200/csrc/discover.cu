// discover.cu -- object discovery of the contrastive head, on device, without host syncs
// (Appendix A).  Reference: roi_heads/weak_head/loss.py:271-345 (two
// triple-nested Python loops issuing ~20-40 tiny kernels and >15 host syncs per (img, branch,
// class) iteration) and weak_head/pseudo_label_generator.py:135-197 (N x G IoU pulled to the
// host through numpy).
//
// Work decomposition: a "pair" = (image b, positive class c).  Pairs are independent except for
// the append order of the SupCon bank, which odwscl_bank_assemble reconstructs afterwards, so
//   phase A  : 1 CTA per pair  (3 argmaxes, 3 IoU rows, union, ordered compaction, hardness)
//   phase B  : select: 1 CTA per (pair, branch) (tau, similarity rows, selection rule incl. the
//              bool-vs-float quirk of loss.py:327, bitonic sort + greedy NMS sweep in shared
//              memory, fallback); merge: 1 CTA per pair, branches in order (ordered set difference
//              against the running membership, membership update)
//   od_layer : 1 CTA per (image, branch)
// Similarity rows are computed only for the <= |pos| query proposals the rule actually reads
// (the reference materialises the whole N x N product per iteration, loss.py:319).
#include "cta_utils.cuh"

namespace {

using odw::kCtaThreads;
constexpr int kD = ODWSCL_SIM_DIM;
constexpr int kMaxPairs = 256;   // (image, positive class) pairs per call: 8 images x 32 classes

// block-wide first-max argmax: larger value wins, ties go to the smaller index (torch.argmax /
// numpy argmax on the CPU reference).  s_v / s_i: 32 entries each.
__device__ __forceinline__ int cta_argmax_first(float v, int idx, float* s_v, int* s_i) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  if (lane == 0) { s_v[wid] = v; s_i[wid] = idx; }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    v = lane < nw ? s_v[lane] : -INFINITY;
    idx = lane < nw ? s_i[lane] : INT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) s_i[0] = idx;
  }
  __syncthreads();
  const int r = s_i[0] == INT_MAX ? 0 : s_i[0];   // all-NaN column (diverged training): stay in range, do not fault
  __syncthreads();
  return r;
}

__device__ __forceinline__ float cta_sum(float v, float* s_v) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = odw_warp_sum(v);
  if (lane == 0) s_v[wid] = v;
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    v = lane < nw ? s_v[lane] : 0.f;
    v = odw_warp_sum(v);
    if (lane == 0) s_v[0] = v;
  }
  __syncthreads();
  const float r = s_v[0];
  __syncthreads();
  return r;
}

__device__ __forceinline__ int col_argmax(const float* __restrict__ S, int off, int N, int C, int col,
                                          float* s_v, int* s_i) {
  float bv = -INFINITY;
  int bi = INT_MAX;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const float v = __ldg(S + (size_t)(off + j) * C + col);
    if (v > bv) { bv = v; bi = j; }          // ascending j per thread: first max kept
  }
  return cta_argmax_first(bv, bi, s_v, s_i);
}

// dot(F[row], s_q) with one warp; every lane gets the result.
__device__ __forceinline__ float warp_dot128(const float* __restrict__ row, const float* s_q, int lane) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(row) + lane);
  const float4 q = reinterpret_cast<const float4*>(s_q)[lane];
  float p = a.x * q.x;
  p = fmaf(a.y, q.y, p);
  p = fmaf(a.z, q.z, p);
  p = fmaf(a.w, q.w, p);
  return odw_warp_sum(p);
}

// ------------------------------------------------------------------------------ phase A
__global__ void __launch_bounds__(kCtaThreads, 1)
phase_a_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ img_off, int C,
               const float* __restrict__ s0, const float* __restrict__ s1, const float* __restrict__ s2,
               const int32_t* __restrict__ pair_img, const int32_t* __restrict__ pair_cls, int Ncap,
               float thres, int32_t* __restrict__ amax, uint8_t* __restrict__ member,
               int32_t* __restrict__ cntA, float* __restrict__ colsum) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint8_t* s_flag = smem;                         // [Ncap]
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  const int p = blockIdx.x;
  const int b = pair_img[p], col = pair_cls[p] + 1;
  const int off = img_off[b], N = img_off[b + 1] - off;
  for (int j = threadIdx.x; j < Ncap; j += blockDim.x) s_flag[j] = 0;
  __syncthreads();
  if (N <= 0) {                                     // empty image: nothing to seed
    for (int j = threadIdx.x; j < Ncap; j += blockDim.x) member[(size_t)p * Ncap + j] = 0;
    if (threadIdx.x == 0) { colsum[p] = 0.f; cntA[p] = 0; amax[p * 3] = amax[p * 3 + 1] = amax[p * 3 + 2] = 0; }
    return;
  }
  const float* S[3] = {s0, s1, s2};
  for (int i = 0; i < 3; ++i) {
    const int m = col_argmax(S[i], off, N, C, col, s_v, s_i);       // loss.py:286
    if (threadIdx.x == 0) amax[p * 3 + i] = m;
    const float4 bm = __ldg(boxes + off + m);
    for (int j = threadIdx.x; j < N; j += blockDim.x)               // utils/utils.py:23-26
      if (odw_iou(__ldg(boxes + off + j), bm, 1.f) >= thres) s_flag[j] = 1;
    __syncthreads();
  }
  float part = 0.f;
  int cnt = 0;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    part += __ldg(s0 + (size_t)(off + j) * C + col);
    cnt += s_flag[j];
  }
  for (int j = threadIdx.x; j < Ncap; j += blockDim.x) member[(size_t)p * Ncap + j] = s_flag[j];
  const float tot = cta_sum(part, s_v);
  const float fc = cta_sum((float)cnt, s_v);
  if (threadIdx.x == 0) { colsum[p] = tot; cntA[p] = (int)fc; }
}

__global__ void __launch_bounds__(kCtaThreads, 1)
phase_a_rows_kernel(const int32_t* __restrict__ img_off, int C, const float* __restrict__ s0,
                    const int32_t* __restrict__ pair_img, const int32_t* __restrict__ pair_cls, int P,
                    int Ncap, const uint8_t* __restrict__ member, const int32_t* __restrict__ cntA,
                    const float* __restrict__ colsum, int32_t* __restrict__ offA,
                    int32_t* __restrict__ rowsA, float* __restrict__ hardA) {
  __shared__ int s_scan[64];
  const int p = blockIdx.x;
  int base = 0;
  for (int q = 0; q < p; ++q) base += cntA[q];
  if (threadIdx.x == 0) {
    offA[p] = base;
    if (p == P - 1) offA[P] = base + cntA[p];
  }
  const int b = pair_img[p], col = pair_cls[p] + 1;
  const int off = img_off[b], N = img_off[b + 1] - off;
  const float cs = colsum[p];
  const uint8_t* mem = member + (size_t)p * Ncap;
  odw::cta_compact(
      N, s_scan, [&](int j) { return mem[j] != 0; },
      [&](int k, int j) {
        rowsA[base + k] = off + j;
        hardA[base + k] = __fdiv_rn(__ldg(s0 + (size_t)(off + j) * C + col), cs);    // loss.py:294
      });
}

// ------------------------------------------------------------------------------ phase B
struct PhaseBSmem {
  float* fq;            // [128] query embedding
  float* sim;           // [Ncap]  (aliases key: the similarity row is dead once `close` is decided)
  uint8_t* close;       // [Ncap]
  float* key;           // [L]
  int* id;              // [L]
  float4* box;          // [L]
  uint8_t* sup;         // [L]
  int* scan;            // [64]
};

__device__ __forceinline__ PhaseBSmem carve_b(unsigned char* smem, int Ncap, int L) {
  PhaseBSmem s;
  s.box = reinterpret_cast<float4*>(smem);
  s.fq = reinterpret_cast<float*>(s.box + L);
  s.key = s.fq + kD;
  s.sim = s.key;                                   // L >= Ncap
  s.id = reinterpret_cast<int*>(s.key + L);
  s.scan = s.id + L;
  s.close = reinterpret_cast<uint8_t*>(s.scan + 64);
  s.sup = s.close + Ncap;
  return s;
}
static size_t phase_b_smem_bytes(int Ncap, int L) {
  return (size_t)L * 16 + kD * 4 + (size_t)L * 8 + 64 * 4 + (size_t)Ncap + L;     // 213.8 KB at Ncap = 8192
}

// phase_b_select_kernel: grid (P, 3) -- one CTA per (pair, refinement branch).  Everything that is independent of
// the running membership: tau, the similarity row(s), the selection rule, sort + greedy NMS -> inst[p][i].
__global__ void __launch_bounds__(kCtaThreads, 1)
phase_b_select_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ img_off, int R, int C,
                      const float* __restrict__ s0, const float* __restrict__ s1, const float* __restrict__ s2,
                      const int32_t* __restrict__ pair_img, const int32_t* __restrict__ pair_cls, int P, int Ncap,
                      int L, const float* __restrict__ F, const float* __restrict__ E,
                      const int32_t* __restrict__ amax, const int32_t* __restrict__ offA,
                      const int32_t* __restrict__ rowsA, float nms_thr, int32_t* __restrict__ inst,
                      int32_t* __restrict__ inst_cnt, float* __restrict__ tau_out,
                      const float* __restrict__ sim_rows_in) {
  extern __shared__ __align__(16) unsigned char smem[];
  const PhaseBSmem s = carve_b(smem, Ncap, L);
  __shared__ float s_v[32];
  const int p = blockIdx.x, i = blockIdx.y;
  const int b = pair_img[p], cls = pair_cls[p], col = cls + 1;
  const int off = img_off[b], N = img_off[b + 1] - off;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int K = offA[P];
  int q_lo = p, q_hi = p + 1;                      // pairs of the same image (contiguous)
  while (q_lo > 0 && pair_img[q_lo - 1] == b) --q_lo;
  while (q_hi < P && pair_img[q_hi] == b) ++q_hi;
  const float* Si = i == 0 ? s0 : (i == 1 ? s1 : s2);
  if (N <= 0) {
    if (threadIdx.x == 0) inst_cnt[p * 3 + i] = 0;
    return;
  }
  const int m = amax[p * 3 + i];
  if (threadIdx.x < kD) s.fq[threadIdx.x] = __ldg(F + (size_t)(off + m) * kD + threadIdx.x);
  __syncthreads();
  // ---- tau = mean_r( F[m] . coll[cls][r] )   (loss.py:320); coll = Phase-A bank of the class
  float part = 0.f;
  int nrows = 0;
  for (int q = 0; q < P; ++q) {
    if (pair_cls[q] != cls) continue;
    const int lo = offA[q], cnt = offA[q + 1] - lo;
    nrows += 3 * cnt;
    for (int t = wid; t < 3 * cnt; t += nwarps) {
      const int seg = t / cnt, k = t - seg * cnt;
      const float* row = seg == 0 ? F + (size_t)__ldg(rowsA + lo + k) * kD
                                  : E + (size_t)((seg - 1) * K + lo + k) * kD;
      const float d = warp_dot128(row, s.fq, lane);
      if (lane == 0) part += d;
    }
  }
  const float tau = __fdiv_rn(cta_sum(part, s_v), (float)nrows);
  if (threadIdx.x == 0) tau_out[p * 3 + i] = tau;
  // ---- similarity row of m and the `>= tau` rule (loss.py:324/330)
  for (int j = wid; j < N; j += nwarps) {
    float sv;
    if (sim_rows_in) sv = __ldg(sim_rows_in + (size_t)(p * 3 + i) * Ncap + j);
    else sv = warp_dot128(F + (size_t)(off + j) * kD, s.fq, lane);
    if (lane == 0) { s.sim[j] = sv; s.close[j] = sv >= tau ? 1 : 0; }
  }
  __syncthreads();
  // ---- other positive classes of the image: close <- (float(close) >= Sim[m_n])  (loss.py:325-327)
  for (int q = q_lo; q < q_hi; ++q) {
    if (q == p) continue;
    const int mn = amax[q * 3 + i];
    if (threadIdx.x < kD) s.fq[threadIdx.x] = __ldg(F + (size_t)(off + mn) * kD + threadIdx.x);
    __syncthreads();
    for (int j = wid; j < N; j += nwarps) {
      const float sn = warp_dot128(F + (size_t)(off + j) * kD, s.fq, lane);
      if (lane == 0) s.close[j] = ((s.close[j] ? 1.f : 0.f) >= sn) ? 1 : 0;
    }
    __syncthreads();
  }
  // ---- candidates (ascending j) keyed by the class score of this branch
  const int ncl = odw::cta_compact(
      N, s.scan, [&](int j) { return s.close[j] != 0; },
      [&](int k, int j) { s.key[k] = __ldg(Si + (size_t)(off + j) * C + col); s.id[k] = j; });
  int* inst_p = inst + (size_t)(p * 3 + i) * Ncap;
  int nk = 0;
  if (ncl > 0) {
    const int L2 = odw::next_pow2(ncl < 32 ? 32 : ncl);
    for (int t = ncl + threadIdx.x; t < L2; t += blockDim.x) { s.key[t] = -INFINITY; s.id[t] = INT_MAX; }
    for (int t = threadIdx.x; t < L2; t += blockDim.x) s.sup[t] = 0;
    __syncthreads();
    odw::cta_bitonic_sort(s.key, s.id, L2);
    for (int t = threadIdx.x; t < ncl; t += blockDim.x) s.box[t] = __ldg(boxes + off + s.id[t]);
    __syncthreads();
    nk = odw::cta_nms_sweep(s.box, s.sup, ncl, nms_thr, 0.f,                 // loss.py:332
                            [&](int k, int pos) { inst_p[k] = s.id[pos]; });
  }
  if (nk == 0) {                                                             // loss.py:333
    if (threadIdx.x == 0) inst_p[0] = m;
    nk = 1;
  }
  if (threadIdx.x == 0) inst_cnt[p * 3 + i] = nk;
}

// phase_b_merge_kernel: grid P -- the order-dependent tail, per pair, branches in sequence:
// new = sorted(kept \ member); fallback [m]; member |= new   (loss.py:336-341)
__global__ void __launch_bounds__(kCtaThreads, 1)
phase_b_merge_kernel(const int32_t* __restrict__ img_off, int C, const float* __restrict__ s0,
                     const int32_t* __restrict__ pair_img, const int32_t* __restrict__ pair_cls, int Ncap,
                     const int32_t* __restrict__ amax, uint8_t* __restrict__ member,
                     const float* __restrict__ colsum, const int32_t* __restrict__ inst,
                     const int32_t* __restrict__ inst_cnt, int32_t* __restrict__ newl,
                     int32_t* __restrict__ new_cnt, float* __restrict__ hardB) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint8_t* s_close = smem;                                   // [Ncap]
  __shared__ int s_scan[64];
  const int p = blockIdx.x;
  const int b = pair_img[p], col = pair_cls[p] + 1;
  const int off = img_off[b], N = img_off[b + 1] - off;
  const float cs = colsum[p];
  uint8_t* mem = member + (size_t)p * Ncap;
  if (N <= 0) {
    if (threadIdx.x < 3) new_cnt[p * 3 + threadIdx.x] = 0;
    return;
  }
  for (int i = 0; i < 3; ++i) {
    const int m = amax[p * 3 + i];
    const int* inst_p = inst + (size_t)(p * 3 + i) * Ncap;
    const int nk = inst_cnt[p * 3 + i];
    for (int j = threadIdx.x; j < N; j += blockDim.x) s_close[j] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < nk; k += blockDim.x) s_close[inst_p[k]] = 1;
    __syncthreads();
    int* new_p = newl + (size_t)(p * 3 + i) * Ncap;
    float* hard_p = hardB + (size_t)(p * 3 + i) * Ncap;
    int nn = odw::cta_compact(
        N, s_scan, [&](int j) { return s_close[j] != 0 && mem[j] == 0; },
        [&](int k, int j) {
          new_p[k] = j;
          hard_p[k] = __fdiv_rn(__ldg(s0 + (size_t)(off + j) * C + col), cs);      // loss.py:343
        });
    if (nn == 0) {
      if (threadIdx.x == 0) {
        new_p[0] = m;
        hard_p[0] = __fdiv_rn(__ldg(s0 + (size_t)(off + m) * C + col), cs);
      }
      nn = 1;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nn; k += blockDim.x) mem[new_p[k]] = 1;
    if (threadIdx.x == 0) new_cnt[p * 3 + i] = nn;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ bank assembly
// kind 0: F rows of phase A, 1: drop, 2: noise, 3: phase B.  Packed (12 bytes) so 2 x 6 x kMaxPairs fit static smem.
struct Seg {
  int kqi, dst, len;
  __device__ Seg() {}
  __device__ Seg(int kind, int q_, int i_, int dst_, int len_) : kqi(kind | (i_ << 2) | (q_ << 4)), dst(dst_), len(len_) {}
  __device__ int src_kind() const { return kqi & 3; }
  __device__ int i() const { return (kqi >> 2) & 3; }
  __device__ int q() const { return kqi >> 4; }
};

__global__ void __launch_bounds__(kCtaThreads, 1)
bank_assemble_kernel(const int32_t* __restrict__ pair_img, const int32_t* __restrict__ pair_cls, int P, int B,
                     int R, int Ncap, int nfg, const int32_t* __restrict__ img_off,
                     const int32_t* __restrict__ offA, const int32_t* __restrict__ rowsA,
                     const float* __restrict__ hardA, const int32_t* __restrict__ newl,
                     const int32_t* __restrict__ new_cnt, const float* __restrict__ hardB, int Mcap,
                     int32_t* __restrict__ row_src, int32_t* __restrict__ row_lab, float* __restrict__ row_w,
                     int32_t* __restrict__ M_out) {
  __shared__ Seg s_bank[6 * kMaxPairs];
  __shared__ Seg s_w[6 * kMaxPairs];
  __shared__ int s_nb, s_nw;
  const int K = offA[P];
  if (threadIdx.x == 0) {
    int nb = 0, dst = 0;
    for (int c = 0; c < nfg; ++c) {                       // class-major rows (sim_loss.py:55-58)
      for (int q = 0; q < P; ++q) {
        if (pair_cls[q] != c) continue;
        const int len = offA[q + 1] - offA[q];
        for (int k = 0; k < 3; ++k) { s_bank[nb++] = Seg(k, q, 0, dst, len); dst += len; }
      }
      for (int q = 0; q < P; ++q) {
        if (pair_cls[q] != c) continue;
        for (int i = 0; i < 3; ++i) {
          const int len = new_cnt[q * 3 + i];
          s_bank[nb++] = Seg(3, q, i, dst, len); dst += len;
        }
      }
    }
    int nw = 0, wd = 0;
    for (int q = 0; q < P; ++q) {                         // execution-order weights (loss.py:296-305)
      const int len = offA[q + 1] - offA[q];
      for (int k = 0; k < 3; ++k) { s_w[nw++] = Seg(k, q, 0, wd, len); wd += len; }
    }
    int q0 = 0;
    while (q0 < P) {                                      // loss.py:311-345: image, branch, class
      int q1 = q0;
      while (q1 < P && pair_img[q1] == pair_img[q0]) ++q1;
      for (int i = 0; i < 3; ++i)
        for (int q = q0; q < q1; ++q) {
          const int len = new_cnt[q * 3 + i];
          s_w[nw++] = Seg(3, q, i, wd, len); wd += len;
        }
      q0 = q1;
    }
    s_nb = nb; s_nw = nw;
    M_out[0] = dst < Mcap ? dst : Mcap;
    M_out[1] = dst;                                       // unclamped: the caller sees a bound that was too small
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int sidx = wid; sidx < s_nb; sidx += nwarps) {
    const Seg g = s_bank[sidx];
    const int gq = g.q(), gi = g.i(), kind = g.src_kind();
    const int cls = pair_cls[gq];
    const int off = img_off[pair_img[gq]];
    for (int k = lane; k < g.len; k += 32) {
      if (g.dst + k >= Mcap) break;
      int src;
      if (kind == 0) src = rowsA[offA[gq] + k];
      else if (kind == 1) src = R + offA[gq] + k;
      else if (kind == 2) src = R + K + offA[gq] + k;
      else src = off + newl[(size_t)(gq * 3 + gi) * Ncap + k];
      row_src[g.dst + k] = src;
      row_lab[g.dst + k] = cls;
    }
  }
  for (int sidx = wid; sidx < s_nw; sidx += nwarps) {
    const Seg g = s_w[sidx];
    const int gq = g.q(), gi = g.i(), kind = g.src_kind();
    for (int k = lane; k < g.len; k += 32) {
      if (g.dst + k >= Mcap) break;
      row_w[g.dst + k] = kind < 3 ? hardA[offA[gq] + k] : hardB[(size_t)(gq * 3 + gi) * Ncap + k];
    }
  }
}

// ------------------------------------------------------------------------------ od_layer
constexpr int kGtChunk = 1024;

__global__ void __launch_bounds__(kCtaThreads, 1)
od_layer_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ img_off, int R, int C,
                const float* __restrict__ s0, const float* __restrict__ s1, const float* __restrict__ s2,
                const int32_t* __restrict__ pair_img, const int32_t* __restrict__ pair_cls, int P, int Ncap,
                const int32_t* __restrict__ inst, const int32_t* __restrict__ inst_cnt, float fg_thr,
                int64_t* __restrict__ labels, float* __restrict__ weights, float4* __restrict__ targets) {
  __shared__ float4 g_box[kGtChunk];
  __shared__ float g_score[kGtChunk];
  __shared__ int g_cls[kGtChunk];
  __shared__ int s_zero[kMaxPairs];               // rows zeroed so far (pseudo_label_generator.py:165)
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  const int b = blockIdx.x / 3, i = blockIdx.x % 3;
  const float* S = i == 0 ? s0 : (i == 1 ? s1 : s2);
  const int off = img_off[b], N = img_off[b + 1] - off;
  int q_lo = 0;
  while (q_lo < P && pair_img[q_lo] != b) ++q_lo;
  int q_hi = q_lo;
  while (q_hi < P && pair_img[q_hi] == b) ++q_hi;
  int64_t* lab = labels + (size_t)i * R + off;
  float* wts = weights + (size_t)i * R + off;
  float4* tgt = targets + (size_t)i * R + off;
  // gridDim.y CTAs share one (image, branch): each labels a contiguous slice of the proposals against ALL pseudo-GT
  // boxes (the class-argmax prelude is recomputed per CTA: N loads per class, negligible)
  const int slice = (N + gridDim.y - 1) / gridDim.y;
  const int j_lo = min(N, (int)blockIdx.y * slice), j_hi = min(N, j_lo + slice);
  if (q_lo == q_hi) {                               // no positive class: all background, zero weight
    for (int j = j_lo + threadIdx.x; j < j_hi; j += blockDim.x) { lab[j] = 0; wts[j] = 0.f; tgt[j] = make_float4(0, 0, 0, 0); }
    return;
  }
  // argmax of each class column on the progressively zeroed score matrix
  for (int q = q_lo; q < q_hi; ++q) {
    const int col = pair_cls[q] + 1;
    float bv = -INFINITY;
    int bi = INT_MAX;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      float v = __ldg(S + (size_t)(off + j) * C + col);
      for (int z = q_lo; z < q; ++z) if (s_zero[z - q_lo] == j) v = 0.f;
      if (v > bv) { bv = v; bi = j; }
    }
    const int m = cta_argmax_first(bv, bi, s_v, s_i);
    if (threadIdx.x == 0) s_zero[q - q_lo] = m;
    __syncthreads();
  }
  // running first-max over GT chunks
  const int per = (j_hi - j_lo + blockDim.x - 1) / blockDim.x;      // proposals per thread (<= 8 for N <= 8192)
  float best[8]; int barg[8]; float bsc[8]; int bcl[8]; float4 bgt[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) { best[u] = -1.f; barg[u] = 0; bsc[u] = 0.f; bcl[u] = 0; bgt[u] = make_float4(0, 0, 0, 0); }
  int q = q_lo, k = 0;                                     // cursor into (pair, position) GT stream
  while (q < q_hi) {
    // fill one chunk
    __syncthreads();
    int filled = 0;
    int qq = q, kk = k;
    while (qq < q_hi && filled < kGtChunk) {
      const int cnt = inst_cnt[qq * 3 + i];
      const int n_gt = cnt > 0 ? cnt : 1;
      const int take = min(n_gt - kk, kGtChunk - filled);
      const int col = pair_cls[qq] + 1;
      for (int t = threadIdx.x; t < take; t += blockDim.x) {
        const int j = cnt > 0 ? inst[(size_t)(qq * 3 + i) * Ncap + kk + t] : s_zero[qq - q_lo];
        float sc = __ldg(S + (size_t)(off + j) * C + col);
        for (int z = q_lo; z < qq; ++z) if (s_zero[z - q_lo] == j) sc = 0.f;
        g_box[filled + t] = __ldg(boxes + off + j);
        g_score[filled + t] = sc;
        g_cls[filled + t] = col;
      }
      filled += take; kk += take;
      if (kk >= n_gt) { ++qq; kk = 0; }
    }
    q = qq; k = kk;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j_lo + threadIdx.x + u * blockDim.x;
      if (u < per && j < j_hi) {
        const float4 pj = __ldg(boxes + off + j);
        for (int g = 0; g < filled; ++g) {
          const float v = odw_iou(pj, g_box[g], 1.f);
          if (v > best[u]) { best[u] = v; bsc[u] = g_score[g]; bcl[u] = g_cls[g]; bgt[u] = g_box[g]; }
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int j = j_lo + threadIdx.x + u * blockDim.x;
    if (u < per && j < j_hi) {
      const float4 pj = __ldg(boxes + off + j);
      lab[j] = best[u] <= fg_thr ? 0 : (int64_t)bcl[u];           // :183 (le)
      wts[j] = bsc[u];
      // BoxCoder(10,10,5,5).encode (modeling/box_coder.py:22-50)
      const float ew = __fadd_rn(__fsub_rn(pj.z, pj.x), 1.f), eh = __fadd_rn(__fsub_rn(pj.w, pj.y), 1.f);
      const float ex = __fadd_rn(pj.x, __fmul_rn(0.5f, ew)), ey = __fadd_rn(pj.y, __fmul_rn(0.5f, eh));
      const float4 gb = bgt[u];
      const float gw = __fadd_rn(__fsub_rn(gb.z, gb.x), 1.f), gh = __fadd_rn(__fsub_rn(gb.w, gb.y), 1.f);
      const float gx = __fadd_rn(gb.x, __fmul_rn(0.5f, gw)), gy = __fadd_rn(gb.y, __fmul_rn(0.5f, gh));
      float4 t;
      t.x = __fdiv_rn(__fmul_rn(10.f, __fsub_rn(gx, ex)), ew);
      t.y = __fdiv_rn(__fmul_rn(10.f, __fsub_rn(gy, ey)), eh);
      t.z = __fmul_rn(5.f, logf(__fdiv_rn(gw, ew)));
      t.w = __fmul_rn(5.f, logf(__fdiv_rn(gh, eh)));
      tgt[j] = t;
    }
  }
}

}  // namespace

ODW_API int odwscl_discover_phase_a_f32(const float* boxes, const int32_t* img_off, int B, int R, int C,
                                        const float* s0, const float* s1, const float* s2,
                                        const int32_t* pair_img, const int32_t* pair_cls, int P, int Ncap,
                                        float thres, int32_t* amax, uint8_t* member, int32_t* cntA,
                                        int32_t* offA, int32_t* rowsA, float* hardA, float* colsum,
                                        odwscl_stream_t stream) {
  if (B < 0 || R < 0 || C < 2 || P < 0 || P > kMaxPairs || Ncap < 0 || Ncap > 8192) return ODWSCL_EINVAL;
  if (P == 0) return 0;
  if (!boxes || !img_off || !s0 || !s1 || !s2 || !pair_img || !pair_cls || !amax || !member || !cntA || !offA ||
      !rowsA || !hardA || !colsum)
    return ODWSCL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  phase_a_kernel<<<P, kCtaThreads, Ncap, st>>>(reinterpret_cast<const float4*>(boxes), img_off, C, s0, s1, s2,
                                               pair_img, pair_cls, Ncap, thres, amax, member, cntA, colsum);
  ODW_LAUNCH_CHECK();
  phase_a_rows_kernel<<<P, kCtaThreads, 0, st>>>(img_off, C, s0, pair_img, pair_cls, P, Ncap, member, cntA,
                                                 colsum, offA, rowsA, hardA);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_discover_phase_b_f32(const float* boxes, const int32_t* img_off, int B, int R, int C,
                                        const float* s0, const float* s1, const float* s2,
                                        const int32_t* pair_img, const int32_t* pair_cls, int P, int Ncap,
                                        const float* F, const float* E, const int32_t* amax, uint8_t* member,
                                        const int32_t* cntA, const int32_t* offA, const int32_t* rowsA,
                                        const float* colsum, float nms_thr, int32_t* inst, int32_t* inst_cnt,
                                        int32_t* newl, int32_t* new_cnt, float* hardB, float* tau_out,
                                        const float* sim_rows_in, odwscl_stream_t stream) {
  (void)cntA; (void)B;
  if (R < 0 || C < 2 || P < 0 || P > kMaxPairs || Ncap <= 0 || Ncap > 8192) return ODWSCL_EINVAL;
  if (P == 0) return 0;
  if (!boxes || !img_off || !s0 || !s1 || !s2 || !pair_img || !pair_cls || !F || !E || !amax || !member || !offA ||
      !rowsA || !colsum || !inst || !inst_cnt || !newl || !new_cnt || !hardB || !tau_out)
    return ODWSCL_EINVAL;
  int L = 32;
  while (L < Ncap) L <<= 1;
  const size_t smem = phase_b_smem_bytes(Ncap, L);
  ODW_CUDA(cudaFuncSetAttribute(phase_b_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  phase_b_select_kernel<<<dim3(P, 3), kCtaThreads, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(boxes), img_off, R, C, s0, s1, s2, pair_img, pair_cls, P, Ncap, L, F, E,
      amax, offA, rowsA, nms_thr, inst, inst_cnt, tau_out, sim_rows_in);
  ODW_LAUNCH_CHECK();
  phase_b_merge_kernel<<<P, kCtaThreads, Ncap, (cudaStream_t)stream>>>(img_off, C, s0, pair_img, pair_cls, Ncap, amax,
                                                                        member, colsum, inst, inst_cnt, newl, new_cnt,
                                                                        hardB);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_bank_assemble(const int32_t* pair_img, const int32_t* pair_cls, int P, int B, int R, int Ncap,
                                 int num_fg_classes, const int32_t* img_off, const int32_t* cntA,
                                 const int32_t* offA, const int32_t* rowsA, const float* hardA,
                                 const int32_t* newl, const int32_t* new_cnt, const float* hardB, int Mcap,
                                 int32_t* row_src, int32_t* row_lab, float* row_w, int32_t* M_out,
                                 odwscl_stream_t stream) {
  (void)cntA;
  if (P < 0 || P > kMaxPairs || Mcap < 0) return ODWSCL_EINVAL;
  if (!M_out) return ODWSCL_EINVAL;
  if (P == 0) {
    ODW_CUDA(cudaMemsetAsync(M_out, 0, 2 * sizeof(int32_t), (cudaStream_t)stream));
    return 0;
  }
  if (!pair_img || !pair_cls || !img_off || !offA || !rowsA || !hardA || !newl || !new_cnt || !hardB || !row_src ||
      !row_lab || !row_w)
    return ODWSCL_EINVAL;
  bank_assemble_kernel<<<1, kCtaThreads, 0, (cudaStream_t)stream>>>(pair_img, pair_cls, P, B, R, Ncap,
                                                                    num_fg_classes, img_off, offA, rowsA, hardA,
                                                                    newl, new_cnt, hardB, Mcap, row_src, row_lab,
                                                                    row_w, M_out);
  ODW_LAUNCH_CHECK();
  return 0;
}

ODW_API int odwscl_od_layer_f32(const float* boxes, const int32_t* img_off, int B, int R, int C, const float* s0,
                                const float* s1, const float* s2, const int32_t* pair_img,
                                const int32_t* pair_cls, int P, int Ncap, const int32_t* inst,
                                const int32_t* inst_cnt, float fg_thr, int64_t* labels, float* weights,
                                float* targets, odwscl_stream_t stream) {
  if (B < 0 || R < 0 || C < 2 || P < 0 || P > kMaxPairs || Ncap < 0 || Ncap > 8192) return ODWSCL_EINVAL;
  if (B == 0 || R == 0) return 0;
  if (!boxes || !img_off || !s0 || !s1 || !s2 || !labels || !weights || !targets) return ODWSCL_EINVAL;
  if (P > 0 && (!pair_img || !pair_cls || !inst || !inst_cnt)) return ODWSCL_EINVAL;
  const int nsplit = max(1, min(16, odw_cdiv(Ncap, 128)));     // 128 proposals per CTA: every thread-proposal loop is short
  od_layer_kernel<<<dim3(B * 3, nsplit), kCtaThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(boxes), img_off, R, C, s0, s1, s2, pair_img, pair_cls, P, Ncap, inst,
      inst_cnt, fg_thr, labels, weights, reinterpret_cast<float4*>(targets));
  ODW_LAUNCH_CHECK();
  return 0;
}
