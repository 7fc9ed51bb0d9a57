// Direct fp32 path — exact-difference scores with gradients, any shape, shared (N,d) or per-query
// (B,N,d) support.  CUDA-core kernels for the latency-bound episodic-training regime
// (reference NWNet.forward, nwhead/nw.py:162-211: B=8 queries x N=n_way*n_shot supports) and for the
// direct `kernel(x, y)` call of NWNet.get_neighbors (nwhead/nw.py:248).
//
// Forward  = NWHead.forward (nwhead/nw.py:266-289) with the kernels of nwhead/kernel.py:13-44.
// Backward = closed form of its autograd; torch.cdist's backward yields zero at
//            coincident points, reproduced by the dist > 0 guard.

#include "nw_common.cuh"

namespace nw {
namespace direct {

__host__ __device__ inline bool kind_normalised(int k) {
  return k == NW_KIND_HYPERSPHERE || k == NW_KIND_COSINE || k == NW_KIND_CLIP;
}
__host__ __device__ inline bool kind_euclid(int k) { return k == NW_KIND_EUCLIDEAN || k == NW_KIND_HYPERSPHERE; }

// inv[r] = 1 / max(|x_r|, 1e-12)
__global__ void inv_norm_kernel(const float* __restrict__ x, long long rows, int d, float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* p = x + r * d;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) ss += p[c] * p[c];
  ss = warp_sum(ss);
  if (lane == 0) inv[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
}

// One warp per TILE of TQ queries x TS supports.  Every (query, support) pair is accumulated exactly as a
// one-warp-per-pair kernel would do it — lane l takes columns l, l+32, ... in order with one fmaf each, then a
// butterfly sum — so a pair's score does not depend on the tile shape or on its neighbours (the neighbour search
// relies on that: candidates scored one by one must reproduce the dense matrix bit for bit).  The tile only
// shares the LOADS: each q and s value is read once per TQ x TS pairs instead of once per pair (the pair-per-warp
// version moved 8 bytes through L1/L2 per fmaf and ran at 2 TFLOP/s).
// batched (per-query supports, TQ == 1): support row j of query b is s[(b * n_support + j) * d].
template <int TQ, int TS>
__global__ void __launch_bounds__(256) scores_kernel(int kind, float scale, const float* __restrict__ q,
                                                     int n_query, int d, const float* __restrict__ s,
                                                     long long n_support, int batched,
                                                     float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const long long tiles_s = (n_support + TS - 1) / TS;
  const long long tiles_q = (n_query + TQ - 1) / TQ;
  const long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= tiles_q * tiles_s) return;
  // shared support: neighbouring warps take the SAME supports and different queries, so the support rows are
  // fetched from HBM once (the whole query batch stays in L2); per-query supports: query-major
  const long long b0 = (batched ? tile / tiles_s : tile % tiles_q) * TQ;
  const long long j0 = (batched ? tile % tiles_s : tile / tiles_q) * TS;
  const float* qp[TQ];
  const float* sp[TS];
#pragma unroll
  for (int i = 0; i < TQ; ++i) qp[i] = q + min(b0 + i, (long long)n_query - 1) * d;  // edge rows: recomputed, not stored
#pragma unroll
  for (int j = 0; j < TS; ++j) sp[j] = s + ((batched ? b0 * n_support : 0) + min(j0 + j, n_support - 1)) * d;

  float iq[TQ], is[TS];
  if (direct::kind_normalised(kind)) {
    float qq[TQ], ss[TS];
#pragma unroll
    for (int i = 0; i < TQ; ++i) qq[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TS; ++j) ss[j] = 0.f;
    for (int c = lane; c < d; c += 32) {
#pragma unroll
      for (int i = 0; i < TQ; ++i) qq[i] = fmaf(qp[i][c], qp[i][c], qq[i]);
#pragma unroll
      for (int j = 0; j < TS; ++j) ss[j] = fmaf(sp[j][c], sp[j][c], ss[j]);
    }
#pragma unroll
    for (int i = 0; i < TQ; ++i) iq[i] = 1.0f / fmaxf(sqrtf(warp_sum(qq[i])), 1e-12f);
#pragma unroll
    for (int j = 0; j < TS; ++j) is[j] = 1.0f / fmaxf(sqrtf(warp_sum(ss[j])), 1e-12f);
  }

  float acc[TQ][TS];
#pragma unroll
  for (int i = 0; i < TQ; ++i)
#pragma unroll
    for (int j = 0; j < TS; ++j) acc[i][j] = 0.f;
  const bool euclid = direct::kind_euclid(kind);
  if (!direct::kind_normalised(kind)) {
#pragma unroll 4  // several support-row loads in flight per lane (the loop is load-latency bound); same order per pair
    for (int c = lane; c < d; c += 32) {
      float qv[TQ], sv[TS];
#pragma unroll
      for (int i = 0; i < TQ; ++i) qv[i] = qp[i][c];
#pragma unroll
      for (int j = 0; j < TS; ++j) sv[j] = sp[j][c];
#pragma unroll
      for (int i = 0; i < TQ; ++i)
#pragma unroll
        for (int j = 0; j < TS; ++j) {
          if (euclid) {
            const float df = qv[i] - sv[j];
            acc[i][j] = fmaf(df, df, acc[i][j]);
          } else {
            acc[i][j] = fmaf(qv[i], sv[j], acc[i][j]);
          }
        }
    }
  } else {
    for (int c = lane; c < d; c += 32) {
      float qv[TQ], sv[TS];
      // separately rounded products (no FMA contraction): identical rows must give exactly 0 (hypersphere)
#pragma unroll
      for (int i = 0; i < TQ; ++i) qv[i] = __fmul_rn(qp[i][c], iq[i]);
#pragma unroll
      for (int j = 0; j < TS; ++j) sv[j] = __fmul_rn(sp[j][c], is[j]);
#pragma unroll
      for (int i = 0; i < TQ; ++i)
#pragma unroll
        for (int j = 0; j < TS; ++j) {
          if (euclid) {
            const float df = __fsub_rn(qv[i], sv[j]);
            acc[i][j] = fmaf(df, df, acc[i][j]);
          } else {
            acc[i][j] = fmaf(qv[i], sv[j], acc[i][j]);
          }
        }
    }
  }
#pragma unroll
  for (int i = 0; i < TQ; ++i)
#pragma unroll
    for (int j = 0; j < TS; ++j) {
      float out = warp_sum(acc[i][j]);
      if (euclid) out = -sqrtf(out);
      else if (kind == NW_KIND_CLIP) out *= scale;
      if (lane == 0 && b0 + i < n_query && j0 + j < n_support) scores[(b0 + i) * n_support + j0 + j] = out;
    }
}

// Shared-support scores for large problems (euclidean / dot): a block of 8 warps owns 32 queries x 32 supports
// and walks the feature axis in 128-column slabs staged in shared memory (cp.async, double buffered); warp w
// keeps 8 x 16 pairs in registers.  Lane l still accumulates columns l, l+32, ... in order with one fmaf each and
// the same butterfly follows, so every pair gets the bits of the warp-tile kernel above; what changes is that an
// operand value is fetched from L2 once per 32 pairs (the warp-tile kernel was L2-bound on the query re-reads).
constexpr int SB_ROWS = 32;    // queries and supports per block
constexpr int SB_COLS = 128;   // columns per slab
constexpr int SB_STAGE_FLOATS = 2 * SB_ROWS * SB_COLS;
constexpr size_t SB_SMEM_BYTES = 2 * SB_STAGE_FLOATS * sizeof(float);

__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

template <bool EUCLID>
__global__ void __launch_bounds__(256, 1) scores_block_kernel(const float* __restrict__ q, int n_query, int d,
                                                              const float* __restrict__ s, long long n_support,
                                                              float* __restrict__ scores) {
  extern __shared__ __align__(16) float slab[];  // [stage][32 query rows | 32 support rows][128]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long tiles_q = (n_query + SB_ROWS - 1) / SB_ROWS;
  const long long b0 = ((long long)blockIdx.x % tiles_q) * SB_ROWS;  // support-major: neighbours share supports
  const long long j0 = ((long long)blockIdx.x / tiles_q) * SB_ROWS;

  auto load_stage = [&](int stage, int c0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = threadIdx.x + i * 256;  // 64 rows x 32 float4
      const int row = f >> 5;
      const int c = c0 + (f & 31) * 4;
      const float* src = row < SB_ROWS ? q + min(b0 + row, (long long)n_query - 1) * d
                                       : s + min(j0 + row - SB_ROWS, n_support - 1) * d;
      float* dst = slab + stage * SB_STAGE_FLOATS + row * SB_COLS + (f & 31) * 4;
      if (c < d) cp_async_16(dst, src + c);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);  // adds exact zeros
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int q_row0 = (warp >> 1) * 8, s_row0 = (warp & 1) * 16;
  float acc[8][16];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;

  const int n_slabs = (d + SB_COLS - 1) / SB_COLS;
  load_stage(0, 0);
  for (int sl = 0; sl < n_slabs; ++sl) {
    if (sl + 1 < n_slabs) {
      load_stage((sl + 1) & 1, (sl + 1) * SB_COLS);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* qs = slab + (sl & 1) * SB_STAGE_FLOATS + q_row0 * SB_COLS + lane;
    const float* ss = slab + (sl & 1) * SB_STAGE_FLOATS + (SB_ROWS + s_row0) * SB_COLS + lane;
#pragma unroll
    for (int kk = 0; kk < SB_COLS / 32; ++kk) {
      float qv[8], sv[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) qv[i] = qs[i * SB_COLS + kk * 32];
#pragma unroll
      for (int j = 0; j < 16; ++j) sv[j] = ss[j * SB_COLS + kk * 32];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (EUCLID) {
            const float df = qv[i] - sv[j];
            acc[i][j] = fmaf(df, df, acc[i][j]);
          } else {
            acc[i][j] = fmaf(qv[i], sv[j], acc[i][j]);
          }
        }
    }
    __syncthreads();  // this stage is refilled two iterations from now
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float keep = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float v = warp_sum(acc[i][j]);
      if (EUCLID) v = -sqrtf(v);
      if (lane == j) keep = v;
    }
    const long long b = b0 + q_row0 + i, col = j0 + s_row0 + lane;
    if (lane < 16 && b < n_query && col < n_support) scores[b * n_support + col] = keep;
  }
}

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) r += red[i];
  return r;
}

// one block per query: softmax statistics + per-class sums (fixed summation order) + log
__global__ void __launch_bounds__(256) aggregate_kernel(const float* __restrict__ scores,
                                                        const int64_t* __restrict__ labels, int labels_batched,
                                                        long long n_support, int n_classes,
                                                        float* __restrict__ logp, float* __restrict__ row_lse,
                                                        int32_t* __restrict__ status) {
  __shared__ float red[8];
  const long long b = blockIdx.x;
  const float* sc = scores + b * n_support;
  const int64_t* lab = labels + (labels_batched ? b * n_support : 0);
  float mx = __int_as_float(0xff800000);
  int bad = 0;
  for (long long j = threadIdx.x; j < n_support; j += blockDim.x) {
    mx = fmaxf(mx, sc[j]);
    const long long y = lab[j];
    if (y < 0 || y >= n_classes) ++bad;
  }
  if (bad) *reinterpret_cast<volatile int32_t*>(status) = 1;  // plain store: the flag may be mapped host memory
  mx = block_max(mx, red);
  float sum = 0.f;
  for (long long j = threadIdx.x; j < n_support; j += blockDim.x) sum += expf(sc[j] - mx);
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) row_lse[b] = mx + logf(sum);
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) {
    float acc = 0.f;
    for (long long j = 0; j < n_support; ++j)
      if (lab[j] == c) acc += expf(sc[j] - mx);
    logp[b * n_classes + c] = logf(acc * inv + 1e-12f);
  }
}

// Same result for LONG rows (n_support > AGG_BINS_MIN_N): the per-class loop above is O(N * C / 256) per thread
// (seconds at N = 1.28M, C = 1000); here every thread walks its supports once and adds exp(score - max) into a
// shared-memory class bin.  The float atomics make the summation order — not the value beyond the last bits — vary
// from run to run; short rows keep the fixed-order loop.
constexpr int AGG_BINS_MIN_N = 4096;
constexpr int AGG_BINS_MAX_C = 8192;

__global__ void __launch_bounds__(1024) aggregate_bins_kernel(const float* __restrict__ scores,
                                                             const int64_t* __restrict__ labels, int labels_batched,
                                                             long long n_support, int n_classes,
                                                             float* __restrict__ logp, float* __restrict__ row_lse,
                                                             int32_t* __restrict__ status) {
  extern __shared__ float bins[];  // n_classes floats
  __shared__ float red[32];
  const long long b = blockIdx.x;
  const float* sc = scores + b * n_support;
  const int64_t* lab = labels + (labels_batched ? b * n_support : 0);
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) bins[c] = 0.f;
  float mx = __int_as_float(0xff800000);
  for (long long j = threadIdx.x; j < n_support; j += blockDim.x) mx = fmaxf(mx, sc[j]);
  mx = block_max(mx, red);  // (its barriers also publish the zeroed bins)
  float sum = 0.f;
  bool bad = false;
  // whole warps walk the row together so that a warp whose 32 supports share one class (the usual case: supports
  // come class-sorted or in long runs) adds ONE value to the bin instead of 32 contended ones
  const long long n_round = (n_support + blockDim.x - 1) / blockDim.x * blockDim.x;
  for (long long j = threadIdx.x; j < n_round; j += blockDim.x) {
    const bool in = j < n_support;
    const float e = in ? expf(sc[j] - mx) : 0.f;
    const long long y = in ? lab[j] : -1;
    sum += e;
    const bool ok = y >= 0 && y < n_classes;
    bad |= in && !ok;
    const long long y0 = __shfl_sync(0xffffffffu, y, 0);
    if (__all_sync(0xffffffffu, y == y0)) {
      const float w = warp_sum(e);
      if ((threadIdx.x & 31) == 0 && ok) atomicAdd(&bins[y], w);
    } else if (ok) {
      atomicAdd(&bins[y], e);
    }
  }
  if (bad) *reinterpret_cast<volatile int32_t*>(status) = 1;
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) row_lse[b] = mx + logf(sum);
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) logp[b * n_classes + c] = logf(bins[c] * inv + 1e-12f);
}

// one block per query: coef[b,j] = dL/dscore[b,j] (linear kinds, times scale) or dL/dscore / dist (euclid kinds)
__global__ void __launch_bounds__(256) coef_kernel(int kind, float scale, const float* __restrict__ scores,
                                                   const float* __restrict__ row_lse,
                                                   const float* __restrict__ logp,
                                                   const float* __restrict__ grad_out,
                                                   const int64_t* __restrict__ labels, int labels_batched,
                                                   long long n_support, int n_classes, float* __restrict__ coef,
                                                   float* __restrict__ grad_scale_rows) {
  extern __shared__ float gP[];  // n_classes floats
  __shared__ float red[8];
  const long long b = blockIdx.x;
  const float* lp = logp + b * n_classes;
  const float* g = grad_out + b * n_classes;
  float dsum = 0.f;
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) {
    const float pe = expf(lp[c]);                 // P + 1e-12
    const float gp = g[c] / pe;                   // d/dP log(P + eps)
    gP[c] = gp;
    dsum += fmaxf(pe - 1e-12f, 0.f) * gp;         // sum_c P_c gP_c = sum_j p_j gP[y_j]
  }
  dsum = block_sum(dsum, red);
  const float* sc = scores + b * n_support;
  const int64_t* lab = labels + (labels_batched ? b * n_support : 0);
  const float z = row_lse[b];
  float gscale = 0.f;
  // gridDim.y slices of the support axis (long rows: one block per query would leave the GPU to B blocks); every
  // slice recomputes the O(C) prologue above.  grad_scale_rows is only produced with a single slice.
  const long long per = (n_support + gridDim.y - 1) / gridDim.y;
  const long long j_lo = (long long)blockIdx.y * per, j_hi = min(j_lo + per, n_support);
  for (long long j = j_lo + threadIdx.x; j < j_hi; j += blockDim.x) {
    const float s = sc[j];
    const float p = expf(s - z);
    // an out-of-range label (rejected by the forward's status flag) contributes to no class: never index gP with it
    const long long y = lab[j];
    const float gs = p * ((y >= 0 && y < n_classes ? gP[y] : 0.f) - dsum);
    float v;
    if (kind_euclid(kind)) {
      const float dist = -s;
      v = dist > 0.f ? gs / dist : 0.f;
    } else {
      v = (kind == NW_KIND_CLIP) ? gs * scale : gs;
      gscale += gs * s;  // dscore/dlogit_scale = score
    }
    coef[b * n_support + j] = v;
  }
  if (grad_scale_rows) {
    gscale = block_sum(gscale, red);
    if (threadIdx.x == 0) grad_scale_rows[b] = gscale;
  }
}

// grad_q: one block per query.  row (d floats) is staged in shared memory for the normalisation Jacobian.
__global__ void __launch_bounds__(256) grad_q_kernel(int kind, const float* __restrict__ q, int d,
                                                     const float* __restrict__ s, long long n_support,
                                                     int batched, const float* __restrict__ coef,
                                                     const float* __restrict__ inv_q,
                                                     const float* __restrict__ inv_s, float* __restrict__ grad_q) {
  extern __shared__ float row[];  // d floats
  __shared__ float red[8];
  const long long b = blockIdx.x;
  const bool norm = kind_normalised(kind);
  const bool euc = kind_euclid(kind);
  const float iq = norm ? inv_q[b] : 1.0f;
  const float* cf = coef + b * n_support;
  const float* sb = s + (batched ? b * n_support * d : 0);
  const float* isb = norm ? inv_s + (batched ? b * n_support : 0) : nullptr;
  float dot = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float qv = q[b * d + c] * iq;
    float acc = 0.f, csum = 0.f;
    for (long long j = 0; j < n_support; ++j) {
      const float w = cf[j];
      const float sv = sb[j * d + c] * (norm ? isb[j] : 1.0f);
      acc = fmaf(w, sv, acc);
      csum += w;
    }
    if (euc) acc -= csum * qv;  // sum_j r_bj (s_j - q_b)
    row[c] = acc;
    dot += acc * qv;
  }
  if (norm) {
    dot = block_sum(dot, red);
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      const float qv = q[b * d + c] * iq;
      grad_q[b * d + c] = (row[c] - dot * qv) * iq;  // (I - q^ q^T) / max(|q|, eps)
    }
  } else {
    for (int c = threadIdx.x; c < d; c += blockDim.x) grad_q[b * d + c] = row[c];
  }
}

// grad_s for a shared support: one block per support row j (sum over queries).
__global__ void __launch_bounds__(256) grad_s_shared_kernel(int kind, const float* __restrict__ q, int n_query,
                                                            int d, const float* __restrict__ s,
                                                            long long n_support, const float* __restrict__ coef,
                                                            const float* __restrict__ inv_q,
                                                            const float* __restrict__ inv_s,
                                                            float* __restrict__ grad_s) {
  extern __shared__ float row[];
  __shared__ float red[8];
  const long long j = blockIdx.x;
  const bool norm = kind_normalised(kind);
  const bool euc = kind_euclid(kind);
  const float is = norm ? inv_s[j] : 1.0f;
  float dot = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float sv = s[j * d + c] * is;
    float acc = 0.f, csum = 0.f;
    for (int b = 0; b < n_query; ++b) {
      const float w = coef[(long long)b * n_support + j];
      const float qv = q[(long long)b * d + c] * (norm ? inv_q[b] : 1.0f);
      acc = fmaf(w, qv, acc);
      csum += w;
    }
    if (euc) acc -= csum * sv;  // sum_b r_bj (q_b - s_j)
    row[c] = acc;
    dot += acc * sv;
  }
  if (norm) {
    dot = block_sum(dot, red);
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      const float sv = s[j * d + c] * is;
      grad_s[j * d + c] = (row[c] - dot * sv) * is;
    }
  } else {
    for (int c = threadIdx.x; c < d; c += blockDim.x) grad_s[j * d + c] = row[c];
  }
}

// ---- large shared supports: grad_q as a split reduction -------------------------------------------------------
// grad_q_kernel above gives every query ONE block that walks all N supports: B blocks on a 148-SM GPU, each
// streaming the whole support (100 ms at N = 1.28M, d = 2048, B = 8).  Here a block owns a chunk of GQ_CHUNK
// supports and 256 columns and accumulates A[b, c] = sum_j w'[b, j] s[j, c] for GQ_BT queries at a time
// (w' = coef, times 1/|s_j| for the normalised kinds), so the support is read from HBM once; the per-chunk partials
// are summed in chunk order by grad_q_finish_kernel (deterministic), which also applies the euclidean "- rowsum(R) q"
// term and the normalisation Jacobian exactly as grad_q_kernel does.
constexpr int GQ_CHUNK = 1024;
constexpr int GQ_BT = 8;

__global__ void __launch_bounds__(256) grad_q_split_kernel(const float* __restrict__ s, long long n_support, int d,
                                                           const float* __restrict__ coef,
                                                           const float* __restrict__ inv_s, int n_query,
                                                           float* __restrict__ partial) {
  __shared__ float w[GQ_BT][GQ_CHUNK];
  const int c = blockIdx.y * 256 + threadIdx.x;
  const long long j0 = (long long)blockIdx.x * GQ_CHUNK;
  const int nj = int(min((long long)GQ_CHUNK, n_support - j0));
  for (int b0 = 0; b0 < n_query; b0 += GQ_BT) {
    const int nb = min(GQ_BT, n_query - b0);
    for (int i = threadIdx.x; i < GQ_BT * GQ_CHUNK; i += 256) {
      const int t = i / GQ_CHUNK, j = i - t * GQ_CHUNK;
      float v = 0.f;
      if (t < nb && j < nj) {
        v = coef[(long long)(b0 + t) * n_support + j0 + j];
        if (inv_s) v *= inv_s[j0 + j];
      }
      w[t][j] = v;
    }
    __syncthreads();
    if (c < d) {
      float acc[GQ_BT];
#pragma unroll
      for (int t = 0; t < GQ_BT; ++t) acc[t] = 0.f;
      const float* sp = s + j0 * d + c;
      // 8 support values in flight per thread before any is consumed (load-latency bound otherwise)
      for (int jb = 0; jb < nj; jb += 8) {
        float sv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) sv[u] = jb + u < nj ? __ldcs(sp + (long long)(jb + u) * d) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
          for (int t = 0; t < GQ_BT; ++t) acc[t] = fmaf(w[t][(jb + u) & (GQ_CHUNK - 1)], sv[u], acc[t]);
        }
      }
      for (int t = 0; t < nb; ++t) partial[((long long)blockIdx.x * n_query + b0 + t) * d + c] = acc[t];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) grad_q_finish_kernel(int kind, const float* __restrict__ q, int d,
                                                            long long n_support, const float* __restrict__ coef,
                                                            const float* __restrict__ inv_q,
                                                            const float* __restrict__ partial, int n_chunks,
                                                            int n_query, float* __restrict__ grad_q) {
  extern __shared__ float row[];  // d floats
  __shared__ float red[8];
  const long long b = blockIdx.x;
  const bool norm = kind_normalised(kind);
  const bool euc = kind_euclid(kind);
  const float iq = norm ? inv_q[b] : 1.0f;
  float csum = 0.f;
  if (euc) {
    const float* cf = coef + b * n_support;
    for (long long j = threadIdx.x; j < n_support; j += 256) csum += cf[j];
    csum = block_sum(csum, red);
  }
  float dot = 0.f;
  for (int c = threadIdx.x; c < d; c += 256) {
    float acc = 0.f;
    for (int g = 0; g < n_chunks; ++g) acc += partial[((long long)g * n_query + b) * d + c];
    const float qv = q[b * d + c] * iq;
    if (euc) acc -= csum * qv;  // sum_j r_bj (s_j - q_b)
    row[c] = acc;
    dot += acc * qv;
  }
  if (norm) {
    dot = block_sum(dot, red);
    for (int c = threadIdx.x; c < d; c += 256) grad_q[b * d + c] = (row[c] - dot * (q[b * d + c] * iq)) * iq;
  } else {
    for (int c = threadIdx.x; c < d; c += 256) grad_q[b * d + c] = row[c];
  }
}

// grad_s for a shared support, several rows per block: grad_s_shared_kernel launches one 256-thread block per
// support row for 64 FMAs per thread; with millions of rows the block turnover, not HBM, was the bound.
constexpr int GS_ROWS = 16;

__global__ void __launch_bounds__(256) grad_s_rows_kernel(int kind, const float* __restrict__ q, int n_query, int d,
                                                          const float* __restrict__ s, long long n_support,
                                                          const float* __restrict__ coef,
                                                          const float* __restrict__ inv_q,
                                                          const float* __restrict__ inv_s,
                                                          float* __restrict__ grad_s) {
  extern __shared__ float row[];  // d floats
  __shared__ float red[8];
  const bool norm = kind_normalised(kind);
  const bool euc = kind_euclid(kind);
  const long long j_end = min((long long)(blockIdx.x + 1) * GS_ROWS, n_support);
  for (long long j = (long long)blockIdx.x * GS_ROWS; j < j_end; ++j) {
    const float is = norm ? inv_s[j] : 1.0f;
    float dot = 0.f;
    for (int c = threadIdx.x; c < d; c += 256) {
      const float sv = s[j * d + c] * is;
      float acc = 0.f, csum = 0.f;
      for (int b = 0; b < n_query; ++b) {
        const float w = coef[(long long)b * n_support + j];
        const float qv = q[(long long)b * d + c] * (norm ? inv_q[b] : 1.0f);
        acc = fmaf(w, qv, acc);
        csum += w;
      }
      if (euc) acc -= csum * sv;  // sum_b r_bj (q_b - s_j)
      row[c] = acc;
      dot += acc * sv;
    }
    if (norm) {
      dot = block_sum(dot, red);
      for (int c = threadIdx.x; c < d; c += 256) grad_s[j * d + c] = (row[c] - dot * (s[j * d + c] * is)) * is;
      __syncthreads();  // row[] and red[] are reused by the next support row
    } else {
      for (int c = threadIdx.x; c < d; c += 256) grad_s[j * d + c] = row[c];
    }
  }
}

// grad_s for a shared support and the kinds without normalisation (euclidean, dot product): a block owns GS_TILE
// support rows x 256 columns.  The coefficients of those rows (all queries) and the query tile are staged in shared
// memory once; every thread then issues the loads of its column for ALL rows of the tile before consuming any
// (the row-at-a-time kernel above is load-latency bound: 1.5 TB/s at N = 160k, d = 2048).
constexpr int GS_TILE = 16;
constexpr int GS_BQ = 32;  // queries staged per pass

__global__ void __launch_bounds__(256) grad_s_tile_kernel(int euclid, const float* __restrict__ q, int n_query, int d,
                                                          const float* __restrict__ s, long long n_support,
                                                          const float* __restrict__ coef,
                                                          float* __restrict__ grad_s) {
  __shared__ float w[GS_BQ][GS_TILE];
  __shared__ float qt[GS_BQ][256];
  const long long j0 = (long long)blockIdx.x * GS_TILE;
  const int c = blockIdx.y * 256 + threadIdx.x;
  const int nj = int(min((long long)GS_TILE, n_support - j0));
  float sv[GS_TILE], acc[GS_TILE], csum[GS_TILE];
#pragma unroll
  for (int j = 0; j < GS_TILE; ++j) {
    sv[j] = (c < d && j < nj) ? __ldcs(s + (j0 + j) * d + c) : 0.f;
    acc[j] = 0.f;
    csum[j] = 0.f;
  }
  for (int b0 = 0; b0 < n_query; b0 += GS_BQ) {
    const int nb = min(GS_BQ, n_query - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < GS_BQ * GS_TILE; i += 256) {
      const int b = i / GS_TILE, j = i - b * GS_TILE;
      w[b][j] = (b < nb && j < nj) ? coef[(long long)(b0 + b) * n_support + j0 + j] : 0.f;
    }
    for (int b = 0; b < nb; ++b) qt[b][threadIdx.x] = c < d ? q[(long long)(b0 + b) * d + c] : 0.f;
    __syncthreads();
    for (int b = 0; b < nb; ++b) {  // queries in ascending order: the summation order of grad_s_shared_kernel
      const float qv = qt[b][threadIdx.x];
#pragma unroll
      for (int j = 0; j < GS_TILE; ++j) {
        acc[j] = fmaf(w[b][j], qv, acc[j]);
        csum[j] += w[b][j];
      }
    }
  }
  if (c < d) {
#pragma unroll
    for (int j = 0; j < GS_TILE; ++j)
      if (j < nj) grad_s[(j0 + j) * d + c] = euclid ? acc[j] - csum[j] * sv[j] : acc[j];  // sum_b r_bj (q_b - s_j)
  }
}

// grad_s for a per-query support (B,N,d): one block per (b, j) pair.
__global__ void __launch_bounds__(128) grad_s_batched_kernel(int kind, const float* __restrict__ q, int d,
                                                             const float* __restrict__ s, long long n_support,
                                                             const float* __restrict__ coef,
                                                             const float* __restrict__ inv_q,
                                                             const float* __restrict__ inv_s,
                                                             float* __restrict__ grad_s) {
  __shared__ float red[8];
  const long long pair = blockIdx.x;
  const long long b = pair / n_support;
  const bool norm = kind_normalised(kind);
  const bool euc = kind_euclid(kind);
  const float is = norm ? inv_s[pair] : 1.0f;
  const float iq = norm ? inv_q[b] : 1.0f;
  const float w = coef[pair];
  float dot = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float sv = s[pair * d + c] * is;
    const float qv = q[b * d + c] * iq;
    const float g = euc ? w * (qv - sv) : w * qv;
    dot += g * sv;
  }
  if (norm) dot = block_sum(dot, red);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float sv = s[pair * d + c] * is;
    const float qv = q[b * d + c] * iq;
    const float g = euc ? w * (qv - sv) : w * qv;
    grad_s[pair * d + c] = norm ? (g - dot * sv) * is : g;
  }
}


// ---------------------------------------------------------------------------------------------
// Fused small-support kernels (episodic training: N <= SMALL_N supports per query).  One launch for the
// whole forward, two for the backward; inverse norms are recomputed in registers instead of staged.
// ---------------------------------------------------------------------------------------------
constexpr int SMALL_N = 1024;

__device__ __forceinline__ float pair_score(int kind, float scale, const float* __restrict__ qp,
                                            const float* __restrict__ sp, int d, int lane) {
  if (kind == NW_KIND_EUCLIDEAN) {
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float df = qp[c] - sp[c];
      acc = fmaf(df, df, acc);
    }
    return -sqrtf(warp_sum(acc));
  }
  if (kind == NW_KIND_DOT) {
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc = fmaf(qp[c], sp[c], acc);
    return warp_sum(acc);
  }
  float qq = 0.f, ss = 0.f;
  for (int c = lane; c < d; c += 32) {
    qq = fmaf(qp[c], qp[c], qq);
    ss = fmaf(sp[c], sp[c], ss);
  }
  const float iq = 1.0f / fmaxf(sqrtf(warp_sum(qq)), 1e-12f);
  const float is = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  float acc = 0.f;
  if (kind == NW_KIND_HYPERSPHERE) {
    for (int c = lane; c < d; c += 32) {
      const float df = __fsub_rn(__fmul_rn(qp[c], iq), __fmul_rn(sp[c], is));
      acc = fmaf(df, df, acc);
    }
    return -sqrtf(warp_sum(acc));
  }
  for (int c = lane; c < d; c += 32) acc = fmaf(__fmul_rn(qp[c], iq), __fmul_rn(sp[c], is), acc);
  const float out = warp_sum(acc);
  return kind == NW_KIND_CLIP ? out * scale : out;
}

// one block per query: scores row -> softmax statistics -> per-class sums -> log-probs
__global__ void __launch_bounds__(256) small_forward_kernel(int kind, float scale, const float* __restrict__ q,
                                                            int d, const float* __restrict__ s, int n_support,
                                                            int batched, const int64_t* __restrict__ labels,
                                                            int labels_batched, int n_classes,
                                                            float* __restrict__ scores, float* __restrict__ logp,
                                                            float* __restrict__ row_lse,
                                                            int32_t* __restrict__ status) {
  __shared__ float sc[SMALL_N];
  __shared__ int lab[SMALL_N];
  __shared__ float red[8];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* qp = q + (long long)b * d;
  const float* sb = s + (batched ? (long long)b * n_support * d : 0);
  const int64_t* lb = labels + (labels_batched ? (long long)b * n_support : 0);
  for (int j = warp; j < n_support; j += 8) {
    const float v = pair_score(kind, scale, qp, sb + (long long)j * d, d, lane);
    if (lane == 0) {
      sc[j] = v;
      scores[(long long)b * n_support + j] = v;
    }
  }
  bool bad = false;
  for (int j = threadIdx.x; j < n_support; j += 256) {
    const long long y = lb[j];
    const bool oob = y < 0 || y >= n_classes;
    bad |= oob;
    lab[j] = oob ? -1 : int(y);  // (a 64-bit label must not alias a class after narrowing)
  }
  if (bad) *reinterpret_cast<volatile int32_t*>(status) = 1;
  __syncthreads();
  float mx = __int_as_float(0xff800000);
  for (int j = threadIdx.x; j < n_support; j += 256) mx = fmaxf(mx, sc[j]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int j = threadIdx.x; j < n_support; j += 256) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = block_sum(sum, red);  // (contains the barriers that publish sc[])
  if (threadIdx.x == 0) row_lse[b] = mx + logf(sum);
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < n_classes; c += 256) {
    float acc = 0.f;
    for (int j = 0; j < n_support; ++j)
      if (lab[j] == c) acc += sc[j];
    logp[(long long)b * n_classes + c] = logf(acc * inv + 1e-12f);
  }
}

// one block per query: dL/dscore coefficients (kept in shared memory and written for the grad_s pass) + grad_q
__global__ void __launch_bounds__(256) small_coef_gradq_kernel(
    int kind, float scale, const float* __restrict__ q, int d, const float* __restrict__ s, int n_support,
    int batched, const int64_t* __restrict__ labels, int labels_batched, int n_classes,
    const float* __restrict__ scores, const float* __restrict__ row_lse, const float* __restrict__ logp,
    const float* __restrict__ grad_out, float* __restrict__ coef, float* __restrict__ grad_q,
    float* __restrict__ grad_scale_rows) {
  extern __shared__ float dyn[];  // [n_classes] gP, then [d] row
  __shared__ float cf[SMALL_N];
  __shared__ float isn[SMALL_N];
  __shared__ float red[8];
  float* gP = dyn;
  float* row = dyn + n_classes;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool norm = kind_normalised(kind);
  const bool euc = kind_euclid(kind);
  const float* lp = logp + (long long)b * n_classes;
  const float* g = grad_out + (long long)b * n_classes;
  const float* qp = q + (long long)b * d;
  const float* sb = s + (batched ? (long long)b * n_support * d : 0);
  const int64_t* lb = labels + (labels_batched ? (long long)b * n_support : 0);
  float dsum = 0.f;
  for (int c = threadIdx.x; c < n_classes; c += 256) {
    const float pe = expf(lp[c]);
    const float gp = g[c] / pe;
    gP[c] = gp;
    dsum += fmaxf(pe - 1e-12f, 0.f) * gp;
  }
  dsum = block_sum(dsum, red);
  const float z = row_lse[b];
  float gscale = 0.f;
  for (int j = threadIdx.x; j < n_support; j += 256) {
    const float sv = scores[(long long)b * n_support + j];
    const long long y = lb[j];  // out-of-range labels belong to no class (see coef_kernel)
    const float gs = expf(sv - z) * ((y >= 0 && y < n_classes ? gP[y] : 0.f) - dsum);
    float v;
    if (euc) {
      const float dist = -sv;
      v = dist > 0.f ? gs / dist : 0.f;
    } else {
      v = (kind == NW_KIND_CLIP) ? gs * scale : gs;
      gscale += gs * sv;
    }
    cf[j] = v;
    coef[(long long)b * n_support + j] = v;
  }
  if (grad_scale_rows) {
    gscale = block_sum(gscale, red);
    if (threadIdx.x == 0) grad_scale_rows[b] = gscale;
  }
  float iq = 1.0f;
  if (norm) {
    for (int j = warp; j < n_support; j += 8) {
      const float* sp = sb + (long long)j * d;
      float ss = 0.f;
      for (int c = lane; c < d; c += 32) ss = fmaf(sp[c], sp[c], ss);
      ss = warp_sum(ss);
      if (lane == 0) isn[j] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }
    float qq = 0.f;
    for (int c = threadIdx.x; c < d; c += 256) qq = fmaf(qp[c], qp[c], qq);
    qq = block_sum(qq, red);
    iq = 1.0f / fmaxf(sqrtf(qq), 1e-12f);
  }
  __syncthreads();
  if (grad_q == nullptr) return;
  float dot = 0.f;
  for (int c = threadIdx.x; c < d; c += 256) {
    const float qv = qp[c] * iq;
    float acc = 0.f, csum = 0.f;
    for (int j = 0; j < n_support; ++j) {
      const float w = cf[j];
      const float sv = sb[(long long)j * d + c] * (norm ? isn[j] : 1.0f);
      acc = fmaf(w, sv, acc);
      csum += w;
    }
    if (euc) acc -= csum * qv;
    row[c] = acc;
    dot += acc * qv;
  }
  if (norm) {
    dot = block_sum(dot, red);
    for (int c = threadIdx.x; c < d; c += 256) grad_q[(long long)b * d + c] = (row[c] - dot * (qp[c] * iq)) * iq;
  } else {
    for (int c = threadIdx.x; c < d; c += 256) grad_q[(long long)b * d + c] = row[c];
  }
}

// grad_s for a shared small support: one block per support row, inverse norms recomputed in the block
__global__ void __launch_bounds__(256) small_grad_s_kernel(int kind, const float* __restrict__ q, int n_query, int d,
                                                           const float* __restrict__ s, int n_support, int batched,
                                                           const float* __restrict__ coef,
                                                           float* __restrict__ grad_s) {
  extern __shared__ float row[];  // d floats
  __shared__ float iqs[256];
  __shared__ float red[8];
  const bool norm = kind_normalised(kind);
  const bool euc = kind_euclid(kind);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pair = blockIdx.x;  // batched: (b, j) pair; shared: j
  const int b0 = batched ? int(pair / n_support) : 0;
  const int nb = batched ? 1 : n_query;
  const int j = batched ? int(pair % n_support) : int(pair);
  const float* sp = s + pair * d;
  float is = 1.0f;
  if (norm) {
    float ss = 0.f;
    for (int c = threadIdx.x; c < d; c += 256) ss = fmaf(sp[c], sp[c], ss);
    ss = block_sum(ss, red);
    is = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    for (int bb = warp; bb < nb; bb += 8) {
      const float* qp = q + (long long)(b0 + bb) * d;
      float qq = 0.f;
      for (int c = lane; c < d; c += 32) qq = fmaf(qp[c], qp[c], qq);
      qq = warp_sum(qq);
      if (lane == 0) iqs[bb] = 1.0f / fmaxf(sqrtf(qq), 1e-12f);
    }
    __syncthreads();
  }
  float dot = 0.f;
  for (int c = threadIdx.x; c < d; c += 256) {
    const float sv = sp[c] * is;
    float acc = 0.f, csum = 0.f;
    for (int bb = 0; bb < nb; ++bb) {
      const float w = coef[(long long)(b0 + bb) * n_support + j];
      const float qv = q[(long long)(b0 + bb) * d + c] * (norm ? iqs[bb] : 1.0f);
      acc = fmaf(w, qv, acc);
      csum += w;
    }
    if (euc) acc -= csum * sv;
    row[c] = acc;
    dot += acc * sv;
  }
  if (norm) {
    dot = block_sum(dot, red);
    for (int c = threadIdx.x; c < d; c += 256) grad_s[pair * d + c] = (row[c] - dot * (sp[c] * is)) * is;
  } else {
    for (int c = threadIdx.x; c < d; c += 256) grad_s[pair * d + c] = row[c];
  }
}

}  // namespace direct
}  // namespace nw

using namespace nw;

static int check_kind(int kind) {
  NW_REQUIRE(kind >= NW_KIND_EUCLIDEAN && kind <= NW_KIND_CLIP, NW_ERR_INVALID, "unknown kernel kind %d", kind);
  return NW_OK;
}

extern "C" int nw_direct_scores(int kind, float scale, const float* q, int n_query, int d, const float* s,
                                int64_t n_support, int support_batched, float* scores, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_kind(kind);
  if (rc != NW_OK) return rc;
  NW_REQUIRE(q && s && scores, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_query > 0 && d > 0 && n_support > 0, NW_ERR_INVALID, "shapes must be positive");
  const bool aligned = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
  if (!support_batched && (kind == NW_KIND_EUCLIDEAN || kind == NW_KIND_DOT) && n_query >= 16 && n_support >= 64 &&
      aligned) {  // 32 x 32 pairs per block through shared-memory slabs
    const long long blocks = ceil_div_ll(n_query, direct::SB_ROWS) * ceil_div_ll(n_support, direct::SB_ROWS);
    NW_REQUIRE(blocks < (1ll << 31), NW_ERR_UNSUPPORTED, "too many (query, support) pairs for the direct path");
    auto kernel = kind == NW_KIND_EUCLIDEAN ? direct::scores_block_kernel<true> : direct::scores_block_kernel<false>;
    NW_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(direct::SB_SMEM_BYTES)));
    kernel<<<unsigned(blocks), 256, direct::SB_SMEM_BYTES, stream>>>(q, n_query, d, s, n_support, scores);
    NW_CUDA_OK(cudaGetLastError());
    return NW_OK;
  }
  const bool wide = !support_batched && n_query >= 3;  // 4 x 4 pairs per warp; else 1 query x 4 supports
  const long long tiles = wide ? ceil_div_ll(n_query, 4) * ceil_div_ll(n_support, 4)
                               : (long long)n_query * ceil_div_ll(n_support, 4);
  const long long blocks = ceil_div_ll(tiles, 8);
  NW_REQUIRE(blocks < (1ll << 31), NW_ERR_UNSUPPORTED, "too many (query, support) pairs for the direct path");
  if (wide)
    direct::scores_kernel<4, 4><<<unsigned(blocks), 256, 0, stream>>>(kind, scale, q, n_query, d, s, n_support, 0,
                                                                      scores);
  else
    direct::scores_kernel<1, 4><<<unsigned(blocks), 256, 0, stream>>>(kind, scale, q, n_query, d, s, n_support,
                                                                      support_batched, scores);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_direct_aggregate(const float* scores, const int64_t* labels, int labels_batched, int n_query,
                                   int64_t n_support, int n_classes, float* logp, float* row_lse,
                                   int32_t* status_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(scores && labels && logp && row_lse && status_out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_query > 0 && n_support > 0 && n_classes > 0, NW_ERR_INVALID, "shapes must be positive");
  if (n_support > direct::AGG_BINS_MIN_N && n_classes <= direct::AGG_BINS_MAX_C)
    direct::aggregate_bins_kernel<<<n_query, 1024, n_classes * sizeof(float), stream>>>(
        scores, labels, labels_batched, n_support, n_classes, logp, row_lse, status_out);
  else
    direct::aggregate_kernel<<<n_query, 256, 0, stream>>>(scores, labels, labels_batched, n_support, n_classes, logp,
                                                          row_lse, status_out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_direct_forward(int kind, float scale, const float* q, int n_query, int d, const float* s,
                                 int64_t n_support, int support_batched, const int64_t* labels, int labels_batched,
                                 int n_classes, float* scores, float* logp, float* row_lse, int32_t* status_flag,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_kind(kind);
  if (rc != NW_OK) return rc;
  NW_REQUIRE(q && s && labels && scores && logp && row_lse && status_flag, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_query > 0 && d > 0 && n_support > 0 && n_classes > 0, NW_ERR_INVALID, "shapes must be positive");
  if (n_support <= direct::SMALL_N) {
    direct::small_forward_kernel<<<n_query, 256, 0, stream>>>(kind, scale, q, d, s, int(n_support), support_batched,
                                                              labels, labels_batched, n_classes, scores, logp,
                                                              row_lse, status_flag);
    NW_CUDA_OK(cudaGetLastError());
    return NW_OK;
  }
  rc = nw_direct_scores(kind, scale, q, n_query, d, s, n_support, support_batched, scores, stream_);
  if (rc != NW_OK) return rc;
  return nw_direct_aggregate(scores, labels, labels_batched, n_query, n_support, n_classes, logp, row_lse, status_flag,
                             stream_);
}

// workspace: coefficients (B*N) | 1/|s| (N or B*N) | 1/|q| (B) | split grad_q partials [chunk][query][d] (large
// shared supports only)
extern "C" int64_t nw_direct_backward_workspace_elems(int n_query, int d, int64_t n_support, int support_batched) {
  const int64_t pairs = int64_t(n_query) * n_support;
  int64_t elems = pairs + (support_batched ? pairs : n_support) + n_query;
  if (!support_batched && n_support > direct::SMALL_N)
    elems += ceil_div_ll(n_support, direct::GQ_CHUNK) * int64_t(n_query) * d;
  return elems;
}

extern "C" int nw_direct_backward(int kind, float scale, const float* q, int n_query, int d, const float* s,
                                  int64_t n_support, int support_batched, const int64_t* labels,
                                  int labels_batched, int n_classes, const float* scores, const float* row_lse,
                                  const float* logp, const float* grad_out, float* workspace, float* grad_q,
                                  float* grad_s, float* grad_scale_rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_kind(kind);
  if (rc != NW_OK) return rc;
  NW_REQUIRE(q && s && labels && scores && row_lse && logp && grad_out && workspace, NW_ERR_INVALID,
             "NULL pointer argument");
  NW_REQUIRE(grad_q || grad_s, NW_ERR_INVALID, "at least one of grad_q / grad_s must be requested");
  NW_REQUIRE(n_query > 0 && d > 0 && n_support > 0 && n_classes > 0, NW_ERR_INVALID, "shapes must be positive");
  NW_REQUIRE(d + n_classes <= NW_DIRECT_BACKWARD_MAX_D_PLUS_C, NW_ERR_UNSUPPORTED,
             "direct backward stages one feature row and one class table in shared memory (d + C <= %d)",
             NW_DIRECT_BACKWARD_MAX_D_PLUS_C);
  // more than the default 48 KB of dynamic shared memory is an opt-in per kernel (and per device: cheap, idempotent)
  const size_t big = size_t(d + n_classes) * sizeof(float);
  if (big > 36 * 1024) {  // (the kernels also hold up to 8.2 KB of static shared memory; the 48 KB default covers both)
    const int cap = NW_DIRECT_BACKWARD_MAX_D_PLUS_C * int(sizeof(float));
    NW_CUDA_OK(cudaFuncSetAttribute(direct::small_coef_gradq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    NW_CUDA_OK(cudaFuncSetAttribute(direct::small_grad_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    NW_CUDA_OK(cudaFuncSetAttribute(direct::coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    NW_CUDA_OK(cudaFuncSetAttribute(direct::grad_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    NW_CUDA_OK(cudaFuncSetAttribute(direct::grad_s_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
  }
  const long long pairs = (long long)n_query * n_support;
  NW_REQUIRE(pairs < (1ll << 31), NW_ERR_UNSUPPORTED, "too many (query, support) pairs for the direct path");
  float* coef = workspace;
  if (n_support <= direct::SMALL_N && (support_batched || n_query <= 256)) {
    // fused small path: coefficients + grad_q in one launch, grad_s in a second
    direct::small_coef_gradq_kernel<<<n_query, 256, (n_classes + d) * sizeof(float), stream>>>(
        kind, scale, q, d, s, int(n_support), support_batched, labels, labels_batched, n_classes, scores, row_lse,
        logp, grad_out, coef, grad_q, kind == NW_KIND_CLIP ? grad_scale_rows : nullptr);
    NW_CUDA_OK(cudaGetLastError());
    if (grad_s) {
      const long long rows = support_batched ? pairs : n_support;
      direct::small_grad_s_kernel<<<unsigned(rows), 256, d * sizeof(float), stream>>>(
          kind, q, n_query, d, s, int(n_support), support_batched, coef, grad_s);
      NW_CUDA_OK(cudaGetLastError());
    }
    return NW_OK;
  }
  float* inv_s = workspace + pairs;
  float* inv_q = inv_s + (support_batched ? pairs : n_support);
  float* split_workspace = (!support_batched && n_support > direct::SMALL_N) ? inv_q + n_query : nullptr;
  const bool norm = direct::kind_normalised(kind);
  if (norm) {
    const long long srows = support_batched ? pairs : n_support;
    direct::inv_norm_kernel<<<unsigned(ceil_div_ll(srows, 8)), 256, 0, stream>>>(s, srows, d, inv_s);
    direct::inv_norm_kernel<<<unsigned(ceil_div(n_query, 8)), 256, 0, stream>>>(q, n_query, d, inv_q);
    NW_CUDA_OK(cudaGetLastError());
  }
  {
    const bool want_scale = kind == NW_KIND_CLIP && grad_scale_rows != nullptr;
    const int slices = want_scale ? 1 : int(ceil_div_ll(n_support, 32768) < 64 ? ceil_div_ll(n_support, 32768) : 64);
    direct::coef_kernel<<<dim3(n_query, slices), 256, n_classes * sizeof(float), stream>>>(
        kind, scale, scores, row_lse, logp, grad_out, labels, labels_batched, n_support, n_classes, coef,
        want_scale ? grad_scale_rows : nullptr);
  }
  NW_CUDA_OK(cudaGetLastError());
  if (grad_q) {
    if (!support_batched && split_workspace != nullptr) {
      // large shared support: split reduction over support chunks (the support is read from HBM once)
      const int n_chunks = int(ceil_div_ll(n_support, direct::GQ_CHUNK));
      dim3 grid(n_chunks, ceil_div(d, 256));
      direct::grad_q_split_kernel<<<grid, 256, 0, stream>>>(s, n_support, d, coef, norm ? inv_s : nullptr, n_query,
                                                            split_workspace);
      NW_CUDA_OK(cudaGetLastError());
      direct::grad_q_finish_kernel<<<n_query, 256, d * sizeof(float), stream>>>(
          kind, q, d, n_support, coef, inv_q, split_workspace, n_chunks, n_query, grad_q);
    } else {
      direct::grad_q_kernel<<<n_query, 256, d * sizeof(float), stream>>>(kind, q, d, s, n_support, support_batched,
                                                                         coef, inv_q, inv_s, grad_q);
    }
    NW_CUDA_OK(cudaGetLastError());
  }
  if (grad_s) {
    if (support_batched)
      direct::grad_s_batched_kernel<<<unsigned(pairs), 128, 0, stream>>>(kind, q, d, s, n_support, coef, inv_q,
                                                                         inv_s, grad_s);
    else if (!norm) {
      dim3 grid(unsigned(ceil_div_ll(n_support, direct::GS_TILE)), ceil_div(d, 256));
      direct::grad_s_tile_kernel<<<grid, 256, 0, stream>>>(direct::kind_euclid(kind) ? 1 : 0, q, n_query, d, s,
                                                           n_support, coef, grad_s);
    } else
      direct::grad_s_rows_kernel<<<unsigned(ceil_div_ll(n_support, direct::GS_ROWS)), 256, d * sizeof(float), stream>>>(
          kind, q, n_query, d, s, n_support, coef, inv_q, inv_s, grad_s);
    NW_CUDA_OK(cudaGetLastError());
  }
  return NW_OK;
}
