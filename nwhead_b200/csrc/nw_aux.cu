// K3 class centroids, K4 support influence, K5 neighbour ranking — HBM-bound streaming kernels.

#include "nw_common.cuh"

namespace nw {
namespace aux {

// ---------------------------------------------------------------------------------------------
// K3: per-class mean over a class-sorted fp32 matrix (reference compute_clusters with n_clusters=1,
// nwhead/utils.py:218-246).  Algorithmic traffic: N*d*4 bytes read once + C*d*4 written.
// Grid (class, column block, row split): every block streams whole 4 KB row segments (coalesced,
// 16-B vector loads, 4 independent accumulators per thread); partials are combined in fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int CENT_SPLITS = 8;
constexpr int CENT_THREADS = 256;

template <bool VEC>
__global__ void __launch_bounds__(CENT_THREADS) centroid_partial_kernel(const float* __restrict__ rows, int d,
                                                                        long long ld,
                                                                        const int64_t* __restrict__ perm,
                                                                        const int32_t* __restrict__ offsets,
                                                                        float* __restrict__ partial, int n_classes) {
  const int c = blockIdx.x;
  const int split = blockIdx.z;
  const int lo = offsets[c], hi = offsets[c + 1];
  const int per = (hi - lo + CENT_SPLITS - 1) / CENT_SPLITS;
  const int r0 = lo + split * per;
  const int r1 = min(r0 + per, hi);
  float* dst = partial + (size_t(split) * n_classes + c) * d;
  if (VEC) {
    const int col = (blockIdx.y * CENT_THREADS + threadIdx.x) * 4;
    if (col >= d) return;
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
    int r = r0;
    for (; r + 3 < r1; r += 4) {
      const float4 v0 = __ldcs(reinterpret_cast<const float4*>(rows + (perm ? perm[r] : r) * ld + col));
      const float4 v1 = __ldcs(reinterpret_cast<const float4*>(rows + (perm ? perm[r + 1] : r + 1) * ld + col));
      const float4 v2 = __ldcs(reinterpret_cast<const float4*>(rows + (perm ? perm[r + 2] : r + 2) * ld + col));
      const float4 v3 = __ldcs(reinterpret_cast<const float4*>(rows + (perm ? perm[r + 3] : r + 3) * ld + col));
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
      a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; r < r1; ++r) {
      const float4 v0 = __ldcs(reinterpret_cast<const float4*>(rows + (perm ? perm[r] : r) * ld + col));
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
    float4 o;
    o.x = (a0.x + a1.x) + (a2.x + a3.x);
    o.y = (a0.y + a1.y) + (a2.y + a3.y);
    o.z = (a0.z + a1.z) + (a2.z + a3.z);
    o.w = (a0.w + a1.w) + (a2.w + a3.w);
    *reinterpret_cast<float4*>(dst + col) = o;
  } else {
    const int col = blockIdx.y * CENT_THREADS + threadIdx.x;
    if (col >= d) return;
    float a = 0.f;
    for (int r = r0; r < r1; ++r) a += rows[(perm ? perm[r] : r) * ld + col];
    dst[col] = a;
  }
}

__global__ void centroid_finish_kernel(const float* __restrict__ partial, const int32_t* __restrict__ offsets,
                                       int n_classes, int d, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_classes * d) return;
  const int c = int(i / d);
  const int cnt = offsets[c + 1] - offsets[c];
  float acc = 0.f;
#pragma unroll
  for (int s = 0; s < CENT_SPLITS; ++s) acc += partial[size_t(s) * n_classes * d + i];
  out[i] = cnt > 0 ? acc / float(cnt) : 0.f;
}

// ---------------------------------------------------------------------------------------------
// K4: support influence (reference util/metric.py:23-50).  8 B per (query, support) pair.
// ---------------------------------------------------------------------------------------------
__global__ void onehot_argmax_kernel(const float* __restrict__ onehot, long long n_rows, int n_classes,
                                     int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const float* p = onehot + r * n_classes;
  float best = __int_as_float(0xff800000);
  int bi = 0x7fffffff;
  for (int c = lane; c < n_classes; c += 32) {
    const float v = p[c];
    if (v > best || (v == best && c < bi)) {
      best = v;
      bi = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
  if (lane == 0) out[r] = bi;
}

constexpr int INFL_VEC_PER_THREAD = 2;  // 16-byte vectors per thread: independent loads in flight

template <bool VEC>
__global__ void __launch_bounds__(256) influence_kernel(const float* __restrict__ softmaxes,
                                                        const int32_t* __restrict__ qlabel,
                                                        const float* __restrict__ sweights,
                                                        const int32_t* __restrict__ slabel, long long n_support,
                                                        int n_classes, int n_sets, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int g = blockIdx.z;
  const float* w = sweights + (long long)b * n_support;
  const int32_t* sl = slabel + (long long)g * n_support;
  float* o = out + ((long long)b * n_sets + g) * n_support;
  if (VEC) {
    const long long base = ((long long)blockIdx.x * blockDim.x * INFL_VEC_PER_THREAD + threadIdx.x) * 4;
    float4 wv[INFL_VEC_PER_THREAD];
    int4 lv[INFL_VEC_PER_THREAD];
#pragma unroll
    for (int k = 0; k < INFL_VEC_PER_THREAD; ++k) {  // all streaming loads first
      const long long j = base + (long long)k * blockDim.x * 4;
      if (j < n_support) {
        wv[k] = __ldcs(reinterpret_cast<const float4*>(w + j));
        lv[k] = __ldg(reinterpret_cast<const int4*>(sl + j));
      }
    }
    const int qy = qlabel[b];
    const float p = softmaxes[(long long)b * n_classes + qy];
#pragma unroll
    for (int k = 0; k < INFL_VEC_PER_THREAD; ++k) {
      const long long j = base + (long long)k * blockDim.x * 4;
      if (j < n_support) {
        float4 r;
        r.x = influence_one(p, wv[k].x, lv[k].x == qy);
        r.y = influence_one(p, wv[k].y, lv[k].y == qy);
        r.z = influence_one(p, wv[k].z, lv[k].z == qy);
        r.w = influence_one(p, wv[k].w, lv[k].w == qy);
        __stcs(reinterpret_cast<float4*>(o + j), r);
      }
    }
  } else {
    const int qy = qlabel[b];
    const float p = softmaxes[(long long)b * n_classes + qy];
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_support) return;
    o[j] = influence_one(p, w[j], sl[j] == qy);
  }
}

// ---------------------------------------------------------------------------------------------
// K5: per-row descending ranking by bitonic sort of 64-bit (score, index) keys.
// ---------------------------------------------------------------------------------------------
constexpr int SORT_CHUNK = 4096;  // keys sorted per block in shared memory (32 KB)

__device__ __forceinline__ uint64_t make_key(float v, uint32_t idx) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
  return (uint64_t(~u) << 32) | idx;               // descending score, ascending index on ties
}

__global__ void rank_init_kernel(const float* __restrict__ scores, long long n_cols, long long padded,
                                 uint64_t* __restrict__ keys) {
  const long long r = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= padded) return;
  keys[r * padded + i] = i < n_cols ? make_key(scores[r * n_cols + i], uint32_t(i)) : ~uint64_t(0);
}

__device__ __forceinline__ void cmp_swap(uint64_t& a, uint64_t& b, bool up) {
  if ((a > b) == up) {
    const uint64_t t = a;
    a = b;
    b = t;
  }
}

// Runs every (size, stride) step with stride < SORT_CHUNK for size in [size_lo, size_hi] on one chunk.
__global__ void __launch_bounds__(1024) rank_smem_kernel(uint64_t* __restrict__ keys, long long padded,
                                                         long long size_lo, long long size_hi) {
  __shared__ uint64_t sm[SORT_CHUNK];
  const long long row_base = (long long)blockIdx.y * padded;
  const long long base = (long long)blockIdx.x * SORT_CHUNK;
  const int n = int(padded < SORT_CHUNK ? padded : SORT_CHUNK);
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = keys[row_base + base + i];
  __syncthreads();
  for (long long size = size_lo; size <= size_hi; size <<= 1) {
    long long stride = size >> 1;
    if (stride >= SORT_CHUNK) stride = SORT_CHUNK >> 1;
    for (; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
        const int i = int(2 * t - (t & (stride - 1)));
        const bool up = ((base + i) & size) == 0;
        cmp_swap(sm[i], sm[i + stride], up);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) keys[row_base + base + i] = sm[i];
}

__global__ void rank_global_step_kernel(uint64_t* __restrict__ keys, long long padded, long long size,
                                        long long stride) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= padded / 2) return;
  uint64_t* row = keys + (long long)blockIdx.y * padded;
  const long long i = 2 * t - (t & (stride - 1));
  const bool up = (i & size) == 0;
  uint64_t a = row[i], b = row[i + stride];
  if ((a > b) == up) {
    row[i] = b;
    row[i + stride] = a;
  }
}

// Top-k selection (k <= SORT_CHUNK / 4): every block sorts one chunk of SORT_CHUNK keys in shared memory and keeps
// its k best; the survivors (n * k / SORT_CHUNK of them) go through the same step again until one chunk is left.
// The keys are unique (score, index) pairs, so the result equals the first k entries of the full sort, at a
// quarter of its compare-exchange stages and with no pass over global memory between them.
template <bool FROM_SCORES>
__global__ void __launch_bounds__(1024) rank_select_kernel(const float* __restrict__ scores,
                                                           const uint64_t* __restrict__ in, long long n_in,
                                                           long long in_stride, uint64_t* __restrict__ out,
                                                           long long out_stride, int k,
                                                           const int32_t* __restrict__ todo) {
  __shared__ uint64_t sm[SORT_CHUNK];
  const long long r = blockIdx.y;
  if (todo && todo[r] == 0) return;  // this row was already ranked by rank_radix_kernel
  const long long base = (long long)blockIdx.x * SORT_CHUNK;
  for (int i = threadIdx.x; i < SORT_CHUNK; i += blockDim.x) {
    const long long g = base + i;
    uint64_t key = ~uint64_t(0);
    if (g < n_in) key = FROM_SCORES ? make_key(scores[r * n_in + g], uint32_t(g)) : in[r * in_stride + g];
    sm[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= SORT_CHUNK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < SORT_CHUNK / 2; t += blockDim.x) {
        const int i = 2 * t - (t & (stride - 1));
        cmp_swap(sm[i], sm[i + stride], (i & size) == 0);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) out[r * out_stride + (long long)blockIdx.x * k + i] = sm[i];
}

__global__ void rank_emit_kernel(const uint64_t* __restrict__ keys, long long padded, long long k,
                                 int64_t* __restrict__ idx_out, const int32_t* __restrict__ todo) {
  const long long r = blockIdx.y;
  if (todo && todo[r] == 0) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  idx_out[r * k + i] = int64_t(keys[r * padded + i] & 0xffffffffull);
}

// Top-k by RADIX SELECTION, one block per row: three histogram passes over the row (11 + 11 + 10 bits of the
// descending-orderable score key) find the exact key of the k-th best score, a fourth pass collects the keys that
// are <= it together with their column indices, and only those (k plus ties, a few dozen) are sorted.  The output is
// the same (score descending, index ascending) order as the bitonic paths.  The sort-and-keep-k selection sorts 4096
// keys for every chunk of 4096 columns: 0.36 ms for the 65 best of 20 000 blocks x 256 queries in topk_exact, 16 ms
// for the 20 best of 1.28M columns x 256 rows.  Rows with more than RS_CAP keys at or above the threshold (massive
// ties) set todo[row] = 1 and are left to the bitonic path that follows.
constexpr int RS_THREADS = 1024;
constexpr int RS_CAP = 4096;

__device__ __forceinline__ uint32_t desc_key(float v) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~u;  // the high word of make_key: smaller key = larger score
}

__global__ void __launch_bounds__(RS_THREADS) rank_radix_kernel(const float* __restrict__ scores, long long n_cols,
                                                                int k, int64_t* __restrict__ idx_out,
                                                                int32_t* __restrict__ todo) {
  __shared__ unsigned hist[2048];
  __shared__ uint64_t keys[RS_CAP];
  __shared__ unsigned s_bin, s_below;
  __shared__ int s_count;
  const float* row = scores + (long long)blockIdx.x * n_cols;
  const int lane = threadIdx.x & 31;
  unsigned prefix = 0, mask = 0;
  int kleft = k;
  for (int level = 0; level < 3; ++level) {
    const int shift = level == 0 ? 21 : (level == 1 ? 10 : 0);
    const int nb = level == 2 ? 1024 : 2048;
    for (int i = threadIdx.x; i < nb; i += RS_THREADS) hist[i] = 0;
    __syncthreads();
    for (long long j = threadIdx.x; j < n_cols; j += RS_THREADS) {
      const uint32_t key = desc_key(row[j]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & (nb - 1)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // first bin whose cumulative count reaches kleft: lane l owns nb/32 consecutive bins
      const int per = nb / 32;
      unsigned mine = 0;
      for (int i = 0; i < per; ++i) mine += hist[lane * per + i];
      unsigned incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned excl = incl - mine;
      if (excl < unsigned(kleft) && incl >= unsigned(kleft)) {
        unsigned run = excl;
        for (int i = 0; i < per; ++i) {
          const unsigned h = hist[lane * per + i];
          if (run + h >= unsigned(kleft)) {
            s_bin = lane * per + i;
            s_below = run;
            break;
          }
          run += h;
        }
      }
    }
    __syncthreads();
    prefix |= s_bin << shift;
    mask |= unsigned(nb - 1) << shift;
    kleft -= int(s_below);
    __syncthreads();
  }
  // prefix is now the key of the k-th best score; collect everything at or above it
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  for (long long j = threadIdx.x; j < n_cols; j += RS_THREADS) {
    const uint32_t key = desc_key(row[j]);
    if (key <= prefix) {
      const int pos = atomicAdd(&s_count, 1);
      if (pos < RS_CAP) keys[pos] = (uint64_t(key) << 32) | uint32_t(j);
    }
  }
  __syncthreads();
  const int n = s_count;
  if (n > RS_CAP) {
    if (threadIdx.x == 0) todo[blockIdx.x] = 1;
    return;
  }
  int padded = 32;
  while (padded < n) padded <<= 1;
  for (int i = n + threadIdx.x; i < padded; i += RS_THREADS) keys[i] = ~uint64_t(0);
  __syncthreads();
  for (int size = 2; size <= padded; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < padded / 2; t += RS_THREADS) {
        const int i = 2 * t - (t & (stride - 1));
        cmp_swap(keys[i], keys[i + stride], (i & size) == 0);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += RS_THREADS) idx_out[(long long)blockIdx.x * k + i] = int64_t(keys[i] & 0xffffffffull);
  if (threadIdx.x == 0) todo[blockIdx.x] = 0;
}

// ---------------------------------------------------------------------------------------------
// Exact top-k refinement (SupportBank.topk_exact): one block per query.  The rows of the query's m best 64-row
// bank blocks (ranked by the tensor-core pass) are scored EXACTLY — per pair the arithmetic of the dense path
// (direct::scores_kernel: lane l takes columns l, l+32, ... with one fmaf each, then the butterfly sum), so a
// candidate's score has the bits of the dense (B, N) matrix — the (score, source index) keys are sorted in shared
// memory, and the query is CERTIFIED in place: a row outside the candidates cannot beat the exact k-th candidate
// score (bounds in nwhead_b200/bank.py::topk_exact).  Gather, re-score, ranking and certificate are one launch with
// no host round trip; every query sizes its own candidate budget m <= m_cap from the same bounds; uncertified queries
// are only counted (the caller takes the dense path for them).
// ---------------------------------------------------------------------------------------------
constexpr int TOPK_THREADS = 512;   // 16 warps x 32 loads in flight per block, 2-3 blocks per SM
constexpr int TOPK_ROWS_PER_WARP = 4;  // candidate rows a warp scores together: 4 independent load streams

__device__ __forceinline__ float key_score(uint64_t key) {
  const uint32_t u = ~uint32_t(key >> 32);
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_refine_kernel(
    const float* __restrict__ q, int d, const float* __restrict__ src, long long n, const int64_t* __restrict__ perm,
    const int64_t* __restrict__ block_order, const float* __restrict__ block_best_sorted, int order_stride, int m_cap,
    long long n_blocks, int k, int padded_cap, const float* __restrict__ q_sq, const float* __restrict__ resid_q,
    const float* __restrict__ smax_sq_ptr, const float* __restrict__ resid_max_ptr, int precision,
    int32_t* __restrict__ done, int64_t* __restrict__ idx_out, int32_t* __restrict__ n_pending) {
  extern __shared__ __align__(16) uint8_t topk_smem[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(topk_smem);
  float* qs = reinterpret_cast<float*>(keys + padded_cap);
  __shared__ int need_smem;
  const int b = blockIdx.x;
  if (done[b]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < d; c += TOPK_THREADS) qs[c] = q[(long long)b * d + c];
  // error bounds of the reduced-precision pass for this query (see SupportBank.topk_exact)
  const float* best_row = block_best_sorted + (long long)b * order_stride;
  const float up = 1.0f + 0.0078125f;                                      // norms of the rounded rows -> of the rows
  const float qn = sqrtf(q_sq[b]) * up;
  const float smax = sqrtf(*smax_sq_ptr) * up;
  const float e2 = 3.814697265625e-06f * (qn * qn + smax * smax);          // 2^-18: fp32 accumulation of the pass
  const float eta = (resid_q[b] + *resid_max_ptr) * (1.0f + 0.0009765625f);
  const float lolo = precision == NW_PREC_BF16 ? 0.0f : (0.00390625f * (qn + smax)) * (0.00390625f * (qn + smax));
  const float fp32_sum = 1.0f - float(d + 8) * 5.9604644775390625e-08f;    // the exact path's own summation error
  auto upper = [&](float beta) {  // largest exact score of a row whose pass score is beta
    return -fmaxf(sqrtf(fmaxf(beta * beta - e2, 0.0f)) - eta, 0.0f) * fp32_sum;
  };
  // Candidate budget of THIS query: the k best blocks each hold a row scoring >= lower(beta_(k)); blocks whose best
  // row cannot reach that are out.  (The certificate below is what guarantees exactness; this only sizes the gather.)
  if (threadIdx.x == 0) need_smem = 0;
  __syncthreads();
  int m = m_cap;
  if (n_blocks > k && order_stride >= k) {
    const float bk = best_row[k - 1];
    const float thr = -(sqrtf(bk * bk + e2 + lolo) + eta) / fp32_sum;
    int cnt = 0;
    for (int j = threadIdx.x; j < order_stride; j += TOPK_THREADS) cnt += upper(best_row[j]) >= thr ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0 && cnt) atomicAdd(&need_smem, cnt);
    __syncthreads();
    m = need_smem;
  }
  m = max(m, (k + 63) / 64);
  m = min(m, m_cap);
  const int n_cand = m * 64;
  int padded = 64;
  while (padded < n_cand) padded <<= 1;
  for (int i = n_cand + threadIdx.x; i < padded; i += TOPK_THREADS) keys[i] = ~uint64_t(0);
  __syncthreads();
  const int64_t* order = block_order + (long long)b * order_stride;
  for (int r0 = warp * TOPK_ROWS_PER_WARP; r0 < n_cand; r0 += (TOPK_THREADS / 32) * TOPK_ROWS_PER_WARP) {
    const float* sp[TOPK_ROWS_PER_WARP];
    long long srow[TOPK_ROWS_PER_WARP];
    bool valid[TOPK_ROWS_PER_WARP];
#pragma unroll
    for (int j = 0; j < TOPK_ROWS_PER_WARP; ++j) {
      const int r = r0 + j;
      const long long bank_row = order[r >> 6] * 64 + (r & 63);
      valid[j] = bank_row < n;
      srow[j] = valid[j] ? (perm ? perm[bank_row] : bank_row) : 0;
      sp[j] = src + srow[j] * d;
    }
    float acc[TOPK_ROWS_PER_WARP] = {0.f, 0.f, 0.f, 0.f};
    // The gather is latency bound (ncu: 1.2 TB/s, every FADD waiting on its own load), so a slab of 8 x 32 columns of
    // all 4 rows — 32 independent 128-byte loads per warp — is issued before anything is consumed.  Columns past d
    // contribute exact zeros; a lane still adds its columns l, l+32, ... in increasing order (the dense path's bits).
    for (int c0 = lane; c0 < d; c0 += 8 * 32) {
      float v[TOPK_ROWS_PER_WARP][8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + 32 * u;
#pragma unroll
        for (int j = 0; j < TOPK_ROWS_PER_WARP; ++j) v[j][u] = c < d ? __ldcs(sp[j] + c) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + 32 * u;
        const float qv = c < d ? qs[c] : 0.0f;
#pragma unroll
        for (int j = 0; j < TOPK_ROWS_PER_WARP; ++j) {
          const float df = qv - v[j][u];
          acc[j] = fmaf(df, df, acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < TOPK_ROWS_PER_WARP; ++j) {
      const float sc = -sqrtf(warp_sum(acc[j]));
      if (lane == 0) keys[r0 + j] = valid[j] ? make_key(sc, uint32_t(srow[j])) : ~uint64_t(0);
    }
  }
  __syncthreads();
  for (int size = 2; size <= padded; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < padded / 2; t += TOPK_THREADS) {
        const int i = 2 * t - (t & (stride - 1));
        cmp_swap(keys[i], keys[i + stride], (i & size) == 0);
      }
      __syncthreads();
    }
  }
  // certificate (all threads evaluate the same scalars): no row outside the candidates can reach the exact k-th score
  bool ok = true;
  if (m < n_blocks) {
    const uint64_t kth = keys[k - 1];
    ok = kth != ~uint64_t(0) && upper(best_row[m]) < key_score(kth);  // best_row[m]: best block left out
  }
  if (ok) {
    for (int i = threadIdx.x; i < k; i += TOPK_THREADS) idx_out[(long long)b * k + i] = int64_t(keys[i] & 0xffffffffull);
    if (threadIdx.x == 0) done[b] = 1;
  } else if (threadIdx.x == 0) {
    atomicAdd(n_pending, 1);
  }
}

static long long next_pow2(long long n) {
  long long p = 1;
  while (p < n) p <<= 1;
  return p;
}


// ---------------------------------------------------------------------------------------------
// K3b: k-means assignment step (reference compute_clusters with n_clusters > 1, nwhead/utils.py:230:
// per-class KMeans).  Every row is compared with the k centroids of ITS OWN class only (exact fp32
// differences).  One warp per PAIR of rows that are consecutive in class-sorted order: the two rows share
// every centroid load (the L1 wavefronts of the centroid reads, not HBM, bound a one-row-per-warp version),
// the rows bypass L1 so that the class's centroids stay there.  Algorithmic traffic: N*d*4 bytes read +
// 8 bytes per row written.
// ---------------------------------------------------------------------------------------------
constexpr int KM_TILE = 4;  // centroids per pass over a row
constexpr int KM_THREADS = 256;

// the rows are read once: keep them out of L1
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ float sqdiff4(const float4 x, const float4 c) {
  const float a = x.x - c.x, b = x.y - c.y, e = x.z - c.z, w = x.w - c.w;
  return a * a + b * b + e * e + w * w;
}

// one row against the k centroids at c0: best squared distance and its index (lowest index on ties)
template <bool VEC>
__device__ __forceinline__ void km_one(const float* __restrict__ x, const float* __restrict__ c0, int d, int k, int lane,
                                       float& best, int& best_j) {
  best = INFINITY;
  best_j = 0;
  for (int j0 = 0; j0 < k; j0 += KM_TILE) {
    const int kk = min(KM_TILE, k - j0);
    float acc[KM_TILE];
#pragma unroll
    for (int j = 0; j < KM_TILE; ++j) acc[j] = 0.f;
    if (VEC) {
      for (int e = lane * 4; e < d; e += 128) {
        const float4 xv = ld_stream_f4(x + e);
#pragma unroll
        for (int j = 0; j < KM_TILE; ++j)
          if (j < kk) acc[j] += sqdiff4(xv, __ldg(reinterpret_cast<const float4*>(c0 + (size_t)(j0 + j) * d + e)));
      }
    } else {
      for (int e = lane; e < d; e += 32) {
        const float xv = x[e];
#pragma unroll
        for (int j = 0; j < KM_TILE; ++j) {
          if (j < kk) {
            const float a = xv - __ldg(c0 + (size_t)(j0 + j) * d + e);
            acc[j] += a * a;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < KM_TILE; ++j) {
      if (j < kk) {
        const float v = warp_sum(acc[j]);
        if (v < best) {
          best = v;
          best_j = j0 + j;
        }
      }
    }
  }
}

// two rows of the same class: every centroid value is loaded once for both (same per-row arithmetic and
// summation order as km_one, so a row's result does not depend on how it was paired)
__device__ __forceinline__ void km_two(const float* __restrict__ xa, const float* __restrict__ xb,
                                       const float* __restrict__ c0, int d, int k, int lane, float& best_a, int& ja,
                                       float& best_b, int& jb) {
  best_a = best_b = INFINITY;
  ja = jb = 0;
  for (int j0 = 0; j0 < k; j0 += KM_TILE) {
    const int kk = min(KM_TILE, k - j0);
    float acc_a[KM_TILE], acc_b[KM_TILE];
#pragma unroll
    for (int j = 0; j < KM_TILE; ++j) acc_a[j] = acc_b[j] = 0.f;
    int e = lane * 4;
    for (; e + 128 < d; e += 256) {  // 4 row loads in flight per lane before the first use
      const float4 a0 = ld_stream_f4(xa + e), a1 = ld_stream_f4(xa + e + 128);
      const float4 b0 = ld_stream_f4(xb + e), b1 = ld_stream_f4(xb + e + 128);
#pragma unroll
      for (int j = 0; j < KM_TILE; ++j) {
        if (j < kk) {
          const float* cj = c0 + (size_t)(j0 + j) * d + e;
          const float4 c_lo = __ldg(reinterpret_cast<const float4*>(cj));
          const float4 c_hi = __ldg(reinterpret_cast<const float4*>(cj + 128));
          acc_a[j] += sqdiff4(a0, c_lo);
          acc_b[j] += sqdiff4(b0, c_lo);
          acc_a[j] += sqdiff4(a1, c_hi);
          acc_b[j] += sqdiff4(b1, c_hi);
        }
      }
    }
    for (; e < d; e += 128) {
      const float4 a0 = ld_stream_f4(xa + e), b0 = ld_stream_f4(xb + e);
#pragma unroll
      for (int j = 0; j < KM_TILE; ++j) {
        if (j < kk) {
          const float4 cv = __ldg(reinterpret_cast<const float4*>(c0 + (size_t)(j0 + j) * d + e));
          acc_a[j] += sqdiff4(a0, cv);
          acc_b[j] += sqdiff4(b0, cv);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < KM_TILE; ++j) {
      if (j < kk) {
        const float va = warp_sum(acc_a[j]), vb = warp_sum(acc_b[j]);
        if (va < best_a) {
          best_a = va;
          ja = j0 + j;
        }
        if (vb < best_b) {
          best_b = vb;
          jb = j0 + j;
        }
      }
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(KM_THREADS) kmeans_assign_kernel(const float* __restrict__ rows, int d, long long ld,
                                                                   const int32_t* __restrict__ group,
                                                                   const int64_t* __restrict__ order,
                                                                   long long n_rows,
                                                                   const float* __restrict__ cent, int k,
                                                                   int32_t* __restrict__ assign,
                                                                   float* __restrict__ dist) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (KM_THREADS / 32);
  const long long pairs = (n_rows + 1) / 2;
  for (long long p = (long long)blockIdx.x * (KM_THREADS / 32) + (threadIdx.x >> 5); p < pairs; p += warps) {
    const bool two = 2 * p + 1 < n_rows;
    const long long ra = order ? order[2 * p] : 2 * p;
    const long long rb = two ? (order ? order[2 * p + 1] : 2 * p + 1) : ra;
    const int ga = group[ra], gb = group[rb];
    float best_a, best_b;
    int ja, jb;
    if (VEC && two && ga == gb) {
      km_two(rows + ra * ld, rows + rb * ld, cent + (size_t)ga * k * d, d, k, lane, best_a, ja, best_b, jb);
    } else {
      km_one<VEC>(rows + ra * ld, cent + (size_t)ga * k * d, d, k, lane, best_a, ja);
      if (two) km_one<VEC>(rows + rb * ld, cent + (size_t)gb * k * d, d, k, lane, best_b, jb);
    }
    if (lane == 0) {
      assign[ra] = ga * k + ja;
      if (dist) dist[ra] = best_a;
      if (two) {
        assign[rb] = gb * k + jb;
        if (dist) dist[rb] = best_b;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k-means++ seeding with scikit-learn's arithmetic (compute_clusters with n_clusters > 1, nwhead/utils.py:230).
// sklearn centres each class (X -= X.mean(0), float32) and evaluates the squared distances of _kmeans_plusplus in
// float64, rounded to float32 (_euclidean_distances_upcast).  seed_dist: for every row, the squared distance to
// each of its class's T candidate rows, with exactly that rounding.  One warp per row, float64 accumulation.
// ---------------------------------------------------------------------------------------------
constexpr int SEED_MAX_T = 8;

__global__ void __launch_bounds__(256) kmeans_seed_dist_kernel(const float* __restrict__ rows, int d, long long ld,
                                                               const int32_t* __restrict__ group,
                                                               const int64_t* __restrict__ order, long long n_rows,
                                                               const float* __restrict__ mean,
                                                               const int64_t* __restrict__ cand, int n_cand,
                                                               float* __restrict__ dist_out) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < n_rows; p += warps) {
    const long long r = order ? order[p] : p;
    const int g = group[r];
    const float* x = rows + r * ld;
    const float* m = mean + (size_t)g * d;
    const float* cp[SEED_MAX_T];
#pragma unroll
    for (int t = 0; t < SEED_MAX_T; ++t) cp[t] = rows + cand[(size_t)g * n_cand + (t < n_cand ? t : 0)] * ld;
    double acc[SEED_MAX_T];
#pragma unroll
    for (int t = 0; t < SEED_MAX_T; ++t) acc[t] = 0.0;
    for (int c = lane; c < d; c += 32) {
      const float mc = m[c];
      const double xc = double(__fsub_rn(__ldcs(x + c), mc));  // the float32 value sklearn holds after X -= mean
#pragma unroll
      for (int t = 0; t < SEED_MAX_T; ++t) {
        if (t < n_cand) {
          const double df = xc - double(__fsub_rn(cp[t][c], mc));
          acc[t] = fma(df, df, acc[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < SEED_MAX_T; ++t) {
      if (t < n_cand) {
        double v = acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) dist_out[(size_t)t * n_rows + r] = fmaxf(float(v), 0.0f);
      }
    }
  }
}

// seed_pick: np.searchsorted(np.cumsum(closest_dist_sq), rand_vals) per class — numpy's cumsum of a float32 array
// is a sequential float32 accumulation, reproduced here by one thread per class walking its rows in class-sorted
// order.  pos_out[c, t] = source row of the first position whose running sum >= targets[c, t] (clipped to the
// class's last row, as sklearn clips).
__global__ void kmeans_seed_pick_kernel(const float* __restrict__ closest, const int64_t* __restrict__ order,
                                        const int32_t* __restrict__ offsets, int n_classes,
                                        const double* __restrict__ targets, int n_targets,
                                        int64_t* __restrict__ pos_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_classes) return;
  const long long lo = offsets[c], hi = offsets[c + 1];
  if (hi <= lo) return;
  long long found[SEED_MAX_T];
  double tg[SEED_MAX_T];
  for (int t = 0; t < SEED_MAX_T; ++t) {
    found[t] = -1;
    tg[t] = t < n_targets ? targets[(size_t)c * n_targets + t] : 0.0;
  }
  float run = 0.0f;
  int open = n_targets;
  for (long long p = lo; p < hi && open > 0; ++p) {
    run = __fadd_rn(run, closest[order ? order[p] : p]);
    for (int t = 0; t < n_targets; ++t)
      if (found[t] < 0 && double(run) >= tg[t]) {
        found[t] = p;
        --open;
      }
  }
  for (int t = 0; t < n_targets; ++t) {
    const long long p = found[t] < 0 ? hi - 1 : found[t];
    pos_out[(size_t)c * n_targets + t] = order ? order[p] : p;
  }
}

}  // namespace aux
}  // namespace nw

using namespace nw;

extern "C" int nw_kmeans_seed_dist(const float* rows, int d, int64_t ld, const int32_t* group, const int64_t* order,
                                   int64_t n_rows, const float* class_mean, const int64_t* cand_rows, int n_cand,
                                   float* dist_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(rows && group && class_mean && cand_rows && dist_out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(d > 0 && ld >= d && n_rows > 0, NW_ERR_INVALID, "bad shape");
  NW_REQUIRE(n_cand >= 1 && n_cand <= aux::SEED_MAX_T, NW_ERR_INVALID, "n_cand must be in [1, %d]", aux::SEED_MAX_T);
  const long long want = ceil_div_ll(n_rows, 8);
  const long long fit = (long long)sm_count() * 8;
  aux::kmeans_seed_dist_kernel<<<unsigned(want < fit ? want : fit), 256, 0, stream>>>(
      rows, d, ld, group, order, n_rows, class_mean, cand_rows, n_cand, dist_out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_kmeans_seed_pick(const float* closest, const int64_t* order, const int32_t* offsets, int n_classes,
                                   const double* targets, int n_targets, int64_t* pos_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(closest && offsets && targets && pos_out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_classes > 0, NW_ERR_INVALID, "n_classes must be positive");
  NW_REQUIRE(n_targets >= 1 && n_targets <= aux::SEED_MAX_T, NW_ERR_INVALID, "n_targets must be in [1, %d]",
             aux::SEED_MAX_T);
  aux::kmeans_seed_pick_kernel<<<ceil_div(n_classes, 64), 64, 0, stream>>>(closest, order, offsets, n_classes, targets,
                                                                           n_targets, pos_out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" size_t nw_class_centroids_workspace_bytes(int n_classes, int d) {
  if (n_classes <= 0 || d <= 0) return 0;
  return size_t(aux::CENT_SPLITS) * size_t(n_classes) * size_t(d) * sizeof(float);
}

extern "C" int nw_class_centroids(const float* rows, int d, int64_t ld, const int64_t* perm,
                                  const int32_t* offsets, int n_classes, float* out, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(rows && offsets && out && workspace, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(d > 0 && ld >= d && n_classes > 0, NW_ERR_INVALID, "bad shape");
  NW_REQUIRE(n_classes <= 65535 * 32, NW_ERR_UNSUPPORTED, "too many classes");
  NW_REQUIRE(workspace_bytes >= nw_class_centroids_workspace_bytes(n_classes, d), NW_ERR_WORKSPACE,
             "workspace too small");
  const bool vec = (d % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
  float* partial = static_cast<float*>(workspace);
  if (vec) {
    dim3 grid(n_classes, ceil_div(d, aux::CENT_THREADS * 4), aux::CENT_SPLITS);
    aux::centroid_partial_kernel<true><<<grid, aux::CENT_THREADS, 0, stream>>>(rows, d, ld, perm, offsets, partial,
                                                                               n_classes);
  } else {
    dim3 grid(n_classes, ceil_div(d, aux::CENT_THREADS), aux::CENT_SPLITS);
    aux::centroid_partial_kernel<false><<<grid, aux::CENT_THREADS, 0, stream>>>(rows, d, ld, perm, offsets,
                                                                                partial, n_classes);
  }
  NW_CUDA_OK(cudaGetLastError());
  const long long total = (long long)n_classes * d;
  aux::centroid_finish_kernel<<<unsigned(ceil_div_ll(total, 256)), 256, 0, stream>>>(partial, offsets, n_classes, d,
                                                                                    out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_kmeans_assign(const float* rows, int d, int64_t ld, const int32_t* group, const int64_t* order,
                                int64_t n_rows, const float* centroids, int k, int32_t* assign_out, float* dist_out,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(rows && group && centroids && assign_out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(d > 0 && ld >= d && n_rows > 0 && k > 0, NW_ERR_INVALID, "bad shape");
  const bool vec = (d % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(centroids) & 15) == 0);
  auto kernel = vec ? aux::kmeans_assign_kernel<true> : aux::kmeans_assign_kernel<false>;
  int resident = 0;  // persistent grid: exactly the blocks that fit, each striding over the row pairs
  NW_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, aux::KM_THREADS, 0));
  const long long want = ceil_div_ll((n_rows + 1) / 2, aux::KM_THREADS / 32);
  const long long fit = (long long)sm_count() * (resident > 0 ? resident : 1);
  kernel<<<unsigned(want < fit ? want : fit), aux::KM_THREADS, 0, stream>>>(rows, d, ld, group, order, n_rows,
                                                                            centroids, k, assign_out, dist_out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_onehot_argmax(const float* onehot, int64_t n_rows, int n_classes, int32_t* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(onehot && out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_rows > 0 && n_classes > 0, NW_ERR_INVALID, "shapes must be positive");
  aux::onehot_argmax_kernel<<<unsigned(ceil_div_ll(n_rows, 8)), 256, 0, stream>>>(onehot, n_rows, n_classes, out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_support_influence(const float* softmaxes, const int32_t* qlabel, const float* sweights,
                                    const int32_t* slabel, int n_query, int n_label_sets, int64_t n_support,
                                    int n_classes, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(softmaxes && qlabel && sweights && slabel && out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_query > 0 && n_support > 0 && n_classes > 0 && n_label_sets > 0, NW_ERR_INVALID,
             "shapes must be positive");
  NW_REQUIRE(n_query <= 65535 && n_label_sets <= 65535, NW_ERR_UNSUPPORTED,
             "n_query and n_label_sets are limited to 65535 per call");
  const bool vec = (n_support % 4 == 0) && ((reinterpret_cast<uintptr_t>(sweights) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(slabel) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec) {
    dim3 grid(unsigned(ceil_div_ll(n_support, 1024 * aux::INFL_VEC_PER_THREAD)), n_query, n_label_sets);
    aux::influence_kernel<true><<<grid, 256, 0, stream>>>(softmaxes, qlabel, sweights, slabel, n_support, n_classes,
                                                          n_label_sets, out);
  } else {
    dim3 grid(unsigned(ceil_div_ll(n_support, 256)), n_query, n_label_sets);
    aux::influence_kernel<false><<<grid, 256, 0, stream>>>(softmaxes, qlabel, sweights, slabel, n_support, n_classes,
                                                           n_label_sets, out);
  }
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" size_t nw_rank_rows_workspace_bytes(int n_rows, int64_t n_cols) {
  if (n_rows <= 0 || n_cols <= 0) return 0;
  return size_t(n_rows) * size_t(aux::next_pow2(n_cols)) * sizeof(uint64_t);
}

extern "C" int nw_rank_rows(const float* scores, int n_rows, int64_t n_cols, int64_t k, int64_t* idx_out,
                            void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(scores && idx_out && workspace, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_rows > 0 && n_rows <= 65535 && n_cols > 0 && n_cols < (1ll << 32), NW_ERR_INVALID, "bad shape");
  NW_REQUIRE(k > 0 && k <= n_cols, NW_ERR_INVALID, "k must be in [1, n_cols]");
  NW_REQUIRE(workspace_bytes >= nw_rank_rows_workspace_bytes(n_rows, n_cols), NW_ERR_WORKSPACE, "workspace too small");
  const long long padded = aux::next_pow2(n_cols);
  uint64_t* keys = static_cast<uint64_t*>(workspace);
  const long long chunk = aux::SORT_CHUNK;
  const long long n1 = ceil_div_ll(n_cols, chunk) * k;       // survivors of the first selection level
  const long long n2 = ceil_div_ll(n1, chunk) * k;           // ... of the second (later levels are smaller)
  if (k <= chunk / 4 && n_cols > chunk && n1 + n2 <= padded) {
    // level outputs alternate between two regions of the workspace: [0, n_rows * n1) and the n_rows * n2 after it
    uint64_t* region[2] = {keys, keys + (long long)n_rows * n1};
    // Radix selection first (one block per row); rows it could not finish (massive ties) are flagged in `todo` and
    // taken by the sort-and-keep-k levels below, which skip every other row.  The flags live behind the two regions.
    int32_t* todo = nullptr;
    if (size_t(n_rows) * size_t(n1 + n2) * sizeof(uint64_t) + size_t(n_rows) * sizeof(int32_t) <= workspace_bytes) {
      todo = reinterpret_cast<int32_t*>(keys + (long long)n_rows * (n1 + n2));
      aux::rank_radix_kernel<<<n_rows, aux::RS_THREADS, 0, stream>>>(scores, n_cols, int(k), idx_out, todo);
      NW_CUDA_OK(cudaGetLastError());
    }
    {
      dim3 grid(unsigned(ceil_div_ll(n_cols, chunk)), n_rows);
      aux::rank_select_kernel<true><<<grid, 1024, 0, stream>>>(scores, nullptr, n_cols, 0, region[0], n1, int(k), todo);
    }
    long long n_cur = n1;
    int cur = 0;
    for (;;) {
      const long long nch = ceil_div_ll(n_cur, chunk);
      dim3 grid(unsigned(nch), n_rows);
      aux::rank_select_kernel<false><<<grid, 1024, 0, stream>>>(nullptr, region[cur], n_cur, n_cur, region[cur ^ 1],
                                                                nch * k, int(k), todo);
      cur ^= 1;
      n_cur = nch * k;
      if (nch == 1) break;
    }
    dim3 grid(unsigned(ceil_div_ll(k, 256)), n_rows);
    aux::rank_emit_kernel<<<grid, 256, 0, stream>>>(region[cur], n_cur, k, idx_out, todo);
    NW_CUDA_OK(cudaGetLastError());
    return NW_OK;
  }
  {
    dim3 grid(unsigned(ceil_div_ll(padded, 256)), n_rows);
    aux::rank_init_kernel<<<grid, 256, 0, stream>>>(scores, n_cols, padded, keys);
  }
  const long long nchunks = padded > chunk ? padded / chunk : 1;
  {
    dim3 grid(unsigned(nchunks), n_rows);
    aux::rank_smem_kernel<<<grid, 1024, 0, stream>>>(keys, padded, 2, padded < chunk ? padded : chunk);
  }
  for (long long size = chunk * 2; size <= padded; size <<= 1) {
    for (long long stride = size >> 1; stride >= chunk; stride >>= 1) {
      dim3 grid(unsigned(ceil_div_ll(padded / 2, 256)), n_rows);
      aux::rank_global_step_kernel<<<grid, 256, 0, stream>>>(keys, padded, size, stride);
    }
    dim3 grid(unsigned(nchunks), n_rows);
    aux::rank_smem_kernel<<<grid, 1024, 0, stream>>>(keys, padded, size, size);
  }
  {
    dim3 grid(unsigned(ceil_div_ll(k, 256)), n_rows);
    aux::rank_emit_kernel<<<grid, 256, 0, stream>>>(keys, padded, k, idx_out, nullptr);
  }
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_topk_refine(const float* q, int n_query, int d, const float* source_rows, int64_t n_rows,
                              const int64_t* perm, const int64_t* block_order, const float* block_best_sorted,
                              int order_stride, int m_cap, int64_t n_blocks, int k, const float* q_sqnorm,
                              const float* resid_q, const float* smax_sq, const float* resid_max, int precision,
                              int32_t* done, int64_t* idx_out, int32_t* n_pending, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(q && source_rows && block_order && block_best_sorted && q_sqnorm && resid_q && smax_sq && resid_max &&
                 done && idx_out && n_pending,
             NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_query > 0 && d > 0 && n_rows > 0 && n_rows < (1ll << 32), NW_ERR_INVALID, "bad shape");
  NW_REQUIRE(m_cap >= 1 && m_cap <= 64 && m_cap <= n_blocks, NW_ERR_INVALID, "m_cap must be in [1, min(64, n_blocks)]");
  NW_REQUIRE(order_stride >= m_cap && (m_cap == n_blocks || order_stride > m_cap), NW_ERR_INVALID,
             "block_order needs m_cap + 1 entries per query");
  NW_REQUIRE(k >= 1 && k <= m_cap * 64 && k <= n_rows, NW_ERR_INVALID, "k must be in [1, min(64 m_cap, n_rows)]");
  const int padded = int(aux::next_pow2(m_cap * 64));
  const size_t smem = size_t(padded) * sizeof(uint64_t) + size_t(d) * sizeof(float);
  NW_REQUIRE(smem <= 200 * 1024, NW_ERR_UNSUPPORTED, "feature dimension %d too large for the refinement kernel", d);
  if (smem > 40 * 1024)
    NW_CUDA_OK(cudaFuncSetAttribute(aux::topk_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  aux::topk_refine_kernel<<<n_query, aux::TOPK_THREADS, smem, stream>>>(
      q, d, source_rows, n_rows, perm, block_order, block_best_sorted, order_stride, m_cap, n_blocks, k, padded,
      q_sqnorm, resid_q, smax_sq, resid_max, precision, done, idx_out, n_pending);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}
