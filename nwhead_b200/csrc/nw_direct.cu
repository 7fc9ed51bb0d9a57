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

// one warp per (query, support) pair
__global__ void __launch_bounds__(256) scores_kernel(int kind, float scale, const float* __restrict__ q,
                                                     int n_query, int d, const float* __restrict__ s,
                                                     long long n_support, int batched,
                                                     float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= (long long)n_query * n_support) return;
  const long long b = pair / n_support;
  const long long j = pair - b * n_support;
  const float* qp = q + b * d;
  const float* sp = s + (batched ? pair : j) * d;
  float out;
  if (kind == NW_KIND_EUCLIDEAN) {
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float df = qp[c] - sp[c];
      acc = fmaf(df, df, acc);
    }
    out = -sqrtf(warp_sum(acc));
  } else if (kind == NW_KIND_DOT) {
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc = fmaf(qp[c], sp[c], acc);
    out = warp_sum(acc);
  } else {
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
        // separately rounded products (no FMA contraction): identical rows must give exactly 0
        const float df = __fsub_rn(__fmul_rn(qp[c], iq), __fmul_rn(sp[c], is));
        acc = fmaf(df, df, acc);
      }
      out = -sqrtf(warp_sum(acc));
    } else {
      for (int c = lane; c < d; c += 32) acc = fmaf(__fmul_rn(qp[c], iq), __fmul_rn(sp[c], is), acc);
      out = warp_sum(acc);
      if (kind == NW_KIND_CLIP) out *= scale;
    }
  }
  if (lane == 0) scores[pair] = out;
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
  if (bad && b == 0) atomicAdd(status, bad);
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
  for (long long j = threadIdx.x; j < n_support; j += blockDim.x) {
    const float s = sc[j];
    const float p = expf(s - z);
    const float gs = p * (gP[lab[j]] - dsum);
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
  const long long pairs = (long long)n_query * n_support;
  const long long blocks = ceil_div_ll(pairs, 8);
  NW_REQUIRE(blocks < (1ll << 31), NW_ERR_UNSUPPORTED, "too many (query, support) pairs for the direct path");
  direct::scores_kernel<<<unsigned(blocks), 256, 0, stream>>>(kind, scale, q, n_query, d, s, n_support,
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
  NW_CUDA_OK(cudaMemsetAsync(status_out, 0, sizeof(int32_t), stream));
  direct::aggregate_kernel<<<n_query, 256, 0, stream>>>(scores, labels, labels_batched, n_support, n_classes, logp,
                                                        row_lse, status_out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int64_t nw_direct_backward_workspace_elems(int n_query, int64_t n_support, int support_batched) {
  const int64_t pairs = int64_t(n_query) * n_support;
  return pairs + (support_batched ? pairs : n_support) + n_query;
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
  NW_REQUIRE(d <= 12288 && n_classes <= 12288, NW_ERR_UNSUPPORTED,
             "direct backward stages one row / one class table in 48 KB of shared memory (d, C <= 12288)");
  const long long pairs = (long long)n_query * n_support;
  NW_REQUIRE(pairs < (1ll << 31), NW_ERR_UNSUPPORTED, "too many (query, support) pairs for the direct path");
  float* coef = workspace;
  float* inv_s = workspace + pairs;
  float* inv_q = inv_s + (support_batched ? pairs : n_support);
  const bool norm = direct::kind_normalised(kind);
  if (norm) {
    const long long srows = support_batched ? pairs : n_support;
    direct::inv_norm_kernel<<<unsigned(ceil_div_ll(srows, 8)), 256, 0, stream>>>(s, srows, d, inv_s);
    direct::inv_norm_kernel<<<unsigned(ceil_div(n_query, 8)), 256, 0, stream>>>(q, n_query, d, inv_q);
    NW_CUDA_OK(cudaGetLastError());
  }
  direct::coef_kernel<<<n_query, 256, n_classes * sizeof(float), stream>>>(
      kind, scale, scores, row_lse, logp, grad_out, labels, labels_batched, n_support, n_classes, coef,
      kind == NW_KIND_CLIP ? grad_scale_rows : nullptr);
  NW_CUDA_OK(cudaGetLastError());
  if (grad_q) {
    direct::grad_q_kernel<<<n_query, 256, d * sizeof(float), stream>>>(kind, q, d, s, n_support, support_batched,
                                                                       coef, inv_q, inv_s, grad_q);
    NW_CUDA_OK(cudaGetLastError());
  }
  if (grad_s) {
    if (support_batched)
      direct::grad_s_batched_kernel<<<unsigned(pairs), 128, 0, stream>>>(kind, q, d, s, n_support, coef, inv_q,
                                                                         inv_s, grad_s);
    else
      direct::grad_s_shared_kernel<<<unsigned(n_support), 256, d * sizeof(float), stream>>>(
          kind, q, n_query, d, s, n_support, coef, inv_q, inv_s, grad_s);
    NW_CUDA_OK(cudaGetLastError());
  }
  return NW_OK;
}
