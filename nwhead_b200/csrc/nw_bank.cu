// K0 — support-bank construction and query preparation (HBM-bound streaming kernels), plus the
// library's error plumbing.
//
// Replaces the CPU fp32 bank of the reference (nwhead/nw.py:213-243, nwhead/support.py:113-120) and
// the per-call whole-bank host->device copy (nwhead/nw.py:156) by a device-resident, class-sorted
// layout: bf16 features stored k-block-major [row_elems/64][N][64] (every 64-element k-block of a row is one
// 128-byte TMA swizzle row and every (row tile, k-block) box is contiguous in HBM), fp32 squared norms of the
// ROUNDED values, int32 labels, int32 class offsets.

#include <stdarg.h>
#include <string.h>

#include "nw_common.cuh"

namespace nw {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in %s", int(e), cudaGetErrorString(e), file, line, what);
  return NW_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

namespace k0 {

__global__ void labels_to_i32_kernel(const int64_t* __restrict__ labels, const int64_t* __restrict__ perm,
                                     long long n, int n_classes, int32_t* __restrict__ out,
                                     int32_t* __restrict__ status) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  int bad = 0, desc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long v = labels[perm ? perm[i] : i];
    if (v < 0 || v >= n_classes) ++bad;
    if (i > 0) {
      const long long pv = labels[perm ? perm[i - 1] : i - 1];
      if (v < pv) ++desc;
    }
    out[i] = int32_t(v);
  }
  if (bad) atomicAdd(status + 0, bad);
  if (desc) atomicAdd(status + 1, desc);
}

__global__ void class_offsets_kernel(const int32_t* __restrict__ lab, long long n, int n_classes,
                                     int32_t* __restrict__ offsets) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
    const int prev = i > 0 ? lab[i - 1] : -1;
    const int cur = i < n ? lab[i] : n_classes;
    for (int c = prev + 1; c <= cur; ++c) offsets[c] = int32_t(i);
  }
}

constexpr int MEAN_SPLITS = 256;

// partial[r][col] = sum over the r-th slab of rows
__global__ void column_partial_kernel(const float* __restrict__ rows, long long n, int d, long long ld,
                                      float* __restrict__ partial) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (n + MEAN_SPLITS - 1) / MEAN_SPLITS;
  const long long r0 = (long long)blockIdx.y * per;
  const long long r1 = r0 + per < n ? r0 + per : n;
  if (col >= d) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  long long r = r0;
  for (; r + 3 < r1; r += 4) {
    a0 += rows[r * ld + col];
    a1 += rows[(r + 1) * ld + col];
    a2 += rows[(r + 2) * ld + col];
    a3 += rows[(r + 3) * ld + col];
  }
  for (; r < r1; ++r) a0 += rows[r * ld + col];
  partial[(long long)blockIdx.y * d + col] = (a0 + a1) + (a2 + a3);
}

__global__ void column_finish_kernel(const float* __restrict__ partial, long long n, int d,
                                     float* __restrict__ mean) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= d) return;
  double acc = 0.0;
  for (int r = 0; r < MEAN_SPLITS; ++r) acc += double(partial[(long long)r * d + col]);
  mean[col] = float(acc / double(n));
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return uint32_t(__bfloat16_as_ushort(a)) | (uint32_t(__bfloat16_as_ushort(b)) << 16);
}

// Where the converted rows go.  One destination for the bank and for single-GPU queries; with the bank sharded
// over several GPUs every rank converts ITS slice of the query batch and stores it straight into every rank's
// query buffer (peer-mapped symmetric memory, NVLink stores): the queries are replicated by the conversion kernel
// itself — no fp32 all-gather, no R-fold redundant conversion.
constexpr int MAX_DEST = 16;
struct RowDest {
  __nv_bfloat16* out[MAX_DEST];  // (row_elems / 64, n_total, 64) each
  float* sqnorm[MAX_DEST];       // (n_total) each, or NULL
  int n_dest;
  long long n_total;             // rows of a destination buffer
  long long row_offset;          // destination row of source row 0
};

// One warp per output row.  VEC: 16-byte loads (d % 4 == 0, ld % 4 == 0, 16-B aligned base).
template <bool VEC>
__global__ void __launch_bounds__(256) rows_to_bf16_kernel(const float* __restrict__ rows, long long n, int d,
                                                           long long ld, const int64_t* __restrict__ perm,
                                                           const float* __restrict__ center, int normalize,
                                                           int layout, int precision, const RowDest dest,
                                                           int row_elems) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* src = rows + (perm ? perm[row] : row) * ld;
  // k-block-major output: element (row, col) lives at ((col / 64) * n_total + row) * 64 + col % 64, so that every
  // (row tile, k-block) box the fused forward loads with TMA is one contiguous run of 128-byte rows in HBM
  const long long drow = dest.row_offset + row;
  auto off = [&](int col) -> long long { return ((long long)(col >> 6) * dest.n_total + drow) * 64 + (col & 63); };

  float inv = 1.0f;
  if (normalize) {
    float ss = 0.f;
    if (VEC) {
      for (int c = lane * 4; c < d; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(src + c);
        if (center) {
          const float4 m = *reinterpret_cast<const float4*>(center + c);
          v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
        }
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    } else {
      for (int c = lane; c < d; c += 32) {
        const float v = src[c] - (center ? center[c] : 0.f);
        ss += v * v;
      }
    }
    ss = warp_sum(ss);
    inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(|x|, eps)
  }

  // segment placement: bank rows [hi | hi | lo], query rows [hi | lo | hi]
  const int seg_hi2 = (layout == NW_ROWS_BANK) ? d : 2 * d;
  const int seg_lo = (layout == NW_ROWS_BANK) ? 2 * d : d;
  float sq = 0.f;
  if (VEC) {
#pragma unroll 4  // several 16-byte loads in flight per lane: a 2048-wide row is 16 iterations
    for (int c = lane * 4; c < d; c += 128) {
      float4 v = *reinterpret_cast<const float4*>(src + c);
      if (center) {
        const float4 m = *reinterpret_cast<const float4*>(center + c);
        v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
      }
      float x[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        lo[i] = __float2bfloat16_rn(0.f);
        hi[i] = __float2bfloat16_rn(x[i]);
        const float h = __bfloat162float(hi[i]);
        if (precision == NW_PREC_BF16X3) {
          lo[i] = __float2bfloat16_rn(x[i] - h);
          const float l = __bfloat162float(lo[i]);
          sq += h * h + 2.0f * h * l;  // exactly what hi.hi + lo.hi + hi.lo contracts to on the diagonal
        } else {
          sq += h * h;
        }
      }
      const uint2 ph = make_uint2(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]));
      const uint2 pl = make_uint2(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]));
      for (int r = 0; r < dest.n_dest; ++r) {
        *reinterpret_cast<uint2*>(dest.out[r] + off(c)) = ph;
        if (precision == NW_PREC_BF16X3) {
          *reinterpret_cast<uint2*>(dest.out[r] + off(seg_hi2 + c)) = ph;
          *reinterpret_cast<uint2*>(dest.out[r] + off(seg_lo + c)) = pl;
        }
      }
    }
  } else {
    for (int c = lane; c < d; c += 32) {
      const float x = (src[c] - (center ? center[c] : 0.f)) * inv;
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      const float h = __bfloat162float(hi);
      __nv_bfloat16 lo = __float2bfloat16_rn(0.f);
      if (precision == NW_PREC_BF16X3) {
        lo = __float2bfloat16_rn(x - h);
        const float l = __bfloat162float(lo);
        sq += h * h + 2.0f * h * l;
      } else {
        sq += h * h;
      }
      for (int r = 0; r < dest.n_dest; ++r) {
        dest.out[r][off(c)] = hi;
        if (precision == NW_PREC_BF16X3) {
          dest.out[r][off(seg_hi2 + c)] = hi;
          dest.out[r][off(seg_lo + c)] = lo;
        }
      }
    }
  }
  for (int c = precision * d + lane; c < row_elems; c += 32)
    for (int r = 0; r < dest.n_dest; ++r) dest.out[r][off(c)] = __float2bfloat16_rn(0.f);
  sq = warp_sum(sq);
  if (lane == 0)
    for (int r = 0; r < dest.n_dest; ++r)
      if (dest.sqnorm[r]) dest.sqnorm[r][drow] = sq;
}

// Squared norm of what the bf16 rounding discards, per row: |x - hi|^2 (NW_PREC_BF16) or |x - hi - lo|^2
// (NW_PREC_BF16X3), with x = rows[i, :] - center exactly as rows_to_bf16_kernel forms it.  One warp per row.
__global__ void __launch_bounds__(256) rounding_residual_kernel(const float* __restrict__ rows, long long n, int d,
                                                                long long ld, const float* __restrict__ center,
                                                                int precision, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* src = rows + row * ld;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float x = src[c] - (center ? center[c] : 0.f);
    float r = x - __bfloat162float(__float2bfloat16_rn(x));
    if (precision == NW_PREC_BF16X3) r -= __bfloat162float(__float2bfloat16_rn(r));
    acc += r * r;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

}  // namespace k0
}  // namespace nw

using namespace nw;

extern "C" const char* nw_last_error(void) { return g_err; }
extern "C" int nw_abi_version(void) { return NW_ABI_VERSION; }

extern "C" int nw_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute", __FILE__, __LINE__);
  NW_REQUIRE(major == 10, NW_ERR_UNSUPPORTED,
             "libnw_sm100 needs a compute-capability 10.x (B200) device, found major %d; there is no fallback",
             major);
  return NW_OK;
}

extern "C" int nw_row_elems(int d, int precision) {
  if (d <= 0 || (precision != NW_PREC_BF16 && precision != NW_PREC_BF16X3)) return NW_ERR_INVALID;
  return ((precision * d + 63) / 64) * 64;
}

extern "C" int nw_labels_to_i32(const int64_t* labels_i64, const int64_t* perm, int64_t n, int n_classes,
                                int32_t* labels_out, int32_t* status_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(labels_i64 && labels_out && status_out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n > 0 && n_classes > 0, NW_ERR_INVALID, "n and n_classes must be positive");
  NW_CUDA_OK(cudaMemsetAsync(status_out, 0, 2 * sizeof(int32_t), stream));
  const int blocks = int(ceil_div_ll(n, 256) < 2048 ? ceil_div_ll(n, 256) : 2048);
  k0::labels_to_i32_kernel<<<blocks, 256, 0, stream>>>(labels_i64, perm, n, n_classes, labels_out, status_out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_class_offsets(const int32_t* labels_sorted, int64_t n, int n_classes, int32_t* offsets,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(labels_sorted && offsets, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n > 0 && n_classes > 0, NW_ERR_INVALID, "n and n_classes must be positive");
  const int blocks = int(ceil_div_ll(n + 1, 256) < 2048 ? ceil_div_ll(n + 1, 256) : 2048);
  k0::class_offsets_kernel<<<blocks, 256, 0, stream>>>(labels_sorted, n, n_classes, offsets);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" size_t nw_column_mean_workspace_bytes(int d) {
  return d > 0 ? size_t(k0::MEAN_SPLITS) * size_t(d) * sizeof(float) : 0;
}

extern "C" int nw_column_mean(const float* rows, int64_t n, int d, int64_t ld, float* mean_out, void* workspace,
                              size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(rows && mean_out && workspace, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n > 0 && d > 0 && ld >= d, NW_ERR_INVALID, "bad shape n=%lld d=%d ld=%lld", (long long)n, d, (long long)ld);
  NW_REQUIRE(workspace_bytes >= nw_column_mean_workspace_bytes(d), NW_ERR_WORKSPACE, "workspace too small");
  dim3 grid(ceil_div(d, 128), k0::MEAN_SPLITS);
  k0::column_partial_kernel<<<grid, 128, 0, stream>>>(rows, n, d, ld, static_cast<float*>(workspace));
  NW_CUDA_OK(cudaGetLastError());
  k0::column_finish_kernel<<<ceil_div(d, 128), 128, 0, stream>>>(static_cast<const float*>(workspace), n, d, mean_out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

static int rows_to_bf16_impl(const float* rows, int64_t n, int d, int64_t ld, const int64_t* perm, const float* center,
                             int normalize, int layout, int precision, const k0::RowDest& dest, int row_elems,
                             cudaStream_t stream) {
  NW_REQUIRE(rows != nullptr, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n > 0 && d > 0 && ld >= d, NW_ERR_INVALID, "bad shape n=%lld d=%d ld=%lld", (long long)n, d, (long long)ld);
  NW_REQUIRE(layout == NW_ROWS_BANK || layout == NW_ROWS_QUERY, NW_ERR_INVALID, "unknown layout %d", layout);
  NW_REQUIRE(precision == NW_PREC_BF16 || precision == NW_PREC_BF16X3, NW_ERR_INVALID, "unknown precision %d", precision);
  NW_REQUIRE(row_elems == nw_row_elems(d, precision), NW_ERR_INVALID, "row_elems %d != nw_row_elems(%d, %d)",
             row_elems, d, precision);
  bool vec = (d % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0) &&
             (!center || (reinterpret_cast<uintptr_t>(center) & 15) == 0);
  for (int r = 0; r < dest.n_dest; ++r) {
    NW_REQUIRE(dest.out[r] != nullptr, NW_ERR_INVALID, "NULL destination %d", r);
    vec = vec && ((reinterpret_cast<uintptr_t>(dest.out[r]) & 7) == 0);
  }
  const int warps_per_block = 8;
  const long long blocks = ceil_div_ll(n, warps_per_block);
  NW_REQUIRE(blocks < (1ll << 31), NW_ERR_UNSUPPORTED, "too many rows");
  if (vec)
    k0::rows_to_bf16_kernel<true><<<unsigned(blocks), warps_per_block * 32, 0, stream>>>(
        rows, n, d, ld, perm, center, normalize, layout, precision, dest, row_elems);
  else
    k0::rows_to_bf16_kernel<false><<<unsigned(blocks), warps_per_block * 32, 0, stream>>>(
        rows, n, d, ld, perm, center, normalize, layout, precision, dest, row_elems);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_rows_to_bf16(const float* rows, int64_t n, int d, int64_t ld, const int64_t* perm,
                               const float* center, int normalize, int layout, int precision, void* out_bf16,
                               int row_elems, float* sqnorm_out, void* stream_) {
  NW_REQUIRE(out_bf16 != nullptr, NW_ERR_INVALID, "NULL pointer argument");
  k0::RowDest dest = {};
  dest.out[0] = static_cast<__nv_bfloat16*>(out_bf16);
  dest.sqnorm[0] = sqnorm_out;
  dest.n_dest = 1;
  dest.n_total = n;
  dest.row_offset = 0;
  return rows_to_bf16_impl(rows, n, d, ld, perm, center, normalize, layout, precision, dest, row_elems,
                           static_cast<cudaStream_t>(stream_));
}

extern "C" int nw_rows_to_bf16_peers(const float* rows, int64_t n, int d, int64_t ld, const float* center,
                                     int normalize, int precision, void* const* out_bf16_host,
                                     float* const* sqnorm_host, int n_dest, int64_t n_total, int64_t row_offset,
                                     int row_elems, void* stream_) {
  NW_REQUIRE(out_bf16_host != nullptr, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_dest >= 1 && n_dest <= k0::MAX_DEST, NW_ERR_INVALID, "n_dest must be in [1, %d]", k0::MAX_DEST);
  NW_REQUIRE(row_offset >= 0 && row_offset + n <= n_total, NW_ERR_INVALID,
             "rows [%lld, %lld) do not fit a destination of %lld rows", (long long)row_offset,
             (long long)(row_offset + n), (long long)n_total);
  k0::RowDest dest = {};
  for (int r = 0; r < n_dest; ++r) {
    dest.out[r] = static_cast<__nv_bfloat16*>(out_bf16_host[r]);
    dest.sqnorm[r] = sqnorm_host ? sqnorm_host[r] : nullptr;
  }
  dest.n_dest = n_dest;
  dest.n_total = n_total;
  dest.row_offset = row_offset;
  return rows_to_bf16_impl(rows, n, d, ld, nullptr, center, normalize, NW_ROWS_QUERY, precision, dest, row_elems,
                           static_cast<cudaStream_t>(stream_));
}

extern "C" int nw_rounding_residual(const float* rows, int64_t n, int d, int64_t ld, const float* center,
                                    int precision, float* resid_sq_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(rows && resid_sq_out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n > 0 && d > 0 && ld >= d, NW_ERR_INVALID, "bad shape n=%lld d=%d ld=%lld", (long long)n, d, (long long)ld);
  NW_REQUIRE(precision == NW_PREC_BF16 || precision == NW_PREC_BF16X3, NW_ERR_INVALID, "unknown precision %d", precision);
  const long long blocks = ceil_div_ll(n, 8);
  NW_REQUIRE(blocks < (1ll << 31), NW_ERR_UNSUPPORTED, "too many rows");
  k0::rounding_residual_kernel<<<unsigned(blocks), 256, 0, stream>>>(rows, n, d, ld, center, precision, resid_sq_out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

// ---------------------------------------------------------------------------------------------
// Operand plumbing of the tensor-core backward (nw_backward_coefficients / nw_dense_products): HBM-bound.
// ---------------------------------------------------------------------------------------------
namespace nw {
namespace k0 {

// in  (kblocks, n_rows, 64) bf16 k-block-major: element f of row r at in[f / 64][r][f % 64]
// out (ceil(n_rows / 64), kblocks * 64, 64): the TRANSPOSED matrix in the same layout, out[r / 64][f][r % 64]
// (rows beyond n_rows are written as zeros).  One block moves one 64 x 64 tile through shared memory: 128-byte
// row segments on both sides.
__global__ void transpose_kblocks_kernel(const __nv_bfloat16* __restrict__ in, long long n_rows, int kblocks,
                                         __nv_bfloat16* __restrict__ out) {
  __shared__ __nv_bfloat16 tile[64][66];
  const long long rb = blockIdx.x;  // row block
  const int kb = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads, 2 elements per thread and pass
  for (int i = ty; i < 64; i += 8) {
    const long long r = rb * 64 + i;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.0f, 0.0f);
    if (r < n_rows) v = *reinterpret_cast<const __nv_bfloat162*>(in + ((long long)kb * n_rows + r) * 64 + 2 * tx);
    tile[i][2 * tx] = v.x;
    tile[i][2 * tx + 1] = v.y;
  }
  __syncthreads();
  const long long f_rows = (long long)kblocks * 64;
  for (int i = ty; i < 64; i += 8) {  // output row f = kb * 64 + i holds column i of the tile
    __nv_bfloat162 v;
    v.x = tile[2 * tx][i];
    v.y = tile[2 * tx + 1][i];
    *reinterpret_cast<__nv_bfloat162*>(out + (rb * f_rows + kb * 64 + i) * 64 + 2 * tx) = v;
  }
}

// out[dst(r)][c] = raw[r][c] - sums[r] * rows[c / 64][r][c % 64]   (sums == NULL: a plain copy),
// dst(r) = perm ? perm[r] : r.  One warp per row, four columns per lane and pass.
__global__ void backward_finish_kernel(const float* __restrict__ raw, long long ld_raw,
                                       const __nv_bfloat16* __restrict__ rows, const float* __restrict__ sums,
                                       const int64_t* __restrict__ perm, long long n_rows, int d,
                                       float* __restrict__ out, long long ld_out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const float s = sums ? __ldg(sums + r) : 0.0f;
  const long long dst = perm ? perm[r] : r;
  const float* src = raw + r * ld_raw;
  float* o = out + dst * ld_out;
  const bool vec = ((ld_raw | ld_out) & 3) == 0 && (d & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    for (int c = lane * 4; c < d; c += 128) {
      float4 v = __ldcs(reinterpret_cast<const float4*>(src + c));
      if (sums) {
        const uint2 b = *reinterpret_cast<const uint2*>(rows + ((long long)(c >> 6) * n_rows + r) * 64 + (c & 63));
        const __nv_bfloat162 b0 = *reinterpret_cast<const __nv_bfloat162*>(&b.x);
        const __nv_bfloat162 b1 = *reinterpret_cast<const __nv_bfloat162*>(&b.y);
        v.x = fmaf(-s, __low2float(b0), v.x);
        v.y = fmaf(-s, __high2float(b0), v.y);
        v.z = fmaf(-s, __low2float(b1), v.z);
        v.w = fmaf(-s, __high2float(b1), v.w);
      }
      __stcs(reinterpret_cast<float4*>(o + c), v);
    }
  } else {
    for (int c = lane; c < d; c += 32) {
      float v = src[c];
      if (sums) v = fmaf(-s, __bfloat162float(rows[((long long)(c >> 6) * n_rows + r) * 64 + (c & 63)]), v);
      o[c] = v;
    }
  }
}

}  // namespace k0
}  // namespace nw

extern "C" int nw_transpose_kblocks(const void* in_bf16, int64_t n_rows, int kblocks, void* out_bf16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(in_bf16 && out_bf16, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_rows > 0 && kblocks > 0 && kblocks <= 65535, NW_ERR_INVALID, "bad shape n_rows=%lld kblocks=%d",
             (long long)n_rows, kblocks);
  const long long row_blocks = ceil_div_ll(n_rows, 64);
  NW_REQUIRE(row_blocks < (1ll << 31), NW_ERR_UNSUPPORTED, "too many rows");
  k0::transpose_kblocks_kernel<<<dim3(unsigned(row_blocks), unsigned(kblocks)), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(in_bf16), n_rows, kblocks, static_cast<__nv_bfloat16*>(out_bf16));
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_backward_finish(const float* raw, int64_t ld_raw, const void* rows_bf16, const float* row_sums,
                                  const int64_t* perm, int64_t n_rows, int d, float* out, int64_t ld_out,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(raw && out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(row_sums == nullptr || rows_bf16 != nullptr, NW_ERR_INVALID, "row_sums needs rows_bf16");
  NW_REQUIRE(n_rows > 0 && d > 0 && ld_raw >= d && ld_out >= d, NW_ERR_INVALID, "bad shape");
  const long long blocks = ceil_div_ll(n_rows, 8);
  NW_REQUIRE(blocks < (1ll << 31), NW_ERR_UNSUPPORTED, "too many rows");
  k0::backward_finish_kernel<<<unsigned(blocks), 256, 0, stream>>>(
      raw, ld_raw, static_cast<const __nv_bfloat16*>(rows_bf16), row_sums, perm, n_rows, d, out, ld_out);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}
