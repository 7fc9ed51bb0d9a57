// K1 — fused NW-head forward on tcgen05 tensor cores (sm_100a).
//
// Replaces NWHead.forward (reference nwhead/nw.py:266-289) for a shared, class-sorted support bank:
//   scores = kernel(q, S)            -> TMA-fed tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) + scalar epilogue
//   softmax over supports            -> online running max / sum in the epilogue warps (exp2 domain)
//   probs @ onehot(labels)           -> per-class segment sums, flushed on (warp-uniform) class change
// The (B, N) score matrix never leaves the SM.  Output is the per-class log-sum-exp table
// class_lse (B, C); nw_logp_from_class_lse turns it into log(P + 1e-12).
//
// Operands are k-block-major bf16 ([row_elems/64][rows][64], see nw_bank.cu): every TMA box is one contiguous
// run of 128-byte swizzle rows.  Two tile shapes: a single CTA computes 128 queries x 256 supports (UMMA
// M=128), a CTA pair (cluster of 2, cta_group::2) 256 x 256 (UMMA M=256) with each CTA staging half of the
// support tile.  TMEM holds two 256-column accumulators so the epilogue of one tile overlaps the MMAs of the next.
//
// Work decomposition (persistent): the bank is cut into `chunks` contiguous ranges of 256-row support tiles; a
// work unit is (chunk, query group of 128 or 256 rows).  Units of one chunk are adjacent in the unit order, so
// the workers of a wave sweep the same support tiles at the same time and the bank is read from HBM ~once (L2
// hits for the other query groups).  Inside a unit the epilogue carries (running max, class sum) in registers
// across tiles; classes cut by a chunk boundary go to `side` and are merged in fixed order by merge_side_kernel
// (deterministic, no atomics).  Results can be stored to several tables (peer GPUs over NVLink) for the
// bank-sharded multi-GPU path.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM allocator, warps 4-7 epilogue
// (one thread per query row = TMEM lane); the emit modes (dense scores / support influence, no state along
// the columns) add warps 8-11 as a second epilogue set.

#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "nw_common.cuh"

namespace nw {
namespace k1 {

constexpr int BM = 128;  // queries per CTA tile (TMEM lanes)
constexpr int BN = 256;  // supports per tile    (UMMA N, TMEM columns)
constexpr int BK = 64;   // bf16 per k-block = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 6;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512 = all of TMEM
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_WARP0 = 4;
constexpr int MAX_EPI_WARPS = 8;
constexpr int MAX_THREADS = (EPI_WARP0 + MAX_EPI_WARPS) * 32;
constexpr int QUAD_SETS = 4;  // epilogue sets of the QUAD variant (very short GEMMs, d <= 512)
constexpr int QUAD_THREADS = (EPI_WARP0 + 4 * QUAD_SETS) * 32;
// Epilogue warps: one thread per query row (4 warps = 128 TMEM lanes).  A second set of 4 warps (same lane
// quarters) takes every other pair of 32-column chunks: always in the emit modes (no state along the columns),
// and in the class-LSE mode when the GEMM is short (d <= 1024) and the epilogue would otherwise be the
// bottleneck.  There each set keeps its OWN running (max, sum) per class and stores to its OWN table; the two
// tables are combined afterwards by one log-add-exp pass, so the sets never synchronise with each other.
// At d <= 512 the two MUFU ops per score take as long as the score's MMAs, and ncu showed both pipes only ~60 %
// busy with two epilogue warps per scheduler (latency-bound: the epilogue cannot keep the MUFU queue full).  The
// QUAD variant runs FOUR sets (16 epilogue warps, one 64-column chunk pair of every tile each, one 32-column chunk
// in registers at a time so that 640 threads fit the register file).
constexpr int NW_MAX_PEERS = 16;

constexpr int MODE_CLASS_LSE = 0;       // online softmax + per-class sums (the NW head)
constexpr int MODE_EMIT_SCORES = 1;     // dense per-pair output: the similarity scores
constexpr int MODE_EMIT_INFLUENCE = 2;  // dense per-pair output: support influence
constexpr int MODE_EMIT_BLOCKBEST = 3;  // best score of every block of 64 support rows (candidate search for top-k)
constexpr int MODE_EMIT_COEF = 4;       // backward coefficients w(row, col), bf16, k-block-major over the columns
// modes whose epilogue goes through the per-warp 32 x 32 transpose buffers (row-major fp32 output)
constexpr int MODE_EMIT_GRADT = 5;      // dense products, corrected and stored TRANSPOSED: the grad_s GEMM of the backward
constexpr bool mode_transposes(int mode) {
  return mode != MODE_CLASS_LSE && mode != MODE_EMIT_COEF && mode != MODE_EMIT_GRADT;
}

// NCTA = 1: one CTA computes a 128 x 256 tile (UMMA M=128).
// NCTA = 2: a CTA pair (cluster of 2, cta_group::2) computes a 256 x 256 tile (UMMA M=256); each CTA stages
//           its own 128 query rows and HALF of the support tile, so L2->SM traffic and shared-memory operand
//           reads per FLOP drop by a third and the smaller stages allow a 6-deep TMA ring.
template <int NCTA, int MODE = 0, bool QUAD = false>
struct Cfg {
  // the emit modes give ring stages up for the per-warp transpose buffers of their epilogue (8 warps: one stage,
  // the 16 warps of the QUAD variant: two)
#ifndef NW_K1_STAGES_PAIR
#define NW_K1_STAGES_PAIR 6  // developer A/B builds: TMA ring depth of the CTA-pair kernels (32 KB per stage and CTA)
#endif
  static constexpr int STAGES = (NCTA == 1 ? 4 : NW_K1_STAGES_PAIR) - (!mode_transposes(MODE) ? 0 : (QUAD ? 2 : 1));
  static constexpr int B_ROWS = BN / NCTA;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};

// Tile metadata slots filled by the TMA producer with two bulk copies per tile (class-LSE mode, 16-byte aligned
// s_sqnorm / labels, tiles that lie completely inside the bank).  The epilogue sets then only wait on an mbarrier:
// the per-set staging below (4 global loads + 5 shared stores per thread, a named barrier and the prefetch of the
// next tile, 11 % of the epilogue's warp time at d = 256 in ncu's source view) remains for edge tiles, unaligned
// arrays and the emit modes.
constexpr int META_SLOTS = 4;
constexpr uint32_t META_CADD_BYTES = BN * 4;
constexpr uint32_t META_LAB_BYTES = (BN + 4) * 4;  // + the look-ahead label, rounded up to 16 bytes

struct __align__(16) TileMeta {
  float cadd[BN];    // per-column additive term: |s|^2 (EUCLID) or 0 (LINEAR); +inf / -inf for padding columns
  int lab[BN + 8];   // labels of the tile's columns plus one look-ahead entry
};

// What the class-end stores need to know about the unit a set is working on; written once per unit by one thread
// of the set (double-buffered by unit parity), read from shared memory at class ends.  It used to live in a
// per-thread struct in LOCAL memory: with 219 KB of shared memory the L1 is a few KB, every class end paid four
// dependent L2 round trips (~880 cycles, 6.5 % of the epilogue at d = 256 in ncu's source view).
struct __align__(16) UnitInfo {
  int cf, cl;  // first / last class of the unit's support range
  int cuts;    // bit 0: class cf continues from the previous chunk; bit 1: class cl continues into the next chunk
  int g;       // chunk index (row block of `side`)
};

template <int MODE, bool QUAD = false>
struct SmemTail {
  TileMeta meta[QUAD ? QUAD_SETS : 2][ACC_STAGES];  // [epilogue set][accumulator stage]: the sets stay independent
  UnitInfo unit[QUAD ? QUAD_SETS : 2][2];           // [epilogue set][unit parity]
  TileMeta mslot[META_SLOTS];                       // class-LSE: tile metadata delivered by the producer (bulk copies)
  uint64_t mfull[META_SLOTS];
  uint64_t mempty[META_SLOTS];
  float stage[!mode_transposes(MODE) ? 1 : (QUAD ? 4 * QUAD_SETS : MAX_EPI_WARPS)][32][33];  // emit modes: per-warp 32x32 transpose buffers
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tfull[ACC_STAGES];
  uint64_t tempty[ACC_STAGES];
  uint32_t tmem_base;
};

template <int NCTA, int MODE, bool QUAD = false>
constexpr size_t smem_bytes() {
  return 1024 /*align slack*/ + size_t(Cfg<NCTA, MODE, QUAD>::STAGES) * Cfg<NCTA, MODE, QUAD>::STAGE_BYTES +
         sizeof(SmemTail<MODE, QUAD>);
}

struct Params {
  const float* q_sqnorm;
  const float* s_sqnorm;
  const int32_t* labels;
  float* lse[NW_MAX_PEERS];  // class-LSE tables the results are stored to (local + peer GPUs over NVLink P2P)
  int n_tables;
  int rows_per_table;  // 0: store every entry to ALL tables; > 0: only to table[row / rows_per_table]
  float* side;
  int n_query;
  int n_support;
  int n_classes;
  int kblocks;
  int q_groups;  // query tiles of 128 * NCTA rows
  int s_keep;    // support tiles are loaded with the L2 evict_last policy (several query groups re-read them)
  int s_tiles;
  int chunks;
  int tiles_per_chunk;
  float scale_log2;  // LINEAR: scale * log2(e)
  int sets;          // epilogue warp sets (1, 2 or 4); class-LSE with several sets: set s stores to lse[s]
  int meta_bulk;     // class-LSE: the producer delivers the metadata of interior tiles (see META_SLOTS)
  int krot_step;            // query group qg walks the k-blocks of every tile starting at (qg * krot_step) % count:
                            // the CTA pairs that share a support tile then ask the L2 for DIFFERENT lines at any
                            // one time (0: all start at k-block 0)
  int* tile_gate;           // [chunk][2] arrival counters, zeroed before the launch, or NULL: the producers of the CTAs
                            // that share a chunk in a wave start every support tile together (see the producer)
  int l2_prefetch;          // developer experiment (NW_B200_L2_PREFETCH=<k-blocks ahead>, 0 = off): support boxes are
                            // prefetched into L2 this far ahead of the ring; > 0: by query group 0 of every chunk
                            // only, < 0: by every group (distance = -value)
  int stagger_ns;           // developer probe (NW_B200_STAGGER_NS): query group g delays its first load by g * this
  int debug_skip_epilogue;  // developer probe (NW_B200_DEBUG_SKIP_EPI=1): accumulators are released unread -> the
                            // speed of the TMA + MMA mainloop alone (results are garbage)
  // ---- MODE_EMIT only: one output value per (query, support) pair
  float* emit_out;           // (B, ld_out)
  long long emit_ld;
  int emit_kind;             // NW_EMIT_SCORES | NW_EMIT_INFLUENCE
  int emit_vec;              // 1: 16-byte stores are aligned
  unsigned long long* clock_probe;  // diagnostics (nw_forward_set_clock_probe) or NULL: per-CTA SM cycles + ns
  const float* row_lse;      // (B) logsumexp_j score(b, j)          [influence]
  const float* p_query;      // (B) softmax mass of the query's class [influence]
  const int32_t* qlabel;     // (B) query labels                      [influence]; row labels [coefficients, orientation 1]
  // ---- split-K (dense products of the tensor-core backward): a unit is (K slice, chunk, query group)
  int kslices;               // >= 1
  int kb_per_slice;          // k-blocks per slice (== kblocks when kslices == 1)
  long long emit_slice_stride;  // floats between the partial outputs of consecutive slices
  // ---- MODE_EMIT_COEF: backward coefficients as the bf16 A operand of the gradient GEMMs
  const float* coef_tab;     // orientation 0: T (rows, coef_ld), value = T[row][label of column]
                             // orientation 1: T^t (classes, coef_ld), value = T^t[label of row][column]
  long long coef_ld;
  int coef_orient;
  float* coef_sums;          // [chunk][epilogue set][row]: sum over the unit's columns of the ROUNDED coefficients
  // ---- MODE_EMIT_GRADT: out[dst(c)][r] = acc(r, c) - sub[c] * gt_rows(r, c), r < gt_valid_rows; sub arrives as the
  // per-column additive term (s_sqnorm), dst(c) as the column label (labels; NULL: dst(c) = c)
  const __nv_bfloat16* gt_rows;  // (ceil(n_support / 64), n_query, 64): element (r, c) at [c / 64][r][c % 64], or NULL
  int gt_valid_rows;
};

// (the emit kind is a template parameter: one kernel with a runtime switch and logf inlined 64 times was > 64 KB
//  of SASS and ran 4x slower on instruction fetch)

// Class-end stores.  Rare (once per class per row), kept out of line so the unrolled column loop stays compact, and
// fed with scalars only: nothing of this lives in local memory.
// A class that lies completely inside the unit is final: it is stored to the local table (this set's table when
// several epilogue sets run) AND, for bank-sharded predict, to the peer GPUs' tables — the exchange happens here,
// tile by tile, as NVLink peer stores that overlap the MMAs.  Classes cut by a chunk boundary go through `side`:
// side[((chunk * B + row) * 2 + slot) * sets + set], slot 0 = head-cut class, slot 1 = tail-cut class.
__device__ __noinline__ void store_class(const Params* p, int row, int set, int cls, float v) {
  const size_t off = size_t(row) * p->n_classes + cls;
  if (p->sets >= 2) p->lse[set][off] = v;
  else if (p->rows_per_table > 0) p->lse[row / p->rows_per_table][off] = v;
  else
    for (int r = 0; r < p->n_tables; ++r) p->lse[r][off] = v;
}

__device__ __forceinline__ int4 lds_int4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

__device__ __noinline__ void flush_class(const Params* p, uint32_t unit_addr, int row, int set, int cls, float m,
                                         float l) {
  if (row < 0) return;
  const float v = (m + lg2_approx(l)) * kLn2;
  const int4 u = lds_int4(unit_addr);  // cf, cl, cuts, chunk
  if (cls == u.x && (u.z & 1)) p->side[((size_t(u.w) * p->n_query + row) * 2) * p->sets + set] = v;
  else if (cls == u.y && (u.z & 2)) p->side[((size_t(u.w) * p->n_query + row) * 2 + 1) * p->sets + set] = v;
  else store_class(p, row, set, cls, v);
}

struct Flusher {
  const Params* p;
  uint32_t unit_addr;  // shared-memory address of this unit's UnitInfo
  int row;             // query row, -1 when it is beyond the batch
  int set;
  __device__ __forceinline__ void operator()(int cls, float m, float l) const {
    flush_class(p, unit_addr, row, set, cls, m, l);
  }
  // single-row class that lies inside the unit: its class log-sum-exp is the score itself
  __device__ __forceinline__ void single(int cls, float v) const {
    if (row >= 0) store_class(p, row, set, cls, v);
  }
  // This row of the ONE table every final class value of this thread goes to, or NULL (peer-GPU routes, row beyond
  // the batch): lets the one-row-per-class path store inline.
  __device__ __forceinline__ float* local_row_table() const {
    if (row < 0) return nullptr;
    if (p->sets >= 2) return p->lse[set] + size_t(row) * p->n_classes;
    if (p->rows_per_table == 0 && p->n_tables == 1) return p->lse[0] + size_t(row) * p->n_classes;
    return nullptr;
  }
  // the class of column label `cls` continues into the next chunk (its sum is not final here)
  __device__ __forceinline__ bool tail_cut_class(int cls) const {
    const int4 u = lds_int4(unit_addr);
    return (u.z & 2) && cls == u.y;
  }
};

// One 32-column chunk of the accumulator for one query row.
//
// The epilogue is instruction-issue bound for short GEMMs (ncu at d = 512: ~11 thread instructions per score, issue
// slots 40 % busy with the fixed-latency `wait` stall on top, MUFU 60 %, tensor pipe 58 %; neither more epilogue
// warps nor a cheaper square root moved it), so the common path is written for few instructions per score:
//   * the running maximum m is LAZY.  exp2(score - m) only needs m close to the true maximum, not equal to it
//     (scores above m give terms > 1, which fp32 holds up to 2^127), so a chunk is summed against the current m
//     right away and m is raised — with the exact chunk maximum, and the open sum rescaled — only when the chunk
//     sum says a score exceeded m by more than ~15 in the exp2 domain (or m is still -inf: the first chunk).
//     This removes the per-chunk min/max pass and the sqrt -> compare -> exp2 chain every chunk used to start with;
//   * sqrt(|d2|) instead of sqrt(max(d2, 0)): the absolute value is a free operand modifier of MUFU.SQRT, and a
//     slightly negative d2 (rounding of near-duplicates) is as close to zero either way;
//   * the per-column additive terms are read as float4.
// Chunks in which a class ends (warp-uniform, 1 in ~40 at 1280 rows per class) take the exact-maximum path.
#ifndef NW_EPI_RSQRT
#define NW_EPI_RSQRT 0
#endif
#ifndef NW_EPI_DEBUG
#define NW_EPI_DEBUG 0  // 2 / 3: developer builds that isolate the TMEM loads / the math of the QUAD epilogue
#endif
#ifndef NW_EPI_POLY_EVERY
#define NW_EPI_POLY_EVERY 0  // k > 0: every k-th exp2 of the QUAD (d <= 512) fast path runs on the FMA pipe
#endif
constexpr float kRaiseMax = 1048576.0f;  // chunk sum above which m is raised (some term exceeded 2^15)

// 2^x on the FMA pipe (the FlashAttention-4 trick for MUFU-bound softmax): round-to-nearest range reduction by the
// 1.5 * 2^23 magic add, degree-4 polynomial on [-0.5, 0.5] (relative error 7e-6), exponent patched in with integer
// arithmetic.  x is clamped to [-126, 126]: -inf / NaN (padding, m still -inf) give ~1e-38, +inf gives 2^126, which
// the caller's "chunk sum too large" test still catches.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fminf(fmaxf(x, -126.0f), 126.0f);
  const float r = x + 12582912.0f;
  const float f = x - (r - 12582912.0f);
  float p = fmaf(0.009666374f, f, 0.05583834f);
  p = fmaf(p, f, 0.24022349f);
  p = fmaf(p, f, 0.69313673f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

template <int EPI, int POLY = 0>
__device__ __forceinline__ void epilogue_chunk(float (&acc)[32], const float* __restrict__ cadd,
                                               const int* __restrict__ lab, uint32_t emask, float qn,
                                               float scale2, float& m, float& l, const Flusher& flush) {
  const float4* __restrict__ cadd4 = reinterpret_cast<const float4*>(cadd);
  if (emask == 0u) {  // warp-uniform fast path: no class ends inside this chunk
    float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // four independent partial sums: no 32-long FADD chain
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
      const float4 c = cadd4[i4];
      const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i4 * 4 + k;
        if (EPI == NW_EPI_EUCLID) {
#if NW_EPI_RSQRT  // A/B build only: sqrt(x) as x * rsqrt(x)
          const float t = fmaxf(fabsf(fmaf(-2.0f, acc[i], qn + cc[k])), 1e-30f);
          float r;
          asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
          acc[i] = t * r;
#else
          acc[i] = sqrt_approx(fabsf(fmaf(-2.0f, acc[i], qn + cc[k])));  // distance (kept for a possible redo)
#endif
          if (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == POLY - 1) part[k] += exp2_poly(fmaf(acc[i], -kLog2e, -m));
          else part[k] += ex2_approx(fmaf(acc[i], -kLog2e, -m));
        } else {
          acc[i] = fmaf(acc[i], scale2, cc[k]);  // score * log2(e)  (or -inf on padding columns)
          if (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == POLY - 1) part[k] += exp2_poly(acc[i] - m);
          else part[k] += ex2_approx(acc[i] - m);
        }
      }
    }
    float sum = (part[0] + part[1]) + (part[2] + part[3]);
    if (!(sum < kRaiseMax)) {  // rare (also inf / NaN while m is -inf): raise m to the exact maximum and redo
      float mx;
      if (EPI == NW_EPI_EUCLID) {
        float dmin = acc[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) dmin = fminf(dmin, acc[i]);
        mx = -dmin * kLog2e;
      } else {
        mx = acc[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, acc[i]);
      }
      if (mx > m) {
        l *= ex2_approx(m - mx);
        m = mx;
      }
      float p2[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (EPI == NW_EPI_EUCLID) p2[i & 3] += ex2_approx(fmaf(acc[i], -kLog2e, -m));
        else p2[i & 3] += ex2_approx(acc[i] - m);
      }
      sum = (p2[0] + p2[1]) + (p2[2] + p2[3]);
    }
    l += sum;
    return;
  }
  // A class ends inside this chunk: exact chunk maximum first, then the exponent arguments of all 32 columns in one
  // branch-free pass (the MUFU latencies overlap), then groups of four columns: exp2 of the group, and either one
  // tree add (no class end in the group) or column by column with the class-end stores.  (The first version did
  // sqrt -> exp2 -> add -> branch per column, a ~55-cycle dependent chain 32 times over: banks with ~100 rows per
  // class, where every fourth chunk comes here, ran 1.6x slower than banks with 1280.)
  float mx;
  if (EPI == NW_EPI_EUCLID) {
    float dmin = __int_as_float(0x7f800000);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      acc[i] = sqrt_approx(fabsf(fmaf(-2.0f, acc[i], qn + cadd[i])));  // distance
      dmin = fminf(dmin, acc[i]);
    }
    mx = -dmin * kLog2e;
  } else {
    mx = __int_as_float(0xff800000);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      acc[i] = fmaf(acc[i], scale2, cadd[i]);  // score * log2(e)  (or -inf on padding columns)
      mx = fmaxf(mx, acc[i]);
    }
  }
  if (mx > m) {  // new running maximum: rescale the open class sum
    l *= ex2_approx(m - mx);
    m = mx;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {  // score * log2(e) - m
    if (EPI == NW_EPI_EUCLID) acc[i] = fmaf(acc[i], -kLog2e, -m);
    else acc[i] = acc[i] - m;
  }
  // cluster / random mode banks hold ONE support per class: every column is a class end, and a call per column
  // (256 per tile and thread) made the config-4 predict kernel 121 us for 9 us of MMAs; those values are stored inline
  float* const row_tab = (emask & (emask >> 1)) != 0u ? flush.local_row_table() : nullptr;
#pragma unroll
  for (int i4 = 0; i4 < 32; i4 += 4) {
    float e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) e[k] = ex2_approx(acc[i4 + k]);
    if (((emask >> i4) & 0xfu) == 0u) {  // warp-uniform
      l += (e[0] + e[1]) + (e[2] + e[3]);
      continue;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i4 + k;
      l += e[k];
      if (emask & (1u << i)) {  // warp-uniform: column i is the last row of its class (in this unit)
        if (i > 0 && (emask & (1u << (i - 1))) && !flush.tail_cut_class(lab[i])) {
          // the previous column closed its class too, so this class has ONE row here (cluster / random mode
          // banks: one support per class): its log-sum-exp is simply its score; stored inline, no call
          if (row_tab != nullptr) row_tab[lab[i]] = (acc[i] + m) * kLn2;
          else flush.single(lab[i], (acc[i] + m) * kLn2);
        } else {
          flush(lab[i], m, l);
        }
        l = 0.0f;
      }
    }
  }
}

template <int EPI, bool INFLUENCE>
__device__ __forceinline__ void emit_chunk(float (&acc)[32], const float* __restrict__ cadd,
                                           const int* __restrict__ lab, float qn, float scale, float z,
                                           float pq, int qy, float (*stage)[33], int lane, float* __restrict__ out,
                                           long long ld, int row0, int n_rows, int col0, int n_valid, bool vec) {
  // phase 1: scores of this thread's query row -> staging tile
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float sc;
    if (EPI == NW_EPI_EUCLID) sc = -sqrt_approx(fmaxf(fmaf(-2.0f, acc[i], qn + cadd[i]), 0.0f));
    else sc = acc[i] * scale;
    stage[lane][i] = sc;
  }
  __syncwarp();
  // phase 2: transposed read-out.  The influence transform runs here, in a rolled loop (4 inline copies instead
  // of 64: the unrolled form was > 64 KB of SASS and instruction-fetch bound); lane rr holds row rr's parameters.
  const int c4 = (lane & 7) * 4;
#pragma unroll 2
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + (lane >> 3);
    float v[4] = {stage[rr][c4], stage[rr][c4 + 1], stage[rr][c4 + 2], stage[rr][c4 + 3]};
    if (INFLUENCE) {
      const float rz = __shfl_sync(0xffffffffu, z, rr);
      const float rp = __shfl_sync(0xffffffffu, pq, rr);
      const int ry = __shfl_sync(0xffffffffu, qy, rr);
      // Branch-free fast path for all four values first (one epilogue warp per scheduler: the per-element branch of
      // influence_one serialised the four dependency chains, 900 cycles per iteration), exact formula afterwards
      // only where |x| is not small.
      float w[4], den[4], x[4];
      bool slow = false;
      // (skipping the division for supports of another class, where the ratio is simply -w, was measured: the
      //  per-lane select costs more than the MUFU.RCP it saves — 2.33 -> 2.72 ms at config 5 from features)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float ind = lab[c4 + k] == ry ? 1.0f : 0.0f;
        w[k] = __expf(v[k] - rz);
        den[k] = rp - ind * w[k];
        x[k] = __fdividef(w[k] * (ind - rp), den[k]);
        v[k] = x[k] * (1.0f + x[k] * (-0.5f + x[k] * (0.33333334f - 0.25f * x[k])));
        slow |= !(fabsf(x[k]) < 0.015625f);
      }
      if (slow) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!(fabsf(x[k]) < 0.015625f)) v[k] = influence_exact(rp, w[k], den[k]);
      }
    }
    if (row0 + rr < n_rows) {
      float* dst = out + (long long)(row0 + rr) * ld + col0 + c4;
      if (vec && c4 + 4 <= n_valid) {
        __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c4 + k < n_valid) dst[k] = v[k];
      }
    }
  }
  __syncwarp();
}

// MODE_EMIT_COEF — the recompute step of the tensor-core backward (closed form: SURVEY.md B.2).  With
//   p(b, j) = exp(score(b, j) - lse_b)            the softmax weight of support j for query b
//   T(b, c) = g(b, c) / (P(b, c) + 1e-12) - sum_c' g(b, c') P(b, c') / (P(b, c') + 1e-12)
// the gradient of the loss with respect to score(b, j) is p(b, j) T(b, y_j), and
//   EUCLID (score = -|q - s|):  w = p T / |q - s|   ->  grad_q = W S - rowsum(W) q,   grad_s = W^t Q - colsum(W) s
//   LINEAR (score = scale q.s): w = p T scale       ->  grad_q = W S,                  grad_s = W^t Q
// (a zero distance contributes nothing, as torch.cdist's backward).  One thread owns one row of the tile and 32
// consecutive columns; w is rounded to bf16 and stored k-block-major over the COLUMNS,
// out[col / 64][row][col % 64], which is exactly the A-operand layout of this kernel: the gradient GEMMs consume it
// with K = the column axis.  Orientation 0: rows = queries, columns = supports (W, for grad_q).  Orientation 1:
// rows = supports, columns = queries (W^t, for grad_s): the per-column log-sum-exp arrives through the label slots
// of the tile metadata (float bits) and the table is indexed [label of the row][column].
// Eight columns of one row: coefficient values -> four packed bf16 pairs; `sum` accumulates the ROUNDED values.
// Branch-free (the first version tested validity and table index per element: 48 BSSY/BSYNC pairs in the SASS
// serialised the rsqrt -> exp2 chains of the 32 columns, and the emit ran at 65 % of the MMA-bound rate).
template <int EPI, bool FULL>
__device__ __forceinline__ uint4 coef_group(const float (&acc)[32], const float* __restrict__ cadd, int c0,
                                            const float (&tb)[8], const float (&zl)[8], float qn, float scale,
                                            int n_valid, float& sum) {
  uint32_t pk[4];
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    float w[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 2 * k2 + h, c = c0 + k;
      float v;
      if (EPI == NW_EPI_EUCLID) {
        const float nn = qn + cadd[c];
        const float d2 = fmaf(-2.0f, acc[c], nn);
        float inv;  // 1 / distance: one MUFU op serves the distance and the division (NaN / inf for d2 <= 0: discarded)
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(d2));
        const float pr = ex2_approx(fmaf(d2 * inv, -kLog2e, zl[k]));
        // a squared distance below the rounding noise of its own terms (~1e-6 (|q|^2 + |s|^2), and |s| ~ |q| for
        // such a pair) is a coincident pair: no gradient
        v = d2 > 2e-6f * qn ? (pr * inv) * tb[k] : 0.0f;
      } else {
        v = ex2_approx(fmaf(acc[c], scale * kLog2e, zl[k])) * (tb[k] * scale);
      }
      w[h] = (FULL || c < n_valid) ? v : 0.0f;
    }
    const __nv_bfloat162 hh = __floats2bfloat162_rn(w[0], w[1]);
    pk[k2] = *reinterpret_cast<const uint32_t*>(&hh);
    sum += __uint_as_float(pk[k2] << 16) + __uint_as_float(pk[k2] & 0xffff0000u);
  }
  return make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

template <int EPI>
__device__ __forceinline__ float coef_chunk(float (&acc)[32], const float* __restrict__ cadd,
                                            const int* __restrict__ lab, float qn, float scale, float row_z,
                                            int row_lab, const Params& p, int row, int col0, int n_valid) {
  if (row < 0) return 0.0f;
  uint4* dst4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.emit_out) +
                                         ((long long)(col0 >> 6) * p.n_query + row) * 64 + (col0 & 63));
  if (n_valid <= 0) {
    // beyond the last column: the output's K axis is padded to a multiple of 64 and the padding must read as zero
    // in the gradient GEMM (the caller need not clear 10 GB for the sake of at most 63 columns)
    if (col0 < ((p.n_support + 63) & ~63)) {
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) dst4[g8] = make_uint4(0u, 0u, 0u, 0u);
    }
    return 0.0f;
  }
  float sum = 0.0f;  // of the values as the gradient GEMM will see them (bf16-rounded)
  const bool by_col = p.coef_orient == 0;  // the table value changes with the column's label
  const float* __restrict__ trow =
      by_col ? p.coef_tab + (long long)row * p.coef_ld : p.coef_tab + (long long)row_lab * p.coef_ld + col0;
  const float zl_row = -row_z * kLog2e;
  // all three tests are warp-uniform (tile metadata and column counts are the same for every row of the warp)
  if (by_col && n_valid >= 32 && lab[0] == lab[31]) {
    // class-sorted bank: the 32 columns share one class (39 of 40 chunks at 1280 rows per class): one table value
    const float t_u = __ldg(trow + lab[0]);
    float tb[8], zl[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      tb[k] = t_u;
      zl[k] = zl_row;
    }
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) dst4[g8] = coef_group<EPI, true>(acc, cadd, g8 * 8, tb, zl, qn, scale, 32, sum);
  } else if (by_col) {
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      float tb[8], zl[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {  // labels of columns beyond the bank are -1: clamped, the value is masked
        tb[k] = __ldg(trow + max(lab[g8 * 8 + k], 0));
        zl[k] = zl_row;
      }
      dst4[g8] = coef_group<EPI, false>(acc, cadd, g8 * 8, tb, zl, qn, scale, n_valid, sum);
    }
  } else if (n_valid >= 32 && (p.coef_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.coef_tab) & 15) == 0) {
    // rows = supports, a full chunk of query columns: table values and column terms as 16-byte loads
    const float4* __restrict__ t4 = reinterpret_cast<const float4*>(trow);
    const int4* __restrict__ l4 = reinterpret_cast<const int4*>(lab);
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      const float4 ta = __ldg(t4 + 2 * g8), tc = __ldg(t4 + 2 * g8 + 1);
      const int4 la = l4[2 * g8], lc = l4[2 * g8 + 1];
      const float tb[8] = {ta.x, ta.y, ta.z, ta.w, tc.x, tc.y, tc.z, tc.w};
      const float zl[8] = {-__int_as_float(la.x) * kLog2e, -__int_as_float(la.y) * kLog2e,
                           -__int_as_float(la.z) * kLog2e, -__int_as_float(la.w) * kLog2e,
                           -__int_as_float(lc.x) * kLog2e, -__int_as_float(lc.y) * kLog2e,
                           -__int_as_float(lc.z) * kLog2e, -__int_as_float(lc.w) * kLog2e};
      dst4[g8] = coef_group<EPI, true>(acc, cadd, g8 * 8, tb, zl, qn, scale, 32, sum);
    }
  } else {
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      float tb[8], zl[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {  // the column's log-sum-exp rides in the label slot (float bits)
        tb[k] = __ldg(trow + min(g8 * 8 + k, n_valid - 1));
        zl[k] = -__int_as_float(lab[g8 * 8 + k]) * kLog2e;
      }
      dst4[g8] = coef_group<EPI, false>(acc, cadd, g8 * 8, tb, zl, qn, scale, n_valid, sum);
    }
  }
  return sum;
}

// MODE_EMIT_GRADT — the grad_s products of the tensor-core backward with their last step fused in.  Operand roles are
// swapped against the forward (rows = features of Q^t, columns = supports, K = queries) so that the big operand, W^t,
// streams once as the "bank" while Q^t stays in L2; the result is stored transposed, out[dst(c)][r] — 32 consecutive
// rows per warp, one coalesced 128-byte store per column — minus colsum(W)[c] times the stored support row, read from
// the transposed bank (64 contiguous bytes per thread and chunk), and in the row order of the caller's support
// tensor.  (Unfused: products 16.8 ms with W^t re-read 8x from HBM, + a 26 GB finishing pass of 4.3 ms, at config 3.)
__device__ __forceinline__ void gradt_chunk(const float (&acc)[32], const float* __restrict__ sub,
                                            const int* __restrict__ dst_row, const Params& p, int row, int col0,
                                            int n_valid) {
  if (row < 0 || row >= p.gt_valid_rows || n_valid <= 0) return;
  const uint4* __restrict__ r4 = reinterpret_cast<const uint4*>(
      p.gt_rows + ((long long)(col0 >> 6) * p.n_query + row) * 64 + (col0 & 63));
  float* __restrict__ out_col = p.emit_out + row;
#pragma unroll
  for (int v = 0; v < 4; ++v) {  // eight columns at a time: one 16-byte load of the stored rows, eight stores
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (p.gt_rows != nullptr) u = __ldg(r4 + v);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = v * 8 + k;
      const float corr = (k & 1) ? __uint_as_float(w[k >> 1] & 0xffff0000u) : __uint_as_float(w[k >> 1] << 16);
      if (i < n_valid) {
        const long long dst = dst_row[i] >= 0 ? dst_row[i] : col0 + i;
        out_col[dst * p.emit_ld] = fmaf(-sub[i], corr, acc[i]);
      }
    }
  }
}

// MODE_EMIT_BLOCKBEST: best score of this thread's query row over one 32-column chunk (padding columns excluded).
template <int EPI>
__device__ __forceinline__ float chunk_best(const float (&acc)[32], const float* __restrict__ cadd, float qn,
                                            float scale, int n_valid) {
  float best = __int_as_float(0xff800000);
  if (EPI == NW_EPI_EUCLID) {
    float dmin = __int_as_float(0x7f800000);  // padding columns carry cadd = +inf and never win
#pragma unroll
    for (int i = 0; i < 32; ++i) dmin = fminf(dmin, fmaf(-2.0f, acc[i], qn + cadd[i]));
    best = -sqrt_approx(fmaxf(dmin, 0.0f));
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < n_valid) best = fmaxf(best, acc[i] * scale);
  }
  return n_valid > 0 ? best : __int_as_float(0xff800000);
}

template <int EPI, int NCTA, int MODE, bool QUAD = false>
__global__ void __launch_bounds__(QUAD ? QUAD_THREADS : MAX_THREADS, 1)
nw_forward_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_s,
                  const __grid_constant__ Params p) {
  using C = Cfg<NCTA, MODE, QUAD>;
  constexpr int STAGES = C::STAGES;
  constexpr int STAGE_BYTES = C::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;  // 1024-B aligned (swizzle-128B atoms); same offset in both CTAs of a pair
  SmemTail<MODE, QUAD>* tail = reinterpret_cast<SmemTail<MODE, QUAD>*>(smem + size_t(STAGES) * STAGE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int worker = blockIdx.x / NCTA;
  const int n_workers = gridDim.x / NCTA;
  const int units_per_slice = p.chunks * p.q_groups;
  const int n_units = units_per_slice * p.kslices;  // kslices == 1 except for the split-K dense products
  // epilogue sets: they share every tile by columns (a compile-time constant where the epilogue is the bottleneck)
  const int n_sets = QUAD ? QUAD_SETS : (MODE == MODE_CLASS_LSE ? p.sets : 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_s);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&tail->full[i]), 1);
      mbar_init(smem_u32(&tail->empty[i]), 1);
    }
    for (int i = 0; i < META_SLOTS; ++i) {
      mbar_init(smem_u32(&tail->mfull[i]), 1);
      mbar_init(smem_u32(&tail->mempty[i]), 4 * n_sets);  // the epilogue warps of THIS CTA
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(smem_u32(&tail->tfull[i]), 1);
      mbar_init(smem_u32(&tail->tempty[i]), NCTA * 4 * n_sets);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (NCTA == 2) {
      tmem_alloc_pair(smem_u32(&tail->tmem_base), TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(smem_u32(&tail->tmem_base), TMEM_COLS);
      tmem_relinquish();
    }
  }
  if (EPI != NW_EPI_EUCLID && MODE == MODE_CLASS_LSE && p.meta_bulk) {
    // LINEAR scores have no additive term: the slots' cadd stay zero, only the labels are copied per tile
    for (int i = threadIdx.x; i < META_SLOTS * BN; i += blockDim.x) tail->mslot[i / BN].cadd[i % BN] = 0.0f;
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;
  // tile whose metadata the producer delivers: completely inside the bank, look-ahead label included
  auto bulk_tile = [&](int t) { return MODE == MODE_CLASS_LSE && p.meta_bulk != 0 && t * BN + BN + 4 <= p.n_support; };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      // Queries are re-read for every support tile; a support tile is read by all query groups of the chunk
      // within a few microseconds.  Both are marked evict_last: same-box A/B on the sustained bench gave
      // supports evict_last +1.3 %, evict_normal 0, evict_first -4 % (fewer re-reads of the bank from HBM).
      // When every worker streams its own support tiles (one query group) there is nothing to keep:
      // evict_last there cost 3-9 % at B <= 256.
      const uint64_t pol_q = l2_policy_evict_last();
      const uint64_t pol_s = p.s_keep ? l2_policy_evict_last() : l2_policy_evict_normal();
      uint32_t it = 0, mc = 0;  // ring stages / metadata slots filled so far
      if (p.stagger_ns > 0 && worker < n_units) {
        const unsigned long long wait_ns = (unsigned long long)(worker % p.q_groups) * p.stagger_ns;
        const unsigned long long t_start = globaltimer_ns();
        while (globaltimer_ns() - t_start < wait_ns) {
        }
      }
      for (int u = worker; u < n_units; u += n_workers) {
        const int ks = u / units_per_slice;
        const int ur = u - ks * units_per_slice;
        const int g = ur / p.q_groups;
        const int qg = ur - g * p.q_groups;
        const int kb0 = ks * p.kb_per_slice;
        const int kb1 = min(kb0 + p.kb_per_slice, p.kblocks);
        const int t0 = g * p.tiles_per_chunk;
        const int t1 = min(t0 + p.tiles_per_chunk, p.s_tiles);
        const int q_row0 = (qg * NCTA + int(cta_rank)) * BM;
        // CTAs that work on this chunk at the same time (its units that fall into this worker's wave)
        int gate_ctas = 0;
        int* gate = nullptr;
        if (p.tile_gate != nullptr) {
          const int wave = u / n_workers;
          const int lo = max(g * p.q_groups, wave * n_workers);
          const int hi = min(g * p.q_groups + p.q_groups, (wave + 1) * n_workers);
          gate_ctas = (hi - lo) * NCTA;
          gate = p.tile_gate + 2 * g + (g * p.q_groups < wave * n_workers ? 1 : 0);
        }
        for (int t = t0; t < t1; ++t) {
          const int s_row0 = t * BN + int(cta_rank) * C::B_ROWS;
          if (gate_ctas > NCTA) {
            // Start the tile together with the other CTAs that read it: a support line is only cheap to share while
            // the requests for it are in flight together (lines prefetched into L2 3 us early were gone when the
            // ring asked for them, profiles/r2_k1_ring_depth_vs_dram.txt), and a 6-stage ring lets the workers drift
            // by a couple of microseconds.  Measured at config 3 on a box that re-read the bank 3.5x: 18.4 -> 6.3 GB
            // from HBM per launch (the bank once + the chunks cut by a wave boundary), 1.17 -> 1.25 GHz under the
            // power cap, 263 k -> 278.6 k queries/s.  The producer runs up to a ring ahead of the MMAs, so the wait
            // costs the tensor pipe ~1 %.  Bounded: a straggler (a CTA that is not resident yet) delays the others
            // by 6 us once, then the gate is ignored for the rest of the unit; nothing can hang.
            atomicAdd(gate, 1);
            const int target = gate_ctas * (t - t0 + 1);
            const long long c_start = clock64();
            while (ld_acquire_gpu_s32(gate) < target) {
              if (clock64() - c_start > 8000) {
                gate_ctas = 0;
                break;
              }
            }
          }
          if (bulk_tile(t)) {
            const uint32_t ms = mc % META_SLOTS;
            mbar_wait(smem_u32(&tail->mempty[ms]), ((mc / META_SLOTS) & 1u) ^ 1u);
            const uint32_t bar = smem_u32(&tail->mfull[ms]);
            mbar_arrive_expect_tx(bar, (EPI == NW_EPI_EUCLID ? META_CADD_BYTES : 0u) + META_LAB_BYTES);
            if (EPI == NW_EPI_EUCLID) bulk_load_1d(smem_u32(tail->mslot[ms].cadd), p.s_sqnorm + t * BN, META_CADD_BYTES, bar);
            bulk_load_1d(smem_u32(tail->mslot[ms].lab), p.labels + t * BN, META_LAB_BYTES, bar);
            ++mc;
          }
          const int nkb = kb1 - kb0;
          const int rot = (qg * p.krot_step) % nkb;  // (the sum over k does not care where it starts)
          const int pf = p.l2_prefetch > 0 ? (qg == 0 ? p.l2_prefetch : 0) : -p.l2_prefetch;
          for (int i = 0; i < nkb; ++i, ++it) {
            const int kb = kb0 + (i + rot < nkb ? i + rot : i + rot - nkb);
            if (pf > 0) {  // the box this worker's ring will ask for `pf` k-blocks from now
              const int ahead = i + pf;
              const int tp = t + ahead / nkb;
              if (tp < t1) tma_prefetch_l2_3d(&map_s, 0, tp * BN + int(cta_rank) * C::B_ROWS, kb0 + ahead % nkb);
            }
            const uint32_t s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u);
            const uint32_t a_dst = smem_u32(smem + size_t(s) * STAGE_BYTES);
            if (NCTA == 2) {
              // both CTAs' bytes are accounted on the LEADER's full barrier (the MMA issuer waits there)
              const uint32_t bar = mapa_u32(smem_u32(&tail->full[s]), 0);
              if (leader) mbar_arrive_expect_tx(smem_u32(&tail->full[s]), 2 * STAGE_BYTES);
              tma_load_3d_pair(a_dst, &map_q, bar, 0, q_row0, kb, pol_q);
              tma_load_3d_pair(a_dst + A_BYTES, &map_s, bar, 0, s_row0, kb, pol_s);
            } else {
              const uint32_t bar = smem_u32(&tail->full[s]);
              mbar_arrive_expect_tx(bar, STAGE_BYTES);
              tma_load_3d(a_dst, &map_q, bar, 0, q_row0, kb, pol_q);
              tma_load_3d(a_dst + A_BYTES, &map_s, bar, 0, s_row0, kb, pol_s);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) ======================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * NCTA, BN);
      uint32_t it = 0, tc = 0;
      for (int u = worker; u < n_units; u += n_workers) {
        const int ks = u / units_per_slice;
        const int g = (u - ks * units_per_slice) / p.q_groups;
        const int kb0 = ks * p.kb_per_slice;
        const int kb1 = min(kb0 + p.kb_per_slice, p.kblocks);
        const int t0 = g * p.tiles_per_chunk;
        const int t1 = min(t0 + p.tiles_per_chunk, p.s_tiles);
        for (int t = t0; t < t1; ++t, ++tc) {
          const uint32_t as = tc & 1u;
          const uint32_t aph = (tc >> 1) & 1u;
          mbar_wait(smem_u32(&tail->tempty[as]), aph ^ 1u);  // epilogues have drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const uint32_t s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(smem_u32(&tail->full[s]), ph);  // TMA bytes (of both CTAs) have landed
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + size_t(s) * STAGE_BYTES);
            const uint64_t adesc = umma_desc_k128(a_addr);
            const uint64_t bdesc = umma_desc_k128(a_addr + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 32 B (= 16 bf16) inside the 128-B swizzled row: +2 in the 16-B address field
              const uint32_t accumulate = (kb > kb0 || k > 0) ? 1u : 0u;
              if (NCTA == 2) umma_bf16_ss_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, accumulate);
              else umma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, accumulate);
            }
            // frees the smem stage (in both CTAs) once these MMAs retire
            if (NCTA == 2) umma_commit_pair(smem_u32(&tail->empty[s]), 3);
            else umma_commit(smem_u32(&tail->empty[s]));
          }
          // accumulator complete -> epilogue warps (of both CTAs)
          if (NCTA == 2) umma_commit_pair(smem_u32(&tail->tfull[as]), 3);
          else umma_commit(smem_u32(&tail->tfull[as]));
        }
      }
    }
    __syncwarp();
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4 * n_sets) {
    // ===================================== epilogue ==========================================
    constexpr int epi_threads = 128;         // threads of ONE epilogue set (each set stages its own metadata)
    const int ew = (warp - EPI_WARP0) & 3;   // == warp % 4: TMEM lane quarter this warp may read
    const int eg = (warp - EPI_WARP0) >> 2;  // epilogue set: which pairs of column chunks this warp handles
    const int et = (threadIdx.x - EPI_WARP0 * 32) & 127;
    const float scale2 = p.scale_log2;
    const float neg_inf = __int_as_float(0xff800000);
    const float pos_inf = __int_as_float(0x7f800000);
    uint32_t tc = 0, uc = 0, mc = 0;  // tiles / units / producer-delivered metadata slots this CTA has worked on
    // Effective SM clock of this launch, measured in the kernel: SM cycles (clock64) over wall time (globaltimer)
    // across the whole epilogue role of one thread per CTA, accumulated per CTA so that a series of launches
    // yields the time-weighted mean.  nvidia-smi / NVML clocks are instantaneous samples; this is the integral.
    const bool probe = p.clock_probe != nullptr && warp == EPI_WARP0 && lane == 0;
    long long probe_c0 = 0;
    unsigned long long probe_t0 = 0;
    if (probe) {
      probe_c0 = clock64();
      probe_t0 = globaltimer_ns();
    }
    for (int u = worker; u < n_units; u += n_workers, ++uc) {
      const int ks = u / units_per_slice;
      const int ur = u - ks * units_per_slice;
      const int g = ur / p.q_groups;
      const int qg = ur - g * p.q_groups;
      float* const emit_out = p.emit_out + (long long)ks * p.emit_slice_stride;  // this K slice's partial output
      const int t0 = g * p.tiles_per_chunk;
      const int t1 = min(t0 + p.tiles_per_chunk, p.s_tiles);
      const int n0 = t0 * BN;
      const int n1 = min(t1 * BN, p.n_support);
      const int row = (qg * NCTA + int(cta_rank)) * BM + ew * 32 + lane;

      const bool row_valid = row < p.n_query;
      Flusher flush;
      flush.p = &p;
      flush.unit_addr = smem_u32(&tail->unit[eg][uc & 1u]);
      flush.row = row_valid ? row : -1;
      flush.set = eg;
      if (MODE == MODE_CLASS_LSE && et == 0) {
        // visible to the set's other threads after the first tile's named barrier (no class ends before that)
        UnitInfo ui;
        ui.cf = __ldg(p.labels + n0);
        ui.cl = __ldg(p.labels + n1 - 1);
        ui.cuts = ((n0 > 0 && __ldg(p.labels + n0 - 1) == ui.cf) ? 1 : 0) |
                  ((n1 < p.n_support && __ldg(p.labels + n1) == ui.cl) ? 2 : 0);
        ui.g = g;
        tail->unit[eg][uc & 1u] = ui;
      }
      // UnitInfo visible to the whole set before its first class end; also keeps any thread from running two units
      // ahead of another (the double buffer above, and the reuse of meta[eg][] by edge tiles, rely on that)
      if (MODE == MODE_CLASS_LSE) named_bar_sync(1 + eg, epi_threads);
      const float qn = (EPI == NW_EPI_EUCLID && row_valid) ? __ldg(p.q_sqnorm + row) : 0.0f;
      float e_z = 0.0f, e_p = 0.0f;
      int e_qy = -1;
      if (MODE == MODE_EMIT_INFLUENCE && row_valid) {
        e_z = __ldg(p.row_lse + row);
        e_p = __ldg(p.p_query + row);
        e_qy = __ldg(p.qlabel + row);
      }
      if (MODE == MODE_EMIT_COEF && row_valid) {
        if (p.coef_orient == 0) e_z = __ldg(p.row_lse + row);
        else e_qy = __ldg(p.qlabel + row);
      }

      // column metadata of a tile: additive term (|s|^2, or 0; +inf / -inf on padding columns) and labels
      float pre_cadd[2];
      int pre_lab[2];
      int pre_lab_next = -1;
      auto load_meta = [&](int t) {
        const int j0 = t * BN;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int i = et + r * epi_threads;
          if (i < BN) {
            const int j = j0 + i;
            if (EPI == NW_EPI_EUCLID) pre_cadd[r] = j < n1 ? __ldg(p.s_sqnorm + j) : pos_inf;
            else if (MODE == MODE_EMIT_GRADT) pre_cadd[r] = (p.s_sqnorm != nullptr && j < n1) ? __ldg(p.s_sqnorm + j) : 0.0f;
            else pre_cadd[r] = j < n1 ? 0.0f : neg_inf;
            pre_lab[r] = (p.labels != nullptr && j < p.n_support) ? __ldg(p.labels + j) : -1;
          }
        }
        if (et == 0) pre_lab_next = (p.labels != nullptr && (j0 + BN) < p.n_support) ? __ldg(p.labels + j0 + BN) : -1;
      };
      if (!bulk_tile(t0)) load_meta(t0);

      float m = neg_inf, l = 0.0f;
      float coef_rsum = 0.0f;  // MODE_EMIT_COEF: this row's coefficient sum over the unit's columns (this set's share)
      int open_cls = -1;  // two sets: class whose partial this set currently holds (warp-uniform)
      for (int t = t0; t < t1; ++t, ++tc) {
        const uint32_t as = tc & 1u;
        const uint32_t aph = (tc >> 1) & 1u;
        const int j0 = t * BN;
        const bool bulk = bulk_tile(t);  // CTA-uniform
        const uint32_t ms = mc % META_SLOTS;
        TileMeta& meta = bulk ? tail->mslot[ms] : tail->meta[eg][as];
        if (bulk) {
          mbar_wait(smem_u32(&tail->mfull[ms]), (mc / META_SLOTS) & 1u);
        } else {
          // publish this tile's column metadata (fetched into registers one tile ahead, so the global-load
          // latency is hidden behind the previous tile's epilogue math)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int i = et + r * epi_threads;
            if (i < BN) {
              meta.cadd[i] = pre_cadd[r];
              meta.lab[i] = pre_lab[r];
            }
          }
          if (et == 0) meta.lab[BN] = pre_lab_next;
#if NW_EPI_DEBUG != 4 && NW_EPI_DEBUG != 5  // developer builds 4 / 5: no per-tile barrier / no barrier and no metadata prefetch
          named_bar_sync(1 + eg, epi_threads);
#endif
        }
#if NW_EPI_DEBUG != 5
        if (t + 1 < t1 && !bulk_tile(t + 1)) load_meta(t + 1);
#endif

        mbar_wait(smem_u32(&tail->tfull[as]), aph);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + as * BN + (uint32_t(ew * 32) << 16);
        // Two 32-column chunks per iteration, NOT unrolled further: the epilogue body is ~1.5k instructions and
        // a fully unrolled tile (8 chunks x 2 paths) overflows the instruction cache (stall_no_inst in ncu).
        // (Software-pipelining the TMEM loads one chunk ahead was measured: no gain, +40 registers.)
        // chunk pairs are dealt round-robin to the epilogue sets (class-LSE: 1 / 2 / 4 of them; emit modes: 2)
        const int c_step = 2 * n_sets;
#pragma unroll 1
        for (int c = 2 * eg; c < BN / 32; c += c_step) {
          if (MODE != MODE_CLASS_LSE) {
            constexpr bool INFL = MODE == MODE_EMIT_INFLUENCE;
            if (QUAD) {
              // 16 epilogue warps, 96 registers: one 32-column chunk in registers at a time
              const int row0 = (qg * NCTA + int(cta_rank)) * BM + ew * 32;
              float(*stg)[33] = tail->stage[warp - EPI_WARP0];
#pragma unroll 1
              for (int cc = c; cc < c + 2; ++cc) {
                float acc[32];
                tmem_ld_32x32(t_addr + cc * 32, acc);
                tmem_ld_wait();
                if (MODE == MODE_EMIT_GRADT) {
                  gradt_chunk(acc, meta.cadd + cc * 32, meta.lab + cc * 32, p, row_valid ? row : -1, j0 + cc * 32,
                              n1 - (j0 + cc * 32));
                  continue;
                }
                if (MODE == MODE_EMIT_COEF) {
                  coef_rsum += coef_chunk<EPI>(acc, meta.cadd + cc * 32, meta.lab + cc * 32, qn, p.scale_log2 * kLn2, e_z,
                                               e_qy, p, row_valid ? row : -1, j0 + cc * 32, n1 - (j0 + cc * 32));
                  continue;
                }
                emit_chunk<EPI, INFL>(acc, meta.cadd + cc * 32, meta.lab + cc * 32, qn, p.scale_log2 * kLn2, e_z, e_p, e_qy,
                                      stg, lane, emit_out, p.emit_ld, row0, p.n_query, j0 + cc * 32, n1 - (j0 + cc * 32),
                                      p.emit_vec != 0);
              }
              continue;
            }
            float acc0[32], acc1[32];
            tmem_ld_32x32(t_addr + c * 32, acc0);
            tmem_ld_32x32(t_addr + (c + 1) * 32, acc1);
            tmem_ld_wait();
            if (MODE == MODE_EMIT_GRADT) {
              const int rr = row_valid ? row : -1;
              gradt_chunk(acc0, meta.cadd + c * 32, meta.lab + c * 32, p, rr, j0 + c * 32, n1 - (j0 + c * 32));
              gradt_chunk(acc1, meta.cadd + (c + 1) * 32, meta.lab + (c + 1) * 32, p, rr, j0 + (c + 1) * 32,
                          n1 - (j0 + (c + 1) * 32));
              continue;
            }
            if (MODE == MODE_EMIT_COEF) {
              const int rr = row_valid ? row : -1;
              coef_rsum += coef_chunk<EPI>(acc0, meta.cadd + c * 32, meta.lab + c * 32, qn, p.scale_log2 * kLn2, e_z, e_qy,
                                           p, rr, j0 + c * 32, n1 - (j0 + c * 32));
              coef_rsum += coef_chunk<EPI>(acc1, meta.cadd + (c + 1) * 32, meta.lab + (c + 1) * 32, qn,
                                           p.scale_log2 * kLn2, e_z, e_qy, p, rr, j0 + (c + 1) * 32,
                                           n1 - (j0 + (c + 1) * 32));
              continue;
            }
            if (MODE == MODE_EMIT_BLOCKBEST) {
              const float b0 = chunk_best<EPI>(acc0, meta.cadd + c * 32, qn, p.scale_log2 * kLn2, n1 - (j0 + c * 32));
              const float b1 = chunk_best<EPI>(acc1, meta.cadd + (c + 1) * 32, qn, p.scale_log2 * kLn2,
                                               n1 - (j0 + (c + 1) * 32));
              // out[block][query]: 32 consecutive query rows per warp -> one coalesced 128-byte store
              if (row_valid) emit_out[(long long)(t * (BN / 64) + (c >> 1)) * p.emit_ld + row] = fmaxf(b0, b1);
              continue;
            }
            const int row0 = (qg * NCTA + int(cta_rank)) * BM + ew * 32;
            float(*stg)[33] = tail->stage[warp - EPI_WARP0];
            emit_chunk<EPI, INFL>(acc0, meta.cadd + c * 32, meta.lab + c * 32, qn, p.scale_log2 * kLn2, e_z, e_p, e_qy,
                                  stg, lane, emit_out, p.emit_ld, row0, p.n_query, j0 + c * 32, n1 - (j0 + c * 32),
                                  p.emit_vec != 0);
            emit_chunk<EPI, INFL>(acc1, meta.cadd + (c + 1) * 32, meta.lab + (c + 1) * 32, qn, p.scale_log2 * kLn2, e_z,
                                  e_p, e_qy, stg, lane, emit_out, p.emit_ld, row0, p.n_query, j0 + (c + 1) * 32,
                                  n1 - (j0 + (c + 1) * 32), p.emit_vec != 0);
            continue;
          }
          if (p.debug_skip_epilogue == 1) continue;
          if (n_sets >= 2) {
            // this set skipped the columns in between: if they ended the class it was accumulating, close its
            // partial now (the other sets close their own; the tables are combined after the kernel)
            const int first = meta.lab[c * 32];
            if (open_cls >= 0 && first != open_cls) {
              flush(open_cls, m, l);
              l = 0.0f;
            }
          }
          // (Deriving "no class ends in this chunk" from three warp-uniform label reads per pair — a class is one
          //  contiguous run of rows — instead of the per-lane comparison + vote below was measured 2 % SLOWER: the
          //  vote overlaps the TMEM load, the uniform test sits in front of it.)
          const int i0 = c * 32 + lane, i1 = i0 + 32;
          const int ja = j0 + i0, jb = j0 + i1;
          uint32_t em1;
          if (QUAD) {
            // one 32-column chunk in registers at a time: 640 threads leave ~100 registers per thread
            float acc[32];
#if NW_EPI_DEBUG == 3  // developer build: the math on register data, no TMEM loads
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = float(i) + qn;
#else
            tmem_ld_32x32(t_addr + c * 32, acc);
#endif
            const uint32_t em0 =
                __ballot_sync(0xffffffffu, ja < n1 && (ja == n1 - 1 || meta.lab[i0] != meta.lab[i0 + 1]));
#if NW_EPI_DEBUG != 3
            tmem_ld_wait();
#endif
#if NW_EPI_DEBUG == 2  // developer build: the TMEM loads only, no math
            l += acc[0] + acc[31];
#else
            epilogue_chunk<EPI, NW_EPI_POLY_EVERY>(acc, meta.cadd + c * 32, meta.lab + c * 32, em0, qn, scale2, m, l, flush);
#endif
#if NW_EPI_DEBUG == 3
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = float(i) + l;
#else
            tmem_ld_32x32(t_addr + (c + 1) * 32, acc);
#endif
            em1 = __ballot_sync(0xffffffffu, jb < n1 && (jb == n1 - 1 || meta.lab[i1] != meta.lab[i1 + 1]));
#if NW_EPI_DEBUG != 3
            tmem_ld_wait();
#endif
#if NW_EPI_DEBUG == 2
            l += acc[0] + acc[31];
#else
            epilogue_chunk<EPI, NW_EPI_POLY_EVERY>(acc, meta.cadd + (c + 1) * 32, meta.lab + (c + 1) * 32, em1, qn, scale2, m, l,
                                                   flush);
#endif
          } else {
            float acc0[32], acc1[32];
            tmem_ld_32x32(t_addr + c * 32, acc0);
            tmem_ld_32x32(t_addr + (c + 1) * 32, acc1);
            // class-end masks of both chunks while the TMEM loads are in flight
            const uint32_t em0 =
                __ballot_sync(0xffffffffu, ja < n1 && (ja == n1 - 1 || meta.lab[i0] != meta.lab[i0 + 1]));
            em1 = __ballot_sync(0xffffffffu, jb < n1 && (jb == n1 - 1 || meta.lab[i1] != meta.lab[i1 + 1]));
            tmem_ld_wait();
            epilogue_chunk<EPI>(acc0, meta.cadd + c * 32, meta.lab + c * 32, em0, qn, scale2, m, l, flush);
            epilogue_chunk<EPI>(acc1, meta.cadd + (c + 1) * 32, meta.lab + (c + 1) * 32, em1, qn, scale2, m, l, flush);
          }
          if (n_sets >= 2) {
            // class left open after this pair (none if its last valid column closed a class or is padding)
            const int j_last = j0 + (c + 2) * 32 - 1;
            open_cls = (j_last < n1 && !(em1 >> 31)) ? meta.lab[(c + 2) * 32 - 1] : -1;
          }
        }
        // all TMEM reads of this accumulator are complete -> hand it back to the (leader's) MMA warp, and the
        // metadata slot to the producer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tail->tempty[as]), 0));
          else mbar_arrive(smem_u32(&tail->tempty[as]));
          if (bulk) mbar_arrive(smem_u32(&tail->mempty[ms]));
        }
        if (bulk) ++mc;
      }
      // two sets: the unit's last columns may belong to the other set; close what this set still holds
      if (MODE == MODE_CLASS_LSE && n_sets >= 2 && open_cls >= 0) flush(open_cls, m, l);
      if (MODE == MODE_EMIT_COEF && row_valid && p.coef_sums != nullptr)
        p.coef_sums[(long long)(g * n_sets + eg) * p.n_query + row] = coef_rsum;
    }
    if (probe) {
      atomicAdd(p.clock_probe + 2 * blockIdx.x, (unsigned long long)(clock64() - probe_c0));
      atomicAdd(p.clock_probe + 2 * blockIdx.x + 1, globaltimer_ns() - probe_t0);
    }
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();  // no CTA may exit while its peer can still signal into its shared memory
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// a[0, n) = v; b[0, nb) = v except b[z0, z1), which is cleared to integer zero (the tile gates inside `side`)
__global__ void fill_kernel(float* __restrict__ a, long long n, float* __restrict__ b, long long nb, float v,
                            long long z0, long long z1) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = v;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += stride)
    b[i] = (i >= z0 && i < z1) ? 0.0f : v;
}

__device__ __forceinline__ float logaddexp_f(float a, float b) {
  const float mx = fmaxf(a, b);
  if (mx == __int_as_float(0xff800000)) return mx;
  return mx + log1pf(expf(-fabsf(a - b)));
}

// Apply the chunk-boundary partials (fixed combination order => bitwise reproducible).
// A class that is cut by a chunk boundary receives ALL of its mass through `side` (never a direct store): every chunk
// contributes two entries, (first class, head partial) and (last class, tail partial), each split over the epilogue
// sets.  Along the chunks the entries of one class are adjacent (a class is one contiguous run of support rows), so
// the merge is a SEGMENTED SCAN: one warp per query row, one entry per lane (the lane folds the sets' partials of
// its entry), five shuffle steps per 32 entries, every segment's last lane stores its class, and a segment that
// continues into the next 32 entries is carried.  (History: one thread per row paid one memory round trip per chunk
// — 170 us of a 950 us small-batch step; a warp replaying the entries serially through log1pf(expf()) was 22 us of a
// 50 us config-1 head call, 9.8 us with the sets pre-combined.)
struct TableList {
  float* t[NW_MAX_PEERS];
  int n;
  int rows_per_table;
};

constexpr int MERGE_ROWS_PER_BLOCK = 4;

__global__ void __launch_bounds__(MERGE_ROWS_PER_BLOCK * 32) merge_side_kernel(
    const TableList tables, const float* __restrict__ side, const int32_t* __restrict__ labels, int n_query,
    int n_support, int n_classes, int chunks, int tiles_per_chunk, int s_tiles, int sets) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * MERGE_ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (b >= n_query) return;  // uniform per warp
  const size_t row_off = size_t(b) * n_classes;
  const float neg_inf = __int_as_float(0xff800000);
  // log(exp(a) + exp(b)), -inf = "nothing"; fast intrinsics (absolute error < 5e-7), symmetric in its arguments
  auto lae = [&](float x, float y) {
    if (x == neg_inf) return y;
    if (y == neg_inf) return x;
    return fmaxf(x, y) + __logf(1.0f + __expf(-fabsf(x - y)));
  };
  auto put = [&](int cls, float v) {  // by the calling lane
    if (tables.rows_per_table > 0) tables.t[b / tables.rows_per_table][row_off + cls] = v;
    else
      for (int r = 0; r < tables.n; ++r) tables.t[r][row_off + cls] = v;
  };
  const int n_entries = 2 * chunks;
  int carry_cls = -1;
  float carry_val = neg_inf;
  for (int base = 0; base < n_entries; base += 32) {
    const int e = base + lane;
    int cls = -2;  // lanes past the last entry: a class nobody has, nothing to add
    float val = neg_inf;
    if (e < n_entries) {
      const int g = e >> 1, slot = e & 1;
      const int t0 = g * tiles_per_chunk;
      const int t1 = min(t0 + tiles_per_chunk, s_tiles);
      cls = slot ? __ldg(labels + min(size_t(t1) * BN, size_t(n_support)) - 1) : __ldg(labels + size_t(t0) * BN);
      const float* sr = side + ((size_t(g) * n_query + b) * 2 + slot) * sets;
#pragma unroll
      for (int i = 0; i < QUAD_SETS; ++i)
        if (i < sets) val = lae(val, sr[i]);
    }
    if (carry_cls >= 0) {  // warp-uniform: the previous 32 entries ended inside a class
      if (__shfl_sync(0xffffffffu, cls, 0) == carry_cls) {
        if (lane == 0) val = lae(carry_val, val);
      } else if (lane == 0 && carry_val != neg_inf) {
        put(carry_cls, carry_val);
      }
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int oc = __shfl_up_sync(0xffffffffu, cls, off);
      const float ov = __shfl_up_sync(0xffffffffu, val, off);
      if (lane >= off && oc == cls) val = lae(ov, val);
    }
    const int next_cls = __shfl_down_sync(0xffffffffu, cls, 1);
    if (lane < 31 && next_cls != cls && cls >= 0 && val != neg_inf) put(cls, val);  // last entry of its class
    carry_cls = __shfl_sync(0xffffffffu, cls, 31);
    carry_val = __shfl_sync(0xffffffffu, val, 31);
  }
  if (lane == 0 && carry_cls >= 0 && carry_val != neg_inf) put(carry_cls, carry_val);
}

// logp[b,c] = log(exp(L[b,c] - logsumexp_c L[b,:]) + 1e-12)   (reference nwhead/nw.py:285-289)
__global__ void __launch_bounds__(256) logp_kernel(const float* __restrict__ class_lse, int n_classes,
                                                   float* __restrict__ logp) {
  __shared__ float red[8];
  __shared__ float bcast;
  const float* row = class_lse + size_t(blockIdx.x) * n_classes;
  float* out = logp + size_t(blockIdx.x) * n_classes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = __int_as_float(0xff800000);
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) mx = fmaxf(mx, row[c]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = red[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, red[i]);
    bcast = v;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.0f;
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) sum += expf(row[c] - mx);
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.0f;
    for (int i = 0; i < 8; ++i) v += red[i];
    bcast = mx + logf(v);
  }
  __syncthreads();
  const float lse = bcast;
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) out[c] = logf(expf(row[c] - lse) + 1e-12f);
}

// The same for rows of up to 32 * PER_LANE classes: one WARP per row, the row held in registers — one read, two
// shuffle reductions, one write, no block barrier (the block-per-row kernel above makes three dependent passes and
// two __syncthreads per row: 19 us for 4096 x 1000, a quarter of a config-4 predict call).
template <int PER_LANE>
__global__ void __launch_bounds__(256) logp_rows_kernel(const float* __restrict__ class_lse, int n_query,
                                                        int n_classes, float* __restrict__ logp) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= n_query) return;  // uniform per warp
  const float* row = class_lse + size_t(b) * n_classes;
  float* out = logp + size_t(b) * n_classes;
  const float neg_inf = __int_as_float(0xff800000);
  float v[PER_LANE];
  float mx = neg_inf;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int c = i * 32 + lane;
    v[i] = c < n_classes ? row[c] : neg_inf;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) sum += expf(v[i] - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int c = i * 32 + lane;
    if (c < n_classes) out[c] = logf(expf(v[i] - lse) + 1e-12f);
  }
}

// row_lse[b] = logsumexp_c L[b,:]; p_query[b] = exp(L[b, qlabel[b]] - row_lse[b])   (inputs of the influence emit pass)
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ class_lse, int n_classes,
                                                        const int32_t* __restrict__ qlabel,
                                                        float* __restrict__ row_lse, float* __restrict__ p_query) {
  __shared__ float red[8];
  __shared__ float bcast;
  const float* row = class_lse + size_t(blockIdx.x) * n_classes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = __int_as_float(0xff800000);
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) mx = fmaxf(mx, row[c]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = red[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, red[i]);
    bcast = v;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.0f;
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) sum += expf(row[c] - mx);
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.0f;
    for (int i = 0; i < 8; ++i) v += red[i];
    const float z = mx + logf(v);
    row_lse[blockIdx.x] = z;
    if (p_query) p_query[blockIdx.x] = expf(row[qlabel[blockIdx.x]] - z);
  }
}

__global__ void lse_merge_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    a[i] = logaddexp_f(a[i], b[i]);
}

// a <- log(exp(a) + sum_t exp(b[t * n + .])), t < n_extra: the tables of the further epilogue sets, in set order
__global__ void lse_merge_sets_kernel(float* __restrict__ a, const float* __restrict__ b, int n_extra, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = a[i];
    for (int t = 0; t < n_extra; ++t) v = logaddexp_f(v, b[(long long)t * n + i]);
    a[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// k-block-major bf16 operand [cols/64][rows][64] seen as a 3-D tensor {64, rows, cols/64}; box = one k-block of
// box_rows rows (box_rows x 128 B, contiguous in HBM), 128-B swizzle, zero fill for rows past the end.
static int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  NW_REQUIRE(fn != nullptr, NW_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[3] = {BK, rows, cols / BK};
  cuuint64_t gstride[2] = {BK * 2, rows * BK * 2};
  cuuint32_t box[3] = {BK, box_rows, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NW_REQUIRE(r == CUDA_SUCCESS, NW_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return NW_OK;
}

}  // namespace k1
}  // namespace nw

using namespace nw;

// diagnostics: per-CTA (SM cycles, nanoseconds) accumulators written by the forward kernels while set
static unsigned long long* g_clock_probe = nullptr;
static long long g_clock_probe_ctas = 0;

extern "C" int nw_forward_set_clock_probe(void* buf, int64_t capacity_ctas) {
  NW_REQUIRE(buf == nullptr || capacity_ctas > 0, NW_ERR_INVALID, "capacity_ctas must be positive");
  g_clock_probe = static_cast<unsigned long long*>(buf);
  g_clock_probe_ctas = buf ? capacity_ctas : 0;
  return NW_OK;
}

static bool force_single_cta() {
  const char* e = getenv("NW_B200_FORCE_1CTA");
  return e && e[0] == '1';
}

// L2 policy of the support tiles: evict_last once `min_groups` query groups share every tile (default 2:
// same-box A/B at B=512 gave +4 % over evict_normal, 0 within noise at B=768/1024;
// NW_B200_S_KEEP_MIN_GROUPS overrides, for A/B runs on one box).
static int support_keep_policy(int q_groups) {
  static const int min_groups = [] {
    const char* e = getenv("NW_B200_S_KEEP_MIN_GROUPS");
    return e && *e ? atoi(e) : 2;
  }();
  return q_groups >= min_groups ? 1 : 0;
}

// Epilogue warp sets the fused forward wants for rows of `row_elems` bf16: the epilogue (2 MUFU ops per score) is
// the bottleneck of short GEMMs.  NW_B200_EPI_SETS=1|2|4 overrides (same-box A/B runs).
extern "C" int nw_forward_epilogue_sets(int row_elems) {
  static const int forced = [] {
    const char* e = getenv("NW_B200_EPI_SETS");
    const int v = e && *e ? atoi(e) : 0;
    return (v == 1 || v == 2 || v == 4) ? v : 0;
  }();
  if (forced) return forced;
  const int kblocks = row_elems / k1::BK;
  return kblocks <= 8 ? k1::QUAD_SETS : (kblocks <= 16 ? 2 : 1);
}

extern "C" int nw_forward_plan(int n_query, int64_t n_support, nw_forward_plan_t* plan) {
  NW_REQUIRE(plan != nullptr, NW_ERR_INVALID, "plan_out is NULL");
  NW_REQUIRE(n_query > 0 && n_support > 0, NW_ERR_INVALID, "n_query and n_support must be positive");
  NW_REQUIRE(n_support < (int64_t(1) << 31) - 512, NW_ERR_UNSUPPORTED, "n_support must be < 2^31 - 512");
  const int sms = sm_count();
  NW_REQUIRE(sms > 0, NW_ERR_CUDA, "no CUDA device");
  // more than one 128-row query tile: pair CTAs (cta_group::2, 256 x 256 tiles)
  const int ncta = (n_query > k1::BM && !force_single_cta() && sms >= 2) ? 2 : 1;
  const int workers = sms / ncta;
  const int q_groups = ceil_div(n_query, k1::BM * ncta);
  const int s_tiles = int(ceil_div_ll(n_support, k1::BN));
  // choose the number of chunks so that (chunks * q_groups) fills whole waves of persistent workers
  long long best_cost = -1;
  int best_tpc = s_tiles;
  for (int w = 1; w <= 64; ++w) {
    long long want = ceil_div_ll((long long)w * workers, q_groups);
    int G = int(want < s_tiles ? want : s_tiles);
    if (G < 1) G = 1;
    const int tpc = ceil_div(s_tiles, G);
    const int Ge = ceil_div(s_tiles, tpc);
    const long long waves = ceil_div_ll((long long)Ge * q_groups, workers);
    const long long cost = waves * (tpc + 1);  // +1 tile-equivalent of per-unit pipeline fill / flush
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_tpc = tpc;
    }
    if (G == s_tiles) break;
  }
  {
    // developer override for scheduling experiments (NW_B200_FORCE_CHUNKS=<chunks>)
    static const int forced = [] {
      const char* e = getenv("NW_B200_FORCE_CHUNKS");
      return e && *e ? atoi(e) : 0;
    }();
    if (forced > 0) best_tpc = ceil_div(s_tiles, forced < s_tiles ? forced : s_tiles);
  }
  plan->q_tiles = q_groups;
  plan->s_tiles = s_tiles;
  plan->tiles_per_chunk = best_tpc;
  plan->chunks = ceil_div(s_tiles, best_tpc);
  const long long units = (long long)plan->chunks * q_groups;
  plan->grid = int(units < workers ? units : workers) * ncta;
  plan->cta_pair = ncta == 2 ? 1 : 0;
  // chunk-boundary partials: [chunk][query][2 slots][up to 4 epilogue sets]  (+ room for the further sets' tables
  // is added by the caller, see nw_forward_epilogue_sets / forward_impl)
  plan->side_elems = int64_t(plan->chunks) * n_query * 2 * k1::QUAD_SETS + 2 * plan->chunks;  // + the tile gates
  return NW_OK;
}

namespace nw {
namespace k1 {
template <int EPI, int NCTA, int MODE = MODE_CLASS_LSE, bool QUAD = false>
static int launch_forward(const CUtensorMap& map_q, const CUtensorMap& map_s, const Params& p, int grid,
                          cudaStream_t stream) {
  static bool attr_set[64] = {false};  // function attributes are per device
  auto kern = nw_forward_kernel<EPI, NCTA, MODE, QUAD>;
  constexpr size_t smem = smem_bytes<NCTA, MODE, QUAD>();
  static_assert(smem <= 232448, "shared memory budget exceeded");
  int dev = 0;
  NW_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    NW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((EPI_WARP0 + 4 * p.sets) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NW_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, map_q, map_s, p));
  return NW_OK;
}
}  // namespace k1
}  // namespace nw

// k-block rotation between the query groups that share a support tile (Params::krot_step).  Developer experiment,
// OFF unless NW_B200_KROT=1: it was meant to keep 16 CTA pairs from asking the L2 for the same line at the same time
// (were the bank re-reads from HBM duplicate fetches of lines still in flight?).  Measured on a box that re-reads
// 13.8 GB per launch: 12.4 GB with the rotation, but 1190 instead of 1204 MHz under the power cap and 269 k instead of
// 272 k queries/s — simultaneous requests are not the cause, and the L2 serves them more cheaply than spread ones.
static int k_rotation_step(int q_groups, int kblocks) {
  static const int enabled = [] {
    const char* e = getenv("NW_B200_KROT");
    return e && *e ? atoi(e) : 0;
  }();
  if (!enabled || q_groups < 2 || kblocks < 2) return 0;
  const int step = kblocks / q_groups;
  return step > 0 ? step : 1;
}

// Developer probe (NW_B200_TRACE_HOST=1): report host-side stages of the forward call that take longer than 0.3 ms.
struct HostStageTimer {
  bool on;
  timespec t0;
  HostStageTimer() {
    static const bool enabled = [] {
      const char* e = getenv("NW_B200_TRACE_HOST");
      return e && e[0] == '1';
    }();
    on = enabled;
    if (on) clock_gettime(CLOCK_MONOTONIC, &t0);
  }
  void lap(const char* what) {
    if (!on) return;
    timespec t1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
    if (ms > 0.3) fprintf(stderr, "[nw host trace] %s took %.2f ms\n", what, ms);
    t0 = t1;
  }
};

static int forward_impl(int epilogue, float scale, const void* q_bf16, const float* q_sqnorm, int n_query,
                        const void* bank_bf16, const float* s_sqnorm, const int32_t* labels, int64_t n_support,
                        int row_elems, int n_classes, float* const* tables, int n_tables, int rows_per_table,
                        bool fill_local, float* side, int64_t side_elems, cudaStream_t stream) {
  NW_REQUIRE(epilogue == NW_EPI_EUCLID || epilogue == NW_EPI_LINEAR, NW_ERR_INVALID, "unknown epilogue %d", epilogue);
  NW_REQUIRE(q_bf16 && bank_bf16 && labels && tables && side, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_tables >= 1 && n_tables <= k1::NW_MAX_PEERS, NW_ERR_INVALID, "n_tables must be in [1, %d]",
             k1::NW_MAX_PEERS);
  for (int r = 0; r < n_tables; ++r) NW_REQUIRE(tables[r] != nullptr, NW_ERR_INVALID, "NULL class-LSE table %d", r);
  NW_REQUIRE(rows_per_table >= 0 && (rows_per_table == 0 || (long long)rows_per_table * n_tables >= n_query),
             NW_ERR_INVALID, "rows_per_table * n_tables must cover n_query");
  NW_REQUIRE(epilogue != NW_EPI_EUCLID || (q_sqnorm && s_sqnorm), NW_ERR_INVALID,
             "the euclidean epilogue needs q_sqnorm and s_sqnorm");
  NW_REQUIRE(row_elems > 0 && row_elems % k1::BK == 0, NW_ERR_INVALID, "row_elems must be a positive multiple of 64");
  NW_REQUIRE(n_classes > 0, NW_ERR_INVALID, "n_classes must be positive");
  NW_REQUIRE((reinterpret_cast<uintptr_t>(q_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(bank_bf16) & 15) == 0,
             NW_ERR_INVALID, "bf16 operands must be 16-byte aligned");
  HostStageTimer host_timer;
  int rc = nw_device_check();
  if (rc != NW_OK) return rc;
  host_timer.lap("nw_device_check");
  nw_forward_plan_t plan;
  rc = nw_forward_plan(n_query, n_support, &plan);
  if (rc != NW_OK) return rc;
  NW_REQUIRE(side_elems >= plan.side_elems, NW_ERR_WORKSPACE, "side scratch too small: %lld < %lld floats",
             (long long)side_elems, (long long)plan.side_elems);
  const int ncta = plan.cta_pair ? 2 : 1;
  // Short GEMMs (d <= 1024) are epilogue-bound with one epilogue warp per scheduler: run two independent epilogue
  // sets, each with its own class-LSE table (the second one lives behind the chunk partials in `side`).
  const long long table_elems = (long long)n_query * n_classes;
  int sets = nw_forward_epilogue_sets(row_elems);
  // the further sets need their own tables behind the chunk partials; the multi-table (peer GPU) routes use one set
  while (sets > 1 && (n_tables != 1 || side_elems < plan.side_elems + (sets - 1) * table_elems)) sets >>= 1;
  float* table1 = side + plan.side_elems;

  CUtensorMap map_q, map_s;
  rc = k1::make_map(&map_q, q_bf16, uint64_t(n_query), uint64_t(row_elems), k1::BM);
  if (rc != NW_OK) return rc;
  rc = k1::make_map(&map_s, bank_bf16, uint64_t(n_support), uint64_t(row_elems), k1::BN / ncta);
  if (rc != NW_OK) return rc;
  host_timer.lap("plan + tensor maps");

  k1::fill_kernel<<<sm_count() * 4, 256, 0, stream>>>(tables[0], fill_local ? table_elems : 0, side,
                                                      (long long)plan.side_elems + (sets - 1) * table_elems,
                                                      -INFINITY, (long long)plan.side_elems - 2 * plan.chunks,
                                                      (long long)plan.side_elems);
  NW_CUDA_OK(cudaGetLastError());
  host_timer.lap("fill launch");

  k1::Params p;
  p.q_sqnorm = q_sqnorm;
  p.s_sqnorm = s_sqnorm;
  p.labels = labels;
  k1::TableList tl;
  for (int r = 0; r < k1::NW_MAX_PEERS; ++r) p.lse[r] = tl.t[r] = (r < n_tables ? tables[r] : nullptr);
  for (int t = 1; t < sets; ++t) p.lse[t] = table1 + (long long)(t - 1) * table_elems;
  p.n_tables = tl.n = n_tables;
  p.rows_per_table = tl.rows_per_table = rows_per_table;
  p.sets = sets;
  p.side = side;
  p.n_query = n_query;
  p.n_support = int(n_support);
  p.n_classes = n_classes;
  p.kblocks = row_elems / k1::BK;
  p.q_groups = plan.q_tiles;
  p.s_keep = support_keep_policy(plan.q_tiles);
  p.s_tiles = plan.s_tiles;
  p.chunks = plan.chunks;
  p.tiles_per_chunk = plan.tiles_per_chunk;
  p.scale_log2 = scale * kLog2e;
  p.emit_out = nullptr;
  p.emit_ld = 0;
  p.emit_kind = 0;
  p.emit_vec = 0;
  p.kslices = 1;
  p.kb_per_slice = p.kblocks;
  p.emit_slice_stride = 0;
  p.coef_tab = nullptr;
  p.coef_ld = 0;
  p.coef_orient = 0;
  p.coef_sums = nullptr;
  p.clock_probe = (g_clock_probe && plan.grid <= g_clock_probe_ctas) ? g_clock_probe : nullptr;
  {
    static const int skip = [] {
      const char* e = getenv("NW_B200_DEBUG_SKIP_EPI");
      return e && *e ? atoi(e) : 0;
    }();
    p.debug_skip_epilogue = skip;
    static const int persist_mb = [] {  // developer probe: L2 set-aside for evict_last (persisting) lines
      const char* e = getenv("NW_B200_PERSIST_L2_MB");
      const int mb = e && *e ? atoi(e) : -1;
      if (mb >= 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, size_t(mb) << 20);
      return mb;
    }();
    (void)persist_mb;
    static const int stagger = [] {
      const char* e = getenv("NW_B200_STAGGER_NS");
      return e && *e ? atoi(e) : 0;
    }();
    p.stagger_ns = stagger;
    p.krot_step = k_rotation_step(plan.q_tiles, p.kblocks);
    static const int l2_prefetch = [] {
      const char* e = getenv("NW_B200_L2_PREFETCH");
      return e && *e ? atoi(e) : 0;
    }();
    p.l2_prefetch = l2_prefetch;
    static const int tile_gate = [] {  // NW_B200_TILE_GATE=0 switches the per-tile producer gate off (A/B runs)
      const char* e = getenv("NW_B200_TILE_GATE");
      return e && *e ? atoi(e) : 1;
    }();
    // Only where the workers of a chunk stay in step by themselves — long, MMA-bound tiles (d >= 1024) and long units:
    // with 10 000 classes at d = 512 (uneven class-end work per tile) the gate cost 6 %, with 17-tile units
    // (N = 160 000) 3 %; at d = 1024 it gains 2 % even off the power cap (profiles/r2_k1_ring_depth_vs_dram.txt).
    const bool gate_shape = p.kblocks >= 16 && plan.tiles_per_chunk >= 32 && plan.q_tiles >= 2 &&
                            plan.q_tiles <= plan.grid / ncta;
    p.tile_gate = (tile_gate != 0 && (gate_shape || tile_gate == 2))
                      ? reinterpret_cast<int*>(side + plan.side_elems) - 2 * plan.chunks
                      : nullptr;
    if (plan.q_tiles < 2 || plan.q_tiles > plan.grid / ncta) p.tile_gate = nullptr;  // (2 = forced: A/B runs)
    static const int bulk_meta = [] {  // developer knob: 0 = every tile's metadata staged by the epilogue sets
      const char* e = getenv("NW_B200_META_BULK");
      return e && *e ? atoi(e) : 1;
    }();
    // bulk copies need 16-byte aligned sources (tile offsets are multiples of 1 KB)
    const bool aligned = (reinterpret_cast<uintptr_t>(labels) & 15u) == 0 &&
                         (epilogue != NW_EPI_EUCLID || (reinterpret_cast<uintptr_t>(s_sqnorm) & 15u) == 0);
    p.meta_bulk = (bulk_meta != 0 && aligned) ? 1 : 0;
  }
  p.row_lse = p.p_query = nullptr;
  p.qlabel = nullptr;

  const bool euc = epilogue == NW_EPI_EUCLID;
  if (sets == k1::QUAD_SETS) {
    rc = euc ? (ncta == 2 ? k1::launch_forward<NW_EPI_EUCLID, 2, k1::MODE_CLASS_LSE, true>(map_q, map_s, p, plan.grid, stream)
                          : k1::launch_forward<NW_EPI_EUCLID, 1, k1::MODE_CLASS_LSE, true>(map_q, map_s, p, plan.grid, stream))
             : (ncta == 2 ? k1::launch_forward<NW_EPI_LINEAR, 2, k1::MODE_CLASS_LSE, true>(map_q, map_s, p, plan.grid, stream)
                          : k1::launch_forward<NW_EPI_LINEAR, 1, k1::MODE_CLASS_LSE, true>(map_q, map_s, p, plan.grid, stream));
  } else {
    rc = euc ? (ncta == 2 ? k1::launch_forward<NW_EPI_EUCLID, 2>(map_q, map_s, p, plan.grid, stream)
                          : k1::launch_forward<NW_EPI_EUCLID, 1>(map_q, map_s, p, plan.grid, stream))
             : (ncta == 2 ? k1::launch_forward<NW_EPI_LINEAR, 2>(map_q, map_s, p, plan.grid, stream)
                          : k1::launch_forward<NW_EPI_LINEAR, 1>(map_q, map_s, p, plan.grid, stream));
  }
  if (rc != NW_OK) return rc;
  host_timer.lap("K1 launch");

  if (sets > 1) {  // combine the epilogue sets' tables
    k1::lse_merge_sets_kernel<<<sm_count() * 4, 256, 0, stream>>>(tables[0], table1, sets - 1, table_elems);
    NW_CUDA_OK(cudaGetLastError());
  }
  if (plan.chunks > 1) {
    k1::merge_side_kernel<<<ceil_div(n_query, k1::MERGE_ROWS_PER_BLOCK), k1::MERGE_ROWS_PER_BLOCK * 32, 0, stream>>>(
        tl, side, labels, n_query, int(n_support), n_classes, plan.chunks, plan.tiles_per_chunk, plan.s_tiles, sets);
    NW_CUDA_OK(cudaGetLastError());
  }
  host_timer.lap("merge launches");
  return NW_OK;
}

extern "C" int nw_forward_class_lse(int epilogue, float scale, const void* q_bf16, const float* q_sqnorm,
                                    int n_query, const void* bank_bf16, const float* s_sqnorm,
                                    const int32_t* labels, int64_t n_support, int row_elems, int n_classes,
                                    float* class_lse, float* side, int64_t side_elems, void* stream_) {
  NW_REQUIRE(class_lse != nullptr, NW_ERR_INVALID, "NULL pointer argument");
  float* tables[1] = {class_lse};
  return forward_impl(epilogue, scale, q_bf16, q_sqnorm, n_query, bank_bf16, s_sqnorm, labels, n_support, row_elems,
                      n_classes, tables, 1, 0, /*fill_local=*/true, side, side_elems, static_cast<cudaStream_t>(stream_));
}

extern "C" int nw_forward_class_lse_peers(int epilogue, float scale, const void* q_bf16, const float* q_sqnorm,
                                          int n_query, const void* bank_bf16, const float* s_sqnorm,
                                          const int32_t* labels, int64_t n_support, int row_elems, int n_classes,
                                          float* const* tables_host, int n_tables, int rows_per_table, float* side,
                                          int64_t side_elems, void* stream_) {
  return forward_impl(epilogue, scale, q_bf16, q_sqnorm, n_query, bank_bf16, s_sqnorm, labels, n_support, row_elems,
                      n_classes, tables_host, n_tables, rows_per_table, /*fill_local=*/false, side, side_elems,
                      static_cast<cudaStream_t>(stream_));
}

// Internal emit kinds beyond include/nw_sm100.h's NW_EMIT_*: the backward coefficients (nw_backward_coefficients).
constexpr int EMIT_COEF_INTERNAL = 100;
constexpr int EMIT_GRADT_INTERNAL = 101;

struct EmitExtra {
  int kslices = 1;                  // split-K (NW_EMIT_SCORES only): partial products, one output per slice
  long long slice_stride = 0;       // floats between the slices' outputs
  const float* coef_tab = nullptr;  // EMIT_COEF_INTERNAL
  long long coef_ld = 0;
  int coef_orient = 0;
  float* coef_ws = nullptr;    // [chunks][COEF_SETS][n_rows] partial row sums
  float* coef_sums = nullptr;  // (n_rows) their fixed-order total
  const void* gt_rows = nullptr;  // EMIT_GRADT_INTERNAL
  int gt_valid_rows = 0;
};

constexpr int COEF_SETS = k1::QUAD_SETS;  // epilogue sets of the coefficient emit

// sums[r] = sum_i ws[i][r] in fixed order (deterministic: no atomics across the chunks / epilogue sets)
__global__ void coef_sum_kernel(const float* __restrict__ ws, int parts, int n_rows, float* __restrict__ sums) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  float acc = 0.0f;
  for (int i = 0; i < parts; ++i) acc += ws[(long long)i * n_rows + r];
  sums[r] = acc;
}

static int emit_impl(int epilogue, float scale, const void* q_bf16, const float* q_sqnorm, int n_query,
                     const void* bank_bf16, const float* s_sqnorm, const int32_t* labels, int64_t n_support,
                     int row_elems, int emit_kind, const float* row_lse, const float* p_query, const int32_t* qlabel,
                     float* out, int64_t ld_out, const EmitExtra& ex, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(epilogue == NW_EPI_EUCLID || epilogue == NW_EPI_LINEAR, NW_ERR_INVALID, "unknown epilogue %d", epilogue);
  NW_REQUIRE(emit_kind == NW_EMIT_SCORES || emit_kind == NW_EMIT_INFLUENCE || emit_kind == NW_EMIT_BLOCK_BEST ||
                 emit_kind == EMIT_COEF_INTERNAL || emit_kind == EMIT_GRADT_INTERNAL,
             NW_ERR_INVALID, "unknown emit kind %d", emit_kind);
  NW_REQUIRE(q_bf16 && bank_bf16 && out, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(epilogue != NW_EPI_EUCLID || (q_sqnorm && s_sqnorm), NW_ERR_INVALID,
             "the euclidean epilogue needs q_sqnorm and s_sqnorm");
  NW_REQUIRE(emit_kind != NW_EMIT_INFLUENCE || (labels && row_lse && p_query && qlabel), NW_ERR_INVALID,
             "influence needs labels, row_lse, p_query and qlabel");
  NW_REQUIRE(row_elems > 0 && row_elems % k1::BK == 0, NW_ERR_INVALID, "row_elems must be a positive multiple of 64");
  NW_REQUIRE(emit_kind == EMIT_COEF_INTERNAL || emit_kind == EMIT_GRADT_INTERNAL ||
                 (emit_kind == NW_EMIT_BLOCK_BEST ? ld_out >= n_query : ld_out >= n_support),
             NW_ERR_INVALID, "ld_out must be >= n_support (>= n_query for NW_EMIT_BLOCK_BEST)");
  NW_REQUIRE(ex.kslices >= 1 && (ex.kslices == 1 || emit_kind == NW_EMIT_SCORES), NW_ERR_INVALID,
             "split-K is available for dense products only");
  NW_REQUIRE((reinterpret_cast<uintptr_t>(q_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(bank_bf16) & 15) == 0,
             NW_ERR_INVALID, "bf16 operands must be 16-byte aligned");
  int rc = nw_device_check();
  if (rc != NW_OK) return rc;
  nw_forward_plan_t plan;
  rc = nw_forward_plan(n_query, n_support, &plan);
  if (rc != NW_OK) return rc;
  const int ncta = plan.cta_pair ? 2 : 1;
  CUtensorMap map_q, map_s;
  rc = k1::make_map(&map_q, q_bf16, uint64_t(n_query), uint64_t(row_elems), k1::BM);
  if (rc != NW_OK) return rc;
  rc = k1::make_map(&map_s, bank_bf16, uint64_t(n_support), uint64_t(row_elems), k1::BN / ncta);
  if (rc != NW_OK) return rc;

  k1::Params p = {};
  p.q_sqnorm = q_sqnorm;
  p.s_sqnorm = s_sqnorm;
  p.labels = labels;
  p.n_tables = 0;
  p.n_query = n_query;
  p.n_support = int(n_support);
  p.n_classes = 1;
  p.kblocks = row_elems / k1::BK;
  p.q_groups = plan.q_tiles;
  p.s_keep = support_keep_policy(plan.q_tiles);
  p.s_tiles = plan.s_tiles;
  p.chunks = plan.chunks;
  p.tiles_per_chunk = plan.tiles_per_chunk;
  p.scale_log2 = scale * kLog2e;
  p.sets = 2;
  p.meta_bulk = 0;
  p.emit_out = out;
  p.emit_ld = ld_out;
  p.emit_kind = emit_kind;
  p.emit_vec = (ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
  p.clock_probe = (g_clock_probe && plan.grid <= g_clock_probe_ctas) ? g_clock_probe : nullptr;
  p.row_lse = row_lse;
  p.p_query = p_query;
  p.qlabel = qlabel;
  // split-K: no slice may be empty (its accumulator would be read without a single MMA)
  p.kb_per_slice = ceil_div(p.kblocks, ex.kslices);
  p.kslices = ceil_div(p.kblocks, p.kb_per_slice);
  p.krot_step = k_rotation_step(plan.q_tiles, p.kb_per_slice);
  p.emit_slice_stride = ex.slice_stride;
  p.coef_tab = ex.coef_tab;
  p.coef_ld = ex.coef_ld;
  p.coef_orient = ex.coef_orient;
  p.coef_sums = ex.coef_ws;
  p.gt_rows = static_cast<const __nv_bfloat16*>(ex.gt_rows);
  p.gt_valid_rows = ex.gt_valid_rows;
  int grid = plan.grid;
  if (p.kslices > 1) {  // more units than the plan knew of: size the persistent grid for all of them
    const long long units = (long long)plan.chunks * plan.q_tiles * p.kslices;
    const int workers = sm_count() / ncta;
    grid = int(units < workers ? units : workers) * ncta;
  }
  const bool euc = epilogue == NW_EPI_EUCLID;
#define NW_LAUNCH_EMIT(MODE_)                                                                                    \
  rc = euc ? (ncta == 2 ? k1::launch_forward<NW_EPI_EUCLID, 2, MODE_>(map_q, map_s, p, grid, stream)           \
                        : k1::launch_forward<NW_EPI_EUCLID, 1, MODE_>(map_q, map_s, p, grid, stream))          \
           : (ncta == 2 ? k1::launch_forward<NW_EPI_LINEAR, 2, MODE_>(map_q, map_s, p, grid, stream)           \
                        : k1::launch_forward<NW_EPI_LINEAR, 1, MODE_>(map_q, map_s, p, grid, stream))
  static const int emit_sets = [] {  // developer knob: 2 = the influence emit runs the 8-warp epilogue as well
    const char* e = getenv("NW_B200_EMIT_SETS");
    return e && *e ? atoi(e) : 4;
  }();
  // The influence transform (exp, reciprocal, polynomial per pair) is latency-bound with two warps per scheduler:
  // four epilogue sets, at the price of two ring stages (config 5 from features, bf16: 2.29 -> 1.70 ms).  Dense
  // scores keep two sets: four gave +4.5 % with one bf16 pass and -7 % with the three passes of bf16x3.
#define NW_LAUNCH_EMIT4(MODE_)                                                                                         \
  rc = euc ? (ncta == 2 ? k1::launch_forward<NW_EPI_EUCLID, 2, MODE_, true>(map_q, map_s, p, grid, stream)           \
                        : k1::launch_forward<NW_EPI_EUCLID, 1, MODE_, true>(map_q, map_s, p, grid, stream))          \
           : (ncta == 2 ? k1::launch_forward<NW_EPI_LINEAR, 2, MODE_, true>(map_q, map_s, p, grid, stream)           \
                        : k1::launch_forward<NW_EPI_LINEAR, 1, MODE_, true>(map_q, map_s, p, grid, stream))
  if (emit_kind == NW_EMIT_SCORES) NW_LAUNCH_EMIT(k1::MODE_EMIT_SCORES);
  else if (emit_kind == EMIT_GRADT_INTERNAL) {
    static const int gradt_sets = [] {  // developer knob (same-box A/B): 2 = the 8-warp epilogue
      const char* e = getenv("NW_B200_GRADT_SETS");
      return e && *e ? atoi(e) : 4;
    }();
    if (gradt_sets == 4) {
      p.sets = k1::QUAD_SETS;
      NW_LAUNCH_EMIT4(k1::MODE_EMIT_GRADT);
    } else {
      NW_LAUNCH_EMIT(k1::MODE_EMIT_GRADT);
    }
  }
  else if (emit_kind == EMIT_COEF_INTERNAL) {
    // exp, reciprocal square root, table lookup, rounding and 64-byte stores per pair: latency-bound like the
    // influence transform, so four epilogue sets
    p.sets = COEF_SETS;
    NW_LAUNCH_EMIT4(k1::MODE_EMIT_COEF);
    if (rc == NW_OK && ex.coef_sums != nullptr) {
      coef_sum_kernel<<<ceil_div(n_query, 256), 256, 0, stream>>>(ex.coef_ws, plan.chunks * COEF_SETS, n_query,
                                                                  ex.coef_sums);
      NW_CUDA_OK(cudaGetLastError());
    }
  }
  else if (emit_kind == NW_EMIT_INFLUENCE && emit_sets == 4) {
    p.sets = k1::QUAD_SETS;
    NW_LAUNCH_EMIT4(k1::MODE_EMIT_INFLUENCE);
  } else if (emit_kind == NW_EMIT_INFLUENCE) NW_LAUNCH_EMIT(k1::MODE_EMIT_INFLUENCE);
  else NW_LAUNCH_EMIT(k1::MODE_EMIT_BLOCKBEST);
#undef NW_LAUNCH_EMIT4
#undef NW_LAUNCH_EMIT
  return rc;
}

extern "C" int nw_forward_emit(int epilogue, float scale, const void* q_bf16, const float* q_sqnorm, int n_query,
                               const void* bank_bf16, const float* s_sqnorm, const int32_t* labels,
                               int64_t n_support, int row_elems, int emit_kind, const float* row_lse,
                               const float* p_query, const int32_t* qlabel, float* out, int64_t ld_out,
                               void* stream_) {
  NW_REQUIRE(emit_kind == NW_EMIT_SCORES || emit_kind == NW_EMIT_INFLUENCE || emit_kind == NW_EMIT_BLOCK_BEST,
             NW_ERR_INVALID, "unknown emit kind %d", emit_kind);
  return emit_impl(epilogue, scale, q_bf16, q_sqnorm, n_query, bank_bf16, s_sqnorm, labels, n_support, row_elems,
                   emit_kind, row_lse, p_query, qlabel, out, ld_out, EmitExtra(), stream_);
}

extern "C" int64_t nw_backward_coefficients_workspace_elems(int64_t n_rows, int64_t n_cols) {
  nw_forward_plan_t plan;
  if (n_rows <= 0 || n_rows >= (int64_t(1) << 31) - 512 || nw_forward_plan(int(n_rows), n_cols, &plan) != NW_OK) return -1;
  return int64_t(plan.chunks) * COEF_SETS * n_rows;
}

// Tensor-core backward, step 1 (recompute): see coef_chunk.  `rows` / `cols` are the two operands of the score
// GEMM in this kernel's k-block-major bf16 layout; orientation 0: rows = queries, cols = the class-sorted bank;
// orientation 1: rows = the bank, cols = queries.
extern "C" int nw_backward_coefficients(int epilogue, float scale, int orientation, const void* rows_bf16,
                                        const float* rows_sqnorm, int64_t n_rows, const void* cols_bf16,
                                        const float* cols_sqnorm, int64_t n_cols, int row_elems,
                                        const float* row_lse, const int32_t* row_labels, const float* col_lse,
                                        const int32_t* col_labels, const float* table, int64_t table_ld,
                                        void* out_bf16, float* row_sums, float* workspace,
                                        int64_t workspace_elems, void* stream_) {
  NW_REQUIRE(orientation == 0 || orientation == 1, NW_ERR_INVALID, "orientation must be 0 or 1");
  NW_REQUIRE(table != nullptr && out_bf16 != nullptr, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_rows > 0 && n_rows < (int64_t(1) << 31) - 512, NW_ERR_UNSUPPORTED, "n_rows must be in (0, 2^31 - 512)");
  NW_REQUIRE((reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0, NW_ERR_INVALID, "out_bf16 must be 16-byte aligned");
  if (orientation == 0)
    NW_REQUIRE(row_lse && col_labels, NW_ERR_INVALID, "orientation 0 needs row_lse and col_labels");
  else
    NW_REQUIRE(col_lse && row_labels, NW_ERR_INVALID, "orientation 1 needs col_lse and row_labels");
  NW_REQUIRE(row_sums != nullptr && workspace != nullptr, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(workspace_elems >= nw_backward_coefficients_workspace_elems(n_rows, n_cols), NW_ERR_WORKSPACE,
             "workspace too small: %lld < %lld floats", (long long)workspace_elems,
             (long long)nw_backward_coefficients_workspace_elems(n_rows, n_cols));
  EmitExtra ex;
  ex.coef_tab = table;
  ex.coef_ld = table_ld;
  ex.coef_orient = orientation;
  ex.coef_ws = workspace;
  ex.coef_sums = row_sums;
  // orientation 1: the per-column log-sum-exp travels through the label slots of the tile metadata (float bits)
  const int32_t* col_meta = orientation == 0 ? col_labels : reinterpret_cast<const int32_t*>(col_lse);
  return emit_impl(epilogue, scale, rows_bf16, rows_sqnorm, int(n_rows), cols_bf16, cols_sqnorm, col_meta, n_cols,
                   row_elems, EMIT_COEF_INTERNAL, row_lse, nullptr, row_labels, reinterpret_cast<float*>(out_bf16),
                   /*ld_out=*/0, ex, stream_);
}

// Tensor-core backward, grad_s in one launch: out[dst(c)][r] = sum_k a[r][k] * b[c][k] - col_sub[c] * rows_t(r, c) for
// r < n_out_cols (see gradt_chunk).  a: Q^t (n_a >= n_out_cols rows, the features), b: W^t (n_b rows, the supports).
extern "C" int nw_dense_products_transposed(const void* a_bf16, int64_t n_a, const void* b_bf16, int64_t n_b,
                                            int k_elems, const float* col_sub, const void* rows_t_bf16,
                                            const int32_t* dst_rows, int n_out_cols, float* out, int64_t ld_out,
                                            void* stream_) {
  NW_REQUIRE(n_a > 0 && n_a < (int64_t(1) << 31) - 512, NW_ERR_UNSUPPORTED, "n_a must be in (0, 2^31 - 512)");
  NW_REQUIRE(n_out_cols > 0 && n_out_cols <= n_a && ld_out >= n_out_cols, NW_ERR_INVALID,
             "need 0 < n_out_cols <= n_a and ld_out >= n_out_cols");
  NW_REQUIRE((col_sub == nullptr) == (rows_t_bf16 == nullptr), NW_ERR_INVALID,
             "col_sub and rows_t_bf16 go together");
  NW_REQUIRE(rows_t_bf16 == nullptr || (reinterpret_cast<uintptr_t>(rows_t_bf16) & 15) == 0, NW_ERR_INVALID,
             "rows_t_bf16 must be 16-byte aligned");
  EmitExtra ex;
  ex.gt_rows = rows_t_bf16;
  ex.gt_valid_rows = n_out_cols;
  return emit_impl(NW_EPI_LINEAR, 1.0f, a_bf16, nullptr, int(n_a), b_bf16, col_sub, dst_rows, n_b, k_elems,
                   EMIT_GRADT_INTERNAL, nullptr, nullptr, nullptr, out, ld_out, ex, stream_);
}

// Tensor-core backward, step 2: out[ks][r][c] = sum over K slice ks of a[r][k] * b[c][k]  (both operands bf16,
// k-block-major), fp32, row-major with leading dimension ld_out.  The slices' partial products are summed by the
// caller (split-K keeps all SMs busy when (n_a / 256) * (n_b / 256) is smaller than the GPU).
extern "C" int nw_dense_products(const void* a_bf16, int64_t n_a, const void* b_bf16, int64_t n_b, int k_elems,
                                 int kslices, float* out, int64_t ld_out, int64_t slice_stride, void* stream_) {
  NW_REQUIRE(n_a > 0 && n_a < (int64_t(1) << 31) - 512, NW_ERR_UNSUPPORTED, "n_a must be in (0, 2^31 - 512)");
  NW_REQUIRE(kslices >= 1 && kslices <= 1024, NW_ERR_INVALID, "kslices must be in [1, 1024]");
  NW_REQUIRE(kslices == 1 || slice_stride >= n_a * ld_out, NW_ERR_INVALID, "slice_stride must cover one output");
  EmitExtra ex;
  ex.kslices = kslices;
  ex.slice_stride = slice_stride;
  return emit_impl(NW_EPI_LINEAR, 1.0f, a_bf16, nullptr, int(n_a), b_bf16, nullptr, nullptr, n_b, k_elems,
                   NW_EMIT_SCORES, nullptr, nullptr, nullptr, out, ld_out, ex, stream_);
}

extern "C" int nw_logp_from_class_lse(const float* class_lse, int n_query, int n_classes, float* logp,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(class_lse && logp, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(n_query > 0 && n_classes > 0, NW_ERR_INVALID, "n_query and n_classes must be positive");
  const int blocks = (n_query + 7) / 8;
  if (n_classes <= 256) k1::logp_rows_kernel<8><<<blocks, 256, 0, stream>>>(class_lse, n_query, n_classes, logp);
  else if (n_classes <= 1024) k1::logp_rows_kernel<32><<<blocks, 256, 0, stream>>>(class_lse, n_query, n_classes, logp);
  else k1::logp_kernel<<<n_query, 256, 0, stream>>>(class_lse, n_classes, logp);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_row_stats(const float* class_lse, int n_query, int n_classes, const int32_t* qlabel, float* row_lse,
                            float* p_query, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(class_lse && row_lse, NW_ERR_INVALID, "NULL pointer argument");
  NW_REQUIRE(p_query == nullptr || qlabel != nullptr, NW_ERR_INVALID, "p_query needs qlabel");
  NW_REQUIRE(n_query > 0 && n_classes > 0, NW_ERR_INVALID, "n_query and n_classes must be positive");
  k1::row_stats_kernel<<<n_query, 256, 0, stream>>>(class_lse, n_classes, qlabel, row_lse, p_query);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}

extern "C" int nw_class_lse_merge(float* a, const float* b, int64_t n_elems, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NW_REQUIRE(a && b && n_elems >= 0, NW_ERR_INVALID, "bad arguments");
  if (n_elems == 0) return NW_OK;
  k1::lse_merge_kernel<<<sm_count() * 4, 256, 0, stream>>>(a, b, (long long)n_elems);
  NW_CUDA_OK(cudaGetLastError());
  return NW_OK;
}
