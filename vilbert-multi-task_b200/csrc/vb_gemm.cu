// vb_gemm.cu — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = alpha * A[M,K] . B[N,K]^T  (bf16 operands, fp32 accumulation in TMEM) + fused epilogue.
//
// Replaces the aten addmm behind every nn.Linear on the ViLBERT hot path and its autograd
// (reference: vilbert/vilbert.py:410-412,466,492,509,553-555,625,653,670,716-725,830,837,865-869 and
// SURVEY.md appendix A). Design:
//   warp 0      TMA producer: cp.async.bulk.tensor 2D boxes (64 x rows, 128B swizzle) into a
//               NUM_STAGES-deep smem ring, completion on mbarriers (complete_tx);
//   warp 1      single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, cta_group::1), accumulators
//               double-buffered in TMEM (2 x BN fp32 columns) so the epilogue of tile i overlaps the
//               main loop of tile i+1; tcgen05.commit releases smem stages / publishes accumulators;
//   warps 2..9  epilogue (two per TMEM lane quadrant, alternate 32-column chunks): tcgen05.ld (32 lanes x 32 columns)
//               -> XOR-swizzled smem transpose -> row-wise,
//               128-bit coalesced pass doing bias / erf-GELU / GELU' / residual / bf16 conversion.
// Operands may be K-major (nn.Linear forward) or MN-major (dgrad / wgrad operands read in place, no
// transposed copies); both use the canonical SWIZZLE_128B UMMA layouts written directly by TMA.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vb_internal.h"
#include "vb_ptx.cuh"

namespace vb {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // one 128-byte swizzle span of bf16
constexpr int UK = 16;           // UMMA K for 16-bit inputs
constexpr int EPI_WARPS = 8;      // two epilogue warps per TMEM lane quadrant: twice the global-memory requests in flight
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
// (Round 2 experiment, rejected: staging the fp32 residual of the F32 epilogue through a TMA ring fed by an 11th warp. The
// epilogue is not latency-bound but bandwidth-bound — all 148 CTAs run their epilogues in lockstep, 64 KB read + 64 KB written
// per tile = 7.5 TB/s chip-wide during that phase — and giving up two operand stages for the ring slowed the main loop:
// 34.3 us vs 29.8 us on 6400x1024x1024, profiles/r02_gemm_timeline_residual_tma_ring_rejected.log.)
// Epilogue staging tile per warp: 32 rows x 32 fp32, dense 128-byte rows whose 16-byte chunks are XOR-swizzled with the
// row index (chunk ^ (row & 7)): thread-per-row 128-bit stores and row-wise 128-bit loads are both bank-conflict free
// without padding (4 KB per warp).
constexpr int STAGING_BYTES_PER_WARP = 32 * 32 * 4;
__device__ __forceinline__ int stg_off(int row, int col) { return row * 32 + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3)); }

// CG = 1: one CTA per 128 x BN tile. CG = 2: CTA pairs (tcgen05 cta_group::2) on a 256 x BN tile; each CTA stages its 128
// rows of A and its BN/2 rows of B, so a stage is smaller and the ring deeper (192 KB of operand stages either way).
template <int BN, int CG>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NUM_STAGES = (192 * 1024) / STAGE_BYTES;   // 6 / 4 (CG 1), 8 / 6 (CG 2)
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = NUM_STAGES * STAGE_BYTES + EPI_WARPS * STAGING_BYTES_PER_WARP + 256 /*barriers*/;
};

struct GemmKernelParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;   // num_k_blocks counts VIRTUAL k-blocks: npass x real_k_blocks
  int real_k_blocks;         // ceil(K / 64)
  int npass;                 // 1, or 2..3 in split precision: pass_sel[i] bit 0 = A_lo, bit 1 = B_lo
  int pass_sel[3];
  int out_fp16;              // format of out_bf16 / out_lo: 0 = bf16, 1 = fp16
  __nv_bfloat16* out_lo;     // split precision: low part of the value written to out_bf16 (same pitch), or NULL
  __nv_bfloat16* out_b16;    // always-bf16 copy of the value written to out_bf16 (same pitch), or NULL
  int split_k, k_blocks_per_split;
  float alpha;
  const float* bias;
  const float* residual;
  long long ld_res;
  const __nv_bfloat16* aux;
  long long ld_aux;
  int act;
  float* out_f32;
  long long ld_of;
  __nv_bfloat16* out_bf16;
  long long ld_ob;
  __nv_bfloat16* out_pre;
  long long ld_op;
  int atomic_out;
  float* out_colsum;
  int vec_f32, vec_bf16, vec_pre, vec_res, vec_aux;  // 128/64-bit access legal for that buffer
  uint64_t desc_base_a, desc_base_b;                 // smem descriptor without the address field
  uint32_t kadv_a, kadv_b;                           // descriptor address advance per UMMA_K (bytes)
  uint32_t idesc;
  DropCfg drop;              // dropout on the epilogue value before the residual add (EPI_F32 / generic)
  int a_mn, b_mn;            // operand majors (runtime: only the TMA producer cares)
  int cluster;               // 1, or 2 = CTA pairs (tcgen05 cta_group::2): the two CTAs take adjacent row blocks of one column
                             // block; the leader issues 256 x BN MMAs for both, each CTA stages its 128 rows of A and its
                             // half of the B tile -> per-SM operand traffic (L2 -> smem and smem -> tensor core) drops by 25-33 %
  int num_m_groups;          // ceil(num_m_blocks / cluster)
  int fast_ok;               // every buffer the specialised epilogue touches allows 128/64-bit accesses
  unsigned long long* dbg;   // optional per-CTA timeline [grid][10] (8 x clock64 + 2 x globaltimer ns), NULL in production
};

// Epilogue specialisations. Each instantiation keeps ONE compact, fully unrolled fast path (whole 32x32 chunk inside
// the matrix, 128/64-bit aligned buffers) plus a shared non-inlined generic path for ragged edges / odd layouts.
// (A single kernel with every variant inlined is ~180 KB of SASS and thrashes the instruction cache: the epilogue
// of one 128x128 tile then costs ~29k cycles instead of ~2k — measured with the clock64 timeline, profiles/.)
enum { EPI_F32 = 0,     // v = alpha*acc (+bias) (ReLU) (+fp32 residual) -> out_f32 (+ 16-bit copy)   (out-proj / FFN2 / dgrad-into-residual / logits / poolers)
       EPI_BF16 = 1,    // v = alpha*acc (+bias) -> out_bf16                                 (QKV, plain dgrads)
       EPI_GELU = 2,    // pre = acc + bias; gelu(pre) -> out_bf16 / out_f32; gelu'(pre) -> out_pre (bf16, saved for backward)
       EPI_DGELU = 3,   // v = acc * aux (aux = saved gelu'(pre)) -> out_bf16 (+ column sums)
       EPI_ATOMIC = 4,  // red.global.add.v4.f32 into out_f32 (split-K wgrad)
       EPI_GENERIC = 5, // runtime flags only (ReLU poolers, unusual output combinations)
       EPI_COUNT = 6 };

struct EpiCtx {
  int m_base, n, nv, rr, cc;
  bool full4;
};

// Generic, compact (non-unrolled) path: any flags, any alignment, ragged rows / columns.
__device__ __noinline__ void epi_generic_chunk(const GemmKernelParams& p, const float* stg, int m_base, int n, int rr, int cc) {
  const int nv = min(4, p.N - n);
  if (nv <= 0) return;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int ps = 0; ps < 8; ++ps) {
    const int row = ps * 4 + rr;
    const long long m = m_base + row;
    if (m >= p.M) break;
#pragma unroll 1
    for (int j = 0; j < nv; ++j) {
      float v = stg[stg_off(row, cc + j)] * p.alpha;
      if (p.bias) v += p.bias[n + j];
      if (p.act == VB_ACT_GELU) {
        float gg, dg;
        gelu_erf_and_grad(v, gg, dg);
        if (p.out_pre) p.out_pre[m * p.ld_op + n + j] = __float2bfloat16(dg);   // saved for backward: gelu'(pre)
        v = gg;
      } else if (p.act == VB_ACT_RELU) {
        v = fmaxf(v, 0.f);
      } else if (p.act == VB_ACT_DGELU) {
        v *= __bfloat162float(p.aux[m * p.ld_aux + n + j]);                      // aux = saved gelu'(pre)
      }
      if (p.drop.ctr) v = drop_apply(v, drop_seed(p.drop), (uint32_t)(m * p.N + n + j), p.drop);
      cs[j] += v;
      if (p.residual) v += p.residual[m * p.ld_res + n + j];
      if (p.out_f32) {
        if (p.atomic_out) atomicAdd(p.out_f32 + m * p.ld_of + n + j, v);
        else p.out_f32[m * p.ld_of + n + j] = v;
      }
      if (p.out_bf16) {
        const uint16_t hi = cvt16(v, p.out_fp16);
        reinterpret_cast<uint16_t*>(p.out_bf16)[m * p.ld_ob + n + j] = hi;
        if (p.out_lo) reinterpret_cast<uint16_t*>(p.out_lo)[m * p.ld_ob + n + j] = cvt16(v - cvt16_to_f32(hi, p.out_fp16), p.out_fp16);
        if (p.out_b16) p.out_b16[m * p.ld_ob + n + j] = __float2bfloat16(v);
      }
    }
  }
  if (p.out_colsum) {
#pragma unroll 1
    for (int j = 0; j < nv; ++j) atomicAdd(p.out_colsum + n + j, cs[j]);
  }
}

// 16-bit output variants of the specialised epilogues (template parameter OUT16; kept out of line / out of the kernels that do
// not need them: the fast paths are 8x unrolled and the epilogue time follows the instruction footprint, see the note above):
//   0 = bf16 (gradient operands, bf16 precision)   1 = fp16 (forward operands)   2 = fp16 + an always-bf16 copy (forward
//   operands the backward's weight-gradient GEMM reads)   3 = split precision: fp16 hi + lo (+ bf16 copy at run time).
__device__ __noinline__ void store16x4_slow(const GemmKernelParams& p, long long off, float v0, float v1, float v2, float v3) {
  if (p.out_lo) {
    uint32_t l01, l23;
    const uint32_t h01 = pack16_split(v0, v1, p.out_fp16, l01), h23 = pack16_split(v2, v3, p.out_fp16, l23);
    *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(p.out_lo + off) = make_uint2(l01, l23);
  } else {
    *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack16(v0, v1, p.out_fp16), pack16(v2, v3, p.out_fp16));
  }
  if (p.out_b16) *reinterpret_cast<uint2*>(p.out_b16 + off) = make_uint2(pack_bf16(v0, v1), pack_bf16(v2, v3));
}
template <int OUT16>
__device__ __forceinline__ void store16x4(const GemmKernelParams& p, long long off, float v0, float v1, float v2, float v3) {
  if (OUT16 == 0) {
    *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_bf16(v0, v1), pack_bf16(v2, v3));
  } else if (OUT16 == 3) {   // split precision: fp16 hi + lo (+ the bf16 copy when asked for)
    uint32_t l01, l23;
    const uint32_t h01 = pack16_split(v0, v1, 1, l01), h23 = pack16_split(v2, v3, 1, l23);
    *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(p.out_lo + off) = make_uint2(l01, l23);
    if (p.out_b16) *reinterpret_cast<uint2*>(p.out_b16 + off) = make_uint2(pack_bf16(v0, v1), pack_bf16(v2, v3));
  } else {
    *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_f16(v0, v1), pack_f16(v2, v3));
    if (OUT16 == 2) *reinterpret_cast<uint2*>(p.out_b16 + off) = make_uint2(pack_bf16(v0, v1), pack_bf16(v2, v3));
  }
}

// Poolers (vilbert.py:1116-1122, 1131-1137): relu(acc + bias) -> fp32 and a 16-bit operand copy. Two tiny launches per step: one
// compact out-of-line routine selected per chunk inside the F32 kernel, so that kernel's unrolled fast path stays as small as it was.
__device__ __noinline__ void epi_pool_chunk(const GemmKernelParams& p, const float* stg, int m_base, int n, int rr, int cc, const float4 b4) {
#pragma unroll 1
  for (int ps = 0; ps < 8; ++ps) {
    const int row = ps * 4 + rr;
    const long long m = m_base + row;
    const float4 a4 = *reinterpret_cast<const float4*>(stg + stg_off(row, cc));
    const float v0 = fmaxf(fmaf(a4.x, p.alpha, b4.x), 0.f), v1 = fmaxf(fmaf(a4.y, p.alpha, b4.y), 0.f);
    const float v2 = fmaxf(fmaf(a4.z, p.alpha, b4.z), 0.f), v3 = fmaxf(fmaf(a4.w, p.alpha, b4.w), 0.f);
    *reinterpret_cast<float4*>(p.out_f32 + m * p.ld_of + n) = make_float4(v0, v1, v2, v3);
    if (p.out_bf16) store16x4_slow(p, m * p.ld_ob + n, v0, v1, v2, v3);
  }
}

template <int EPI, int OUT16>
__device__ __forceinline__ void epi_fast_chunk(const GemmKernelParams& p, const float* stg, int m_base, int n, int rr, int cc,
                                               const float4 (&resv)[8], const uint2 (&auxv)[8], const float4 b4) {
  float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
  const uint32_t dseed = (EPI == EPI_F32 && p.drop.ctr) ? drop_seed(p.drop) : 0u;
#pragma unroll
  for (int ps = 0; ps < 8; ++ps) {
    const int row = ps * 4 + rr;
    const long long m = m_base + row;
    const float4 a4 = *reinterpret_cast<const float4*>(stg + stg_off(row, cc));
    float v0 = fmaf(a4.x, p.alpha, b4.x), v1 = fmaf(a4.y, p.alpha, b4.y), v2 = fmaf(a4.z, p.alpha, b4.z), v3 = fmaf(a4.w, p.alpha, b4.w);
    if (EPI == EPI_GELU) {
      float d0, d1, d2, d3;
      gelu_erf_and_grad(v0, v0, d0); gelu_erf_and_grad(v1, v1, d1); gelu_erf_and_grad(v2, v2, d2); gelu_erf_and_grad(v3, v3, d3);
      *reinterpret_cast<uint2*>(p.out_pre + m * p.ld_op + n) = make_uint2(pack_bf16(d0, d1), pack_bf16(d2, d3));   // gelu'(pre) for backward
      if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + m * p.ld_of + n) = make_float4(v0, v1, v2, v3);
      if (p.out_bf16) store16x4<OUT16>(p, m * p.ld_ob + n, v0, v1, v2, v3);
    } else if (EPI == EPI_DGELU) {
      const float2 x01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&auxv[ps].x));
      const float2 x23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&auxv[ps].y));
      v0 *= x01.x; v1 *= x01.y; v2 *= x23.x; v3 *= x23.y;   // aux = gelu'(pre) saved by the forward epilogue
      cs0 += v0; cs1 += v1; cs2 += v2; cs3 += v3;
      store16x4<0>(p, m * p.ld_ob + n, v0, v1, v2, v3);   // a gradient operand: always bf16
    } else if (EPI == EPI_F32) {
      if (p.drop.ctr) {   // LN(dropout(dense(x)) + residual): mask the dense output, element index m*N + n
        const uint32_t e0 = (uint32_t)(m * p.N + n);
        v0 = drop_apply(v0, dseed, e0, p.drop); v1 = drop_apply(v1, dseed, e0 + 1, p.drop);
        v2 = drop_apply(v2, dseed, e0 + 2, p.drop); v3 = drop_apply(v3, dseed, e0 + 3, p.drop);
      }
      if (p.residual) { v0 += resv[ps].x; v1 += resv[ps].y; v2 += resv[ps].z; v3 += resv[ps].w; }
      float* dst = p.out_f32 + m * p.ld_of + n;
      if (p.vec_f32) {
        *reinterpret_cast<float4*>(dst) = make_float4(v0, v1, v2, v3);
      } else {   // row pitch not a multiple of 4 floats (30522-/1601-/3129-wide logits): same bytes, 32-bit stores
        dst[0] = v0; dst[1] = v1; dst[2] = v2; dst[3] = v3;
      }
    } else if (EPI == EPI_BF16) {
      store16x4<OUT16>(p, m * p.ld_ob + n, v0, v1, v2, v3);
    } else if (EPI == EPI_ATOMIC) {
      asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(p.out_f32 + m * p.ld_of + n), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
    }
  }
  if (EPI == EPI_DGELU && p.out_colsum) {
    // reduce over the 4 row-lanes sharing these columns (lane bits 3,4), then one atomic per column
    cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
    cs2 += __shfl_xor_sync(0xffffffffu, cs2, 8);  cs3 += __shfl_xor_sync(0xffffffffu, cs3, 8);
    cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16); cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
    cs2 += __shfl_xor_sync(0xffffffffu, cs2, 16); cs3 += __shfl_xor_sync(0xffffffffu, cs3, 16);
    if (rr == 0) {
      atomicAdd(p.out_colsum + n, cs0); atomicAdd(p.out_colsum + n + 1, cs1);
      atomicAdd(p.out_colsum + n + 2, cs2); atomicAdd(p.out_colsum + n + 3, cs3);
    }
  }
}

template <int BN, int EPI, int CG, int OUT16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_a_lo, const __grid_constant__ CUtensorMap tmap_b_lo,
                    const GemmKernelParams p) {
  using Cfg = GemmCfg<BN, CG>;
  constexpr bool pair = (CG == 2);
  constexpr int NUM_STAGES = Cfg::NUM_STAGES;

  // SWIZZLE_128B tiles need 1024-byte alignment: the kernel has no static shared memory, so the dynamic window starts at
  // the CTA's (1024-aligned) shared base; checked once below.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_tiles = smem;
  float* staging = reinterpret_cast<float*>(smem + NUM_STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NUM_STAGES * Cfg::STAGE_BYTES + EPI_WARPS * STAGING_BYTES_PER_WARP);
  uint64_t* full_bar = bars;                       // [NUM_STAGES]
  uint64_t* empty_bar = bars + NUM_STAGES;         // [NUM_STAGES]
  uint64_t* tmem_full_bar = bars + 2 * NUM_STAGES; // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#define VB_DBG(slot) do { if (p.dbg) p.dbg[blockIdx.x * 10 + (slot)] = clock64(); } while (0)
#define VB_DBG_NS(slot) do { if (p.dbg) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.dbg[blockIdx.x * 10 + (slot)] = t_; } } while (0)
  if (threadIdx.x == 0) { VB_DBG(0); VB_DBG_NS(8); }

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.npass > 1) { tma_prefetch_desc(&tmap_a_lo); tma_prefetch_desc(&tmap_b_lo); }
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tmem_full_bar[s]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[s]), EPI_WARPS * CG);   // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp_idx == 1) {
    if constexpr (pair) { tmem_alloc_pair(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS); tmem_relinquish_pair(); }
    else                { tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // the barriers of both CTAs of a pair must exist before the peer's TMA / commits / arrives reach them
  if constexpr (pair) cluster_sync_all();
  pdl_entry();   // everything above (barrier init, TMEM alloc, descriptor prefetch) overlapped the previous kernel's tail
  if (threadIdx.x == 0) VB_DBG(1);

  // work items = (row-block group, column block, k split); the CTAs of a pair walk the same items in lockstep, CTA
  // `crank` takes row block group * cluster + crank (possibly past the matrix: it then loads zero rows and stores
  // nothing, but still stages its half of B)
  const int crank = pair ? (int)cluster_ctarank() : 0;
  const int group = blockIdx.x / CG;
  const int num_groups = gridDim.x / CG;
  const int total_work = p.num_m_groups * p.num_n_blocks * p.split_k;
  const bool leader = (crank == 0);

  if (warp_idx == 0) {
    // ================================================================ TMA producer (one elected lane issues; the warp stays converged).
    // The guard must be elect.sync, not `lane == 0`: only then does ptxas know a single thread is active and emit the
    // UTMALDG / UTCHMMA / UTCBAR once instead of inside a per-thread BRA.U.ANY serialisation loop (~80 cycles per MMA)
    int stage = 0;
    uint32_t phase = 0;
    for (int w = group; w < total_work; w += num_groups) {
      const int split = w % p.split_k;
      const int t2 = w / p.split_k;
      const int m_blk = (t2 % p.num_m_groups) * CG + crank;
      const int n_blk = t2 / p.num_m_groups;
      const int kb0 = split * p.k_blocks_per_split;
      const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
      // loads of one (real) k-block from the given tensor maps. Single-pass launches call it with the kernel's own
      // __grid_constant__ maps (compile-time parameter addresses, as in round 1); split precision selects hi / lo maps per pass.
      auto issue_loads = [&](const CUtensorMap* tma_a, const CUtensorMap* tma_b, int kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t sa = smem_u32(smem_tiles + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
        if constexpr (!pair) {
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES);
          if (p.a_mn) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), tma_a, m_blk * BM + j * 64, kb * BK, fb);
          } else {
            tma_load_2d(sa, tma_a, kb * BK, m_blk * BM, fb);
          }
          if (p.b_mn) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), tma_b, n_blk * BN + j * 64, kb * BK, fb);
          } else {
            tma_load_2d(sb, tma_b, kb * BK, n_blk * BN, fb);
          }
        } else {
          // both CTAs complete their bytes on the LEADER's full barrier (the leader issues the MMAs for the pair);
          // this CTA stages its 128 rows of A and columns [crank * BN/2, +BN/2) of the B tile
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), 2 * Cfg::STAGE_BYTES);
          if (p.a_mn) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * (BK * 128), tma_a, m_blk * BM + j * 64, kb * BK, fb);
          } else {
            tma_load_2d_pair(sa, tma_a, kb * BK, m_blk * BM, fb);
          }
          const int n0 = n_blk * BN + crank * (BN / 2);
          if (p.b_mn) {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_2d_pair(sb + j * (BK * 128), tma_b, n0 + j * 64, kb * BK, fb);
          } else {
            tma_load_2d_pair(sb, tma_b, kb * BK, n0, fb);   // tensor-map box = BN/2 rows
          }
        }
      };
      for (int kbv = kb0; kbv < kb1; ++kbv) {
        if (elect_one()) {
          if (p.npass == 1) {
            issue_loads(&tmap_a, &tmap_b, kbv);
          } else {
            // virtual k-block -> (pass, real k-block): split precision walks K once per pass with the hi / lo tensor maps
            const int ps = kbv / p.real_k_blocks;
            const int sel = p.pass_sel[ps];
            issue_loads((sel & 1) ? &tmap_a_lo : &tmap_a, (sel & 2) ? &tmap_b_lo : &tmap_b, kbv - ps * p.real_k_blocks);
          }
          if (kbv == kb0 && w == group) VB_DBG(2);
        }
        __syncwarp();
        if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_idx == 1) {
    // ================================================================ MMA issuer (one elected lane issues; the warp stays converged)
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int w = group; w < total_work && leader; w += num_groups, ++it) {   // in a pair only the leader issues
      const int split = w % p.split_k;
      const int kb0 = split * p.k_blocks_per_split;
      const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const uint32_t tmem_d = tmem_base + as * BN;
      if (elect_one()) {
        mbar_wait(smem_u32(&tmem_empty_bar[as]), aphase ^ 1);
        tc_fence_after();
      }
      __syncwarp();
      for (int kb = kb0; kb < kb1; ++kb) {
        if (elect_one()) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          if (kb == kb0 && w == group) VB_DBG(3);
          const uint32_t sa = smem_u32(smem_tiles + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = umma_desc_at(p.desc_base_a, sa + k * p.kadv_a);
            const uint64_t db = umma_desc_at(p.desc_base_b, sb + k * p.kadv_b);
            if constexpr (pair) umma_bf16_pair(tmem_d, da, db, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else                umma_bf16(tmem_d, da, db, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // smem slot free once these MMAs retire (in both CTAs of a pair)
          if constexpr (pair) umma_commit_pair(smem_u32(&empty_bar[stage]), 0x3);
          else                umma_commit(smem_u32(&empty_bar[stage]));
        }
        __syncwarp();
        if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) {
        if constexpr (pair) umma_commit_pair(smem_u32(&tmem_full_bar[as]), 0x3);   // accumulator complete (both halves)
        else                umma_commit(smem_u32(&tmem_full_bar[as]));
        if (w == group) VB_DBG(4);
      }
      __syncwarp();
    }
  } else {
    // ================================================================ epilogue (warps 2..9)
    // Per 32-column chunk: (1) issue the global reads of this chunk (fp32 residual / bf16 GELU pre-activation) so their
    // latency overlaps the TMEM read, (2) tcgen05.ld 32 lanes x 32 columns -> registers (thread = row), (3) 128-bit
    // stores into a padded smem tile, (4) row-wise pass where a lane owns 4 consecutive columns of rows {rr, rr+4, ..}:
    // 128-bit smem reads, fused math, 128-bit coalesced global stores.
    const int lane_grp = warp_idx & 3;  // TMEM lanes [32*lane_grp, +32) are visible to this warp
    float* stg = staging + (warp_idx - 2) * (32 * 32);
    const int half = (warp_idx - 2) >> 2;   // the two warps of a lane quadrant take alternate 32-column chunks
    const int rr = lane >> 3;           // row within a 4-row group of the coalesced pass
    const int cc = (lane & 7) * 4;      // first of 4 columns handled by this lane
    int it = 0;
    for (int w = group; w < total_work; w += num_groups, ++it) {
      const int t2 = w / p.split_k;
      const int m_blk = (t2 % p.num_m_groups) * CG + crank;
      const int n_blk = t2 / p.num_m_groups;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m_base = m_blk * BM + lane_grp * 32;
      const uint32_t taddr = tmem_base + (uint32_t(lane_grp * 32) << 16) + as * BN;
      const bool rows_full = (m_base + 32 <= p.M);
      const bool rows_live = (m_base < p.M);
      bool waited = false;
      constexpr int NC = BN / 32;
      // software pipeline over chunk pairs: the global operands of chunk c+1 are in flight while chunk c is processed
      float4 res0[8], res1[8];
      uint2 aux0[8], aux1[8];
      float4 bia0, bia1;
      auto chunk_fast = [&](int c) -> bool {
        return (EPI != EPI_GENERIC) && p.fast_ok && rows_full && (n_blk * BN + c * 32 + 32 <= p.N);
      };
      auto prefetch = [&](int c, float4 (&resv)[8], uint2 (&auxv)[8], float4& b4) {
        b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c >= NC || !rows_live || !chunk_fast(c)) return;
        const int n = n_blk * BN + c * 32 + cc;
        if (EPI == EPI_F32 && p.residual) {
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const float* src = p.residual + (long long)(m_base + ps * 4 + rr) * p.ld_res + n;
            if (p.vec_res) resv[ps] = *reinterpret_cast<const float4*>(src);
            else resv[ps] = make_float4(src[0], src[1], src[2], src[3]);
          }
        }
        if (EPI == EPI_DGELU) {
#pragma unroll
          for (int ps = 0; ps < 8; ++ps)
            auxv[ps] = *reinterpret_cast<const uint2*>(p.aux + (long long)(m_base + ps * 4 + rr) * p.ld_aux + n);
        }
        if (p.bias) b4 = *reinterpret_cast<const float4*>(p.bias + n);
      };
      auto process = [&](int c, const float4 (&resv)[8], const uint2 (&auxv)[8], const float4 b4) {
        const int n_chunk = n_blk * BN + c * 32;
        const bool chunk_live = (n_chunk < p.N) && rows_live;  // warp-uniform
        if (!waited) {
          mbar_wait(smem_u32(&tmem_full_bar[as]), aphase);
          tc_fence_after();
          waited = true;
          if (w == group && warp_idx == 2 && lane == 0) VB_DBG(5);
        }
        // TMEM -> registers -> padded smem tile (row = lane)
        if (chunk_live) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + stg_off(lane, j * 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
        if (c == NC - 2 + half) {
          // every TMEM read this warp makes of the accumulator stage has landed in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (pair) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[as]), 0));   // the leader's MMA warp waits for both CTAs
            else                mbar_arrive(smem_u32(&tmem_empty_bar[as]));
          }
        }
        if (!chunk_live) return;
        __syncwarp();
        // coalesced row pass
        if (EPI == EPI_F32 && p.act == VB_ACT_RELU && chunk_fast(c)) epi_pool_chunk(p, stg, m_base, n_chunk + cc, rr, cc, b4);
        else if (chunk_fast(c)) epi_fast_chunk<EPI, OUT16>(p, stg, m_base, n_chunk + cc, rr, cc, resv, auxv, b4);
        else               epi_generic_chunk(p, stg, m_base, n_chunk + cc, rr, cc);
        __syncwarp();
      };
      prefetch(half, res0, aux0, bia0);
#pragma unroll 1
      for (int c = half; c < NC; c += 4) {
        prefetch(c + 2, res1, aux1, bia1);
        process(c, res0, aux0, bia0);
        prefetch(c + 4, res0, aux0, bia0);
        process(c + 2, res1, aux1, bia1);
      }
      if (w == group && warp_idx == 2 && lane == 0) VB_DBG(6);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  // the peer may still arrive on this CTA's barriers / the leader's MMAs may still write the peer's TMEM until both are done
  if constexpr (pair) cluster_sync_all();
  if (threadIdx.x == 0) { VB_DBG(7); VB_DBG_NS(9); }
  if (warp_idx == 1) {
    tc_fence_after();
    if constexpr (pair) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else                tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 2D bf16 tensor map: inner extent `inner` (contiguous), outer extent `outer` with row pitch ld elements.
static int make_tmap(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return VB_OK;
}

template <int BN, int EPI, int CG, int OUT16 = 0>
static int launch_gemm(const CUtensorMap* tm, GemmKernelParams& p, long long total_work, int max_ctas,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG>;
  auto kern = gemm_tcgen05_kernel<BN, EPI, CG, OUT16>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  // persistent grid: one CTA per SM; with clusters, as many clusters as the device can keep resident at once (a GPC
  // with an odd number of free SMs cannot host a pair there) so that no cluster waits for a second wave
  int groups_cap = max_ctas;
  if (CG > 1) {
    static int max_clusters = -1;   // per template instantiation; cluster size is 2 whenever it is not 1
    if (max_clusters < 0) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2 * 148); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      cudaError_t eo = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      if (eo != cudaSuccess || n <= 0) return set_error(VB_ERR_CUDA, "cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(eo));
      max_clusters = n;
    }
    groups_cap = max_ctas / CG < max_clusters ? max_ctas / CG : max_clusters;
  }
  const int groups = (int)(total_work < groups_cap ? total_work : groups_cap);
  const int grid = groups * CG;
  cudaError_t e = launch_pdl_cluster(kern, dim3(grid), dim3(GEMM_THREADS), (size_t)Cfg::SMEM_BYTES, stream, CG, tm[0], tm[1], tm[2], tm[3], p);
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  return VB_OK;
}

template <int BN, int CG>
static int launch_gemm_epi(int epi, int out16, const CUtensorMap* tm, GemmKernelParams& p, long long work, int max_ctas,
                           cudaStream_t stream) {
  switch (epi) {
    case EPI_F32: return launch_gemm<BN, EPI_F32, CG>(tm, p, work, max_ctas, stream);
    case EPI_BF16:
      if (out16 == 1) return launch_gemm<BN, EPI_BF16, CG, 1>(tm, p, work, max_ctas, stream);
      if (out16 == 2) return launch_gemm<BN, EPI_BF16, CG, 2>(tm, p, work, max_ctas, stream);
      if (out16 == 3) return launch_gemm<BN, EPI_BF16, CG, 3>(tm, p, work, max_ctas, stream);
      return launch_gemm<BN, EPI_BF16, CG, 0>(tm, p, work, max_ctas, stream);
    case EPI_GELU:
      if (out16 == 1) return launch_gemm<BN, EPI_GELU, CG, 1>(tm, p, work, max_ctas, stream);
      if (out16 == 2) return launch_gemm<BN, EPI_GELU, CG, 2>(tm, p, work, max_ctas, stream);
      if (out16 == 3) return launch_gemm<BN, EPI_GELU, CG, 3>(tm, p, work, max_ctas, stream);
      return launch_gemm<BN, EPI_GELU, CG, 0>(tm, p, work, max_ctas, stream);
    case EPI_DGELU: return launch_gemm<BN, EPI_DGELU, CG>(tm, p, work, max_ctas, stream);
    case EPI_ATOMIC: return launch_gemm<BN, EPI_ATOMIC, CG>(tm, p, work, max_ctas, stream);
    default: return launch_gemm<BN, EPI_GENERIC, CG>(tm, p, work, max_ctas, stream);
  }
}

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// cluster_m = 0 resolves to this (env VB_GEMM_CLUSTER=1|2 forces one mode, for experiments)
static int default_cluster() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VB_GEMM_CLUSTER");
    v = (e && (e[0] == '1' || e[0] == '2')) ? (e[0] - '0') : 0;   // 0 = chosen per problem by the cost model
  }
  return v;
}

// Modelled fixed cost of running a problem as CTA pairs. In isolation a pair launch costs only ~600 cycles more (cluster
// sync, leader-only issue), but inside the two-stream step a pair needs both SMs of a TPC free at once while kernels of the
// other stream hold SMs: sweeping this constant on the full training step (profiles/r01_bench_v15_pair_penalty_sweep.txt:
// 600 -> 11.24 ms, 2500 -> 11.15, 6000 -> 11.07, 10000 -> 11.11, 16000 -> 11.17, 30000 -> 11.18) puts the optimum at ~6000,
// i.e. pairs only where they save at least ~3 us.
static long long pair_penalty() {
  static long long v = -1;
  if (v < 0) {
    const char* e = getenv("VB_GEMM_PAIR_PENALTY");   // cycles; development override
    v = e ? atoll(e) : 6000;
  }
  return v;
}

// Chooses (tile width, CTAs per tile group, k splits) for a problem; honours the values the caller fixed. Pure host code.
static int choose_config(const vb_gemm_args* a, int max_ctas, int* bn_out, int* cluster_out, int* split_out) {
  const int num_m = (a->M + BM - 1) / BM;
  const int num_k = (a->K + BK - 1) / BK * (1 + (a->A_lo ? 1 : 0) + (a->B_lo ? 1 : 0));   // virtual k-blocks (split precision passes)
  // Tile configuration = (tile width bn, CTAs per tile group cg, k splits): minimise the modelled time of the busiest CTA,
  // in SM cycles, with constants measured on B200 (clock64 timelines / feed probes in profiles/):
  //   main loop per 64-deep k-block: 128x128 ~430 (bound by the SM's operand ingest, ~98 B/clk of TMA writes competing with
  //   the tensor core's smem reads), 128x256 ~650 (L2 -> SM bandwidth with all SMs pulling), CTA pair 256x128 ~400,
  //   CTA pair 256x256 ~505 (= the tcgen05 floor); the epilogue of a tile overlaps the next tile's main loop, the last one
  //   is exposed; ~3k cycles of prologue + first-load latency per launch.
  if (a->block_n != 0 && a->block_n != 128 && a->block_n != 256) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: block_n must be 0, 128 or 256");
  int cluster_req = a->cluster_m ? a->cluster_m : default_cluster();
  if (cluster_req != 0 && cluster_req != 1 && cluster_req != 2) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: cluster_m must be 0, 1 or 2");
  if (a->split_k > 1 && (!a->atomic_out || a->act != VB_ACT_NONE || a->bias || a->residual || a->out_colsum))
    return set_error(VB_ERR_INVALID, "vb_gemm_bf16: split_k > 1 needs atomic_out and a plain epilogue");
  const bool can_split = a->atomic_out && a->act == VB_ACT_NONE && !a->bias && !a->residual && !a->out_colsum;
  long long epi_base = 9000;   // generic epilogue
  if (a->atomic_out) epi_base = 3000;
  else if (a->act == VB_ACT_GELU || a->act == VB_ACT_DGELU) epi_base = 4200;
  else if (a->act == VB_ACT_NONE && a->out_f32 && !a->out_bf16) epi_base = a->residual ? 4800 : 3000;
  else if (a->act == VB_ACT_NONE && a->out_bf16 && !a->out_f32) epi_base = 3000;
  int bn = 128, cluster = 1, split_k = 1;
  {
    long long best = -1;
    static const int split_cand[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32};
    for (int w = 128; w <= 256; w += 128) {
      if (a->block_n && a->block_n != w) continue;
      if (!a->block_n && w == 256 && a->N <= 128) continue;
      for (int cg = 1; cg <= 2; ++cg) {
        if (cluster_req && cluster_req != cg && !(cluster_req == 2 && cg == 1 && (num_m < 2 || max_ctas < 2))) continue;
        if (cg == 2 && (num_m < 2 || max_ctas < 2)) continue;
        const long long t_kb = (w == 128) ? (cg == 1 ? 430 : 400) : (cg == 1 ? 650 : 505);
        const long long epi = epi_base * (w / 128);
        const long long groups = (cg == 1) ? max_ctas : max_ctas / 2;
        const long long tiles = (long long)((num_m + cg - 1) / cg) * ((a->N + w - 1) / w);
        for (int sc : split_cand) {
          if (a->split_k > 0 && sc != 1) break;
          int sp = a->split_k > 0 ? a->split_k : sc;
          if (sp > 1 && !can_split) break;
          if (sp > num_k) { if (a->split_k > 0) sp = num_k; else break; }
          const long long kps_ = (num_k + sp - 1) / sp;
          sp = (int)((num_k + kps_ - 1) / kps_);   // no empty splits
          const long long rounds = (tiles * sp + groups - 1) / groups;
          const long long ml = kps_ * t_kb;
          const long long c = 3000 + (cg == 2 ? pair_penalty() : 0) + rounds * (ml > epi ? ml : epi) + epi;
          if (best < 0 || c < best) { best = c; bn = w; cluster = cg; split_k = sp; }
        }
      }
    }
    if (best < 0) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: no tile configuration for block_n=%d cluster_m=%d split_k=%d", a->block_n, a->cluster_m, a->split_k);
  }
  *bn_out = bn; *cluster_out = cluster; *split_out = split_k;
  return VB_OK;
}

}  // namespace vb

extern "C" vb_status vb_gemm_bf16(const vb_gemm_args* a, void* stream_) {
  using namespace vb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: null args");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: empty problem %dx%dx%d", a->M, a->N, a->K);
  if (!a->A || !a->B) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: null operand");
  if ((a->lda % 8) || (a->ldb % 8) || !aligned(a->A, 16) || !aligned(a->B, 16))
    return set_error(VB_ERR_INVALID, "vb_gemm_bf16: operands need ld %% 8 == 0 and 16-byte aligned bases (lda=%lld ldb=%lld)",
                     (long long)a->lda, (long long)a->ldb);
  if (!a->out_f32 && !a->out_bf16) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: no output");
  if (a->act == VB_ACT_DGELU && !a->aux) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: DGELU needs aux");
  if (a->bias && !aligned(a->bias, 16)) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: bias must be 16-byte aligned");
  if (a->atomic_out && (!a->out_f32 || a->out_bf16 || a->out_pre))
    return set_error(VB_ERR_INVALID, "vb_gemm_bf16: atomic_out supports only out_f32");
  if ((a->A_lo && !aligned(a->A_lo, 16)) || (a->B_lo && !aligned(a->B_lo, 16)))
    return set_error(VB_ERR_INVALID, "vb_gemm_bf16: A_lo / B_lo must be 16-byte aligned");
  if ((a->out_lo || a->out_b16) && !a->out_bf16) return set_error(VB_ERR_INVALID, "vb_gemm_bf16: out_lo / out_b16 need out_bf16");
  if ((a->a_fp16 != 0) != (a->b_fp16 != 0))
    return set_error(VB_ERR_UNSUPPORTED, "vb_gemm_bf16: A and B must have the same 16-bit format (fp16 x bf16 faults on sm_100)");
  int dev_sms = 0, cc = 0;
  if (int s = vb_device_info(&dev_sms, &cc)) return s;
  if (cc / 10 != 10) return set_error(VB_ERR_UNSUPPORTED, "vb_gemm_bf16: needs an sm_100 device (found sm_%d)", cc);

  const int num_m = (a->M + BM - 1) / BM;
  const int real_k = (a->K + BK - 1) / BK;
  const int npass = 1 + (a->A_lo ? 1 : 0) + (a->B_lo ? 1 : 0);
  const int num_k = real_k * npass;   // virtual k-blocks (split precision: K is walked once per pass)
  int max_ctas = a->max_ctas > 0 ? a->max_ctas : dev_sms;

  int bn = 128, cluster = 1, split_k = 1;
  if (int st = choose_config(a, max_ctas, &bn, &cluster, &split_k)) return st;
  const int num_n = (a->N + bn - 1) / bn;
  const int kps = (num_k + split_k - 1) / split_k;

  GemmKernelParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_m_blocks = num_m; p.num_n_blocks = num_n; p.num_k_blocks = num_k;
  p.real_k_blocks = real_k; p.npass = npass;
  p.pass_sel[0] = 0; p.pass_sel[1] = a->A_lo ? 1 : 2; p.pass_sel[2] = 2;
  p.out_fp16 = a->out_fp16 ? 1 : 0;
  p.out_lo = static_cast<__nv_bfloat16*>(a->out_lo);
  p.out_b16 = static_cast<__nv_bfloat16*>(a->out_b16);
  p.split_k = split_k; p.k_blocks_per_split = kps;
  p.alpha = a->alpha;
  p.bias = a->bias;
  p.residual = a->residual; p.ld_res = a->ld_res;
  p.aux = static_cast<const __nv_bfloat16*>(a->aux); p.ld_aux = a->ld_aux;
  p.act = a->act;
  p.out_f32 = a->out_f32; p.ld_of = a->ld_out_f32;
  p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16); p.ld_ob = a->ld_out_bf16;
  p.out_pre = static_cast<__nv_bfloat16*>(a->out_pre); p.ld_op = a->ld_out_pre;
  p.atomic_out = a->atomic_out;
  p.out_colsum = a->out_colsum;
  p.vec_f32 = a->out_f32 && aligned(a->out_f32, 16) && (a->ld_out_f32 % 4 == 0);
  p.vec_bf16 = a->out_bf16 && aligned(a->out_bf16, 8) && (a->ld_out_bf16 % 4 == 0) && (!a->out_lo || aligned(a->out_lo, 8)) &&
               (!a->out_b16 || aligned(a->out_b16, 8));
  p.vec_pre = a->out_pre && aligned(a->out_pre, 8) && (a->ld_out_pre % 4 == 0);
  p.vec_res = a->residual && aligned(a->residual, 16) && (a->ld_res % 4 == 0);
  p.vec_aux = a->aux && aligned(a->aux, 8) && (a->ld_aux % 4 == 0);

  // smem matrix descriptors (see vb_ptx.cuh). K-major: rows of 128 B, 8-row groups 1024 B apart (SBO),
  // K advance of 16 elements = 32 B inside the swizzle span. MN-major: 64-element (128 B) rows indexed
  // by k, 8-k groups 1024 B apart (SBO), next 64 MN elements BK*128 B further (LBO); K advance of 16 = 2 groups.
  const uint32_t lbo_a = a->dbg_lbo_a ? a->dbg_lbo_a : (a->a_mn_major ? BK * 128 : 16);
  const uint32_t sbo_a = a->dbg_sbo_a ? a->dbg_sbo_a : 1024;
  const uint32_t lbo_b = a->dbg_lbo_b ? a->dbg_lbo_b : (a->b_mn_major ? BK * 128 : 16);
  const uint32_t sbo_b = a->dbg_sbo_b ? a->dbg_sbo_b : 1024;
  p.desc_base_a = umma_desc_base(lbo_a, sbo_a);
  p.desc_base_b = umma_desc_base(lbo_b, sbo_b);
  p.kadv_a = a->a_mn_major ? 2 * 1024 : UK * 2;
  p.kadv_b = a->b_mn_major ? 2 * 1024 : UK * 2;
  p.idesc = umma_idesc_bf16(BM * cluster, bn, a->a_mn_major ? 1 : 0, a->b_mn_major ? 1 : 0, a->a_fp16, a->b_fp16);   // pairs: 256 x bn MMAs
  p.dbg = reinterpret_cast<unsigned long long*>(a->dbg_timeline);
  p.a_mn = a->a_mn_major ? 1 : 0;
  p.b_mn = a->b_mn_major ? 1 : 0;
  p.cluster = cluster;
  p.num_m_groups = (num_m + cluster - 1) / cluster;
  p.fast_ok = 0;
  p.drop.ctr = (a->dropout.step && a->dropout.p > 0.f) ? a->dropout.step : nullptr;
  p.drop.site = a->dropout.site;
  p.drop.thresh = (uint32_t)((double)a->dropout.p * 4294967296.0);
  p.drop.scale = a->dropout.p < 1.f ? 1.f / (1.f - a->dropout.p) : 0.f;

  CUtensorMap tm[4];   // A, B, A_lo, B_lo (the lo maps alias the hi ones when a low part is absent)
  int st;
  for (int i = 0; i < 4; ++i) {
    const bool is_a = (i & 1) == 0;
    const void* ptr = is_a ? (i < 2 ? a->A : a->A_lo) : (i < 2 ? a->B : a->B_lo);
    if (!ptr) { tm[i] = tm[i - 2]; continue; }
    if (is_a) {
      if (a->a_mn_major) st = make_tmap(&tm[i], ptr, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BK);
      else               st = make_tmap(&tm[i], ptr, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BK, BM);
    } else {
      if (a->b_mn_major) st = make_tmap(&tm[i], ptr, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK);
      else               st = make_tmap(&tm[i], ptr, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BK, (uint32_t)(bn / cluster));   // a pair CTA stages half the tile
    }
    if (st) return st;
  }

  const long long total_work = (long long)p.num_m_groups * num_n * split_k;
  // pick the epilogue specialisation; anything unusual runs the generic one
  int epi = EPI_GENERIC;
  const bool no_extra = !a->out_colsum;
  const bool has_drop = p.drop.ctr != nullptr;   // only the F32 specialisation (and the generic path) implement it
  if (a->atomic_out) {
    if (a->act == VB_ACT_NONE && !a->bias && !a->residual && no_extra && !has_drop) { epi = EPI_ATOMIC; p.fast_ok = p.vec_f32; }
  } else if (a->act == VB_ACT_GELU) {
    if (a->out_pre && !a->residual && no_extra && !has_drop && (a->out_bf16 || a->out_f32)) {
      epi = EPI_GELU; p.fast_ok = p.vec_pre && (!a->out_bf16 || p.vec_bf16) && (!a->out_f32 || p.vec_f32);
    }
  } else if (a->act == VB_ACT_DGELU) {
    if (a->out_bf16 && !a->out_f32 && !a->residual && !a->bias && !has_drop) { epi = EPI_DGELU; p.fast_ok = p.vec_bf16 && p.vec_aux; }
  } else if (a->act == VB_ACT_RELU) {
    if (a->out_f32 && no_extra && !a->residual && !has_drop) { epi = EPI_F32; p.fast_ok = p.vec_f32 && (!a->out_bf16 || p.vec_bf16); }   // poolers: fp32 + operand copy
  } else if (a->act == VB_ACT_NONE) {
    if (a->out_f32 && !a->out_bf16 && no_extra) { epi = EPI_F32; p.fast_ok = 1; }   // unaligned pitches use 32-bit accesses
    else if (a->out_bf16 && !a->out_f32 && !a->residual && no_extra && !has_drop) { epi = EPI_BF16; p.fast_ok = p.vec_bf16; }
  }
  // 16-bit output variant of the BF16 / GELU specialisations; split precision (out_lo) and other combinations run the generic epilogue
  int out16 = 0;
  if (epi == EPI_BF16 || epi == EPI_GELU) {
    if ((a->out_lo || a->out_b16) && !a->out_fp16) epi = EPI_GENERIC;     // bf16 hi + lo: not a combination the engine uses
    else out16 = a->out_lo ? 3 : (a->out_fp16 ? (a->out_b16 ? 2 : 1) : 0);
  } else if (epi == EPI_DGELU && (a->out_fp16 || a->out_lo || a->out_b16)) {
    epi = EPI_GENERIC;
  }
  if (cluster == 2) {
    if (bn == 256) return launch_gemm_epi<256, 2>(epi, out16, tm, p, total_work, max_ctas, stream);
    return launch_gemm_epi<128, 2>(epi, out16, tm, p, total_work, max_ctas, stream);
  }
  if (bn == 256) return launch_gemm_epi<256, 1>(epi, out16, tm, p, total_work, max_ctas, stream);
  return launch_gemm_epi<128, 1>(epi, out16, tm, p, total_work, max_ctas, stream);
}

extern "C" vb_status vb_gemm_plan(const vb_gemm_args* a, int32_t sm_count_, int32_t* block_n, int32_t* cluster_m, int32_t* split_k) {
  using namespace vb;
  if (!a || !block_n || !cluster_m || !split_k) return set_error(VB_ERR_INVALID, "vb_gemm_plan: null argument");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error(VB_ERR_INVALID, "vb_gemm_plan: empty problem %dx%dx%d", a->M, a->N, a->K);
  int sms = sm_count_;
  if (sms <= 0) {
    int cc = 0;
    if (int s = vb_device_info(&sms, &cc)) return s;
  }
  const int max_ctas = a->max_ctas > 0 ? a->max_ctas : sms;
  int bn, cl, sp;
  if (int s = choose_config(a, max_ctas, &bn, &cl, &sp)) return s;
  *block_n = bn; *cluster_m = cl; *split_k = sp;
  return VB_OK;
}
