// vb_attn_tc.cu — attention forward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) for the sequence lengths of
// ViLBERT's standard tasks (Nq, Nk <= 128: 36-38 tokens, 100-101 regions; head dim 64 or 128). Same math as
// vb_attn.cu::attn_fwd_kernel (reference: vilbert.py:424-460, 571-619, 771-809): P = softmax(Q K^T * scale + mask[b, key]),
// O = dropout(P) V, heads read in place from the packed QKV buffer and merged in the output; row log-sum-exp saved.
//
// One persistent CTA per SM walks (batch, head) problems:
//   warp 0      TMA producer: Q, K ([rows x 64] SWIZZLE_128B boxes, K-major operands) and V ([64 keys x 64] boxes, MN-major B
//               operand) of problem i+1 land in the other half of a 2-stage shared-memory ring while problem i computes
//               (3-D tensor maps [batch][row][column]: rows past the sequence are zero-filled, never the next sample's);
//   warps 1..4  one thread per query row (= TMEM lane). An elected thread issues S = Q K^T as tcgen05.mma (128 x Nk x D, fp32
//               accumulator in TMEM); every thread reads its row with tcgen05.ld (32 columns at a time), takes the row maximum,
//               exponentiates, sums, applies the dropout mask and writes P as a 16-bit K-major SWIZZLE_128B operand into shared
//               memory; the elected thread issues O = P V (128 x D x Nk) into a second TMEM region; the rows are scaled by
//               1 / sum, converted and stored. S and P never leave the SM.
// Longer sequences (the 12-in-1 mix's 306 x 257), head dims 16 / 32 and the split-precision mode stay on the mma.sync kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>

#include "vb_internal.h"
#include "vb_ptx.cuh"

namespace vb {

constexpr float TC_LOG2E = 1.4426950408889634f;
constexpr int TC_THREADS = 160;   // warp 0 producer + 4 compute warps

struct AttnTcParams {
  int B, H, Nq, Nk;
  const float* mask;
  float scale;
  uint16_t* O; long long ldo;
  uint16_t* Ob;          // optional always-bf16 copy of O
  float* lse;            // [B, H, Nq] log2 domain, or NULL
  DropCfg drop;
  int fp16;              // format of Q / K / V / O
  uint32_t idesc_s, idesc_o;
  int nk16;              // Nk rounded up to a multiple of 16 (UMMA N of S, number of 16-key steps of O)
};

// 3-D tiled load: coordinates (column, row, batch).
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int D>
struct TcCfg {
  static constexpr int SPANS = D / 64;                 // 64-element (128-byte) column spans of a head
  static constexpr int QK_BYTES = SPANS * 128 * 128;   // Q or K: SPANS tiles of [128 rows][128 B]
  static constexpr int V_BYTES = 2 * SPANS * 64 * 128; // V: 2 key blocks x SPANS tiles of [64 keys][128 B]
  static constexpr int STAGE_BYTES = 2 * QK_BYTES + V_BYTES;
  static constexpr int P_BYTES = 2 * 128 * 128;        // P: 2 key spans of [128 rows][128 B]
  static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + P_BYTES + 2 * 128 * 4 /*mask*/ + 128 /*barriers*/ + 1024 /*alignment slack*/;
  static constexpr int TMEM_COLS = 256;                // S: columns [0, 128), O: [128, 128 + D)
};

template <int D>
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  using Cfg = TcCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem + 2 * Cfg::STAGE_BYTES;
  float* sMask = reinterpret_cast<float*>(sP + Cfg::P_BYTES);          // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + 2 * 128);
  uint64_t* full_bar = bars;          // [2] operands of a stage landed
  uint64_t* empty_bar = bars + 2;     // [2] the MMAs reading a stage retired
  uint64_t* s_bar = bars + 4;         // S accumulator complete
  uint64_t* o_bar = bars + 5;         // O accumulator complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(s_bar), 1);
    mbar_init(smem_u32(o_bar), 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp_idx == 1) { tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_entry();

  const int n_items = p.B * p.H;
  if (warp_idx == 0) {
    // ================================================================ TMA producer
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int b = w / p.H, h = w % p.H;
      const int stage = it & 1;
      if (elect_one()) {
        mbar_wait(smem_u32(&empty_bar[stage]), ((it >> 1) & 1) ^ 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES);
        const uint32_t sQ = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sK = sQ + Cfg::QK_BYTES;
        const uint32_t sV = sK + Cfg::QK_BYTES;
#pragma unroll
        for (int j = 0; j < Cfg::SPANS; ++j) {
          tma_load_3d(sQ + j * (128 * 128), &tmQ, h * D + j * 64, 0, b, fb);
          tma_load_3d(sK + j * (128 * 128), &tmK, h * D + j * 64, 0, b, fb);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
            tma_load_3d(sV + (kb * Cfg::SPANS + j) * (64 * 128), &tmV, h * D + j * 64, kb * 64, b, fb);
        }
      }
      __syncwarp();
    }
  } else {
    // ================================================================ compute: thread = query row = TMEM lane
    const int row = (warp_idx & 3) * 32 + lane;
    const int tid = (warp_idx - 1) * 32 + lane;         // 0..127 within the compute group
    const uint32_t taddr = tmem_base + (uint32_t((warp_idx & 3) * 32) << 16);
    const bool issuer = (warp_idx == 1);
    const float c = p.scale * TC_LOG2E;
    const uint64_t desc_k = umma_desc_base(16, 1024);            // K-major SWIZZLE_128B (Q, K, P)
    const uint64_t desc_v = umma_desc_base(64 * 128, 1024);      // MN-major SWIZZLE_128B (V): 64-column sub-tiles 8 KB apart
    const uint32_t dseed = p.drop.ctr ? drop_seed(p.drop) : 0u;
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int b = w / p.H, h = w % p.H;
      const int stage = it & 1;
      const uint32_t ph = it & 1;
      float* mk = sMask + stage * 128;
      mk[tid] = (tid < p.Nk) ? (p.mask ? p.mask[(long long)b * p.Nk + tid] * TC_LOG2E : 0.f) : -CUDART_INF_F;
      const uint32_t sQ = smem_u32(smem + stage * Cfg::STAGE_BYTES);
      const uint32_t sK = sQ + Cfg::QK_BYTES;
      const uint32_t sV = sK + Cfg::QK_BYTES;
      compute_bar();   // mask staged; every thread is done with the previous problem's TMEM regions and P tile
      // ---- S = Q K^T
      if (issuer) {
        if (elect_one()) {
          mbar_wait(smem_u32(&full_bar[stage]), (it >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < Cfg::SPANS; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base, umma_desc_at(desc_k, sQ + j * (128 * 128) + k * 32), umma_desc_at(desc_k, sK + j * (128 * 128) + k * 32),
                        p.idesc_s, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(s_bar));
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(s_bar), ph);
      tc_fence_after();
      // ---- row softmax: pass 1 = maximum, pass 2 = exponentials, row sum, dropout, P -> shared memory (16-bit, swizzled)
      const int nchunk = (p.nk16 + 31) / 32;
      float mx = -CUDART_INF_F;
      for (int ch = 0; ch < nchunk; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + ch * 32, r);
        tmem_ld_wait();
        // columns >= nk16 of the last chunk were never written by the MMA (stale TMEM bits, possibly NaN): select, do not compute
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (ch * 32 + j < p.nk16) ? __uint_as_float(r[j]) * c + mk[ch * 32 + j] : -CUDART_INF_F);
      }
      float l = 0.f;
      const uint32_t e_row = (uint32_t)((((long long)b * p.H + h) * p.Nq + row) * p.Nk);
      for (int ch = 0; ch < nchunk; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + ch * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float p0 = (ch * 32 + j < p.nk16) ? exp2f(__uint_as_float(r[j]) * c + mk[ch * 32 + j] - mx) : 0.f;       // nk16 is even
          float p1 = (ch * 32 + j < p.nk16) ? exp2f(__uint_as_float(r[j + 1]) * c + mk[ch * 32 + j + 1] - mx) : 0.f;
          l += p0 + p1;
          if (p.drop.ctr) {   // nn.Dropout on the probabilities; the row sum stays undropped
            p0 *= drop_factor(dseed, e_row + ch * 32 + j, p.drop);
            p1 *= drop_factor(dseed, e_row + ch * 32 + j + 1, p.drop);
          }
          pk[j >> 1] = pack16(p0, p1, p.fp16);
        }
        // keys [32 ch, +32) of row `row`: 16-byte chunks 4 (ch & 1) .. +4 of the 128-byte row in key span ch >> 1
        uint8_t* prow = sP + (ch >> 1) * (128 * 128) + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((ch & 1) * 4 + q) ^ (row & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      fence_proxy_async();    // generic-proxy writes of P -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      compute_bar();
      // ---- O = P V
      if (issuer) {
        if (elect_one()) {
          tc_fence_after();
          const int ksteps = p.nk16 / 16;
          for (int ks = 0; ks < ksteps; ++ks) {
            const int kb = ks >> 2, k = ks & 3;
            umma_bf16(tmem_base + 128, umma_desc_at(desc_k, smem_u32(sP) + kb * (128 * 128) + k * 32),
                      umma_desc_at(desc_v, sV + kb * Cfg::SPANS * (64 * 128) + k * 2048), p.idesc_o, ks > 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(o_bar));
          umma_commit(smem_u32(&empty_bar[stage]));   // Q / K / V of this stage are free once these MMAs retire
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(o_bar), ph);
      tc_fence_after();
      // ---- epilogue: O row / l -> 16-bit, heads merged in place
      const float inv = 1.f / l;
      const bool live = row < p.Nq;
      uint16_t* orow = p.O + ((long long)b * p.Nq + row) * p.ldo + h * D;
      uint16_t* brow = p.Ob ? p.Ob + ((long long)b * p.Nq + row) * p.ldo + h * D : nullptr;
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + 128 + ch * 32, r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t o4[4], b4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float v0 = __uint_as_float(r[q * 8 + 2 * j]) * inv, v1 = __uint_as_float(r[q * 8 + 2 * j + 1]) * inv;
              o4[j] = pack16(v0, v1, p.fp16);
              b4[j] = pack_bf16(v0, v1);
            }
            *reinterpret_cast<uint4*>(orow + ch * 32 + q * 8) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            if (brow) *reinterpret_cast<uint4*>(brow + ch * 32 + q * 8) = make_uint4(b4[0], b4[1], b4[2], b4[3]);
          }
        }
      }
      if (p.lse && live) p.lse[((long long)b * p.H + h) * p.Nq + row] = mx + log2f(l);
      tc_fence_before();
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled_t tc_encode_fn() {
  static PFN_encodeTiled_t fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled_t>(ptr);
  }
  return fn;
}

// [B][N][cols] view of a head-packed activation: element (b, i, col) at base[(b*N + i)*ld + col]; box = 64 columns x box_rows rows.
static bool make_tmap3(CUtensorMap* tm, const void* base, int cols, int N, int B, long long ld, int box_rows) {
  PFN_encodeTiled_t enc = tc_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)N * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool attn_fwd_tc_eligible(const vb_attn_args* a) {
  // Opt-in (VB_ATTN_TC=1): measured on B200 (profiles/r02_attn_fwd_probe_*.log) the kernel is faster than the mma.sync one only on
  // the 100 x 100 x 128 image self-attention (24.0 vs 30.0 us) and slower on the short text-query shapes (17.8 vs 8.8 us at
  // 36 x 36 x 64), where one 128-row tcgen05 tile per problem is mostly padding and the per-problem latency chain (TMA -> MMA ->
  // softmax -> MMA -> store) is not hidden by the 2.6 - 5 problems an SM gets; inside the training step it costs 0.25 ms.
  static const bool on = getenv("VB_ATTN_TC") && getenv("VB_ATTN_TC")[0] == '1';
  if (!on) return false;
  if (a->Q_lo || a->K_lo || a->V_lo || a->O_lo) return false;
  if (a->Nq > 128 || a->Nk > 128 || (a->D != 64 && a->D != 128)) return false;
  return true;
}

template <int D>
static int launch_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p, cudaStream_t st) {
  auto kern = attn_fwd_tc_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<D>::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_attention_fwd(tc): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int grid = sm_count();
  if (grid <= 0) grid = 148;
  if (grid > p.B * p.H) grid = p.B * p.H;
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(TC_THREADS), (size_t)TcCfg<D>::SMEM_BYTES, st, tq, tk, tv, p);
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_attention_fwd(tc): %s", cudaGetErrorString(e));
  return VB_OK;
}

int attn_fwd_tc_launch(const vb_attn_args* a, cudaStream_t st) {
  AttnTcParams p;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk;
  p.mask = a->mask; p.scale = a->scale;
  p.O = static_cast<uint16_t*>(a->O); p.ldo = a->ldo;
  p.Ob = static_cast<uint16_t*>(a->O_b16);
  p.lse = a->lse;
  const bool on = a->dropout.step && a->dropout.p > 0.f;
  p.drop.ctr = on ? a->dropout.step : nullptr;
  p.drop.site = a->dropout.site;
  p.drop.thresh = on ? (uint32_t)((double)a->dropout.p * 4294967296.0) : 0u;
  p.drop.scale = on && a->dropout.p < 1.f ? 1.f / (1.f - a->dropout.p) : 1.f;
  p.fp16 = a->qkv_fp16 ? 1 : 0;
  p.nk16 = (a->Nk + 15) / 16 * 16;
  p.idesc_s = umma_idesc_bf16(128, p.nk16, 0, 0, p.fp16, p.fp16);
  p.idesc_o = umma_idesc_bf16(128, a->D, 0, 1, p.fp16, p.fp16);
  CUtensorMap tq, tk, tv;
  const int cols = a->H * a->D;
  if (!make_tmap3(&tq, a->Q, cols, a->Nq, a->B, a->ldq, 128) || !make_tmap3(&tk, a->K, cols, a->Nk, a->B, a->ldk, 128) ||
      !make_tmap3(&tv, a->V, cols, a->Nk, a->B, a->ldv, 64))
    return set_error(VB_ERR_CUDA, "vb_attention_fwd(tc): cuTensorMapEncodeTiled failed");
  if (a->D == 64) return launch_tc<64>(tq, tk, tv, p, st);
  return launch_tc<128>(tq, tk, tv, p, st);
}

}  // namespace vb
