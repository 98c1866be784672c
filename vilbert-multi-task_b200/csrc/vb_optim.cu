// vb_optim.cu — fused multi-tensor AdamW on the engine's flat buffers.
//
// Reference semantics: pytorch_transformers==1.0.0 AdamW as constructed at train_tasks.py:425-426
// (AdamW(optimizer_grouped_parameters, lr=base_lr, correct_bias=False), one param group PER TENSOR with its own
// lr / weight_decay, train_tasks.py:401-421), stepped at train_tasks.py:550-551 followed by model.zero_grad():
//     m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g g;  p -= step_size m / (sqrt(v) + eps);  p -= lr wd p
// with step_size = lr (correct_bias False) or lr sqrt(1 - b2^t) / (1 - b1^t). The decoupled weight decay is applied
// AFTER the Adam update and uses the updated p, like the reference.
//
// All parameters live in ONE flat fp32 buffer (engine.ParamStore), so the whole optimizer step is one HBM-bound
// launch: per element read p, g, m, v (16 B), write p, m, v (12 B) + the 16-bit tensor-core operand copy of the new
// weight (2 B, + 2 B low part in split precision) + the zeroed gradient (4 B). That removes the separate weight
// cast and gradient memset kernels of the training step. Work is described by a chunk table (contiguous ranges that
// do not cross tensor boundaries, each pointing at its hyper-parameter group).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vb_internal.h"
#include "vb_ptx.cuh"

namespace vb {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             uint16_t* __restrict__ p16, uint16_t* __restrict__ p16_lo, __nv_bfloat16* __restrict__ p16_b, int fp16,
             const long long* __restrict__ chunk_start,
             const int* __restrict__ chunk_count, const int* __restrict__ chunk_group, int n_chunks,
             const vb_adamw_group* __restrict__ groups, const int* __restrict__ step, float grad_scale, int zero_grad) {
  pdl_entry();
  const int t = step ? *step : 1;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const long long s0 = chunk_start[c];
    const int n = chunk_count[c];
    const vb_adamw_group G = groups[chunk_group[c]];
    float step_size = G.lr;
    if (G.correct_bias) step_size = G.lr * sqrtf(1.f - powf(G.beta2, (float)t)) / (1.f - powf(G.beta1, (float)t));
    const float decay = 1.f - G.lr * G.weight_decay;   // p <- p - lr wd p  (weight_decay > 0 only)
    const float ob1 = 1.f - G.beta1, ob2 = 1.f - G.beta2;
    auto upd = [&](float pv, float gv, float& mv, float& vv) -> float {
      gv *= grad_scale;
      mv = G.beta1 * mv + ob1 * gv;
      vv = G.beta2 * vv + ob2 * gv * gv;
      pv = pv - step_size * (mv / (sqrtf(vv) + G.eps));
      if (G.weight_decay > 0.f) pv *= decay;
      return pv;
    };
    const int n4 = n >> 2;   // chunk starts are multiples of 4 elements (tensors start on 8-element boundaries)
    float4* p4 = reinterpret_cast<float4*>(p + s0);
    float4* g4 = reinterpret_cast<float4*>(g + s0);
    float4* m4 = reinterpret_cast<float4*>(m + s0);
    float4* v4 = reinterpret_cast<float4*>(v + s0);
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 pv = p4[i], mv = m4[i], vv = v4[i];
      const float4 gv = g4[i];
      pv.x = upd(pv.x, gv.x, mv.x, vv.x); pv.y = upd(pv.y, gv.y, mv.y, vv.y);
      pv.z = upd(pv.z, gv.z, mv.z, vv.z); pv.w = upd(pv.w, gv.w, mv.w, vv.w);
      p4[i] = pv; m4[i] = mv; v4[i] = vv;
      if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p16) {
        if (p16_lo) {
          uint32_t l01, l23;
          const uint32_t h01 = pack16_split(pv.x, pv.y, fp16, l01), h23 = pack16_split(pv.z, pv.w, fp16, l23);
          reinterpret_cast<uint2*>(p16 + s0)[i] = make_uint2(h01, h23);
          reinterpret_cast<uint2*>(p16_lo + s0)[i] = make_uint2(l01, l23);
        } else {
          reinterpret_cast<uint2*>(p16 + s0)[i] = make_uint2(pack16(pv.x, pv.y, fp16), pack16(pv.z, pv.w, fp16));
        }
      }
      if (p16_b) reinterpret_cast<uint2*>(p16_b + s0)[i] = make_uint2(pack_bf16(pv.x, pv.y), pack_bf16(pv.z, pv.w));
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) {   // ragged tail of a tensor (e.g. a 3129-entry bias)
      const long long e = s0 + i;
      float mv = m[e], vv = v[e];
      const float pv = upd(p[e], g[e], mv, vv);
      p[e] = pv; m[e] = mv; v[e] = vv;
      if (zero_grad) g[e] = 0.f;
      if (p16) {
        const uint16_t hi = cvt16(pv, fp16);
        p16[e] = hi;
        if (p16_lo) p16_lo[e] = cvt16(pv - cvt16_to_f32(hi, fp16), fp16);
      }
      if (p16_b) p16_b[e] = __float2bfloat16(pv);
    }
  }
}

}  // namespace vb

extern "C" vb_status vb_adamw_step(float* p, float* g, float* m, float* v, void* p16, void* p16_lo, void* p16_b, int32_t p16_fp16,
                                   const int64_t* chunk_start, const int32_t* chunk_count, const int32_t* chunk_group, int32_t n_chunks,
                                   const vb_adamw_group* groups, const int32_t* step, float grad_scale, int32_t zero_grad, void* stream) {
  using namespace vb;
  if (n_chunks <= 0) return VB_OK;
  if (!p || !g || !m || !v || !chunk_start || !chunk_count || !chunk_group || !groups)
    return set_error(VB_ERR_INVALID, "vb_adamw_step: null argument");
  auto al = [](const void* q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  if (!al(p, 16) || !al(g, 16) || !al(m, 16) || !al(v, 16) || (p16 && !al(p16, 8)) || (p16_lo && (!al(p16_lo, 8) || !p16)) || !al(p16_b, 8))
    return set_error(VB_ERR_INVALID, "vb_adamw_step: buffers must be 16-byte aligned (16-bit copies 8-byte)");
  int grid = sm_count() * 8;
  if (grid <= 0) grid = 148 * 8;
  if (grid > n_chunks) grid = n_chunks;
  cudaError_t e = launch_pdl(adamw_kernel, dim3(grid), dim3(OPT_THREADS), (size_t)0, static_cast<cudaStream_t>(stream), p, g, m, v,
                             static_cast<uint16_t*>(p16), static_cast<uint16_t*>(p16_lo), static_cast<__nv_bfloat16*>(p16_b), (int)(p16_fp16 ? 1 : 0),
                             reinterpret_cast<const long long*>(chunk_start), chunk_count, chunk_group, (int)n_chunks, groups, step,
                             grad_scale, (int)(zero_grad ? 1 : 0));
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "vb_adamw_step: %s", cudaGetErrorString(e));
  return VB_OK;
}
