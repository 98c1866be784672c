/*
 * vilbert_b200.h — C ABI of libvilbert_b200.so: the sm_100a kernels behind the ViLBERT two-stream
 * co-attentional encoder hot path (reference: vilbert/vilbert.py:396-1107 BertLayer / BertImageLayer /
 * BertConnectionLayer / BertEncoder, :320-367 + :1409-1432 embeddings, :1110-1137 poolers,
 * :1140-1258 + :1638-1722 heads).
 *
 * The reference has no FFI of its own (it is pure PyTorch); every entry point below
 * cites the reference nn.Module / expression whose arithmetic it replaces. The Python host
 * (vilbert-multi-task_b200/) binds these with ctypes — see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - the caller owns every buffer; the library never allocates or frees device memory;
 *   - every launch is asynchronous on the caller-supplied stream (cudaStream_t passed as void*);
 *   - every function returns a vb_status (0 = ok); vb_last_error() gives a message for the calling
 *     thread; no C++ exception crosses the boundary;
 *   - "bf16" buffers are raw uint16 bfloat16; "f32" are IEEE binary32; matrices are row-major with
 *     an explicit leading dimension in ELEMENTS.
 */
#ifndef VILBERT_B200_H_
#define VILBERT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int vb_status;
enum {
  VB_OK = 0,
  VB_ERR_INVALID = 1,     /* bad shape / alignment / argument                     */
  VB_ERR_UNSUPPORTED = 2, /* device is not sm_100 or an unsupported configuration */
  VB_ERR_CUDA = 3         /* a CUDA runtime / driver call failed                  */
};

enum { VB_ACT_NONE = 0, VB_ACT_GELU = 1, VB_ACT_RELU = 2, VB_ACT_DGELU = 3 };

/* ABI version of this header (bumped on incompatible change). */
int vb_version(void);
/* Message for the last non-OK status returned on this thread ("" if none). */
const char* vb_last_error(void);
/* Number of SMs / compute capability (major*10+minor) of the current device. */
vb_status vb_device_info(int* sm_count, int* cc);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction on the tcgen05 tensor cores (TMA -> 128B-swizzled smem -> tcgen05.mma -> TMEM).
 *   D[M,N] = alpha * sum_k A(m,k) * B(n,k)   followed by the fused epilogue
 *   v = D (+ bias[n]); act; (+ residual[m,n]); -> out_f32 / out_bf16
 * Replaces every nn.Linear on the path (vilbert.py:410-412,466,492,509 text; :553-555,625,653,670
 * image; :716-725,830,837,865-869 connection; :1116,1131 poolers; heads :1143,1163,1183,1250,1714-1719)
 * and their autograd (dgrad / wgrad):
 *   forward   y = x W^T + b          A = x  [M,K] k-major,  B = W  [N,K] k-major
 *   dgrad     dx = dy W              A = dy [M,N'] k-major, B = W  [N',K'] stored [red, out] -> b_mn_major
 *   wgrad     dW = dy^T x            A = dy [red, out] -> a_mn_major, B = x [red, in] -> b_mn_major
 * Operand storage:
 *   k-major  : element (row r of the M/N extent, reduction index k) at ptr[r*ld + k]
 *   mn-major : element (r, k) at ptr[k*ld + r]
 * Requirements: ld % 8 == 0, base pointers 16-byte aligned. M, N, K arbitrary (TMA zero-fills the
 * edges, stores are predicated).
 * act: VB_ACT_GELU = exact erf GELU (vilbert.py:111-117), out_pre receives the pre-activation;
 *      VB_ACT_DGELU multiplies by gelu'(aux[m,n]) (aux = saved pre-activation);
 * atomic_out: accumulate into out_f32 with red.global.add (needed when split_k > 1).
 */
typedef struct vb_gemm_args {
  int32_t M, N, K;
  const void* A;      /* bf16 */
  int64_t lda;
  int32_t a_mn_major;
  const void* B;      /* bf16 */
  int64_t ldb;
  int32_t b_mn_major;
  float alpha;
  const float* bias;     /* [N] or NULL */
  const float* residual; /* f32 [M,N] or NULL; may alias out_f32 */
  int64_t ld_res;
  const void* aux;       /* bf16 [M,N] for VB_ACT_DGELU, else NULL */
  int64_t ld_aux;
  int32_t act;
  float* out_f32;        /* or NULL */
  int64_t ld_out_f32;
  void* out_bf16;        /* or NULL */
  int64_t ld_out_bf16;
  void* out_pre;         /* bf16 pre-activation (GELU) or NULL */
  int64_t ld_out_pre;
  int32_t atomic_out;    /* 0 store, 1 red.add into out_f32 */
  int32_t split_k;       /* >= 1; > 1 requires atomic_out and no act / bf16 outputs */
  int32_t block_n;       /* 0 = auto, else 128 or 256 */
  int32_t max_ctas;      /* 0 = one persistent CTA per SM */
  /* debug/test overrides for the smem matrix descriptors (0 = library default) */
  uint32_t dbg_lbo_a, dbg_sbo_a, dbg_lbo_b, dbg_sbo_b;
} vb_gemm_args;

vb_status vb_gemm_bf16(const vb_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VILBERT_B200_H_ */
