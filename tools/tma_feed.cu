// TMA -> smem -> tcgen05.mma feed-rate microbenchmark (development tool, not part of the library).
// Each CTA streams KB k-blocks of an A tile (128 rows x 64 bf16) and a B tile (BN rows x 64 bf16) through an S-stage
// mbarrier ring, exactly like the GEMM main loop, and reports cycles per k-block for:
//   mode 0: consumer releases the slot at once (pure TMA feed rate)
//   mode 1: consumer issues the 4 MMAs of the k-block and releases the slot with tcgen05.commit (the GEMM main loop)
//   mode 2: like 1 but the consumer does not wait for the data (stale operands): MMA and TMA share the SM at the MMA's
//           pace without the load-latency dependency
// as a function of S, BN and the number of CTAs. Operands live in a 32 MB buffer (L2 resident after the first pass).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vilbert-multi-task_b200/csrc tools/tma_feed.cu -o tools/_bin/tma_feed -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "vb_ptx.cuh"
using namespace vb;

struct P { int S, BN, KB, mode, rows_per_cta; long long* out; };

__global__ void __launch_bounds__(128) feed_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[8], empty_bar[8], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const int stage_bytes = 16384 + p.BN * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&done_bar), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 2) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int row0 = blockIdx.x * p.rows_per_cta;
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < p.KB; ++kb) {
        if (p.mode != 2) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        } else {   // bounded: the consumer does not wait for data in this mode and may lap the producer
          for (int spin = 0; spin < 4096 && !mbar_try_wait(smem_u32(&empty_bar[stage]), phase ^ 1); ++spin) {}
        }
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_arrive_expect_tx(fb, stage_bytes);
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        tma_load_2d(sa, &ta, (kb * 64) % 8192, row0 % 2048, fb);
        tma_load_2d(sa + 16384, &tb, (kb * 64) % 8192, (row0 * 2) % 1024, fb);
        if (++stage == p.S) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    long long t0 = 0;
    if (elect_one()) {
      const uint64_t base = umma_desc_base(16, 1024);
      const uint32_t idesc = umma_idesc_bf16(128, p.BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < p.KB; ++kb) {
        if (p.mode != 2) mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        if (kb == 0) t0 = clock64();
        if (p.mode == 0) {
          mbar_arrive(smem_u32(&empty_bar[stage]));
        } else {
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem, umma_desc_at(base, sa + k * 32), umma_desc_at(base, sa + 16384 + k * 32), idesc, 1u);
          umma_commit(smem_u32(&empty_bar[stage]));
        }
        if (++stage == p.S) { stage = 0; phase ^= 1; }
      }
      umma_commit(smem_u32(&done_bar));
      mbar_wait(smem_u32(&done_bar), 0);
      p.out[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fnp;
  // A: [2048 + 128 rows][8192] bf16 (32 MB), B: [1024 + 256 rows][8192]
  void *A, *B; cudaMalloc(&A, (size_t)2304 * 8192 * 2); cudaMalloc(&B, (size_t)1536 * 8192 * 2);
  cudaMemset(A, 0, (size_t)2304 * 8192 * 2); cudaMemset(B, 0, (size_t)1536 * 8192 * 2);
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
  auto mk = [&](CUtensorMap* tm, void* ptr, uint64_t rows, uint32_t box_rows) {
    cuuint64_t dims[2] = {8192, rows}; cuuint64_t strides[1] = {8192 * 2}; cuuint32_t box[2] = {64, box_rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  };
  const int KB = 1024;
  printf("cycles per k-block (A 128x64 + B BNx64 bf16 per stage), %d k-blocks, median-ish = CTA 0 / max over CTAs\n", KB);
  for (int BN : {128, 256}) {
    CUtensorMap ta, tb; mk(&ta, A, 2304, 128); mk(&tb, B, 1536, (uint32_t)BN);
    const int stage_bytes = 16384 + BN * 128;
    for (int mode : {0, 1, 2}) {
      for (int S = 2; S <= 6; ++S) {
        if (S * stage_bytes > 200 * 1024) continue;
        for (int ctas : {1, 148}) {
          P p{S, BN, KB, mode, 128, d};
          long long h[148];
          for (int rep = 0; rep < 2; ++rep) {
            feed_kernel<<<ctas, 128, S * stage_bytes>>>(ta, tb, p);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, d, ctas * 8, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("  BN=%3d mode=%d stages=%d ctas=%3d: %6.0f / %6.0f   (%5.1f B/clk/SM)\n", BN, mode, S, ctas, double(h[0]) / KB, double(mx) / KB,
                 stage_bytes / (double(mx) / KB));
        }
      }
    }
  }
  return 0;
}
