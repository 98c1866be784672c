// tcgen05.mma issue-rate microbenchmark (development tool, not part of the library).
// One CTA, one issuing thread, operands resident in shared memory (zeros), no TMA: measures cycles per
// 128 x N x 16 bf16 MMA for N in {64,128,256}, with one accumulator or alternating between two, with a commit per
// 4 MMAs (like the GEMM main loop), issued under `if (lane == 0)` or under `if (elect.sync)`: with the former ptxas wraps every
// UTCHMMA in a per-thread BRA.U.ANY loop (~83 cycles per MMA regardless of N); with the latter they issue back to back. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -I vilbert-multi-task_b200/csrc tools/mma_rate.cu -o gpurun_out/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "vb_ptx.cuh"
using namespace vb;

struct Res { long long cycles; };

template <bool ELECT>
__global__ void __launch_bounds__(128) rate_kernel(int N, int n_acc, int commit_each, int n_stage, int groups, int a_tmem_unused,
                                                    long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < n_stage * 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); fence_mbar_init(); }
  fence_proxy_async();
  __syncthreads();
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint64_t base = umma_desc_base(16, 1024);
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    long long t0 = 0, t1 = 0;
    if (ELECT ? elect_one() : (lane == 0)) {
      t0 = clock64();
      int stage = 0;
      for (int g = 0; g < groups; ++g) {
        const uint32_t sa = smem_u32(smem + stage * 49152);
        const uint32_t sb = sa + 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tmem + ((n_acc == 2) ? ((k & 1) * 256) : (n_acc == 4 ? (k * 128) : 0));
          umma_bf16(d, umma_desc_at(base, sa + k * 32), umma_desc_at(base, sb + k * 32), idesc, 1u);
        }
        if (commit_each) umma_commit(smem_u32(&bar[1]));
        if (++stage == n_stage) stage = 0;
      }
      umma_commit(smem_u32(&bar[0]));
      mbar_wait(smem_u32(&bar[0]), 0);
      t1 = clock64();
      out[0] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int groups = 512;
  printf("tcgen05.mma kind::f16 cta_group::1 M=128, SS operands, %d groups of 4 MMAs (K=16 each)\n", groups);
  for (int elect = 0; elect < 2; ++elect)
  for (int N : {64, 128, 256}) {
    for (int n_acc : {1, 2, 4}) {
      if (n_acc == 4 && N > 128) continue;
      for (int commit_each : {1}) {
        for (int n_stage : {4}) {
          long long h = 0;
          for (int rep = 0; rep < 2; ++rep) {
            if (elect) rate_kernel<true><<<1, 128, 4 * 49152>>>(N, n_acc, commit_each, n_stage, groups, 0, d);
            else rate_kernel<false><<<1, 128, 4 * 49152>>>(N, n_acc, commit_each, n_stage, groups, 0, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          }
          printf("  guard=%s N=%3d accumulators=%d commit/group=%d stages=%d: %7.1f cycles per MMA (floor %d)\n", elect ? "elect.sync" : "lane==0   ", N, n_acc, commit_each, n_stage,
                 double(h) / (groups * 4), N / 2);
        }
      }
    }
  }
  return 0;
}
