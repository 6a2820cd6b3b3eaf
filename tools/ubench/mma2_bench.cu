// Microbenchmark: issue rate of cta_group::2 (CTA-pair) TS-form UMMA, M=256 (128 rows per CTA), K=16.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_b_desc(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a), "l"(b), "r"(id), "r"(acc)
               : "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_rate(int batches, int reps, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  uint32_t phase = 0;
  const unsigned long long t0 = clock64();
  for (int it = 0; it < batches; ++it) {
    if (rank == 0 && tid == 0) {
      const uint64_t b0 = make_b_desc(smem_u32(smem));
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 12; ++k) mma2_ts(base, base + 480 + (k & 3) * 8, b0 + (uint64_t)((k & 3) * 2), idesc(256, N), 1);
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(&bar)),
                   "h"((uint16_t)3)
                   : "memory");
    }
    if (tid == 0) {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(&bar)), "r"(phase)
                     : "memory");
      }
    }
    phase ^= 1;
    __syncthreads();
  }
  if (tid == 0) out[rank] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

template <int N>
void run(unsigned long long* out) {
  cudaFuncSetAttribute(mma2_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int reps : {1, 4, 16}) {
    out[0] = out[1] = 0;
    mma2_rate<N><<<2, 128, 65536>>>(100, reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); exit(1); }
    printf("cta_group::2 M=256 N=%3d mmas/commit=%3d : %7llu cyc per batch (leader), %7llu (peer), %6.1f cyc per MMA\n", N, 12 * reps,
           out[0] / 100, out[1] / 100, (double)out[0] / (100.0 * 12 * reps));
  }
}

int main() {
  unsigned long long* out;
  cudaMallocManaged(&out, 64);
  run<32>(out);
  run<64>(out);
  run<128>(out);
  run<256>(out);
  return 0;
}
