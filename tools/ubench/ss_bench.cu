// Microbenchmark: issue rate of SS-form tcgen05.mma (both operands in shared memory, NO-swizzle K-major "interleaved"
// layout: rows of 16 bytes, 8 rows = one 128-byte core matrix) for the operand shapes of the image-encoder trunk.
//   orientation 1 (k_enc_trunk): A = 128 positions (activation window), B = N weight rows (N = 32 / 64 / 96 / 192)
//   orientation 2 (swapped):     A = 128 (or 64) weight rows, B = N positions (N = 128 / 224 / 256)
// Prints cycles per MMA for M in {64, 128} x several N, with the LBO / SBO strides the kernels use.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_il(uint32_t a, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(id), "r"(acc)
               : "memory");
}

// a_lbo / b_lbo: byte distance between the two 16-byte K chunks of a K = 16 step (plane stride); SBO = 128
__global__ void __launch_bounds__(128, 1) ss_rate(int M, int N, uint32_t a_lbo, uint32_t b_lbo, int batches, int per_batch,
                                                  unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t id = idesc(M, N);
  uint32_t phase = 0;
  const unsigned long long t0 = clock64();
  for (int it = 0; it < batches; ++it) {
    if (warp == 0) {
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      if (pred) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
        for (int k = 0; k < per_batch; ++k) {
          // walk the operands like the kernels do: a few distinct start addresses, 16-byte granular shifts
          const uint64_t da = desc_il(a0 + (uint32_t)(k % 6) * 3072, a_lbo, 128);
          const uint64_t db = desc_il(b0 + (uint32_t)(k % 3) * 34 * 16 + (uint32_t)(k % 2) * 16, b_lbo, 128);
          mma_ss(base + (uint32_t)(k & 1) * 256, da, db, id, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
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
  if (tid == 0) out[0] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

static void run(const char* what, int M, int N, uint32_t a_lbo, uint32_t b_lbo, unsigned long long* out, int grid) {
  const int batches = 50, per = 72;
  out[0] = 0;
  ss_rate<<<grid, 128, 200 * 1024>>>(M, N, a_lbo, b_lbo, batches, per, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s M=%d N=%d: %s\n", what, M, N, cudaGetErrorString(e)); exit(1); }
  printf("%-34s M=%3d N=%3d grid=%3d : %6.1f cyc per MMA  (floor %5.1f)\n", what, M, N, grid,
         (double)out[0] / (batches * (double)per), 128.0 * N / 256.0);
}

int main() {
  unsigned long long* out;
  cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(ss_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {1, 148}) {
    // orientation 1: A = activation window planes (LBO = 2 x 3328), B = weights [2 NPAD rows][8] per chunk
    run("o1 A=positions B=weights", 128, 32, 2 * 3328, 32 * 16, out, grid);
    run("o1 A=positions B=weights", 128, 64, 2 * 3328, 64 * 16, out, grid);
    run("o1 dx-stacked", 128, 96, 2 * 3328, 96 * 16, out, grid);
    run("o1 dx-stacked", 128, 192, 2 * 3328, 192 * 16, out, grid);
    // orientation 2: A = weights [rows][8] per chunk (LBO = rows x 16), B = activation window planes (LBO = 2 x 4864)
    run("o2 A=weights(96 rows) B=positions", 128, 128, 96 * 16, 2 * 4864, out, grid);
    run("o2 A=weights(96 rows) B=positions", 128, 224, 96 * 16, 2 * 4864, out, grid);
    run("o2 A=weights(96 rows) B=positions", 128, 256, 96 * 16, 2 * 4864, out, grid);
    run("o2 A=weights(48 rows) B=positions", 64, 224, 48 * 16, 2 * 4864, out, grid);
    run("o2 A=weights(48 rows) B=positions", 64, 256, 48 * 16, 2 * 4864, out, grid);
  }
  return 0;
}
