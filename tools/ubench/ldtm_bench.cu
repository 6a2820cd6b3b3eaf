// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps (B200).
// Every warp reads (writes) its own 32-lane quadrant; R back-to-back instructions, then tcgen05.wait.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// MODE 0: loads only (4 x ld.x16 = one 64-column fp32 row block per iteration, like one epilogue)
// MODE 1: stores only (8 x st.x8 = 64 columns)
// MODE 2: the epilogue's mix: 4 x ld.x16 + 8 x st.x8 per iteration
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_tmem(int iters, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t d[4][16], acc = 0;
  uint32_t s[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16(base + c * 16, d[c]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 4; ++c) acc += d[c][0] ^ d[c][15];
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int c = 0; c < 8; ++c) tmem_st8(base + 64 + c * 8, s);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const unsigned long long t1 = clock64();
  __syncthreads();
  if (tid == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
  unsigned long long* out;
  uint32_t* sink;
  cudaMallocManaged(&out, 64);
  cudaMallocManaged(&sink, 64);
  const int iters = 2000;
  const char* names[3] = {"ld 4 x x16 (8 KB / 4 warps)", "st 8 x x8 (8 KB / 4 warps)", "ld 4 x x16 + st 8 x x8"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {1, 4, 8, 16}) {
      out[0] = 0;
      if (mode == 0) k_tmem<0><<<1, warps * 32>>>(iters, out, sink);
      if (mode == 1) k_tmem<1><<<1, warps * 32>>>(iters, out, sink);
      if (mode == 2) k_tmem<2><<<1, warps * 32>>>(iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
      const double cyc = (double)out[0] / iters;
      const double bytes = warps * 32.0 * 64 * 4 * (mode == 2 ? 2 : 1);
      printf("%-32s warps=%2d : %7.1f cyc / iteration, %6.1f B/cyc/SM\n", names[mode], warps, cyc, bytes / cyc);
    }
  }
  return 0;
}
