// tc_common.cuh -- tcgen05 / TMEM / mbarrier helpers and the operand-image layout shared by the forward
// (particle_chain_tc.cu) and backward (particle_chain_bwd.cu) tensor-core chain kernels.
#pragma once
#include <cuda_bf16.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace mmf {

constexpr int TC_MAX_GROUPS = 4;
constexpr int TILE_B = 64 * 128;      // one 64(N) x 64(K) bf16 operand tile: 64 rows of 128 bytes
constexpr int OUT_TILE_B = 16 * 128;  // the output layer, N padded to 16
constexpr int OUT_PAD = 16;

// ---- image of one chain as it sits in shared memory (built once by k_pack_chain_mma) --------------
//   [layer 0 hi | layer 0 lo | ... | layer L-1 hi | layer L-1 lo | out hi | out lo]   bf16, SW128 K-major
//   [in_Wt[in_dim][64] | in_b[64] | bias[L][64] (zeros for the mid layer) | out_b[16]]  fp32
// nsplit = 1: one CTA holds all 64 weight rows of a tile (cta_group::1).
// nsplit = 2: CTA-pair build (cta_group::2): CTA `rank` holds rows [32 rank, 32 rank + 32) of every tile and
//             rows [16 rank, 16 rank + 16) of the output layer, which is padded to N = 32.
// The buffer behind mmf_chain.w_mma is [nsplit=1 image | nsplit=2 rank 0 image | nsplit=2 rank 1 image].
__host__ __device__ inline int chain_layers(const ChainDev& c) { return 2 * c.n_pre + 1 + 2 * c.n_post; }
__host__ __device__ inline size_t image_tiles_bytes(const ChainDev& c, int nsplit = 1) {
  return (size_t)chain_layers(c) * 2 * (TILE_B / nsplit) + 2 * OUT_TILE_B;
}
__host__ __device__ inline size_t image_bytes(const ChainDev& c, int nsplit = 1) {
  const size_t b = image_tiles_bytes(c, nsplit) + sizeof(float) * (size_t)(c.in_dim * U + U + chain_layers(c) * U + OUT_PAD);
  return (b + 1023) & ~(size_t)1023;  // keep the concatenated images 1024-byte aligned
}
__host__ __device__ inline size_t image_offset(const ChainDev& c, int nsplit, int rank) {
  return nsplit == 1 ? 0 : image_bytes(c, 1) + (size_t)rank * image_bytes(c, 2);
}
__host__ __device__ inline size_t image_total_bytes(const ChainDev& c) { return image_bytes(c, 1) + 2 * image_bytes(c, 2); }

// byte offset of element (n, k) inside a K-major SWIZZLE_128B tile whose rows are 64 bf16 = 128 B
__host__ __device__ inline int sw128_offset(int n, int k) {
  return (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + (k & 7) * 2;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)  // suspend-time hint (ns): park instead of spinning
      : "memory");
  return ok != 0;
}
// try_wait suspends in hardware for a bounded time; the spin bound turns a lost arrival into a trap
// (a CUDA error the host sees) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity, hint_ns)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// true in exactly one (converged) lane of the warp; lets ptxas keep the MMA operands in uniform registers
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void group_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// CTA-pair form: one instruction covers 256 rows (128 per CTA); issued by the leader CTA only.
__device__ __forceinline__ void mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `target_rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t target_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(target_rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();
  }
}

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

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1):
// start address >> 4 | LBO (16 B) << 16 | SBO (8 rows x 128 B = 1024 B) << 32 | version << 46 | layout 2 << 61
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// bf16 split of 8 fp32 pairs, stored as the next A operand (columns [8*chunk, 8*chunk+8) of the hi and lo
// regions):  hi = rz_bf16(v) (truncation == the top 16 bits of v), lo = rn_bf16(v - hi); v - hi is exact.
// The remainder is formed by the mixed-precision FMA (PTX fma.rn.f32.bf16 -> FHFMA.BF16), which reads the packed
// bf16 halves in place: 4 instructions per pair (F2FP, 2 x FHFMA, F2FP), only the two conversions on the
// half-rate ALU pipe (ncu: the ALU pipe was the busiest pipe of the epilogue with the mask-and-subtract form).
// With RELU the activation is folded into the two conversions: v < 0 gives hi = 0, then v - 0 < 0 gives lo = 0.
template <bool RELU>
__device__ __forceinline__ void store_a_chunk(const float2 (&v)[8], uint32_t tAhi, uint32_t tAlo, int chunk,
                                              bool single_pass) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (RELU) asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi[j]) : "f"(v[j].y), "f"(v[j].x));
    else asm("cvt.rz.bf16x2.f32 %0, %1, %2;" : "=r"(hi[j]) : "f"(v[j].y), "f"(v[j].x));
    float rx, ry;
    asm("{\n\t"
        ".reg .b16 l, h, m1;\n\t"
        "mov.b32 {l, h}, %2;\n\t"
        "mov.b16 m1, 0xBF80;\n\t"  // -1.0 in bf16
        "fma.rn.f32.bf16 %0, l, m1, %3;\n\t"
        "fma.rn.f32.bf16 %1, h, m1, %4;\n\t"
        "}"
        : "=f"(rx), "=f"(ry)
        : "r"(hi[j]), "f"(v[j].x), "f"(v[j].y));
    if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo[j]) : "f"(ry), "f"(rx));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo[j]) : "f"(ry), "f"(rx));
  }
  tmem_st8(tAhi + chunk * 8, hi);
  if (!single_pass) tmem_st8(tAlo + chunk * 8, lo);
}


}  // namespace mmf
