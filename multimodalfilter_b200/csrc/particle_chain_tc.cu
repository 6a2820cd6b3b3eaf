// particle_chain_tc.cu -- R3 + R4 + R5 + first half of R6 on the 5th-generation tensor cores
// (MMF_PREC_BF16X3 / MMF_PREC_BF16): the throughput build of the per-particle MLP chain.
//
// Layout of the computation (one persistent CTA per SM, 4 groups x 128 threads):
//   * a group owns one TILE of 128 particles at a time; thread r of the group owns particle r of
//     the tile for the whole chain, and TMEM lane r;
//   * every 64->64 layer is D[128x64] (TMEM, fp32) = A[128x64] (TMEM, bf16) x W^T (smem, bf16,
//     K-major SWIZZLE_128B), issued by ONE thread as tcgen05.mma.cta_group::1.kind::f16 with the
//     A operand read from tensor memory (TS form);
//   * fp32-grade accuracy comes from split operands: a = a_hi + a_lo, w = w_hi + w_lo (bf16 each),
//     D = a_hi w_hi + a_hi w_lo + a_lo w_hi  (3 MMAs per K step, fp32 accumulate) -- MMF_PREC_BF16X3;
//     MMF_PREC_BF16 issues only the first product;
//   * the epilogue (tcgen05.ld -> +bias / +per-trajectory row / +residual -> ReLU -> bf16 split ->
//     tcgen05.st of the next layer's A operand) runs on the group's own 128 threads, so activations
//     never leave the SM: no shared-memory or HBM round trip between layers;
//   * the 4 groups are independent pipelines: while one group's accumulator is in the tensor pipe
//     the other three run their epilogues on the CUDA cores;
//   * weights: one chain at a time is resident in shared memory (dynamics 148 KiB, a head 116 KiB,
//     both bf16 halves), brought in by cp.async.bulk (TMA, mbarrier complete_tx) once per phase;
//     the kernel makes one pass over its tiles per chain ("phase"): dynamics -> moved state to HBM
//     -> head 0 -> running fused log-likelihood to HBM -> head 1 ... (+24 B/particle of traffic,
//     three orders of magnitude below the tensor time).
//
// Replaces the same reference lines as particle_chain_ffma.cu.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace mmf {

// GROUPS independent tile pipelines per CTA; TPR threads share one particle row (each owns 64/TPR
// activation columns), so a group is 128*TPR threads = 4*TPR warps, all four TMEM lane quadrants
// covered TPR times.  TMEM: 128 columns per group (64 accumulator + 32 A_hi + 32 A_lo) => GROUPS <= 4.
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

__global__ void k_pack_chain_mma(ChainDev ch, uint8_t* __restrict__ base) {
  // blockIdx.y selects the image: 0 -> nsplit 1; 1, 2 -> nsplit 2 rank 0, 1
  const int nsplit = blockIdx.y == 0 ? 1 : 2, rank = blockIdx.y == 0 ? 0 : (int)blockIdx.y - 1;
  uint8_t* dst = base + image_offset(ch, nsplit, rank);
  const int rows = U / nsplit, tile_b = TILE_B / nsplit;
  const int L = chain_layers(ch);
  const float* w = ch.w;
  // fp32 pack offsets (include/mmf_b200.h): in, pre-res, mid, post-res, out
  const int off_in = 0;
  const int off_first_res = ch.in_dim * U + U;
  const int off_mid = off_first_res + ch.n_pre * RES_FLOATS;
  const int off_post = off_mid + U * U;
  const int off_out = off_post + ch.n_post * RES_FLOATS;
  float* fdst = reinterpret_cast<float*>(dst + image_tiles_bytes(ch, nsplit));
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;

  for (int layer = 0; layer < L; ++layer) {
    const float* Wt;  // transposed [k][j]
    const float* b;   // may be null (mid layer: bias lives in the per-trajectory row)
    if (layer < 2 * ch.n_pre) {
      const float* r = w + off_first_res + (layer >> 1) * RES_FLOATS;
      Wt = r + (layer & 1) * (U * U + U);
      b = Wt + U * U;
    } else if (layer == 2 * ch.n_pre) {
      Wt = w + off_mid;
      b = nullptr;
    } else {
      const int rel = layer - 2 * ch.n_pre - 1;
      const float* r = w + off_post + (rel >> 1) * RES_FLOATS;
      Wt = r + (rel & 1) * (U * U + U);
      b = Wt + U * U;
    }
    uint8_t* hi = dst + (size_t)layer * 2 * tile_b;
    uint8_t* lo = hi + tile_b;
    for (int e = tid; e < rows * U; e += nth) {
      const int nl = e / U, k = e % U, n = nl + rank * rows;  // B[n][k] = W[n][k] = Wt[k][n]
      const float v = Wt[k * U + n];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      const int o = sw128_offset(nl, k);
      *reinterpret_cast<__nv_bfloat16*>(hi + o) = h;
      *reinterpret_cast<__nv_bfloat16*>(lo + o) = l;
    }
    for (int j = tid; j < U; j += nth) fdst[ch.in_dim * U + U + layer * U + j] = b ? b[j] : 0.0f;
  }
  {  // output layer: out_W[out_dim][64] row-major, rows >= out_dim are zero
    uint8_t* hi = dst + (size_t)L * 2 * tile_b;
    uint8_t* lo = hi + OUT_TILE_B;
    const float* W = w + off_out;
    for (int e = tid; e < OUT_PAD * U; e += nth) {
      const int nl = e / U, k = e % U, n = nl + rank * OUT_PAD;
      const float v = n < ch.out_dim ? W[n * U + k] : 0.0f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      const int o = sw128_offset(nl, k);
      *reinterpret_cast<__nv_bfloat16*>(hi + o) = h;
      *reinterpret_cast<__nv_bfloat16*>(lo + o) = l;
    }
    for (int j = tid; j < OUT_PAD; j += nth)
      fdst[ch.in_dim * U + U + L * U + j] = j < ch.out_dim ? W[ch.out_dim * U + j] : 0.0f;
  }
  for (int e = tid; e < ch.in_dim * U + U; e += nth) fdst[e] = w[off_in + e];
}

size_t chain_mma_bytes(const mmf_chain* chain) { return image_total_bytes(to_dev(*chain)); }

int pack_chain_mma(const mmf_chain* chain, void* dst, cudaStream_t stream) {
  MMF_REQUIRE(chain->w != nullptr, "pack_chain_mma: chain has no fp32 weights");
  MMF_REQUIRE(((uintptr_t)dst & 15) == 0, "pack_chain_mma: destination must be 16-byte aligned");
  k_pack_chain_mma<<<dim3(16, 3), 256, 0, stream>>>(to_dev(*chain), static_cast<uint8_t*>(dst));
  MMF_LAUNCH_CHECK("k_pack_chain_mma");
  return MMF_OK;
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

// Debug-only phase timestamps of (CTA 0, group 0, thread 0): enabled with MMF_TC_TIMESTAMPS=1, read back with
// mmf_debug_tc_timestamps().  Layout: per layer iteration 5 clock64 stamps.
__device__ unsigned long long g_tc_stamps[8192];
__device__ unsigned int g_tc_stamp_count;

struct TcParams {
  ChainDev chains[1 + MMF_MAX_HEADS];
  const uint8_t* images[1 + MMF_MAX_HEADS];
  int K;
  uint32_t enabled;
  int sd, N, M, single_pass;
  uint32_t wait_hint_ns;
  int stamps, use_lock;
  long long total;
  size_t image_cap;  // bytes reserved for the resident image (1024-aligned)
  const float* states_in;
  const float* eps;
  const float* rowbias;
  const float* logw_in;
  const float* modw;
  float* states_out;
  float* logw_out;
  float* ll_out;
  float q[MMF_MAX_SD * MMF_MAX_SD];
};

// bf16 split of 8 fp32 pairs, stored as the next A operand (columns [8*chunk, 8*chunk+8) of the hi and lo
// regions):  hi = rz_bf16(v) (truncation == the top 16 bits of v), lo = rn_bf16(v - hi).  With RELU the
// activation is folded into the two conversions: v < 0 gives hi = 0, and v - trunc(v) <= 0 gives lo = 0,
// so no separate max() is needed.  5 instructions per pair: F2FP, 2 x LOP, FFMA2, F2FP.
template <bool RELU>
__device__ __forceinline__ void store_a_chunk(const float2 (&v)[8], uint32_t tAhi, uint32_t tAlo, int chunk,
                                              bool single_pass) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 t = make_float2(__uint_as_float(__float_as_uint(v[j].x) & 0xffff0000u),
                                 __uint_as_float(__float_as_uint(v[j].y) & 0xffff0000u));
    const float2 r = __ffma2_rn(t, make_float2(-1.0f, -1.0f), v[j]);
    if (RELU) {
      asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi[j]) : "f"(v[j].y), "f"(v[j].x));
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo[j]) : "f"(r.y), "f"(r.x));
    } else {
      asm("cvt.rz.bf16x2.f32 %0, %1, %2;" : "=r"(hi[j]) : "f"(v[j].y), "f"(v[j].x));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo[j]) : "f"(r.y), "f"(r.x));
    }
  }
  tmem_st8(tAhi + chunk * 8, hi);
  if (!single_pass) tmem_st8(tAlo + chunk * 8, lo);
}

enum { EPI_RES_A = 0, EPI_RES_B = 1, EPI_MID_RELU = 2, EPI_MID_LINEAR = 3 };

// Epilogue of one 64-wide layer for this thread's row: accumulator (TMEM) -> +bias (+residual) ->
// activation -> residual stream update -> bf16 split -> next A operand (TMEM).
//   RES_A : t = relu(D + b1)          residual stream xr untouched
//   RES_B : y = relu(D + b2 + xr)     xr = y
//   MID_* : v = [relu](D + rowbias)   xr = v      (bias4 then points at the per-trajectory row in global memory)
// tD / tAhi / tAlo / bias4 already point at this thread's first column.
template <int KIND, int COLS>
__device__ __forceinline__ void epilogue(uint32_t tD, uint32_t tAhi, uint32_t tAlo, const float4* __restrict__ bias4,
                                         float2 (&xr)[COLS / 2], bool single_pass) {
  constexpr int CHUNKS = COLS / 16;
  // software pipeline over the accumulator chunks: the tcgen05.ld of chunk c+1 is in flight while chunk c
  // is processed (tcgen05.wait::ld waits for ALL outstanding loads, so it is issued after the compute block)
  uint32_t d[2][16];
  tmem_ld16(tD, d[0]);
  tc_wait_ld();
#pragma unroll
  for (int chunk = 0; chunk < CHUNKS; ++chunk) {
    if (chunk + 1 < CHUNKS) tmem_ld16(tD + (chunk + 1) * 16, d[(chunk + 1) & 1]);
    float2 b[8];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 t = (KIND >= EPI_MID_RELU) ? __ldg(bias4 + chunk * 4 + q4) : bias4[chunk * 4 + q4];
      b[2 * q4] = make_float2(t.x, t.y);
      b[2 * q4 + 1] = make_float2(t.z, t.w);
    }
    float2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float2 a = __fadd2_rn(make_float2(__uint_as_float(d[chunk & 1][2 * j]), __uint_as_float(d[chunk & 1][2 * j + 1])), b[j]);
      if (KIND == EPI_RES_B) a = __fadd2_rn(a, xr[chunk * 8 + j]);
      if (KIND == EPI_RES_B || KIND == EPI_MID_RELU) {  // the fp32 residual stream needs the real max()
        a.x = fmaxf(a.x, 0.0f);
        a.y = fmaxf(a.y, 0.0f);
      }
      if (KIND != EPI_RES_A) xr[chunk * 8 + j] = a;
      v[j] = a;
    }
    if (KIND == EPI_RES_A) store_a_chunk<true>(v, tAhi, tAlo, chunk, single_pass);  // relu folded into the cvt
    else store_a_chunk<false>(v, tAhi, tAlo, chunk, single_pass);
    if (chunk + 1 < CHUNKS) tc_wait_ld();
  }
}

// PAIR = false: cta_group::1, every CTA works alone.
// PAIR = true : CTAs are launched as clusters of two; group g of both CTAs advance in lock step and
//               ONE tcgen05.mma.cta_group::2 (M = 256) issued by the leader covers both tiles, halving
//               the number of MMA instructions per tile (an N <= 64 MMA costs ~47 cycles of tensor
//               front-end time whatever its N or M).
template <int TC_GROUPS, bool PAIR>
__global__ void __launch_bounds__(TC_GROUPS * 128, 1) k_particle_chain_tc(const __grid_constant__ TcParams P) {
  constexpr int TC_COLS = U;
  constexpr int TC_CHUNKS = TC_COLS / 16;
  constexpr int NSPLIT = PAIR ? 2 : 1;
  constexpr int TILE = TILE_B / NSPLIT;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem + P.image_cap);
  uint64_t* gbar = wbar + 1;                      // MMA-done, one per group (both CTAs in PAIR mode)
  uint64_t* ready = gbar + TC_MAX_GROUPS;         // PAIR: A operand ready in both CTAs (lives in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + TC_MAX_GROUPS);
  // Turnstile for the tensor pipe: a group issues its whole batch of MMAs (one layer) while holding it.
  // Without it the four issuing threads interleave their MMAs one by one, all groups finish their layers at
  // the same moment and the kernel degenerates into "everybody in the epilogue, then everybody queueing at
  // the tensor pipe" (measured with the phase timestamps below); batches served one at a time stagger the groups.
  uint32_t* mma_lock = tmem_slot + 1;

  const int tid = threadIdx.x, gt = tid & 127;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
  const int g = warp >> 2;                                   // group of 4 warps
  const int row = gt;  // particle row inside the tile == TMEM lane
  const int sd = P.sd;
  const bool single_pass = P.single_pass != 0;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;

  if (tid == 0) {
    *mma_lock = 0;
    mbar_init(wbar, 1);
    for (int i = 0; i < TC_GROUPS; ++i) {
      mbar_init(gbar + i, 1);
      mbar_init(ready + i, 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t grp_cols = tmem_base + g * 128;                // lane 0 view (for the MMA issuer)
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;   // this warp's 32-lane quadrant
  const uint32_t tD = grp_cols + lane_off;
  const uint32_t tAhi = tD + 64, tAlo = tD + 96;
  uint32_t wphase = 0, gphase = 0, rphase = 0;

  const long long tiles = (P.total + 127) / 128;
  constexpr uint32_t IDESC_L = make_idesc(64, PAIR ? 256 : 128);
  constexpr uint32_t IDESC_O = make_idesc(OUT_PAD * NSPLIT, PAIR ? 256 : 128);
  // work distribution: a "slot" is one tile per CTA of the unit (unit = CTA, or CTA pair)
  const long long unit = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const long long units = PAIR ? (gridDim.x >> 1) : gridDim.x;

  for (int c = 0; c <= P.K; ++c) {
    if (c > 0 && !((P.enabled >> (c - 1)) & 1u)) continue;
    const ChainDev ch = P.chains[c];
    const int L = chain_layers(ch);
    // is this the last enabled head?  (decides whether logw_out holds a running value or the result)
    bool last_head = false, first_head = false;
    if (c > 0) {
      last_head = (P.enabled >> c) == 0;
      first_head = (P.enabled & ((1u << (c - 1)) - 1u)) == 0;
    }

    // ---- bring this chain's image into shared memory (TMA bulk copy) -------------------------------
    // (PAIR: a CTA only gets here when every MMA that read its image has been committed and waited for,
    //  and the leader cannot issue next-phase MMAs before this CTA's groups signal `ready` again.)
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)image_bytes(ch, NSPLIT);
      const uint8_t* src = P.images[c] + image_offset(ch, NSPLIT, rank);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(wbar, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768) {
        const uint32_t n = bytes - off < 32768 ? bytes - off : 32768;
        bulk_g2s(smem + off, src + off, n, wbar);
      }
    }
    mbar_wait(wbar, wphase);
    wphase ^= 1;

    const uint32_t tiles_addr = smem_u32(smem);
    const float* fsm = reinterpret_cast<const float*>(smem + image_tiles_bytes(ch, NSPLIT));
    const float* in_Wt = fsm;
    const float* in_b = fsm + ch.in_dim * U;
    const float* biases = in_b + U;
    const float* out_b = biases + L * U;
    const int mid_at = 2 * ch.n_pre;

    for (long long it = 0;; ++it) {
      const long long tile_base = ((it * units + unit) * TC_GROUPS + g) * NSPLIT;
      if (tile_base >= tiles) break;  // identical for both CTAs of a pair
      const long long tile = tile_base + rank;
      const long long p_raw = tile * 128 + row;
      const bool live = p_raw < P.total;
      const long long p = live ? p_raw : P.total - 1;
      const int n = (int)(p / P.M);
      const float* xsrc = (c == 0) ? P.states_in : P.states_out;
      float x[MMF_MAX_SD];
#pragma unroll
      for (int i = 0; i < MMF_MAX_SD; ++i) x[i] = (i < sd) ? xsrc[p * sd + i] : 0.0f;

      // ---- input layer on the CUDA cores: xr = relu(in_W x + in_b) -> A operand ------------------------
      float2 xr[TC_COLS / 2];
      {
        const float4* b4 = reinterpret_cast<const float4*>(in_b);
        const float4* w4 = reinterpret_cast<const float4*>(in_Wt);
#pragma unroll
        for (int chunk = 0; chunk < TC_CHUNKS; ++chunk) {
          float2 v[8];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 t = b4[chunk * 4 + q4];
            v[2 * q4] = make_float2(t.x, t.y);
            v[2 * q4 + 1] = make_float2(t.z, t.w);
          }
#pragma unroll
          for (int i = 0; i < MMF_MAX_SD; ++i) {
            if (i < sd) {
              const float2 xi = make_float2(x[i], x[i]);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const float4 t = w4[i * (U / 4) + chunk * 4 + q4];
                v[2 * q4] = __ffma2_rn(make_float2(t.x, t.y), xi, v[2 * q4]);
                v[2 * q4 + 1] = __ffma2_rn(make_float2(t.z, t.w), xi, v[2 * q4 + 1]);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j].x = fmaxf(v[j].x, 0.0f);
            v[j].y = fmaxf(v[j].y, 0.0f);
            xr[chunk * 8 + j] = v[j];
          }
          store_a_chunk<false>(v, tAhi, tAlo, chunk, single_pass);
        }
      }

      // ---- 64x64 layers ----------------------------------------------------------------------------------
      for (int layer = 0; layer <= L; ++layer) {
        const bool is_out = (layer == L);
        const bool stamp = P.stamps && blockIdx.x == 0 && tid == 0 && c == 0 && it == 1;
        unsigned long long ts0 = 0, ts1 = 0, ts2 = 0, ts3 = 0;
        if (stamp) ts0 = clock64();
        // hand the A operand to the tensor core
        tc_wait_st();
        tc_fence_before();
        if (stamp) ts1 = clock64();
        group_bar(1 + g, 128);
        if (stamp) ts2 = clock64();
        if ((warp & 3) == 0 && elect_one_sync()) {  // one lane of the group's first warp issues
          bool issue = true;
          if (PAIR) {
            mbar_arrive_remote(ready + g, 0);  // "my A operand is in TMEM and I am done reading D"
            issue = (rank == 0);
            if (issue) mbar_wait_cluster(ready + g, rphase);
          }
          if (issue) {
            if (P.use_lock) {
              while (atomicCAS(mma_lock, 0u, 1u) != 0u) __nanosleep(20);
            }
            tc_fence_after();
            const uint32_t hi_addr = tiles_addr + (is_out ? (uint32_t)L * 2 * TILE : (uint32_t)layer * 2 * TILE);
            const uint32_t lo_addr = hi_addr + (is_out ? OUT_TILE_B : TILE);
            const uint64_t bhi = make_b_desc(hi_addr), blo = make_b_desc(lo_addr);
            const uint32_t idesc = is_out ? IDESC_O : IDESC_L;
            const uint32_t a_hi = grp_cols + 64, a_lo = grp_cols + 96;
            if (PAIR) {
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_ts2(grp_cols, a_hi + k * 8, bhi + (uint64_t)(k * 2), idesc, k > 0);
              if (!single_pass) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ts2(grp_cols, a_hi + k * 8, blo + (uint64_t)(k * 2), idesc, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ts2(grp_cols, a_lo + k * 8, bhi + (uint64_t)(k * 2), idesc, 1);
              }
              tc_commit2(gbar + g);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_ts(grp_cols, a_hi + k * 8, bhi + (uint64_t)(k * 2), idesc, k > 0);
              if (!single_pass) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ts(grp_cols, a_hi + k * 8, blo + (uint64_t)(k * 2), idesc, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ts(grp_cols, a_lo + k * 8, bhi + (uint64_t)(k * 2), idesc, 1);
              }
              tc_commit(gbar + g);
            }
            if (P.use_lock) atomicExch(mma_lock, 0u);
          }
        }
        rphase ^= 1;
        if (stamp) ts3 = clock64();
        mbar_wait(gbar + g, gphase, P.wait_hint_ns);
        gphase ^= 1;
        tc_fence_after();
        if (stamp) {
          const unsigned int slot = atomicAdd(&g_tc_stamp_count, 1u);
          if (slot < 1600) {
            g_tc_stamps[slot * 5 + 0] = ts0;
            g_tc_stamps[slot * 5 + 1] = ts1;
            g_tc_stamps[slot * 5 + 2] = ts2;
            g_tc_stamps[slot * 5 + 3] = ts3;
            g_tc_stamps[slot * 5 + 4] = clock64();
          }
        }
        if (is_out) break;

        // ---- epilogue of this layer = producer of the next layer's A operand ---------------------------
        if (layer == mid_at) {
          const float4* brow = reinterpret_cast<const float4*>(P.rowbias + ((size_t)c * P.N + n) * U);
          if (ch.mid_relu) epilogue<EPI_MID_RELU, TC_COLS>(tD, tAhi, tAlo, brow, xr, single_pass);
          else epilogue<EPI_MID_LINEAR, TC_COLS>(tD, tAhi, tAlo, brow, xr, single_pass);
        } else {
          const int rel = (layer < mid_at) ? layer : layer - mid_at - 1;
          const float4* bsm = reinterpret_cast<const float4*>(biases + layer * U);
          if ((rel & 1) == 0) epilogue<EPI_RES_A, TC_COLS>(tD, tAhi, tAlo, bsm, xr, single_pass);
          else epilogue<EPI_RES_B, TC_COLS>(tD, tAhi, tAlo, bsm, xr, single_pass);
        }
      }

      // ---- output layer result: y[o] = D[o] + out_b[o] -----------------------------------------------------
      float y[MMF_MAX_SD + 1];
      {
        uint32_t d[16];
        tmem_ld16(tD, d);
        tc_wait_ld();
#pragma unroll
        for (int o = 0; o < MMF_MAX_SD + 1; ++o) y[o] = __uint_as_float(d[o]) + out_b[o];
      }

      if (c == 0) {
        float gsel = 0.0f;
#pragma unroll
        for (int o = 0; o < MMF_MAX_SD + 1; ++o)
          if (o == sd) gsel = y[o];
        const float gate = 1.0f / (1.0f + expf(-gsel));
        float e[MMF_MAX_SD];
#pragma unroll
        for (int i = 0; i < MMF_MAX_SD; ++i) e[i] = (i < sd) ? P.eps[p * sd + i] : 0.0f;
#pragma unroll
        for (int i = 0; i < MMF_MAX_SD; ++i) {
          if (i < sd) {
            float noise = 0.0f;
#pragma unroll
            for (int j = 0; j < MMF_MAX_SD; ++j)
              if (j <= i && j < sd) noise = fmaf(P.q[i * sd + j], e[j], noise);
            const float moved = (x[i] + y[i] * gate) + noise;
            if (live) P.states_out[p * sd + i] = moved;
          }
        }
      } else {
        const float ll = y[0];
        if (P.ll_out != nullptr && live) P.ll_out[(size_t)(c - 1) * P.total + p] = ll;
        const float v = ll + (P.modw != nullptr ? __ldg(P.modw + (size_t)n * P.K + (c - 1)) : 0.0f);
        float fused = v;
        if (!first_head) {  // running log-sum-exp kept in logw_out between head phases
          const float prev = P.logw_out[p];
          const float mx = fmaxf(prev, v);
          fused = (mx == -INFINITY) ? -INFINITY : mx + logf(expf(prev - mx) + expf(v - mx));
        }
        if (live) P.logw_out[p] = last_head ? P.logw_in[p] + fused : fused;
      }
      // the next tile's input layer overwrites the A region: the out-layer MMA that read it has completed
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int launch_particle_chain_tc(const mmf_pf_model* model, int N, int M, const float* states_in, const float* eps,
                             const float* rowbias, const float* logw_in, const float* modw, uint32_t enabled,
                             int precision, float* states_out, float* logw_out, float* ll_out, cudaStream_t stream) {
  // pipeline shape: MMF_TC_VARIANT = <groups><cta group>: 41 = 4 groups, cta_group::1; 42 = 4 groups, CTA pairs
  int variant = 41;
  if (const char* env = getenv("MMF_TC_VARIANT")) variant = atoi(env);
  const bool pair = (variant % 10) == 2;
  TcParams P;
  P.K = model->num_heads;
  size_t cap = 0;
  for (int c = 0; c <= P.K; ++c) {
    const mmf_chain& src = (c == 0) ? model->dynamics : model->heads[c - 1];
    P.chains[c] = to_dev(src);
    P.images[c] = static_cast<const uint8_t*>(src.w_mma);
    if (c == 0 || ((enabled >> (c - 1)) & 1u)) {
      MMF_REQUIRE(src.w_mma != nullptr, "tensor-core chain %d has no operand image: call mmf_pack_chain_mma first", c);
      MMF_REQUIRE(((uintptr_t)src.w_mma & 15) == 0, "operand image %d must be 16-byte aligned", c);
      const size_t b = image_bytes(P.chains[c], pair ? 2 : 1);
      cap = b > cap ? b : cap;
    }
  }
  P.image_cap = cap;
  P.enabled = enabled;
  P.sd = model->state_dim;
  P.N = N;
  P.M = M;
  P.single_pass = (precision == MMF_PREC_BF16) ? 1 : 0;
  P.total = (long long)N * M;
  P.states_in = states_in;
  P.eps = eps;
  P.rowbias = rowbias;
  P.logw_in = logw_in;
  P.modw = modw;
  P.states_out = states_out;
  P.logw_out = logw_out;
  P.ll_out = ll_out;
  for (int i = 0; i < MMF_MAX_SD * MMF_MAX_SD; ++i) P.q[i] = model->q_tril[i];
  P.wait_hint_ns = 0;
  P.stamps = getenv("MMF_TC_TIMESTAMPS") != nullptr;
  P.use_lock = 0;
  if (const char* env = getenv("MMF_TC_LOCK")) P.use_lock = atoi(env);

  const size_t smem = cap + 1024;  // + barriers, TMEM slot
  int sms = 148, dev = 0;
  MMF_CUDA(cudaGetDevice(&dev));
  MMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tiles = (P.total + 127) / 128;
#define MMF_TC_LAUNCH(G, PAIRED)                                                                  \
  do {                                                                                            \
    static thread_local int configured_dev = -1;                                                  \
    static thread_local size_t window = 0;                                                        \
    if (configured_dev != dev) {                                                                  \
      int rc = opt_in_shared_memory(k_particle_chain_tc<G, PAIRED>, &window);                     \
      if (rc) return rc;                                                                          \
      configured_dev = dev;                                                                       \
    }                                                                                             \
    MMF_REQUIRE(smem <= window, "tensor-core chain needs %zu B of shared memory (window %zu B)", smem, window); \
    const int per_unit = G * (PAIRED ? 2 : 1);                                                    \
    long long units = (tiles + per_unit - 1) / per_unit;                                          \
    const long long max_units = PAIRED ? sms / 2 : sms;                                           \
    if (units > max_units) units = max_units;                                                     \
    cudaLaunchConfig_t cfg = {};                                                                  \
    cfg.gridDim = dim3((unsigned)(units * (PAIRED ? 2 : 1)));                                     \
    cfg.blockDim = dim3(G * 128);                                                                 \
    cfg.dynamicSmemBytes = smem;                                                                  \
    cfg.stream = stream;                                                                          \
    cudaLaunchAttribute attr[1];                                                                  \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                             \
    attr[0].val.clusterDim.x = PAIRED ? 2 : 1;                                                    \
    attr[0].val.clusterDim.y = 1;                                                                 \
    attr[0].val.clusterDim.z = 1;                                                                 \
    cfg.attrs = attr;                                                                             \
    cfg.numAttrs = 1;                                                                             \
    MMF_CUDA(cudaLaunchKernelEx(&cfg, k_particle_chain_tc<G, PAIRED>, P));                        \
  } while (0)
  switch (variant) {
    case 41: MMF_TC_LAUNCH(4, false); break;
    case 31: MMF_TC_LAUNCH(3, false); break;
    case 42: MMF_TC_LAUNCH(4, true); break;
    case 32: MMF_TC_LAUNCH(3, true); break;
    default: set_error("unknown MMF_TC_VARIANT %d", variant); return MMF_E_INVALID;
  }
#undef MMF_TC_LAUNCH
  MMF_LAUNCH_CHECK("k_particle_chain_tc");
  return MMF_OK;
}

}  // namespace mmf

// debug-only export (not part of the public ABI): copies the recorded phase timestamps to the host
extern "C" int mmf_debug_tc_timestamps(unsigned long long* out, int max_entries) {
  unsigned int count = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&count, mmf::g_tc_stamp_count, sizeof(count));
  int n = (int)count < max_entries ? (int)count : max_entries;
  if (n > 1600) n = 1600;
  cudaMemcpyFromSymbol(out, mmf::g_tc_stamps, sizeof(unsigned long long) * 5 * n);
  unsigned int zero = 0;
  cudaMemcpyToSymbol(mmf::g_tc_stamp_count, &zero, sizeof(zero));
  return n;
}
