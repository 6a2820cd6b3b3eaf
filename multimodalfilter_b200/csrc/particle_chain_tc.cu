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
#include "tc_common.cuh"

#ifndef MMF_TC_ABLATE
#define MMF_TC_ABLATE 0
#endif
#if MMF_TC_ABLATE == 3  // measurement build (tools/ubench): clock64 stamps of CTA 0's hand-overs, second tile of chain 0
__device__ unsigned long long g_ws_stamps[2 * 16 * 4 * 4];
#define MMF_STAMP(cond, role, layer, grp, k) \
  do { if (cond) g_ws_stamps[(((role) * 16 + (layer)) * 4 + (grp)) * 4 + (k)] = clock64(); } while (0)
#else
#define MMF_STAMP(cond, role, layer, grp, k) do { } while (0)
#endif

namespace mmf {

struct TcParams {
  ChainDev chains[1 + MMF_MAX_HEADS];
  const uint8_t* images[1 + MMF_MAX_HEADS];
  int K;
  uint32_t enabled;
  int sd, N, M, single_pass;
  uint32_t wait_hint_ns;
  int first_chain;   // 0: dynamics + heads; 1: heads only (states_out already holds the moved particles)
  float* act_out;    // training: (K, L+1, N*M, 64) fp32 activations feeding every GEMM layer of every head, or null
  long long total;
  size_t image_cap;  // bytes reserved for the resident image (1024-aligned)
  const float* states_in;
  const float* eps;
  const float* rowbias;
  int rb_stride;      // trajectories per rowbias plane (N, or T * N when the rows of a whole sequence were hoisted)
  const float* logw_in;
  const float* modw;
  float* states_out;
  float* logw_out;
  float* ll_out;
  float q[MMF_MAX_SD * MMF_MAX_SD];
};

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

enum { EPI_RES_A = 0, EPI_RES_B = 1, EPI_MID_RELU = 2, EPI_MID_LINEAR = 3 };

// Epilogue of one 64-wide layer for this thread's row: accumulator (TMEM) -> +bias (+residual) ->
// activation -> residual stream update -> bf16 split -> next A operand (TMEM).
//   RES_A : t = relu(D + b1)          residual stream xr untouched
//   RES_B : y = relu(D + b2 + xr)     xr = y
//   MID_* : v = [relu](D + rowbias)   xr = v      (bias4 then points at the per-trajectory row in global memory)
// tD / tAhi / tAlo / bias4 already point at this thread's first column.
// store 16 activations of this thread's row (training: kept for the backward pass)
// The saved activations (and the deltas of the backward kernel) are stored CHUNK-MAJOR: plane[c4][row][4 floats],
// c4 = column / 4, so that the 32 rows of a warp write 512 contiguous bytes per store instruction (row-major
// 256-byte rows cost one 32-byte sector per thread per instruction and saturated the LSU: ncu, r01).
// act_row points at this row's float4 of column chunk 0; stride4 = rows * 4 floats separates column chunks.
__device__ __forceinline__ void store_act_chunk(float* act_row, int chunk, const float2 (&v)[8], size_t stride4) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<float4*>(act_row + (size_t)(chunk * 4 + q) * stride4) =
        make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
}

template <int KIND, int COLS>
__device__ __forceinline__ void epilogue(uint32_t tD, uint32_t tAhi, uint32_t tAlo, const float4* __restrict__ bias4,
                                         float2 (&xr)[COLS / 2], bool single_pass, float* act_row, size_t act_stride4) {
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
    if (act_row != nullptr) {
      if (KIND == EPI_RES_A) {  // the stored activation needs the real max() (it is folded into the cvt below)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = make_float2(fmaxf(v[j].x, 0.0f), fmaxf(v[j].y, 0.0f));
      }
      store_act_chunk(act_row, chunk, v, act_stride4);
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

  const int tid = threadIdx.x, gt = tid & 127;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
  const int g = warp >> 2;                                   // group of 4 warps
  const int row = gt;  // particle row inside the tile == TMEM lane
  const int sd = P.sd;
  const bool single_pass = P.single_pass != 0;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;

  if (tid == 0) {
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

  for (int c = P.first_chain; c <= P.K; ++c) {
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

      // training: base of this particle's saved activations for head c (layer index selects the plane)
      float* act_base = (P.act_out != nullptr && c > 0 && live)
                            ? P.act_out + (size_t)(c - 1) * (L + 1) * P.total * U + (size_t)p * 4
                            : nullptr;
      const size_t act_plane = (size_t)P.total * U;
      const size_t act_stride4 = (size_t)P.total * 4;

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
          if (act_base != nullptr) store_act_chunk(act_base, chunk, v, act_stride4);
          store_a_chunk<false>(v, tAhi, tAlo, chunk, single_pass);
        }
      }

      // ---- 64x64 layers ----------------------------------------------------------------------------------
      for (int layer = 0; layer <= L; ++layer) {
        const bool is_out = (layer == L);
        // hand the A operand to the tensor core
        tc_wait_st();
        tc_fence_before();
        group_bar(1 + g, 128);
        if ((warp & 3) == 0 && elect_one_sync()) {  // one lane of the group's first warp issues
          bool issue = true;
          if (PAIR) {
            mbar_arrive_remote(ready + g, 0);  // "my A operand is in TMEM and I am done reading D"
            issue = (rank == 0);
            if (issue) mbar_wait_cluster(ready + g, rphase);
          }
          if (issue) {
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
          }
        }
        rphase ^= 1;
        mbar_wait(gbar + g, gphase, P.wait_hint_ns);
        gphase ^= 1;
        tc_fence_after();
        if (is_out) break;

        // ---- epilogue of this layer = producer of the next layer's A operand ---------------------------
        if (layer == mid_at) {
          const float4* brow = reinterpret_cast<const float4*>(P.rowbias + ((size_t)c * P.rb_stride + n) * U);
          float* arow = act_base ? act_base + (size_t)(layer + 1) * act_plane : nullptr;
          if (ch.mid_relu) epilogue<EPI_MID_RELU, TC_COLS>(tD, tAhi, tAlo, brow, xr, single_pass, arow, act_stride4);
          else epilogue<EPI_MID_LINEAR, TC_COLS>(tD, tAhi, tAlo, brow, xr, single_pass, arow, act_stride4);
        } else {
          const int rel = (layer < mid_at) ? layer : layer - mid_at - 1;
          const float4* bsm = reinterpret_cast<const float4*>(biases + layer * U);
          float* arow = act_base ? act_base + (size_t)(layer + 1) * act_plane : nullptr;
          if ((rel & 1) == 0) epilogue<EPI_RES_A, TC_COLS>(tD, tAhi, tAlo, bsm, xr, single_pass, arow, act_stride4);
          else epilogue<EPI_RES_B, TC_COLS>(tD, tAhi, tAlo, bsm, xr, single_pass, arow, act_stride4);
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


// ---- warp-specialised kernel (the default): dedicated issuer warps feed the tensor pipe ---------------------------
// In the symmetric kernel above every group issues its own MMAs from inside its epilogue warps.  Measured (ncu r01 and the
// clock64 timeline of the measurement build, profiles/r02_*): a layer costs epilogue time PLUS MMA time instead of their
// maximum.  Three facts from the timeline shape this kernel:
//   * the tensor pipe's queue is one or two MMAs deep: the issuing lane blocks for ~32 cycles per MMA, so "issue 12 MMAs"
//     takes as long as executing them (370-480 cycles) -> the issuing lane cannot do anything else;
//   * waiting on an mbarrier that has ALREADY completed still costs ~130 cycles: one issuer serving the groups in turn
//     leaves the pipe idle for that long between batches (600 cycles per tile-layer against 384 of tensor work);
//   * tcgen05.commit -> waiting warps running again: ~190 cycles.
// So: NI = 2 issuer warps (warps 4 G and 4 G + 1, issuer w serves groups w, w + NI, ...), each blocked on the pipe while
// the other sits in its barrier wait; the pipe always has a batch to run, and at most NI accumulators complete together,
// so the groups' epilogues stay staggered against the MMAs.  The epilogue warps lose the issue code, the per-layer named
// barrier and the divergent single-lane region from their instruction stream, and hand over through mbarriers only:
//   a_ready[g] (count 4): a warp of group g arrives once its part of the next A operand is in TMEM
//   d_ready[g] (count 1): tcgen05.commit after the MMAs of group g's accumulator
// At a tile boundary the next tile's input layer is computed and published BEFORE the finished tile's output arithmetic
// (gate, noise, fusion, global loads and stores), whose latency would otherwise sit on the group's critical path.
// mbarrier helpers on precomputed 32-bit shared-memory addresses (the generic-pointer wrappers re-derive the address,
// cluster CTA id included, at every use: ~10 instructions per wait / arrive in the hot loop).  The wait parks the warp
// in hardware for up to `hint` ns per attempt (it is woken by the completing arrive), so a waiting warp costs the
// epilogue warps that share its scheduler almost no issue slots; the attempt bound turns a lost arrival into a trap.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok, attempts = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    if (!ok && ++attempts > (1u << 22)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int KIND>
__device__ __forceinline__ void epilogue_ws(uint32_t tD, uint32_t tAhi, uint32_t tAlo, const float4* __restrict__ bias4,
                                            float2 (&xr)[U / 2], bool single_pass, float* act_row, size_t act_stride4,
                                            uint32_t dbar, uint32_t& dph, uint32_t hint_ns,
                                            unsigned long long* stamp = nullptr) {
  mbar_wait_a(dbar, dph, hint_ns);
  dph ^= 1;
  tc_fence_after();
#if MMF_TC_ABLATE == 3
  if (stamp) *stamp = clock64();
#endif
#if MMF_TC_ABLATE != 2  // measurement build 2 (tools/ubench): no epilogue arithmetic -> MMA pipeline + hand-over alone
  epilogue<KIND, U>(tD, tAhi, tAlo, bias4, xr, single_pass, act_row, act_stride4);
#endif
}

// hand the A operand to the issuer: my TMEM stores have retired and are ordered before the arrive it observes
__device__ __forceinline__ void ws_publish(uint32_t abar_g, bool lane0) {
  tc_wait_st();
  tc_fence_before();
  __syncwarp();
  if (lane0) mbar_arrive_a(abar_g);
}

// start of a chain phase, executed by ALL threads of the CTA: barrier, bulk copy of the chain's operand image into
// shared memory (one thread), wait.  Every MMA that read the previous image has completed: each group waited for
// its last accumulator before it got here.
__device__ __forceinline__ void ws_load_image(const TcParams& P, int c, uint8_t* smem, uint64_t* wbar, uint32_t& wphase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)image_bytes(P.chains[c], 1);
    const uint8_t* src = P.images[c];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(wbar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768) {
      const uint32_t n = bytes - off < 32768 ? bytes - off : 32768;
      bulk_g2s(smem + off, src + off, n, wbar);
    }
  }
  mbar_wait(wbar, wphase);
  wphase ^= 1;
}

template <int G, int NI, bool TRAIN>
__global__ void __launch_bounds__(G * 128 + 128, 1) k_particle_chain_ws(const __grid_constant__ TcParams P) {
  constexpr int TC_CHUNKS = U / 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem + P.image_cap);
  uint64_t* dbar = wbar + 1;                  // [G]
  uint64_t* abar = dbar + TC_MAX_GROUPS;      // [G]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(abar + TC_MAX_GROUPS);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
  const bool service = warp >= 4 * G;  // warps 4 G .. 4 G + 3: NI issuers and parked warps (a full warpgroup)
  const int g = warp >> 2;             // worker: group
  const int row = tid & 127;           // worker: particle row inside the tile == TMEM lane
  const int sd = P.sd;
  const bool single_pass = P.single_pass != 0;

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int i = 0; i < G; ++i) {
      mbar_init(dbar + i, 1);
      mbar_init(abar + i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t dbar_a = smem_u32(dbar), abar_a = smem_u32(abar);  // barrier i at + 8 i
  const uint32_t hint_ns = P.wait_hint_ns;
  uint32_t wphase = 0;
  const long long tiles = (P.total + 127) / 128;
  constexpr uint32_t IDESC_L = make_idesc(U, 128);
  constexpr uint32_t IDESC_O = make_idesc(OUT_PAD, 128);
  const long long unit = blockIdx.x, units = gridDim.x;

  if (service) {
    // Register budget: the CTA is launched with 96 registers per thread (640 threads).  The service warpgroup hands
    // 64 x 128 registers back to the CTA pool and the four worker warpgroups take 16 x 128 each: 112 registers for
    // the epilogue (fp32 residual stream 64 + two accumulator chunks in flight 32 + the split).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    const int me = warp - 4 * G;  // issuer index (< NI) or parked
    uint32_t aph = 0;             // bit i = parity of a_ready[i]
    for (int c = P.first_chain; c <= P.K; ++c) {
      if (c > 0 && !((P.enabled >> (c - 1)) & 1u)) continue;
      const int L = chain_layers(P.chains[c]);
      ws_load_image(P, c, smem, wbar, wphase);
      if (me >= NI) continue;
      const uint32_t tiles_addr = smem_u32(smem);
      for (long long it = 0;; ++it) {
        const long long tile0 = (it * units + unit) * G;
        if (tile0 >= tiles) break;
        const long long left = tiles - tile0;
        const int ng = left < G ? (int)left : G;  // groups that have a tile in this slot
        for (int layer = 0; layer <= L; ++layer) {
          const bool is_out = layer == L;
          const uint32_t hi_addr = tiles_addr + (is_out ? (uint32_t)L * 2 * TILE_B : (uint32_t)layer * 2 * TILE_B);
          const uint32_t lo_addr = hi_addr + (is_out ? OUT_TILE_B : TILE_B);
          const uint64_t bhi = make_b_desc(hi_addr), blo = make_b_desc(lo_addr);
          const uint32_t idesc = is_out ? IDESC_O : IDESC_L;
#pragma unroll
          for (int k0 = 0; k0 < G; k0 += NI) {
            const int i = k0 + me;
            if (i >= ng) break;
            MMF_STAMP(blockIdx.x == 0 && c == 0 && it == 1 && (tid & 31) == 0, 0, layer, i, 0);
            mbar_wait_a(abar_a + 8 * i, (aph >> i) & 1u, hint_ns);
            aph ^= 1u << i;
            tc_fence_after();
            MMF_STAMP(blockIdx.x == 0 && c == 0 && it == 1 && (tid & 31) == 0, 0, layer, i, 1);
            if (elect_one_sync()) {
              const uint32_t d = tmem_base + i * 128, a_hi = d + 64, a_lo = d + 96;
#if MMF_TC_ABLATE != 1  // measurement build 1 (tools/ubench): no MMAs, the commit arrives at once -> epilogue pipeline alone
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_ts(d, a_hi + k * 8, bhi + (uint64_t)(k * 2), idesc, k > 0);
              if (!single_pass) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ts(d, a_hi + k * 8, blo + (uint64_t)(k * 2), idesc, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ts(d, a_lo + k * 8, bhi + (uint64_t)(k * 2), idesc, 1);
              }
#endif
              tc_commit_a(dbar_a + 8 * i);
            }
            __syncwarp();
            MMF_STAMP(blockIdx.x == 0 && c == 0 && it == 1 && (tid & 31) == 0, 0, layer, i, 2);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;  // this warp's 32-lane quadrant
    const uint32_t tD = tmem_base + g * 128 + lane_off;
    const uint32_t tAhi = tD + 64, tAlo = tD + 96;
    uint32_t dph = 0;  // parity of d_ready[g]
    const uint32_t my_dbar = dbar_a + 8 * g, my_abar = abar_a + 8 * g;
    const bool lane0 = (tid & 31) == 0;
    for (int c = P.first_chain; c <= P.K; ++c) {
      if (c > 0 && !((P.enabled >> (c - 1)) & 1u)) continue;
      const ChainDev ch = P.chains[c];
      const int L = chain_layers(ch);
      bool last_head = false, first_head = false;
      if (c > 0) {
        last_head = (P.enabled >> c) == 0;
        first_head = (P.enabled & ((1u << (c - 1)) - 1u)) == 0;
      }
      ws_load_image(P, c, smem, wbar, wphase);

      const float* fsm = reinterpret_cast<const float*>(smem + image_tiles_bytes(ch, 1));
      const float* in_Wt = fsm;
      const float* in_b = fsm + ch.in_dim * U;
      const float* biases = in_b + U;
      const float* out_b = biases + L * U;
      const int mid_at = 2 * ch.n_pre;
      const float* xsrc = (c == 0) ? P.states_in : P.states_out;
      const size_t act_plane = (size_t)P.total * U;
      const size_t act_stride4 = (size_t)P.total * 4;
      const long long tile_step = units * G;

      float2 xr[U / 2];  // fp32 residual stream of this thread's row
      // input layer on the CUDA cores: xr = relu(in_W x + in_b) -> A operand (and, training, activation plane 0)
      auto input_layer = [&](const float (&x)[MMF_MAX_SD], float* act_base) {
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
          if (TRAIN && act_base != nullptr) store_act_chunk(act_base, chunk, v, act_stride4);
          store_a_chunk<false>(v, tAhi, tAlo, chunk, single_pass);
        }
      };
      // training: base of a particle's saved activations for head c (the layer index selects the plane)
      auto act_of = [&](long long p, bool live) -> float* {
        return (TRAIN && P.act_out != nullptr && c > 0 && live)
                   ? P.act_out + (size_t)(c - 1) * (L + 1) * P.total * U + (size_t)p * 4
                   : nullptr;
      };

      long long tile = unit * G + g;
      if (tile >= tiles) continue;
      long long p_raw = tile * 128 + row;
      bool live = p_raw < P.total;
      long long p = live ? p_raw : P.total - 1;
      float x[MMF_MAX_SD];
#pragma unroll
      for (int i = 0; i < MMF_MAX_SD; ++i) x[i] = (i < sd) ? xsrc[p * sd + i] : 0.0f;
      float* act_base = act_of(p, live);
      input_layer(x, act_base);
      ws_publish(my_abar, lane0);

      for (long long it = 0;; ++it) {
        const int n = (int)(p / P.M);
        {  // the mid layer reads this trajectory's hoisted row from global memory right after its barrier wait: have
           // it in L1 by then (two 128-byte lines; the 128 rows of a tile share a handful of trajectories)
          const float* brow = P.rowbias + ((size_t)c * P.rb_stride + n) * U;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(brow));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(brow + 32));
        }
        // ---- 64x64 layers: wait for the accumulator, epilogue = next A operand, hand it over -------------------------
        for (int layer = 0; layer < L; ++layer) {
          MMF_STAMP(blockIdx.x == 0 && c == 0 && it == 1 && (tid & 127) == 0, 1, layer, g, 0);
          float* arow = (TRAIN && act_base) ? act_base + (size_t)(layer + 1) * act_plane : nullptr;
#if MMF_TC_ABLATE == 3
          unsigned long long* st = (blockIdx.x == 0 && c == 0 && it == 1 && (tid & 127) == 0) ? &g_ws_stamps[((16 + layer) * 4 + g) * 4 + 1] : nullptr;
#else
          unsigned long long* st = nullptr;
#endif
          if (layer == mid_at) {
            const float4* brow = reinterpret_cast<const float4*>(P.rowbias + ((size_t)c * P.rb_stride + n) * U);
            if (ch.mid_relu) epilogue_ws<EPI_MID_RELU>(tD, tAhi, tAlo, brow, xr, single_pass, arow, act_stride4, my_dbar, dph, hint_ns, st);
            else epilogue_ws<EPI_MID_LINEAR>(tD, tAhi, tAlo, brow, xr, single_pass, arow, act_stride4, my_dbar, dph, hint_ns, st);
          } else {
            const int rel = (layer < mid_at) ? layer : layer - mid_at - 1;
            const float4* bsm = reinterpret_cast<const float4*>(biases + layer * U);
            if ((rel & 1) == 0) epilogue_ws<EPI_RES_A>(tD, tAhi, tAlo, bsm, xr, single_pass, arow, act_stride4, my_dbar, dph, hint_ns, st);
            else epilogue_ws<EPI_RES_B>(tD, tAhi, tAlo, bsm, xr, single_pass, arow, act_stride4, my_dbar, dph, hint_ns, st);
          }
          MMF_STAMP(blockIdx.x == 0 && c == 0 && it == 1 && (tid & 127) == 0, 1, layer, g, 2);
          ws_publish(my_abar, lane0);  // after the last layer: the A operand of the output layer
        }
        // ---- tile boundary: fetch the next tile's particle while the output layer is in the tensor pipe -------------
        const long long tile_n = tile + tile_step;
        const bool have_next = tile_n < tiles;
        const long long pn_raw = tile_n * 128 + row;
        const bool live_n = have_next && pn_raw < P.total;
        const long long p_n = live_n ? pn_raw : P.total - 1;
        float xn[MMF_MAX_SD];
#pragma unroll
        for (int i = 0; i < MMF_MAX_SD; ++i) xn[i] = (have_next && i < sd) ? xsrc[p_n * sd + i] : 0.0f;
        // operands of the output arithmetic, fetched early as well
        float e[MMF_MAX_SD], prev = 0.0f, lw_in = 0.0f, mw = 0.0f;
        if (c == 0) {
#pragma unroll
          for (int i = 0; i < MMF_MAX_SD; ++i) e[i] = (i < sd) ? P.eps[p * sd + i] : 0.0f;
        } else {
          if (!first_head) prev = P.logw_out[p];
          if (last_head) lw_in = P.logw_in[p];
          if (P.modw != nullptr) mw = __ldg(P.modw + (size_t)n * P.K + (c - 1));
        }
        mbar_wait_a(my_dbar, dph, hint_ns);
        dph ^= 1;
        tc_fence_after();
        float y[MMF_MAX_SD + 1];
        {
          uint32_t d[16];
          tmem_ld16(tD, d);
          tc_wait_ld();
#pragma unroll
          for (int o = 0; o < MMF_MAX_SD + 1; ++o) y[o] = __uint_as_float(d[o]) + out_b[o];
        }
        // the output-layer MMA has completed: its A operand and accumulator are free -> next tile's input layer first
        float* act_base_n = nullptr;
        if (have_next) {
          act_base_n = act_of(p_n, live_n);
          input_layer(xn, act_base_n);
          ws_publish(my_abar, lane0);
        }
        // ---- output arithmetic of the finished tile -----------------------------------------------------------------
        if (c == 0) {
          float gsel = 0.0f;
#pragma unroll
          for (int o = 0; o < MMF_MAX_SD + 1; ++o)
            if (o == sd) gsel = y[o];
          const float gate = 1.0f / (1.0f + expf(-gsel));
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
          const float v = ll + mw;
          float fused = v;
          if (!first_head) {  // running log-sum-exp kept in logw_out between head phases
            const float mx = fmaxf(prev, v);
            fused = (mx == -INFINITY) ? -INFINITY : mx + logf(expf(prev - mx) + expf(v - mx));
          }
          if (live) P.logw_out[p] = last_head ? lw_in + fused : fused;
        }
        if (!have_next) break;
        tile = tile_n;
        p = p_n;
        live = live_n;
        act_base = act_base_n;
#pragma unroll
        for (int i = 0; i < MMF_MAX_SD; ++i) x[i] = xn[i];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// Pipeline shape, fixed when the library is loaded (environment MMF_TC_VARIANT, read once: no getenv on the launch
// path): 72 = warp-specialised, 4 groups + 2 issuer warps (default); 71 / 74 = same with 1 / 4 issuer warps;
// 41 / 31 = symmetric kernel, every group issues its own MMAs; 42 / 32 = symmetric, CTA pairs.
static int tc_variant() {
  static const int variant = [] {
    const char* env = getenv("MMF_TC_VARIANT");
    return env ? atoi(env) : 72;
  }();
  return variant;
}

// Hardware park time per mbarrier wait attempt in the warp-specialised kernel (ns; MMF_TC_WAIT_HINT, read once).
static uint32_t tc_wait_hint() {
  static const uint32_t hint = [] {
    const char* env = getenv("MMF_TC_WAIT_HINT");
    return env ? (uint32_t)atoi(env) : 20000u;
  }();
  return hint;
}

int launch_particle_chain_tc(const mmf_pf_model* model, int N, int M, const float* states_in, const float* eps,
                             const float* rowbias, const float* logw_in, const float* modw, uint32_t enabled,
                             int precision, float* states_out, float* logw_out, float* ll_out, cudaStream_t stream,
                             int first_chain, float* act_out, int rb_stride) {
  const int variant = tc_variant();
  const bool pair = variant < 70 && (variant % 10) == 2;
  TcParams P;
  P.K = model->num_heads;
  size_t cap = 0;
  for (int c = 0; c <= P.K; ++c) {
    const mmf_chain& src = (c == 0) ? model->dynamics : model->heads[c - 1];
    P.chains[c] = to_dev(src);
    P.images[c] = static_cast<const uint8_t*>(src.w_mma);
    if ((c == 0 && first_chain == 0) || (c > 0 && ((enabled >> (c - 1)) & 1u))) {
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
  P.rb_stride = rb_stride > 0 ? rb_stride : N;
  P.logw_in = logw_in;
  P.modw = modw;
  P.states_out = states_out;
  P.logw_out = logw_out;
  P.ll_out = ll_out;
  for (int i = 0; i < MMF_MAX_SD * MMF_MAX_SD; ++i) P.q[i] = model->q_tril[i];
  P.wait_hint_ns = tc_wait_hint();
  P.first_chain = first_chain;
  P.act_out = act_out;

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
#define MMF_WS_LAUNCH(G, NI)                                                                       \
  do {                                                                                            \
    static thread_local int configured_dev = -1;                                                  \
    static thread_local size_t window = 0;                                                        \
    if (configured_dev != dev) {                                                                  \
      int rc = opt_in_shared_memory(k_particle_chain_ws<G, NI, false>, &window);               \
      if (rc) return rc;                                                                          \
      rc = opt_in_shared_memory(k_particle_chain_ws<G, NI, true>, &window);                    \
      if (rc) return rc;                                                                          \
      configured_dev = dev;                                                                       \
    }                                                                                             \
    MMF_REQUIRE(smem <= window, "tensor-core chain needs %zu B of shared memory (window %zu B)", smem, window); \
    long long units = (tiles + G - 1) / G;                                                        \
    if (units > sms) units = sms;                                                                 \
    if (act_out != nullptr)                                                                       \
      k_particle_chain_ws<G, NI, true><<<(unsigned)units, G * 128 + 128, smem, stream>>>(P);   \
    else                                                                                          \
      k_particle_chain_ws<G, NI, false><<<(unsigned)units, G * 128 + 128, smem, stream>>>(P);  \
    MMF_LAUNCH_CHECK("k_particle_chain_ws");                                                      \
    return MMF_OK;                                                                                \
  } while (0)
  if (variant == 72) MMF_WS_LAUNCH(4, 2);
  if (variant == 71) MMF_WS_LAUNCH(4, 1);
  if (variant == 74) MMF_WS_LAUNCH(4, 4);
#undef MMF_WS_LAUNCH
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

#if MMF_TC_ABLATE == 3
extern "C" int mmf_debug_ws_stamps(unsigned long long* out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, ::g_ws_stamps, sizeof(unsigned long long) * 2 * 16 * 4 * 4);
}
#endif
