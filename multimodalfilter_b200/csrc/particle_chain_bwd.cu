// particle_chain_bwd.cu -- backward of the per-particle measurement heads for BPTT training
// (BASELINE config C4: push crossmodal PF, no resampling, MSE on the estimates).
//
// With the dynamics frozen -- as in every training curriculum of the reference
// (ref: scripts/push_task/train_push.py:154,213) -- the particle states carry no gradient to a trainable
// leaf, so what BPTT needs per filter step is, for every enabled head k and particle:
//     delta_l = d loss / d (pre-activation of layer l),  l = input layer, the 64x64 layers
// from which the host forms dW_l = delta_l^T a_l, db_l = sum delta_l, and d rowbias = sum_m delta_mid.
// The activations a_l come from the forward kernel (k_particle_chain_tc, act_out).
//
// Same machine as the forward: a group of 128 threads owns a 128-particle tile, thread = particle row = TMEM
// lane; delta_l (bf16 hi/lo split, TMEM) x W_l (smem, K-major SW128, here with K = OUTPUT features) ->
// gradient w.r.t. the layer input (TMEM, fp32), tcgen05.mma TS form, 3 MMAs per K step; the epilogue adds the
// residual branch, applies the ReLU mask read from the saved activation, stores delta_l and feeds the next GEMM.
//
// Replaces (for the heads) what autograd does for ref: crossmodal/push_models/pf.py:91-109 in
// torchfilter.train.train_filter (A.7; call site ref: crossmodal/train_helpers.py:155-162).
#include "tc_common.cuh"

namespace mmf {

// ---- backward operand image of one chain ----------------------------------------------------------------
//   [layer 0 hi | layer 0 lo | ... | layer L-1 hi | layer L-1 lo]   B[n = input feature][k = output feature] = W[k][n]
//   [out_W[out_dim][64]]                                            fp32 (gradient of the output layer, CUDA cores)
__host__ __device__ inline size_t bwd_image_tiles_bytes(const ChainDev& c) { return (size_t)chain_layers(c) * 2 * TILE_B; }
__host__ __device__ inline size_t bwd_image_bytes(const ChainDev& c) {
  const size_t b = bwd_image_tiles_bytes(c) + sizeof(float) * (size_t)(c.out_dim * U);
  return (b + 1023) & ~(size_t)1023;
}

__global__ void k_pack_chain_bwd(ChainDev ch, uint8_t* __restrict__ dst) {
  const int L = chain_layers(ch);
  const float* w = ch.w;
  const int off_first_res = ch.in_dim * U + U;
  const int off_mid = off_first_res + ch.n_pre * RES_FLOATS;
  const int off_post = off_mid + U * U;
  const int off_out = off_post + ch.n_post * RES_FLOATS;
  float* fdst = reinterpret_cast<float*>(dst + bwd_image_tiles_bytes(ch));
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int layer = 0; layer < L; ++layer) {
    const float* Wt;  // packed transposed weights: Wt[k_in][j_out] = W[j_out][k_in]
    if (layer < 2 * ch.n_pre) {
      Wt = w + off_first_res + (layer >> 1) * RES_FLOATS + (layer & 1) * (U * U + U);
    } else if (layer == 2 * ch.n_pre) {
      Wt = w + off_mid;
    } else {
      const int rel = layer - 2 * ch.n_pre - 1;
      Wt = w + off_post + (rel >> 1) * RES_FLOATS + (rel & 1) * (U * U + U);
    }
    uint8_t* hi = dst + (size_t)layer * 2 * TILE_B;
    uint8_t* lo = hi + TILE_B;
    for (int e = tid; e < U * U; e += nth) {
      const int n = e / U, k = e % U;  // B[n = in][k = out] = W[k][n] = Wt[n][k]
      const float v = Wt[n * U + k];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      const int o = sw128_offset(n, k);
      *reinterpret_cast<__nv_bfloat16*>(hi + o) = h;
      *reinterpret_cast<__nv_bfloat16*>(lo + o) = l;
    }
  }
  for (int e = tid; e < ch.out_dim * U; e += nth) fdst[e] = w[off_out + e];
}

size_t chain_bwd_bytes(const mmf_chain* chain) { return bwd_image_bytes(to_dev(*chain)); }

int pack_chain_bwd(const mmf_chain* chain, void* dst, cudaStream_t stream) {
  MMF_REQUIRE(chain->w != nullptr, "pack_chain_bwd: chain has no fp32 weights");
  MMF_REQUIRE(((uintptr_t)dst & 15) == 0, "pack_chain_bwd: destination must be 16-byte aligned");
  k_pack_chain_bwd<<<16, 256, 0, stream>>>(to_dev(*chain), static_cast<uint8_t*>(dst));
  MMF_LAUNCH_CHECK("k_pack_chain_bwd");
  return MMF_OK;
}

struct BwdParams {
  ChainDev chains[MMF_MAX_HEADS];
  const uint8_t* images[MMF_MAX_HEADS];
  int K;
  uint32_t enabled;
  long long total;
  size_t image_cap;
  const float* act;     // (K, L+1, P, 64): activation feeding GEMM layer l (index l), the output layer (index L)
  const float* d_ll;    // (K, P): d loss / d log-likelihood of head k
  float* delta_out;     // (K, L+1, P, 64): delta of GEMM layer l (index l), of the input layer (index L)
};

enum { LK_RES_A = 0, LK_RES_B = 1, LK_MID = 2 };
__device__ __forceinline__ int layer_kind(const ChainDev& ch, int layer) {
  const int mid_at = 2 * ch.n_pre;
  if (layer == mid_at) return LK_MID;
  const int rel = layer < mid_at ? layer : layer - mid_at - 1;
  return (rel & 1) ? LK_RES_B : LK_RES_A;
}

// One 16-column slice of "next delta = (grad [+ skip]) * relu'(activation)": stores it, keeps the skip branch,
// and (feed != 0) writes it as the next A operand.
// act_row / delta_row point at this row's float4 of column chunk 0 in the chunk-major planes (see store_act_chunk in
// particle_chain_tc.cu); stride4 = rows * 4 floats separates the column chunks.
__device__ __forceinline__ void delta_chunk(float2 (&v)[8], const float* __restrict__ act_row, float* __restrict__ delta_row,
                                            size_t stride4, float2 (&gres)[U / 2], int chunk, bool add_res, bool mask,
                                            bool set_res, bool feed, uint32_t tAhi, uint32_t tAlo) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 a = mask ? __ldg(reinterpret_cast<const float4*>(act_row + (size_t)(chunk * 4 + q) * stride4))
                          : make_float4(1.f, 1.f, 1.f, 1.f);
    float2 g0 = v[2 * q], g1 = v[2 * q + 1];
    if (add_res) {
      g0 = __fadd2_rn(g0, gres[chunk * 8 + 2 * q]);
      g1 = __fadd2_rn(g1, gres[chunk * 8 + 2 * q + 1]);
    }
    g0.x = a.x > 0.0f ? g0.x : 0.0f;
    g0.y = a.y > 0.0f ? g0.y : 0.0f;
    g1.x = a.z > 0.0f ? g1.x : 0.0f;
    g1.y = a.w > 0.0f ? g1.y : 0.0f;
    v[2 * q] = g0;
    v[2 * q + 1] = g1;
    if (set_res) {
      gres[chunk * 8 + 2 * q] = g0;
      gres[chunk * 8 + 2 * q + 1] = g1;
    }
  }
  if (delta_row != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(delta_row + (size_t)(chunk * 4 + q) * stride4) =
          make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
  }
  if (feed) store_a_chunk<false>(v, tAhi, tAlo, chunk, false);
}

template <int TC_GROUPS>
__global__ void __launch_bounds__(TC_GROUPS * 128, 1) k_head_chain_bwd(const __grid_constant__ BwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem + P.image_cap);
  uint64_t* gbar = wbar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbar + TC_MAX_GROUPS);

  const int tid = threadIdx.x, row = tid & 127;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = warp >> 2;
  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int i = 0; i < TC_GROUPS; ++i) mbar_init(gbar + i, 1);
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
  const uint32_t grp_cols = tmem_base + g * 128;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tD = grp_cols + lane_off;
  const uint32_t tAhi = tD + 64, tAlo = tD + 96;
  uint32_t wphase = 0, gphase = 0;
  const long long tiles = (P.total + 127) / 128;
  constexpr uint32_t IDESC = make_idesc(64, 128);
  const size_t plane = (size_t)P.total * U;

  for (int k = 0; k < P.K; ++k) {
    if (!((P.enabled >> k) & 1u)) continue;
    const ChainDev ch = P.chains[k];
    const int L = chain_layers(ch);
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)bwd_image_bytes(ch);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(wbar, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768) {
        const uint32_t n = bytes - off < 32768 ? bytes - off : 32768;
        bulk_g2s(smem + off, P.images[k] + off, n, wbar);
      }
    }
    mbar_wait(wbar, wphase);
    wphase ^= 1;
    const uint32_t tiles_addr = smem_u32(smem);
    const float* out_W = reinterpret_cast<const float*>(smem + bwd_image_tiles_bytes(ch));  // row 0: the head's scalar output

    for (long long tile = (long long)blockIdx.x * TC_GROUPS + g; tile < tiles; tile += (long long)gridDim.x * TC_GROUPS) {
      const long long p_raw = tile * 128 + row;
      const bool live = p_raw < P.total;
      const long long p = live ? p_raw : P.total - 1;
      const float* act_base = P.act + (size_t)k * (L + 1) * P.total * U + (size_t)p * 4;
      float* delta_base = P.delta_out + (size_t)k * (L + 1) * P.total * U + (size_t)p * 4;
      const size_t stride4 = (size_t)P.total * 4;
      const float dll = live ? P.d_ll[(size_t)k * P.total + p] : 0.0f;

      // ---- output layer: grad w.r.t. its input = dll * out_W[0]; delta of layer L-1 -------------------------
      float2 gres[U / 2];
#pragma unroll
      for (int j = 0; j < U / 2; ++j) gres[j] = make_float2(0.0f, 0.0f);
      {
        const int kind = layer_kind(ch, L - 1);
        const bool mask = (kind != LK_MID) || ch.mid_relu;
        const float4* w4 = reinterpret_cast<const float4*>(out_W);
#pragma unroll
        for (int chunk = 0; chunk < 4; ++chunk) {
          float2 v[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 w = w4[chunk * 4 + q];
            v[2 * q] = make_float2(w.x * dll, w.y * dll);
            v[2 * q + 1] = make_float2(w.z * dll, w.w * dll);
          }
          delta_chunk(v, act_base + (size_t)L * plane, live ? delta_base + (size_t)(L - 1) * plane : nullptr, stride4, gres,
                      chunk, false, mask, kind == LK_RES_B, true, tAhi, tAlo);
        }
      }

      // ---- 64x64 layers, last to first -----------------------------------------------------------------------------
      for (int layer = L - 1; layer >= 0; --layer) {
        tc_wait_st();
        tc_fence_before();
        group_bar(1 + g, 128);
        if ((warp & 3) == 0 && elect_one_sync()) {
          tc_fence_after();
          const uint32_t hi_addr = tiles_addr + (uint32_t)layer * 2 * TILE_B;
          const uint64_t bhi = make_b_desc(hi_addr), blo = make_b_desc(hi_addr + TILE_B);
          const uint32_t a_hi = grp_cols + 64, a_lo = grp_cols + 96;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) mma_ts(grp_cols, a_hi + kk * 8, bhi + (uint64_t)(kk * 2), IDESC, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) mma_ts(grp_cols, a_hi + kk * 8, blo + (uint64_t)(kk * 2), IDESC, 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) mma_ts(grp_cols, a_lo + kk * 8, bhi + (uint64_t)(kk * 2), IDESC, 1);
          tc_commit(gbar + g);
        }
        mbar_wait(gbar + g, gphase);
        gphase ^= 1;
        tc_fence_after();

        // D = gradient w.r.t. the input of `layer`; turn it into the delta of the producer of that input
        const int kind = layer_kind(ch, layer);
        const bool add_res = (kind == LK_RES_A);  // the block input also feeds the skip connection
        const int prod = layer - 1;               // producer: GEMM layer `prod`, or the input layer when < 0
        const int pkind = prod >= 0 ? layer_kind(ch, prod) : LK_RES_A;
        const bool mask = prod < 0 || pkind != LK_MID || ch.mid_relu;
        const bool set_res = prod >= 0 && pkind == LK_RES_B;
        float* drow = live ? delta_base + (size_t)(prod >= 0 ? prod : L) * plane : nullptr;
        const float* arow = act_base + (size_t)layer * plane;
#pragma unroll
        for (int chunk = 0; chunk < 4; ++chunk) {
          uint32_t d[16];
          tmem_ld16(tD + chunk * 16, d);
          tc_wait_ld();
          float2 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = make_float2(__uint_as_float(d[2 * j]), __uint_as_float(d[2 * j + 1]));
          delta_chunk(v, arow, drow, stride4, gres, chunk, add_res, mask, set_res, prod >= 0, tAhi, tAlo);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int launch_head_chain_bwd(const mmf_pf_model* model, int N, int M, const float* act, const float* d_ll, uint32_t enabled,
                          float* delta_out, cudaStream_t stream) {
  BwdParams P;
  P.K = model->num_heads;
  size_t cap = 0;
  int L0 = -1;
  for (int k = 0; k < P.K; ++k) {
    P.chains[k] = to_dev(model->heads[k]);
    P.images[k] = static_cast<const uint8_t*>(model->heads[k].w_bwd);
    if ((enabled >> k) & 1u) {
      MMF_REQUIRE(P.images[k] != nullptr, "head %d has no backward operand image: call mmf_pack_chain_bwd first", k);
      const size_t b = bwd_image_bytes(P.chains[k]);
      cap = b > cap ? b : cap;
    }
    const int L = chain_layers(P.chains[k]);
    MMF_REQUIRE(L0 < 0 || L == L0, "heads_backward: all heads must have the same depth");
    L0 = L;
  }
  P.image_cap = cap;
  P.enabled = enabled;
  P.total = (long long)N * M;
  P.act = act;
  P.d_ll = d_ll;
  P.delta_out = delta_out;
  const size_t smem = cap + 1024;
  static thread_local int configured_dev = -1;
  static thread_local size_t window = 0;
  int dev = 0, sms = 148;
  MMF_CUDA(cudaGetDevice(&dev));
  MMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (configured_dev != dev) {
    int rc = opt_in_shared_memory(k_head_chain_bwd<4>, &window);
    if (rc) return rc;
    configured_dev = dev;
  }
  MMF_REQUIRE(smem <= window, "heads_backward needs %zu B of shared memory (window %zu B)", smem, window);
  const long long tiles = (P.total + 127) / 128;
  long long grid = (tiles + 3) / 4;
  if (grid > sms) grid = sms;
  k_head_chain_bwd<4><<<(int)grid, 4 * 128, smem, stream>>>(P);
  MMF_LAUNCH_CHECK("k_head_chain_bwd");
  return MMF_OK;
}

}  // namespace mmf

// ---- weight gradients: dW[k][l] = delta[k][l]^T act[k][l]  (64 x 64, reduction over the N*M rows) ----------------
// HBM-bound (2 x 256 B per row per layer against 4096 FMAs).  The rows are split over CTAs; every CTA writes its partial
// 64 x 64 tile (and partial column sums) to a workspace and a second kernel adds the partials IN A FIXED ORDER, so the
// gradients are bit-reproducible from run to run (hard part 4: no floating-point atomics).
namespace mmf {

// ---- dW on the tensor cores -----------------------------------------------------------------------------------
// dW = delta^T act is a (64 x rows) x (rows x 64) contraction with the ROWS as K.  Both operands are "MN-major" for
// the MMA (the 64 features are the contiguous dimension), which tcgen05 takes directly from shared memory: a tile of
// 64 rows is converted to bf16 hi/lo on the fly and written as 8(K) x 8(MN) core matrices,
//   element (m, k) at (m / 8) * 1024 + (k / 8) * 128 + (k % 8) * 16 + (m % 8) * 2   [SBO = 1024, LBO = 128],
// with the hi and lo halves STACKED along MN: A = [delta_hi ; delta_lo] (M = 128), B = [act_hi | act_lo] (N = 128),
// so ONE M=128 N=128 K=16 MMA per 16 rows yields all four split products; the accumulator (128 lanes x 128 columns,
// fp32, TMEM) lives across the CTA's whole row range and its quadrants are summed on the way out.
// 0.92 ms (FFMA2 kernel above, 35 % of the fp32 FMA peak) -> HBM-bound.
constexpr int DWT_ROWS = 64;                      // rows (K) per tile
constexpr int DWT_OPERAND_B = 16 * 1024;          // [16 MN groups][8 K groups][8][8] bf16
constexpr int DWT_THREADS = 256;

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {  // MN-major, no swizzle: LBO 128 B, SBO 1024 B
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ss_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// 8 fp32 -> 8 bf16 hi (rn) + 8 bf16 lo (rn of the exact remainder), 16 bytes each
__device__ __forceinline__ void split8_rn(const float4& x0, const float4& x1, uint4& hi4, uint4& lo4) {
  const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
    float rx, ry;
    asm("{\n\t"
        ".reg .b16 l, h, m1;\n\t"
        "mov.b32 {l, h}, %2;\n\t"
        "mov.b16 m1, 0xBF80;\n\t"
        "fma.rn.f32.bf16 %0, l, m1, %3;\n\t"
        "fma.rn.f32.bf16 %1, h, m1, %4;\n\t"
        "}"
        : "=f"(rx), "=f"(ry)
        : "r"(hi[j]), "f"(v[2 * j]), "f"(v[2 * j + 1]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo[j]) : "f"(ry), "f"(rx));
  }
  hi4 = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  lo4 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// part_dW: [head][layer][chunk][64][64], part_db: [head][layer][chunk][64]   (chunk = blockIdx.x)
__global__ void __launch_bounds__(DWT_THREADS, 1) k_heads_dw_tc(const float* __restrict__ act, const float* __restrict__ delta,
                                                                float* __restrict__ part_dW, float* __restrict__ part_db,
                                                                long long P, int planes_per_head, int L, int rows_per_cta) {
  __shared__ float s_colsum[DWT_THREADS / 32][2][8];
  extern __shared__ __align__(1024) uint8_t smem[];  // [buffer 0: A | B][buffer 1: A | B] + barriers
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 4 * DWT_OPERAND_B);  // [buffer]: the MMAs reading it are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int layer = blockIdx.y, head = blockIdx.z;
  const size_t plane = ((size_t)head * planes_per_head + layer) * (size_t)P * U;
  const float* A = act + plane;
  const float* D = delta + plane;
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  const long long row1 = row0 + rows_per_cta < P ? row0 + rows_per_cta : P;
  if (tid == 0) {
    mbar_init(done, 1);
    mbar_init(done + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // D fp32, A/B bf16, both MN-major (bits 15, 16), N = 128, M = 128
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

  // unit u = (column group of 8, row): this thread owns units tid and tid + 256 of every tile
  float colsum[2][8];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) colsum[q][j] = 0.0f;
  uint32_t phase[2] = {0, 0};
  long long t = 0;
  for (long long r = row0; r < row1; r += DWT_ROWS, ++t) {
    const int buf = (int)(t & 1);
    uint8_t* sA = smem + buf * 2 * DWT_OPERAND_B;
    uint8_t* sB = sA + DWT_OPERAND_B;
    float4 dv[2][2], av[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int u = tid + q * DWT_THREADS;
      const int m8 = u >> 6, rr = u & 63;
      const bool ok = r + rr < row1;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const size_t off = ((size_t)(2 * m8 + h) * P + (size_t)(r + rr)) * 4;
        dv[q][h] = ok ? __ldg(reinterpret_cast<const float4*>(D + off)) : make_float4(0, 0, 0, 0);
        av[q][h] = ok ? __ldg(reinterpret_cast<const float4*>(A + off)) : make_float4(0, 0, 0, 0);
      }
    }
    if (t >= 2) {  // the MMAs of tile t - 2 read this buffer
      mbar_wait(done + buf, phase[buf]);
      phase[buf] ^= 1;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int u = tid + q * DWT_THREADS;
      const int m8 = u >> 6, rr = u & 63;
      const uint32_t core = (uint32_t)(rr >> 3) * 128 + (uint32_t)(rr & 7) * 16;
      uint4 hi4, lo4;
      split8_rn(dv[q][0], dv[q][1], hi4, lo4);
      *reinterpret_cast<uint4*>(sA + (uint32_t)m8 * 1024 + core) = hi4;
      *reinterpret_cast<uint4*>(sA + (uint32_t)(8 + m8) * 1024 + core) = lo4;
      split8_rn(av[q][0], av[q][1], hi4, lo4);
      *reinterpret_cast<uint4*>(sB + (uint32_t)m8 * 1024 + core) = hi4;
      *reinterpret_cast<uint4*>(sB + (uint32_t)(8 + m8) * 1024 + core) = lo4;
      colsum[q][0] += dv[q][0].x; colsum[q][1] += dv[q][0].y; colsum[q][2] += dv[q][0].z; colsum[q][3] += dv[q][0].w;
      colsum[q][4] += dv[q][1].x; colsum[q][5] += dv[q][1].y; colsum[q][6] += dv[q][1].z; colsum[q][7] += dv[q][1].w;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA
    __syncthreads();
    if (warp == 0 && elect_one_sync()) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
#pragma unroll
      for (int ks = 0; ks < DWT_ROWS / 16; ++ks)
        mma_ss_f16(tmem_base, make_desc_mn(a_addr + ks * 256), make_desc_mn(b_addr + ks * 256), IDESC, (t > 0 || ks > 0) ? 1u : 0u);
      tc_commit(done + buf);
    }
  }
  // drain: the last one or two commits
  for (int b = 0; b < 2; ++b) {
    const long long used = (t + 1 - b) / 2;  // tiles that went through buffer b
    if (used > 0) mbar_wait(done + b, phase[b]);
  }
  tc_fence_after();
  if (t > 0 && warp < 4) {
    // lanes 0..63: rows of delta_hi (quadrants hi*hi | hi*lo), lanes 64..127: rows of delta_lo (lo*hi | lo*lo)
    const int lane_row = warp * 32 + (tid & 31);
    const int j = lane_row & 63;
    // rows j (delta_hi) and 64 + j (delta_lo) of the accumulator both belong to output row j: the hi half goes to
    // slot 0 of the partial tile pair, the lo half to slot 1; the reduce kernel adds them in that order
    float* out = part_dW + ((((size_t)head * L + layer) * gridDim.x + blockIdx.x) * 2 + (lane_row >> 6)) * U * U + (size_t)j * U;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 64; c += 16) {
      uint32_t x[16], y[16];
      tmem_ld16(taddr + c, x);
      tmem_ld16(taddr + 64 + c, y);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        *reinterpret_cast<float4*>(out + c + i) =
            make_float4(__uint_as_float(x[i]) + __uint_as_float(y[i]), __uint_as_float(x[i + 1]) + __uint_as_float(y[i + 1]),
                        __uint_as_float(x[i + 2]) + __uint_as_float(y[i + 2]), __uint_as_float(x[i + 3]) + __uint_as_float(y[i + 3]));
    }
  } else if (t == 0 && warp < 4) {  // a CTA without rows still owns its slots of the workspace
    const int lane_row = warp * 32 + (tid & 31);
    float* out = part_dW + ((((size_t)head * L + layer) * gridDim.x + blockIdx.x) * 2 + (lane_row >> 6)) * U * U + (size_t)(lane_row & 63) * U;
    for (int c = 0; c < 64; c += 4) *reinterpret_cast<float4*>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // bias gradient: reduce the per-thread column sums over the 32 rows of the warp; the two warps that share a column
  // group (rows 0..31 and 32..63 of the tiles) are added in warp order
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float v = colsum[q][jj];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) s_colsum[warp][q][jj] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < U) {  // column tid = 8 m8 + jj; unit u = tid' + 256 q has m8 = u >> 6: q = m8 >> 2, warps 2 (m8 & 3), 2 (m8 & 3) + 1
    const int m8 = tid >> 3, jj = tid & 7, q = m8 >> 2, w0 = 2 * (m8 & 3);
    part_db[(((size_t)head * L + layer) * gridDim.x + blockIdx.x) * U + tid] = s_colsum[w0][q][jj] + s_colsum[w0 + 1][q][jj];
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
}

// Gradients of the two thin layers at the ends of a head, one pass over plane L of delta and act:
//   g_in[k][j][d] = sum_p delta_in[p][j] * x[p][d]   (input layer weight, x = particle states)
//   db[k][L][j]   = sum_p delta_in[p][j]             (input layer bias)
//   g_out[k][j]   = sum_p d_ll[p] * a_L[p][j]        (output layer weight)
// thread = (column chunk c4, row lane): 16 consecutive rows of one chunk per half warp (coalesced 256 B).
__global__ void __launch_bounds__(256) k_heads_edge_grads(const float* __restrict__ act, const float* __restrict__ delta,
                                                          const float* __restrict__ x, const float* __restrict__ d_ll,
                                                          float* __restrict__ part_edge, long long P, int L, int sd,
                                                          int rows_per_cta) {  // part_edge: [head][chunk][64][2 + MAX_SD]
  const int head = blockIdx.y;
  const size_t plane = ((size_t)head * (L + 1) + L) * (size_t)P * U;
  const float* A = act + plane;
  const float* D = delta + plane;
  const int c4 = threadIdx.x >> 4, lane16 = threadIdx.x & 15;
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  const long long row1 = row0 + rows_per_cta < P ? row0 + rows_per_cta : P;
  float gi[4][MMF_MAX_SD], gb[4] = {0.f, 0.f, 0.f, 0.f}, go[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int d = 0; d < MMF_MAX_SD; ++d) gi[j][d] = 0.0f;
  for (long long p = row0 + lane16; p < row1; p += 16) {
    const float4 dv = __ldg(reinterpret_cast<const float4*>(D + ((size_t)c4 * P + (size_t)p) * 4));
    const float4 av = __ldg(reinterpret_cast<const float4*>(A + ((size_t)c4 * P + (size_t)p) * 4));
    const float w = __ldg(d_ll + (size_t)head * P + p);
    const float dj[4] = {dv.x, dv.y, dv.z, dv.w}, aj[4] = {av.x, av.y, av.z, av.w};
    float xs[MMF_MAX_SD];
#pragma unroll
    for (int d = 0; d < MMF_MAX_SD; ++d) xs[d] = d < sd ? __ldg(x + (size_t)p * sd + d) : 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gb[j] += dj[j];
      go[j] = fmaf(w, aj[j], go[j]);
#pragma unroll
      for (int d = 0; d < MMF_MAX_SD; ++d) gi[j][d] = fmaf(dj[j], xs[d], gi[j][d]);
    }
  }
  // reduce over the 16 row lanes of this column chunk (they share a half warp)
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gb[j] += __shfl_xor_sync(0xffffffffu, gb[j], o);
      go[j] += __shfl_xor_sync(0xffffffffu, go[j], o);
#pragma unroll
      for (int d = 0; d < MMF_MAX_SD; ++d) gi[j][d] += __shfl_xor_sync(0xffffffffu, gi[j][d], o);
    }
  }
  if (lane16 == 0) {
    float* out = part_edge + ((size_t)head * gridDim.x + blockIdx.x) * U * (2 + MMF_MAX_SD);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* o = out + (size_t)(c4 * 4 + j) * (2 + MMF_MAX_SD);
      o[0] = gb[j];
      o[1] = go[j];
#pragma unroll
      for (int d = 0; d < MMF_MAX_SD; ++d) o[2 + d] = gi[j][d];
    }
  }
}

// second pass: adds the partials of every output element in chunk order (fixed => bit-reproducible)
__global__ void k_heads_grads_reduce(const float* __restrict__ part_dW, const float* __restrict__ part_db,
                                     const float* __restrict__ part_edge, int K, int L, int sd, int dw_chunks,
                                     int edge_chunks, float* __restrict__ dW, float* __restrict__ db,
                                     float* __restrict__ g_in, float* __restrict__ g_out) {
  const long long n_dw = (long long)K * L * U * U, n_db = (long long)K * (L + 1) * U;
  const long long n_in = (long long)K * U * sd, n_out = (long long)K * U;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_dw + n_db + n_in + n_out;
       e += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    if (e < n_dw) {
      const long long hl = e / (U * U), ij = e % (U * U);
      const float* p = part_dW + (size_t)hl * dw_chunks * 2 * U * U + ij;
      for (int c = 0; c < 2 * dw_chunks; ++c) s += p[(size_t)c * U * U];
      dW[e] = s;
    } else if (e < n_dw + n_db) {
      const long long r = e - n_dw, head = r / ((L + 1) * U), layer = (r / U) % (L + 1), col = r % U;
      if (layer < L) {
        const float* p = part_db + ((size_t)(head * L + layer) * dw_chunks) * U + col;
        for (int c = 0; c < dw_chunks; ++c) s += p[(size_t)c * U];
      } else {
        const float* p = part_edge + ((size_t)head * edge_chunks * U + col) * (2 + MMF_MAX_SD);
        for (int c = 0; c < edge_chunks; ++c) s += p[(size_t)c * U * (2 + MMF_MAX_SD)];
      }
      db[r] = s;
    } else if (e < n_dw + n_db + n_in) {
      const long long r = e - n_dw - n_db, head = r / (U * sd), col = (r / sd) % U, d = r % sd;
      const float* p = part_edge + ((size_t)head * edge_chunks * U + col) * (2 + MMF_MAX_SD) + 2 + d;
      for (int c = 0; c < edge_chunks; ++c) s += p[(size_t)c * U * (2 + MMF_MAX_SD)];
      g_in[r] = s;
    } else {
      const long long r = e - n_dw - n_db - n_in, head = r / U, col = r % U;
      const float* p = part_edge + ((size_t)head * edge_chunks * U + col) * (2 + MMF_MAX_SD) + 1;
      for (int c = 0; c < edge_chunks; ++c) s += p[(size_t)c * U * (2 + MMF_MAX_SD)];
      g_out[r] = s;
    }
  }
}

// row partition of the weight-gradient kernels (a function of the device's SM count only)
struct DwPartition {
  long long dw_chunks, dw_rows, edge_chunks, edge_rows;
};
static DwPartition dw_partition(int K, int L, long long P, int sms) {
  DwPartition q;
  // tensor-core dW kernel: 3 CTAs of 64 KB fit an SM
  q.dw_chunks = ((long long)sms * 3 + (long long)K * L - 1) / ((long long)K * L);
  q.dw_rows = (P + q.dw_chunks - 1) / q.dw_chunks;
  q.dw_rows = ((q.dw_rows + DWT_ROWS - 1) / DWT_ROWS) * DWT_ROWS;
  q.dw_chunks = (P + q.dw_rows - 1) / q.dw_rows;
  q.edge_chunks = ((long long)sms * 4 + K - 1) / K;
  q.edge_rows = (P + q.edge_chunks - 1) / q.edge_chunks;
  q.edge_rows = ((q.edge_rows + 15) / 16) * 16;
  q.edge_chunks = (P + q.edge_rows - 1) / q.edge_rows;
  return q;
}
static int current_sms(int* sms) {
  int dev = 0;
  *sms = 148;
  MMF_CUDA(cudaGetDevice(&dev));
  MMF_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  return MMF_OK;
}

size_t heads_dw_workspace_bytes(int K, int L, long long P) {
  int sms = 148;
  if (P <= 0 || current_sms(&sms) != MMF_OK) return 0;
  const DwPartition q = dw_partition(K, L, P, sms);
  const size_t floats = (size_t)K * L * q.dw_chunks * (2 * U * U + U) + (size_t)K * q.edge_chunks * U * (2 + MMF_MAX_SD);
  return floats * sizeof(float);
}

int launch_heads_dw(int K, int L, long long P, int sd, const float* act, const float* delta, const float* x,
                    const float* d_ll, float* dW, float* db, float* g_in, float* g_out, void* workspace,
                    cudaStream_t stream) {
  if (P == 0) return MMF_OK;
  int dev = 0, sms = 148;
  MMF_CUDA(cudaGetDevice(&dev));
  int rc = current_sms(&sms);
  if (rc) return rc;
  MMF_REQUIRE(workspace != nullptr && ((uintptr_t)workspace & 15) == 0,
              "heads_weight_grads: pass a 16-byte aligned workspace of mmf_pf_heads_weight_grads_workspace_bytes() bytes");
  const DwPartition q = dw_partition(K, L, P, sms);
  float* part_dW = static_cast<float*>(workspace);
  float* part_db = part_dW + (size_t)K * L * q.dw_chunks * 2 * U * U;
  float* part_edge = part_db + (size_t)K * L * q.dw_chunks * U;
  {
    const size_t smem = 4 * DWT_OPERAND_B + 64;
    static thread_local int configured_dev = -1;
    static thread_local size_t window = 0;
    if (configured_dev != dev) {
      rc = opt_in_shared_memory(k_heads_dw_tc, &window);
      if (rc) return rc;
      configured_dev = dev;
    }
    MMF_REQUIRE(smem <= window, "heads_weight_grads needs %zu B of shared memory (window %zu B)", smem, window);
    k_heads_dw_tc<<<dim3((unsigned)q.dw_chunks, (unsigned)L, (unsigned)K), DWT_THREADS, smem, stream>>>(
        act, delta, part_dW, part_db, P, L + 1, L, (int)q.dw_rows);
    MMF_LAUNCH_CHECK("k_heads_dw_tc");
  }
  k_heads_edge_grads<<<dim3((unsigned)q.edge_chunks, (unsigned)K), 256, 0, stream>>>(act, delta, x, d_ll, part_edge, P, L, sd,
                                                                                   (int)q.edge_rows);
  MMF_LAUNCH_CHECK("k_heads_edge_grads");
  const long long outputs = (long long)K * L * U * U + (long long)K * (L + 1) * U + (long long)K * U * (sd + 1);
  k_heads_grads_reduce<<<(unsigned)((outputs + 255) / 256), 256, 0, stream>>>(part_dW, part_db, part_edge, K, L, sd,
                                                                            (int)q.dw_chunks, (int)q.edge_chunks, dW, db, g_in,
                                                                            g_out);
  MMF_LAUNCH_CHECK("k_heads_grads_reduce");
  return MMF_OK;
}

}  // namespace mmf
