// pf_loop_small.cu -- R1 for SMALL problems: the whole T-step particle-filter recursion in ONE kernel launch
// (replaces the Python loop of A.2 `Filter.forward_loop` over A.3 `ParticleFilter.forward`,
// call site ref: crossmodal/eval_helpers.py:139-142).
//
// At BASELINE config C1 (32 trajectories x 30 particles) the per-step kernels are pure latency: 2 launches per step, each
// a handful of CTAs, 49 us per filter step.  Every reduction of the recursion is inside ONE trajectory, so a CTA that owns
// a trajectory never has to talk to another CTA: here CTA n carries trajectory n through all T steps by itself --
//   predict (dynamics chain) -> K measurement heads -> fusion -> normalise -> estimate -> resample + gather
// with the particle set in shared / L1-resident global memory and no grid-wide synchronisation at all.
//
// Two variants (launch_pf_loop_small picks by precision):
//   k_pf_loop_small      MMF_PREC_FP32: the chains on the CUDA cores in exactly the accumulation order of
//                        particle_chain_ffma.cu, so the one-launch and the per-step fp32 paths produce identical bits.
//                        NW (default 8) warps, warp = (particle chunk of 32, block of OPW output features), lane = particle;
//                        a lane keeps its particle's 64 input activations in registers, the weights of the current layer
//                        are warp-broadcast LDS.128 reads, two FMAs per FFMA2.  A layer's weights (16.6 KB, fp32 pack of
//                        mmf_chain) are fetched by cp.async into a 4-deep ring three layers ahead; one __syncthreads per
//                        layer exchanges the activations (two ping-pong buffers, layout [feature / 4][particle][4]).
//   k_pf_loop_small_mma  MMF_PREC_BF16X3 / BF16: the layers on mma.sync with split bf16 operands (further down).
// Normalise / estimate / resample is nr_trajectory<32> of normalize_resample.cuh -- for M <= 32 its register-resident
// twin nr_small -- run by warp 0: the pinned arithmetic and therefore the indices are those of k_normalize_resample bit
// for bit.
#include "normalize_resample.cuh"
#include "tc_common.cuh"

namespace mmf {

constexpr int LS_NBUF = 4;          // weight ring: the copy of a layer is issued three layers before its arithmetic
constexpr int LS_WMAX = U * U + U;  // floats of the largest stage (a residual half: 64x64 matrix + bias)
constexpr int LS_MAX_STAGES = (1 + MMF_MAX_HEADS) * 24;

enum { LS_IN = 0, LS_RES_A = 1, LS_RES_B = 2, LS_MID = 3, LS_OUT = 4 };

struct LoopParams {
  ChainDev chains[1 + MMF_MAX_HEADS];
  int K;
  uint32_t enabled;
  int sd, N, M, T;
  float* states;      // (N, M, sd)  in: the set entering step 0, out: the set after step T-1
  float* logw;        // (N, M)
  float* states_ws;   // (N, M, sd)  moved particles
  float* logw_ws;     // (N, M)      un-normalised log-weights
  const float* eps;       // (T, N*M, sd)
  const float* rowbias;   // (1+K, T*N, 64)
  const float* modw;      // (T, N, K) or null
  const double* uniforms; // (T, N, M) / (T, N) / null
  float* est_out;         // (T, N, sd)
  int estimation, mode;
  float q[MMF_MAX_SD * MMF_MAX_SD];
  const uint8_t* images[1 + MMF_MAX_HEADS];  // tensor-core variant: the chains' bf16 hi/lo operand images (mmf_chain.w_mma)
  int single_pass;                           // MMF_PREC_BF16: hi x hi products only
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Normalise / estimate / resample of ONE trajectory with M <= 32 particles, in registers: lane j owns particle j and draw j.
// Same operations on the same values in the same order as nr_trajectory<32> (normalize_resample.cuh) for the cases it
// covers -- hard resampling that keeps the particle count, modes NONE / MULTINOMIAL_STRICT / SYSTEMATIC_STRICT -- so the
// outputs are identical bit for bit (tests/test_gpu_parity.py::test_one_launch_forward_loop_matches_the_per_step_kernels
// compares against the per-step kernels, which run nr_trajectory).  What it sheds: the generic path's shared-memory
// round trips, guide table and padded loops, ~4x the instructions at this size (ncu: a third of a C1 step).
//   * sum exp / estimate: every `for (i = lane; i < M; i += 32)` loop of the generic code has at most one iteration here;
//   * strict CDF: c_j = fl(c_{j-1} + e_j), the sequential sum, formed redundantly by all lanes from shuffled e_k;
//   * inverse CDF: idx = #{k : c_k < c*} with c* = cdf_threshold(total, u), the exact form of the pinned predicate.
// All pointers of R already address this trajectory (the generic function is called with n = 0 the same way).
__device__ __forceinline__ bool nr_small_applies(const ResampleParams& R) {
  return R.M <= 32 && R.M_out == R.M && !(R.alpha < 1.0f) && R.logits_in == nullptr && R.logw_norm_out == nullptr &&
         R.logits_out == nullptr && R.idx_out == nullptr &&
         (R.mode == MMF_RESAMPLE_NONE || R.mode == MMF_RESAMPLE_MULTINOMIAL_STRICT || R.mode == MMF_RESAMPLE_SYSTEMATIC_STRICT);
}

__device__ __forceinline__ void nr_small(const ResampleParams& R, int lane) {
  const int M = R.M, sd = R.sd;
  const bool mine = lane < M;
  const bool resample = R.mode != MMF_RESAMPLE_NONE;
  // ---- normalise ---------------------------------------------------------------------------------------------------------
  const float lraw = mine ? R.logw_unnorm[lane] : -INFINITY;
  const float mx = warp_max(lraw);
  const float shift = (mx == -INFINITY || mx == INFINITY) ? 0.0f : mx;
  float s = 0.0f;
  if (mine) s += expf(lraw - shift);
  s = warp_sum(s);
  const float lse = shift + logf(s);
  const float l = lraw - lse;  // normalised log-weight of my particle
  // ---- estimate ----------------------------------------------------------------------------------------------------------
  float x[MMF_MAX_SD];
#pragma unroll
  for (int d = 0; d < MMF_MAX_SD; ++d) x[d] = (mine && d < sd) ? R.states[lane * sd + d] : 0.0f;
  if (mine && !resample) R.logw_out[lane] = l;
  if (R.estimation == MMF_ESTIMATE_WEIGHTED_AVERAGE) {
    const float wgt = mine ? expf(l) : 0.0f;
#pragma unroll
    for (int d = 0; d < MMF_MAX_SD; ++d) {
      if (d < sd) {
        float acc = 0.0f;
        if (mine) acc = fmaf(wgt, x[d], acc);
        const float v = warp_sum(acc);
        if (lane == 0) R.est_out[d] = v;
      }
    }
  } else {
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    if (mine && l > best) {
      best = l;
      best_i = lane;
    }
    const float gbest = warp_max(best);
    int cand = (best == gbest) ? best_i : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    const int win = cand == 0x7fffffff ? 0 : cand;
    if (lane < sd) R.est_out[lane] = R.states[win * sd + lane];
  }
  if (!resample) return;
  // ---- pinned softmax numerators and the strict (sequential) CDF ---------------------------------------------------------------
  const float lmax = warp_max(mine ? l : -INFINITY);
  const float e = mine ? exp_pinned(l - lmax) : 0.0f;
  float run = 0.0f, c = 0.0f;
  for (int k = 0; k < M; ++k) {
    run = __fadd_rn(run, __shfl_sync(0xffffffffu, e, k));
    if (lane == k) c = run;
  }
  const float total = run;  // c_{M-1}
  // ---- inverse CDF of my draw + gather ----------------------------------------------------------------------------------
  const bool systematic = R.mode == MMF_RESAMPLE_SYSTEMATIC_STRICT;
  double u = 0.5;
  if (mine) u = systematic ? (R.uniforms[0] + (double)lane) / (double)R.M_out : R.uniforms[lane];
  const float cstar = cdf_threshold(total, u);
  int idx = 0;
  for (int k = 0; k < M; ++k) idx += __shfl_sync(0xffffffffu, c, k) < cstar ? 1 : 0;
  idx = idx < M - 1 ? idx : M - 1;
  float sv[MMF_MAX_SD];
#pragma unroll
  for (int d = 0; d < MMF_MAX_SD; ++d) sv[d] = (mine && d < sd) ? R.states[idx * sd + d] : 0.0f;
  if (mine) {
#pragma unroll
    for (int d = 0; d < MMF_MAX_SD; ++d)
      if (d < sd) R.states_out[lane * sd + d] = sv[d];
    R.logw_out[lane] = -logf((float)M);
  }
}

template <int CH, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_pf_loop_small(const __grid_constant__ LoopParams P) {
  constexpr int LS_THREADS = NW * 32;
  constexpr int MP = 32 * CH;       // particle slots of the CTA
  constexpr int OPW = U * CH / NW;  // output features per warp (NW / CH warps share a particle chunk)
  static_assert(NW % CH == 0 && OPW >= 4 && OPW % 4 == 0 && OPW <= U, "warp tiling");
  constexpr int OQ = OPW / 4;       // ... in float4 units
  extern __shared__ __align__(16) float sm[];
  __shared__ const float* st_src[LS_MAX_STAGES];
  __shared__ int st_nf[LS_MAX_STAGES];
  __shared__ unsigned char st_kind[LS_MAX_STAGES], st_chain[LS_MAX_STAGES];
  __shared__ int n_stages_s;
  float* wbuf = sm;                                 // [LS_NBUF][LS_WMAX]
  float* act = wbuf + LS_NBUF * LS_WMAX;            // [2][64 / 4][MP][4]
  float* xmv = act + 2 * U * MP;                    // [MP][4] moved particle (input of the heads)
  float* nrs = xmv + MP * 4;                        // nr_trajectory scratch slice

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int chunk = warp % CH, ob = warp / CH;      // particle chunk, output block
  const int pslot = chunk * 32 + lane;              // particle slot of this lane
  const int n = blockIdx.x;                         // trajectory
  const int M = P.M, sd = P.sd, T = P.T;
  const bool live = pslot < M;
  const int p = live ? pslot : M - 1;               // clamped particle (loads only)
  const size_t base = (size_t)n * M;

  if (tid == 0) {  // the step's stages: every enabled chain = input layer, its 64x64 layers, output layer
    int g = 0;
    for (int c = 0; c <= P.K; ++c) {
      if (c > 0 && !((P.enabled >> (c - 1)) & 1u)) continue;
      const ChainDev ch = P.chains[c];
      const float* w = ch.w;
      st_src[g] = w; st_nf[g] = ch.in_dim * U + U; st_kind[g] = LS_IN; st_chain[g] = (unsigned char)c; ++g;
      w += ch.in_dim * U + U;
      const int L = 2 * ch.n_pre + 1 + 2 * ch.n_post, mid_at = 2 * ch.n_pre;
      for (int s = 0; s < L; ++s) {
        const bool is_mid = s == mid_at;
        const int rel = s < mid_at ? s : s - mid_at - 1;
        st_src[g] = w; st_nf[g] = is_mid ? U * U : U * U + U;
        st_kind[g] = is_mid ? LS_MID : ((rel & 1) == 0 ? LS_RES_A : LS_RES_B);
        st_chain[g] = (unsigned char)c; ++g;
        w += is_mid ? U * U : U * U + U;
      }
      st_src[g] = w; st_nf[g] = ch.out_dim * U + ch.out_dim; st_kind[g] = LS_OUT; st_chain[g] = (unsigned char)c; ++g;
    }
    n_stages_s = g;
  }
  __syncthreads();
  const int G = n_stages_s;
  int last_chain = 0;
  for (int c = 1; c <= P.K; ++c)
    if ((P.enabled >> (c - 1)) & 1u) last_chain = c;

  auto prefetch = [&](int stage, int buf) {
    const float* src = st_src[stage];
    const int nf = st_nf[stage], n16 = nf >> 2;
    float* dst = wbuf + buf * LS_WMAX;
    for (int i = tid; i < n16; i += LS_THREADS) cp_async16(dst + 4 * i, src + 4 * i);
    for (int i = (n16 << 2) + tid; i < nf; i += LS_THREADS) cp_async4(dst + i, src + i);
    cp_async_commit();
  };

  ResampleParams R;
  R.N = P.N; R.M = M; R.sd = sd; R.M_out = M;
  R.estimation = P.estimation; R.mode = P.mode; R.alpha = 1.0f;
  R.logits_in = nullptr; R.logw_norm_out = nullptr; R.logits_out = nullptr; R.idx_out = nullptr;
  R.N = 1;  // nr_trajectory / nr_small are called with n = 0 on pointers that already address this trajectory
  R.logw_unnorm = P.logw_ws + base; R.logw_out = P.logw + base;
  const bool resample = P.mode != MMF_RESAMPLE_NONE;
  const bool systematic = P.mode == MMF_RESAMPLE_SYSTEMATIC_STRICT || P.mode == MMF_RESAMPLE_SYSTEMATIC_FAST;

  float* cur = P.states;       // particle set entering the step
  float* moved = P.states_ws;  // after the dynamics
  // weight ring: stage number s (counted over the whole sequence) lives in buffer s % LS_NBUF; one cp.async group per
  // stage (possibly empty), issued LS_NBUF - 1 stages ahead
  const long long total_stages = (long long)T * G;
  long long sq = 0;            // sequence number of the current stage
  int cur_buf = 0, ahead_stage = 0, ahead_buf = 0;  // ring position of the current stage / of the next stage to fetch
  for (int a = 0; a < LS_NBUF - 1; ++a) {
    if (a < total_stages) prefetch(ahead_stage, ahead_buf); else cp_async_commit();
    ahead_stage = ahead_stage + 1 == G ? 0 : ahead_stage + 1;
    ahead_buf = ahead_buf + 1 == LS_NBUF ? 0 : ahead_buf + 1;
  }

  for (int t = 0; t < T; ++t) {
    float x[MMF_MAX_SD];       // this lane's particle entering the step
#pragma unroll
    for (int i = 0; i < MMF_MAX_SD; ++i) x[i] = 0.0f;
    float lse_m = -INFINITY, lse_s = 0.0f;
    int ab = 0;                // activation buffer holding the current layer's input
    // operands from global memory are fetched at the chain's input layer, layers before their use
    float4 rb[OQ];             // this warp's slice of the chain's per-trajectory row (bias of the mid layer)
    float e[MMF_MAX_SD] = {0.f, 0.f, 0.f, 0.f}, mw = 0.0f, lw_in = 0.0f;

    for (int gi = 0; gi < G; ++gi) {
      asm volatile("cp.async.wait_group %0;" ::"n"(LS_NBUF - 2) : "memory");  // this stage's group has completed
      __syncthreads();  // its weights are visible to all; the previous stage's activations / particle set too
      // fetch the stage LS_NBUF - 1 ahead: its buffer was last read by the stage before this barrier
      if (sq + LS_NBUF - 1 < total_stages) prefetch(ahead_stage, ahead_buf); else cp_async_commit();
      ahead_stage = ahead_stage + 1 == G ? 0 : ahead_stage + 1;
      ahead_buf = ahead_buf + 1 == LS_NBUF ? 0 : ahead_buf + 1;
      const float* wb = wbuf + cur_buf * LS_WMAX;
      cur_buf = cur_buf + 1 == LS_NBUF ? 0 : cur_buf + 1;
      ++sq;
      const int kind = st_kind[gi], c = st_chain[gi];
      const ChainDev ch = P.chains[c];

      if (kind == LS_IN) {
        // ---- input layer: relu(in_W x + in_b), this warp's OPW features of its particle --------------------------------
        float xi[MMF_MAX_SD];
        if (c == 0) {
#pragma unroll
          for (int i = 0; i < MMF_MAX_SD; ++i) {
            x[i] = i < sd ? cur[(base + p) * sd + i] : 0.0f;
            xi[i] = x[i];
          }
        } else {
          const float4 v = *reinterpret_cast<const float4*>(xmv + pslot * 4);
          xi[0] = v.x; xi[1] = v.y; xi[2] = v.z; xi[3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < OQ; ++q)
          rb[q] = __ldg(reinterpret_cast<const float4*>(P.rowbias + ((size_t)c * T * P.N + (size_t)t * P.N + n) * U +
                                                         ob * OPW + 4 * q));
        if (ob == 0) {
          if (c == 0) {
#pragma unroll
            for (int i = 0; i < MMF_MAX_SD; ++i)
              if (i < sd) e[i] = __ldg(P.eps + ((size_t)t * P.N * M + base + p) * sd + i);
          } else {
            mw = P.modw != nullptr ? __ldg(P.modw + ((size_t)t * P.N + n) * P.K + (c - 1)) : 0.0f;
            if (c == last_chain) lw_in = P.logw[base + p];
          }
        }
        const float* inb = wb + ch.in_dim * U;
#pragma unroll
        for (int q = 0; q < OQ; ++q) {
          const int j = ob * OPW + 4 * q;
          float4 a = *reinterpret_cast<const float4*>(inb + j);
#pragma unroll
          for (int i = 0; i < MMF_MAX_SD; ++i) {
            if (i < ch.in_dim) {
              const float4 w = *reinterpret_cast<const float4*>(wb + i * U + j);
              a.x = fmaf(w.x, xi[i], a.x);
              a.y = fmaf(w.y, xi[i], a.y);
              a.z = fmaf(w.z, xi[i], a.z);
              a.w = fmaf(w.w, xi[i], a.w);
            }
          }
          a.x = fmaxf(a.x, 0.0f); a.y = fmaxf(a.y, 0.0f); a.z = fmaxf(a.z, 0.0f); a.w = fmaxf(a.w, 0.0f);
          *reinterpret_cast<float4*>(act + ((size_t)((j >> 2) * MP + pslot)) * 4) = a;  // buffer 0
        }
        ab = 0;
      } else if (kind != LS_OUT) {
        // ---- 64 -> 64 layer ------------------------------------------------------------------------------------------
        const float* ain = act + ab * (U * MP);
        float* aout = act + (ab ^ 1) * (U * MP);
        float h[U];
#pragma unroll
        for (int k4 = 0; k4 < U / 4; ++k4) {
          const float4 v = *reinterpret_cast<const float4*>(ain + ((size_t)(k4 * MP + pslot)) * 4);
          h[4 * k4] = v.x; h[4 * k4 + 1] = v.y; h[4 * k4 + 2] = v.z; h[4 * k4 + 3] = v.w;
        }
        float2 acc[2 * OQ];
#pragma unroll
        for (int q = 0; q < OQ; ++q) {
          const int j = ob * OPW + 4 * q;
          float4 b;
          if (kind == LS_MID) {
            b = rb[q];
          } else {
            b = *reinterpret_cast<const float4*>(wb + U * U + j);
            if (kind == LS_RES_B) {  // + the block's input, parked in the other buffer at this lane's own positions
              const float4 r = *reinterpret_cast<const float4*>(aout + ((size_t)((j >> 2) * MP + pslot)) * 4);
              b.x += r.x; b.y += r.y; b.z += r.z; b.w += r.w;
            }
          }
          acc[2 * q] = make_float2(b.x, b.y);
          acc[2 * q + 1] = make_float2(b.z, b.w);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const float2 hk = make_float2(h[k], h[k]);
#pragma unroll
          for (int q = 0; q < OQ; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(wb + k * U + ob * OPW + 4 * q);  // warp broadcast
            acc[2 * q] = __ffma2_rn(make_float2(w.x, w.y), hk, acc[2 * q]);
            acc[2 * q + 1] = __ffma2_rn(make_float2(w.z, w.w), hk, acc[2 * q + 1]);
          }
        }
        const bool relu = kind == LS_MID ? ch.mid_relu != 0 : true;
#pragma unroll
        for (int q = 0; q < OQ; ++q) {
          float4 v = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
          if (relu) {
            v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
          }
          *reinterpret_cast<float4*>(aout + ((size_t)(((ob * OPW) >> 2) + q) * MP + pslot) * 4) = v;
        }
        ab ^= 1;
      } else if (ob == 0) {
        // ---- output layer (row-major out_W[out_dim][64] | out_b): one warp per particle chunk -----------------------------
        const float* ain = act + ab * (U * MP);
        float h[U];
#pragma unroll
        for (int k4 = 0; k4 < U / 4; ++k4) {
          const float4 v = *reinterpret_cast<const float4*>(ain + ((size_t)(k4 * MP + pslot)) * 4);
          h[4 * k4] = v.x; h[4 * k4 + 1] = v.y; h[4 * k4 + 2] = v.z; h[4 * k4 + 3] = v.w;
        }
        float y[MMF_MAX_SD + 1];
#pragma unroll
        for (int o = 0; o < MMF_MAX_SD + 1; ++o) {
          y[o] = 0.0f;
          if (o < ch.out_dim) {
            float a = wb[ch.out_dim * U + o];
#pragma unroll
            for (int k = 0; k < U; ++k) a = fmaf(wb[o * U + k], h[k], a);
            y[o] = a;
          }
        }
        if (c == 0) {
          float gsel = 0.0f;
#pragma unroll
          for (int o = 0; o < MMF_MAX_SD + 1; ++o)
            if (o == sd) gsel = y[o];
          const float gate = 1.0f / (1.0f + expf(-gsel));
          float xn[MMF_MAX_SD];
#pragma unroll
          for (int i = 0; i < MMF_MAX_SD; ++i) {
            xn[i] = 0.0f;
            if (i < sd) {
              const float pred = x[i] + y[i] * gate;
              float noise = 0.0f;
#pragma unroll
              for (int j = 0; j < MMF_MAX_SD; ++j)
                if (j <= i && j < sd) noise = fmaf(P.q[i * sd + j], e[j], noise);
              xn[i] = pred + noise;
              if (live) moved[(base + p) * sd + i] = xn[i];
            }
          }
          *reinterpret_cast<float4*>(xmv + pslot * 4) = make_float4(xn[0], xn[1], xn[2], xn[3]);
        } else {
          const float v = y[0] + mw;
          if (v > lse_m) {
            lse_s = lse_s * expf(lse_m - v) + 1.0f;
            lse_m = v;
          } else if (v > -INFINITY) {
            lse_s += expf(v - lse_m);
          }
          if (c == last_chain && live) {
            const float fused = (lse_m == -INFINITY) ? -INFINITY : lse_m + logf(lse_s);
            P.logw_ws[base + p] = lw_in + fused;
          }
        }
      }
    }

    // ---- normalise, estimate, resample + gather: warp 0, the arithmetic of k_normalize_resample ------------------------
    __syncthreads();  // moved particles and un-normalised log-weights of all chunks are in (L1-coherent) global memory
    if (warp == 0) {
      R.states = moved + base * sd;
      R.uniforms = P.uniforms ? P.uniforms + (size_t)t * (systematic ? (size_t)P.N : (size_t)P.N * M) +
                                    (systematic ? (size_t)n : base) : nullptr;
      R.states_out = resample ? cur + base * sd : nullptr;
      R.est_out = P.est_out + ((size_t)t * P.N + n) * sd;
      if (nr_small_applies(R)) nr_small(R, lane); else nr_trajectory<32>(R, 0, nrs, nullptr, nullptr);
    }
    if (!resample) {  // the moved set IS the next step's input
      float* tmp = cur; cur = moved; moved = tmp;
    }
    // (the barrier at the top of the next stage orders warp 0's stores before the next step's loads)
  }
  __syncthreads();
  if (cur != P.states)
    for (int i = tid; i < M * sd; i += LS_THREADS) P.states[base * sd + i] = cur[base * sd + i];
}

// ---- tensor-core variant (MMF_PREC_BF16X3 / MMF_PREC_BF16): same kernel structure, the 64 -> 64 layers on mma.sync ---------
// The CUDA-core variant above is bound by shared-memory bandwidth: a warp-broadcast weight read feeds 32 lanes only, and
// ncu counts 2560 shared-memory wavefronts per layer against 1024 FFMA2 issue slots (profiles/r02_summary.md).  Here a
// layer is D[32 CH particles x 64] = A x W^T on the warp-level tensor-core path (mma.sync.m16n8k16, bf16 operands, fp32
// accumulate; tcgen05 needs a 128-row tile and a TMEM round trip per layer, which is all latency at 30 particles) with
// the split operands of the throughput kernel (a_hi w_hi + a_hi w_lo + a_lo w_hi):
//   * B = the chain's tcgen05 operand image AS IT IS (bf16 hi / lo tiles, K-major SWIZZLE_128B): ldmatrix on it yields
//     mma.sync B fragments directly and conflict-free, so no weight conversion happens in this kernel; a layer's two tiles
//     (16 KB) stream through the same cp.async ring, the fp32 tails (input layer, biases) and output tiles stay resident;
//   * A = the activations as bf16 hi / lo planes in shared memory ([particle][64 + 8 pad] bf16, ldmatrix conflict-free),
//     written by the previous layer's epilogue, plus an fp32 copy of the residual stream;
//   * warp = (chunk of 32 particles, quarter of the 64 features): 2 m-tiles x 2 n-tiles, 48 MMAs per layer.
constexpr int LM_AST = 36;     // 32-bit words per particle row of a bf16 plane (64 bf16 + 8 pad)
constexpr int LM_XST = 72;     // floats per particle row of the fp32 residual buffer
constexpr int LM_TAIL = 1024;  // floats reserved per chain for the fp32 tail of its image

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// v = (features c, c + 1 of one particle) -> bf16 hi / lo words: hi = rn(v), lo = rn(v - hi) (v - hi is exact)
__device__ __forceinline__ void split_pair(float2 v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.y), "f"(v.x));
  float rx, ry;
  asm("{\n\t"
      ".reg .b16 l, h, m1;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "mov.b16 m1, 0xBF80;\n\t"
      "fma.rn.f32.bf16 %0, l, m1, %3;\n\t"
      "fma.rn.f32.bf16 %1, h, m1, %4;\n\t"
      "}"
      : "=f"(rx), "=f"(ry)
      : "r"(hi), "f"(v.x), "f"(v.y));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(ry), "f"(rx));
}

// NT = n-tiles (8 output features each) per warp: 2 -> 4 warps per particle chunk, 1 -> 8 warps per chunk (two warps per
// scheduler hide each other's ldmatrix / HMMA latency; every warp then re-reads the chunk's A operand).
template <int CH, int NT>
__global__ void __launch_bounds__(256 * CH / NT, 1) k_pf_loop_small_mma(const __grid_constant__ LoopParams P) {
  constexpr int NQ = 8 / NT;  // warps per particle chunk
  constexpr int THREADS = 32 * NQ * CH, MP = 32 * CH;
  constexpr int RING_B = 2 * TILE_B;  // bytes of one layer: hi tile | lo tile
  extern __shared__ __align__(1024) uint8_t smb[];
  __shared__ const uint8_t* ring_src[LS_MAX_STAGES];
  __shared__ unsigned char st_kind[LS_MAX_STAGES], st_chain[LS_MAX_STAGES], st_layer[LS_MAX_STAGES];
  __shared__ int n_stages_s, n_dense_s;
  uint8_t* ring = smb;                                                       // [LS_NBUF][RING_B]
  uint8_t* outt = ring + LS_NBUF * RING_B;                                   // [1 + K][2 * OUT_TILE_B]
  float* tails = reinterpret_cast<float*>(outt + (1 + MMF_MAX_HEADS) * 2 * OUT_TILE_B);  // [1 + K][LM_TAIL]
  uint32_t* planes = reinterpret_cast<uint32_t*>(tails + (1 + MMF_MAX_HEADS) * LM_TAIL);  // [2][hi | lo][MP][LM_AST]
  float* xres = reinterpret_cast<float*>(planes + 2 * 2 * MP * LM_AST);      // [MP][LM_XST] fp32 residual stream
  float* xmv = xres + MP * LM_XST;                                           // [MP][4] moved particle (input of the heads)
  float* p_a = xmv + MP * 4;                                                 // [M][sd] packed particle set (entering the step)
  float* p_b = p_a + MP * 4;                                                 // [M][sd] packed moved set
  float* p_lw = p_b + MP * 4;                                                // [M] normalised log-weights
  float* p_lwu = p_lw + MP;                                                  // [M] un-normalised log-weights of the step
  float* nrs = p_lwu + MP;                                                   // nr_trajectory scratch slice
  __shared__ __align__(8) uint64_t full[LS_NBUF];                            // ring: bytes of a layer have landed

  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int chunk = warp / NQ, nq = warp % NQ;
  const int n = blockIdx.x;
  const int M = P.M, sd = P.sd, T = P.T;
  const size_t base = (size_t)n * M;
  const bool single_pass = P.single_pass != 0;

  if (tid == 0) {
    int gcount = 0, d = 0;
    for (int c = 0; c <= P.K; ++c) {
      if (c > 0 && !((P.enabled >> (c - 1)) & 1u)) continue;
      const ChainDev ch = P.chains[c];
      const int L = chain_layers(ch), mid_at = 2 * ch.n_pre;
      st_kind[gcount] = LS_IN; st_chain[gcount] = (unsigned char)c; st_layer[gcount] = 0; ++gcount;
      for (int l = 0; l < L; ++l) {
        const int rel = l < mid_at ? l : l - mid_at - 1;
        st_kind[gcount] = l == mid_at ? LS_MID : ((rel & 1) == 0 ? LS_RES_A : LS_RES_B);
        st_chain[gcount] = (unsigned char)c; st_layer[gcount] = (unsigned char)l; ++gcount;
        ring_src[d++] = P.images[c] + (size_t)l * RING_B;
      }
      st_kind[gcount] = LS_OUT; st_chain[gcount] = (unsigned char)c; st_layer[gcount] = (unsigned char)L; ++gcount;
    }
    n_stages_s = gcount;
    n_dense_s = d;
  }
  if (tid == 0) {
    for (int i = 0; i < LS_NBUF; ++i) mbar_init(full + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the trajectory's particle set lives in shared memory for the whole sequence
  for (int i = tid; i < M * sd; i += THREADS) p_a[i] = P.states[base * sd + i];
  for (int i = tid; i < M; i += THREADS) p_lw[i] = P.logw[base + i];
  // resident parts of the images: fp32 tails [in_Wt | in_b | bias[L][64] | out_b[16]] and the output-layer tiles
  for (int c = 0; c <= P.K; ++c) {
    const ChainDev ch = P.chains[c];
    const int L = chain_layers(ch);
    const float* tsrc = reinterpret_cast<const float*>(P.images[c] + image_tiles_bytes(ch));
    const int nt = ch.in_dim * U + U + L * U + OUT_PAD;
    for (int i = tid; i < nt; i += THREADS) tails[c * LM_TAIL + i] = tsrc[i];
    const uint4* osrc = reinterpret_cast<const uint4*>(P.images[c] + (size_t)L * RING_B);
    uint4* odst = reinterpret_cast<uint4*>(outt + c * 2 * OUT_TILE_B);
    for (int i = tid; i < 2 * OUT_TILE_B / 16; i += THREADS) odst[i] = osrc[i];
  }
  __syncthreads();
  const int G = n_stages_s, GD = n_dense_s;
  int last_chain = 0;
  for (int c = 1; c <= P.K; ++c)
    if ((P.enabled >> (c - 1)) & 1u) last_chain = c;

  // a layer's two tiles arrive by ONE bulk async copy (TMA) issued by thread 0; the per-layer __syncthreads is what frees a
  // ring buffer (every warp has read it), the mbarrier is what publishes it
  auto prefetch = [&](int dense_stage, int buf) {
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(full + buf, RING_B);
      bulk_g2s(ring + buf * RING_B, ring_src[dense_stage], RING_B, full + buf);
    }
  };
  // bf16 planes of activation buffer b: hi at +0, lo at + MP * LM_AST words
  auto plane = [&](int b) -> uint32_t* { return planes + (size_t)b * 2 * MP * LM_AST; };
  // features (c, c + 1) of particle row r: fp32 copy (residual stream) and bf16 hi / lo words for the next layer's A operand
  auto store_act = [&](int b, int r, int c, float2 v, bool write_x) {
    if (write_x) *reinterpret_cast<float2*>(xres + r * LM_XST + c) = v;
    uint32_t hi, lo;
    split_pair(v, hi, lo);
    uint32_t* pl = plane(b);
    pl[r * LM_AST + (c >> 1)] = hi;
    pl[MP * LM_AST + r * LM_AST + (c >> 1)] = lo;
  };

  ResampleParams R;
  R.N = P.N; R.M = M; R.sd = sd; R.M_out = M;
  R.estimation = P.estimation; R.mode = P.mode; R.alpha = 1.0f;
  R.logits_in = nullptr; R.logw_norm_out = nullptr; R.logits_out = nullptr; R.idx_out = nullptr;
  R.N = 1;  // nr_trajectory is called with n = 0 on pointers that already address this trajectory
  R.logw_unnorm = p_lwu; R.logw_out = p_lw;
  const bool resample = P.mode != MMF_RESAMPLE_NONE;
  const bool systematic = P.mode == MMF_RESAMPLE_SYSTEMATIC_STRICT || P.mode == MMF_RESAMPLE_SYSTEMATIC_FAST;

  float* cur = p_a;
  float* moved = p_b;
  const long long total_dense = (long long)T * GD;
  long long dq = 0;  // dense stages done so far (over the whole sequence)
  int cur_buf = 0, ahead_stage = 0, ahead_buf = 0;
  uint32_t full_par = 0;  // bit b = parity of the next completion of ring buffer b
  for (int a = 0; a < LS_NBUF - 1; ++a) {
    if (a < total_dense) prefetch(ahead_stage, ahead_buf);
    ahead_stage = ahead_stage + 1 == GD ? 0 : ahead_stage + 1;
    ahead_buf = ahead_buf + 1 == LS_NBUF ? 0 : ahead_buf + 1;
  }
  // the particles whose output arithmetic this lane performs (lanes t == 0 of the warps nq == 0): rows g, g + 8 of both m-tiles
  const bool out_lane = nq == 0 && t == 0;
  int oslot[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) oslot[q] = chunk * 32 + (q >> 1) * 16 + g + (q & 1) * 8;

  for (int step = 0; step < T; ++step) {
    // ---- operands of the output arithmetic -> registers; this step's uniforms on their way to L1 ---------------------------
    if (warp == 0 && P.uniforms != nullptr) {
      const double* un = P.uniforms + (size_t)step * (systematic ? (size_t)P.N : (size_t)P.N * M) + (systematic ? (size_t)n : base);
      const int count = systematic ? 1 : M;
      for (int i = lane * 16; i < count; i += 32 * 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(un + i));
    }
    float eps[4][MMF_MAX_SD], lw_in[4], lse_m[4], lse_s[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      lse_m[q] = -INFINITY;
      lse_s[q] = 0.0f;
      lw_in[q] = 0.0f;
#pragma unroll
      for (int i = 0; i < MMF_MAX_SD; ++i) eps[q][i] = 0.0f;
      if (out_lane) {
        const int p = oslot[q] < M ? oslot[q] : M - 1;
        lw_in[q] = p_lw[p];
#pragma unroll
        for (int i = 0; i < MMF_MAX_SD; ++i)
          if (i < sd) eps[q][i] = __ldg(P.eps + ((size_t)step * P.N * M + base + p) * sd + i);
      }
    }
    int ab = 0;
    float2 rb[NT];  // this thread's columns of the chain's per-trajectory row
#pragma unroll
    for (int j = 0; j < NT; ++j) rb[j] = make_float2(0.f, 0.f);
    float mw = 0.0f;

    for (int gi = 0; gi < G; ++gi) {
      const int kind = st_kind[gi], c = st_chain[gi], layer = st_layer[gi];
      const ChainDev ch = P.chains[c];
      const int L = chain_layers(ch);
      const float* tl = tails + c * LM_TAIL;
      const float* in_b = tl + ch.in_dim * U;
      const float* biases = in_b + U;
      const float* out_b = biases + L * U;
      const bool dense = kind != LS_IN && kind != LS_OUT;
      __syncthreads();  // the previous stage's activations are visible; its ring buffer is free
      const uint8_t* wt = ring + cur_buf * RING_B;
      if (dense) {
        if (dq + LS_NBUF - 1 < total_dense) prefetch(ahead_stage, ahead_buf);
        mbar_wait(full + cur_buf, (full_par >> cur_buf) & 1u);
        full_par ^= 1u << cur_buf;
        ahead_stage = ahead_stage + 1 == GD ? 0 : ahead_stage + 1;
        ahead_buf = ahead_buf + 1 == LS_NBUF ? 0 : ahead_buf + 1;
        cur_buf = cur_buf + 1 == LS_NBUF ? 0 : cur_buf + 1;
        ++dq;
      }

      if (kind == LS_IN) {
        // ---- input layer on the CUDA cores: relu(in_W x + in_b) for this thread's 4 rows x 4 columns ---------------------------
#pragma unroll
        for (int j = 0; j < NT; ++j)
          rb[j] = __ldg(reinterpret_cast<const float2*>(P.rowbias + ((size_t)c * T * P.N + (size_t)step * P.N + n) * U +
                                                        8 * (NT * nq + j) + 2 * t));
        if (c > 0) mw = P.modw != nullptr ? __ldg(P.modw + ((size_t)step * P.N + n) * P.K + (c - 1)) : 0.0f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int r = chunk * 32 + mt * 16 + g + hf * 8;
            float xi[4] = {0.f, 0.f, 0.f, 0.f};
            if (c == 0) {
              const int pr = r < M ? r : M - 1;
#pragma unroll
              for (int i = 0; i < MMF_MAX_SD; ++i)
                if (i < sd) xi[i] = cur[pr * sd + i];
            } else {
              const float4 xv = *reinterpret_cast<const float4*>(xmv + r * 4);
              xi[0] = xv.x; xi[1] = xv.y; xi[2] = xv.z; xi[3] = xv.w;
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              const int col = 8 * (NT * nq + j) + 2 * t;
              float2 v = *reinterpret_cast<const float2*>(in_b + col);
#pragma unroll
              for (int i = 0; i < MMF_MAX_SD; ++i) {
                if (i < ch.in_dim) {
                  const float2 w = *reinterpret_cast<const float2*>(tl + i * U + col);
                  v.x = fmaf(w.x, xi[i], v.x);
                  v.y = fmaf(w.y, xi[i], v.y);
                }
              }
              v.x = fmaxf(v.x, 0.0f);
              v.y = fmaxf(v.y, 0.0f);
              store_act(0, r, col, v, true);
            }
          }
        }
        ab = 0;
      } else if (dense) {
        // ---- 64 -> 64 layer: 2 m-tiles x 2 n-tiles per warp, K = 64 in four k16 steps, split operands ----------------------------
        const uint32_t a_hi = smem_u32(plane(ab)), a_lo = a_hi + MP * LM_AST * 4;
        const uint32_t w_hi = smem_u32(wt), w_lo = w_hi + TILE_B;
        float acc[2][NT][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.0f;
        const int lm = lane >> 3, lr = lane & 7;  // ldmatrix: this lane addresses row lr of matrix lm
        uint32_t bh[4], bl[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          uint32_t ah[2][4], al[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t off = (uint32_t)(((chunk * 32 + mt * 16 + (lm & 1) * 8 + lr) * LM_AST + 8 * s + (lm >> 1) * 4) * 4);
            ldmatrix_x4(ah[mt], a_hi + off);
            if (!single_pass) ldmatrix_x4(al[mt], a_lo + off);
          }
          if (NT == 2) {  // matrices: (n-tile 2 nq, k chunk 2 s), (2 nq, 2 s + 1), (2 nq + 1, 2 s), (2 nq + 1, 2 s + 1)
            const uint32_t off = (uint32_t)((2 * nq + (lm >> 1)) * 1024 + lr * 128 + (((2 * s + (lm & 1)) ^ lr) << 4));
            ldmatrix_x4(bh, w_hi + off);
            if (!single_pass) ldmatrix_x4(bl, w_lo + off);
          } else if ((s & 1) == 0) {  // one n-tile: the k chunks 2 s .. 2 s + 3 of two steps in one load
            const uint32_t off = (uint32_t)(nq * 1024 + lr * 128 + (((2 * s + lm) ^ lr) << 4));
            ldmatrix_x4(bh, w_hi + off);
            if (!single_pass) ldmatrix_x4(bl, w_lo + off);
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              const int q = NT == 2 ? 2 * j : 2 * (s & 1);  // registers of this (n-tile, step) in bh / bl
              mma_bf16(acc[mt][j], ah[mt], bh[q], bh[q + 1]);
              if (!single_pass) {
                mma_bf16(acc[mt][j], ah[mt], bl[q], bl[q + 1]);
                mma_bf16(acc[mt][j], al[mt], bh[q], bh[q + 1]);
              }
            }
          }
        }
        const bool relu = kind == LS_MID ? ch.mid_relu != 0 : true;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const int col = 8 * (NT * nq + j) + 2 * t;
            const float2 b = kind == LS_MID ? rb[j] : *reinterpret_cast<const float2*>(biases + layer * U + col);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int r = chunk * 32 + mt * 16 + g + hf * 8;
              float2 v = make_float2(acc[mt][j][2 * hf] + b.x, acc[mt][j][2 * hf + 1] + b.y);
              if (kind == LS_RES_B) {
                const float2 x = *reinterpret_cast<const float2*>(xres + r * LM_XST + col);
                v.x += x.x;
                v.y += x.y;
              }
              if (relu) {
                v.x = fmaxf(v.x, 0.0f);
                v.y = fmaxf(v.y, 0.0f);
              }
              store_act(ab ^ 1, r, col, v, kind != LS_RES_A);
            }
          }
        }
        ab ^= 1;
      } else if (nq == 0) {
        // ---- output layer (N padded to 16 in the image; n-tile 0 holds every output), then the per-particle arithmetic -------------
        const uint32_t a_hi = smem_u32(plane(ab)), a_lo = a_hi + MP * LM_AST * 4;
        const uint32_t w_hi = smem_u32(outt + c * 2 * OUT_TILE_B), w_lo = w_hi + OUT_TILE_B;
        float acc[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][e] = 0.0f;
        const int lm = lane >> 3, lr = lane & 7;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          uint32_t bh[2], bl[2];
          const uint32_t boff = (uint32_t)(lr * 128 + (((2 * s + (lm & 1)) ^ lr) << 4));
          ldmatrix_x2(bh, w_hi + boff);
          if (!single_pass) ldmatrix_x2(bl, w_lo + boff);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            uint32_t ah[4], al[4];
            const uint32_t off = (uint32_t)(((chunk * 32 + mt * 16 + (lm & 1) * 8 + lr) * LM_AST + 8 * s + (lm >> 1) * 4) * 4);
            ldmatrix_x4(ah, a_hi + off);
            mma_bf16(acc[mt], ah, bh[0], bh[1]);
            if (!single_pass) {
              ldmatrix_x4(al, a_lo + off);
              mma_bf16(acc[mt], ah, bl[0], bl[1]);
              mma_bf16(acc[mt], al, bh[0], bh[1]);
            }
          }
        }
        // lane (g, t) holds outputs 2 t, 2 t + 1 of rows g (acc[.][0..1]) and g + 8 (acc[.][2..3]): bring outputs 2, 3 to t == 0
        float y[4][MMF_MAX_SD + 1];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const float o2 = __shfl_down_sync(0xffffffffu, acc[mt][2 * hf], 1);
            const float o3 = __shfl_down_sync(0xffffffffu, acc[mt][2 * hf + 1], 1);
            const float o4 = __shfl_down_sync(0xffffffffu, acc[mt][2 * hf], 2);
            const float o[MMF_MAX_SD + 1] = {acc[mt][2 * hf], acc[mt][2 * hf + 1], o2, o3, o4};
#pragma unroll
            for (int k = 0; k < MMF_MAX_SD + 1; ++k) y[mt * 2 + hf][k] = o[k] + out_b[k];
          }
        }
        if (t == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int slot = oslot[q];
            const bool live = slot < M;
            const int p = live ? slot : M - 1;
            if (c == 0) {
              float gsel = 0.0f;
#pragma unroll
              for (int k = 0; k < MMF_MAX_SD + 1; ++k)
                if (k == sd) gsel = y[q][k];
              const float gate = 1.0f / (1.0f + expf(-gsel));
              float x[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int i = 0; i < MMF_MAX_SD; ++i)
                if (i < sd) x[i] = cur[p * sd + i];
              float xn[MMF_MAX_SD];
#pragma unroll
              for (int i = 0; i < MMF_MAX_SD; ++i) {
                xn[i] = 0.0f;
                if (i < sd) {
                  const float pred = x[i] + y[q][i] * gate;
                  float noise = 0.0f;
#pragma unroll
                  for (int j = 0; j < MMF_MAX_SD; ++j)
                    if (j <= i && j < sd) noise = fmaf(P.q[i * sd + j], eps[q][j], noise);
                  xn[i] = pred + noise;
                  if (live) moved[p * sd + i] = xn[i];
                }
              }
              *reinterpret_cast<float4*>(xmv + slot * 4) = make_float4(xn[0], xn[1], xn[2], xn[3]);
            } else {
              const float v = y[q][0] + mw;
              if (v > lse_m[q]) {
                lse_s[q] = lse_s[q] * expf(lse_m[q] - v) + 1.0f;
                lse_m[q] = v;
              } else if (v > -INFINITY) {
                lse_s[q] += expf(v - lse_m[q]);
              }
              if (c == last_chain && live) {
                const float fused = (lse_m[q] == -INFINITY) ? -INFINITY : lse_m[q] + logf(lse_s[q]);
                p_lwu[p] = lw_in[q] + fused;
              }
            }
          }
        }
      }
    }

    // ---- normalise, estimate, resample + gather: warp 0, the arithmetic of k_normalize_resample --------------------------------
    __syncthreads();
    if (warp == 0) {
      R.states = moved;
      R.uniforms = P.uniforms ? P.uniforms + (size_t)step * (systematic ? (size_t)P.N : (size_t)P.N * M) +
                                    (systematic ? (size_t)n : base) : nullptr;
      R.states_out = resample ? cur : nullptr;
      R.est_out = P.est_out + ((size_t)step * P.N + n) * sd;
      if (nr_small_applies(R)) nr_small(R, lane); else nr_trajectory<32>(R, 0, nrs, nullptr, nullptr);
    }
    if (!resample) {
      float* tmp = cur; cur = moved; moved = tmp;
    }
    __syncthreads();  // the next step reads the resampled set and the new log-weights before its first barrier
  }
  for (int i = tid; i < M * sd; i += THREADS) P.states[base * sd + i] = cur[i];
  for (int i = tid; i < M; i += THREADS) P.logw[base + i] = p_lw[i];
}


static bool loop_small_forced(int* forced) {
  static const int mode = [] {
    const char* env = getenv("MMF_PF_LOOP_SMALL");  // 0: never, 1: whenever the shape fits; unset: small problems only
    return env ? atoi(env) : -1;
  }();
  *forced = mode;
  return mode >= 0;
}

// Does mmf_pf_forward_loop run this (N, M) through the one-launch kernel?  It CAN for M <= 128 particles (a CTA per
// trajectory); it DOES by default for M <= 64 and N <= one CTA per SM, where it beats the per-step kernels (measured on
// B200, us per filter step, one-launch vs per-step: 32 x 30 particles 29 vs 51; 148 x 100: 58 vs 55; 296 x 30: 58 vs 55).
bool pf_loop_small_applies(int N, int M) {
  int forced = -1;
  if (M < 1 || M > 128 || N < 1) return false;
  if (loop_small_forced(&forced)) return forced != 0;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return M <= 64 && N <= sms;
}

template <int CH, int NW>
static int launch_ls(const LoopParams& P, size_t smem, cudaStream_t stream) {
  static thread_local int configured_dev = -1;
  static thread_local size_t window = 0;
  int dev = 0;
  MMF_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    int rc = opt_in_shared_memory(k_pf_loop_small<CH, NW>, &window);
    if (rc) return rc;
    configured_dev = dev;
  }
  MMF_REQUIRE(smem <= window, "pf_loop_small needs %zu B of shared memory (window %zu B)", smem, window);
  k_pf_loop_small<CH, NW><<<P.N, NW * 32, smem, stream>>>(P);
  MMF_LAUNCH_CHECK("k_pf_loop_small");
  return MMF_OK;
}

static size_t mma_smem_bytes(int ch, int M) {
  return (size_t)LS_NBUF * 2 * TILE_B + (size_t)(1 + MMF_MAX_HEADS) * (2 * OUT_TILE_B + LM_TAIL * sizeof(float)) +
         (size_t)32 * ch * ((size_t)2 * 2 * LM_AST * 4 + LM_XST * 4 + 3 * 4 * 4 + 2 * 4) + trajectory_scratch_floats(M, false) * sizeof(float);
}

// returns MMF_OK, an error, or -1 when the shared-memory window is too small for this shape (caller: CUDA-core variant)
template <int CH, int NT>
static int launch_ls_mma(const LoopParams& P, int M, cudaStream_t stream) {
  static thread_local int configured_dev = -1;
  static thread_local size_t window = 0;
  int dev = 0;
  MMF_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    int rc = opt_in_shared_memory(k_pf_loop_small_mma<CH, NT>, &window);
    if (rc) return rc;
    configured_dev = dev;
  }
  const size_t smem = mma_smem_bytes(CH, M);
  if (smem > window) return -1;
  k_pf_loop_small_mma<CH, NT><<<P.N, 256 * CH / NT, smem, stream>>>(P);
  MMF_LAUNCH_CHECK("k_pf_loop_small_mma");
  return MMF_OK;
}

int launch_pf_loop_small(const mmf_pf_model* model, int T, int N, int M, float* states, float* logw, const float* rowbias,
                         const float* modw, uint32_t enabled, int precision, const float* eps, int estimation, int mode,
                         const double* uniforms, float* states_ws, float* logw_ws, float* est_out, cudaStream_t stream) {
  LoopParams P;
  P.K = model->num_heads;
  P.chains[0] = to_dev(model->dynamics);
  P.images[0] = static_cast<const uint8_t*>(model->dynamics.w_mma);
  for (int k = 0; k < P.K; ++k) {
    P.chains[1 + k] = to_dev(model->heads[k]);
    P.images[1 + k] = static_cast<const uint8_t*>(model->heads[k].w_mma);
  }
  int stages = 0;
  bool have_images = true, tails_fit = true;
  for (int c = 0; c <= P.K; ++c) {
    MMF_REQUIRE(((uintptr_t)P.chains[c].w & 15) == 0, "pf_loop_small: chain %d weights are not 16-byte aligned", c);
    const int L = 2 * P.chains[c].n_pre + 1 + 2 * P.chains[c].n_post;
    stages += L + 2;
    have_images = have_images && P.images[c] != nullptr && ((uintptr_t)P.images[c] & 15) == 0;
    tails_fit = tails_fit && P.chains[c].in_dim * U + U + L * U + OUT_PAD <= LM_TAIL;
  }
  MMF_REQUIRE(stages <= LS_MAX_STAGES, "pf_loop_small: %d layers exceed the stage table (%d)", stages, LS_MAX_STAGES);
  P.enabled = enabled;
  P.sd = model->state_dim; P.N = N; P.M = M; P.T = T;
  P.states = states; P.logw = logw; P.states_ws = states_ws; P.logw_ws = logw_ws;
  P.eps = eps; P.rowbias = rowbias; P.modw = modw; P.uniforms = uniforms; P.est_out = est_out;
  P.estimation = estimation; P.mode = mode;
  P.single_pass = precision == MMF_PREC_BF16 ? 1 : 0;
  for (int i = 0; i < MMF_MAX_SD * MMF_MAX_SD; ++i) P.q[i] = model->q_tril[i];
  const int ch = M <= 32 ? 1 : (M <= 64 ? 2 : 4);
  if (precision != MMF_PREC_FP32 && have_images && tails_fit) {  // tensor-core variant (split bf16 operands)
    // M <= 32: eight warps of one n-tile each (MMF_LS_NT=2 keeps four warps of two n-tiles, for A/B timing)
    static const int nt1 = [] {
      const char* env = getenv("MMF_LS_NT");
      return env == nullptr || atoi(env) == 1;
    }();
    int rc = ch == 1 ? (nt1 ? launch_ls_mma<1, 1>(P, M, stream) : launch_ls_mma<1, 2>(P, M, stream))
             : ch == 2 ? launch_ls_mma<2, 2>(P, M, stream) : launch_ls_mma<4, 2>(P, M, stream);
    if (rc != -1) return rc;
  }
  const size_t floats = (size_t)LS_NBUF * LS_WMAX + 2 * (size_t)U * 32 * ch + (size_t)32 * ch * 4 + trajectory_scratch_floats(M, false);
  const size_t smem = floats * sizeof(float);
  // 8 warps per CTA: every warp re-reads its chunk's activations from shared memory, so fewer, wider warps save
  // shared-memory bandwidth (the bound of this kernel) and more warps hide more latency (measured at C1: 4 warps 50.2,
  // 8 warps 51.0, 16 warps 59.9 us per step)
  switch (ch) {
    case 1: return launch_ls<1, 8>(P, smem, stream);
    case 2: return launch_ls<2, 8>(P, smem, stream);
    default: return launch_ls<4, 8>(P, smem, stream);
  }
}

}  // namespace mmf
