// particle_chain_ffma.cu -- R3 + R4 + R5 + first half of R6 on CUDA cores, fp32 throughout
// (MMF_PREC_FP32: the parity build of the per-particle MLP chain).
//
// One thread owns one particle for the whole step: its 64-wide activation row lives in registers,
// weights of the current chain are staged once per CTA tile in shared memory and read as warp
// broadcasts (one LDS.128 feeds four FFMAs), the residual input of a resblock is parked in a
// thread-private shared-memory column.  Nothing but the particle state, its noise and its
// log-weight touches HBM (12*sd + 8 bytes per particle-step).
//
// Replaces, per particle: ref: crossmodal/push_models/dynamics.py:44-63 (state branch, shared
// layers, sigmoid gate), A.3 predict (`loc + scale_tril @ eps`), ref: crossmodal/push_models/pf.py:
// 91-109 per head, ref: crossmodal/base_models/crossmodal_pf.py:132-139 (fusion), A.3 `logw + ll`.
#include "kernels.cuh"

namespace mmf {

constexpr int TPB = 256;

struct PredictParams {
  ChainDev chains[1 + MMF_MAX_HEADS];  // [0] dynamics, [1+k] head k
  int K;
  uint32_t enabled;
  int sd;
  int N, M;
  long long total;
  const float* states_in;
  const float* eps;
  const float* rowbias;  // (1+K, rb_stride, 64)
  int rb_stride;
  const float* logw_in;
  const float* modw;  // (N, K) or null
  float* states_out;
  float* logw_out;
  float* ll_out;  // (K, N*M) or null
  float q[MMF_MAX_SD * MMF_MAX_SD];
};

// acc[j] += sum_k Wt[k][j] * h[k]; Wt in shared memory, every lane reads the same address.
__device__ __forceinline__ void dense64(const float* __restrict__ Wt, const float (&h)[U], float (&acc)[U]) {
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const float4* row = reinterpret_cast<const float4*>(Wt + k * U);
    const float hk = h[k];
#pragma unroll
    for (int j4 = 0; j4 < U / 4; ++j4) {
      const float4 w = row[j4];
      acc[4 * j4 + 0] = fmaf(w.x, hk, acc[4 * j4 + 0]);
      acc[4 * j4 + 1] = fmaf(w.y, hk, acc[4 * j4 + 1]);
      acc[4 * j4 + 2] = fmaf(w.z, hk, acc[4 * j4 + 2]);
      acc[4 * j4 + 3] = fmaf(w.w, hk, acc[4 * j4 + 3]);
    }
  }
}

__global__ void __launch_bounds__(TPB, 1) k_particle_chain_ffma(const __grid_constant__ PredictParams P) {
  extern __shared__ __align__(16) float smem[];
  float* wsm = smem;                       // staged chain weights (packed layout, verbatim)
  float* xs;                               // residual park: xs[k * TPB + tid]
  {
    int maxf = 0;
    for (int c = 0; c <= P.K; ++c) maxf = max(maxf, P.chains[c].floats());
    xs = smem + ((maxf + 3) & ~3);
  }
  const int tid = threadIdx.x;
  const int sd = P.sd;
  const long long tiles = (P.total + TPB - 1) / TPB;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long p_raw = tile * TPB + tid;
    const bool live = p_raw < P.total;
    const long long p = live ? p_raw : P.total - 1;
    const int n = (int)(p / P.M);

    float x[MMF_MAX_SD], xn[MMF_MAX_SD];
#pragma unroll
    for (int i = 0; i < MMF_MAX_SD; ++i) {
      x[i] = (i < sd) ? P.states_in[p * sd + i] : 0.0f;
      xn[i] = 0.0f;
    }
    float lse_m = -INFINITY, lse_s = 0.0f;

    for (int c = 0; c <= P.K; ++c) {
      if (c > 0 && !((P.enabled >> (c - 1)) & 1u)) continue;
      const ChainDev ch = P.chains[c];
      // ---- stage this chain's weights ---------------------------------------------------------
      __syncthreads();
      {
        const int nf = ch.floats();
        const float4* src = reinterpret_cast<const float4*>(ch.w);
        float4* dst = reinterpret_cast<float4*>(wsm);
        for (int i = tid; i < nf / 4; i += TPB) dst[i] = __ldg(src + i);
        for (int i = (nf & ~3) + tid; i < nf; i += TPB) wsm[i] = __ldg(ch.w + i);
      }
      __syncthreads();

      const float* w = wsm;
      float h[U], acc[U];
      // ---- input layer: h = relu(in_W x + in_b) -----------------------------------------------
      {
        const float* inb = w + ch.in_dim * U;
#pragma unroll
        for (int j = 0; j < U; ++j) acc[j] = inb[j];
#pragma unroll
        for (int i = 0; i < MMF_MAX_SD; ++i) {
          if (i < ch.in_dim) {
            const float xi = (c == 0) ? x[i] : xn[i];
#pragma unroll
            for (int j = 0; j < U; ++j) acc[j] = fmaf(w[i * U + j], xi, acc[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) h[j] = fmaxf(acc[j], 0.0f);
        w += ch.in_dim * U + U;
      }
      // ---- 64x64 layers: [res]*n_pre, mid, [res]*n_post ---------------------------------------
      const int n_dense = 2 * ch.n_pre + 1 + 2 * ch.n_post;
      const int mid_at = 2 * ch.n_pre;
      for (int s = 0; s < n_dense; ++s) {
        const bool is_mid = (s == mid_at);
        const int rel = (s < mid_at) ? s : s - mid_at - 1;
        const bool first_half = !is_mid && ((rel & 1) == 0);
        const float* Wt = w;
        if (is_mid) {
          const float4* rb = reinterpret_cast<const float4*>(P.rowbias + ((size_t)c * P.rb_stride + n) * U);
#pragma unroll
          for (int j4 = 0; j4 < U / 4; ++j4) {
            const float4 b = __ldg(rb + j4);
            acc[4 * j4 + 0] = b.x; acc[4 * j4 + 1] = b.y; acc[4 * j4 + 2] = b.z; acc[4 * j4 + 3] = b.w;
          }
          w += U * U;
        } else {
          const float* b = w + U * U;
          if (first_half) {
#pragma unroll
            for (int j = 0; j < U; ++j) {
              xs[j * TPB + tid] = h[j];
              acc[j] = b[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < U; ++j) acc[j] = b[j] + xs[j * TPB + tid];
          }
          w += U * U + U;
        }
        dense64(Wt, h, acc);
        const bool relu = is_mid ? (ch.mid_relu != 0) : true;
        if (relu) {
#pragma unroll
          for (int j = 0; j < U; ++j) h[j] = fmaxf(acc[j], 0.0f);
        } else {
#pragma unroll
          for (int j = 0; j < U; ++j) h[j] = acc[j];
        }
      }
      // ---- output layer (row-major out_W[out_dim][64]) ----------------------------------------
      float y[MMF_MAX_SD + 1];
#pragma unroll
      for (int o = 0; o < MMF_MAX_SD + 1; ++o) {
        y[o] = 0.0f;
        if (o < ch.out_dim) {
          float a = w[ch.out_dim * U + o];
#pragma unroll
          for (int k = 0; k < U; ++k) a = fmaf(w[o * U + k], h[k], a);
          y[o] = a;
        }
      }

      if (c == 0) {
        // gate, residual update, reparameterised process noise
        float g = 0.0f;
#pragma unroll
        for (int o = 0; o < MMF_MAX_SD + 1; ++o)
          if (o == sd) g = y[o];
        const float gate = 1.0f / (1.0f + expf(-g));
        float e[MMF_MAX_SD];
#pragma unroll
        for (int i = 0; i < MMF_MAX_SD; ++i) e[i] = (i < sd) ? P.eps[p * sd + i] : 0.0f;
#pragma unroll
        for (int i = 0; i < MMF_MAX_SD; ++i) {
          if (i < sd) {
            const float pred = x[i] + y[i] * gate;
            float noise = 0.0f;
#pragma unroll
            for (int j = 0; j < MMF_MAX_SD; ++j)
              if (j <= i && j < sd) noise = fmaf(P.q[i * sd + j], e[j], noise);
            xn[i] = pred + noise;
          }
        }
      } else {
        const float ll = y[0];
        if (P.ll_out != nullptr && live) P.ll_out[(size_t)(c - 1) * P.total + p] = ll;
        const float v = ll + (P.modw != nullptr ? __ldg(P.modw + (size_t)n * P.K + (c - 1)) : 0.0f);
        if (v > lse_m) {
          lse_s = lse_s * expf(lse_m - v) + 1.0f;
          lse_m = v;
        } else if (v > -INFINITY) {
          lse_s += expf(v - lse_m);
        }
      }
    }

    if (live) {
#pragma unroll
      for (int i = 0; i < MMF_MAX_SD; ++i)
        if (i < sd) P.states_out[p * sd + i] = xn[i];
      const float fused = (lse_m == -INFINITY) ? -INFINITY : lse_m + logf(lse_s);
      P.logw_out[p] = P.logw_in[p] + fused;
    }
  }
}

int launch_particle_chain_ffma(const mmf_pf_model* model, int N, int M, const float* states_in,
                               const float* eps, const float* rowbias, const float* logw_in,
                               const float* modw, uint32_t enabled, float* states_out,
                               float* logw_out, float* ll_out, cudaStream_t stream, int rb_stride) {
  PredictParams P;
  P.K = model->num_heads;
  P.chains[0] = to_dev(model->dynamics);
  int maxf = P.chains[0].floats();
  for (int k = 0; k < P.K; ++k) {
    P.chains[1 + k] = to_dev(model->heads[k]);
    maxf = maxf > P.chains[1 + k].floats() ? maxf : P.chains[1 + k].floats();
  }
  P.enabled = enabled;
  P.sd = model->state_dim;
  P.N = N;
  P.M = M;
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

  const size_t smem = ((size_t)((maxf + 3) & ~3) + (size_t)U * TPB) * sizeof(float);
  static thread_local int configured_dev = -1;
  static thread_local size_t window = 0;
  int dev = 0;
  MMF_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    int rc = opt_in_shared_memory(k_particle_chain_ffma, &window);
    if (rc) return rc;
    configured_dev = dev;
  }
  MMF_REQUIRE(smem <= window, "particle chain needs %zu B of shared memory (window %zu B)", smem, window);
  int sms = 148;
  MMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tiles = (P.total + TPB - 1) / TPB;
  const int grid = (int)(tiles < sms ? tiles : sms);
  k_particle_chain_ffma<<<grid, TPB, smem, stream>>>(P);
  MMF_LAUNCH_CHECK("k_particle_chain_ffma");
  return MMF_OK;
}

}  // namespace mmf
