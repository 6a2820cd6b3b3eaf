// reweight_train.cu -- R5 + R6 of the BPTT step as ONE forward and ONE backward kernel:
//   fused   = logsumexp_k(ll[k] + w[n,k]) over the enabled heads    (ref: crossmodal/base_models/crossmodal_pf.py:132-139)
//   u       = logw_in + fused                                        (A.3 `self.particle_log_weights + log-likelihoods`)
//   logw_n  = u - logsumexp_m(u)                                     (A.3 normalisation)
//   est     = sum_m exp(logw_n) x_m                                  (A.3 weighted-average estimate)
// and the reverse-mode derivative of exactly that: given d_est (N, sd) and d_logw_n (N, M) it returns d_ll (K, N, M),
// d_w (N, K) and d_logw_in (N, M).  As torch ops the same thing is ~12 small kernels forward and ~25 backward per filter
// step of the training graph (stack / add / logsumexp x 2 / sub / exp / mul / sum and their autograd); the particle states
// carry no gradient here (frozen dynamics, the reference's training setting: ref: scripts/push_task/train_push.py:154,213).
// A warp owns a trajectory; nothing is saved between the two kernels: the backward recomputes v, fused, u and the weights.
#include "kernels.cuh"

namespace mmf {

struct ReweightParams {
  int N, M, K, sd;
  uint32_t enabled;
  const float* ll;       // (K, N, M) planes (disabled planes are ignored)
  const float* w;        // (N, K) modality log-weights or null
  const float* logw_in;  // (N, M)
  const float* states;   // (N, M, sd)
  float* logw_out;       // fwd: (N, M)
  float* est_out;        // fwd: (N, sd)
  const float* d_est;    // bwd: (N, sd) or null
  const float* d_logw;   // bwd: (N, M) or null
  float* d_ll;           // bwd: (K, N, M)   (zeros in disabled planes)
  float* d_w;            // bwd: (N, K) or null
  float* d_logw_in;      // bwd: (N, M)
};

// u_m = logw_in + logsumexp_k(v_k); also returns the per-head v_k and the fused value (for the backward's softmax)
__device__ __forceinline__ float fused_logw(const ReweightParams& P, int n, int m, float (&v)[MMF_MAX_HEADS], float& fused) {
  const size_t nm = (size_t)n * P.M + m, plane = (size_t)P.N * P.M;
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < MMF_MAX_HEADS; ++k) {
    v[k] = -INFINITY;
    if (k < P.K && ((P.enabled >> k) & 1u)) {
      v[k] = P.ll[k * plane + nm] + (P.w ? P.w[(size_t)n * P.K + k] : 0.0f);
      mx = fmaxf(mx, v[k]);
    }
  }
  const float shift = (mx == -INFINITY || mx == INFINITY) ? 0.0f : mx;
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < MMF_MAX_HEADS; ++k)
    if (k < P.K && ((P.enabled >> k) & 1u)) s += expf(v[k] - shift);
  fused = shift + logf(s);
  return P.logw_in[nm] + fused;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(128) k_reweight_train(const __grid_constant__ ReweightParams P) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < P.N; n += warps) {
    // ---- pass 1: u and its log-sum-exp over the particles ------------------------------------------------------------
    float mx = -INFINITY;
    for (int m = lane; m < P.M; m += 32) {
      float v[MMF_MAX_HEADS], f;
      mx = fmaxf(mx, fused_logw(P, n, m, v, f));
    }
    mx = warp_max(mx);
    const float shift = (mx == -INFINITY || mx == INFINITY) ? 0.0f : mx;
    float s = 0.0f;
    for (int m = lane; m < P.M; m += 32) {
      float v[MMF_MAX_HEADS], f;
      s += expf(fused_logw(P, n, m, v, f) - shift);
    }
    const float lse = shift + logf(warp_sum(s));
    if (!BACKWARD) {
      // ---- normalised log-weights and the estimate ----------------------------------------------------------------------
      float acc[MMF_MAX_SD] = {0.f, 0.f, 0.f, 0.f};
      for (int m = lane; m < P.M; m += 32) {
        float v[MMF_MAX_HEADS], f;
        const float l = fused_logw(P, n, m, v, f) - lse;
        P.logw_out[(size_t)n * P.M + m] = l;
        const float wgt = expf(l);
#pragma unroll
        for (int d = 0; d < MMF_MAX_SD; ++d)
          if (d < P.sd) acc[d] = fmaf(wgt, P.states[((size_t)n * P.M + m) * P.sd + d], acc[d]);
      }
#pragma unroll
      for (int d = 0; d < MMF_MAX_SD; ++d) {
        if (d < P.sd) {
          const float t = warp_sum(acc[d]);
          if (lane == 0) P.est_out[(size_t)n * P.sd + d] = t;
        }
      }
    } else {
      // ---- g_m = d_logw_n[m] + w_m (x_m . d_est);  d_u[m] = g_m - w_m sum_j g_j ------------------------------------------------
      float de[MMF_MAX_SD];
#pragma unroll
      for (int d = 0; d < MMF_MAX_SD; ++d) de[d] = (P.d_est && d < P.sd) ? P.d_est[(size_t)n * P.sd + d] : 0.0f;
      float gsum = 0.0f;
      for (int m = lane; m < P.M; m += 32) {
        float v[MMF_MAX_HEADS], f;
        const float wgt = expf(fused_logw(P, n, m, v, f) - lse);
        float dot = 0.0f;
#pragma unroll
        for (int d = 0; d < MMF_MAX_SD; ++d)
          if (d < P.sd) dot = fmaf(P.states[((size_t)n * P.M + m) * P.sd + d], de[d], dot);
        gsum += (P.d_logw ? P.d_logw[(size_t)n * P.M + m] : 0.0f) + wgt * dot;
      }
      gsum = warp_sum(gsum);
      float dw[MMF_MAX_HEADS] = {0.f, 0.f, 0.f, 0.f};
      const size_t plane = (size_t)P.N * P.M;
      for (int m = lane; m < P.M; m += 32) {
        float v[MMF_MAX_HEADS], f;
        const float wgt = expf(fused_logw(P, n, m, v, f) - lse);
        float dot = 0.0f;
#pragma unroll
        for (int d = 0; d < MMF_MAX_SD; ++d)
          if (d < P.sd) dot = fmaf(P.states[((size_t)n * P.M + m) * P.sd + d], de[d], dot);
        const float g = (P.d_logw ? P.d_logw[(size_t)n * P.M + m] : 0.0f) + wgt * dot;
        const float du = g - wgt * gsum;
        const size_t nm = (size_t)n * P.M + m;
        P.d_logw_in[nm] = du;
#pragma unroll
        for (int k = 0; k < MMF_MAX_HEADS; ++k) {
          if (k < P.K) {
            float dv = 0.0f;
            if (((P.enabled >> k) & 1u) && v[k] > -INFINITY) dv = du * expf(v[k] - f);  // softmax over the enabled heads
            P.d_ll[k * plane + nm] = dv;
            dw[k] += dv;
          }
        }
      }
      if (P.d_w) {
#pragma unroll
        for (int k = 0; k < MMF_MAX_HEADS; ++k) {
          if (k < P.K) {
            const float t = warp_sum(dw[k]);
            if (lane == 0) P.d_w[(size_t)n * P.K + k] = t;
          }
        }
      }
    }
  }
}

int launch_reweight_train(int N, int M, int K, int sd, uint32_t enabled, const float* ll, const float* w, const float* logw_in,
                          const float* states, float* logw_out, float* est_out, const float* d_est, const float* d_logw,
                          float* d_ll, float* d_w, float* d_logw_in, bool backward, cudaStream_t stream) {
  if (N == 0 || M == 0) return MMF_OK;
  ReweightParams P;
  P.N = N; P.M = M; P.K = K; P.sd = sd; P.enabled = enabled;
  P.ll = ll; P.w = w; P.logw_in = logw_in; P.states = states; P.logw_out = logw_out; P.est_out = est_out;
  P.d_est = d_est; P.d_logw = d_logw; P.d_ll = d_ll; P.d_w = d_w; P.d_logw_in = d_logw_in;
  int dev = 0, sms = 148;
  MMF_CUDA(cudaGetDevice(&dev));
  MMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long blocks = ((long long)N + 3) / 4;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  if (backward) k_reweight_train<true><<<(int)blocks, 128, 0, stream>>>(P);
  else k_reweight_train<false><<<(int)blocks, 128, 0, stream>>>(P);
  MMF_LAUNCH_CHECK("k_reweight_train");
  return MMF_OK;
}

}  // namespace mmf
