// resample_big.cu -- R6 (normalise, estimate) + R7 (resample, gather) for LONG trajectories (M > 2048 particles: BASELINE
// config C5's sweep from 4 K to 1 M particles per trajectory).
//
// The CTA-per-trajectory kernel (normalize_resample.cu) walks a whole trajectory with 256 threads: at M = 1 M and the 128
// trajectories a C5 tile holds that is 128 CTAs making ~8 dependent passes over 4 MB each -- 65 ms per step, 66 GB/s, half
// of the sweep point's time (profiles/r02_summary.md).  Here every pass is a grid over (chunk of 4096 particles,
// trajectory), so the whole GPU streams; only what the pinned arithmetic makes serial stays serial:
//   1  k_big_max        per-chunk max of the log-weights
//   2  k_big_sum        per-chunk sum exp(l - max)                                  -> lse (fixed order over the chunks)
//   3  k_big_norm_est   l - lse (normalised log-weights / logits), per-chunk estimate partials, e = EXP(logit - max logit)
//   4  k_big_finalize   estimate = sum of the partials in chunk order (or the first arg max)
//   5  STRICT: k_big_scan_strict   c_j = fl(c_{j-1} + e_j), the sequential fp32 sum torch.multinomial performs: ONE lane per
//              trajectory adds (4 cycles per particle: 2.2 ms at M = 1 M, the floor of this mode), the other lanes of
//              its warp stream the blocks in and out around it
//      FAST:   k_big_group_totals -> k_big_group_scan -> k_big_apply   the blocked order (segments of 8, Kogge-Stone over
//              the 32 segment totals of a 256-group, serial over the group totals), groups in parallel
//   6  k_big_coarse + k_big_search_gather   inverse CDF by a two-level lock-step batched binary search: every stride-th CDF
//              entry (<= 8192 of them) sits in shared memory, so only the last log2(stride) probes of a draw touch global
//              memory, inside one window of stride entries; same decisions
//              as lower_bound_batch (exact-threshold fallback near ties), then gather of the particle states
// Arithmetic per particle is that of nr_trajectory (same device functions); the log-sum-exp and the estimate are summed
// in another (fixed) order, so log-weights / estimates agree to rounding and the indices are bit-exact given the logits.
// Hard resampling only (alpha == 1); soft resampling keeps the CTA-per-trajectory kernel.
#include "normalize_resample.cuh"

namespace mmf {

constexpr int BG_TPB = 256;
constexpr int BG_PER_THREAD = 16;
constexpr int BG_CHUNK = BG_TPB * BG_PER_THREAD;  // 4096 particles = 16 groups of 256
constexpr int BG_SCAN_BLOCK = 1024;               // particles per iteration of the strict scan
constexpr int BG_COARSE = 8192;                   // entries of the shared-memory level of the search

struct BigWs {
  float* cdf;     // [N][Mpad]   e_j, then c_j
  float* pmax;    // [N][Mc]
  float* psum;    // [N][Mc]
  float* pest;    // [N][Mc][4]
  float* pbestv;  // [N][Mc]
  int* pbesti;    // [N][Mc]
  float* gtot;    // [N][Mpad / 256]
  float* coarse;  // [N][BG_COARSE]  every stride-th CDF entry (the last of each window)
  int Mc;
  size_t Mpad;
};

__host__ __device__ inline size_t big_mpad(int M) { return ((size_t)M + BG_CHUNK - 1) / BG_CHUNK * BG_CHUNK; }
static size_t big_ws_floats(int N, int M) {
  const size_t Mpad = big_mpad(M), Mc = Mpad / BG_CHUNK;
  return (size_t)N * (Mpad + Mc * 8 + Mpad / GROUP + BG_COARSE) + 64;
}
static BigWs big_ws(float* base, int N, int M) {
  BigWs w;
  w.Mpad = big_mpad(M);
  w.Mc = (int)(w.Mpad / BG_CHUNK);
  w.cdf = base;
  w.pmax = w.cdf + (size_t)N * w.Mpad;
  w.psum = w.pmax + (size_t)N * w.Mc;
  w.pest = w.psum + (size_t)N * w.Mc;
  w.pbestv = w.pest + (size_t)N * w.Mc * 4;
  w.pbesti = reinterpret_cast<int*>(w.pbestv + (size_t)N * w.Mc);
  w.gtot = reinterpret_cast<float*>(w.pbesti + (size_t)N * w.Mc);
  w.coarse = w.gtot + (size_t)N * (w.Mpad / GROUP);
  return w;
}

__device__ __forceinline__ float bg_block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < BG_TPB / 32; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float bg_block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < BG_TPB / 32; ++w) r += red[w];
  return r;
}
// max over the per-chunk maxima of trajectory n (every CTA of the trajectory computes the same value)
__device__ __forceinline__ float bg_traj_max(const BigWs& W, int n, float* red) {
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < W.Mc; c += BG_TPB) mx = fmaxf(mx, W.pmax[(size_t)n * W.Mc + c]);
  return bg_block_max(mx, red);
}

__global__ void __launch_bounds__(BG_TPB) k_big_max(const __grid_constant__ ResampleParams P, const BigWs W) {
  __shared__ float red[BG_TPB / 32];
  const int n = blockIdx.y, c = blockIdx.x;
  const float* src = (P.logits_in ? P.logits_in : P.logw_unnorm) + (size_t)n * P.M;
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < BG_PER_THREAD; ++k) {
    const int i = c * BG_CHUNK + k * BG_TPB + threadIdx.x;
    if (i < P.M) mx = fmaxf(mx, src[i]);
  }
  mx = bg_block_max(mx, red);
  if (threadIdx.x == 0) W.pmax[(size_t)n * W.Mc + c] = mx;
}

__global__ void __launch_bounds__(BG_TPB) k_big_sum(const __grid_constant__ ResampleParams P, const BigWs W) {
  __shared__ float red[BG_TPB / 32];
  const int n = blockIdx.y, c = blockIdx.x;
  const float* src = P.logw_unnorm + (size_t)n * P.M;
  const float mx = bg_traj_max(W, n, red);
  const float shift = (mx == -INFINITY || mx == INFINITY) ? 0.0f : mx;
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < BG_PER_THREAD; ++k) {
    const int i = c * BG_CHUNK + k * BG_TPB + threadIdx.x;
    if (i < P.M) s += expf(src[i] - shift);
  }
  s = bg_block_sum(s, red);
  if (threadIdx.x == 0) W.psum[(size_t)n * W.Mc + c] = s;
}

__global__ void __launch_bounds__(BG_TPB) k_big_norm_est(const __grid_constant__ ResampleParams P, const BigWs W) {
  __shared__ float red[BG_TPB / 32];
  __shared__ float lse_s;
  __shared__ int cand_s[BG_TPB / 32];
  const int n = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
  const int M = P.M, sd = P.sd;
  const bool given = P.logits_in != nullptr;
  const bool resample = P.mode != MMF_RESAMPLE_NONE;
  const float* src = (given ? P.logits_in : P.logw_unnorm) + (size_t)n * M;
  const float mx = bg_traj_max(W, n, red);
  float lse = 0.0f, lmax = mx;
  if (!given) {
    if (tid == 0) {  // the same sequential sum over the chunks in every CTA of the trajectory
      float s = 0.0f;
      for (int k = 0; k < W.Mc; ++k) s += W.psum[(size_t)n * W.Mc + k];
      const float shift = (mx == -INFINITY || mx == INFINITY) ? 0.0f : mx;
      lse_s = shift + logf(s);
    }
    __syncthreads();
    lse = lse_s;
    lmax = mx - lse;  // = max_i fl(l_i - lse): rounding is monotone
  }
  const float* xs = P.states ? P.states + (size_t)n * M * sd : nullptr;
  const bool weighted = P.estimation == MMF_ESTIMATE_WEIGHTED_AVERAGE;
  float acc[MMF_MAX_SD] = {0.f, 0.f, 0.f, 0.f};
  float best = -INFINITY;
  int best_i = 0x7fffffff;
#pragma unroll 4
  for (int k = 0; k < BG_PER_THREAD; ++k) {
    const int i = c * BG_CHUNK + k * BG_TPB + tid;
    float e = 0.0f;
    if (i < M) {
      const float lg = given ? src[i] : src[i] - lse;
      if (!given) {
        if (P.logw_norm_out) P.logw_norm_out[(size_t)n * M + i] = lg;
        if (!resample) P.logw_out[(size_t)n * M + i] = lg;
        if (weighted) {
          const float wgt = expf(lg);
#pragma unroll
          for (int d = 0; d < MMF_MAX_SD; ++d)
            if (d < sd) acc[d] = fmaf(wgt, xs[(size_t)i * sd + d], acc[d]);
        } else if (lg > best) {
          best = lg;
          best_i = i;
        }
      }
      if (P.logits_out) P.logits_out[(size_t)n * M + i] = lg;
      e = exp_pinned(lg - lmax);
    }
    if (resample) W.cdf[(size_t)n * W.Mpad + i] = e;  // padded entries: zeros
  }
  if (given) return;
  if (weighted) {
#pragma unroll
    for (int d = 0; d < MMF_MAX_SD; ++d) {
      const float v = bg_block_sum(acc[d], red);
      if (tid == 0) W.pest[((size_t)n * W.Mc + c) * 4 + d] = v;
    }
  } else {
    const float gbest = bg_block_max(best, red);
    int cand = (best == gbest) ? best_i : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    __syncthreads();
    if ((tid & 31) == 0) cand_s[tid >> 5] = cand;
    __syncthreads();
    if (tid == 0) {
      int win = cand_s[0];
      for (int w = 1; w < BG_TPB / 32; ++w) win = min(win, cand_s[w]);
      W.pbestv[(size_t)n * W.Mc + c] = gbest;
      W.pbesti[(size_t)n * W.Mc + c] = win;
    }
  }
}

__global__ void k_big_finalize(const __grid_constant__ ResampleParams P, const BigWs W) {
  const int n = blockIdx.x, lane = threadIdx.x;
  const int sd = P.sd;
  if (P.estimation == MMF_ESTIMATE_WEIGHTED_AVERAGE) {
    if (lane < sd) {
      float s = 0.0f;
      for (int c = 0; c < W.Mc; ++c) s += W.pest[((size_t)n * W.Mc + c) * 4 + lane];
      P.est_out[(size_t)n * sd + lane] = s;
    }
  } else {
    float best = -INFINITY;
    int win = 0x7fffffff;
    for (int c = 0; c < W.Mc; ++c) {  // chunks in index order: a later chunk only wins with a strictly larger value
      const float v = W.pbestv[(size_t)n * W.Mc + c];
      if (v > best) {
        best = v;
        win = W.pbesti[(size_t)n * W.Mc + c];
      }
    }
    if (win == 0x7fffffff) win = 0;
    if (lane < sd) P.est_out[(size_t)n * sd + lane] = P.states[((size_t)n * P.M + win) * sd + lane];
  }
}

// STRICT: the sequential fp32 sum.  One warp per trajectory: every lane brings 32 floats of the next block in (registers)
// while lane 0 adds through the current block in shared memory, 16 values per iteration as in nr_trajectory.
__global__ void __launch_bounds__(32) k_big_scan_strict(const __grid_constant__ ResampleParams P, const BigWs W) {
  __shared__ __align__(16) float buf[BG_SCAN_BLOCK];
  const int n = blockIdx.x, lane = threadIdx.x;
  float4* g4 = reinterpret_cast<float4*>(W.cdf + (size_t)n * W.Mpad);
  float4* b4 = reinterpret_cast<float4*>(buf);
  const int blocks = (int)(((size_t)P.M + BG_SCAN_BLOCK - 1) / BG_SCAN_BLOCK);  // Mpad is a multiple of the block: zeros behind M
  float4 nxt[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) nxt[q] = g4[q * 32 + lane];
  float run = 0.0f;
  for (int b = 0; b < blocks; ++b) {
#pragma unroll
    for (int q = 0; q < 8; ++q) b4[q * 32 + lane] = nxt[q];
    __syncwarp();
    if (b + 1 < blocks) {
#pragma unroll
      for (int q = 0; q < 8; ++q) nxt[q] = g4[(size_t)(b + 1) * (BG_SCAN_BLOCK / 4) + q * 32 + lane];
    }
    if (lane == 0) {
      for (int blk = 0; blk < BG_SCAN_BLOCK / 16; ++blk) {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = b4[blk * 4 + q];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          run = __fadd_rn(run, v[q].x); v[q].x = run;
          run = __fadd_rn(run, v[q].y); v[q].y = run;
          run = __fadd_rn(run, v[q].z); v[q].z = run;
          run = __fadd_rn(run, v[q].w); v[q].w = run;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) b4[blk * 4 + q] = v[q];
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) g4[(size_t)b * (BG_SCAN_BLOCK / 4) + q * 32 + lane] = b4[q * 32 + lane];
    __syncwarp();
  }
}

// FAST, pass a: total of every 256-group (segments of 8 summed serially, Kogge-Stone over the 32 segment totals)
template <bool APPLY>
__global__ void __launch_bounds__(BG_TPB) k_big_groups(const __grid_constant__ ResampleParams P, const BigWs W) {
  const int n = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* cdf = W.cdf + (size_t)n * W.Mpad;
  float* gtot = W.gtot + (size_t)n * (W.Mpad / GROUP);
  for (int gl = wid; gl < BG_CHUNK / GROUP; gl += BG_TPB / 32) {
    const size_t g = (size_t)blockIdx.x * (BG_CHUNK / GROUP) + gl;
    if (g * GROUP >= (size_t)P.M) break;
    float4* e4 = reinterpret_cast<float4*>(cdf + g * GROUP + lane * SEG);
    float4 a = e4[0], b = e4[1];
    float e[SEG] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float run = 0.0f;
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
      run = __fadd_rn(run, e[i]);
      e[i] = run;
    }
    float t = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float v = __shfl_up_sync(0xffffffffu, t, d);
      if (lane >= d) t = __fadd_rn(v, t);
    }
    if (!APPLY) {
      if (lane == 31) gtot[g] = t;
    } else {  // c_j = fl(fl(G[g] + S[s]) + local_j)
      const float excl = __shfl_up_sync(0xffffffffu, t, 1);
      const float base = __fadd_rn(gtot[g], lane == 0 ? 0.0f : excl);
#pragma unroll
      for (int i = 0; i < SEG; ++i) e[i] = __fadd_rn(base, e[i]);
      e4[0] = make_float4(e[0], e[1], e[2], e[3]);
      e4[1] = make_float4(e[4], e[5], e[6], e[7]);
    }
  }
}

// FAST, pass b: exclusive serial sum over the group totals of a trajectory
__global__ void k_big_group_scan(const __grid_constant__ ResampleParams P, const BigWs W) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= P.N) return;
  float* gtot = W.gtot + (size_t)n * (W.Mpad / GROUP);
  const int groups = (P.M + GROUP - 1) / GROUP;
  float run = 0.0f;
  for (int g = 0; g < groups; ++g) {
    const float t = gtot[g];
    gtot[g] = run;
    run = __fadd_rn(run, t);
  }
}

// window width of the two-level search: the smallest power of two with ceil(M / stride) <= BG_COARSE
__host__ __device__ inline int big_stride_log2(int M) {
  int s = 0;
  while (((long long)M + (1 << s) - 1) >> s > BG_COARSE) ++s;
  return s;
}

__global__ void __launch_bounds__(BG_TPB) k_big_coarse(const __grid_constant__ ResampleParams P, const BigWs W) {
  const int n = blockIdx.y, sl = big_stride_log2(P.M), G = (int)(((long long)P.M + (1 << sl) - 1) >> sl);
  const float* cdf = W.cdf + (size_t)n * W.Mpad;
  for (int g = blockIdx.x * BG_TPB + threadIdx.x; g < G; g += gridDim.x * BG_TPB) {
    const long long last = (((long long)g + 1) << sl) - 1;
    W.coarse[(size_t)n * BG_COARSE + g] = cdf[last < P.M ? last : P.M - 1];
  }
}

constexpr int BNB = 4;  // draws a thread searches in lock step (8 measured slower: 128 registers, two CTAs per SM: 13.9 vs 12.6 ms at M = 1 M)

// lower_bound_batch with the first levels in shared memory: `coarse[g]` = the last CDF entry of window g.  Same probes'
// decisions, same near-tie rule (any probed entry within the band of the key sends the draw to the exact threshold).
__device__ __forceinline__ void lower_bound_two_level(const float* cdf, const float* coarse, int G, int sl, int M, float total,
                                                      const double (&u)[BNB], int (&idx)[BNB], bool monotone) {
  float c0[BNB], band[BNB];
  int lo[BNB], hi[BNB];
  bool near[BNB];
#pragma unroll
  for (int b = 0; b < BNB; ++b) {
    c0[b] = (float)(u[b] * (double)total);
    band[b] = fmaxf(c0[b] * 4.8e-7f, 1e-37f);
    lo[b] = 0;
    hi[b] = G;
    near[b] = !(u[b] > 0.0);
  }
  const int it1 = 32 - __clz(G);
  for (int it = 0; it < it1; ++it) {
#pragma unroll
    for (int b = 0; b < BNB; ++b) {
      const bool open = lo[b] < hi[b];
      const int mid = open ? lo[b] + ((hi[b] - lo[b]) >> 1) : 0;
      const float c = coarse[mid];
      near[b] |= open && fabsf(c - c0[b]) <= band[b];
      if (open) {
        if (c < c0[b]) lo[b] = mid + 1; else hi[b] = mid;
      }
    }
  }
  bool past[BNB];  // the key lies beyond the last entry: lower bound = M
#pragma unroll
  for (int b = 0; b < BNB; ++b) {
    past[b] = lo[b] >= G;
    const long long first = (long long)lo[b] << sl, end = first + (1 << sl) - 1;  // the window's last entry is >= key
    lo[b] = past[b] ? M : (int)first;
    hi[b] = past[b] ? M : (int)(end < M ? end : M - 1);
  }
  for (int it = 0; it < sl; ++it) {
#pragma unroll
    for (int b = 0; b < BNB; ++b) {
      const bool open = lo[b] < hi[b];
      const int mid = open ? lo[b] + ((hi[b] - lo[b]) >> 1) : 0;
      const float c = cdf[mid];
      near[b] |= open && fabsf(c - c0[b]) <= band[b];
      if (open) {
        if (c < c0[b]) lo[b] = mid + 1; else hi[b] = mid;
      }
    }
  }
  // A probed entry within a few ulp of the key: decide with the exact threshold c*.  At M ~ 1 M the CDF entries are about
  // an ulp apart, so this is the COMMON case here (the CTA-per-trajectory kernel redoes a full binary search for it:
  // most of its 65 ms at M = 1 M).  c* is within a few ulp of the key, hence the exact lower bound is within a few entries
  // of the one just found: walk there.  Only for the STRICT modes, whose sequential sums of non-negative terms are
  // monotone, so that the walk ends exactly where the definition's binary search does; the blocked sums of the FAST modes
  // can dip by an ulp at a segment boundary, and there the definition IS the probe sequence of the full search.  A long
  // plateau of equal entries (dead particles) falls back to the full search as well.
#pragma unroll
  for (int b = 0; b < BNB; ++b) {
    if (near[b]) {
      const float cstar = cdf_threshold(total, u[b]);
      int l = lo[b] < M ? lo[b] : M;
      int steps = monotone ? 0 : 32;
      while (l > 0 && !(cdf[l - 1] < cstar) && steps < 32) { --l; ++steps; }
      while (l < M && cdf[l] < cstar && steps < 32) { ++l; ++steps; }
      if (steps >= 32) {
        int h = M;
        l = 0;
        while (l < h) {
          const int mid = l + ((h - l) >> 1);
          if (cdf[mid] < cstar) l = mid + 1; else h = mid;
        }
      }
      lo[b] = l;
    }
    idx[b] = lo[b] < M - 1 ? lo[b] : M - 1;
  }
}

__global__ void __launch_bounds__(BG_TPB) k_big_search_gather(const __grid_constant__ ResampleParams P, const BigWs W) {
  __shared__ float coarse[BG_COARSE];
  const int n = blockIdx.y, tid = threadIdx.x;
  const int M = P.M, sd = P.sd, S = P.M_out;
  const float* cdf = W.cdf + (size_t)n * W.Mpad;
  const int sl = big_stride_log2(M), G = (int)(((long long)M + (1 << sl) - 1) >> sl);
  for (int g = tid; g < G; g += BG_TPB) coarse[g] = W.coarse[(size_t)n * BG_COARSE + g];
  __syncthreads();
  const float total = cdf[M - 1];
  const bool systematic = P.mode == MMF_RESAMPLE_SYSTEMATIC_STRICT || P.mode == MMF_RESAMPLE_SYSTEMATIC_FAST;
  const float uniform_lw = -logf((float)M);
  const double u0 = systematic ? P.uniforms[n] : 0.0;
  const int j0 = blockIdx.x * BG_CHUNK;
  for (int base = j0; base < j0 + BG_CHUNK && base < S; base += BG_TPB * BNB) {
    double u[BNB];
#pragma unroll
    for (int b = 0; b < BNB; ++b) {
      const int j = base + b * BG_TPB + tid;
      u[b] = j >= S ? 0.5 : systematic ? (u0 + (double)j) / (double)S : P.uniforms[(size_t)n * S + j];
    }
    int idx[BNB];
    lower_bound_two_level(cdf, coarse, G, sl, M, total, u, idx,
                          P.mode == MMF_RESAMPLE_MULTINOMIAL_STRICT || P.mode == MMF_RESAMPLE_SYSTEMATIC_STRICT);
    float sv[BNB][MMF_MAX_SD];
    if (P.states_out) {
#pragma unroll
      for (int b = 0; b < BNB; ++b) {
        const float* src = P.states + ((size_t)n * M + idx[b]) * sd;
#pragma unroll
        for (int d = 0; d < MMF_MAX_SD; ++d) sv[b][d] = d < sd ? src[d] : 0.0f;
      }
    }
#pragma unroll
    for (int b = 0; b < BNB; ++b) {
      const int j = base + b * BG_TPB + tid;
      if (j < S) {
        if (P.idx_out) P.idx_out[(size_t)n * S + j] = idx[b];
        if (P.states_out) {
          float* dst = P.states_out + ((size_t)n * S + j) * sd;
#pragma unroll
          for (int d = 0; d < MMF_MAX_SD; ++d)
            if (d < sd) dst[d] = sv[b][d];
        }
        if (P.logw_out) P.logw_out[(size_t)n * S + j] = uniform_lw;
      }
    }
  }
}

constexpr int BIG_M_MIN = 2048;  // trajectories longer than the warp-per-trajectory limit take the multi-pass path (measured
                                 // against the CTA-per-trajectory kernel: M = 4 K 5.1 vs 6.8 ms, 16 K 6.0 vs 11.3, 64 K 7.0
                                 // vs 16.3, 256 K 8.9 vs 31.1, 1 M 12.6 vs 65.3 ms per step of 134 M particles)

// MMF_RESAMPLE_BIG (read once when the library is loaded): 0 keeps the CTA-per-trajectory kernel (A/B timing), a value
// > 1 replaces the threshold BIG_M_MIN.
static int big_threshold() {
  static const int t = [] {
    const char* env = getenv("MMF_RESAMPLE_BIG");
    const int v = env ? atoi(env) : 1;
    return v == 0 ? 0x7fffffff : (v > 1 ? v : BIG_M_MIN);
  }();
  return t;
}

bool resample_big_applies(int N, int M, bool soft) { return M > big_threshold() && !soft && N <= 65535; }

size_t resample_big_workspace_bytes(int N, int M) { return big_ws_floats(N, M) * sizeof(float); }

int launch_resample_big(const ResampleParams& P, void* workspace, cudaStream_t stream) {
  MMF_REQUIRE(workspace != nullptr, "normalize_resample: M=%d needs a workspace of mmf_pf_resample_workspace_bytes() bytes", P.M);
  MMF_REQUIRE(((uintptr_t)workspace & 15) == 0, "normalize_resample: the workspace must be 16-byte aligned");
  const BigWs W = big_ws(static_cast<float*>(workspace), P.N, P.M);
  const bool given = P.logits_in != nullptr;
  const bool resample = P.mode != MMF_RESAMPLE_NONE;
  const dim3 grid(W.Mc, P.N);
  k_big_max<<<grid, BG_TPB, 0, stream>>>(P, W);
  if (!given) k_big_sum<<<grid, BG_TPB, 0, stream>>>(P, W);
  k_big_norm_est<<<grid, BG_TPB, 0, stream>>>(P, W);
  if (!given) k_big_finalize<<<P.N, 32, 0, stream>>>(P, W);
  if (resample) {
    const bool strict = P.mode == MMF_RESAMPLE_MULTINOMIAL_STRICT || P.mode == MMF_RESAMPLE_SYSTEMATIC_STRICT;
    if (strict) {
      k_big_scan_strict<<<P.N, 32, 0, stream>>>(P, W);
    } else {
      k_big_groups<false><<<grid, BG_TPB, 0, stream>>>(P, W);
      k_big_group_scan<<<(P.N + 31) / 32, 32, 0, stream>>>(P, W);
      k_big_groups<true><<<grid, BG_TPB, 0, stream>>>(P, W);
    }
    k_big_coarse<<<dim3(4, P.N), BG_TPB, 0, stream>>>(P, W);
    const dim3 sgrid((unsigned)(((size_t)P.M_out + BG_CHUNK - 1) / BG_CHUNK), P.N);
    k_big_search_gather<<<sgrid, BG_TPB, 0, stream>>>(P, W);
  }
  MMF_LAUNCH_CHECK("k_big_*");
  return MMF_OK;
}

}  // namespace mmf
