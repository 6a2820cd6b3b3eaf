// normalize_resample.cuh -- device side of R6 (normalise, estimate) and R7 (resample + gather) for ONE trajectory,
// shared by k_normalize_resample (normalize_resample.cu) and the whole-sequence kernel (pf_loop_small.cu).
#pragma once

#include "kernels.cuh"
#include "pinned_math.cuh"

namespace mmf {

constexpr int NR_TPB = 256;
constexpr int SEG = 8;
constexpr int GROUP = 32 * SEG;

__device__ __forceinline__ float block_max(float v, float* scratch) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = scratch[0];
  for (int w = 1; w < NR_TPB / 32; ++w) r = fmaxf(r, scratch[w]);
  return r;
}

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = scratch[0];
  for (int w = 1; w < NR_TPB / 32; ++w) r += scratch[w];
  return r;
}

// The pinned search predicate is  q(c) = [(double) fl(c / total) < u]  (DESIGN.md "Resampling arithmetic").
// fl(c / total) is monotone in c, so q(c) <=> c < c*, with c* the smallest non-negative float for which q is
// false.  c* is found once per draw with a handful of exact evaluations of q around fl(u * total); the binary
// search then compares raw CDF entries with c*: every probe takes the same decision as the division-based
// definition (so indices stay bit-identical), at one FSETP per probe instead of an IEEE division.
__device__ __forceinline__ bool q_pred(float c, float total, double u) { return (double)__fdiv_rn(c, total) < u; }

__device__ __forceinline__ float cdf_threshold(float total, double u) {
  if (!(u > 0.0)) return 0.0f;  // q is false everywhere
  float c = (float)(u * (double)total);
  if (!(c > 0.0f)) c = __int_as_float(1);
  for (int it = 0; it < 64; ++it) {  // walk down while the predecessor already fails q
    const float pm = __int_as_float(__float_as_int(c) - 1);
    if (c > 0.0f && !q_pred(pm, total, u)) c = pm; else break;
  }
  for (int it = 0; it < 64 && q_pred(c, total, u); ++it) c = __int_as_float(__float_as_int(c) + 1);
  return c;
}

__device__ __forceinline__ int lower_bound_cdf(const float* cdf, int M, float total, double u) {
  // Fast path: search against c0 = fl(u * total), which is within ~2 ulp of c*.  A probe can only decide
  // differently from the definition if the probed entry lies between c0 and c*, i.e. within a few ulp of c0;
  // if any probed entry was that close, redo the search with the exact threshold.
  const float c0 = (float)(u * (double)total);
  const float band = fmaxf(c0 * 4.8e-7f, 1e-37f);  // >= 4 ulp of c0
  int lo = 0, hi = M;
  bool near = false;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    const float c = cdf[mid];
    near |= fabsf(c - c0) <= band;
    if (c < c0) lo = mid + 1; else hi = mid;
  }
  if (near || !(u > 0.0)) {
    const float cstar = cdf_threshold(total, u);
    lo = 0;
    hi = M;
    while (lo < hi) {
      const int mid = lo + ((hi - lo) >> 1);
      if (cdf[mid] < cstar) lo = mid + 1; else hi = mid;
    }
  }
  return lo < M - 1 ? lo : M - 1;
}

// NB independent draws searched in lock step: every probe step issues NB independent shared-memory loads, so the
// dependent load -> compare -> load chain of one binary search overlaps with the others (per-lane memory-level
// parallelism; the single-draw form left the kernel latency-bound).  Same decisions as lower_bound_cdf.
constexpr int NB = 4;
__device__ __forceinline__ void lower_bound_batch(const float* cdf, int M, float total, const double (&u)[NB],
                                                  int (&idx)[NB], int iters) {
  float c0[NB], band[NB];
  int lo[NB], hi[NB];
  bool near[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    c0[b] = (float)(u[b] * (double)total);
    band[b] = fmaxf(c0[b] * 4.8e-7f, 1e-37f);
    lo[b] = 0;
    hi[b] = M;
    near[b] = !(u[b] > 0.0);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const bool open = lo[b] < hi[b];
      const int mid = open ? lo[b] + ((hi[b] - lo[b]) >> 1) : 0;
      const float c = cdf[mid];
      near[b] |= open && fabsf(c - c0[b]) <= band[b];
      if (open) {
        if (c < c0[b]) lo[b] = mid + 1; else hi[b] = mid;
      }
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (near[b]) {  // rare: an entry within a few ulp of the threshold was probed -> exact predicate
      const float cstar = cdf_threshold(total, u[b]);
      int l = 0, h = M;
      while (l < h) {
        const int mid = l + ((h - l) >> 1);
        if (cdf[mid] < cstar) l = mid + 1; else h = mid;
      }
      lo[b] = l;
    }
    idx[b] = lo[b] < M - 1 ? lo[b] : M - 1;
  }
}

// Guide-table (bucket) inverse CDF for the warp-per-trajectory path.  f(c) = min(int(c * scale), K) is monotone
// in c, and guide[b] = number of CDF entries with f < b; hence for a draw with threshold c0 and b = f(c0) the
// lower bound lies in [guide[b], guide[b + 1]]: a bucket is hit with probability 1/K whatever it holds, so the
// expected range is M / K ~ 1 entry, against log2(M) dependent probes of a binary search.  Decisions are those
// of lower_bound_cdf: the final neighbours decide whether the exact threshold c* has to be consulted.
__device__ __forceinline__ int guided_lower_bound(const float* cdf, const uint16_t* guide, int M, int K, float scale,
                                                  float total, double u) {
  // the search key only has to be within a few ulp of the exact threshold (the band check below sends close
  // calls to the exact predicate), so it is formed in fp32: no fp64 arithmetic on the per-draw path
  const float uf = (float)u;
  const float c0 = uf * total;
  int b = (int)(c0 * scale);
  b = b < 0 ? 0 : (b > K ? K : b);
  int lo = guide[b], hi = guide[b + 1];
  while (hi - lo > 4) {  // crowded bucket: bisect down to a short run first
    const int mid = lo + ((hi - lo) >> 1);
    if (cdf[mid] < c0) lo = mid + 1; else hi = mid;
  }
  while (lo < hi && cdf[lo] < c0) ++lo;
  const float band = fmaxf(c0 * 9.6e-7f, 1e-37f);  // >= 8 ulp: fp32 key (<= 2 ulp off) + distance key..c* (<= 2 ulp)
  bool near = !(uf > 0.0f);
  if (lo < M) near |= fabsf(cdf[lo] - c0) <= band;
  if (lo > 0) near |= fabsf(cdf[lo - 1] - c0) <= band;
  if (near) {
    const float cstar = cdf_threshold(total, u);
    lo = 0;
    hi = M;
    while (lo < hi) {
      const int mid = lo + ((hi - lo) >> 1);
      if (cdf[mid] < cstar) lo = mid + 1; else hi = mid;
    }
  }
  return lo < M - 1 ? lo : M - 1;
}

// per-trajectory scratch arrays, in floats: [cdf Mpad | segoff Mpad/8 | gtot Mpad/256 + 4 | diff Mpad | guide (u16) Mpad + 2]
__host__ __device__ inline size_t trajectory_scratch_floats(int M, bool soft) {
  const size_t Mpad = ((size_t)(M + GROUP - 1) / GROUP) * GROUP;
  const size_t guide = Mpad <= 65536 ? (Mpad + 2 + 1) / 2 : 0;  // u16 guide table of the warp-per-trajectory search
  const size_t n = Mpad + Mpad / SEG + Mpad / GROUP + 4 + (soft ? Mpad : 0) + guide;
  return (n + 63) & ~(size_t)63;  // slices stay 256-byte aligned (vector loads in the serial scan)
}

// COOP = threads that cooperate on one trajectory:
//   COOP = NR_TPB: one CTA per trajectory (any M; block-wide reductions through shared memory);
//   COOP = 32    : one WARP per trajectory (M <= NR_WARP_MAX_M): eight independent trajectories per CTA, warp
//                  shuffles instead of block barriers, so the serial CDF chain of one trajectory (one lane, a
//                  dependent fp32 add every ~4 cycles) hides behind the streaming phases of the 40+ other warps
//                  resident on the SM.  This is the path of BASELINE configs C1/C3/C4 (M = 30 ... 1000).
// GLOBAL_WS = false: the trajectory's arrays live in shared memory (M up to ~48 k);
// GLOBAL_WS = true : they live in a caller-provided global workspace, one slice per CTA (any M: the
//                    1 k ... 1 M particle sweep of BASELINE config C5 runs through this instantiation).
template <int COOP>
__device__ __forceinline__ void coop_sync() {
  if (COOP == 32) __syncwarp(); else __syncthreads();
}
template <int COOP>
__device__ __forceinline__ float coop_max(float v, float* scratch) {
  return COOP == 32 ? warp_max(v) : block_max(v, scratch);
}
template <int COOP>
__device__ __forceinline__ float coop_sum(float v, float* scratch) {
  return COOP == 32 ? warp_sum(v) : block_sum(v, scratch);
}

constexpr int NR_WARP_TPB = 128;      // warp-per-trajectory CTAs: 4 trajectories each
constexpr int NR_WARP_CTAS_PER_SM = 7;  // 28 resident warps per SM: C3's 4096 trajectories are ONE wave on 148 SMs

// One trajectory `n`, executed by the COOP threads that cooperate on it (a warp, or the whole CTA): `cdf` is the base of
// the trajectory's scratch slice (trajectory_scratch_floats(M, soft) floats: shared memory or a global workspace),
// `scratch` / `cand_w` the block-reduction cells (unused when COOP == 32).  The caller synchronises the set before the call.
template <int COOP>
__device__ __forceinline__ void nr_trajectory(const ResampleParams& P, int n, float* cdf, float* scratch, int* cand_w) {
  const int M = P.M, sd = P.sd, tid = threadIdx.x;
  const int cid = tid % COOP;  // index inside the cooperating set
  const int Mpad = ((M + GROUP - 1) / GROUP) * GROUP;
  const bool soft = P.alpha < 1.0f;
  float* segoff = cdf + Mpad;     // Mpad / SEG
  float* gtot = segoff + Mpad / SEG;  // Mpad / GROUP (+4)
  float* diff = gtot + Mpad / GROUP + 4;  // Mpad (only when alpha < 1): logw - logits
  uint16_t* guide = reinterpret_cast<uint16_t*>(diff + (soft ? Mpad : 0));  // Mpad + 2 entries (warp path only)
  const bool resample = P.mode != MMF_RESAMPLE_NONE;
  float lmax;
  if (P.logits_in == nullptr) {
    // ---- normalise: logw = l - logsumexp(l) ----------------------------------------------------
    const float* lw = P.logw_unnorm + (size_t)n * M;
    float mx = -INFINITY;
    for (int base = 0; base < M; base += COOP * NB) {  // NB independent global loads in flight per thread
      float l[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int i = base + b * COOP + cid;
        l[b] = i < M ? lw[i] : -INFINITY;
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int i = base + b * COOP + cid;
        if (i < M) cdf[i] = l[b];
        mx = fmaxf(mx, l[b]);
      }
    }
    mx = coop_max<COOP>(mx, scratch);
    const float shift = (mx == -INFINITY || mx == INFINITY) ? 0.0f : mx;
    float s = 0.0f;
    for (int i = cid; i < M; i += COOP) s += expf(cdf[i] - shift);
    s = coop_sum<COOP>(s, scratch);
    const float lse = shift + logf(s);

    // ---- estimate ------------------------------------------------------------------------------
    const float* xs = P.states + (size_t)n * M * sd;
    float acc[MMF_MAX_SD] = {0.f, 0.f, 0.f, 0.f};
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    const bool weighted = P.estimation == MMF_ESTIMATE_WEIGHTED_AVERAGE;
    for (int base = 0; base < M; base += COOP * NB) {
      float xv[NB][MMF_MAX_SD];
      if (weighted) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int i = base + b * COOP + cid;
#pragma unroll
          for (int d = 0; d < MMF_MAX_SD; ++d) xv[b][d] = (d < sd && i < M) ? xs[(size_t)i * sd + d] : 0.0f;
        }
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int i = base + b * COOP + cid;
        if (i < M) {
          const float l = cdf[i] - lse;
          cdf[i] = l;
          if (P.logw_norm_out) P.logw_norm_out[(size_t)n * M + i] = l;
          if (!resample) P.logw_out[(size_t)n * M + i] = l;
          if (weighted) {
            const float wgt = expf(l);
#pragma unroll
            for (int d = 0; d < MMF_MAX_SD; ++d)
              if (d < sd) acc[d] = fmaf(wgt, xv[b][d], acc[d]);
          } else if (l > best) {
            best = l;
            best_i = i;
          }
        }
      }
    }
    if (P.estimation == MMF_ESTIMATE_WEIGHTED_AVERAGE) {
#pragma unroll
      for (int d = 0; d < MMF_MAX_SD; ++d) {
        if (d < sd) {
          const float v = coop_sum<COOP>(acc[d], scratch);
          if (cid == 0) P.est_out[(size_t)n * sd + d] = v;
        }
      }
    } else {
      const float gbest = coop_max<COOP>(best, scratch);
      int cand = (best == gbest) ? best_i : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
      int win = cand;
      if (COOP != 32) {
        __syncthreads();
        if ((tid & 31) == 0) cand_w[tid >> 5] = cand;
        __syncthreads();
        win = cand_w[0];
        for (int w = 1; w < NR_TPB / 32; ++w) win = min(win, cand_w[w]);
      }
      if (win == 0x7fffffff) win = 0;
      if (cid < sd) P.est_out[(size_t)n * sd + cid] = xs[(size_t)win * sd + cid];
    }
    if (!resample) return;

    // ---- logits (soft resampling mixes in the uniform, A.3) ------------------------------------
    float lm = -INFINITY;
    if (soft) {
      const float la = logf(P.alpha), lb = -logf((float)M) + logf(1.0f - P.alpha);
      for (int i = cid; i < M; i += COOP) {
        const float a = cdf[i] + la;
        const float m2 = fmaxf(a, lb);
        const float lg = m2 + logf(expf(a - m2) + expf(lb - m2));
        diff[i] = cdf[i] - lg;
        cdf[i] = lg;
        lm = fmaxf(lm, lg);
      }
    } else {
      for (int i = cid; i < M; i += COOP) lm = fmaxf(lm, cdf[i]);
    }
    lmax = coop_max<COOP>(lm, scratch);
  } else {
    const float* lg = P.logits_in + (size_t)n * M;
    float lm = -INFINITY;
    for (int i = cid; i < M; i += COOP) {
      const float l = lg[i];
      cdf[i] = l;
      lm = fmaxf(lm, l);
    }
    lmax = coop_max<COOP>(lm, scratch);
  }
  if (P.logits_out)
    for (int i = cid; i < M; i += COOP) P.logits_out[(size_t)n * M + i] = cdf[i];

  // ---- pinned softmax numerators ------------------------------------------------------------------
  for (int i = cid; i < Mpad; i += COOP) cdf[i] = (i < M) ? exp_pinned(cdf[i] - lmax) : 0.0f;
  coop_sync<COOP>();

  // ---- CDF --------------------------------------------------------------------------------------------
  const bool strict = (P.mode == MMF_RESAMPLE_MULTINOMIAL_STRICT || P.mode == MMF_RESAMPLE_SYSTEMATIC_STRICT);
  if (strict) {
    // The chain of M dependent fp32 adds IS the definition (torch.multinomial's CPU order), so it cannot be
    // parallelised; what can be done is keep everything but the adds off the critical path: 16 values are
    // fetched per iteration with vector loads (independent of the running sum) and written back vectorised,
    // leaving 4 cycles per element.  Other warps / CTAs on the SM overlap their streaming phases with it.
    if (cid == 0) {
      float run = 0.0f;
      float4* c4 = reinterpret_cast<float4*>(cdf);
      const int blocks16 = (M + 15) / 16;  // padded entries are zero: adding them is exact and harmless
      for (int b = 0; b < blocks16; ++b) {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = c4[b * 4 + q];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          run = __fadd_rn(run, v[q].x); v[q].x = run;
          run = __fadd_rn(run, v[q].y); v[q].y = run;
          run = __fadd_rn(run, v[q].z); v[q].z = run;
          run = __fadd_rn(run, v[q].w); v[q].w = run;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) c4[b * 4 + q] = v[q];
      }
    }
    coop_sync<COOP>();
  } else {
    const int lane = tid & 31, wid = (tid % COOP) >> 5;
    const int groups = Mpad / GROUP;
    for (int g = wid; g < groups; g += COOP / 32) {
      const int s = g * 32 + lane;
      float* e = cdf + (size_t)s * SEG;
      float run = 0.0f;
#pragma unroll
      for (int i = 0; i < SEG; ++i) {
        run = __fadd_rn(run, e[i]);
        e[i] = run;
      }
      float t = run;  // Kogge-Stone inclusive scan of the 32 segment totals
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, t, d);
        if (lane >= d) t = __fadd_rn(v, t);
      }
      const float excl = __shfl_up_sync(0xffffffffu, t, 1);
      segoff[s] = (lane == 0) ? 0.0f : excl;
      if (lane == 31) gtot[g] = t;
    }
    coop_sync<COOP>();
    if (cid == 0) {
      float run = 0.0f;
      for (int g = 0; g < groups; ++g) {
        const float t = gtot[g];
        gtot[g] = run;  // exclusive
        run = __fadd_rn(run, t);
      }
    }
    coop_sync<COOP>();
    for (int i = cid; i < M; i += COOP) {
      const float base = __fadd_rn(gtot[i / GROUP], segoff[i / SEG]);
      cdf[i] = __fadd_rn(base, cdf[i]);
    }
    coop_sync<COOP>();
  }
  const float total = cdf[M - 1];
  const int K = ((M + 31) / 32) * 32;  // guide buckets (<= Mpad)
  const float scale = (float)K / total;
  if (COOP == 32) {
    for (int i = cid; i < M; i += 32) {
      int fi = (int)(cdf[i] * scale);
      fi = fi < 0 ? 0 : (fi > K ? K : fi);
      int fp = -1;
      if (i > 0) {
        fp = (int)(cdf[i - 1] * scale);
        fp = fp < 0 ? 0 : (fp > K ? K : fp);
      }
      for (int b = fp + 1; b <= fi; ++b) guide[b] = (uint16_t)i;
    }
    int fl = (int)(total * scale);
    fl = fl < 0 ? 0 : (fl > K ? K : fl);
    for (int b = fl + 1 + cid; b <= K + 1; b += 32) guide[b] = (uint16_t)M;
    __syncwarp();
  }

  // ---- inverse CDF + gather -----------------------------------------------------------------------
  const bool systematic = (P.mode == MMF_RESAMPLE_SYSTEMATIC_STRICT || P.mode == MMF_RESAMPLE_SYSTEMATIC_FAST);
  const float uniform_lw = -logf((float)M);
  const double u0 = systematic ? P.uniforms[n] : 0.0;
  const int iters = 32 - __clz(M);  // probe steps until every [lo, hi) interval is empty
  for (int base = 0; base < P.M_out; base += COOP * NB) {
    double u[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int j = base + b * COOP + cid;
      u[b] = j >= P.M_out ? 0.5 : systematic ? (u0 + (double)j) / (double)P.M_out : P.uniforms[(size_t)n * P.M_out + j];
    }
    int idx[NB];
    if (COOP == 32) {
#pragma unroll
      for (int b = 0; b < NB; ++b) idx[b] = guided_lower_bound(cdf, guide, M, K, scale, total, u[b]);
    } else {
      lower_bound_batch(cdf, M, total, u, idx, iters);
    }
    float sv[NB][MMF_MAX_SD];
    if (P.states_out) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float* src = P.states + ((size_t)n * M + idx[b]) * sd;
#pragma unroll
        for (int d = 0; d < MMF_MAX_SD; ++d) sv[b][d] = d < sd ? src[d] : 0.0f;
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int j = base + b * COOP + cid;
      if (j < P.M_out) {
        if (P.idx_out) P.idx_out[(size_t)n * P.M_out + j] = idx[b];
        if (P.states_out) {
          float* dst = P.states_out + ((size_t)n * P.M_out + j) * sd;
#pragma unroll
          for (int d = 0; d < MMF_MAX_SD; ++d)
            if (d < sd) dst[d] = sv[b][d];
        }
        if (P.logw_out) P.logw_out[(size_t)n * P.M_out + j] = soft ? diff[idx[b]] : uniform_lw;
      }
    }
  }
}

}  // namespace mmf
