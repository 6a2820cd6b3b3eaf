/* ORACLE (test infrastructure) -- C restatement of the PINNED resampling arithmetic.
 *
 * "Bit-exact resampling indices" is only meaningful against a fully specified arithmetic
 * (SURVEY.md section 7 hard part 1; Appendix A.6).  The specification is DESIGN.md "Resampling
 * arithmetic"; this file restates it independently of the CUDA sources with plain C99
 * (fmaf / rintf, -ffp-contract=off), every operation a single IEEE-754 binary32
 * round-to-nearest-even op:
 *
 *   m   = max_j logit_j                                   (exact)
 *   e_j = EXP(logit_j - m)                                (mmf_exp_pinned below)
 *   STRICT: c_j = fl(c_{j-1} + e_j)                       sequential fp32 running sum -- what
 *           torch.multinomial does on CPU (A.6: 0 mismatches / 256,000 draws)
 *   FAST  : segments of 8 consecutive elements (serial running sum, zero padded to a multiple
 *           of 256), groups of 32 segments (Kogge-Stone inclusive scan of the segment totals,
 *           steps 1,2,4,8,16), serial exclusive running sum over the group totals;
 *           c_j = fl(fl(G_excl[g] + S_excl[s]) + local_j)
 *   idx = lower_bound_j { (double) fl(c_j / c_{M-1}) >= u }, clamped to M-1
 *   multinomial: u = the float64 uniforms as given, (N, S) trajectory-major;
 *   systematic : u_j = (u0 + j) / S in float64, one u0 per trajectory.
 *
 * Built by oracle/pinned/Makefile into oracle/pinned/libmmf_pinned.so (git-ignored).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float as_float(int32_t bits) {
  float f;
  memcpy(&f, &bits, sizeof f);
  return f;
}

float mmf_exp_pinned(float x) {
  if (!(x >= -87.0f)) return 0.0f;
  const float t = x * 1.44269504088896341f;
  const float n = rintf(t);
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  const float r2 = r * r;
  p = fmaf(p, r2, r);
  p = p + 1.0f;
  return p * as_float(((int32_t)n + 127) << 23);
}

void mmf_exp_pinned_array(const float* x, float* y, int64_t n) {
  for (int64_t i = 0; i < n; ++i) y[i] = mmf_exp_pinned(x[i]);
}

#define SEG 8
#define GROUP 256

/* cdf must hold Mpad = roundup(M, 256) floats */
static void build_cdf(const float* logits, int M, int fast, float* cdf) {
  const int Mpad = ((M + GROUP - 1) / GROUP) * GROUP;
  float m = -INFINITY;
  for (int j = 0; j < M; ++j) m = logits[j] > m ? logits[j] : m;
  for (int j = 0; j < Mpad; ++j) cdf[j] = j < M ? mmf_exp_pinned(logits[j] - m) : 0.0f;
  if (!fast) {
    float run = 0.0f;
    for (int j = 0; j < M; ++j) {
      run = run + cdf[j];
      cdf[j] = run;
    }
    return;
  }
  const int groups = Mpad / GROUP;
  float* segoff = (float*)malloc(sizeof(float) * (size_t)(Mpad / SEG));
  float* gtot = (float*)malloc(sizeof(float) * (size_t)groups);
  for (int g = 0; g < groups; ++g) {
    float t[32], nxt[32];
    for (int lane = 0; lane < 32; ++lane) {
      float* e = cdf + ((size_t)g * 32 + lane) * SEG;
      float run = 0.0f;
      for (int i = 0; i < SEG; ++i) {
        run = run + e[i];
        e[i] = run;
      }
      t[lane] = run;
    }
    for (int d = 1; d < 32; d <<= 1) { /* Kogge-Stone: every lane reads the previous round */
      for (int lane = 0; lane < 32; ++lane) nxt[lane] = lane >= d ? t[lane - d] + t[lane] : t[lane];
      memcpy(t, nxt, sizeof t);
    }
    for (int lane = 0; lane < 32; ++lane) segoff[g * 32 + lane] = lane == 0 ? 0.0f : t[lane - 1];
    gtot[g] = t[31];
  }
  float run = 0.0f;
  for (int g = 0; g < groups; ++g) {
    const float t = gtot[g];
    gtot[g] = run;
    run = run + t;
  }
  for (int j = 0; j < M; ++j) {
    const float base = gtot[j / GROUP] + segoff[j / SEG];
    cdf[j] = base + cdf[j];
  }
  free(segoff);
  free(gtot);
}

static int64_t lower_bound(const float* cdf, int M, float total, double u) {
  int lo = 0, hi = M;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    const float c = cdf[mid] / total;
    if ((double)c < u) lo = mid + 1; else hi = mid;
  }
  return lo < M - 1 ? lo : M - 1;
}

/* mode: 1 multinomial strict, 2 multinomial fast, 3 systematic strict, 4 systematic fast
 * (the MMF_RESAMPLE_* values of include/mmf_b200.h).  cdf_out (optional) receives the
 * un-normalised CDF, N x M. */
int mmf_pinned_resample(int32_t N, int32_t M, int32_t S, const float* logits, int32_t mode, const double* uniforms,
                        int64_t* idx_out, float* cdf_out) {
  if (mode < 1 || mode > 4 || M < 1 || S < 1) return -1;
  const int fast = (mode == 2 || mode == 4), systematic = (mode >= 3);
  const int Mpad = ((M + GROUP - 1) / GROUP) * GROUP;
  float* cdf = (float*)malloc(sizeof(float) * (size_t)Mpad);
  for (int n = 0; n < N; ++n) {
    build_cdf(logits + (size_t)n * M, M, fast, cdf);
    if (cdf_out) memcpy(cdf_out + (size_t)n * M, cdf, sizeof(float) * (size_t)M);
    const float total = cdf[M - 1];
    for (int j = 0; j < S; ++j) {
      const double u = systematic ? (uniforms[n] + (double)j) / (double)S : uniforms[(size_t)n * S + j];
      idx_out[(size_t)n * S + j] = lower_bound(cdf, M, total, u);
    }
  }
  free(cdf);
  return 0;
}
