// pinned_math.cuh -- arithmetic that DEFINES "bit-exact resampling" (DESIGN.md "Resampling
// arithmetic").  Every operation is a single IEEE-754 binary32 op with round-to-nearest-even
// (explicit intrinsics so nvcc cannot contract or reassociate), hence reproducible on any CPU:
// oracle/pinned/mmf_pinned.c restates the same sequence with fmaf()/rintf().
#pragma once
#include <cuda_runtime.h>

namespace mmf {

// exp(x) for x <= 0 (x = logit - max).  Returns exactly 0 below -87 (keeps the result normal).
__device__ __forceinline__ float exp_pinned(float x) {
  if (!(x >= -87.0f)) return 0.0f;  // also catches -inf and NaN
  const float t = __fmul_rn(x, 1.44269504088896341f);
  const float n = rintf(t);                       // round half to even
  float r = __fmaf_rn(n, -0.693359375f, x);       // Cody-Waite, ln2 = 0.693359375 - 2.12194440e-4
  r = __fmaf_rn(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = __fmaf_rn(p, r, 1.3981999507e-3f);
  p = __fmaf_rn(p, r, 8.3334519073e-3f);
  p = __fmaf_rn(p, r, 4.1665795894e-2f);
  p = __fmaf_rn(p, r, 1.6666665459e-1f);
  p = __fmaf_rn(p, r, 5.0000001201e-1f);
  const float r2 = __fmul_rn(r, r);
  p = __fmaf_rn(p, r2, r);
  p = __fadd_rn(p, 1.0f);
  const int e = (int)n + 127;                     // n in [-126, 0] here => e in [1, 127]
  return __fmul_rn(p, __int_as_float(e << 23));
}

}  // namespace mmf
