// common.cuh -- shared helpers for libmmf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmf_b200.h"

namespace mmf {

constexpr int U = MMF_UNITS;          // hidden width of every hot MLP (ref: units=64 everywhere)
constexpr int RES_FLOATS = 2 * (U * U + U);

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);

#define MMF_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      mmf::set_error(__VA_ARGS__);              \
      return MMF_E_INVALID;                     \
    }                                           \
  } while (0)

#define MMF_CUDA(call)                                          \
  do {                                                          \
    cudaError_t err__ = (call);                                 \
    if (err__ != cudaSuccess) return mmf::cuda_fail(err__, #call); \
  } while (0)

#define MMF_LAUNCH_CHECK(name)                                   \
  do {                                                           \
    cudaError_t err__ = cudaGetLastError();                      \
    if (err__ != cudaSuccess) return mmf::cuda_fail(err__, name); \
  } while (0)

// Device-side view of an mmf_chain (POD copy so it can travel in kernel parameters).
struct ChainDev {
  int in_dim, n_pre, mid_relu, n_post, out_dim;
  const float* w;
  __host__ __device__ int floats() const {
    return in_dim * U + U + n_pre * RES_FLOATS + U * U + n_post * RES_FLOATS + out_dim * U + out_dim;
  }
};

inline ChainDev to_dev(const mmf_chain& c) {
  ChainDev d;
  d.in_dim = c.in_dim;
  d.n_pre = c.n_pre_res;
  d.mid_relu = c.mid_relu;
  d.n_post = c.n_post_res;
  d.out_dim = c.out_dim;
  d.w = c.w;
  return d;
}

struct TrajRowsDev {
  int in_dim, has_encoder;
  const float* w;
};

inline TrajRowsDev to_dev(const mmf_traj_rows& r) {
  TrajRowsDev d;
  d.in_dim = r.in_dim;
  d.has_encoder = r.has_encoder;
  d.w = r.w;
  return d;
}

int validate_chain(const mmf_chain& c, int sd, int out_dim, const char* what);

// Opt the kernel in to the largest dynamic shared-memory window the device offers (227 KiB on
// sm_100 minus the kernel's static shared memory); returns that window in *avail.
template <typename Kernel>
int opt_in_shared_memory(Kernel kernel, size_t* avail) {
  int dev = 0, optin = 0;
  MMF_CUDA(cudaGetDevice(&dev));
  MMF_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  cudaFuncAttributes fa;
  MMF_CUDA(cudaFuncGetAttributes(&fa, kernel));
  const size_t window = (size_t)optin - fa.sharedSizeBytes;
  MMF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)window));
  *avail = window;
  return MMF_OK;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mmf
