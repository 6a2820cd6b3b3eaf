// row_mlp.cu -- the per-trajectory MLPs of the path as ONE kernel per module (R9):
// position / sensor observation encoders (ref: crossmodal/push_models/layers.py:107-136), the particle-filter and
// Kalman-filter crossmodal weight models (ref: crossmodal/push_models/crossmodal_pf.py:72-104,
// crossmodal/door_models/crossmodal_kf.py:134-167) and the virtual-sensor head (ref: crossmodal/door_models/kf.py:81-126).
//
// These are stacks of Linear / ReLU / Sigmoid / residual blocks, 3 ... 192 inputs wide, evaluated once per (step,
// trajectory) row: 10-90 k MAC per row, all weights L2 resident.  As torch modules each stack is 10-40 library
// launches (cuBLAS SIMT sgemm + ATen elementwise) per call; here the host compiles a stack into a short program of
// fused ops   dst = act(W src + b [+ residual])   over per-row scratch slots and one launch runs the whole program:
// a warp owns R = 8 rows, lane j computes output features j, j + 32, ... for its rows (weights are read once per 8 rows,
// coalesced, through L1; the rows' activations sit row-minor in shared memory, S[slot][row], so the 8 rows' value of
// one input is two broadcast LDS.128: 4 loads per 16 FFMA), fp32 FFMA in the reference's summation order over the
// inputs -- so the results are torch's to rounding (parity tests: 2e-5).  First version (4 rows per warp, row-major
// scratch: 6 loads per 8 FFMA): 0.57 ms per 16,384 rows of the virtual-sensor program (ncu launch list of C2).
#include "kernels.cuh"

namespace mmf {

constexpr int RM_ROWS = 8;       // rows per warp
constexpr int RM_WARPS = 16;     // most warps per CTA (fewer when the program's scratch is large)
constexpr int RM_MAX_OUT = 256;  // widest layer output

struct RowMlpParams {
  mmf_mlp_op ops[MMF_MLP_MAX_OPS];
  int n_ops, n_inputs, n_outputs, scratch;
  long long rows;
  const float* weights;
  const float* inputs[MMF_MLP_MAX_IO];
  int in_dims[MMF_MLP_MAX_IO], in_slots[MMF_MLP_MAX_IO];
  float* outputs[MMF_MLP_MAX_IO];
  int out_dims[MMF_MLP_MAX_IO], out_slots[MMF_MLP_MAX_IO];
};

constexpr int RM_KC = 64;                      // input rows of a staged weight tile
constexpr int RM_TILE = RM_KC * 64;            // floats of a weight tile: RM_KC inputs x 64 output features

__device__ __forceinline__ void rm_cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// Weight tile (k0 .. k0 + RM_KC) x (j0 .. j0 + 64) of an op -> shared memory, by all threads of the CTA.
__device__ __forceinline__ void rm_stage(float* buf, const float* Wt, int in_dim, int out_dim, int k0, int j0) {
  const int cols = min(64, out_dim - j0), rows = min(RM_KC, in_dim - k0);
  for (int e = threadIdx.x; e < rows * 64; e += blockDim.x) {
    const int kk = e >> 6, c = e & 63;
    if (c < cols) rm_cp_async4(buf + e, Wt + (size_t)(k0 + kk) * out_dim + j0 + c);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// The CTA's warps run the program in lock step: a weight tile is fetched from L2 ONCE per CTA (cp.async, double
// buffered) and read by every warp from shared memory.  (First version: every warp streamed the program's ~300 KB of
// weights through L1 for its own 4 rows -- 1.4 GB of L2 traffic per launch, 0.57 ms per 16,384 rows.)
__global__ void __launch_bounds__(RM_WARPS * 32) k_row_mlp(const __grid_constant__ RowMlpParams P) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  float* wbuf = sm;                                                        // [2][RM_TILE]
  float* S = sm + 2 * RM_TILE + (size_t)warp * RM_ROWS * P.scratch;        // S[slot * RM_ROWS + r]: a slot's rows are contiguous
  const long long cta_rows = (long long)warps * RM_ROWS;
  const long long passes = (P.rows + cta_rows - 1) / cta_rows;
  for (long long pass = blockIdx.x; pass < passes; pass += gridDim.x) {  // CTA-uniform trip count (barriers inside)
    const long long row0 = pass * cta_rows + (long long)warp * RM_ROWS;
    __syncwarp();
    // ---- inputs -> scratch (rows past the end read row P.rows - 1: computed, never written) -------------------------
    for (int i = 0; i < P.n_inputs; ++i) {
      const int d = P.in_dims[i];
      for (int e = lane; e < RM_ROWS * d; e += 32) {
        const int r = e / d, k = e - r * d;
        long long row = row0 + r;
        row = row < P.rows ? row : P.rows - 1;
        S[(P.in_slots[i] + k) * RM_ROWS + r] = __ldg(P.inputs[i] + row * d + k);
      }
    }
    __syncwarp();
    // ---- the program -------------------------------------------------------------------------------------------------
    for (int o = 0; o < P.n_ops; ++o) {
      const mmf_mlp_op op = P.ops[o];
      const float* Wt = P.weights + op.w_off;              // [in_dim][out_dim], input-major
      const float* b = Wt + (size_t)op.in_dim * op.out_dim;  // [out_dim]
      const int kchunks = (op.in_dim + RM_KC - 1) / RM_KC;
      const int tiles = ((op.out_dim + 63) / 64) * kchunks;
      __syncthreads();  // every warp is done with the previous op's last tile
      rm_stage(wbuf, Wt, op.in_dim, op.out_dim, 0, 0);
      float acc_a[RM_ROWS], acc_b[RM_ROWS];
      for (int t = 0; t < tiles; ++t) {
        const int j0 = (t / kchunks) * 64, k0 = (t % kchunks) * RM_KC;
        if (t + 1 < tiles) {
          rm_stage(wbuf + ((t + 1) & 1) * RM_TILE, Wt, op.in_dim, op.out_dim, ((t + 1) % kchunks) * RM_KC, ((t + 1) / kchunks) * 64);
          asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();  // tile t has landed for everybody
        const float* w = wbuf + (t & 1) * RM_TILE;
        const int ja = j0 + lane, jb = j0 + 32 + lane;
        const bool va = ja < op.out_dim, vb = jb < op.out_dim;
        if (k0 == 0) {
          const float ba = va ? __ldg(b + ja) : 0.0f, bb = vb ? __ldg(b + jb) : 0.0f;
#pragma unroll
          for (int r = 0; r < RM_ROWS; ++r) {
            acc_a[r] = ba;
            acc_b[r] = bb;
          }
        }
        const float4* src4 = reinterpret_cast<const float4*>(S + (size_t)(op.src + k0) * RM_ROWS);
        const int kn = min(RM_KC, op.in_dim - k0);
#pragma unroll 4
        for (int k = 0; k < kn; ++k) {
          const float wa = va ? w[k * 64 + lane] : 0.0f;
          const float wb = vb ? w[k * 64 + 32 + lane] : 0.0f;
          const float4 x0 = src4[2 * k], x1 = src4[2 * k + 1];  // the 8 rows' input k: two broadcast loads
          const float x[RM_ROWS] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int r = 0; r < RM_ROWS; ++r) {
            acc_a[r] = fmaf(wa, x[r], acc_a[r]);
            acc_b[r] = fmaf(wb, x[r], acc_b[r]);
          }
        }
        if (k0 + RM_KC >= op.in_dim) {
          // residual, activation, store: the 8 rows of a feature are 32 contiguous bytes -> vector accesses, conflict-free
          float4* dst_a = reinterpret_cast<float4*>(S + (size_t)(op.dst + ja) * RM_ROWS);
          float4* dst_b = reinterpret_cast<float4*>(S + (size_t)(op.dst + jb) * RM_ROWS);
          if (op.res >= 0) {
            if (va) {
              const float4* r4 = reinterpret_cast<const float4*>(S + (size_t)(op.res + ja) * RM_ROWS);
              const float4 r0 = r4[0], r1 = r4[1];
              acc_a[0] += r0.x; acc_a[1] += r0.y; acc_a[2] += r0.z; acc_a[3] += r0.w;
              acc_a[4] += r1.x; acc_a[5] += r1.y; acc_a[6] += r1.z; acc_a[7] += r1.w;
            }
            if (vb) {
              const float4* r4 = reinterpret_cast<const float4*>(S + (size_t)(op.res + jb) * RM_ROWS);
              const float4 r0 = r4[0], r1 = r4[1];
              acc_b[0] += r0.x; acc_b[1] += r0.y; acc_b[2] += r0.z; acc_b[3] += r0.w;
              acc_b[4] += r1.x; acc_b[5] += r1.y; acc_b[6] += r1.z; acc_b[7] += r1.w;
            }
          }
#pragma unroll
          for (int r = 0; r < RM_ROWS; ++r) {
            if (op.act == MMF_MLP_RELU) {
              acc_a[r] = fmaxf(acc_a[r], 0.0f);
              acc_b[r] = fmaxf(acc_b[r], 0.0f);
            } else if (op.act == MMF_MLP_SIGMOID) {
              acc_a[r] = 1.0f / (1.0f + expf(-acc_a[r]));
              acc_b[r] = 1.0f / (1.0f + expf(-acc_b[r]));
            }
          }
          if (va) {  // dst never overlaps src / res (the host allocates the slots)
            dst_a[0] = make_float4(acc_a[0], acc_a[1], acc_a[2], acc_a[3]);
            dst_a[1] = make_float4(acc_a[4], acc_a[5], acc_a[6], acc_a[7]);
          }
          if (vb) {
            dst_b[0] = make_float4(acc_b[0], acc_b[1], acc_b[2], acc_b[3]);
            dst_b[1] = make_float4(acc_b[4], acc_b[5], acc_b[6], acc_b[7]);
          }
        }
        __syncthreads();  // tile t's buffer may be refilled (tile t + 2)
      }
    }
    __syncwarp();
    // ---- outputs -----------------------------------------------------------------------------------------------------
    for (int i = 0; i < P.n_outputs; ++i) {
      const int d = P.out_dims[i];
      for (int e = lane; e < RM_ROWS * d; e += 32) {
        const int r = e / d, k = e - r * d;
        if (row0 + r < P.rows) P.outputs[i][(row0 + r) * d + k] = S[(P.out_slots[i] + k) * RM_ROWS + r];
      }
    }
  }
}

int launch_row_mlp(long long rows, const mmf_mlp_op* ops, int n_ops, const float* weights, const float* const* inputs,
                   const int32_t* in_dims, const int32_t* in_slots, int n_inputs, float* const* outputs,
                   const int32_t* out_dims, const int32_t* out_slots, int n_outputs, int scratch, cudaStream_t stream) {
  if (rows == 0) return MMF_OK;
  MMF_REQUIRE(n_ops >= 1 && n_ops <= MMF_MLP_MAX_OPS, "row_mlp: %d ops outside 1..%d", n_ops, MMF_MLP_MAX_OPS);
  MMF_REQUIRE(n_inputs >= 1 && n_inputs <= MMF_MLP_MAX_IO && n_outputs >= 1 && n_outputs <= MMF_MLP_MAX_IO,
              "row_mlp: %d inputs / %d outputs outside 1..%d", n_inputs, n_outputs, MMF_MLP_MAX_IO);
  MMF_REQUIRE(scratch >= 1 && scratch <= 1536, "row_mlp: scratch of %d floats per row outside 1..1536", scratch);
  MMF_REQUIRE(weights != nullptr && ops != nullptr, "row_mlp: NULL program or weights");
  RowMlpParams P;
  P.n_ops = n_ops; P.n_inputs = n_inputs; P.n_outputs = n_outputs; P.scratch = scratch; P.rows = rows;
  P.weights = weights;
  auto inside = [&](int slot, int dim) { return slot >= 0 && dim >= 1 && slot + dim <= scratch; };
  for (int o = 0; o < n_ops; ++o) {
    const mmf_mlp_op& op = ops[o];
    MMF_REQUIRE(op.in_dim >= 1 && op.out_dim >= 1 && op.out_dim <= RM_MAX_OUT && op.w_off >= 0,
                "row_mlp: op %d has bad dimensions (%d -> %d)", o, op.in_dim, op.out_dim);
    MMF_REQUIRE(inside(op.src, op.in_dim) && inside(op.dst, op.out_dim) && (op.res < 0 || inside(op.res, op.out_dim)),
                "row_mlp: op %d addresses scratch outside [0, %d)", o, scratch);
    MMF_REQUIRE(op.dst + op.out_dim <= op.src || op.src + op.in_dim <= op.dst, "row_mlp: op %d: dst overlaps src", o);
    MMF_REQUIRE(op.res < 0 || op.dst + op.out_dim <= op.res || op.res + op.out_dim <= op.dst,
                "row_mlp: op %d: dst overlaps the residual", o);
    MMF_REQUIRE(op.act >= MMF_MLP_NONE && op.act <= MMF_MLP_SIGMOID, "row_mlp: op %d: unknown activation %d", o, op.act);
    P.ops[o] = op;
  }
  for (int i = 0; i < n_inputs; ++i) {
    MMF_REQUIRE(inputs[i] != nullptr && inside(in_slots[i], in_dims[i]), "row_mlp: input %d is NULL or outside the scratch", i);
    P.inputs[i] = inputs[i]; P.in_dims[i] = in_dims[i]; P.in_slots[i] = in_slots[i];
  }
  for (int i = 0; i < n_outputs; ++i) {
    MMF_REQUIRE(outputs[i] != nullptr && inside(out_slots[i], out_dims[i]), "row_mlp: output %d is NULL or outside the scratch", i);
    P.outputs[i] = outputs[i]; P.out_dims[i] = out_dims[i]; P.out_slots[i] = out_slots[i];
  }
  const size_t wtiles = 2 * (size_t)RM_TILE * sizeof(float);
  const size_t per_warp = (size_t)RM_ROWS * scratch * sizeof(float);
  static thread_local int configured_dev = -1;
  static thread_local size_t window = 0;
  int dev = 0, sms = 148;
  MMF_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    int rc = opt_in_shared_memory(k_row_mlp, &window);
    if (rc) return rc;
    configured_dev = dev;
  }
  // The per-row scratch lives in shared memory, which therefore decides how many warps an SM holds: take the better of
  // one CTA per SM (whole window) and two CTAs per SM (half of it each); every CTA carries its own two weight tiles.
  auto fit = [&](size_t budget) {
    const long long w = budget > wtiles + 1024 ? (long long)((budget - wtiles - 1024) / per_warp) : 0;
    return (int)(w > RM_WARPS ? RM_WARPS : w);
  };
  const int w1 = fit(window), w2 = fit(window / 2);
  const int per_sm = 2 * w2 >= w1 ? 2 : 1;
  const int warps = per_sm == 2 ? w2 : w1;
  MMF_REQUIRE(warps >= 1, "row_mlp: a scratch of %d floats per row does not fit shared memory (window %zu B)", scratch, window);
  const size_t smem = wtiles + (size_t)warps * per_warp;
  MMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long cta_rows = (long long)warps * RM_ROWS;
  long long grid = (rows + cta_rows - 1) / cta_rows;
  const long long cap = (long long)sms * per_sm;
  if (grid > cap) grid = cap;
  k_row_mlp<<<(unsigned)grid, warps * 32, smem, stream>>>(P);
  MMF_LAUNCH_CHECK("k_row_mlp");
  return MMF_OK;
}

}  // namespace mmf
