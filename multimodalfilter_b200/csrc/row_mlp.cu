// row_mlp.cu -- the per-trajectory MLPs of the path as ONE kernel per module (R9):
// position / sensor observation encoders (ref: crossmodal/push_models/layers.py:107-136), the particle-filter and
// Kalman-filter crossmodal weight models (ref: crossmodal/push_models/crossmodal_pf.py:72-104,
// crossmodal/door_models/crossmodal_kf.py:134-167) and the virtual-sensor head (ref: crossmodal/door_models/kf.py:81-126).
//
// These are stacks of Linear / ReLU / Sigmoid / residual blocks, 3 ... 192 inputs wide, evaluated once per (step,
// trajectory) row: 10-90 k MAC per row, all weights L2 resident.  As torch modules each stack is 10-40 library
// launches (cuBLAS SIMT sgemm + ATen elementwise) per call; here the host compiles a stack into a short program of
// fused ops   dst = act(W src + b [+ residual])   over per-row scratch slots and one launch runs the whole program:
// a warp owns R = 4 rows, lane j computes output features j, j + 32, ... for its rows (weights are read once per 4 rows,
// coalesced, through L1; the rows' activations are shared-memory broadcasts), fp32 FFMA in the reference's summation
// order over the inputs -- so the results are torch's to rounding (parity tests: 1e-5).
#include "kernels.cuh"

namespace mmf {

constexpr int RM_ROWS = 4;       // rows per warp
constexpr int RM_WARPS = 8;      // warps per CTA
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

__global__ void __launch_bounds__(RM_WARPS * 32) k_row_mlp(const __grid_constant__ RowMlpParams P) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* S = sm + (size_t)warp * RM_ROWS * P.scratch;  // S[r * scratch + slot]
  const long long groups = (P.rows + RM_ROWS - 1) / RM_ROWS;
  for (long long grp = (long long)blockIdx.x * RM_WARPS + warp; grp < groups; grp += (long long)gridDim.x * RM_WARPS) {
    const long long row0 = grp * RM_ROWS;
    __syncwarp();
    // ---- inputs -> scratch (rows past the end read row P.rows - 1: computed, never written) -------------------------
    for (int i = 0; i < P.n_inputs; ++i) {
      const int d = P.in_dims[i];
      for (int e = lane; e < RM_ROWS * d; e += 32) {
        const int r = e / d, k = e - r * d;
        long long row = row0 + r;
        row = row < P.rows ? row : P.rows - 1;
        S[r * P.scratch + P.in_slots[i] + k] = __ldg(P.inputs[i] + row * d + k);
      }
    }
    __syncwarp();
    // ---- the program -------------------------------------------------------------------------------------------------
    for (int o = 0; o < P.n_ops; ++o) {
      const mmf_mlp_op op = P.ops[o];
      const float* Wt = P.weights + op.w_off;              // [in_dim][out_dim], input-major
      const float* b = Wt + (size_t)op.in_dim * op.out_dim;  // [out_dim]
      for (int j0 = 0; j0 < op.out_dim; j0 += 64) {        // two output features per lane per pass
        const int ja = j0 + lane, jb = j0 + 32 + lane;
        const bool va = ja < op.out_dim, vb = jb < op.out_dim;
        float acc_a[RM_ROWS], acc_b[RM_ROWS];
        const float ba = va ? __ldg(b + ja) : 0.0f, bb = vb ? __ldg(b + jb) : 0.0f;
#pragma unroll
        for (int r = 0; r < RM_ROWS; ++r) {
          acc_a[r] = ba;
          acc_b[r] = bb;
        }
        const float* src = S + op.src;
#pragma unroll 4
        for (int k = 0; k < op.in_dim; ++k) {
          const float wa = va ? __ldg(Wt + (size_t)k * op.out_dim + ja) : 0.0f;
          const float wb = vb ? __ldg(Wt + (size_t)k * op.out_dim + jb) : 0.0f;
#pragma unroll
          for (int r = 0; r < RM_ROWS; ++r) {
            const float x = src[r * P.scratch + k];  // broadcast
            acc_a[r] = fmaf(wa, x, acc_a[r]);
            acc_b[r] = fmaf(wb, x, acc_b[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < RM_ROWS; ++r) {
          float ya = acc_a[r], yb = acc_b[r];
          if (op.res >= 0) {
            if (va) ya += S[r * P.scratch + op.res + ja];
            if (vb) yb += S[r * P.scratch + op.res + jb];
          }
          if (op.act == MMF_MLP_RELU) {
            ya = fmaxf(ya, 0.0f);
            yb = fmaxf(yb, 0.0f);
          } else if (op.act == MMF_MLP_SIGMOID) {
            ya = 1.0f / (1.0f + expf(-ya));
            yb = 1.0f / (1.0f + expf(-yb));
          }
          if (va) S[r * P.scratch + op.dst + ja] = ya;  // dst never overlaps src / res (the host allocates the slots)
          if (vb) S[r * P.scratch + op.dst + jb] = yb;
        }
      }
      __syncwarp();
    }
    // ---- outputs -----------------------------------------------------------------------------------------------------
    for (int i = 0; i < P.n_outputs; ++i) {
      const int d = P.out_dims[i];
      for (int e = lane; e < RM_ROWS * d; e += 32) {
        const int r = e / d, k = e - r * d;
        if (row0 + r < P.rows) P.outputs[i][(row0 + r) * d + k] = S[r * P.scratch + P.out_slots[i] + k];
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
  const size_t smem = (size_t)RM_WARPS * RM_ROWS * scratch * sizeof(float);
  static thread_local int configured_dev = -1;
  static thread_local size_t window = 0;
  int dev = 0, sms = 148;
  MMF_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    int rc = opt_in_shared_memory(k_row_mlp, &window);
    if (rc) return rc;
    configured_dev = dev;
  }
  MMF_REQUIRE(smem <= window, "row_mlp needs %zu B of shared memory (window %zu B)", smem, window);
  MMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long groups = (rows + RM_ROWS - 1) / RM_ROWS;
  long long grid = (groups + RM_WARPS - 1) / RM_WARPS;
  const long long cap = (long long)sms * 4;
  if (grid > cap) grid = cap;
  k_row_mlp<<<(unsigned)grid, RM_WARPS * 32, smem, stream>>>(P);
  MMF_LAUNCH_CHECK("k_row_mlp");
  return MMF_OK;
}

}  // namespace mmf
