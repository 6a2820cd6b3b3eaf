// particle_chain_tc.cu -- tcgen05 / TMEM build of the per-particle MLP chain (placeholder until
// the tensor-core path lands; the FP32 CUDA-core chain in particle_chain_ffma.cu is the parity build).
#include "kernels.cuh"

namespace mmf {

size_t chain_mma_bytes(const mmf_chain* chain) {
  (void)chain;
  return 0;
}

int pack_chain_mma(const mmf_chain* chain, void* dst, cudaStream_t stream) {
  (void)chain; (void)dst; (void)stream;
  set_error("tcgen05 operand packing is not built yet");
  return MMF_E_UNSUPPORTED;
}

int launch_particle_chain_tc(const mmf_pf_model*, int, int, const float*, const float*, const float*, const float*,
                             const float*, uint32_t, int, float*, float*, float*, cudaStream_t) {
  set_error("the tcgen05 particle chain is not built yet; use MMF_PREC_FP32");
  return MMF_E_UNSUPPORTED;
}

}  // namespace mmf
