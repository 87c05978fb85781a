// tcgen05 / TMEM / TMA implicit-GEMM path (TIC_COMPUTE_TENSOR_*).  Placeholder interface until
// the kernels land: nothing is "supported", so every layer runs on the fp32 CUDA-core kernels.
#pragma once
#include <string>
#include "tic_common.cuh"

namespace tic {
struct UmmaWeights {
  void release() {}
};
inline bool umma_supported(const LayerArgs&, int, int) { return false; }
inline int launch_umma(cudaStream_t, const LayerArgs&, int, int, const float*, UmmaWeights*, bool, int, std::string*) {
  return -5;
}
}  // namespace tic
