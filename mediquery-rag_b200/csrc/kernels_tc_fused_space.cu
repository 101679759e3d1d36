// kernels_tc_fused_space.cu -- the single-launch tensor pipeline instantiated for Chroma's l2 / ip spaces over raw (non-unit)
// vectors (tc_fused_device.cuh with XF = true: the epilogue maps a tile's approximate cosines to ranking scores with the
// stored row norms, the tail rescoring uses common.cuh:rank_score).  A translation unit of its own so that it builds in
// parallel with the cosine instantiations of kernels_tc_fused.cu.
#include "tc_fused_device.cuh"

namespace mqvs {
namespace tc {

template cudaError_t launch_fused_any<true>(int, int, int, const CUtensorMap&, const CUtensorMap&, const FusedArgs&, int,
                                            cudaStream_t);

}  // namespace tc
}  // namespace mqvs
