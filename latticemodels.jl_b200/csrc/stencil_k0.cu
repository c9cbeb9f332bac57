// Stencil kernels of compiled pattern 0 (see stencil.cu for the pattern table).
#include "stencil_inst.cuh"

namespace lm {
int stencil_launch_0(int variant, bool c64, int mode, const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s) {
#if defined(LM_STENCIL_EXPLORE) && (0 == 1 || 0 == 2)
    return -1;      // exploration builds skip the catch-all / honeycomb-NN patterns
#else
    return launch_var<1, LM_ST_MASK0>(variant, c64, mode, a, tmx, grid, s);
#endif
}
int stencil_observe_0(bool c64, const StencilObsArgs& a, const CUtensorMap& tmx, unsigned grid, cudaStream_t s) {
    return launch_obs<1, LM_ST_MASK0>(c64, a, tmx, grid, s);
}
void stencil_obs_shape_0(int* P1, int* P2, int* nf) { obs_shape<1, LM_ST_MASK0>(P1, P2, nf); }
}  // namespace lm
