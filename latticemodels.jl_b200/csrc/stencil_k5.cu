// Stencil kernels of compiled pattern 5 (see stencil.cu for the pattern table).
#include "stencil_inst.cuh"

namespace lm {
int stencil_launch_5(int variant, bool c64, int mode, const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s) {
#if defined(LM_STENCIL_EXPLORE) && (5 == 1 || 5 == 2)
    return -1;      // exploration builds skip the catch-all / honeycomb-NN patterns
#else
    return launch_var<2, LM_ST_MASK5>(variant, c64, mode, a, tmx, grid, s);
#endif
}
int stencil_observe_5(bool c64, const StencilObsArgs& a, const CUtensorMap& tmx, unsigned grid, cudaStream_t s) {
    return launch_obs<2, LM_ST_MASK5>(c64, a, tmx, grid, s);
}
void stencil_obs_shape_5(int* P1, int* P2, int* nf) { obs_shape<2, LM_ST_MASK5>(P1, P2, nf); }
}  // namespace lm
