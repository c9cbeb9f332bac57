// Stencil kernels of compiled pattern 2 (see stencil.cu for the pattern table).
#include "stencil_inst.cuh"

namespace lm {
int stencil_launch_2(int variant, bool c64, int mode, const StencilArgs& a, dim3 grid, cudaStream_t s) {
#if defined(LM_STENCIL_EXPLORE) && (2 == 1 || 2 == 2)
    return -1;      // exploration builds skip the catch-all / honeycomb-NN patterns
#else
    return launch_var<2, LM_ST_MASK2>(variant, c64, mode, a, grid, s);
#endif
}
}  // namespace lm
