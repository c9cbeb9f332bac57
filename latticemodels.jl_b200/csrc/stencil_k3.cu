// Stencil kernels of compiled pattern 3 (see stencil.cu for the pattern table).
#define LM_ST_ID 3
#include "stencil_unit.inc"
