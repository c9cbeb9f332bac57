// Stencil kernels of compiled pattern 6 (see stencil.cu for the pattern table).
#define LM_ST_ID 6
#include "stencil_unit.inc"
