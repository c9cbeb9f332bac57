// Stencil kernels of compiled pattern 2 (see stencil.cu for the pattern table).
#define LM_ST_ID 2
#include "stencil_unit.inc"
