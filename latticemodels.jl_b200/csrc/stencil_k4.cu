// Stencil kernels of compiled pattern 4 (see stencil.cu for the pattern table).
#define LM_ST_ID 4
#include "stencil_unit.inc"
