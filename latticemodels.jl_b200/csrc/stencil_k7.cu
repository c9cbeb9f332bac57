// Stencil kernels of compiled pattern 7 (see stencil.cu for the pattern table).
#define LM_ST_ID 7
#include "stencil_unit.inc"
