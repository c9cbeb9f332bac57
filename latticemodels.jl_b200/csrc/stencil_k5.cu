// Stencil kernels of compiled pattern 5 (see stencil.cu for the pattern table).
#define LM_ST_ID 5
#include "stencil_unit.inc"
