// Stencil kernels of compiled pattern 9 (see stencil.cu for the pattern table).
#define LM_ST_ID 9
#include "stencil_unit.inc"
