// Stencil kernels of compiled pattern 1 (see stencil.cu for the pattern table).
#define LM_ST_ID 1
#include "stencil_unit.inc"
