// Stencil kernels of compiled pattern 0 (see stencil.cu for the pattern table).
#define LM_ST_ID 0
#include "stencil_unit.inc"
