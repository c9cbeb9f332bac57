// Stencil kernels of compiled pattern 8 (see stencil.cu for the pattern table).
#define LM_ST_ID 8
#include "stencil_unit.inc"
