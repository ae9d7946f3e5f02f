// Part 1 of the generic kernels' instantiations (see the end of riccati_generic.cu): compiled in parallel with the other parts.
#define O2C_GENERIC_PART 1
#include "riccati_generic.cu"
