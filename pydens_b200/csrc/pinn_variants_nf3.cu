// step_kernel instantiations for NF = 3 first-order directions (see pinn_variants.inc)
#define PINN_VARIANT_NF 3
#include "pinn_variants.inc"
