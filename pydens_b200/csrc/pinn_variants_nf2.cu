// step_kernel instantiations for NF = 2 first-order directions (see pinn_variants.inc)
#define PINN_VARIANT_NF 2
#include "pinn_variants.inc"
