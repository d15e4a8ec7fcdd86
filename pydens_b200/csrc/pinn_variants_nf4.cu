// step_kernel instantiations for NF = 4 first-order directions (see pinn_variants.inc)
#define PINN_VARIANT_NF 4
#include "pinn_variants.inc"
