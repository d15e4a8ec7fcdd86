// step_kernel instantiations for NF = 0 first-order directions, general problems (see pinn_variants.inc)
#define PINN_VARIANT_NF 0
#define PINN_VARIANT_GEN 1
#include "pinn_variants.inc"
