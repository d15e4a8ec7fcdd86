// step_kernel instantiations for NF = 4 first-order directions, general problems, NS = 0..2 (see pinn_variants.inc)
#define PINN_VARIANT_NF 4
#define PINN_VARIANT_GEN 1
#define PINN_VARIANT_NS_HI 2
#define PINN_VARIANT_NEXT pinn_variants_gen_nf4_hi
#include "pinn_variants.inc"
