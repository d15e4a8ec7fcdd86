// step_kernel instantiations for NF = 3 first-order directions, general problems, NS = 0..1 (see pinn_variants.inc)
#define PINN_VARIANT_NF 3
#define PINN_VARIANT_GEN 1
#define PINN_VARIANT_NS_HI 1
#define PINN_VARIANT_NEXT pinn_variants_gen_nf3_hi
#include "pinn_variants.inc"
