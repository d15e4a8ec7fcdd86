// step_kernel instantiations for NF = 3 first-order directions, general problems, NS = 2..3 (see pinn_variants.inc)
#define PINN_VARIANT_NF 3
#define PINN_VARIANT_GEN 1
#define PINN_VARIANT_NS_LO 2
#define PINN_VARIANT_NAME pinn_variants_gen_nf3_hi
#include "pinn_variants.inc"
