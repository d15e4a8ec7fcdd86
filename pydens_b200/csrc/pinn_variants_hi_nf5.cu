// step_kernel instantiation for NF = NS = 5 derivative directions (see pinn_variants_hi_nf6.cu)
#define PINN_VARIANT_HI_NF 5
#include "pinn_variants_hi.inc"
