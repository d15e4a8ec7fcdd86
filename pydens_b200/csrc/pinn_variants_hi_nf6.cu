// step_kernel instantiation for NF = NS = 6 derivative directions (the ABI's PINN_MAX_DIRS): full Hessians in three
// dimensions, Laplacians in five / six.  One kernel only — the general form with the per-point state in global memory
// (13 jet channels per unit do not fit shared memory for any network worth the name); the tracer promotes every
// direction of such a problem to second order, so NS = NF is the only combination that occurs.
#define PINN_VARIANT_HI_NF 6
#include "pinn_variants_hi.inc"
