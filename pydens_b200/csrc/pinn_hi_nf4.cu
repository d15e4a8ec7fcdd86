// hi_step_kernel instantiations for 4 derivative direction(s), orders 3 and 4 (see pinn_hi_kernel.cuh): two arguments
// and their two diagonals — the jet set of the biharmonic operator in two dimensions
#include "pinn_hi_kernel.cuh"

pinn::StepKernelFn pinn_hi_variant_nf4(int order) {
    if (order == 3) return pinn::hi::hi_step_kernel<4, 3>;
    if (order == 4) return pinn::hi::hi_step_kernel<4, 4>;
    return nullptr;
}
