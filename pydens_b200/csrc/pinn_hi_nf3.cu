// hi_step_kernel instantiations for 3 derivative direction(s), orders 3 and 4 (see pinn_hi_kernel.cuh)
#include "pinn_hi_kernel.cuh"

pinn::StepKernelFn pinn_hi_variant_nf3(int order) {
    if (order == 3) return pinn::hi::hi_step_kernel<3, 3>;
    if (order == 4) return pinn::hi::hi_step_kernel<3, 4>;
    return nullptr;
}
