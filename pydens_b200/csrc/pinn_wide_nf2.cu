// wide_step_kernel (tcgen05 / TMEM tile kernel for wide networks) instantiations for NF = 2 first-order directions
#include "pinn_wide_kernel.cuh"

pinn::StepKernelFn pinn_wide_variant_nf2(int ns) {
    using namespace pinn::wide;
    switch (ns) {
        case 0: return wide_step_kernel<2, 0>;
        case 1: return wide_step_kernel<2, 1>;
        case 2: return wide_step_kernel<2, 2>;
        default: return nullptr;
    }
}
