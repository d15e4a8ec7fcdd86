// wide_step_kernel (tcgen05 / TMEM tile kernel for wide networks) instantiations for NF = 3 first-order directions
#include "pinn_wide_kernel.cuh"

pinn::StepKernelFn pinn_wide_variant_nf3(int ns, int threads) {
    using namespace pinn::wide;
    if (threads == 512) {
        switch (ns) {
            case 0: return wide_step_kernel<3, 0, 512>;
            case 1: return wide_step_kernel<3, 1, 512>;
            case 2: return wide_step_kernel<3, 2, 512>;
            case 3: return wide_step_kernel<3, 3, 512>;
            default: return nullptr;
        }
    }
    switch (ns) {
        case 0: return wide_step_kernel<3, 0, 256>;
        case 1: return wide_step_kernel<3, 1, 256>;
        case 2: return wide_step_kernel<3, 2, 256>;
        case 3: return wide_step_kernel<3, 3, 256>;
        default: return nullptr;
    }
}
