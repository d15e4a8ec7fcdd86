// small_step_kernel (tiny-batch fit loop, work split over (point, unit) pairs) instantiations for NF = 2
#include "pinn_small_kernel.cuh"

pinn::MultiKernelFn pinn_small_variant_nf2(int ns) {
    using namespace pinn::small;
    switch (ns) {
        case 0: return small_step_kernel<2, 0>;
        case 1: return small_step_kernel<2, 1>;
        case 2: return small_step_kernel<2, 2>;
        default: return nullptr;
    }
}
