// small_step_kernel (tiny-batch fit loop, work split over (point, unit) pairs) instantiations for NF = 3
#include "pinn_small_kernel.cuh"

pinn::MultiKernelFn pinn_small_variant_nf3(int ns) {
    using namespace pinn::small;
    switch (ns) {
        case 0: return small_step_kernel<3, 0>;
        case 1: return small_step_kernel<3, 1>;
        case 2: return small_step_kernel<3, 2>;
        case 3: return small_step_kernel<3, 3>;
        default: return nullptr;
    }
}
