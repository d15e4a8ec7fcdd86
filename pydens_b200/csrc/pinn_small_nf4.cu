// small_step_kernel (tiny-batch fit loop, work split over (point, unit) pairs) instantiations for NF = 4
#include "pinn_small_kernel.cuh"

pinn::MultiKernelFn pinn_small_variant_nf4(int ns) {
    using namespace pinn::small;
    switch (ns) {
        case 0: return small_step_kernel<4, 0>;
        case 1: return small_step_kernel<4, 1>;
        case 2: return small_step_kernel<4, 2>;
        case 3: return small_step_kernel<4, 3>;
        case 4: return small_step_kernel<4, 4>;
        default: return nullptr;
    }
}
