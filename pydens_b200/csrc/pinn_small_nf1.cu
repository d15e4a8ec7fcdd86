// small_step_kernel (tiny-batch fit loop, work split over (point, unit) pairs) instantiations for NF = 1
#include "pinn_small_kernel.cuh"

pinn::MultiKernelFn pinn_small_variant_nf1(int ns) {
    using namespace pinn::small;
    switch (ns) {
        case 0: return small_step_kernel<1, 0>;
        case 1: return small_step_kernel<1, 1>;
        default: return nullptr;
    }
}
