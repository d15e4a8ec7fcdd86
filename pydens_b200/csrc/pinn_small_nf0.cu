// small_step_kernel (tiny-batch fit loop, work split over (point, unit) pairs) instantiations for NF = 0
#include "pinn_small_kernel.cuh"

pinn::MultiKernelFn pinn_small_variant_nf0(int ns) {
    using namespace pinn::small;
    switch (ns) {
        case 0: return small_step_kernel<0, 0>;

        default: return nullptr;
    }
}
