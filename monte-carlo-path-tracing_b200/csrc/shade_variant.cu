// shade_variant.cu — one instantiation pair (path / volpath) of k_shade.  Compiled once per BSDF model with
// -DB200PT_SHADE_ONLY=<model> -DB200PT_SHADE_NAME=<suffix> (see the Makefile), so the models build in parallel and each
// kernel only carries the code of its own model.
#include "shade_kernel.cuh"

#ifndef B200PT_SHADE_ONLY
#error "compile with -DB200PT_SHADE_ONLY=<bsdf type or -1> -DB200PT_SHADE_NAME=<suffix>"
#endif
#define B200PT_CAT2(a, b) a##b
#define B200PT_CAT(a, b) B200PT_CAT2(a, b)

namespace b200pt {

B200PT_DECLARE_SHADE_VARIANT(B200PT_CAT(LaunchShadeVariant_, B200PT_SHADE_NAME)) {
    if (vol)
        k_shade<true, B200PT_SHADE_ONLY><<<blocks, kShadeThreads, 0, stream>>>(scene, bp, depth, qin, which_in, qout, sq, radiance, counters, capacity, bin_list, bin);
    else
        k_shade<false, B200PT_SHADE_ONLY><<<blocks, kShadeThreads, 0, stream>>>(scene, bp, depth, qin, which_in, qout, sq, radiance, counters, capacity, bin_list, bin);
}

} // namespace b200pt
