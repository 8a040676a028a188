// Kernel instantiations for 2-limb fields (compiled as its own translation unit so the four
// limb counts build in parallel).
#include "launch_impl.cuh"
template struct Launch<2>;
