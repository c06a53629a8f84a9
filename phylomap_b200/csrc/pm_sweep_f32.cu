// Sweep kernels, float production arithmetic, 2 and 4 states.
#include "pm_launch_impl.cuh"
template struct pm::Sweep<float, 2, false>;
template struct pm::Sweep<float, 4, false>;
