// Sweep kernels, double production arithmetic, 2 and 4 states.
#include "pm_launch_impl.cuh"
template struct pm::Sweep<double, 2, false>;
template struct pm::Sweep<double, 4, false>;
