// Sweep kernels, double production arithmetic, run-time state count (<= PM_NMAX).
#include "pm_launch_impl.cuh"
template struct pm::Sweep<double, 0, false>;
