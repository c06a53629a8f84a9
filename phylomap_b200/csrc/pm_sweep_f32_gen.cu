// Sweep kernels, float production arithmetic, run-time state count (<= PM_NMAX).
#include "pm_launch_impl.cuh"
template struct pm::Sweep<float, 0, false>;
