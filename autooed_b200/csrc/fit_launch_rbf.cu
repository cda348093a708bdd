#include "fit_launch_impl.cuh"

AOED_FIT_LAUNCH_FAMILY(rbf, -1)
