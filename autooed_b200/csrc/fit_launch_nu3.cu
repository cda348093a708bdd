#include "fit_launch_impl.cuh"

AOED_FIT_LAUNCH_FAMILY(nu3, 3)
