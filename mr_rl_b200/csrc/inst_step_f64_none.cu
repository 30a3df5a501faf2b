// Instantiation unit: storage double, noise MR_NOISE_NONE — step + reset kernels.
#define MR_T double
#define MR_MODE MR_NOISE_NONE
#include "mr_step.inl"
