// Instantiation unit: storage double, noise MR_NOISE_TABLE — step + reset kernels.
#define MR_T double
#define MR_MODE MR_NOISE_TABLE
#include "mr_step.inl"
