// Instantiation unit: storage double, noise MR_NOISE_PHILOX — step + reset kernels.
#define MR_T double
#define MR_MODE MR_NOISE_PHILOX
#include "mr_step.inl"
