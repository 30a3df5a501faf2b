// Instantiation unit: storage float, noise MR_NOISE_TABLE — step + reset kernels.
#define MR_T float
#define MR_MODE MR_NOISE_TABLE
#include "mr_step.inl"
