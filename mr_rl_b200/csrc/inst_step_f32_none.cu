// Instantiation unit: storage float, noise MR_NOISE_NONE — step + reset kernels.
#define MR_T float
#define MR_MODE MR_NOISE_NONE
#include "mr_step.inl"
