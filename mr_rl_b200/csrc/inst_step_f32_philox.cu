// Instantiation unit: storage float, noise MR_NOISE_PHILOX — step + reset kernels.
#define MR_T float
#define MR_MODE MR_NOISE_PHILOX
#include "mr_step.inl"
