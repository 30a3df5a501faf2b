// Instantiation unit: storage double, noise MR_NOISE_NONE — fused rollout kernel.
#define MR_T double
#define MR_MODE MR_NOISE_NONE
#include "mr_rollout.inl"
