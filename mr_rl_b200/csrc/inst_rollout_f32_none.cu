// Instantiation unit: storage float, noise MR_NOISE_NONE — fused rollout kernel.
#define MR_T float
#define MR_MODE MR_NOISE_NONE
#include "mr_rollout.inl"
