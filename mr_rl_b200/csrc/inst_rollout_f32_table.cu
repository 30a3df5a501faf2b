// Instantiation unit: storage float, noise MR_NOISE_TABLE — fused rollout kernel.
#define MR_T float
#define MR_MODE MR_NOISE_TABLE
#include "mr_rollout.inl"
