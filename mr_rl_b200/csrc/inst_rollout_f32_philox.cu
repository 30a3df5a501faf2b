// Instantiation unit: storage float, noise MR_NOISE_PHILOX — fused rollout kernel.
#define MR_T float
#define MR_MODE MR_NOISE_PHILOX
#include "mr_rollout.inl"
