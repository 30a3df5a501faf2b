// Instantiation unit: storage double, noise MR_NOISE_PHILOX — fused rollout kernel.
#define MR_T double
#define MR_MODE MR_NOISE_PHILOX
#include "mr_rollout.inl"
