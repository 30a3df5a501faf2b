// Instantiation unit: storage double, noise MR_NOISE_TABLE — fused rollout kernel.
#define MR_T double
#define MR_MODE MR_NOISE_TABLE
#include "mr_rollout.inl"
