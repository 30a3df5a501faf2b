// DDPG actor forward (RL/MR_ddpg.py:124-149): obs(5) -> FC64 -> BN -> ReLU -> FC64 -> BN -> ReLU
// -> FC2 tanh -> * action_bound.  fp32 like the TensorFlow reference.  Weights live in shared
// memory (20.3 KB); every thread evaluates the MLP of its own env, so all weight reads are
// shared-memory broadcasts.
//
// Packed parameter layout (float32, input-major weight matrices W[in][out]):
//   w1[5][64] b1[64] gamma1[64] beta1[64] mean1[64] var1[64]
//   w2[64][64] b2[64] gamma2[64] beta2[64] mean2[64] var2[64]
//   w3[64][2] b3[2]
#pragma once

namespace mr {

constexpr int kActorIn = 5, kActorHidden = 64, kActorOut = 2;
constexpr int kOffW1 = 0;
constexpr int kOffB1 = kOffW1 + kActorIn * kActorHidden;
constexpr int kOffG1 = kOffB1 + kActorHidden;
constexpr int kOffBe1 = kOffG1 + kActorHidden;
constexpr int kOffM1 = kOffBe1 + kActorHidden;
constexpr int kOffV1 = kOffM1 + kActorHidden;
constexpr int kOffW2 = kOffV1 + kActorHidden;
constexpr int kOffB2 = kOffW2 + kActorHidden * kActorHidden;
constexpr int kOffG2 = kOffB2 + kActorHidden;
constexpr int kOffBe2 = kOffG2 + kActorHidden;
constexpr int kOffM2 = kOffBe2 + kActorHidden;
constexpr int kOffV2 = kOffM2 + kActorHidden;
constexpr int kOffW3 = kOffV2 + kActorHidden;
constexpr int kOffB3 = kOffW3 + kActorHidden * kActorOut;
constexpr int kActorParams = kOffB3 + kActorOut;          // 5186 floats
constexpr float kBnEps = 1e-5f;                           // tflearn batch_normalization epsilon

#if defined(__CUDACC__)
__device__ __forceinline__ float bn_relu(float x, const float* w, int off_g, int o) {
    // tflearn inference form: gamma * (x - moving_mean) / sqrt(moving_var + eps) + beta
    const float g = w[off_g + o], be = w[off_g + kActorHidden + o];
    const float m = w[off_g + 2 * kActorHidden + o], v = w[off_g + 3 * kActorHidden + o];
    const float y = g * (x - m) / sqrtf(v + kBnEps) + be;
    return y > 0.f ? y : 0.f;
}

__device__ __forceinline__ void actor_forward_smem(const float* __restrict__ w, const float obs[5], float hi0, float hi1,
                                                   float act[2]) {
    float h1[kActorHidden];
#pragma unroll
    for (int o = 0; o < kActorHidden; ++o) h1[o] = w[kOffB1 + o];
#pragma unroll
    for (int i = 0; i < kActorIn; ++i) {
        const float x = obs[i];
#pragma unroll
        for (int o4 = 0; o4 < kActorHidden / 4; ++o4) {
            const float4 ww = *reinterpret_cast<const float4*>(w + kOffW1 + i * kActorHidden + o4 * 4);
            h1[o4 * 4 + 0] = fmaf(x, ww.x, h1[o4 * 4 + 0]); h1[o4 * 4 + 1] = fmaf(x, ww.y, h1[o4 * 4 + 1]);
            h1[o4 * 4 + 2] = fmaf(x, ww.z, h1[o4 * 4 + 2]); h1[o4 * 4 + 3] = fmaf(x, ww.w, h1[o4 * 4 + 3]);
        }
    }
#pragma unroll
    for (int o = 0; o < kActorHidden; ++o) h1[o] = bn_relu(h1[o], w, kOffG1, o);

    float h2[kActorHidden];
#pragma unroll
    for (int o = 0; o < kActorHidden; ++o) h2[o] = w[kOffB2 + o];
#pragma unroll
    for (int i = 0; i < kActorHidden; ++i) {
        const float x = h1[i];
#pragma unroll
        for (int o4 = 0; o4 < kActorHidden / 4; ++o4) {
            const float4 ww = *reinterpret_cast<const float4*>(w + kOffW2 + i * kActorHidden + o4 * 4);
            h2[o4 * 4 + 0] = fmaf(x, ww.x, h2[o4 * 4 + 0]); h2[o4 * 4 + 1] = fmaf(x, ww.y, h2[o4 * 4 + 1]);
            h2[o4 * 4 + 2] = fmaf(x, ww.z, h2[o4 * 4 + 2]); h2[o4 * 4 + 3] = fmaf(x, ww.w, h2[o4 * 4 + 3]);
        }
    }
    float o0 = w[kOffB3], o1 = w[kOffB3 + 1];
#pragma unroll
    for (int i = 0; i < kActorHidden; ++i) {
        const float x = bn_relu(h2[i], w, kOffG2, i);
        const float2 ww = *reinterpret_cast<const float2*>(w + kOffW3 + i * kActorOut);
        o0 = fmaf(x, ww.x, o0); o1 = fmaf(x, ww.y, o1);
    }
    act[0] = tanhf(o0) * hi0;   // scaled_out = out * action_bound (RL/MR_ddpg.py:137)
    act[1] = tanhf(o1) * hi1;
}
#endif

}  // namespace mr
