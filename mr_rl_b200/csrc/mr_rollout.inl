// Fused K-step rollout kernel.  Included by the per-(dtype, noise mode) instantiation units
// with MR_T and MR_MODE defined.
#include <cstdlib>

#include "mr_actor.cuh"
#include "mr_actor_tc.cuh"
#include "mr_actor_tc16.cuh"
#include "mr_common.cuh"

namespace mr {

// =============================================================================================
// Fused rollout: K env steps per launch, state in registers, one env per thread.
// FP64-pipe bound (no HBM traffic for state between steps).
// =============================================================================================
constexpr int kSrcActorTc = 100;     // internal: MR_ACTIONS_ACTOR with the hidden layer on the tensor cores, 3xTF32 (mr_actor_tc.cuh)
constexpr int kSrcActorTc16 = 101;   // internal: both dense layers on the tensor cores, 3xFP16, 4 CTAs per SM (mr_actor_tc16.cuh)

#ifndef MR_ROLLOUT_MINB
#define MR_ROLLOUT_MINB 7   // measured: 72 registers, 28 warps/SM.  Round 1: 6 CTAs (85 registers) 56 vs 46 Genv-steps/s (sigma = 0)
                            // against unconstrained (143 registers).  Round 2 (sigma = 1, K = 64; 6 / 7 / 8 CTAs per SM): 2^20 envs random
                            // actions 48.1 / 49.2 / 48.8, tensor actions 49.3 / 52.0 / 52.1 Genv-steps/s; 131072 envs (the per-rank
                            // batch of 8 GPUs: 1024 CTAs, one wave at 7 per SM instead of two) 235 / 192 / 195 us
#endif
// SMALL: a batch that does not fill the GPU (BASELINE configs[1]: 4096 envs = 32 CTAs) is bound by the latency of one
// env's dependent chain, not by occupancy: the register cap that buys 24 warps per SM only adds spill traffic to that
// chain.  Measured (4096 envs, K = 64, sigma = 1): 76 us with the cap, 60 us without.
template <int SRC, bool PERENV, bool SMALL> struct RolloutMinCtas {
    static constexpr int value = SRC == kSrcActorTc16 ? 4 : (SRC == MR_ACTIONS_ACTOR || SRC == kSrcActorTc || PERENV || SMALL) ? 1 : MR_ROLLOUT_MINB;
};
// PERENV: a0 / noise_var / is_mismatched come from the per-env rows of the state (the model flag is then a run-time
// branch and MISM is ignored).
template <class T, int MODE, bool MISM, int SRC, bool PERENV = false, bool SMALL = false>
__global__ void __launch_bounds__(128, RolloutMinCtas<SRC, PERENV, SMALL>::value)
env_rollout_kernel(StateView<T> st, RolloutView<T> io, OutView<T> out, NoiseView nv, TimeView tv, Params p, int64_t n) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ double s_stats[MR_STATS_LEN];
    float* s_actor = reinterpret_cast<float*>(s_dyn);
    if constexpr (SRC == MR_ACTIONS_ACTOR) {
        for (int k = threadIdx.x; k < kActorParams; k += blockDim.x) s_actor[k] = io.actor[k];
    }
    if constexpr (SRC == kSrcActorTc) actor_tc_setup(*reinterpret_cast<ActorTcSmem*>(s_dyn), io.actor);
    if constexpr (SRC == kSrcActorTc16) actor_tc16_setup(*reinterpret_cast<ActorTc16Smem*>(s_dyn), io.actor);
    if (threadIdx.x < MR_STATS_LEN) s_stats[threadIdx.x] = 0.0;
    __syncthreads();

    // per-thread statistics, kept narrow inside the step loop (the kernel is register-bound): the episode ends by reason
    // and the summed episode lengths are integers, the episode count is their sum and the env-step count is k_steps per
    // live env; they become the doubles of the statistics vector only after the loop
    // one test per step instead of eight when nothing is recorded (the throughput case)
    const bool recording = io.traj_xy || io.traj_sp || io.traj_done || io.traj_actions || io.traj_rew || io.traj_episode ||
                           io.traj_step || io.traj_reset_xy;
    double sum_rew = 0.0;
    int sum_len = 0, n_goal = 0, n_oob = 0, n_timeout = 0, n_live_steps = 0, n_failed = 0;
    const uint64_t off = step_offset(nv);
    const Params p_launch = p;
    int actor_calls = 0;                           // mbarrier phase of the tensor-core actor paths

    // One tile of blockDim.x envs per iteration.  The 3xFP16 actor kernel is persistent (grid = resident CTAs: its
    // per-CTA set-up — weights split into fp16 pairs in UMMA layout, TMEM allocation — is paid once per CTA instead of
    // once per 128 envs); every other variant is launched with one CTA per tile.
    const int64_t n_tiles = (n + blockDim.x - 1) / blockDim.x;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t i = tile * blockDim.x + threadIdx.x;
    const bool live = i < n;                       // every thread runs the loop (CTA-wide barriers in the actor path)
    if constexpr (PERENV) p = p_launch;
    Env e;
    e.x = 110.0; e.y = 110.0; e.fx = 0.0; e.fy = 0.0; e.h = p.dt; e.counter = 0;   // harmless state for padding lanes
    e.status = 0; e.spx = e.spy = 0.0;
    int32_t cur = 0;
    bool mism_i = MISM;
    if constexpr (PERENV) {
        if (live) { p.a0 = st.a0[i]; p.sigma = st.sigma[i]; mism_i = st.mism[i] != 0; }
    }
    if (live) {
        e.x = (double)st.x[i]; e.y = (double)st.y[i]; e.fx = (double)st.fx[i]; e.fy = (double)st.fy[i];
        e.counter = st.counter[i];
        const double t_in = time_at(tv, e.counter, p.dt);
        e.h = decode_h<T>(st.h[i], (t_in + p.dt) - t_in);
        if constexpr (MODE == MR_NOISE_TABLE) cur = st.cursor[i];
    }
    int32_t ep = 0;                                  // episodes this env has finished (recording key)
    if (live && io.episode_counter) ep = io.episode_counter[i];
    // t_k is carried from step to step: the table holds t_{k+1} = fl(t_k + dt) (mr_fill_time_table_host), which is this
    // step's tb, so the per-step table read (a dependent L1 load at the head of every step) is only needed after a reset
    double t_cur = time_at(tv, e.counter, p.dt);
    Observation o = observe(e, p);
    // an env that is stepped past its terminal step (run_sim does, utils.py:46-54; the reference never resets by
    // itself) stays `done` on every later step: only the transition counts as an episode end in the statistics.
    // A freshly reset env (counter 0) has not ended anything yet, wherever it starts.
    bool was_done = o.done && e.counter > 0;
    o.rew = 0.0; o.done = false; o.why = 0;
    bool overflow = false;

    // tensor-core actor + generated noise, matched model: the step's eight normals are drawn while the actor's MMAs are in
    // flight (same values as PhiloxNoise::draw8 inside sim_step, so the results are bit-identical to the composed path)
    constexpr bool kFillDraws = SRC == kSrcActorTc16 && MODE == MR_NOISE_PHILOX && !MISM && !PERENV;
#ifndef MR_ACTOR_FILL
#define MR_ACTOR_FILL 3          // bit 0: block 0 behind the layer-1 MMAs, bit 1: block 1 behind the layer-2 MMAs
#endif
    for (int k = 0; k < io.k_steps; ++k) {
        double f_t = 0.0, al = 0.0;
        float z8[8];
        if constexpr (SRC == MR_ACTIONS_TENSOR) {
            if (live) {
                const T* a = io.actions + ((int64_t)k * n + i) * 2;
                if constexpr (sizeof(T) == 8) { const double2 v = *reinterpret_cast<const double2*>(a); f_t = v.x; al = v.y; }
                else { const float2 v = *reinterpret_cast<const float2*>(a); f_t = v.x; al = v.y; }
            }
        } else if constexpr (SRC == MR_ACTIONS_BROADCAST) {
            f_t = (double)io.actions[2 * k]; al = (double)io.actions[2 * k + 1];
        } else if constexpr (SRC == MR_ACTIONS_PHILOX) {
            double u[4];
            philox_uniform4(p, nv.env_base + (uint64_t)i, off + (uint64_t)k, kPurposeAction, u);
            f_t = p.act_hi[0] * u[0]; al = p.act_hi[1] * u[1];   // U[0,20) x U[0,2pi)
        } else {
            const float obs5[5] = {(float)e.x, (float)e.y, 0.f, 0.f, (float)o.d};
            float a2[2];
            if constexpr (SRC == kSrcActorTc)
                actor_tc_forward(*reinterpret_cast<ActorTcSmem*>(s_dyn), obs5, (float)p.act_hi[0], (float)p.act_hi[1], actor_calls++, a2);
            else if constexpr (SRC == kSrcActorTc16) {
                if constexpr (kFillDraws) {
                    PhiloxNoise nzd;
                    nzd.seek(nv.env_base + (uint64_t)(live ? i : 0), off + (uint64_t)k);
                    auto d0 = [&]() { nzd.draw4_at(p, 0u, z8); };
                    auto d1 = [&]() { nzd.draw4_at(p, 1u, z8 + 4); };
                    if constexpr ((MR_ACTOR_FILL & 1) == 0) d0();
                    if constexpr ((MR_ACTOR_FILL & 2) == 0) d1();
                    auto f1 = [&]() { if constexpr ((MR_ACTOR_FILL & 1) != 0) d0(); };
                    auto f2 = [&]() { if constexpr ((MR_ACTOR_FILL & 2) != 0) d1(); };
                    actor_tc16_forward(*reinterpret_cast<ActorTc16Smem*>(s_dyn), obs5, (float)p.act_hi[0], (float)p.act_hi[1],
                                       actor_calls++, a2, f1, f2);
                } else {
                    actor_tc16_forward(*reinterpret_cast<ActorTc16Smem*>(s_dyn), obs5, (float)p.act_hi[0], (float)p.act_hi[1], actor_calls++, a2);
                }
            }
            else
                actor_forward_smem(s_actor, obs5, (float)p.act_hi[0], (float)p.act_hi[1], a2);
            f_t = (double)a2[0]; al = (double)a2[1];
        }
        auto nz = make_noise<MODE>(nv, n, live ? i : 0, cur, off + (uint64_t)k);
        const double t = t_cur;
        const double tb = t + p.dt, tb2 = tb + p.dt;
        t_cur = tb;
        e.counter += 1;
        if constexpr (PERENV) {
            if (mism_i) sim_step<true>(e, t, tb, tb2, f_t, al, p, nz);
            else sim_step<false>(e, t, tb, tb2, f_t, al, p, nz);
        } else if constexpr (kFillDraws) {
            nz.blk += 2;                                     // where draw8() would have left the stream
            sim_step_drawn(e, t, tb, tb2, f_t, al, p, nz, z8);
        } else {
            sim_step<MISM>(e, t, tb, tb2, f_t, al, p, nz);
        }
        o = observe(e, p);
        if constexpr (MODE == MR_NOISE_TABLE) { cur = nz.cursor; overflow |= nz.overflow != 0; }
        if (live && recording) {
            if (io.traj_xy) {
                io.traj_xy[((int64_t)k * 2) * n + i] = (T)e.x;
                io.traj_xy[((int64_t)k * 2 + 1) * n + i] = (T)e.y;
            }
            if (io.traj_sp) {
                io.traj_sp[((int64_t)k * 2) * n + i] = (T)e.spx;
                io.traj_sp[((int64_t)k * 2 + 1) * n + i] = (T)e.spy;
            }
            if (io.traj_done) io.traj_done[(int64_t)k * n + i] = o.done ? 1 : 0;
            if (io.traj_actions) { io.traj_actions[((int64_t)k * n + i) * 2] = (T)f_t; io.traj_actions[((int64_t)k * n + i) * 2 + 1] = (T)al; }
            if (io.traj_rew) io.traj_rew[(int64_t)k * n + i] = (T)o.rew;
            if (io.traj_episode) io.traj_episode[(int64_t)k * n + i] = ep;
            if (io.traj_step) io.traj_step[(int64_t)k * n + i] = e.counter;
        }
        if (live) sum_rew += o.rew;
        if (o.done) {
            if (live && !was_done) {
                sum_len += e.counter;
                if (o.why == 1) ++n_goal;
                else if (o.why == 2) ++n_oob;
                else ++n_timeout;
            }
            if (!was_done) ++ep;
            if (p.auto_reset) {
                int ov = 0;
                double x0, y0;
                if (io.reset_init) {                        // the caller's start positions, one per finished episode
                    const T* r = io.reset_init + ((int64_t)((ep - 1) % io.reset_init_len) * n + (live ? i : 0)) * 2;
                    x0 = (double)r[0]; y0 = (double)r[1];
                } else {
                    sample_init(nv, live ? i : 0, off + (uint64_t)k, p, x0, y0);
                }
                if constexpr (PERENV) {
                    if (mism_i) auto_reset_at<MODE, true>(e, x0, y0, nv, n, live ? i : 0, cur, off + (uint64_t)k, p, ov);
                    else auto_reset_at<MODE, false>(e, x0, y0, nv, n, live ? i : 0, cur, off + (uint64_t)k, p, ov);
                } else {
                    auto_reset_at<MODE, MISM>(e, x0, y0, nv, n, live ? i : 0, cur, off + (uint64_t)k, p, ov);
                }
                overflow |= ov != 0;
                t_cur = time_at(tv, e.counter, p.dt);   // counter 0 again
                o.d = sqrt(e.x * e.x + e.y * e.y);   // the policy's next input is the new episode's first observation (env.reset())
                if (live && recording && io.traj_reset_xy) {
                    io.traj_reset_xy[((int64_t)k * 2) * n + i] = (T)e.x;
                    io.traj_reset_xy[((int64_t)k * 2 + 1) * n + i] = (T)e.y;
                }
            }
        }
        was_done = o.done && !p.auto_reset;
    }

    if (live) {
        n_live_steps += io.k_steps;
        if (overflow) e.status |= kNoiseOverflow;
        if (e.status) ++n_failed;
        const double t_out = time_at(tv, e.counter, p.dt);
        st.x[i] = (T)e.x; st.y[i] = (T)e.y; st.fx[i] = (T)e.fx; st.fy[i] = (T)e.fy;
        st.h[i] = encode_h<T>(e.h, (t_out + p.dt) - t_out);
        st.counter[i] = e.counter;
        if constexpr (MODE == MR_NOISE_TABLE) st.cursor[i] = cur;
        if (e.status) st.status[i] |= (uint8_t)e.status;
        if (io.episode_counter) io.episode_counter[i] = ep;
        // outputs of the LAST step (after an auto reset: the first observation of the new episode)
        if (out.obs) {
            out.obs[i] = (T)e.x; out.obs[out.stride + i] = (T)e.y;
            if (out.goal) { out.obs[2 * out.stride + i] = (T)0; out.obs[3 * out.stride + i] = (T)0; }
            out.obs[4 * out.stride + i] = (T)sqrt(e.x * e.x + e.y * e.y);
        }
        if (out.rew) out.rew[i] = (T)o.rew;
        if (out.done) out.done[i] = o.done ? 1 : 0;
        if (out.sp) { out.sp[i] = (T)e.spx; out.sp[out.stride + i] = (T)e.spy; }
    }
    }   // tiles
    if constexpr (SRC == kSrcActorTc) actor_tc_teardown(*reinterpret_cast<ActorTcSmem*>(s_dyn));
    if constexpr (SRC == kSrcActorTc16) actor_tc16_teardown(*reinterpret_cast<ActorTc16Smem*>(s_dyn));

    if (io.stats) {   // warp shuffle -> shared -> one atomic per block per statistic
        double acc[MR_STATS_LEN];
#pragma unroll
        for (int k = 0; k < MR_STATS_LEN; ++k) acc[k] = 0.0;
        acc[MR_STAT_ENV_STEPS] = (double)n_live_steps;
        acc[MR_STAT_SUM_REWARD] = sum_rew;
        acc[MR_STAT_EPISODES] = (double)(n_goal + n_oob + n_timeout);
        acc[MR_STAT_SUM_LENGTH] = (double)sum_len;
        acc[MR_STAT_GOAL] = (double)n_goal;
        acc[MR_STAT_OUT_OF_BOUNDS] = (double)n_oob;
        acc[MR_STAT_TIMEOUT] = (double)n_timeout;
        acc[MR_STAT_FAILED] = (double)n_failed;
#pragma unroll
        for (int k = 0; k < MR_STATS_LEN; ++k) {
            double v = acc[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&s_stats[k], v);
        }
        __syncthreads();
        if (threadIdx.x < MR_STATS_LEN && s_stats[threadIdx.x] != 0.0) atomicAdd(io.stats + threadIdx.x, s_stats[threadIdx.x]);
    }
}


template <class T, int MODE, bool MISM>
static int rollout_src(const StateView<T>& sv, const RolloutView<T>& rv, const OutView<T>& ov, const NoiseView& nv,
                       const TimeView& tv, const Params& p, int64_t n, cudaStream_t s) {
    const int threads = 128;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    if (sv.a0) {                                            // per-env parameter rows (instantiated once, under MISM = false)
        if constexpr (!MISM) {
            switch (rv.action_source) {
                case MR_ACTIONS_TENSOR:
                    env_rollout_kernel<T, MODE, false, MR_ACTIONS_TENSOR, true><<<blocks, threads, 0, s>>>(sv, rv, ov, nv, tv, p, n); break;
                case MR_ACTIONS_BROADCAST:
                    env_rollout_kernel<T, MODE, false, MR_ACTIONS_BROADCAST, true><<<blocks, threads, 0, s>>>(sv, rv, ov, nv, tv, p, n); break;
                case MR_ACTIONS_PHILOX:
                    env_rollout_kernel<T, MODE, false, MR_ACTIONS_PHILOX, true><<<blocks, threads, 0, s>>>(sv, rv, ov, nv, tv, p, n); break;
                default:
                    return fail(MR_ERR_UNSUPPORTED, "mr_env_rollout: the in-kernel actor policy does not take per-env parameter rows");
            }
            return check_launch("mr_env_rollout");
        }
    }
    // one wave of the uncapped kernel (2 CTAs per SM at ~145-250 registers) covers the batch: latency-bound, no register
    // cap.  Measured (tensor actions, sigma = 1, K = 64; capped / uncapped): 4096 envs 76.6 / 68.3 us, 32768 envs 86.8 / 78.5 us,
    // 49152 envs (more than one wave of the uncapped kernel) 95 / 150 us
    int sms_small = 0;
    cudaDeviceGetAttribute(&sms_small, cudaDevAttrMultiProcessorCount, current_device());
    static const int small_env = [] { const char* e = getenv("MR_ROLLOUT_SMALL"); return !e ? -1 : (e[0] == '1' ? 1 : 0); }();   // A/B runs
    const bool small = small_env >= 0 ? small_env == 1 : blocks <= (unsigned)(2 * (sms_small > 0 ? sms_small : 148));
    switch (rv.action_source) {
        case MR_ACTIONS_TENSOR:
            if (small) env_rollout_kernel<T, MODE, MISM, MR_ACTIONS_TENSOR, false, true><<<blocks, threads, 0, s>>>(sv, rv, ov, nv, tv, p, n);
            else env_rollout_kernel<T, MODE, MISM, MR_ACTIONS_TENSOR><<<blocks, threads, 0, s>>>(sv, rv, ov, nv, tv, p, n);
            break;
        case MR_ACTIONS_BROADCAST:
            if (small) env_rollout_kernel<T, MODE, MISM, MR_ACTIONS_BROADCAST, false, true><<<blocks, threads, 0, s>>>(sv, rv, ov, nv, tv, p, n);
            else env_rollout_kernel<T, MODE, MISM, MR_ACTIONS_BROADCAST><<<blocks, threads, 0, s>>>(sv, rv, ov, nv, tv, p, n);
            break;
        case MR_ACTIONS_PHILOX:
            if (small) env_rollout_kernel<T, MODE, MISM, MR_ACTIONS_PHILOX, false, true><<<blocks, threads, 0, s>>>(sv, rv, ov, nv, tv, p, n);
            else env_rollout_kernel<T, MODE, MISM, MR_ACTIONS_PHILOX><<<blocks, threads, 0, s>>>(sv, rv, ov, nv, tv, p, n);
            break;
        case MR_ACTIONS_ACTOR: {
            // default: both dense layers on the tensor cores (tcgen05, 3xFP16, 4 CTAs/SM); MR_ACTOR_PATH=tf32 selects the
            // 3xTF32 hidden-layer kernel (fp32 operand range), MR_ACTOR_PATH=simt the CUDA-core MLP
            const int path = g_actor_path;              // mr_set_actor_path(), initialised from MR_ACTOR_PATH
            static bool attr[kMaxDevices] = {};
            const int dev = current_device();
            if (!attr[dev]) {
                cudaFuncSetAttribute(env_rollout_kernel<T, MODE, MISM, kSrcActorTc>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(ActorTcSmem));
                cudaFuncSetAttribute(env_rollout_kernel<T, MODE, MISM, kSrcActorTc16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(ActorTc16Smem));
                attr[dev] = true;
            }
            if (path == 1)
                env_rollout_kernel<T, MODE, MISM, MR_ACTIONS_ACTOR><<<blocks, threads, kActorParams * sizeof(float), s>>>(sv, rv, ov, nv, tv, p, n);
            else if (path == 2)
                env_rollout_kernel<T, MODE, MISM, kSrcActorTc><<<blocks, kTcRows, sizeof(ActorTcSmem), s>>>(sv, rv, ov, nv, tv, p, n);
            else {
                int sms = 0;                                   // persistent: the CTAs that are co-resident (4 per SM)
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                const unsigned resident = (unsigned)(4 * (sms > 0 ? sms : 148));
                env_rollout_kernel<T, MODE, MISM, kSrcActorTc16><<<blocks < resident ? blocks : resident, kT16Rows, sizeof(ActorTc16Smem), s>>>(sv, rv, ov, nv, tv, p, n);
            }
            break;
        }
        default: return fail(MR_ERR_ARG, "mr_env_rollout: unknown action source %d", rv.action_source);
    }
    return check_launch("mr_env_rollout");
}

template <>
int launch_rollout<MR_T, MR_MODE>(const StateView<MR_T>& sv, const RolloutView<MR_T>& rv, const OutView<MR_T>& ov,
                                  const NoiseView& nv, const TimeView& tv, const Params& p, int64_t n, cudaStream_t s) {
    return (p.mism && !sv.a0) ? rollout_src<MR_T, MR_MODE, true>(sv, rv, ov, nv, tv, p, n, s)
                              : rollout_src<MR_T, MR_MODE, false>(sv, rv, ov, nv, tv, p, n, s);
}

}  // namespace mr
