// TEST ARTEFACT: compiles mr_rl_b200/csrc/mr_core.cuh (the __host__ __device__ per-env
// algorithm the CUDA kernels run) with g++ so the control flow can be checked against the
// golden vectors in the GPU-less build container.  Not part of the product library.
#include <cstdint>
#include <cstring>

#include "../mr_rl_b200/csrc/mr_core.cuh"

extern "C" int host_core_rollout(const double* actions /*[T][2]*/, int T, double x0, double y0, double sigma, double a0,
                                 int mism, int mism_at_reset, const double* z, int zlen,
                                 double* pos /*[T][2]*/, double* obs_d, uint8_t* done, int32_t* counter,
                                 int64_t* cursor, int32_t* attempts, double* carry /*[T][3] fx fy h*/,
                                 double* sp /*[T][2]*/, double* reset_out /*[4] fx fy h cursor*/) {
    using namespace mr;
    Params p;
    p.a0 = a0; p.sigma = sigma; p.dt = 0.030; p.rtol = 0.030 / 100; p.atol = 1e-4;
    p.min_dist = 30; p.bound_xy = 5000; p.bound_d = 80000;
    p.init_lo[0] = p.init_lo[1] = 100; p.init_hi[0] = p.init_hi[1] = 120; p.act_hi[0] = 20; p.act_hi[1] = 6.283185307179586;
    p.mism = mism; p.mism_reset = mism_at_reset; p.max_steps = 50; p.reward_mode = 0; p.auto_reset = 0;
    finalize_params(p);
    philox_make_keys(0, p.keys);
    TableNoise nz;
    nz.col = z; nz.stride = 1; nz.cursor = 0; nz.len = zlen; nz.overflow = 0;
    Env e;
    e.spx = e.spy = 0;
    if (mism_at_reset) env_reset<true>(e, x0, y0, p.dt, p, nz); else env_reset<false>(e, x0, y0, p.dt, p, nz);
    reset_out[0] = e.fx; reset_out[1] = e.fy; reset_out[2] = e.h; reset_out[3] = (double)nz.cursor;
    double t = 0.0;
    for (int k = 0; k < T; ++k) {
        const double tb = t + p.dt, tb2 = tb + p.dt;
        e.counter += 1;
        int att;
        if (mism) att = sim_step<true>(e, t, tb, tb2, actions[2 * k], actions[2 * k + 1], p, nz);
        else att = sim_step<false>(e, t, tb, tb2, actions[2 * k], actions[2 * k + 1], p, nz);
        const Observation o = observe(e, p);
        pos[2 * k] = e.x; pos[2 * k + 1] = e.y; obs_d[k] = o.d; done[k] = o.done; counter[k] = e.counter;
        cursor[k] = nz.cursor; attempts[k] = att; carry[3 * k] = e.fx; carry[3 * k + 1] = e.fy; carry[3 * k + 2] = e.h;
        sp[2 * k] = e.spx; sp[2 * k + 1] = e.spy;
        t = tb;
    }
    return e.status | (nz.overflow ? kNoiseOverflow : 0);
}
