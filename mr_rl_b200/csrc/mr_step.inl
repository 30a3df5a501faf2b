// Single-step and reset kernels.  Included by the per-(dtype, noise mode) instantiation units
// with MR_T and MR_MODE defined.
#include <cstdlib>

#include "mr_common.cuh"
#include "mr_step_tma.cuh"
#include "mr_step_ws.cuh"

namespace mr {

// =============================================================================================
// Single step: one launch = MR_Env.step for n envs.  HBM-bound: 60 B read + 93 B written per
// env-step in fp64 storage; VEC consecutive envs per thread so every row access is 16 bytes.
// =============================================================================================
#ifndef MR_STEP_MINB
#define MR_STEP_MINB 1
#endif

// MR_Env.step for one env held in registers (shared by the scalar / vector / per-env-parameter kernels).
template <int MODE, bool MISM, class T>
__device__ __forceinline__ Observation step_one(Env& e, T h_raw, double f_t, double al, const NoiseView& nv, uint64_t off,
                                                const TimeView& tv, const Params& p, int64_t n, int64_t i, int32_t& cur,
                                                double& d_out, double& il_next) {
    const double t = time_at(tv, e.counter, p.dt);
    const double tb = t + p.dt, tb2 = tb + p.dt;
    e.h = decode_h<T>(h_raw, tb - t);
    auto nz = make_noise<MODE>(nv, n, i, cur, off);
    e.counter += 1;                                               // MR_env.py:80
    sim_step<MISM>(e, t, tb, tb2, f_t, al, p, nz);
    const Observation o = observe(e, p);
    if constexpr (MODE == MR_NOISE_TABLE) { cur = nz.cursor; if (nz.overflow) e.status |= kNoiseOverflow; }
    d_out = o.d; il_next = tb2 - tb;
    if (p.auto_reset && o.done) {   // reported obs = first obs of the new episode; done/rew = terminal step
        int ov = 0;
        auto_reset_env<MODE, MISM>(e, nv, n, i, cur, off, p, ov);
        if (ov) e.status |= kNoiseOverflow;
        d_out = sqrt(e.x * e.x + e.y * e.y);
        il_next = p.dt;
    }
    return o;
}

template <class T, int VEC, int MODE, bool MISM>
__global__ void __launch_bounds__(256, MR_STEP_MINB)
env_step_kernel(StateView<T> st, const T* __restrict__ actions, OutView<T> out, NoiseView nv, TimeView tv,
                Params p, int64_t n) {
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i0 >= n) return;
    if (VEC > 1 && i0 + VEC > n) return;   // host launches a VEC=1 tail kernel for the remainder

    // all loads first: independent 16-byte requests in flight per thread
    const Pack<T, VEC> px = load_pack<T, VEC>(st.x, i0), py = load_pack<T, VEC>(st.y, i0);
    const Pack<T, VEC> pfx = load_pack<T, VEC>(st.fx, i0), pfy = load_pack<T, VEC>(st.fy, i0);
    const Pack<T, VEC> ph = load_pack<T, VEC>(st.h, i0);
    const Pack<int32_t, VEC> pc = load_pack<int32_t, VEC>(st.counter, i0);
    Pack<T, VEC> pa0, pa1;
    if constexpr (VEC == 1) {
        if (sizeof(T) == 8 && p.act_f32) {                                 // float32 actions, fp64 storage (scalar kernel only)
            const float2 a = reinterpret_cast<const float2*>(actions)[i0];
            pa0.v[0] = (T)a.x; pa1.v[0] = (T)a.y;
        } else { pa0.v[0] = actions[2 * i0]; pa1.v[0] = actions[2 * i0 + 1]; }
    } else {
        pa0 = load_pack<T, VEC>(actions, 2 * i0);                           // [n][2] -> 2*VEC values
        pa1 = load_pack<T, VEC>(actions, 2 * i0 + VEC);
    }
    Pack<int32_t, VEC> pcur;
    if constexpr (MODE == MR_NOISE_TABLE) pcur = load_pack<int32_t, VEC>(st.cursor, i0);
    const uint64_t off = step_offset(nv);

    Pack<T, VEC> ox, oy, ofx, ofy, oh, od, orew, ospx, ospy;
    Pack<int32_t, VEC> oc, ocur;
    Pack<uint8_t, VEC> odone;
    int any_status = 0;
    int status_v[VEC];

#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        Env e;
        e.x = (double)px.v[j]; e.y = (double)py.v[j]; e.fx = (double)pfx.v[j]; e.fy = (double)pfy.v[j];
        e.counter = pc.v[j]; e.status = 0; e.spx = e.spy = 0.0;
        double f_t, al;
        if constexpr (VEC == 1) { f_t = (double)pa0.v[0]; al = (double)pa1.v[0]; }
        else {
            const int q = 2 * j;   // position inside the 2*VEC action values of this thread
            f_t = (double)(q < VEC ? pa0.v[q] : pa1.v[q - VEC]);
            al = (double)(q + 1 < VEC ? pa0.v[q + 1] : pa1.v[q + 1 - VEC]);
        }
        int32_t cur = 0;
        if constexpr (MODE == MR_NOISE_TABLE) cur = pcur.v[j];
        double d_out, il_next;
        const Observation o = step_one<MODE, MISM, T>(e, ph.v[j], f_t, al, nv, off, tv, p, n, i0 + j, cur, d_out, il_next);
        if constexpr (MODE == MR_NOISE_TABLE) ocur.v[j] = cur;
        ox.v[j] = (T)e.x; oy.v[j] = (T)e.y; ofx.v[j] = (T)e.fx; ofy.v[j] = (T)e.fy; oh.v[j] = encode_h<T>(e.h, il_next);
        oc.v[j] = e.counter; od.v[j] = (T)d_out; orew.v[j] = (T)o.rew; odone.v[j] = o.done ? 1 : 0;
        ospx.v[j] = (T)e.spx; ospy.v[j] = (T)e.spy;
        status_v[j] = e.status; any_status |= e.status;
    }

    store_pack<T, VEC>(st.x, i0, ox); store_pack<T, VEC>(st.y, i0, oy);
    store_pack<T, VEC>(st.fx, i0, ofx); store_pack<T, VEC>(st.fy, i0, ofy);
    store_pack<T, VEC>(st.h, i0, oh);
    store_pack<int32_t, VEC>(st.counter, i0, oc);
    if constexpr (MODE == MR_NOISE_TABLE) store_pack<int32_t, VEC>(st.cursor, i0, ocur);
    if (sizeof(T) == 8 && out.f32) {                                      // float32 output rows (element-wise: rare path)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (out.obs) {
                put_out<T>(out.obs, true, i0 + j, (double)ox.v[j]);
                put_out<T>(out.obs, true, out.stride + i0 + j, (double)oy.v[j]);
                if (out.goal) { put_out<T>(out.obs, true, 2 * out.stride + i0 + j, 0.0); put_out<T>(out.obs, true, 3 * out.stride + i0 + j, 0.0); }
                put_out<T>(out.obs, true, 4 * out.stride + i0 + j, (double)od.v[j]);
            }
            if (out.rew) put_out<T>(out.rew, true, i0 + j, (double)orew.v[j]);
            if (out.sp) { put_out<T>(out.sp, true, i0 + j, (double)ospx.v[j]); put_out<T>(out.sp, true, out.stride + i0 + j, (double)ospy.v[j]); }
        }
    } else {
        if (out.obs) {
            Pack<T, VEC> zero;
#pragma unroll
            for (int j = 0; j < VEC; ++j) zero.v[j] = (T)0;
            store_pack<T, VEC>(out.obs, i0, ox);
            store_pack<T, VEC>(out.obs + out.stride, i0, oy);
            if (out.goal) {
                store_pack<T, VEC>(out.obs + 2 * out.stride, i0, zero);   // goal is always (0,0), MR_env.py:57
                store_pack<T, VEC>(out.obs + 3 * out.stride, i0, zero);
            }
            store_pack<T, VEC>(out.obs + 4 * out.stride, i0, od);
        }
        if (out.rew) store_pack<T, VEC>(out.rew, i0, orew);
        if (out.sp) { store_pack<T, VEC>(out.sp, i0, ospx); store_pack<T, VEC>(out.sp + out.stride, i0, ospy); }
    }
    if (out.done) store_pack<uint8_t, VEC>(out.done, i0, odone);
    if (any_status) {                                                 // rare: sticky flags
#pragma unroll
        for (int j = 0; j < VEC; ++j) if (status_v[j]) st.status[i0 + j] |= (uint8_t)status_v[j];
    }
}

// The same step with PER-ENV a0 / noise_var / is_mismatched rows (MR_simulator.py:16-19 are instance attributes): one env
// per thread, the model flag is a run-time branch.  A parameter sweep over a million envs is one launch.
template <class T, int MODE>
__global__ void __launch_bounds__(128)
env_step_perenv_kernel(StateView<T> st, const T* __restrict__ actions, OutView<T> out, NoiseView nv, TimeView tv,
                       Params p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Env e;
    e.x = (double)st.x[i]; e.y = (double)st.y[i]; e.fx = (double)st.fx[i]; e.fy = (double)st.fy[i];
    e.counter = st.counter[i]; e.status = 0; e.spx = e.spy = 0.0;
    const T h_raw = st.h[i];
    double f_t, al;
    if (sizeof(T) == 8 && p.act_f32) { const float2 a = reinterpret_cast<const float2*>(actions)[i]; f_t = a.x; al = a.y; }
    else { f_t = (double)actions[2 * i]; al = (double)actions[2 * i + 1]; }
    int32_t cur = 0;
    if constexpr (MODE == MR_NOISE_TABLE) cur = st.cursor[i];
    p.a0 = st.a0[i]; p.sigma = st.sigma[i];                           // this thread's copy of the launch parameters
    const uint64_t off = step_offset(nv);
    double d_out, il_next;
    Observation o;
    if (st.mism[i]) o = step_one<MODE, true, T>(e, h_raw, f_t, al, nv, off, tv, p, n, i, cur, d_out, il_next);
    else o = step_one<MODE, false, T>(e, h_raw, f_t, al, nv, off, tv, p, n, i, cur, d_out, il_next);
    st.x[i] = (T)e.x; st.y[i] = (T)e.y; st.fx[i] = (T)e.fx; st.fy[i] = (T)e.fy; st.h[i] = encode_h<T>(e.h, il_next);
    st.counter[i] = e.counter;
    if constexpr (MODE == MR_NOISE_TABLE) st.cursor[i] = cur;
    if (out.obs) {
        put_out<T>(out.obs, out.f32, i, e.x); put_out<T>(out.obs, out.f32, out.stride + i, e.y);
        if (out.goal) { put_out<T>(out.obs, out.f32, 2 * out.stride + i, 0.0); put_out<T>(out.obs, out.f32, 3 * out.stride + i, 0.0); }
        put_out<T>(out.obs, out.f32, 4 * out.stride + i, d_out);
    }
    if (out.rew) put_out<T>(out.rew, out.f32, i, o.rew);
    if (out.done) out.done[i] = o.done ? 1 : 0;
    if (out.sp) { put_out<T>(out.sp, out.f32, i, e.spx); put_out<T>(out.sp, out.f32, out.stride + i, e.spy); }
    if (e.status) st.status[i] |= (uint8_t)e.status;
}

// =============================================================================================
// Reset: MR_Env.reset for the masked envs.
// =============================================================================================
template <class T, int MODE>
__global__ void __launch_bounds__(256)
env_reset_kernel(StateView<T> st, const T* __restrict__ init_xy, const uint8_t* __restrict__ mask, int reset_cursor,
                 ResetRows rr, OutView<T> out, NoiseView nv, Params p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) return;
    const uint64_t off = step_offset(nv);
    double x0, y0;
    if (init_xy) { x0 = (double)init_xy[2 * i]; y0 = (double)init_xy[2 * i + 1]; }
    else {
        double u[4];
        philox_uniform4(p, nv.env_base + (uint64_t)i, off, kPurposeInit, u);
        // gym Box.sample: uniform(low, high).astype(float32)
        x0 = (double)(float)(p.init_lo[0] + (p.init_hi[0] - p.init_lo[0]) * u[0]);
        y0 = (double)(float)(p.init_lo[1] + (p.init_hi[1] - p.init_lo[1]) * u[1]);
    }
    // MR_env.py:179-183: noise_var and a0 are assigned BEFORE the integrator is built, is_mismatched AFTER it
    bool mism_old = p.mism_reset != 0;
    if (st.a0) {
        mism_old = st.mism[i] != 0;
        p.a0 = rr.a0 ? rr.a0[i] : p.a0;
        p.sigma = rr.sigma ? rr.sigma[i] : p.sigma;
        st.a0[i] = p.a0; st.sigma[i] = p.sigma;
        st.mism[i] = rr.mism ? (rr.mism[i] ? 1 : 0) : (p.mism ? 1 : 0);
    }
    int32_t cur = 0;
    if constexpr (MODE == MR_NOISE_TABLE) cur = reset_cursor ? 0 : st.cursor[i];
    auto nz = make_noise<MODE>(nv, n, i, cur, off, kPurposeResetNoise);
    Env e;
    e.spx = e.spy = 0.0;
    if (mism_old) env_reset<true>(e, x0, y0, p.dt, p, nz);
    else env_reset<false>(e, x0, y0, p.dt, p, nz);
    if constexpr (MODE == MR_NOISE_TABLE) { st.cursor[i] = nz.cursor; if (nz.overflow) e.status |= kNoiseOverflow; }
    st.x[i] = (T)e.x; st.y[i] = (T)e.y; st.fx[i] = (T)e.fx; st.fy[i] = (T)e.fy; st.h[i] = encode_h<T>(e.h, p.dt);
    st.counter[i] = 0;
    st.status[i] = (uint8_t)e.status;
    if (out.obs) {
        out.obs[i] = (T)e.x; out.obs[out.stride + i] = (T)e.y;
        if (out.goal) { out.obs[2 * out.stride + i] = (T)0; out.obs[3 * out.stride + i] = (T)0; }
        out.obs[4 * out.stride + i] = (T)sqrt(e.x * e.x + e.y * e.y);
    }
    if (out.rew) out.rew[i] = (T)0;
    if (out.done) out.done[i] = 0;
    if (out.sp) { out.sp[i] = (T)e.spx; out.sp[out.stride + i] = (T)e.spy; }
}


template <class T, int VEC, int MODE, bool MISM>
static void launch_step_range(const StateView<T>& sv, const T* actions, const OutView<T>& ov, const NoiseView& nv,
                              const TimeView& tv, const Params& p, int64_t n, cudaStream_t s) {
    const int threads = 256;
    const int64_t items = (n + VEC - 1) / VEC;
    const int64_t blocks = (items + threads - 1) / threads;
    if (blocks > 0)
        env_step_kernel<T, VEC, MODE, MISM><<<(unsigned)blocks, threads, 0, s>>>(sv, actions, ov, nv, tv, p, n);
}

template <class T>
static StateView<T> offset_state(StateView<T> v, int64_t o) {
    v.x += o; v.y += o; v.fx += o; v.fy += o; v.h += o; v.counter += o;
    if (v.cursor) v.cursor += o;
    v.status += o;
    if (v.a0) { v.a0 += o; v.sigma += o; v.mism += o; }
    return v;
}

template <class T>
static OutView<T> offset_out(OutView<T> v, int64_t o) {
    if (v.obs) v.obs = out_row(v.obs, v.f32, o);
    if (v.rew) v.rew = out_row(v.rew, v.f32, o);
    if (v.done) v.done += o;
    if (v.sp) v.sp = out_row(v.sp, v.f32, o);
    return v;
}

static int sm_count(int dev) {
    static int cached[kMaxDevices] = {};
    if (!cached[dev]) {
        cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
        if (cached[dev] <= 0) cached[dev] = 148;
    }
    return cached[dev];
}

// kernel choice override (tuning / A-B measurements / tests): mr_set_step_path(), initialised from the environment
// variable MR_STEP_PATH=tma|ws|vec|scalar ("tma" = the plain TMA kernel also where the warp-specialised one would be
// picked, "ws" = the default choice)
// (5 = "tmap": same as the default choice; 1 = "tma": 1-D bulk copies only)
static int step_path_override() { return g_step_path; }

template <class T, class K, class... Extra>
static void launch_persistent(K kernel, int threads, size_t smem, int* ctas_per_sm, int64_t n_tiles, cudaStream_t s,
                              const StateView<T>& sv, const T* act, const OutView<T>& ov, const NoiseView& nv,
                              const TimeView& tv, const Params& p, int64_t n, const Extra&... extra) {
    const int dev = current_device();
    if (!ctas_per_sm[dev]) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
        ctas_per_sm[dev] = occ > 0 ? occ : 1;
    }
    // persistent grid: exactly the CTAs that are co-resident, so there is never a second wave
    const int64_t max_ctas = (int64_t)sm_count(dev) * ctas_per_sm[dev];
    unsigned grid = (unsigned)(n_tiles < max_ctas ? n_tiles : max_ctas);
    // host-buffer step (mr_env_step_host, direct mode): the kernel is bound by the PCIe link, not by occupancy, and fewer CTAs
    // keep fewer reads and writes in flight on it: 592 CTAs 0.671 ms, 148 (one per SM) 0.637 ms, 24 0.638 ms, 16 0.76 ms per
    // 2^20-env host step (profiles/r02_e2e_host_modes.txt).  MR_STEP_MAX_CTAS overrides for A/B runs.
    static const int grid_cap_env = [] { const char* e = getenv("MR_STEP_MAX_CTAS"); return e ? atoi(e) : -1; }();
    const int grid_cap = grid_cap_env >= 0 ? grid_cap_env : g_step_cta_cap * sm_count(dev);
    if (grid_cap > 0 && grid > (unsigned)grid_cap) grid = (unsigned)grid_cap;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, sv, act, ov, nv, tv, p, n_tiles, n, extra...);
}

template <class T, int MODE, bool MISM>
static void step_ranges(const StateView<T>& sv, const T* act, const OutView<T>& ov, const NoiseView& nv, const TimeView& tv,
                        const Params& p, int64_t n, bool vec_ok, cudaStream_t s) {
    constexpr int VEC = 16 / sizeof(T);
    const int force = step_path_override();
    int64_t done = 0;
    if (sv.a0) {                                            // per-env a0 / noise_var / is_mismatched rows
        const int threads = 128;
        env_step_perenv_kernel<T, MODE><<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(sv, act, ov, nv, tv, p, n);
        return;
    }
    const bool act32 = sizeof(T) == 8 && p.act_f32;         // float32 actions with fp64 storage: tiled and scalar kernels
    auto act_at = [&](int64_t env) { return act32 ? (const T*)((const float*)act + 2 * env) : act + 2 * env; };
    {
        constexpr int kTile = TileOf<T>::value;
        if (vec_ok && n >= kTile && (force == 0 || force == 1 || force == 4 || force == 5)) {
            // Blackwell path: persistent CTAs, TMA bulk copies through shared memory
            const int64_t n_tiles = n / kTile;
            static int ctas_per_sm[kMaxDevices] = {};      // per template instantiation and device
            bool ws = false;
            if constexpr (MODE == MR_NOISE_PHILOX && !MISM) ws = force == 4 && !act32 && !ov.f32;   // measured slower than the plain kernel: opt-in only
            if constexpr (MODE == MR_NOISE_PHILOX && !MISM) {
                static int ctas_ws[kMaxDevices] = {};
                if (ws) launch_persistent<T>(env_step_tma_ws_kernel<T>, WsCfg<T>::kThreads, sizeof(StepSmemWs<T>), ctas_ws,
                                             n_tiles, s, sv, act, ov, nv, tv, p, n);
            }
            if (!ws) {
                // 2-D tensor maps for the equally strided rows (state, obs pairs, state_prime, noise table) when the
                // buffers allow it, else one 1-D bulk copy per row; force == 1 ("tma") keeps the 1-D form for A/B runs
                static int ctas_tmap[kMaxDevices] = {};
                StepMaps maps;
                // measured (B200, 2^20 envs, us per launch, 1-D -> tensor maps): generated noise fp64 30.8 -> 29.8, fp32 28.2 ->
                // 26.5; table noise 61.3 -> 51.2; with the state_prime rows 32.2 -> 31.3; noise-free fp32 23.7 -> 22.9 but
                // noise-free fp64 27.7 -> 28.6 when the row stride is a large power of two (the five rows of the box then
                // hit the same DRAM channels at once; VecMREnv pads its rows: 27.9 both ways), so that case keeps 1-D copies
                const ptrdiff_t row_stride = (const char*)sv.y - (const char*)sv.x;
                const bool want = force == 5 || !(MODE == MR_NOISE_NONE && sizeof(T) == 8 && row_stride % 65536 == 0);
                const bool tmap = force != 1 && want && build_step_maps<T>(maps, sv, ov, nv, n, kTile, NoiseRows<MODE, MISM>::value);
                if (tmap)
                    launch_persistent<T>(env_step_tma_kernel<T, MODE, MISM, true>, kTile, sizeof(StepSmem<T, MODE, MISM>), ctas_tmap,
                                         n_tiles, s, sv, act, ov, nv, tv, p, n, maps);
                else
                    launch_persistent<T>(env_step_tma_kernel<T, MODE, MISM, false>, kTile, sizeof(StepSmem<T, MODE, MISM>), ctas_per_sm,
                                         n_tiles, s, sv, act, ov, nv, tv, p, n, maps);
            }
            done = n_tiles * kTile;
        } else if (vec_ok && !act32 && n >= VEC && force != 3 && force != 1 && force != 4 && force != 5) {
            if constexpr (MODE != MR_NOISE_TABLE) {        // (the table mode has the tiled kernel and the scalar one)
                const int64_t n_vec = (n / VEC) * VEC;
                launch_step_range<T, VEC, MODE, MISM>(sv, act, ov, nv, tv, p, n_vec, s);
                done = n_vec;
            }
        }
    }
    if (done < n) {
        NoiseView nv2 = nv; nv2.env_base += (uint64_t)done; nv2.table_col0 += done;
        if (done == 0) launch_step_range<T, 1, MODE, MISM>(sv, act, ov, nv, tv, p, n, s);
        else launch_step_range<T, 1, MODE, MISM>(offset_state(sv, done), act_at(done), offset_out(ov, done), nv2, tv, p,
                                                 n - done, s);
    }
}

template <>
int launch_step<MR_T, MR_MODE>(const StateView<MR_T>& sv, const MR_T* actions, const OutView<MR_T>& ov, const NoiseView& nv,
                               const TimeView& tv, const Params& p, int64_t n, bool vec_ok, cudaStream_t s) {
    if (p.mism) step_ranges<MR_T, MR_MODE, true>(sv, actions, ov, nv, tv, p, n, vec_ok, s);
    else step_ranges<MR_T, MR_MODE, false>(sv, actions, ov, nv, tv, p, n, vec_ok, s);
    return check_launch("mr_env_step");
}

template <>
int launch_reset<MR_T, MR_MODE>(const StateView<MR_T>& sv, const MR_T* init_xy, const uint8_t* mask, int reset_cursor,
                                const ResetRows& rr, const OutView<MR_T>& ov, const NoiseView& nv, const Params& p, int64_t n,
                                cudaStream_t s) {
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    env_reset_kernel<MR_T, MR_MODE><<<blocks, threads, 0, s>>>(sv, init_xy, mask, reset_cursor, rr, ov, nv, p, n);
    return check_launch("mr_env_reset");
}

}  // namespace mr
